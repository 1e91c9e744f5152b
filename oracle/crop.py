"""Oracle: box -> 256x256 normalised crop (numpy, integer-exact).  TEST INFRASTRUCTURE ONLY.

Restates
  hamer/infer.py:154-259                       hamer_inference.prepare_batch_bbox
  hamer/hamer/datasets/utils.py:15-34          expand_to_aspect_ratio
  hamer/hamer/datasets/utils.py:82-129         gen_trans_from_patch_cv (rot = 0, scale = 1)
  hamer/hamer/datasets/utils.py:318-376        generate_image_patch_cv2
  hamer/hamer/datasets/utils.py:379-392        convert_cvimg_to_tensor
and the two third-party OpenCV calls on that path (opencv-python, any 4.x; 4.13 in this image):
  cv2.getAffineTransform  (utils.py:127)  -> closed form for the axis-aligned 3-point case
  cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT, 0) on uint8 (utils.py:359-363) -> the classic
      fixed-point algorithm: inverse map in double, 10-bit coordinates (AB_BITS), 5-bit bilinear
      fractions (INTER_BITS), int16 weights scaled by 2^15, rounding (sum + 2^14) >> 15.
Pinned against cv2 itself in tests/test_oracle_crop.py.
"""
import numpy as np

IMAGE_SIZE = 256
BBOX_SHAPE = (192, 256)          # models/__init__.py:43
RESCALE = 2.5                    # infer.py:190
IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)

AB_BITS, INTER_BITS = 10, 5
AB_SCALE, INTER_TAB = 1 << AB_BITS, 1 << INTER_BITS


def expand_to_aspect_ratio(wh, target=BBOX_SHAPE):
    """utils.py:15-34."""
    w, h = float(wh[0]), float(wh[1])
    w_t, h_t = target
    if h / w < h_t / w_t:
        h_new, w_new = max(w * h_t / w_t, h), w
    else:
        h_new, w_new = h, max(h * w_t / h_t, w)
    return np.array([w_new, h_new])


def box_to_center_size(x1, y1, x2, y2):
    """infer.py:182-199 -> (center_x, center_y, final_bbox_size) as python floats (float64)."""
    cx, cy = (x1 + x2) / 2.0, (y1 + y2) / 2.0
    scale = np.array([RESCALE * (x2 - x1) / 200.0, RESCALE * (y2 - y1) / 200.0])
    size = expand_to_aspect_ratio(scale * 200, BBOX_SHAPE).max()
    return float(cx), float(cy), float(size)


def _lu_solve6(A, b):
    """cv::solve(DECOMP_LU) for a 6x6 double system (OpenCV hal LU: partial pivoting, no FMA) - the
    solver behind cv2.getAffineTransform.  Plain python floats = IEEE double, one rounding per op."""
    A = [[float(v) for v in r] for r in A]
    b = [float(v) for v in b]
    m = 6
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(A[j][i]) > abs(A[k][i]):
                k = j
        if k != i:
            for j in range(i, m):
                A[i][j], A[k][j] = A[k][j], A[i][j]
            b[i], b[k] = b[k], b[i]
        d = -1 / A[i][i]
        for j in range(i + 1, m):
            alpha = A[j][i] * d
            for c in range(i + 1, m):
                A[j][c] += alpha * A[i][c]
            b[j] += alpha * b[i]
    for i in range(m - 1, -1, -1):
        s = b[i]
        for c in range(i + 1, m):
            s -= A[i][c] * b[c]
        b[i] = s / A[i][i]
    return b


def affine_from_box(cx, cy, size, patch=IMAGE_SIZE):
    """gen_trans_from_patch_cv with rot=0, scale=1 (utils.py:82-129): the three source points are
    stored in a float32 array (centre, centre + (0, f32(size/2)), centre + (f32(size/2), 0)), the
    destination points are (p/2,p/2), (p/2,p), (p,p/2); cv2.getAffineTransform builds the 6x6
    system [x y 1 0 0 0; 0 0 0 x y 1] and solves it by LU in double."""
    half = np.float32(size * 0.5)
    s0 = (np.float32(cx), np.float32(cy))
    s1 = (np.float32(np.float64(cx) + 0.0), np.float32(np.float64(cy) + np.float64(half)))
    s2 = (np.float32(np.float64(cx) + np.float64(half)), np.float32(np.float64(cy) + 0.0))
    h = patch * 0.5
    dst = ((h, h), (h, 2 * h), (2 * h, h))
    A, b = [], []
    for (sx, sy), (dx, dy) in zip((s0, s1, s2), dst):
        A.append([sx, sy, 1, 0, 0, 0])
        A.append([0, 0, 0, sx, sy, 1])
        b += [dx, dy]
    return np.array(_lu_solve6(A, b), dtype=np.float64).reshape(2, 3)


def warp_affine_u8(img, M, out_w=IMAGE_SIZE, out_h=IMAGE_SIZE):
    """cv2.warpAffine(img, M, (out_w,out_h), INTER_LINEAR, BORDER_CONSTANT, 0) for uint8 HxWxC."""
    img = np.ascontiguousarray(img)
    H, W = img.shape[:2]
    m = np.array(M, dtype=np.float64).reshape(2, 3).copy()
    # invert (cv::warpAffine without WARP_INVERSE_MAP)
    D = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[1, 1] * D, m[0, 0] * D
    m[0, 0] = A11
    m[0, 1] *= -D
    m[1, 0] *= -D
    m[1, 1] = A22
    b1 = -m[0, 0] * m[0, 2] - m[0, 1] * m[1, 2]
    b2 = -m[1, 0] * m[0, 2] - m[1, 1] * m[1, 2]
    m[0, 2], m[1, 2] = b1, b2

    xs = np.arange(out_w, dtype=np.float64)
    ys = np.arange(out_h, dtype=np.float64)
    adelta = np.rint(m[0, 0] * xs * AB_SCALE).astype(np.int64)
    bdelta = np.rint(m[1, 0] * xs * AB_SCALE).astype(np.int64)
    rd = AB_SCALE // INTER_TAB // 2
    X0 = np.rint((m[0, 1] * ys + m[0, 2]) * AB_SCALE).astype(np.int64) + rd
    Y0 = np.rint((m[1, 1] * ys + m[1, 2]) * AB_SCALE).astype(np.int64) + rd
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    fx = (X & (INTER_TAB - 1)).astype(np.int64)
    fy = (Y & (INTER_TAB - 1)).astype(np.int64)
    # int16 bilinear weights scaled by 2^15 (products of /32 fractions are exact multiples of 32)
    w00 = (INTER_TAB - fx) * (INTER_TAB - fy) * 32
    w01 = fx * (INTER_TAB - fy) * 32
    w10 = (INTER_TAB - fx) * fy * 32
    w11 = fx * fy * 32

    def tap(yy, xx):
        inside = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64)
        return v * inside[..., None]

    acc = (tap(sy, sx) * w00[..., None] + tap(sy, sx + 1) * w01[..., None] +
           tap(sy + 1, sx) * w10[..., None] + tap(sy + 1, sx + 1) * w11[..., None])
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)


def gaussian_kernel1d(sigma, radius):
    """scipy.ndimage._filters._gaussian_kernel1d (order 0).  sigma keeps the caller's type: ViTDetDataset hands skimage a
    numpy float32 (float32 boxes, vitdet_dataset.py:37-42), so sigma * sigma and -0.5 / sigma2 round to float32 there."""
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    return phi / phi.sum()


def skimage_gaussian(img_u8, sigma, truncate=4.0):
    """skimage.filters.gaussian(img, sigma=sigma, channel_axis=2, preserve_range=True) for a uint8 HxWx3 image
    (vitdet_dataset.py:71, infer.py:316).  scikit-image is not installed here; its gaussian() is a thin wrapper:
    image.astype(float64), then scipy.ndimage.gaussian_filter(image, (sigma, sigma, 0), mode='nearest', cval=0, truncate=4.0).
    This is the numpy statement of what scipy computes (ni_filters.c NI_Correlate1D with symmetric weights: centre tap first,
    then (x[-k] + x[+k]) * w_k for k = radius .. 1; rows, then columns) - pinned bit for bit to scipy itself in
    tests/test_oracle_crop.py."""
    r = int(truncate * float(sigma) + 0.5)
    w = gaussian_kernel1d(sigma, r)
    x = img_u8.astype(np.float64)

    def pass1d(a, axis):
        n = a.shape[axis]
        idx = np.arange(n)
        out = a * w[r]
        for k in range(r, 0, -1):
            out = out + (np.take(a, np.clip(idx - k, 0, n - 1), axis=axis) + np.take(a, np.clip(idx + k, 0, n - 1), axis=axis)) * w[r - k]
        return out
    return pass1d(pass1d(x, 0), 1)


def warp_affine_f64(img, M):
    """cv2.warpAffine(img float64 HxWxC, M, (256, 256), INTER_LINEAR, BORDER_CONSTANT 0): the same fixed-point coordinates
    as the 8-bit path (1/32-pixel fractions) but float weights (the products {1 - f/32, f/32} x {1 - g/32, g/32} in float32)
    and double accumulation S00 w00 + S01 w01 + S10 w10 + S11 w11 (imgproc/imgwarp.cpp remapBilinear<Cast<double, double>, ..., float>).
    Pinned to cv2 in tests/test_oracle_crop.py."""
    H, W = img.shape[:2]
    m = np.array(M, dtype=np.float64).reshape(2, 3).copy()
    D = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]                 # cv::warpAffine's own inversion (as in warp_affine_u8)
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[1, 1] * D, m[0, 0] * D
    m[0, 0] = A11
    m[0, 1] *= -D
    m[1, 0] *= -D
    m[1, 1] = A22
    b1 = -m[0, 0] * m[0, 2] - m[0, 1] * m[1, 2]
    b2 = -m[1, 0] * m[0, 2] - m[1, 1] * m[1, 2]
    m[0, 2], m[1, 2] = b1, b2
    x = np.arange(IMAGE_SIZE)
    y = np.arange(IMAGE_SIZE)
    adelta = np.rint(m[0, 0] * x * AB_SCALE).astype(np.int64)
    bdelta = np.rint(m[1, 0] * x * AB_SCALE).astype(np.int64)
    rd = AB_SCALE // INTER_TAB // 2
    X0 = np.rint((m[0, 1] * y + m[0, 2]) * AB_SCALE).astype(np.int64) + rd
    Y0 = np.rint((m[1, 1] * y + m[1, 2]) * AB_SCALE).astype(np.int64) + rd
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    fx = X & (INTER_TAB - 1)
    fy = Y & (INTER_TAB - 1)
    f = np.arange(INTER_TAB, dtype=np.float32) * np.float32(1.0 / INTER_TAB)
    t1 = np.stack([np.float32(1.0) - f, f], 1)

    def tap(yy, xx):
        inside = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(inside[..., None], v, 0.0)

    def wgt(a, b):
        return (t1[fy, a] * t1[fx, b]).astype(np.float64)[..., None]
    return tap(sy, sx) * wgt(0, 0) + tap(sy, sx + 1) * wgt(0, 1) + tap(sy + 1, sx) * wgt(1, 0) + tap(sy + 1, sx + 1) * wgt(1, 1)


def crop_one(img_bgr, label, box):
    """One iteration of the loop in prepare_batch_bbox (infer.py:162-248).
    Returns (img (3,256,256) f32, center (2,), size, trans (2,3) f64, do_flip)."""
    x1, y1, x2, y2 = box
    do_flip = 0.0 if label == 'right' else 1.0
    cx, cy, size = box_to_center_size(x1, y1, x2, y2)
    M = affine_from_box(cx, cy, size)
    patch = warp_affine_u8(img_bgr, M)
    patch = patch[:, :, ::-1]                       # BGR -> RGB (infer.py:228)
    if label != 'right':
        patch = patch[:, ::-1, :]                   # cv2.flip(.., 1) (infer.py:229-230)
    t = np.transpose(patch, (2, 0, 1)).astype(np.float32)   # convert_cvimg_to_tensor
    mean = 255.0 * np.array(IMAGE_MEAN)
    std = 255.0 * np.array(IMAGE_STD)
    for c in range(3):                               # infer.py:235-238 (float32 array, python floats)
        t[c] = (t[c] - float(mean[c])) / float(std[c])
    return t, np.array([cx, cy], np.float32), np.float32(size), M, np.float32(do_flip)


def prepare_batch_bbox(img_bgr, boxes):
    """infer.py:154-259.  boxes = [[label, [x1,y1,x2,y2]], ...] -> dict of numpy arrays with the
    reference's keys/shapes (img, box_center, box_size, img_size, inv_trans, trans, do_flip)."""
    imgs, cen, siz, isz, tr, flips = [], [], [], [], [], []
    for label, box in boxes:
        t, c, s, M, f = crop_one(img_bgr, label, box)
        imgs.append(t); cen.append(c); siz.append(s); tr.append(M.astype(np.float32)); flips.append(f)
        isz.append(np.array([img_bgr.shape[1], img_bgr.shape[0]], np.float32))
    return {'img': np.stack(imgs), 'box_center': np.stack(cen), 'box_size': np.stack(siz),
            'img_size': np.stack(isz), 'inv_trans': np.stack(tr), 'trans': np.stack(tr),
            'do_flip': np.stack(flips)}


def vitdet_item(img_bgr, box, right, rescale_factor=2.5):
    """ViTDetDataset.__getitem__ (hamer/hamer/datasets/vitdet_dataset.py:48-95), anti-alias blur included:
    float32 box arithmetic (:37-42, 52-58), flip INSIDE the warp (generate_image_patch_cv2 do_flip,
    datasets/utils.py:347-349: the frame is mirrored and c_x = img_width - c_x - 1), BGR->RGB, CHW float32, normalise.
    Returns (img (3,256,256) f32, box_center (2,) f32, box_size f32)."""
    b = np.asarray(box, dtype=np.float32)
    center = (b[2:4] + b[0:2]) / 2.0
    scale = rescale_factor * (b[2:4] - b[0:2]) / 200.0
    wh = scale * 200
    w, h = wh
    w_t, h_t = BBOX_SHAPE
    if h / w < h_t / w_t:
        new = np.array([w, max(w * h_t / w_t, h)])
    else:
        new = np.array([max(h * w_t / h_t, w), h])
    bbox_size = new.max()
    c_x, c_y = center[0], center[1]
    img = img_bgr
    downsampling_factor = (bbox_size * 1.0) / IMAGE_SIZE / 2.0
    if downsampling_factor > 1.1:                       # vitdet_dataset.py:62-71: blur the frame, warp the float64 image
        img = skimage_gaussian(img, (downsampling_factor - 1) / 2)
    if right == 0:
        img = img[:, ::-1, :]
        c_x = img.shape[1] - c_x - 1
    M = affine_from_box(float(c_x), float(c_y), float(bbox_size))
    if img.dtype == np.uint8:
        patch = warp_affine_u8(np.ascontiguousarray(img), M)[:, :, ::-1]
    else:
        patch = warp_affine_f64(np.ascontiguousarray(img), M)[:, :, ::-1]
    t = np.transpose(patch, (2, 0, 1)).astype(np.float32)
    mean = 255.0 * np.array(IMAGE_MEAN)
    std = 255.0 * np.array(IMAGE_STD)
    for c in range(3):     # float32 arithmetic (numpy 1.x value-based casting of the float64 scalars, as in infer.py:235-238)
        t[c] = (t[c] - np.float32(mean[c])) / np.float32(std[c])
    return t, center, bbox_size
