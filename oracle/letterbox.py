"""Oracle: detector hand-off - YOLOv7 letterbox and box rescaling (numpy / torch).  TEST INFRASTRUCTURE ONLY.

Restates
  yolo/yolov7/utils/datasets.py:999-1030   letterbox(img, new_shape, auto=True, stride)
  yolo/yolov7/utils/datasets.py:137-141    LoadImage.process_img (letterbox, BGR -> RGB, HWC -> CHW)
  yolo/detector.py:119-125                 uint8 -> float32 / float16, / 255
  yolo/yolov7/utils/general.py:323-345     scale_coords / clip_coords
  yolo/detector.py:136-145                 .round(), 'right' if cls == 1 else 'left'
and the third-party call on that path, cv2.resize(INTER_LINEAR) on uint8 (opencv-python 4.x imgproc/resize.cpp,
HResizeLinear / VResizeLinear with 11-bit coefficients): ``resize_linear_u8`` is the numpy statement of that fixed-point
algorithm (what the CUDA kernel implements) and is pinned to cv2.resize itself in tests/test_oracle_letterbox.py;
``letterbox`` calls cv2 directly, exactly like the reference.
"""
import numpy as np


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 HxWxC, bit for bit."""
    sh, sw = src.shape[:2]

    def axis(ssize, dsize, clamp_weights):
        scale = ssize / dsize                                   # double
        d = np.arange(dsize)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)        # (float)((dx + 0.5) * scale_x - 0.5)
        s = np.floor(f).astype(np.int32)
        f = (f - s).astype(np.float32)
        if clamp_weights:                                       # columns: the weight moves onto the clamped tap
            lo = s < 0
            f[lo] = 0
            s[lo] = 0
            hi = s >= ssize - 1
            f[hi] = 0
            s[hi] = ssize - 1
        a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32)
        a1 = np.rint(f * np.float32(2048)).astype(np.int32)
        return np.clip(s, 0, ssize - 1), np.clip(s + 1, 0, ssize - 1), a0, a1

    sx0, sx1, ax0, ax1 = axis(sw, dw, True)
    sy0, sy1, ay0, ay1 = axis(sh, dh, False)                    # rows: taps clamped, weights kept
    S = src.astype(np.int32)
    hrow = S[:, sx0] * ax0[None, :, None] + S[:, sx1] * ax1[None, :, None]
    r0, r1 = hrow[sy0], hrow[sy1]
    out = (((ay0[:, None, None] * (r0 >> 4)) >> 16) + ((ay1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def letterbox(img, new_shape=(640, 640), color=(114, 114, 114), auto=True, scaleFill=False, scaleup=True, stride=32):
    """datasets.py:999-1030."""
    import cv2
    shape = img.shape[:2]
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    if not scaleup:
        r = min(r, 1.0)
    ratio = r, r
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    elif scaleFill:
        dw, dh = 0.0, 0.0
        new_unpad = (new_shape[1], new_shape[0])
        ratio = new_shape[1] / shape[1], new_shape[0] / shape[0]
    dw /= 2
    dh /= 2
    if shape[::-1] != new_unpad:
        img = cv2.resize(img, new_unpad, interpolation=cv2.INTER_LINEAR)
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    img = cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=color)
    return img, ratio, (dw, dh)


def process_img(img0, img_size=640, stride=32, half=False):
    """LoadImage.process_img + the tensor conversion of Detector.detect: [3, H', W'] torch tensor in 0..1 (no batch axis)."""
    import torch
    img = letterbox(img0, img_size, stride=stride)[0]
    img = np.ascontiguousarray(img[:, :, ::-1].transpose(2, 0, 1))
    img = torch.from_numpy(img)
    img = img.half() if half else img.float()
    img /= 255.0
    return img


def scale_coords(img1_shape, coords, img0_shape):
    """general.py:323-345 (ratio_pad=None), in place on a torch tensor like the reference."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    coords[:, [0, 2]] -= pad[0]
    coords[:, [1, 3]] -= pad[1]
    coords[:, :4] /= gain
    coords[:, 0].clamp_(0, img0_shape[1])
    coords[:, 1].clamp_(0, img0_shape[0])
    coords[:, 2].clamp_(0, img0_shape[1])
    coords[:, 3].clamp_(0, img0_shape[0])
    return coords


def detections_to_list(det, img1_shape, img0_shape):
    """detector.py:136-145 for one image: det [n, 6] (x1 y1 x2 y2 conf cls) -> [[label, [x1, y1, x2, y2]], ...]."""
    det = det.clone()
    dets = []
    if len(det):
        det[:, :4] = scale_coords(img1_shape, det[:, :4], img0_shape).round()
        for row in det.tolist():
            dets.append(['right' if row[-1] == 1 else 'left', row[:4]])
    return dets
