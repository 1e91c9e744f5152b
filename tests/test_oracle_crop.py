"""Crop oracle (numpy fixed-point restatement) vs OpenCV itself and vs the stored cv2 goldens."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE
from oracle import crop as oc

cv2 = pytest.importorskip('cv2')


def synthetic_frame(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (w + h))], axis=-1)
    return np.clip(base + rng.integers(-40, 40, (h, w, 3)), 0, 255).astype(np.uint8)


def reference_patch(img, box):
    """generate_image_patch_cv2 (hamer/hamer/datasets/utils.py:318-376) via cv2, as infer.py calls it."""
    cx, cy, size = oc.box_to_center_size(*box)
    src = np.zeros((3, 2), np.float32)
    c = np.zeros(2); c[0] = cx; c[1] = cy
    src[0, :] = c
    src[1, :] = c + np.array([0, size * 0.5], dtype=np.float32)
    src[2, :] = c + np.array([size * 0.5, 0], dtype=np.float32)
    dst = np.array([[128, 128], [128, 256], [256, 128]], np.float32)
    M = cv2.getAffineTransform(src, dst)
    return cv2.warpAffine(img, M, (256, 256), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT,
                          borderValue=0), M


def test_golden_crops():
    g = np.load(os.path.join(GOLDEN, 'crop_cv2.npz'))
    h, w = [int(v) for v in g['frame_hw']]
    frame = synthetic_frame(h, w, int(g['frame_seed']))
    for box, M_ref, digest, sample in zip(g['boxes'], g['M'], g['sha256'], g['samples']):
        cx, cy, size = oc.box_to_center_size(*box)
        M = oc.affine_from_box(cx, cy, size)
        assert np.array_equal(M, M_ref)                       # bit-exact cv2.getAffineTransform
        patch = oc.warp_affine_u8(frame, M)
        assert np.array_equal(patch[::37, ::41], sample)
        assert hashlib.sha256(patch.tobytes()).hexdigest() == str(digest)   # bit-exact cv2.warpAffine


def test_random_boxes_match_cv2_bit_exact():
    rng = np.random.default_rng(5)
    frame = synthetic_frame(480, 640, 9)
    for i in range(40):
        x1 = float(rng.integers(-60, 600)); y1 = float(rng.integers(-60, 440))
        bw = float(rng.integers(10, 700)); bh = float(rng.integers(10, 700))
        if i % 2:
            x1 += rng.random(); y1 += rng.random(); bw += rng.random(); bh += rng.random()
        box = (x1, y1, x1 + bw, y1 + bh)
        ref, M_ref = reference_patch(frame, box)
        cx, cy, size = oc.box_to_center_size(*box)
        M = oc.affine_from_box(cx, cy, size)
        assert np.array_equal(M, M_ref)
        assert np.array_equal(oc.warp_affine_u8(frame, M), ref)


def test_fully_outside_and_tiny_boxes():
    frame = synthetic_frame(100, 120, 1)
    for box in [(-500., -500., -400., -420.), (1000., 50., 1100., 90.), (10., 10., 12., 11.)]:
        ref, _ = reference_patch(frame, box)
        cx, cy, size = oc.box_to_center_size(*box)
        assert np.array_equal(oc.warp_affine_u8(frame, oc.affine_from_box(cx, cy, size)), ref)


def test_prepare_batch_layout_and_flip():
    frame = synthetic_frame(300, 400, 2)
    boxes = [['right', [50, 60, 180, 220]], ['left', [200, 40, 330, 250]]]
    b = oc.prepare_batch_bbox(frame, boxes)
    assert b['img'].shape == (2, 3, 256, 256) and b['img'].dtype == np.float32
    assert b['box_center'].shape == (2, 2) and b['box_size'].shape == (2,)
    assert b['trans'].shape == (2, 2, 3) and b['do_flip'].tolist() == [0.0, 1.0]
    assert np.array_equal(b['img_size'], np.array([[400, 300], [400, 300]], np.float32))
    # same pipeline written directly with cv2 (infer.py:217-238)
    for i, (label, box) in enumerate(boxes):
        ref, _ = reference_patch(frame, box)
        ref = ref[:, :, ::-1]
        if label != 'right':
            ref = cv2.flip(ref, 1)
        t = np.transpose(ref.copy(), (2, 0, 1)).astype(np.float32)
        mean = 255. * np.array([0.485, 0.456, 0.406]); std = 255. * np.array([0.229, 0.224, 0.225])
        for c in range(3):
            t[c, :, :] = (t[c, :, :] - float(mean[c])) / float(std[c])
        assert np.array_equal(b['img'][i], t)


@pytest.mark.skipif(not os.path.exists(f'{REFERENCE}/hamer/example_data/test1.jpg'),
                    reason='reference example images not available on this box')
def test_example_images_config1():
    """BASELINE config 1 inputs: hamer/example_data/test1..5.jpg with hand-written boxes (the YOLO
    weights are not in the tree)."""
    boxes = {1: [(300, 180, 470, 400), (520, 200, 640, 380)], 2: [(100, 100, 400, 500)],
             3: [(700, 300, 1000, 700)], 4: [(50, 40, 260, 300)], 5: [(900, 500, 1300, 1000)]}
    for i, bl in boxes.items():
        img = cv2.imread(f'{REFERENCE}/hamer/example_data/test{i}.jpg')
        for box in bl:
            ref, _ = reference_patch(img, tuple(float(v) for v in box))
            cx, cy, size = oc.box_to_center_size(*[float(v) for v in box])
            assert np.array_equal(oc.warp_affine_u8(img, oc.affine_from_box(cx, cy, size)), ref)


def test_vitdet_item_against_reference_functions():
    """oracle.crop.vitdet_item (ViTDetDataset.__getitem__, vitdet_dataset.py:48-95) against the reference's own
    expand_to_aspect_ratio / gen_trans_from_patch_cv recipe executed with cv2 (flip inside the warp: the frame is
    mirrored and c_x = W - c_x - 1, datasets/utils.py:347-349)."""
    frame = synthetic_frame(300, 400, 5)
    for box, right in [((50, 60, 180, 220), 1), ((200.5, 40.25, 330, 250), 0), ((-30, 10, 90, 140), 0)]:
        t, center, size = oc.vitdet_item(frame, box, right)
        b = np.asarray(box, np.float32)
        c = (b[2:4] + b[0:2]) / 2.0
        sc = 2.5 * (b[2:4] - b[0:2]) / 200.0
        bs = oc.expand_to_aspect_ratio(sc * 200).max()          # float64 restatement: same value for these boxes
        assert np.float32(bs) == np.float32(size) and np.array_equal(c, center)
        img, cx = frame, c[0]
        if right == 0:
            img = np.ascontiguousarray(frame[:, ::-1, :])
            cx = frame.shape[1] - cx - 1
        src = np.zeros((3, 2), np.float32)
        src[0] = [cx, c[1]]
        src[1] = src[0] + np.array([0, np.float32(size) * 0.5], np.float32)
        src[2] = src[0] + np.array([np.float32(size) * 0.5, 0], np.float32)
        dst = np.array([[128, 128], [128, 256], [256, 128]], np.float32)
        M = cv2.getAffineTransform(src, dst)
        patch = cv2.warpAffine(img, M, (256, 256), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        ref = np.transpose(patch[:, :, ::-1], (2, 0, 1)).astype(np.float32)
        mean = 255. * np.array([0.485, 0.456, 0.406]); std = 255. * np.array([0.229, 0.224, 0.225])
        for ch in range(3):
            ref[ch] = (ref[ch] - np.float32(mean[ch])) / np.float32(std[ch])
        assert np.array_equal(t, ref)


def test_antialias_blur_and_float_warp_against_scipy_and_cv2():
    """The anti-aliased item (vitdet_dataset.py:62-71): oracle.crop.skimage_gaussian states what scipy.ndimage.gaussian_filter
    computes (skimage.filters.gaussian's backend; mode 'nearest', truncate 4) and warp_affine_f64 what cv2.warpAffine does on
    a float64 image - both bit for bit, so the CUDA path (which evaluates the same formulas) is pinned to the libraries the
    reference calls."""
    from scipy import ndimage as ndi
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (300, 420, 3), dtype=np.uint8)
    for sigma in (0.135, 0.28, 0.67, 1.94, 3.3):
        ref = ndi.gaussian_filter(img.astype(np.float64), sigma=(sigma, sigma, 0), mode='nearest', cval=0, truncate=4.0)
        assert np.array_equal(oc.skimage_gaussian(img, sigma), ref)
    big = rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    for cx, cy, size in [(600.0, 300.0, 800.0), (100.0, 100.0, 1200.0), (1200.5, 700.25, 650.3)]:
        sigma = (size / 256 / 2 - 1) / 2
        bl = oc.skimage_gaussian(big, sigma)
        M = oc.affine_from_box(cx, cy, size)
        ref = cv2.warpAffine(bl, M, (256, 256), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        assert np.array_equal(oc.warp_affine_f64(bl, M), ref)
    # and a whole anti-aliased ViTDetDataset item, left hand (flip inside the warp)
    t, center, size = oc.vitdet_item(big, (300, 100, 700, 640), 0)
    assert size / 256 / 2 > 1.1 and t.shape == (3, 256, 256) and t.dtype == np.float32
    fr = ndi.gaussian_filter(big.astype(np.float64), sigma=((size / 256 / 2 - 1) / 2,) * 2 + (0,), mode='nearest', truncate=4.0)[:, ::-1, :]
    M = oc.affine_from_box(float(big.shape[1] - center[0] - 1), float(center[1]), float(size))
    patch = cv2.warpAffine(np.ascontiguousarray(fr), M, (256, 256), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    ref = np.transpose(patch[:, :, ::-1], (2, 0, 1)).astype(np.float32)
    mean = 255. * np.array([0.485, 0.456, 0.406]); std = 255. * np.array([0.229, 0.224, 0.225])
    for ch in range(3):
        ref[ch] = (ref[ch] - np.float32(mean[ch])) / np.float32(std[ch])
    assert np.array_equal(t, ref)
