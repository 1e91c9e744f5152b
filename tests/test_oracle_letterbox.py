"""The detector hand-off oracle against OpenCV itself (what the reference calls): cv2.resize INTER_LINEAR on uint8 is the
fixed-point algorithm oracle.letterbox.resize_linear_u8 states - and hm_letterbox implements - bit for bit."""
import numpy as np
import pytest

from oracle import letterbox as ol

cv2 = pytest.importorskip('cv2')

SIZES = [(1080, 1920, 640, 360), (480, 640, 640, 480), (240, 320, 640, 480), (720, 1280, 640, 360), (1000, 750, 480, 640),
         (333, 517, 640, 412), (100, 100, 640, 640), (37, 1000, 640, 24), (1080, 1920, 960, 540), (2160, 3840, 640, 360)]


@pytest.mark.parametrize('h,w,dw,dh', SIZES)
def test_resize_restatement_matches_cv2(h, w, dw, dh):
    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(ol.resize_linear_u8(img, dw, dh), ref)


def test_letterbox_geometry_of_a_1080p_frame():
    """1920 x 1080 -> r = 1/3 -> 640 x 360, padded to a multiple of 32: 640 x 384 (12 rows of 114 above and below)."""
    img = np.random.default_rng(0).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    out, ratio, (dw, dh) = ol.letterbox(img, 640, stride=32)
    assert out.shape == (384, 640, 3) and ratio == (1 / 3, 1 / 3) and (dw, dh) == (0.0, 12.0)
    assert (out[:12] == 114).all() and (out[-12:] == 114).all()
    assert np.array_equal(out[12:-12], ol.resize_linear_u8(img, 640, 360))


def test_scale_coords_round_trip():
    import torch
    det = torch.tensor([[100.4, 50.2, 300.7, 200.9, 0.9, 1.0], [0.0, 0.0, 700.0, 400.0, 0.5, 0.0]])
    dets = ol.detections_to_list(det, (384, 640), (1080, 1920, 3))
    assert dets[0][0] == 'right' and dets[1][0] == 'left'
    assert dets[0][1] == [301.0, 115.0, 902.0, 567.0]         # (x - 0) * 3, (y - 12) * 3, rounded half to even
    assert dets[1][1] == [0.0, 0.0, 1920.0, 1080.0]           # clipped to the frame


def test_hm_letterbox_shape_matches_the_reference_geometry():
    """hm_letterbox_shape is pure host arithmetic of libhamer_b200.so (no GPU): the padded output shape of
    letterbox(img0, new_size, stride, auto=True) (yolo/yolov7/utils/datasets.py:999-1030) for a sweep of frame sizes."""
    import ctypes as C
    from hamer_yolo_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)
    cases = [(1080, 1920), (480, 640), (720, 1280), (2160, 3840), (100, 100), (37, 1000), (1000, 37), (641, 641), (1, 1)]
    cases += [(int(rng.integers(8, 2200)), int(rng.integers(8, 4000))) for _ in range(60)]
    for new_size, stride in ((640, 32), (960, 32), (416, 64)):
        for h, w in cases:
            oh, ow = C.c_int32(), C.c_int32()
            rc = lib.hm_letterbox_shape(h, w, new_size, stride, C.byref(oh), C.byref(ow))
            img = np.zeros((h, w, 3), np.uint8)
            ref = ol.letterbox(img, new_size, stride=stride)[0].shape[:2]
            assert rc == 0 and (oh.value, ow.value) == ref, (h, w, new_size, stride, (oh.value, ow.value), ref)
