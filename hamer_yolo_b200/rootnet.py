"""Drop-in mirror of the RootNet / SAR depth call of the reference pipeline:
``EstimateRGB.estimate_root_depth_custom(img, K, bbox)`` (rootnet/Model_RGB.py:572-638; called per hand at
hamer/infer.py:1275).  The box arithmetic below is the reference's (numpy float64 -> float32, torch float32 for k_value);
crop, ResNet-34 trunk and depth head run on the GPU through the C ABI (hm_rootnet_depth)."""
import ctypes as C

import numpy as np
import torch

INPUT_IMG_SHAPE = (256, 256)      # rootnet/sar_config_stage_1.py:17
BBOX_REAL = (0.3, 0.3)            # rootnet/sar_config_stage_1.py:19


def sanitize_bbox(bbox, img_width, img_height):
    """rootnet/preprocessing.py:154-165."""
    x, y, w, h = bbox
    x1 = np.max((0, x))
    y1 = np.max((0, y))
    x2 = np.min((img_width - 1, x1 + np.max((0, w - 1))))
    y2 = np.min((img_height - 1, y1 + np.max((0, h - 1))))
    if w * h > 0 and x2 > x1 and y2 > y1:
        bbox = np.array([x1, y1, x2 - x1, y2 - y1])
    else:
        bbox = None
    return bbox


def process_bbox(bbox, img_width, img_height, input_img_shape, ratio=1.25):
    """rootnet/preprocessing.py:167-188."""
    bbox = sanitize_bbox(bbox, img_width, img_height)
    if bbox is None:
        return bbox
    w = bbox[2]
    h = bbox[3]
    c_x = bbox[0] + w / 2.
    c_y = bbox[1] + h / 2.
    aspect_ratio = input_img_shape[1] / input_img_shape[0]
    if w > aspect_ratio * h:
        h = w / aspect_ratio
    elif w < aspect_ratio * h:
        w = h * aspect_ratio
    bbox[2] = w * ratio
    bbox[3] = h * ratio
    bbox[0] = c_x - bbox[2] / 2.
    bbox[1] = c_y - bbox[3] / 2.
    return bbox.astype(np.float32)


class RootNet:
    """``RootNet(engine, backbone_state_dict, rootnet_state_dict)``: ``backbone_state_dict`` holds the ``backbone.*`` entries of
    the SAR checkpoint's ``'model'`` (keys with or without the ``backbone.`` prefix), ``rootnet_state_dict`` its ``'rootnet'``
    (``depth_layer.weight`` / ``.bias``).  ``engine`` is the hamer_yolo_b200 Engine (``model.engine``)."""

    def __init__(self, engine, backbone_state_dict, rootnet_state_dict):
        self.engine = engine
        n = 0
        for k, v in backbone_state_dict.items():
            if k.endswith('num_batches_tracked'):
                continue
            key = k if k.startswith('backbone.') else 'backbone.' + k
            engine.set_weight('rootnet.' + key, v)
            n += 1
        for k, v in rootnet_state_dict.items():
            engine.set_weight('rootnet.rootnet.' + k, v)
        with torch.cuda.device(engine.device):
            engine.check(engine.lib.hm_rootnet_finalize(engine.h))

    @staticmethod
    def calculate_k(bbox, fx, fy):
        """Model_RGB.py:494-498."""
        area = bbox[-1] * bbox[-2]
        real_area = torch.tensor(BBOX_REAL[0] * BBOX_REAL[1])
        return torch.sqrt(real_area * fx * fy / area).unsqueeze(0)

    def estimate_root_depths(self, img, K, bboxes):
        """Batched form: one frame, several [x1, y1, x2, y2] boxes -> CUDA float32 tensor of root depths."""
        eng = self.engine
        height, width = img.shape[:2]
        if isinstance(K, np.ndarray):
            fx, fy = K[0, 0], K[1, 1]
        else:
            fx, fy = K[0][0], K[1][1]
        proc, kval = [], []
        for bbox in bboxes:
            x1, y1, x2, y2 = bbox
            b = process_bbox([x1, y1, x2 - x1, y2 - y1], width, height, INPUT_IMG_SHAPE, 1.5)      # Model_RGB.py:587-595
            if b is None:
                raise ValueError(f'degenerate hand box {bbox}')
            proc.append(b)
            kval.append(float(self.calculate_k(b, fx, fy)[0]))
        proc = np.ascontiguousarray(np.stack(proc), dtype=np.float32)
        kval = np.ascontiguousarray(kval, dtype=np.float32)
        n = len(proc)
        depth = torch.empty(n, device=eng.device, dtype=torch.float32)
        if isinstance(img, torch.Tensor):
            assert img.is_cuda and img.dtype == torch.uint8 and img.stride(-1) == 1
            fptr, stride, on_dev, keep = img.data_ptr(), img.stride(0), 1, img
        else:
            keep = np.ascontiguousarray(img)
            assert keep.dtype == np.uint8 and keep.ndim == 3 and keep.shape[2] == 3
            fptr, stride, on_dev = keep.ctypes.data, keep.strides[0], 0
        with torch.cuda.device(eng.device):
            eng.check(eng.lib.hm_rootnet_depth(eng.h, C.c_void_p(fptr), height, width, stride, on_dev,
                                               C.c_void_p(proc.ctypes.data), C.c_void_p(kval.ctypes.data), n,
                                               C.c_void_p(depth.data_ptr()), eng.stream()))
        return depth

    def estimate_root_depth_custom(self, img, K, bbox):
        """Model_RGB.py:572-638: frame (HxWx3 BGR uint8), intrinsics (3x3), hand box [x1, y1, x2, y2] -> root depth (float)."""
        return self.estimate_root_depths(img, K, [bbox]).item()
