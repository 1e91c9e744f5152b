"""hamer_yolo_b200 - B200-native (sm_100a) implementation of HaMeR's per-crop hand reconstruction
hot path behind the reference's Python API (2646207530/hamer-yolo).

    from hamer_yolo_b200 import load_hamer, hamer_inference
    model, cfg = load_hamer(ckpt_path, mano=load_mano_pkl('.../MANO_RIGHT.pkl'))
    out, params = model({'img': img})          # same dicts as hamer/hamer/models/hamer.py:124-156

Everything numerical runs in hand-written CUDA kernels inside ``libhamer_b200.so`` (C ABI in
include/hamer_b200.h); there is no CPU or library fallback.
"""
from .configs import CfgNode, default_config
from .mano_data import load_mano_pkl, synthetic_mano

__all__ = ['CfgNode', 'default_config', 'load_mano_pkl', 'synthetic_mano', 'HAMER', 'load_hamer',
           'hamer_inference', 'Engine', 'HAMER_INFER', 'ViTDetDataset', 'RootNet']


def __getattr__(name):
    # torch / the CUDA library are imported lazily so that config + data helpers work without them
    if name in ('HAMER', 'load_hamer', 'Engine', 'HAMER_INFER'):
        from . import hamer
        return getattr(hamer, name)
    if name == 'ViTDetDataset':
        from .datasets import ViTDetDataset
        return ViTDetDataset
    if name == 'RootNet':
        from .rootnet import RootNet
        return RootNet
    if name == 'hamer_inference':
        from .inference import hamer_inference
        return hamer_inference
    raise AttributeError(name)
