"""CPU-side checks: the C-ABI library loads and exports everything include/hamer_b200.h declares,
host-side logic (config, state-dict layout, MANO loader, sharding / packing, 2-rank gloo gather)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT


def header_functions():
    txt = open(os.path.join(ROOT, 'include', 'hamer_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(hm_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    from hamer_yolo_b200 import _lib
    lib = _lib.load()
    declared = header_functions()
    assert sorted(_lib.EXPORTS) == declared, (set(_lib.EXPORTS) ^ set(declared))
    for name in declared:
        assert getattr(lib, name) is not None
    assert b'sm_100a' in lib.hm_version()


def test_struct_layouts_match_header():
    from hamer_yolo_b200 import _lib
    assert ctypes.sizeof(_lib.HmConfig) == 8 * 4
    assert ctypes.sizeof(_lib.HmOutputs) == 10 * ctypes.sizeof(ctypes.c_void_p)
    txt = open(os.path.join(ROOT, 'include', 'hamer_b200.h')).read()
    body = txt[txt.index('typedef struct hm_outputs {'):txt.index('} hm_outputs;')]
    fields = re.findall(r'float\*\s+(\w+);', body)
    assert fields == [n for n, _ in _lib.HmOutputs._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_create_fails_loudly_without_gpu():
    from hamer_yolo_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    cfg = _lib.HmConfig(max_batch=1)
    rc = lib.hm_create(ctypes.byref(h), 0, ctypes.byref(cfg))
    assert rc < 0 and not h.value
    assert len(lib.hm_last_error(None)) > 0
    from hamer_yolo_b200.hamer import Engine
    with pytest.raises(RuntimeError):
        Engine('cpu')


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'hamer_yolo_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_state_dict_layout_matches_reference_tree(oracle_model):
    """weights.state_dict_shapes() == the oracle module tree == the reference module tree
    (make_golden.py asserts oracle == reference)."""
    from hamer_yolo_b200.weights import state_dict_shapes, random_state_dict
    sd = {k: v for k, v in oracle_model.state_dict().items()}
    shapes = state_dict_shapes()
    assert set(shapes) == set(sd)
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == tuple(s), k
    rs = random_state_dict(seed=1, device='cpu')
    assert set(rs) == set(shapes)
    w = rs['backbone.blocks.3.mlp.fc1.weight']
    assert torch.equal(w, w.to(torch.bfloat16).float()) and w.abs().max() <= 1 / np.sqrt(1280) + 1e-3


def test_config_mirror():
    from hamer_yolo_b200.configs import default_config
    cfg = default_config()
    assert cfg.MODEL.IMAGE_SIZE == 256 and cfg.MODEL.BBOX_SHAPE == [192, 256]
    assert cfg.EXTRA.FOCAL_LENGTH == 5000 and cfg.MANO.NUM_HAND_JOINTS == 15
    assert cfg.MODEL.get('BBOX_SHAPE', None) == [192, 256]
    assert cfg.MODEL.MANO_HEAD.TRANSFORMER_DECODER.context_dim == 1280


def test_synthetic_mano_invariants():
    from hamer_yolo_b200.mano_data import synthetic_mano
    m = synthetic_mano(0)
    assert m['v_template'].shape == (778, 3) and m['posedirs'].shape == (135, 2334)
    assert np.allclose(m['lbs_weights'].sum(1), 1, atol=1e-6) and (m['lbs_weights'] != 0).sum(1).max() <= 8
    assert np.allclose(m['J_regressor'].sum(1), 1, atol=1e-5)
    assert all(m['parents'][j] < j for j in range(16))
    m2 = synthetic_mano(0)
    assert all(np.array_equal(m[k], m2[k]) for k in m)


def test_shard_range_and_pack_roundtrip():
    from hamer_yolo_b200 import dist as hd
    for n, w in [(512, 8), (512, 4), (10, 4), (3, 8), (64, 1)]:
        spans = [hd.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    B = 5
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g)
    out = {'pred_vertices': r(B, 778, 3), 'pred_keypoints_3d': r(B, 21, 3), 'pred_keypoints_2d': r(B, 21, 2),
           'pred_cam': r(B, 3), 'pred_cam_t': r(B, 3),
           'pred_mano_params': {'global_orient': r(B, 1, 3, 3), 'hand_pose': r(B, 15, 3, 3), 'betas': r(B, 10)}}
    rec = hd.pack_outputs(out)
    assert rec.shape == (B, hd.RECORD_FLOATS) and hd.RECORD_FLOATS * 4 == 10396
    back = hd.unpack_outputs(rec)
    for k in ('pred_vertices', 'pred_keypoints_3d', 'pred_keypoints_2d', 'pred_cam', 'pred_cam_t'):
        assert torch.equal(back[k], out[k])
    for k in ('global_orient', 'hand_pose', 'betas'):
        assert torch.equal(back['pred_mano_params'][k], out['pred_mano_params'][k])


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from hamer_yolo_b200 import dist as hd
dist.init_process_group('gloo', init_method='env://')
rank, world = dist.get_rank(), dist.get_world_size()
n = int(sys.argv[2])
full = torch.arange(n * hd.RECORD_FLOATS, dtype=torch.float32).reshape(n, hd.RECORD_FLOATS)
lo, hi = hd.shard_range(n, rank, world)
counts = [hd.shard_range(n, r, world)[1] - hd.shard_range(n, r, world)[0] for r in range(world)]
got = hd.all_gather_records(full[lo:hi].clone(), counts)
assert torch.equal(got, full), (rank, got.shape)
dist.barrier()
dist.destroy_process_group()
print('ok', rank)
'''


@pytest.mark.parametrize('n', [8, 7])
def test_allgather_world2_gloo(tmp_path, n):
    """N>1 host path on CPU: two gloo ranks shard n hands (equal and ragged) and all-gather."""
    script = tmp_path / 'w.py'
    script.write_text(_GLOO_WORKER)
    port = 29500 + (os.getpid() % 500) + n
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(n)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out.decode()[-2000:]


_GLOO_SHARDED_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from hamer_yolo_b200 import dist as hd
dist.init_process_group('gloo', init_method='env://')
rank, world = dist.get_rank(), dist.get_world_size()
n = int(sys.argv[2])

def fake_model(batch):
    """Stands in for HAMER.forward on CPU: every output is a function of the crop's first pixel; like hm_forward it
    rejects an empty batch."""
    x = batch['img']
    B = x.shape[0]
    if B == 0:
        raise RuntimeError('batch size must be positive')
    v = x[:, 0, 0, 0]
    f = lambda *s: v.reshape(B, *([1] * len(s))).expand(B, *s).clone()
    out = {'pred_vertices': f(778, 3), 'pred_keypoints_3d': f(21, 3) + 1, 'pred_keypoints_2d': f(21, 2) + 2,
           'pred_cam': f(3) + 3, 'pred_cam_t': f(3) + 4,
           'pred_mano_params': {'global_orient': f(1, 3, 3) + 5, 'hand_pose': f(15, 3, 3) + 6, 'betas': f(10) + 7}}
    return out, None

img = torch.arange(n, dtype=torch.float32).reshape(n, 1, 1, 1).expand(n, 3, 4, 4).contiguous()
got = hd.forward_sharded(fake_model, img)
want, _ = fake_model({'img': img})
assert got['pred_vertices'].shape == (n, 778, 3)
for k in ('pred_vertices', 'pred_keypoints_3d', 'pred_keypoints_2d', 'pred_cam', 'pred_cam_t'):
    assert torch.equal(got[k], want[k]), k
for k in ('global_orient', 'hand_pose', 'betas'):
    assert torch.equal(got['pred_mano_params'][k], want['pred_mano_params'][k]), k
dist.barrier()
dist.destroy_process_group()
print('ok', rank)
'''


@pytest.mark.parametrize('n', [1, 5])
def test_forward_sharded_world2_gloo(tmp_path, n):
    """forward_sharded on two gloo ranks; n = 1 leaves rank 1 with an empty shard (fewer hands than ranks): it must
    skip the model call and still join the all-gather instead of raising while rank 0 blocks."""
    script = tmp_path / 'ws.py'
    script.write_text(_GLOO_SHARDED_WORKER)
    port = 30100 + (os.getpid() % 500) + n
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(n)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out.decode()[-2000:]


def test_npy_record_layout(tmp_path):
    """infer.py:1296-1312 layout, written and read back the way infer.py:1350 reads it."""
    pytest.importorskip('cv2')
    from hamer_yolo_b200.inference import hand_record, save_image_results
    from oracle.head import rot6d_to_rotmat
    R = rot6d_to_rotmat(torch.randn(16, 6, generator=torch.Generator().manual_seed(1))).reshape(1, 16, 3, 3)
    out = {'pred_mano_params': {'global_orient': R[:, :1], 'hand_pose': R[:, 1:], 'betas': torch.randn(1, 10)},
           'pred_cam_t_full': torch.tensor([[0.1, -0.2, 12.0]])}
    d = hand_record(out, 0, True)
    assert set(d) == {'betas', 'theta', 'pose_hand', 'pose_global', 'cam_t', 'is_right'}
    assert d['betas'].shape == (10,) and d['theta'].shape == (48,) and d['pose_hand'].shape == (45,)
    assert d['pose_global'].shape == (3,) and d['cam_t'].shape == (3,) and d['is_right'] is True
    assert np.array_equal(d['theta'], np.concatenate([d['pose_global'], d['pose_hand']]))
    path = save_image_results(str(tmp_path), 'frame0', {'left': None, 'right': d})
    back = np.load(path, allow_pickle=True).item()
    assert back['left'] is None and np.array_equal(back['right']['theta'], d['theta'])
    # axis-angle round trip through cv2.Rodrigues
    import cv2
    assert np.allclose(cv2.Rodrigues(d['pose_global'])[0], R[0, 0].numpy(), atol=1e-6)


@pytest.mark.parametrize('B', [1, 2, 3, 5, 7, 8, 64])
def test_graph_output_buffer_fields_are_aligned(B):
    """The packed output buffer of a CUDA graph: every field 64-byte aligned for every batch size, fields disjoint, shapes as
    hm_outputs expects.  (With tightly packed fields an odd batch put pred_vertices 4 bytes off and hm_mano_lbs silently took its
    three-kernel path: same tolerances, different bits.)"""
    from hamer_yolo_b200.hamer import Engine
    flat = torch.zeros(Engine._flat_floats(B), dtype=torch.float32)
    views = Engine._views(Engine, flat, B)
    spans = []
    for name, shp in Engine._FIELDS:
        v = views[name]
        assert tuple(v.shape) == (B, *shp) and v.is_contiguous()
        off = (v.data_ptr() - flat.data_ptr())
        assert off % 64 == 0, (name, off)
        spans.append((off, off + v.numel() * 4))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    assert spans[-1][1] <= flat.numel() * 4
