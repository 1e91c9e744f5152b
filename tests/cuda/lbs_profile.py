"""Runs hm_mano_lbs alone at BASELINE config[3] size (65536 hands) - for ncu launch lists."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hamer_yolo_b200.hamer import Engine
from hamer_yolo_b200 import synthetic_mano

nh = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
eng = Engine('cuda:0', max_batch=2)
eng.set_mano(synthetic_mano(0))
dev = eng.device
g = torch.Generator(device=dev).manual_seed(5)
betas = torch.randn(nh, 10, generator=g, device=dev)
r6 = torch.randn(nh * 16, 6, generator=g, device=dev)
b1 = torch.nn.functional.normalize(r6[:, :3])
b2 = torch.nn.functional.normalize(r6[:, 3:] - (b1 * r6[:, 3:]).sum(-1, keepdim=True) * b1)
rot = torch.stack((b1, b2, torch.linalg.cross(b1, b2)), dim=-1).reshape(nh, 16, 3, 3).contiguous()
for _ in range(3):
    v, j = eng.mano_lbs(betas, rot)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    v, j = eng.mano_lbs(betas, rot)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f'mano_lbs {nh} hands: {ms:.4f} ms  {nh * 10204 / ms / 1e6:.1f} GB/s')
if os.environ.get('HM_MANO_CLOCKS'):
    import ctypes as C
    import numpy as np
    n_ctas = 148
    clk = np.zeros((n_ctas, 8, 6), np.int64)
    eng.check(eng.lib.hm_test_mano_clocks(eng.h, C.c_void_p(betas.data_ptr()), C.c_void_p(rot.data_ptr()), nh, C.c_void_p(v.data_ptr()),
                                          C.c_void_p(j.data_ptr()), C.c_void_p(clk.ctypes.data), n_ctas, eng.stream()))
    d = np.diff(clk, axis=2).astype(np.float64)
    names = ['wait prev unit (bar)', 'wait inputs', 'prologue', 'joint flush + bar', 'tiles']
    live = clk[:, :, 5] > 0
    for u in range(8):
        m = live[:, u]
        if m.sum() == 0:
            continue
        print(f'unit {u} ({int(m.sum())} CTAs): ' + ', '.join(f'{names[k]} {np.median(d[m, u, k]):.0f}' for k in range(5)))
    tot = clk[:, :, 5].max(axis=1) - clk[:, 0, 0]
    print('per-CTA first-unit-start to last-unit-end: median', np.median(tot), 'max', tot.max())
