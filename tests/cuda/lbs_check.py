"""hm_mano_lbs against the oracle at a few batch sizes (quick GPU-box check while developing the kernel).
usage: python -u tests/cuda/lbs_check.py [B ...]"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from hamer_yolo_b200.hamer import Engine
from hamer_yolo_b200 import synthetic_mano
from oracle.mano import mano_forward
from oracle.pipeline import perspective_projection
from test_gpu_parity import random_mano_inputs

sizes = [int(a) for a in sys.argv[1:]] or [1, 2, 16, 37, 300, 1000]
mano = synthetic_mano(0)
eng = Engine('cuda:0', max_batch=2)
eng.set_mano(mano)
dev = eng.device
for B in sizes:
    betas, rot, cam_t = random_mano_inputs(B, 100 + B)
    print('B', B, 'launch', flush=True)
    v, j, k = eng.mano_lbs(betas.to(dev), rot.to(dev), cam_t.to(dev))
    torch.cuda.synchronize()
    n = min(B, 512)
    v_ref, j_ref = mano_forward(rot[:n, :1], rot[:n, 1:], betas[:n], mano)
    k_ref = perspective_projection(j_ref, cam_t[:n], torch.full((n, 2), 5000.0 / 256))
    ev = (v[:n].cpu() - v_ref).norm(dim=-1)
    print(f'  verts max {ev.max().item() * 1e3:.5f} mm (vertex {ev.max(0).values.argmax().item()}, hand {ev.max(1).values.argmax().item()})'
          f'  joints {(j[:n].cpu() - j_ref).norm(dim=-1).max().item() * 1e3:.5f} mm  kp2d {(k[:n].cpu() - k_ref).abs().max().item() * 256:.5f} px',
          flush=True)
    if B > n:
        v2, j2, _ = eng.mano_lbs(betas[B - 40:].to(dev), rot[B - 40:].to(dev), cam_t[B - 40:].to(dev))
        print('  tail vs re-run max diff', (v[B - 40:] - v2).abs().max().item(), (j[B - 40:] - j2).abs().max().item(), flush=True)
print('done', flush=True)
