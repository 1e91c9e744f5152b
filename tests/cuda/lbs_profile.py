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
