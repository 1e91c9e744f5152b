"""Diagnostics: does a hand's MANO output depend on the batch it is computed in?  (GPU box)"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hamer_yolo_b200.hamer import Engine
from hamer_yolo_b200 import synthetic_mano
from test_gpu_parity import random_mano_inputs
eng = Engine('cuda:0', max_batch=2)
eng.set_mano(synthetic_mano(0))
dev = eng.device
betas, rot, cam_t = random_mano_inputs(40, 5)
betas, rot, cam_t = betas.to(dev), rot.to(dev), cam_t.to(dev)
ref = [t.clone() for t in eng.mano_lbs(betas, rot, cam_t)]
for B in (1, 2, 3, 4, 5, 6, 7, 8, 9, 16, 31, 32, 33):
    v, j, k = eng.mano_lbs(betas[:B].contiguous(), rot[:B].contiguous(), cam_t[:B].contiguous())
    dv = (v - ref[0][:B]).abs().amax(dim=(1, 2))
    print(f'B={B:3d} verts max diff {dv.max().item():.3e} (hands differing: {(dv > 0).nonzero().flatten().tolist()[:12]})  joints {(j - ref[1][:B]).abs().max().item():.3e}  kp2d {(k - ref[2][:B]).abs().max().item():.3e}')
    if dv.max().item() > 0:
        h = int(dv.argmax()); dd = (v[h] - ref[0][h]).abs().amax(dim=1)
        print('      hand', h, 'vertices differing', int((dd > 0).sum()), 'first', (dd > 0).nonzero().flatten().tolist()[:10])
