"""Decoder head alone (hm_mano_head) as a CUDA-graph replay: device time per call at the given batch sizes.  The per-launch
events of hm_profile_* run eager and are launch-bound for this stage (50 short kernels), so this is the number that
matches what the head costs inside the replayed forward graph."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hamer_yolo_b200 import load_hamer, synthetic_mano
from hamer_yolo_b200.weights import random_state_dict

dev = torch.device('cuda:0')
model, _ = load_hamer(state_dict=random_state_dict(seed=0, device=dev), mano=synthetic_mano(0), device='cuda:0', max_batch=64)
eng = model.engine
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for B in [int(a) for a in sys.argv[1:]] or [64, 8, 1]:
    feats = torch.randn(B, 192, 1280, device=dev)
    for _ in range(3):
        eng.mano_head(feats)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = eng.mano_head(feats)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(10):
        flush.zero_()                    # the features (63 MB at 64 crops) start in HBM, as after the last LayerNorm of a step
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f'B={B}: head graph replay {tot / 10 * 1e3:.1f} us cold L2, {e0.elapsed_time(e1) / 20 * 1e3:.1f} us back to back '
          f'({eng.launch_count()} launches per forward)')
