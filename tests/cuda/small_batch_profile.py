"""Per-stage device time of hm_forward at small batch sizes (the video-stream regime, BASELINE configs[4])."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hamer_yolo_b200 import load_hamer, synthetic_mano
from hamer_yolo_b200.weights import random_state_dict

dev = torch.device('cuda:0')
model, _ = load_hamer(state_dict=random_state_dict(seed=0, device=dev), mano=synthetic_mano(0), device="cuda:0", max_batch=max([int(a) for a in sys.argv[1:]] or [16]))
eng = model.engine
for B in [int(a) for a in sys.argv[1:]] or [2, 5, 8]:
    img = torch.randn(B, 3, 256, 256, device=dev)
    for _ in range(3):
        eng.forward(img)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.forward(img)
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / 20
    eng.profile_enable(True)
    for _ in range(2):
        eng.forward(img)
    eng.profile_collect()
    for _ in range(10):
        eng.forward(img)
    prof = eng.profile_collect()
    eng.profile_enable(False)
    print(f'B={B}: graph replay {graph_ms:.3f} ms/forward; stages (ms/forward): ' +
          ', '.join(f'{k} {v[0] / 10:.3f}' for k, v in prof.items() if v[1]))
