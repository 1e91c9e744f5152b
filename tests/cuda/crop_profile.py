"""hm_preprocess alone on a device-resident 1080p frame (crop / warp / normalise kernel) - event-timed and for ncu captures.
usage: python tests/cuda/crop_profile.py [crops per call (default 64)] [box side in pixels (default 320)]
Algorithmic bytes per crop: 3 x 256 x 256 x 4 written + the 3 x (1.2 box)^2 source bytes of the window read."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hamer_yolo_b200.hamer import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
side = float(sys.argv[2]) if len(sys.argv) > 2 else 320.0
eng = Engine('cuda:0', max_batch=max(n, 2))
rng = np.random.default_rng(0)
H, W = 1080, 1920
frame = torch.from_numpy(rng.integers(0, 256, (H, W, 3), dtype=np.uint8)).to(eng.device)
cx = rng.uniform(side, W - side, n)
cy = rng.uniform(side / 2, H - side / 2, n)
wh = side * rng.uniform(0.8, 1.0, (n, 2))
boxes = np.stack([cx - wh[:, 0] / 2, cy - wh[:, 1] / 2, cx + wh[:, 0] / 2, cy + wh[:, 1] / 2], 1)
right = rng.integers(0, 2, n)
out = torch.empty(n, 3, 256, 256, device=eng.device)
for _ in range(3):
    eng.preprocess(frame, boxes, right, out=out)
torch.cuda.synchronize()
eng.profile_enable(True)
for _ in range(20):
    eng.preprocess(frame, boxes, right, out=out)
prof = eng.profile_collect()
eng.profile_enable(False)
ms = prof['crop'][0] / 20
win = 3.0 * (1.2 * np.maximum(wh[:, 0], wh[:, 1] * 0.75) * np.array([1.0])[0]) ** 2      # window of expand_to_aspect_ratio x 1.2 rescale (approx.)
bytes_call = n * 3 * 256 * 256 * 4 + float(win.sum())
print(f'hm_preprocess {n} crops of a 1080p device frame (boxes ~{side:.0f} px): {ms * 1e3:.1f} us per call (device time, '
      f'includes the per-call parameter upload)  {bytes_call / ms / 1e6:.1f} GB/s algorithmic')
