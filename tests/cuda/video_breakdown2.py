"""Stage-by-stage host timing of one video frame: preprocess (+ sync), forward (+ sync), for pageable / pinned / device frames."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hamer_yolo_b200 import load_hamer, synthetic_mano
from hamer_yolo_b200.weights import random_state_dict
from bench_video import synthetic_frame
dev = torch.device('cuda', 0)
model, _ = load_hamer(state_dict=random_state_dict(seed=0, device=dev), mano=synthetic_mano(0), device=dev, max_batch=8)
eng = model.engine
frames = [synthetic_frame(1080, 1920, s) for s in range(4)]
pinned = [torch.from_numpy(f).pin_memory().numpy() for f in frames]
on_dev = [torch.from_numpy(f).to(dev) for f in frames]
for n in range(2, 9):
    eng.forward(torch.randn(n, 3, 256, 256, device=dev))
for kind, src in (('host', frames), ('pinned', pinned), ('device', on_dev), ('host', frames)):
    rng = np.random.default_rng(0)
    tp, tps, tf = [], [], []
    for i in range(220):
        n = int(rng.integers(2, 9))
        boxes = []
        for _ in range(n):
            s = float(rng.integers(80, 401)); x1 = float(rng.integers(0, 1840)); y1 = float(rng.integers(0, 1000))
            boxes.append([x1, y1, x1 + s, y1 + s * float(rng.uniform(0.8, 1.3))])
        right = rng.integers(0, 2, n)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        img, *_ = eng.preprocess(src[i % 4], np.array(boxes), right)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out = eng.forward(img)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if i >= 20:
            tp.append(t1 - t0); tps.append(t2 - t0); tf.append(t3 - t2)
    f = lambda a: np.percentile(np.array(a) * 1e3, 50)
    print(f'{kind:7s} preprocess call {f(tp):.3f} ms, + sync {f(tps):.3f} ms, forward + sync {f(tf):.3f} ms')
