"""Where a video frame's latency goes (BASELINE configs[4] loop of bench_video.py): host frame vs device-resident frame vs
pinned host frame through estimate_from_rgb, and the forward alone at the same hand counts."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hamer_yolo_b200 import load_hamer, synthetic_mano, hamer_inference
from hamer_yolo_b200.weights import random_state_dict
from bench_video import synthetic_frame
dev = torch.device('cuda', 0)
model, _ = load_hamer(state_dict=random_state_dict(seed=0, device=dev), mano=synthetic_mano(0), device=dev, max_batch=8)
hi = hamer_inference(model=model)
frames = [synthetic_frame(1080, 1920, s) for s in range(4)]
pinned = [torch.from_numpy(f).pin_memory() for f in frames]
on_dev = [torch.from_numpy(f).to(dev) for f in frames]
k_real = np.array([[1400., 0, 960.], [0, 1400., 540.], [0, 0, 1]], np.float32)
for n in range(2, 9):
    hi.estimate_from_rgb(frames[0], [['right', [100.0 * n, 100.0, 100.0 * n + 150, 300.0]]] * n, k_real)
def loop(kind, nframes=200):
    rng = np.random.default_rng(0)
    lat = []
    for i in range(nframes + 20):
        n = int(rng.integers(2, 9))
        dets = []
        for _ in range(n):
            s = float(rng.integers(80, 401))
            x1 = float(rng.integers(0, 1920 - 80)); y1 = float(rng.integers(0, 1080 - 80))
            dets.append(['right' if rng.random() < 0.5 else 'left', [x1, y1, x1 + s, y1 + s * float(rng.uniform(0.8, 1.3))]])
        f = {'host': frames, 'pinned': [p.numpy() for p in pinned], 'device': on_dev}[kind][i % 4]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if kind == 'forward':
            pass
        out, _ = hi.estimate_from_rgb(f, dets, k_real)
        v = out['pred_vertices'].cpu(); k = out['pred_keypoints_2d_full'].cpu()
        dt = (time.perf_counter() - t0) * 1e3
        if i >= 20:
            lat.append(dt)
    lat = np.array(lat)
    print(f'{kind:8s} frame: p50 {np.percentile(lat, 50):.3f} ms  p99 {np.percentile(lat, 99):.3f}  mean {lat.mean():.3f}')
for kind in ('host', 'pinned', 'device', 'host'):
    loop(kind)
eng = model.engine
for B in (2, 5, 8):
    img = torch.randn(B, 3, 256, 256, device=dev)
    for _ in range(3): eng.forward(img)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50):
        eng.forward(img); torch.cuda.synchronize()
    print(f'forward B={B}: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per call + sync')
