#!/usr/bin/env python
"""Benchmark of the HaMeR per-crop hot path (BASELINE.json metric: hand crops/sec, ViT-H + MANO,
256x192 crops).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 64]

A "step" = one pass of the hot path over one batch of synthetic crops per GPU:
ViT-H/16 backbone -> MANO decoder head -> MANO LBS -> projection (+ for N > 1 the NCCL all-gather of
the packed per-hand outputs).  Workload at N = 1 is BASELINE config[1] (batch 64); for N > 1 each rank
keeps 64 crops (weak scaling; N = 8 is config[2]: 512 crops, 64 per GPU).  Weights are random-init of
the reference architecture, data is synthetic (no network for checkpoints / datasets).

Rank 0 prints ONE JSON line (see the keys at the bottom).  `--impl reference` times the CPU oracle
(the restatement of the reference's PyTorch path - the reference itself needs packages that are not
installable here and does not travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'hand crops/sec (ViT-H+MANO, 256x192)'
UNIT = 'crops/s'
GFLOP_BACKBONE = 248.01          # per crop, SURVEY 8d / BASELINE.md section 3
GFLOP_HEAD_REF = 3.085           # reference formulation of the decoder head
LBS_BYTES_PER_HAND = 10204       # 616 in + 9336 verts + 252 joints
# algorithmic FLOPs per crop of the four tcgen05 GEMM families + patch embed (2*M*N*K, M = 192)
GEMM_GFLOP = {'gemm_patch': 2 * 192 * 1280 * 768 / 1e9, 'gemm_qkv': 32 * 2 * 192 * 3840 * 1280 / 1e9,
              'gemm_proj': 32 * 2 * 192 * 1280 * 1280 / 1e9, 'gemm_fc1': 32 * 2 * 192 * 5120 * 1280 / 1e9,
              'gemm_fc2': 32 * 2 * 192 * 1280 * 5120 / 1e9}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'],
                'tf_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap'}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons)}


def build_oracle(seed=0):
    import torch
    from oracle.pipeline import OHamer
    from hamer_yolo_b200.mano_data import synthetic_mano
    torch.manual_seed(seed)
    return OHamer(synthetic_mano(0))


def time_oracle_cpu(sample_crops, iters, warmup=1):
    """The reference's CPU PyTorch path (oracle restatement), fp32, all host threads."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_oracle()
    g = torch.Generator().manual_seed(1)
    img = torch.randn(sample_crops, 3, 256, 256, generator=g)
    for _ in range(warmup):
        model({'img': img})
    t0 = time.perf_counter()
    for _ in range(iters):
        model({'img': img})
    dt = (time.perf_counter() - t0) / iters
    return sample_crops / dt, dt, cores


def gpu_eager_baseline(dev, sd, batch, iters=3):
    """The "cuBLAS via PyTorch" bar: the reference's PyTorch path (oracle restatement,
    same weights) run by torch eager on this B200 - fp32 (TF32 off, the reference default) and fp16 autocast."""
    import torch
    from oracle.pipeline import OHamer
    from hamer_yolo_b200.mano_data import synthetic_mano
    model = OHamer(synthetic_mano(0))
    model.load_state_dict({k: v.to('cpu') for k, v in sd.items()})
    model = model.to(dev)
    g = torch.Generator(device=dev).manual_seed(3)
    img = torch.randn(batch, 3, 256, 256, generator=g, device=dev)
    out = {}
    for name, ctx in (('fp32', None), ('fp16_autocast', torch.autocast('cuda', dtype=torch.float16))):
        def once():
            if ctx is None:
                return model({'img': img})
            with ctx:
                return model({'img': img})
        for _ in range(2):
            once()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            o, _ = once()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / iters
        out[name] = {'value': batch / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms}
    out['batch'] = batch
    out['what'] = 'oracle restatement of the reference PyTorch path, torch eager on this GPU (cuBLAS / ATen kernels), device-resident inputs'
    del model
    torch.cuda.empty_cache()
    return out


def config0_pass(model, sd):
    """BASELINE configs[0] shape of work (infer.py on example images: one image at a time, 1-2 boxes, batch = hands of the
    image) on five synthetic 480x640 frames with fixed boxes (the example JPEGs do not travel to the GPU box): ms per
    image through hamer_inference.estimate_from_rgb, checked against the CPU oracle on the same crops."""
    import numpy as np
    import torch
    from bench_video import synthetic_frame
    from hamer_yolo_b200 import hamer_inference
    from oracle import crop as oc
    from oracle.pipeline import OHamer
    from hamer_yolo_b200.mano_data import synthetic_mano
    hi = hamer_inference(model=model)
    ora = OHamer(synthetic_mano(0))
    ora.load_state_dict({k: v.to('cpu') for k, v in sd.items()})
    cases = [[['right', [100, 80, 260, 300]], ['left', [330, 120, 520, 420]]], [['left', [40, 60, 300, 400]]],
             [['right', [300, 200, 420, 330]], ['right', [60, 50, 180, 220]]], [['left', [200, 100, 460, 440]]],
             [['right', [380, 260, 600, 470]]]]
    frames = [synthetic_frame(480, 640, 40 + i) for i in range(5)]
    for f, d in zip(frames, cases):                                   # warm-up (graphs per batch size)
        hi.estimate_from_rgb(f, d)
    torch.cuda.synchronize()
    lat, err = [], 0.0
    for f, d in zip(frames, cases):
        t0 = time.perf_counter()
        out, _ = hi.estimate_from_rgb(f, d)
        v = out['pred_vertices'].cpu()
        lat.append((time.perf_counter() - t0) * 1e3)
        batch = {k: torch.from_numpy(x) for k, x in oc.prepare_batch_bbox(f, d).items()}
        ref, _ = ora(batch)
        err = max(err, (v - ref['pred_vertices']).norm(dim=-1).max().item() * 1e3)
    return {'images': 5, 'hands': sum(len(d) for d in cases), 'ms_per_image': float(np.mean(lat)),
            'max_vertex_error_mm_vs_cpu_oracle': err,
            'what': 'BASELINE configs[0] shape: one image at a time, batch = hands of the image, host frame in / vertices out'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample = 8
    # warm-up steps and timed steps are each one bounded sample (8 crops) of the workload
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_oracle()
    g = torch.Generator().manual_seed(1)
    img = torch.randn(sample, 3, 256, 256, generator=g)
    for _ in range(max(1, min(args.warmup, 2))):
        model({'img': img})
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        model({'img': img})
    dt = (time.perf_counter() - t0) / steps
    value = sample / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': {'workload': 'HaMeR ViT-H/16 + MANO decoder head + MANO LBS on synthetic 256x192 crops '
                                   '(random-init weights); CPU oracle of the reference PyTorch path',
                       'batch_per_step': sample},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': f'{steps} steps x {sample} crops, fp32, torch CPU, {cores} threads'},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from hamer_yolo_b200 import load_hamer, synthetic_mano
    from hamer_yolo_b200 import dist as hd
    from hamer_yolo_b200.weights import random_state_dict

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B = args.batch
    if args.scaling == 'strong':                   # BASELINE configs[2]: 512 crops in total, sharded over the ranks
        B = max(1, args.global_batch // world)
    sd = random_state_dict(seed=0, device=dev)
    model, _ = load_hamer(state_dict=sd, mano=synthetic_mano(0), device=dev, max_batch=B)
    eng = model.engine
    eng.use_graphs = not args.no_graphs
    collective = 'none'
    if world > 1:
        try:                                       # the library's own communicator: hm_allgather = ncclAllGather from the C ABI
            hd.init_comm(eng)
            collective = 'hm_allgather (ncclAllGather issued by libhamer_b200.so)'
        except Exception as exc:                   # noqa: BLE001
            collective = f'torch.distributed all_gather_into_tensor (hm_comm_init failed: {exc})'[:200]
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    img = torch.randn(B, 3, 256, 256, generator=g, device=dev)
    img_host = img.cpu().pin_memory()
    counts = [B] * world

    def step(x):
        out, _ = model({'img': x})
        rec = hd.pack_outputs(out)
        if world > 1:
            rec = hd.all_gather_records(rec, counts, engine=eng)
        return rec

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident throughput ("value") ----------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        rec = step(img)
    assert torch.isfinite(rec).all(), 'non-finite outputs'
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        rec = step(img)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    # Second timed pass of the same K steps with one CUDA-event pair around every kernel launch (library
    # side, on the launch stream): per-kernel durations for the roofline.  Kept apart from the pass above
    # because the ~620 event records per step cost ~4 % and cannot be captured in a CUDA graph.
    eng.profile_enable(True)
    for _ in range(2):
        step(img)
    eng.profile_collect()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        rec = step(img)
    p1.record()
    sync_all()
    sampler.stop_flag = True
    ms_profiled = p0.elapsed_time(p1) / args.steps
    prof = eng.profile_collect()
    eng.profile_enable(False)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers ("e2e") ------------------------------
    # Every step copies its inputs from pinned host memory (50 MB) and reads its packed result back to
    # the host.  The H2D copy of step i+1 runs on a copy stream while step i computes (two device input
    # buffers), the D2H read is asynchronous into pinned memory and is consumed one step later.
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    bufs = [torch.empty_like(img) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    host_out = [torch.empty(world * B, hd.RECORD_FLOATS, dtype=torch.float32).pin_memory() for _ in range(2)]
    read_done = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(n):
        with torch.cuda.stream(copy_stream):
            bufs[0].copy_(img_host, non_blocking=True)
            copied[0].record(copy_stream)
        for i in range(n):
            cur, nxt = i % 2, (i + 1) % 2
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(consumed[nxt])       # step i-1 finished reading bufs[nxt]
                    bufs[nxt].copy_(img_host, non_blocking=True)
                    copied[nxt].record(copy_stream)
            main.wait_event(copied[cur])
            rec = step(bufs[cur])
            consumed[cur].record(main)
            if i >= 2:
                read_done[cur].synchronize()                         # host buffer free again (result of step i-2 read)
            host_out[cur].copy_(rec, non_blocking=True)
            read_done[cur].record(main)
        read_done[(n - 1) % 2].synchronize()
        return host_out[(n - 1) % 2]

    e2e_run(2)
    sync_all()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    host_rec = e2e_run(args.steps)
    t1.record()
    sync_all()
    ms_e2e = t0.elapsed_time(t1)
    assert torch.isfinite(host_rec).all()
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item()
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    gemm_ms = sum(prof[k][0] for k in GEMM_GFLOP)
    gemm_launches = sum(prof[k][1] for k in GEMM_GFLOP)
    gemm_tflop_per_launch = sum(GEMM_GFLOP.values()) * B / 1e3 * args.steps / max(gemm_launches, 1)
    gemm_ms_per_launch = gemm_ms / max(gemm_launches, 1)
    achieved = gemm_tflop_per_launch / (gemm_ms_per_launch * 1e-3) if gemm_ms_per_launch else 0.0
    per_family = {k: round(GEMM_GFLOP[k] * B / 1e3 * args.steps / (prof[k][0] * 1e-3), 1)
                  for k in GEMM_GFLOP if prof[k][0] > 0}
    breakdown = {k: round(v[0] / args.steps, 4) for k, v in prof.items() if v[1]}

    # ---- MANO LBS alone at BASELINE config[3] size (outside the timed region) -----------------------
    nh = 65536
    gb = torch.Generator(device=dev).manual_seed(5)
    betas = torch.randn(nh, 10, generator=gb, device=dev)
    r6 = torch.randn(nh * 16, 6, generator=gb, device=dev)
    a1, a2 = r6[:, :3], r6[:, 3:]
    b1 = torch.nn.functional.normalize(a1)
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1)
    rot = torch.stack((b1, b2, torch.linalg.cross(b1, b2)), dim=-1).reshape(nh, 16, 3, 3).contiguous()
    for _ in range(3):
        eng.mano_lbs(betas, rot)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    l0.record()
    for _ in range(10):
        eng.mano_lbs(betas, rot)
    l1.record()
    torch.cuda.synchronize(dev)
    lbs_ms = l0.elapsed_time(l1) / 10
    lbs_gbs = nh * LBS_BYTES_PER_HAND / (lbs_ms * 1e-3) / 1e9

    # ---- the reference's CPU path on this box's host cores (bounded sample) ------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, dt, cores = time_oracle_cpu(sample_crops=8, iters=2, warmup=1)
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': f'2 iterations x 8 crops (of the {B}-crop step), fp32 oracle of the reference PyTorch '
                         f'path, torch CPU with {cores} threads, {dt:.2f} s per iteration'}

    launches = eng.launch_count() + 3 + (1 if world > 1 else 0)       # + input copy, result clone, packing (+ all-gather)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get('per_kernel')                # MB per launch per GEMM family, all from one build (tag in `source`)
        traffic_src = tj.get('source')

    # ---- other configurations of BASELINE.json, reported as extra keys (N = 1 only) -----------------
    extras = {}
    if world == 1 and not args.no_extras:
        try:
            extras['gpu_eager_baseline'] = gpu_eager_baseline(dev, sd, 16)
        except Exception as exc:                   # noqa: BLE001
            extras['gpu_eager_baseline'] = {'error': str(exc)[:200]}
        import bench_video
        v = bench_video.run(model, args.video_frames, 10)
        extras['video'] = {k: v[k] for k in ('p50_ms', 'p99_ms', 'mean_ms', 'frames', 'mean_hands_per_frame', 'hands_per_s')}
        extras['video']['what'] = 'BASELINE configs[4]: 1920x1080 host frames, 2-8 hands per frame, estimate_from_rgb, per-frame sync'
        extras['config0'] = config0_pass(model, sd)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'fp16', 'data': 'synthetic',
        'config': {'workload': f'HaMeR ViT-H/16 (32 blocks, 192 tokens) + MANO decoder head + MANO LBS, synthetic '
                               f'256x192 crops, batch {B} per GPU, random-init weights (BASELINE config[1]'
                               + (f'; {world * B} crops sharded over {world} GPUs + all-gather of MANO outputs' if world > 1 else '') + ')',
                   'batch_per_gpu': B, 'global_batch': world * B,
                   'parallelism': f'dp{world}' if world > 1 else 'single', 'collective': collective,
                   'dtype_detail': 'fp16 tensor-core operands (tcgen05 kind::f16, same rate as bf16), fp32 '
                                   'accumulate / residual / LayerNorm / softmax; decoder head and MANO in fp32',
                   'launch': 'CUDA graph of hm_forward per batch size (Engine.use_graphs); roofline pass un-graphed with per-launch events',
                   'l2': 'per-step working set (1.3 GB fp16 weights + 0.45 GB activations) exceeds the 126 MB L2; no flush needed',
                   'gflop_per_crop': GFLOP_BACKBONE + GFLOP_HEAD_REF},
        'clocks': sampler.result(),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(img_host.numel() * 4),
                'd2h_bytes_per_step': int(host_rec.numel() * 4)},
        'gpu_launches': launches * args.steps,
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['tf_sustained'], 'unit': 'TFLOP/s',
                     'frac': achieved / peaks['tf_sustained'], 'traffic': traffic, 'traffic_unit': 'MB per launch per GEMM family (ncu dram__bytes_read.sum + dram__bytes_write.sum)',
                     'traffic_source': traffic_src,
                     'kernel': 'gemm_bf16_tn_kernel (tcgen05/TMA GEMM, all epilogues; 129 launches per step)',
                     'peak_source': f"{peaks['source']} cuBLAS bf16 sustained (kernel timed inside a long step)",
                     'per_family_tflops': per_family,
                     'end_to_end_tflops': value * (GFLOP_BACKBONE + GFLOP_HEAD_REF) / 1e3,
                     'frac_of_2250_nominal': achieved / 2250.0},
        'stage_ms_per_step': breakdown,
        'profiled_pass_ms_per_step': ms_profiled,
        'mano_lbs': {'hands': nh, 'ms': lbs_ms, 'bound': 'hbm', 'achieved': lbs_gbs, 'unit': 'GB/s',
                     'peak': peaks['hbm_gbs'], 'frac': lbs_gbs / peaks['hbm_gbs'],
                     'bytes_per_hand': LBS_BYTES_PER_HAND},
    }
    if cpu is not None:
        line['cpu_baseline'] = cpu
    line.update(extras)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graphs', action='store_true', help='eager launches (for ncu launch lists)')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: --batch crops per GPU (default); strong: --global-batch crops in total (BASELINE configs[2])')
    ap.add_argument('--global-batch', type=int, default=512)
    ap.add_argument('--no-extras', action='store_true', help='skip the gpu_eager_baseline / video / config0 keys')
    ap.add_argument('--video-frames', type=int, default=200)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
