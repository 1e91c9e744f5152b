"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test calls the CUDA kernels through
the C ABI (libhamer_b200.so) and compares with the CPU oracle / golden fixtures on the same seeded
inputs.  Tolerances (north_star): vertices & 3D joints <= 0.5 mm, 2D keypoints <= 0.5 px,
MANO params <= 1e-3 relative; integer work (crops) bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

MM = 1e-3


def rel_err(a, b):
    """max |a-b| / max |b|  (the stricter of the usual readings of 'relative')."""
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.fixture(scope='session')
def engine(oracle_model):
    from hamer_yolo_b200.hamer import Engine
    eng = Engine('cuda:0', max_batch=8, max_frame_bytes=1080 * 1920 * 3)
    eng.set_mano(oracle_model.mano_data)
    for k, v in oracle_model.state_dict().items():
        eng.set_weight(k, v)
    eng.finalize_weights()
    yield eng
    eng.close()


@pytest.fixture(scope='session')
def oracle_out(oracle_model, seeded_images):
    return oracle_model({'img': seeded_images})


# ---------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------
# The last three shapes are the batch-64 step's own schedules (M = 64 x 192 = 12288 rows, CTA-pair 256 x 256 tiles):
# N = 1280 gives 240 tiles for 74 CTA pairs = 222 whole tiles + 18 tail tiles split into K-slices (gemm_host.cuh),
# N = 3840 / 5120 the plain multi-wave persistent loop.
@pytest.mark.parametrize('cta_group', [2, 1])
@pytest.mark.parametrize('epi,M,N,K', [(0, 192, 3840, 1280), (1, 384, 5120, 1280), (2, 200, 1280, 5120),
                                       (3, 192, 1280, 1536), (2, 1536, 1280, 1280),
                                       (2, 12288, 1280, 1280), (2, 12288, 1280, 5120), (1, 12288, 5120, 1280)])
def test_gemm_epilogues(engine, cta_group, epi, M, N, K):
    dev = engine.device
    g = torch.Generator(device=dev).manual_seed(M + N + K + epi)
    A = (torch.randn(M, K, generator=g, device=dev) * 0.5).half()
    W = (torch.randn(N, K, generator=g, device=dev) * 0.03).half()
    bias = torch.randn(N, generator=g, device=dev)
    pos = torch.randn(192, N, generator=g, device=dev)
    acc = A.float() @ W.float().t() + bias
    if epi in (0, 1):
        out = torch.zeros(M, N, device=dev, dtype=torch.float16)
        ref = torch.nn.functional.gelu(acc) if epi == 1 else acc
    else:
        out = torch.randn(M, N, generator=g, device=dev)
        ref = acc + (out if epi == 2 else pos[torch.arange(M, device=dev) % 192])
    engine.check(engine.lib.hm_test_gemm(engine.h, cta_group, 256, epi, 0, C.c_void_p(A.data_ptr()),
                                         C.c_void_p(W.data_ptr()), C.c_void_p(bias.data_ptr()),
                                         C.c_void_p(pos.data_ptr()), C.c_void_p(out.data_ptr()), M, N, K,
                                         engine.stream()))
    torch.cuda.synchronize()
    if epi in (0, 1):
        assert torch.allclose(out.float(), ref, rtol=2e-3, atol=2e-3)
    else:
        assert torch.allclose(out, ref, rtol=1e-4, atol=2e-4)


def test_fp16_conversions_saturate(engine):
    """Activations above the fp16 range (65,504) must clamp, not become inf: GEMM epilogues (bias, GELU) and the
    LayerNorm output."""
    dev = engine.device
    g = torch.Generator(device=dev).manual_seed(77)
    M, N, K = 256, 256, 1280
    A = (torch.randn(M, K, generator=g, device=dev) * 100).half()
    W = (torch.randn(N, K, generator=g, device=dev) * 10).half()
    bias = torch.randn(N, generator=g, device=dev)
    acc = A.float() @ W.float().t() + bias
    assert (acc.abs() > 65504).float().mean().item() > 0.01          # the case is live
    for epi in (0, 1):
        out = torch.zeros(M, N, device=dev, dtype=torch.float16)
        engine.check(engine.lib.hm_test_gemm(engine.h, 2, 256, epi, 0, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()),
                                             C.c_void_p(bias.data_ptr()), C.c_void_p(bias.data_ptr()),
                                             C.c_void_p(out.data_ptr()), M, N, K, engine.stream()))
        ref = (torch.nn.functional.gelu(acc) if epi == 1 else acc).clamp(-65504, 65504)
        assert torch.isfinite(out).all()
        assert torch.allclose(out.float(), ref, rtol=2e-3, atol=32.0)     # one fp16 ulp at 32768..65504 is 32
    x = torch.randn(64, 1280, generator=g, device=dev)
    gamma = torch.full((1280,), 1e5, device=dev)
    beta = torch.zeros(1280, device=dev)
    y = torch.empty(64, 1280, device=dev, dtype=torch.float16)
    engine.check(engine.lib.hm_test_layernorm(engine.h, C.c_void_p(x.data_ptr()), C.c_void_p(gamma.data_ptr()),
                                              C.c_void_p(beta.data_ptr()), C.c_void_p(y.data_ptr()), 64, 1e-6, 1, engine.stream()))
    ref = torch.nn.functional.layer_norm(x, (1280,), gamma, beta, 1e-6).clamp(-65504, 65504)
    assert torch.isfinite(y).all() and torch.allclose(y.float(), ref, rtol=2e-3, atol=1.0)


@pytest.mark.parametrize('rows,f16', [(192, 1), (1000, 1), (192, 0)])
def test_layernorm(engine, rows, f16):
    dev = engine.device
    g = torch.Generator(device=dev).manual_seed(rows)
    x = torch.randn(rows, 1280, generator=g, device=dev) * 3 + 0.5
    gamma = torch.randn(1280, generator=g, device=dev)
    beta = torch.randn(1280, generator=g, device=dev)
    y = torch.empty(rows, 1280, device=dev, dtype=torch.float16 if f16 else torch.float32)
    engine.check(engine.lib.hm_test_layernorm(engine.h, C.c_void_p(x.data_ptr()), C.c_void_p(gamma.data_ptr()),
                                              C.c_void_p(beta.data_ptr()), C.c_void_p(y.data_ptr()), rows,
                                              1e-6, f16, engine.stream()))
    ref = torch.nn.functional.layer_norm(x.cpu(), (1280,), gamma.cpu(), beta.cpu(), 1e-6)
    if f16:
        assert torch.allclose(y.float().cpu(), ref, rtol=1e-3, atol=2e-3)
    else:
        assert torch.allclose(y.cpu(), ref, rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize('variant', [1, 0])
@pytest.mark.parametrize('B', [1, 3])
def test_attention_core(engine, B, variant):
    dev = engine.device
    g = torch.Generator(device=dev).manual_seed(B)
    qkv = (torch.randn(B * 192, 3840, generator=g, device=dev) * 1.5).half()
    out = torch.empty(B * 192, 1280, device=dev, dtype=torch.float16)
    engine.check(engine.lib.hm_test_attention(engine.h, variant, C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()),
                                              B, engine.stream()))
    q, k, v = qkv.float().cpu().reshape(B, 192, 3, 16, 80).permute(2, 0, 3, 1, 4)
    attn = ((q * 80 ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
    ref = (attn @ v).transpose(1, 2).reshape(B * 192, 1280)
    assert torch.allclose(out.float().cpu(), ref, rtol=2e-3, atol=2e-3)


# ---------------------------------------------------------------------------------------------
# crop / warp / normalise: bit-exact
# ---------------------------------------------------------------------------------------------
def synthetic_frame(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (w + h))], axis=-1)
    return np.clip(base + rng.integers(-40, 40, (h, w, 3)), 0, 255).astype(np.uint8)


def test_crop_bit_exact_vs_oracle_and_cv2(engine):
    import cv2
    from oracle import crop as oc
    frame = synthetic_frame(720, 1280, 4)
    rng = np.random.default_rng(8)
    boxes, labels = [], []
    for i in range(8):
        x1 = float(rng.integers(-60, 1200)); y1 = float(rng.integers(-60, 660))
        w = float(rng.integers(20, 500)); h = float(rng.integers(20, 500))
        if i % 2:
            x1 += rng.random(); y1 += rng.random(); w += rng.random(); h += rng.random()
        boxes.append([x1, y1, x1 + w, y1 + h]); labels.append('right' if i % 3 else 'left')
    img, center, size, trans = engine.preprocess(frame, boxes, [l == 'right' for l in labels])
    ref = oc.prepare_batch_bbox(frame, [[l, b] for l, b in zip(labels, boxes)])
    assert np.array_equal(img.cpu().numpy(), ref['img'])              # every float identical
    assert np.array_equal(center, ref['box_center']) and np.array_equal(size, ref['box_size'])
    assert np.array_equal(trans, ref['trans'])
    # and directly against OpenCV for one box (BGR->RGB, no flip)
    b = boxes[1]
    cx, cy, s = oc.box_to_center_size(*b)
    patch = cv2.warpAffine(frame, oc.affine_from_box(cx, cy, s), (256, 256), flags=cv2.INTER_LINEAR,
                           borderMode=cv2.BORDER_CONSTANT, borderValue=0)[:, :, ::-1]
    assert labels[1] == 'right'
    mean = 255. * np.array([0.485, 0.456, 0.406]); std = 255. * np.array([0.229, 0.224, 0.225])
    t = np.transpose(patch.copy(), (2, 0, 1)).astype(np.float32)
    for c in range(3):
        t[c] = (t[c] - float(mean[c])) / float(std[c])
    assert np.array_equal(img[1].cpu().numpy(), t)


def test_crop_edge_cases(engine):
    from oracle import crop as oc
    frame = synthetic_frame(90, 130, 6)
    cases = [['left', [-400., -300., -350., -200.]],      # fully outside -> constant border
             ['right', [0., 0., 130., 90.]],               # whole frame
             ['right', [60., 40., 62., 41.]],              # tiny box (heavy magnification)
             ['left', [-20., -20., 400., 300.]]]           # larger than the frame
    img, *_ = engine.preprocess(frame, [c[1] for c in cases], [c[0] == 'right' for c in cases])
    ref = oc.prepare_batch_bbox(frame, cases)
    assert np.array_equal(img.cpu().numpy(), ref['img'])
    # device-resident frame path gives the same bytes
    img2, *_ = engine.preprocess(torch.from_numpy(frame).to(engine.device), [c[1] for c in cases],
                                 [c[0] == 'right' for c in cases])
    assert torch.equal(img, img2)
    from hamer_yolo_b200._lib import HmError
    with pytest.raises(HmError):
        engine.preprocess(frame, [[10., 10., 10., 50.]], [True])      # empty box


def test_crop_rectangle_upload_branch(engine):
    """A 1080p host frame with a few small, scattered boxes: hm_preprocess uploads one rectangle per crop instead of the
    row band (the rectangles + 64 KB each are much smaller than the band).  The device frame buffer is first filled
    with a different frame, so a pixel read outside the uploaded rectangles would show."""
    from oracle import crop as oc
    other = np.full((1080, 1920, 3), 255, np.uint8)
    engine.preprocess(other, [[0., 0., 1920., 1080.]], [True])            # band path: the whole poison frame is resident
    frame = synthetic_frame(1080, 1920, 21)
    cases = [['right', [100., 60., 180., 150.]], ['left', [1700., 900., 1790., 1000.]], ['right', [905.5, 500.25, 990., 590.75]],
             ['left', [1880., 20., 1960., 100.]]]                          # the last one hangs over the right edge
    boxes = [c[1] for c in cases]
    # the branch condition of hm_api.cu (hm_preprocess), restated: rectangles (+4 px margin) + 64 KB each < row band
    rect = 0
    lo, hi = 1e9, -1e9
    for b in boxes:
        cx, cy, s = oc.box_to_center_size(*b)
        x0, x1 = max(0, np.floor(cx - s / 2) - 4), min(1919, np.ceil(cx + s / 2) + 4)
        y0, y1 = max(0, np.floor(cy - s / 2) - 4), min(1079, np.ceil(cy + s / 2) + 4)
        rect += (x1 - x0 + 1) * 3 * (y1 - y0 + 1)
        lo, hi = min(lo, cy - s / 2), max(hi, cy + s / 2)
    band = (min(1079, np.ceil(hi) + 4) - max(0, np.floor(lo) - 4)) * 1920 * 3
    assert rect + len(boxes) * 65536 < band
    img, center, size, trans = engine.preprocess(frame, boxes, [c[0] == 'right' for c in cases])
    ref = oc.prepare_batch_bbox(frame, cases)
    assert np.array_equal(img.cpu().numpy(), ref['img'])
    assert np.array_equal(center, ref['box_center']) and np.array_equal(size, ref['box_size'])
    # same frame through the band path (one box spanning the frame forces it) gives the same crops
    img2, *_ = engine.preprocess(frame, boxes + [[0., 0., 1920., 1080.]], [c[0] == 'right' for c in cases] + [True])
    assert torch.equal(img2[:4], img)


# ---------------------------------------------------------------------------------------------
# MANO layer
# ---------------------------------------------------------------------------------------------
def random_mano_inputs(B, seed):
    from oracle.head import rot6d_to_rotmat
    g = torch.Generator().manual_seed(seed)
    betas = torch.randn(B, 10, generator=g)
    rot = rot6d_to_rotmat(torch.randn(B * 16, 6, generator=g)).reshape(B, 16, 3, 3).contiguous()
    cam_t = torch.stack([torch.randn(B, generator=g) * 0.1, torch.randn(B, generator=g) * 0.1,
                         20 + torch.rand(B, generator=g) * 30], dim=1)
    return betas, rot, cam_t


@pytest.mark.parametrize('B', [1, 16, 37, 300])
def test_mano_lbs_vs_oracle(engine, oracle_model, B):
    from oracle.mano import mano_forward
    from oracle.pipeline import perspective_projection
    betas, rot, cam_t = random_mano_inputs(B, 100 + B)
    dev = engine.device
    verts, joints, kp2d = engine.mano_lbs(betas.to(dev), rot.to(dev), cam_t.to(dev))
    v_ref, j_ref = mano_forward(rot[:, :1], rot[:, 1:], betas, oracle_model.mano_data)
    k_ref = perspective_projection(j_ref, cam_t, torch.full((B, 2), 5000.0 / 256))
    assert (verts.cpu() - v_ref).norm(dim=-1).max().item() < 0.03 * MM         # fp16 blend coefficients + fp16 pose basis
    assert (joints.cpu() - j_ref).norm(dim=-1).max().item() < 0.02 * MM
    assert (kp2d.cpu() - k_ref).abs().max().item() * 256 < 0.01                # px


def test_mano_lbs_three_kernel_path_agrees(engine, oracle_model):
    """hm_config.mano_legacy selects the prep / GEMM / skinning kernels (pose basis split hi + lo, exact to 2^-22);
    the fused kernel rounds the pose basis to fp16: the two agree to 0.03 mm (0.004 mm on the real MANO asset)."""
    from hamer_yolo_b200.hamer import Engine
    betas, rot, cam_t = random_mano_inputs(700, 3)
    dev = engine.device
    v, j, k = engine.mano_lbs(betas.to(dev), rot.to(dev), cam_t.to(dev))
    eng = Engine('cuda:0', max_batch=2, mano_legacy=1)
    try:
        eng.set_mano(oracle_model.mano_data)
        v2, j2, k2 = eng.mano_lbs(betas.to(dev), rot.to(dev), cam_t.to(dev))
    finally:
        eng.close()
    assert (v - v2).abs().max().item() < 3e-5 and (j - j2).abs().max().item() < 3e-5
    assert (k - k2).abs().max().item() * 256 < 1e-2            # px


@pytest.mark.parametrize('B', [1, 37, 129, 300])
def test_mano_lbs_writes_stay_inside_the_outputs(engine, B):
    """Guard bands around every output of hm_mano_lbs (ragged hand tiles, the partial last vertex group, fingertip and
    projection writes): nothing outside [B,778,3] / [B,21,3] / [B,21,2] may change, everything inside must."""
    import ctypes as C
    betas, rot, cam_t = random_mano_inputs(B, 50 + B)
    dev = engine.device
    pad = 4096
    bufs = {}
    for name, numel in (('verts', B * 778 * 3), ('joints', B * 63), ('kp2d', B * 42)):
        bufs[name] = torch.full((pad + numel + pad,), float('nan'), device=dev)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    b_d, r_d, c_d = betas.to(dev), rot.to(dev), cam_t.to(dev)
    with torch.cuda.device(dev):
        engine.check(engine.lib.hm_mano_lbs(engine.h, ptr(b_d), ptr(r_d), ptr(c_d), B, ptr(bufs['verts'][pad:]),
                                            ptr(bufs['joints'][pad:]), ptr(bufs['kp2d'][pad:]), engine.stream()))
    torch.cuda.synchronize()
    for name, t in bufs.items():
        assert torch.isnan(t[:pad]).all() and torch.isnan(t[-pad:]).all(), name
        assert torch.isfinite(t[pad:-pad]).all(), name
    v_ref, j_ref, k_ref = engine.mano_lbs(b_d, r_d, c_d)
    assert torch.equal(bufs['verts'][pad:-pad].view(B, 778, 3), v_ref)
    assert torch.equal(bufs['joints'][pad:-pad].view(B, 21, 3), j_ref) and torch.equal(bufs['kp2d'][pad:-pad].view(B, 21, 2), k_ref)


def test_mano_lbs_full_size_properties(engine):
    """BASELINE config 4 size (65536 hands): batch-position independence (bit-exact) and
    global-rotation equivariance about the root joint."""
    B = 65536
    betas, rot, _ = random_mano_inputs(B, 7)
    dev = engine.device
    verts, joints = engine.mano_lbs(betas.to(dev), rot.to(dev))
    idx = torch.tensor([0, 1, 15, 16, 17, 4095, 40000, 65535])
    v_s, j_s = engine.mano_lbs(betas[idx].to(dev), rot[idx].to(dev))
    assert torch.equal(verts[idx.to(dev)], v_s) and torch.equal(joints[idx.to(dev)], j_s)
    assert torch.isfinite(verts).all()
    # a random 256-hand subset of the full-size run against the oracle
    from oracle.mano import mano_forward
    from hamer_yolo_b200.mano_data import synthetic_mano
    sub = torch.randperm(B, generator=torch.Generator().manual_seed(3))[:256]
    v_ref, j_ref = mano_forward(rot[sub, :1], rot[sub, 1:], betas[sub], synthetic_mano(0))
    assert (verts[sub.to(dev)].cpu() - v_ref).norm(dim=-1).max().item() < 0.03 * MM
    assert (joints[sub.to(dev)].cpu() - j_ref).norm(dim=-1).max().item() < 0.02 * MM
    from oracle.head import rot6d_to_rotmat
    Q = rot6d_to_rotmat(torch.randn(B, 6, generator=torch.Generator().manual_seed(9)))
    rot2 = rot.clone()
    rot2[:, 0] = Q @ rot[:, 0]
    v2, j2 = engine.mano_lbs(betas.to(dev), rot2.to(dev))
    root = joints[:, :1]
    Qd = Q.to(dev)
    assert ((v2 - root) - (verts - root) @ Qd.transpose(1, 2)).abs().max().item() < 2e-6
    assert (j2[:, :1] - root).abs().max().item() < 1e-6


# ---------------------------------------------------------------------------------------------
# backbone / head / full forward
# ---------------------------------------------------------------------------------------------
def test_backbone_features(engine, seeded_images, oracle_out):
    feats = engine.backbone(seeded_images.to(engine.device)).cpu()
    ref = oracle_out[0]['conditioning_feats']
    err = (feats - ref).abs().max().item()
    assert err < 6e-3, err          # fp16-activation rounding budget (emulation: 1.7e-3)
    g = np.load(os.path.join(GOLDEN, 'hamer_fwd_seed0.npz'))
    assert np.abs(feats[:, ::16, ::16].numpy() - g['feats_sub']).max() < 6e-3


def test_decoder_head_fp32(engine, oracle_model, oracle_out):
    """The head alone, fed the oracle's features: fp32 re-association must be ~1e-6 exact."""
    ref, _ = oracle_out
    t = engine.mano_head(ref['conditioning_feats'].to(engine.device))
    p = ref['pred_mano_params']
    assert (t['global_orient'].cpu() - p['global_orient']).abs().max().item() < 2e-5
    assert (t['hand_pose'].cpu() - p['hand_pose']).abs().max().item() < 2e-5
    assert (t['betas'].cpu() - p['betas']).abs().max().item() < 2e-5
    assert (t['pred_cam'].cpu() - ref['pred_cam']).abs().max().item() < 2e-5
    assert torch.allclose(t['pred_cam_t'].cpu(), ref['pred_cam_t'], rtol=1e-4, atol=1e-4)


def check_forward(t, ref):
    p = ref['pred_mano_params']
    rot = torch.cat([t['global_orient'], t['hand_pose']], dim=1).cpu()
    rot_ref = torch.cat([p['global_orient'], p['hand_pose']], dim=1)
    res = {
        'verts_mm': (t['pred_vertices'].cpu() - ref['pred_vertices']).norm(dim=-1).max().item() / MM,
        'joints_mm': (t['pred_keypoints_3d'].cpu() - ref['pred_keypoints_3d']).norm(dim=-1).max().item() / MM,
        'kp2d_px': (t['pred_keypoints_2d'].cpu() - ref['pred_keypoints_2d']).abs().max().item() * 256,
        'rot_rel': rel_err(rot, rot_ref), 'betas_rel': rel_err(t['betas'].cpu(), p['betas']),
        'cam_rel': rel_err(t['pred_cam'].cpu(), ref['pred_cam']),
    }
    print('parity:', {k: round(v, 6) for k, v in res.items()})
    assert res['verts_mm'] <= 0.5 and res['joints_mm'] <= 0.5, res
    assert res['kp2d_px'] <= 0.5, res
    assert res['rot_rel'] <= 1e-3 and res['betas_rel'] <= 1e-3 and res['cam_rel'] <= 1e-3, res
    return res


def test_full_forward_parity(engine, seeded_images, oracle_out):
    t = engine.forward(seeded_images.to(engine.device))
    check_forward(t, oracle_out[0])
    # and against the stored golden (reference modules' own outputs)
    g = np.load(os.path.join(GOLDEN, 'hamer_fwd_seed0.npz'))
    assert np.abs(t['pred_vertices'].cpu().numpy() - g['pred_vertices']).max() < 0.5 * MM
    assert rel_err(t['betas'].cpu(), torch.tensor(g['betas'])) <= 1e-3
    rot = torch.cat([t['global_orient'], t['hand_pose']], dim=1).cpu()
    assert rel_err(rot, torch.tensor(np.concatenate([g['global_orient'], g['hand_pose']], axis=1))) <= 1e-3


def test_batch_one_and_single_cta_tiles(oracle_model, seeded_images, oracle_out):
    """B = 1 (192 rows, ragged tiles) and the single-CTA GEMM variant give the same parity."""
    from hamer_yolo_b200.hamer import Engine
    eng = Engine('cuda:0', max_batch=2, gemm_cta_group=1)
    try:
        eng.set_mano(oracle_model.mano_data)
        for k, v in oracle_model.state_dict().items():
            eng.set_weight(k, v)
        eng.finalize_weights()
        t = eng.forward(seeded_images[:1].to(eng.device))
        ref = {k: (v[:1] if torch.is_tensor(v) else {kk: vv[:1] for kk, vv in v.items()})
               for k, v in oracle_out[0].items()}
        check_forward(t, ref)
    finally:
        eng.close()
