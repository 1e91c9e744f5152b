"""GPU parity at the configuration bench.py measures (BASELINE configs[1]: 64 crops, max_batch 64) and on
checkpoint-like weights (not bf16/fp16 representable, large activations).

At 64 crops (M = 12288 rows) hm_forward runs a different schedule from the 2-crop tests: CTA-pair 256 x 256 GEMM tiles,
the "222 whole tiles + 18 tail tiles x K-slices" plan of the residual GEMMs, the attention grid with one CTA per
(crop, head), the 25-tile MANO units.  Tolerances are the north-star bars (tests/test_gpu_parity.py::check_forward)."""
import numpy as np
import pytest
import torch

from test_gpu_parity import check_forward, rel_err

pytestmark = pytest.mark.gpu


def _engine(model, max_batch, **kw):
    from hamer_yolo_b200.hamer import Engine
    eng = Engine('cuda:0', max_batch=max_batch, **kw)
    eng.set_mano(model.mano_data)
    for k, v in model.state_dict().items():
        eng.set_weight(k, v)
    eng.finalize_weights()
    return eng


@pytest.fixture(scope='module')
def images64():
    g = torch.Generator().manual_seed(4321)
    return torch.randn(64, 3, 256, 256, generator=g).to(torch.bfloat16).float()


@pytest.fixture(scope='module')
def oracle_out64(oracle_model, images64):
    return oracle_model({'img': images64})[0]


@pytest.fixture(scope='module')
def engine64(oracle_model):
    eng = _engine(oracle_model, 64)
    yield eng
    eng.close()


def test_forward_parity_batch64(engine64, images64, oracle_out64):
    """The benchmarked schedule, eager launches and CUDA-graph replay."""
    dev = engine64.device
    img = images64.to(dev)
    engine64.use_graphs = False
    t = engine64.forward(img)
    res = check_forward(t, oracle_out64)
    engine64.use_graphs = True
    t2 = engine64.forward(img)            # captures the graph
    t3 = engine64.forward(img)            # replays it
    check_forward(t2, oracle_out64)
    check_forward(t3, oracle_out64)
    # the tail-wave K-slices are reduce-added in arrival order: runs agree to fp32 summation order (~0.02 mm), not bitwise
    assert (t3['pred_vertices'] - t['pred_vertices']).abs().max().item() < 1e-4
    assert res['verts_mm'] < 0.1                                          # and with margin: 5x below the bar


def test_backbone_features_batch64(engine64, images64, oracle_out64):
    feats = engine64.backbone(images64.to(engine64.device)).cpu()
    err = (feats - oracle_out64['conditioning_feats']).abs().max().item()
    assert err < 6e-3, err


def test_forward_batch64_deterministic_mode(oracle_model, images64, oracle_out64):
    """hm_config.deterministic at 64 crops: whole tiles only - bit-identical across runs and equal to a 16-crop sub-batch."""
    eng = _engine(oracle_model, 64, deterministic=1)
    try:
        img = images64.to(eng.device)
        a = eng.forward(img)
        b = eng.forward(img)
        c = eng.forward(img[16:32].contiguous())
        check_forward(a, oracle_out64)
        assert torch.equal(a['pred_vertices'], b['pred_vertices'])
        assert torch.equal(a['pred_vertices'][16:32], c['pred_vertices'])
    finally:
        eng.close()


def test_graph_cache_is_dropped_when_weights_or_mano_change(oracle_model, seeded_images):
    """Captured graphs hold weight pointers and tensor maps by value: reloading weights or MANO tables must
    invalidate them, and the replay after the reload must use the new tensors."""
    from hamer_yolo_b200.mano_data import synthetic_mano
    eng = _engine(oracle_model, 2)
    try:
        img = seeded_images.to(eng.device)
        a = eng.forward(img)
        eng.forward(img)
        assert 2 in eng._graphs
        sd = oracle_model.state_dict()
        eng.set_weight('mano_head.init_betas', sd['mano_head.init_betas'] + 0.5)
        assert not eng._graphs
        eng.finalize_weights()
        b = eng.forward(img)
        assert torch.allclose(b['betas'], a['betas'] + 0.5, atol=1e-3)          # (the K-split makes runs differ by ~1e-5)
        eng.forward(img)
        eng.set_mano(synthetic_mano(1))
        assert not eng._graphs
        c = eng.forward(img)
        assert (c['pred_vertices'] - b['pred_vertices']).abs().max().item() > 1e-3      # different hand model
        for k, v in sd.items():                                                        # full reload, same size
            eng.set_weight(k, v)
        eng.finalize_weights()
        eng.set_mano(oracle_model.mano_data)
        d = eng.forward(img)
        assert (d['pred_vertices'] - a['pred_vertices']).abs().max().item() < 1e-4
        # LRU bound of the cache
        eng.MAX_GRAPHS = 2
        eng.forward(img[:1].contiguous())
        eng.forward(img)
        eng.forward(img[:1].contiguous())
        assert len(eng._graphs) <= 2
        # empty batch: empty tensors, no launch
        e = eng.forward(img[:0])
        assert e['pred_vertices'].shape == (0, 778, 3)
    finally:
        eng.close()


# ---------------------------------------------------------------------------------------------
# checkpoint-like weights: arbitrary fp32 values (not representable in 16 bits), un-rounded images
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def raw_oracle():
    """Seed-0 default init of the reference constructors WITHOUT the bf16 rounding of SURVEY H1: what a trained
    checkpoint looks like to the 16-bit tensor-core operands."""
    from oracle.pipeline import OHamer
    from hamer_yolo_b200.mano_data import synthetic_mano
    torch.manual_seed(0)
    return OHamer(synthetic_mano(0))


@pytest.fixture(scope='module')
def raw_images():
    g = torch.Generator().manual_seed(99)
    return torch.randn(4, 3, 256, 256, generator=g)


@pytest.fixture(scope='module')
def raw_out(raw_oracle, raw_images):
    return raw_oracle({'img': raw_images})[0]


def _errors(t, ref):
    p = ref['pred_mano_params']
    rot = torch.cat([t['global_orient'], t['hand_pose']], dim=1).cpu()
    rot_ref = torch.cat([p['global_orient'], p['hand_pose']], dim=1)
    return {'verts_mm': (t['pred_vertices'].cpu() - ref['pred_vertices']).norm(dim=-1).max().item() * 1e3,
            'kp2d_px': (t['pred_keypoints_2d'].cpu() - ref['pred_keypoints_2d']).abs().max().item() * 256,
            'rot_rel': rel_err(rot, rot_ref), 'betas_rel': rel_err(t['betas'].cpu(), p['betas']),
            'cam_rel': rel_err(t['pred_cam'].cpu(), ref['pred_cam'])}


def test_unrounded_fp32_weights_split_mode(raw_oracle, raw_images, raw_out):
    """hm_config.split_weights: W = W_hi + W_lo (two fp16 matrices), every backbone GEMM over K' = 2K.  All north-star
    bars hold for weights that are not representable in 16 bits."""
    eng = _engine(raw_oracle, 4, split_weights=1)
    try:
        t = eng.forward(raw_images.to(eng.device))
        res = check_forward(t, raw_out)
        assert res['rot_rel'] < 6e-4, res          # same margin as with 16-bit representable weights
        # single-CTA plan (M <= 384 rows) and the 2-CTA plan both take the [hi | lo] path
        t1 = eng.forward(raw_images[:1].to(eng.device))
        ref1 = {k: (v[:1] if torch.is_tensor(v) else {kk: vv[:1] for kk, vv in v.items()}) for k, v in raw_out.items()}
        check_forward(t1, ref1)
    finally:
        eng.close()


def test_unrounded_fp32_weights_default_mode(raw_oracle, raw_images, raw_out):
    """Default mode (weights rounded once to fp16) on the same un-rounded weights.  The geometric bars (0.5 mm, 0.5 px)
    hold; the 1e-3 parameter bar is at its edge (CPU emulation: 1.3e-3 on the rotation matrices):
    the figure is printed and the test is an expected failure when it exceeds the bar - split_weights is the mode that
    meets it (test above)."""
    eng = _engine(raw_oracle, 4)
    try:
        t = eng.forward(raw_images.to(eng.device))
    finally:
        eng.close()
    res = _errors(t, raw_out)
    print('un-rounded fp32 weights, default (single fp16) mode:', {k: round(v, 6) for k, v in res.items()})
    assert res['verts_mm'] <= 0.5 and res['kp2d_px'] <= 0.5, res
    assert max(res['rot_rel'], res['betas_rel'], res['cam_rel']) <= 3e-3, res
    if max(res['rot_rel'], res['betas_rel'], res['cam_rel']) > 1e-3:
        pytest.xfail(f'parameter bar 1e-3 missed with single-fp16 weights: {res}')


def test_large_weights_and_activations_stay_finite(raw_oracle, raw_images):
    """Trained ViT-H blocks have outlier fc1 activations.  With one fc1 scaled so that GELU(fc1) exceeds the fp16 range
    (and its weights exceed it too), the saturating conversions keep every output finite; un-saturated, the inf would
    turn the whole residual stream into NaN at the next GEMM."""
    sd = {k: v.clone() for k, v in raw_oracle.state_dict().items()}
    sd['backbone.blocks.5.mlp.fc1.weight'] *= 4e6           # |w| up to ~1e5 > 65504: weight conversion saturates
    sd['backbone.blocks.9.mlp.fc1.weight'] *= 3e3           # activations ~1e5: epilogue conversion saturates
    from hamer_yolo_b200.hamer import Engine
    eng = Engine('cuda:0', max_batch=4)
    try:
        eng.set_mano(raw_oracle.mano_data)
        for k, v in sd.items():
            eng.set_weight(k, v)
        eng.finalize_weights()
        feats = eng.backbone(raw_images.to(eng.device))
        assert torch.isfinite(feats).all()
        t = eng.forward(raw_images.to(eng.device))
        for k, v in t.items():
            assert torch.isfinite(v).all(), k
    finally:
        eng.close()
