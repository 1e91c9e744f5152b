"""Generate the golden fixtures under tests/golden/ by importing the REFERENCE's own Python modules
from /root/reference (authoring container only - the reference does not travel to the GPU box).

    python tests/golden/make_golden.py

What it pins
  1. oracle.backbone.OViT / oracle.head.OManoHead have the same parameter names, shapes and (under
     the same torch seed) the same random-init values as the reference's ``vit()`` and
     ``MANOTransformerDecoderHead`` (asserted here, not stored).
  2. hamer_fwd_seed0.npz : outputs of the reference backbone + head (real modules, fp32, CPU) on a
     seeded batch with bf16-representable >=2-D weights, plus the forward_step outputs
     obtained by feeding the reference head's output through oracle.pipeline (the reference's own
     forward_step needs smplx, absent) on the synthetic MANO.
  3. mano_lbs_real.npz   : outputs of the reference's in-tree manopth ManoLayer.forward on the real
     MANO_RIGHT.pkl for seeded poses (joints + a 64-vertex subsample only; the licensed model data
     itself is never stored).
  4. crop_cv2.npz        : cv2.getAffineTransform / cv2.warpAffine results (hashes + samples) for
     seeded boxes on a seeded synthetic frame.
"""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference'
OUT = os.path.join(ROOT, 'tests', 'golden')


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference_modules():
    """Bare ``hamer.*`` package skeleton + a 3-symbol timm shim."""
    timm = types.ModuleType('timm'); tm = types.ModuleType('timm.models'); tl = types.ModuleType('timm.models.layers')
    tl.drop_path = lambda x, p=0., training=False: x
    tl.to_2tuple = lambda v: v if isinstance(v, tuple) else (v, v)
    tl.trunc_normal_ = torch.nn.init.trunc_normal_
    sys.modules.update({'timm': timm, 'timm.models': tm, 'timm.models.layers': tl})
    for pkg in ['hamer', 'hamer.models', 'hamer.models.backbones', 'hamer.models.components',
                'hamer.models.heads', 'hamer.utils']:
        m = types.ModuleType(pkg); m.__path__ = []
        sys.modules[pkg] = m
    base = f'{REF}/hamer/hamer'
    geometry = _load('hamer.utils.geometry', f'{base}/utils/geometry.py')
    _load('hamer.models.components.t_cond_mlp', f'{base}/models/components/t_cond_mlp.py')
    _load('hamer.models.components.pose_transformer', f'{base}/models/components/pose_transformer.py')
    vit = _load('hamer.models.backbones.vit', f'{base}/models/backbones/vit.py')
    mano_head = _load('hamer.models.heads.mano_head', f'{base}/models/heads/mano_head.py')
    return vit, mano_head, geometry


class AttrDict(dict):
    __getattr__ = dict.__getitem__


def ref_cfg():
    return AttrDict(MODEL=AttrDict(MANO_HEAD=AttrDict(
        TYPE='transformer_decoder', IN_CHANNELS=2048,
        TRANSFORMER_DECODER=AttrDict(depth=6, heads=8, mlp_dim=1024, dim_head=64, dropout=0.0,
                                     emb_dropout=0.0, norm='layer', context_dim=1280))),
        MANO=AttrDict(NUM_HAND_JOINTS=15))


def round_weights_bf16_(module):
    with torch.no_grad():
        for p in module.parameters():
            if p.dim() >= 2:
                p.copy_(p.to(torch.bfloat16).float())


def make_forward_golden():
    from oracle.backbone import OViT
    from oracle.head import OManoHead, default_mean_params
    from oracle.pipeline import OHamer
    from hamer_yolo_b200.mano_data import synthetic_mano
    vit, mano_head, geometry = import_reference_modules()
    mean = default_mean_params()
    orig_load = np.load
    np.load = lambda *a, **k: mean if str(a[0]).endswith('mano_mean_params.npz') else orig_load(*a, **k)
    try:
        torch.manual_seed(0)
        ref_bb = vit.vit(None)
        ref_bb.eval()
        ref_head = mano_head.MANOTransformerDecoderHead(ref_cfg())
        ref_head.eval()
    finally:
        np.load = orig_load
    torch.manual_seed(0)
    model = OHamer(synthetic_mano(0), mean)
    # 1. identical parameter trees
    for ref, mine, tag in ((ref_bb, model.backbone, 'backbone'), (ref_head, model.mano_head, 'head')):
        rs, ms = ref.state_dict(), mine.state_dict()
        assert list(rs.keys()) == list(ms.keys()), (tag, set(rs) ^ set(ms))
        for k in rs:
            assert torch.equal(rs[k], ms[k]), (tag, k)
    print('parameter trees identical:', sum(p.numel() for p in ref_bb.parameters()) / 1e6, 'M +',
          sum(p.numel() for p in ref_head.parameters()) / 1e6, 'M')
    for m in (ref_bb, ref_head, model.backbone, model.mano_head):
        round_weights_bf16_(m)
    g = torch.Generator().manual_seed(1234)
    img = torch.randn(2, 3, 256, 256, generator=g).to(torch.bfloat16).float()
    with torch.no_grad():
        feats = ref_bb(img[:, :, :, 32:-32])
        params, cam, _ = ref_head(feats)
    out, _ = model({'img': img})
    # the oracle must agree with the reference modules it restates
    d_feat = (out['conditioning_feats'] - feats).abs().max().item()
    d_rot = max((out['pred_mano_params'][k] - params[k]).abs().max().item() for k in params)
    d_cam = (out['pred_cam'] - cam).abs().max().item()
    print(f'oracle vs reference modules: feats {d_feat:.3e} mano params {d_rot:.3e} cam {d_cam:.3e}')
    np.savez_compressed(
        os.path.join(OUT, 'hamer_fwd_seed0.npz'),
        seed_weights=0, seed_input=1234,
        feats_sub=feats[:, ::16, ::16].numpy(), feats_absmean=feats.abs().mean().item(),
        global_orient=params['global_orient'].numpy(), hand_pose=params['hand_pose'].numpy(),
        betas=params['betas'].numpy(), pred_cam=cam.numpy(),
        pred_cam_t=out['pred_cam_t'].numpy(), pred_vertices=out['pred_vertices'].numpy(),
        pred_keypoints_3d=out['pred_keypoints_3d'].numpy(),
        pred_keypoints_2d=out['pred_keypoints_2d'].numpy())


def make_mano_golden():
    from hamer_yolo_b200.mano_data import load_mano_pkl
    mp = f'{REF}/rootnet/KeypointFusion/manopth'
    # package skeleton "kf.manopth.*" so that manolayer's "from ..mano.webuser..." resolves to a stub
    # (the real one needs chumpy); every other manopth module is the reference's own file.
    for name, path in (('kf', mp), ('kf.manopth', f'{mp}/manopth'), ('kf.mano', ''), ('kf.mano.webuser', '')):
        pkg = types.ModuleType(name); pkg.__path__ = [path] if path else []
        sys.modules[name] = pkg
    wh = types.ModuleType('kf.mano.webuser.smpl_handpca_wrapper_HAND_only'); wh.ready_arguments = None
    sys.modules[wh.__name__] = wh
    for sub in ('argutils', 'rodrigues_layer', 'tensutils', 'rot6d', 'rotproj'):
        _load(f'kf.manopth.{sub}', f'{mp}/manopth/{sub}.py')
    manolayer = _load('kf.manopth.manolayer', f'{mp}/manopth/manolayer.py')
    rod = sys.modules['kf.manopth.rodrigues_layer']
    m = load_mano_pkl(f'{REF}/rootnet/KeypointFusion/MANO/MANO_RIGHT.pkl')
    layer = manolayer.ManoLayer.__new__(manolayer.ManoLayer)
    torch.nn.Module.__init__(layer)
    layer.center_idx = None; layer.robust_rot = False; layer.rot = 3; layer.flat_hand_mean = True
    layer.side = 'right'; layer.use_pca = False; layer.joint_rot_mode = 'axisang'
    layer.root_rot_mode = 'axisang'; layer.ncomps = 45
    layer.register_buffer('th_betas', torch.zeros(1, 10))
    layer.register_buffer('th_shapedirs', torch.tensor(m['shapedirs']))
    layer.register_buffer('th_posedirs', torch.tensor(m['posedirs'].T.reshape(778, 3, 135).copy()))
    layer.register_buffer('th_v_template', torch.tensor(m['v_template']).unsqueeze(0))
    layer.register_buffer('th_J_regressor', torch.tensor(m['J_regressor']))
    layer.register_buffer('th_weights', torch.tensor(m['lbs_weights']))
    layer.register_buffer('th_hands_mean', torch.zeros(1, 45))
    layer.kintree_parents = [4294967295] + [int(p) for p in m['parents'][1:]]
    g = torch.Generator().manual_seed(7)
    B = 5
    pose_aa = torch.randn(B, 48, generator=g) * 0.6
    betas = torch.randn(B, 10, generator=g)
    with torch.no_grad():
        verts_mm, joints_mm = layer(pose_aa, betas)
        rot = rod.batch_rodrigues(pose_aa.reshape(-1, 3)).reshape(B, 16, 3, 3)
    sub = np.arange(0, 778, 13)[:60].tolist() + [744, 320, 443, 554]
    np.savez_compressed(os.path.join(OUT, 'mano_lbs_real.npz'),
                        betas=betas.numpy(), rotmats=rot.numpy(), vert_ids=np.array(sub),
                        verts_sub=(verts_mm / 1000.0)[:, sub].numpy(),
                        joints21_manopth_tips=(joints_mm / 1000.0).numpy())
    print('manopth golden written; |verts| max', (verts_mm / 1000).abs().max().item())


def synthetic_frame(h=1080, w=1920, seed=0):
    """Seeded noise + gradients, uint8 BGR (SURVEY 8d cfg 5)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (w + h))], axis=-1)
    noise = rng.integers(-40, 40, (h, w, 3))
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def make_crop_golden():
    import cv2
    frame = synthetic_frame(540, 960, 3)
    rng = np.random.default_rng(11)
    boxes, hashes, Ms, samples = [], [], [], []
    for i in range(24):
        x1 = float(rng.integers(-40, 900)); y1 = float(rng.integers(-40, 500))
        w = float(rng.integers(24, 420)); h = float(rng.integers(24, 420))
        if i % 3 == 0:
            x1 += rng.random(); y1 += rng.random(); w += rng.random(); h += rng.random()
        x2, y2 = x1 + w, y1 + h
        cx, cy = (x1 + x2) / 2.0, (y1 + y2) / 2.0
        bw, bh = 2.5 * (x2 - x1), 2.5 * (y2 - y1)
        size = max(bw, bh * 192 / 256) if bh / bw >= 256 / 192 else max(bw, bw * 256 / 192, bh)
        # reference arithmetic for the size (expand_to_aspect_ratio(...).max())
        s = np.array([2.5 * (x2 - x1) / 200.0, 2.5 * (y2 - y1) / 200.0]) * 200
        ww, hh = s
        if hh / ww < 256 / 192:
            size = max(max(ww * 256 / 192, hh), ww)
        else:
            size = max(hh, max(hh * 192 / 256, ww))
        src = np.zeros((3, 2), np.float32)
        c = np.zeros(2); c[0] = cx; c[1] = cy
        src[0, :] = c
        src[1, :] = c + np.array([0, size * 0.5], dtype=np.float32)
        src[2, :] = c + np.array([size * 0.5, 0], dtype=np.float32)
        dst = np.array([[128, 128], [128, 256], [256, 128]], np.float32)
        M = cv2.getAffineTransform(src, dst)
        patch = cv2.warpAffine(frame, M, (256, 256), flags=cv2.INTER_LINEAR,
                               borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        boxes.append([x1, y1, x2, y2]); Ms.append(M)
        hashes.append(hashlib.sha256(patch.tobytes()).hexdigest())
        samples.append(patch[::37, ::41].copy())
    np.savez_compressed(os.path.join(OUT, 'crop_cv2.npz'), frame_seed=3, frame_hw=np.array([540, 960]),
                        boxes=np.array(boxes), M=np.array(Ms), sha256=np.array(hashes),
                        samples=np.array(samples), cv2_version=cv2.__version__)
    print('crop golden written')


if __name__ == '__main__':
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ['crop', 'mano', 'fwd']
    if 'crop' in which:
        make_crop_golden()
    if 'mano' in which:
        make_mano_golden()
    if 'fwd' in which:
        make_forward_golden()
