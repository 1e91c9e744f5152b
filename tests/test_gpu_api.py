"""GPU tests of the reference-facing Python API (HAMER, hamer_inference, npy writer, sharding)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def synthetic_frame(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (w + h))], axis=-1)
    return np.clip(base + rng.integers(-40, 40, (h, w, 3)), 0, 255).astype(np.uint8)


@pytest.fixture(scope='module')
def model(oracle_model):
    from hamer_yolo_b200 import load_hamer
    m, cfg = load_hamer(state_dict=oracle_model.state_dict(), mano=oracle_model.mano_data,
                        device='cuda:0', max_batch=4)
    assert cfg.MODEL.BBOX_SHAPE == [192, 256]
    return m


def test_forward_dict_layout(model, seeded_images, oracle_model):
    out, params = model({'img': seeded_images.cuda(), 'box_center': torch.zeros(2, 2)})   # extra keys ignored
    ref, ref_params = oracle_model({'img': seeded_images})
    assert set(out) == set(ref) - {'conditioning_feats'}
    assert set(params) == set(ref_params)
    for k in ('pred_cam', 'pred_cam_t', 'focal_length', 'pred_keypoints_3d', 'pred_vertices', 'pred_keypoints_2d'):
        assert out[k].shape == ref[k].shape and out[k].dtype == torch.float32 and out[k].is_cuda
    for k in ('global_orient', 'hand_pose', 'betas'):
        assert out['pred_mano_params'][k].shape == ref['pred_mano_params'][k].shape
        assert out['pred_mano_params'][k].data_ptr() != params[k].data_ptr()        # clones (hamer.py:126)
    assert torch.equal(params['trans'], out['pred_cam_t'])
    assert model.mano.faces.shape == (1538, 3)
    # batches larger than max_batch are chunked transparently.  The K-slices of the residual GEMMs are reduce-added in
    # arrival order and the split depends on the batch size: a different fp32 summation order flips fp16 roundings of
    # later activations, so a crop's vertices move by ~0.02 mm between batchings (the distance to the fp32 oracle is
    # the same 0.02 mm; bar 0.5 mm).  Bit-exactness: test_deterministic_mode_is_batch_invariant.
    big = torch.cat([seeded_images, seeded_images, seeded_images]).cuda()
    out6, _ = model({'img': big})
    assert (out6['pred_vertices'][:2] - out['pred_vertices']).abs().max().item() < 1e-4
    assert (out6['pred_vertices'][4:] - out['pred_vertices']).abs().max().item() < 1e-4
    with pytest.raises(ValueError):
        model({'img': torch.zeros(1, 3, 224, 224).cuda()})


@pytest.mark.parametrize('use_k', [False, True])
def test_estimate_from_rgb_vs_oracle(model, oracle_model, use_k):
    from hamer_yolo_b200 import hamer_inference
    from oracle import crop as oc
    from oracle.pipeline import postprocess
    frame = synthetic_frame(480, 640, 12)
    dets = [['right', [100, 80, 260, 300]], ['left', [330, 120, 520, 420]]]
    k_real = np.array([[600., 0, 320.], [0, 610., 240.], [0, 0, 1]], np.float32) if use_k else None
    hi = hamer_inference(model=model)
    out, params = hi.estimate_from_rgb(frame, dets, k_real)
    batch = {k: torch.from_numpy(v) for k, v in oc.prepare_batch_bbox(frame, dets).items()}
    ref, _ = oracle_model(batch)
    ref = postprocess(ref, batch, k_real)
    assert torch.equal(out['img'].cpu(), batch['img'])                       # crops bit-exact
    assert torch.equal(out['trans'].cpu(), batch['trans']) and torch.equal(out['do_flip'].cpu(), batch['do_flip'])
    assert (out['pred_vertices'].cpu() - ref['pred_vertices']).norm(dim=-1).max().item() < 0.5e-3
    assert (out['pred_keypoints_3d'].cpu() - ref['pred_keypoints_3d']).norm(dim=-1).max().item() < 0.5e-3
    rel = ((out['pred_cam_t_full'].cpu() - ref['pred_cam_t_full']).abs().max() / ref['pred_cam_t_full'].abs().max()).item()
    assert rel < 1e-3, rel
    assert (out['pred_keypoints_2d_full'].cpu() - ref['pred_keypoints_2d_full']).abs().max().item() < 0.5   # px
    # the reference's quirk: x of every right hand (do_flip = 0) is zeroed in place (infer.py:392)
    assert torch.all(out['pred_keypoints_3d'][0, :, 0] == 0)
    with pytest.raises(ValueError):
        hi.estimate_from_rgb(frame, [])


def test_process_image_npy(model, tmp_path):
    from hamer_yolo_b200 import hamer_inference
    from hamer_yolo_b200.inference import process_image, save_image_results
    hi = hamer_inference(model=model)
    frame = synthetic_frame(480, 640, 13)
    res = process_image(hi, frame, [['right', [100, 80, 260, 300]], ['left', [330, 120, 520, 420]]])
    path = save_image_results(str(tmp_path), 'img0', res)
    back = np.load(path, allow_pickle=True).item()
    assert back['right']['is_right'] is True and back['left']['is_right'] is False
    assert back['left']['theta'].shape == (48,) and back['left']['cam_t'].shape == (3,)
    assert np.isfinite(back['right']['theta']).all()


def test_sharded_forward_matches_unsharded(seeded_images, oracle_model):
    """Single-process stand-in for the N-rank path: shards processed one after another must equal
    the unsharded batch bit for bit (no cross-crop coupling anywhere in the path; deterministic mode so
    that the K-split of the residual GEMMs does not depend on the shard size)."""
    from hamer_yolo_b200 import dist as hd, load_hamer
    model, _ = load_hamer(state_dict=oracle_model.state_dict(), mano=oracle_model.mano_data, device='cuda:0',
                          max_batch=8, deterministic=True)
    img4 = torch.cat([seeded_images, seeded_images.flip(0)]).cuda()        # 4 crops
    # 4 = 2 + 2 and 7 = 4 + 3: odd and even batches must run the same kernels (the packed output buffer of a graph once
    # put pred_vertices 4 bytes off for odd batches, which sent hm_mano_lbs down its three-kernel path: equal within the
    # tolerances, not bit for bit - found by the 2-GPU test)
    for img in (img4, torch.cat([img4, img4.flip(-1)[:3]])):
        n = img.shape[0]
        full, _ = model({'img': img})
        full = {k: (v.clone() if torch.is_tensor(v) else {kk: vv.clone() for kk, vv in v.items()}) for k, v in full.items()}
        recs = []
        for r in range(2):
            lo, hi = hd.shard_range(n, r, 2)
            o, _ = model({'img': img[lo:hi]})
            recs.append(hd.pack_outputs(o))
        back = hd.unpack_outputs(torch.cat(recs))
        assert torch.equal(back['pred_vertices'], full['pred_vertices']), n
        assert torch.equal(back['pred_keypoints_2d'], full['pred_keypoints_2d']), n
        assert torch.equal(back['pred_mano_params']['hand_pose'], full['pred_mano_params']['hand_pose']), n
        assert torch.equal(back['pred_mano_params']['betas'], full['pred_mano_params']['betas']), n


def test_deterministic_mode_is_batch_invariant(seeded_images, oracle_model):
    """hm_config.deterministic: no K-split, so a crop's result does not depend on the batch it is in or on the run."""
    from hamer_yolo_b200 import load_hamer
    m, _ = load_hamer(state_dict=oracle_model.state_dict(), mano=oracle_model.mano_data, device='cuda:0', max_batch=4,
                      deterministic=True)
    a, _ = m({'img': seeded_images.cuda()})
    b, _ = m({'img': torch.cat([seeded_images, seeded_images, seeded_images]).cuda()})
    c, _ = m({'img': seeded_images.cuda()})
    assert torch.equal(a['pred_vertices'], c['pred_vertices'])
    assert torch.equal(b['pred_vertices'][:2], a['pred_vertices']) and torch.equal(b['pred_vertices'][4:], a['pred_vertices'])


@pytest.mark.parametrize('use_k', [False, True])
def test_camera_postprocess_kernel_matches_torch_bitwise(model, use_k):
    """hm_camera_postprocess against the reference's PyTorch expressions (renderer.py:24-74, infer.py:392-486):
    separately rounded fp32 operations in the same order -> identical bits."""
    from hamer_yolo_b200.inference import custom_cam_crop_to_full, perspective_projection, META_COLS
    g = torch.Generator().manual_seed(11)
    n = 37
    cam = torch.stack([0.6 + torch.rand(n, generator=g), torch.randn(n, generator=g) * 0.2, torch.randn(n, generator=g) * 0.2], 1)
    kp3d = torch.randn(n, 21, 3, generator=g) * 0.08
    meta = torch.zeros(n, META_COLS)
    meta[:, 0] = torch.rand(n, generator=g) * 1920
    meta[:, 1] = torch.rand(n, generator=g) * 1080
    meta[:, 2] = 80 + torch.rand(n, generator=g) * 400
    meta[:, 3], meta[:, 4] = 1920, 1080
    meta[:, 5] = (torch.rand(n, generator=g) < 0.5).float()
    if use_k:
        meta[:, 12], meta[:, 13], meta[:, 14], meta[:, 15] = 1400.0, 1390.0, 955.0, 541.0
    else:
        meta[:, 12] = meta[:, 13] = 5000.0 / 256 * 1920
        meta[:, 14], meta[:, 15] = 960.0, 540.0
    dev = model.device
    cam_d, kp_d, meta_d = cam.to(dev), kp3d.to(dev), meta.to(dev)
    # reference expressions on the same device
    flip = meta_d[:, 5]
    kp_ref = kp_d.clone()
    kp_ref[:, :, 0] = kp_ref[:, :, 0] * flip.unsqueeze(1)
    c = cam_d.clone()
    c[:, 1] = c[:, 1] * (1.0 - 2.0 * flip)
    t_ref = custom_cam_crop_to_full(c, meta_d[:, 0:2], meta_d[:, 2], meta_d[:, 3:5], meta_d[:, 12], meta_d[:, 13],
                                    meta_d[:, 14], meta_d[:, 15])
    if use_k:
        p = kp_ref + t_ref.unsqueeze(1)
        depth = p[:, :, 2:3] + 1e-9
        uv_ref = torch.cat([p[:, :, 0:1] / depth * meta_d[:, 12].reshape(-1, 1, 1) + meta_d[:, 14].reshape(-1, 1, 1),
                            p[:, :, 1:2] / depth * meta_d[:, 13].reshape(-1, 1, 1) + meta_d[:, 15].reshape(-1, 1, 1)], dim=-1)
    else:
        uv_ref = perspective_projection(kp_ref, t_ref, torch.stack([meta_d[:, 12], meta_d[:, 13]], dim=1))
    differ = not torch.allclose(meta_d[:, 12], meta_d[:, 13])
    t, uv = model.engine.camera_postprocess(cam_d, kp_d, meta_d, differ, use_k)
    assert torch.equal(t, t_ref) and torch.equal(uv, uv_ref) and torch.equal(kp_d, kp_ref)


def test_programmatic_dependent_launch_changes_no_bits(tmp_path):
    """Programmatic dependent launch (the launches carry the programmatic-stream-serialization attribute; every kernel waits
    on griddepcontrol.wait) is automatic for small batches; HM_PDL=0 / 1 force it off / on: identical results, eager and
    graph-replayed."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / 'pdl_run.py'
    script.write_text(
        "import sys, torch\n"
        f"sys.path.insert(0, {root!r})\n"
        "from hamer_yolo_b200 import load_hamer, synthetic_mano\n"
        "from hamer_yolo_b200.weights import random_state_dict\n"
        "dev = torch.device('cuda:0')\n"
        "m, _ = load_hamer(state_dict=random_state_dict(seed=3, device=dev), mano=synthetic_mano(0), device='cuda:0',\n"
        "                  max_batch=4, deterministic=True)\n"
        "img = torch.randn(3, 3, 256, 256, generator=torch.Generator().manual_seed(5)).to(dev)\n"
        "a, _ = m({'img': img})\n"
        "b, _ = m({'img': img})\n"                      # second call replays the captured graph
        "assert torch.equal(a['pred_vertices'], b['pred_vertices'])\n"
        "torch.save(a['pred_vertices'].cpu(), sys.argv[1])\n")
    outs = []
    for pdl in ('0', '1'):
        out = tmp_path / f'verts_{pdl}.pt'
        env = dict(os.environ, HM_PDL=pdl)
        subprocess.run([sys.executable, str(script), str(out)], check=True, env=env, timeout=240)
        outs.append(torch.load(out))
    assert torch.equal(outs[0], outs[1])


def test_vitdet_dataset_and_hamer_infer(model, oracle_model):
    """ViTDetDataset mirror (vitdet_dataset.py:16-95: float32 box arithmetic, flip inside the warp) is bit-exact against
    the oracle, and HAMER_INFER returns the single dict of hamer.py:536-553."""
    from hamer_yolo_b200 import ViTDetDataset, HAMER_INFER
    from hamer_yolo_b200.configs import default_config
    from oracle import crop as oc
    frame = synthetic_frame(480, 640, 31)
    boxes = np.array([[100, 80, 260, 280], [330.5, 120.25, 490, 330], [-20, 30, 120, 200]], np.float32)   # crop windows < 563 px
    right = np.array([1, 0, 0])
    ds = ViTDetDataset(default_config(), frame, boxes, right, engine=model.engine)
    assert len(ds) == 3
    b = ds.batch()
    for i in range(3):
        t, center, size = oc.vitdet_item(frame, boxes[i], right[i])
        item = ds[i]
        assert np.array_equal(item['img'], t) and np.array_equal(b['img'][i].cpu().numpy(), t)
        assert np.array_equal(item['box_center'], center) and item['box_size'] == size and item['right'] == right[i]
        assert item['personid'] == i and np.array_equal(item['img_size'], np.array([640., 480.]))
    # crop windows > 563 px: the frame is blurred (skimage gaussian = scipy gaussian_filter) and the float64 image is warped
    # (vitdet_dataset.py:62-71) - bit-exact as well, both hands, frame borders inside the blur's support
    big = synthetic_frame(1080, 1920, 32)
    bboxes = np.array([[500, 200, 900, 640], [1500.5, 600.25, 1915, 1075], [0, 0, 700, 900], [300, 300, 450, 500]], np.float32)
    bright = np.array([1, 0, 0, 1])
    dsb = ViTDetDataset(default_config(), big, bboxes, bright, engine=model.engine)
    bb = dsb.batch()
    for i in range(4):
        t, center, size = oc.vitdet_item(big, bboxes[i], bright[i])
        assert (size / 256 / 2 > 1.1) == (i < 3)                     # the last box takes the 8-bit path inside the same call
        assert np.array_equal(bb['img'][i].cpu().numpy(), t), i
        assert dsb[i]['box_size'] == size
    infer = HAMER_INFER.__new__(HAMER_INFER)
    infer.__dict__.update(model.__dict__)                    # same engine / weights, the inference-only surface
    out = infer({'img': b['img'][:2]})
    ref, ref_params = model({'img': b['img'][:2]})
    assert isinstance(out, dict) and set(out) == set(ref) | {'pose', 'beta'}
    assert out['pose'].shape == (2, 16, 3, 3) and out['beta'].shape == (2, 10)
    assert torch.allclose(out['pose'][:, :1], out['pred_mano_params']['global_orient'])


def test_allgather_entry_point_single_rank(model):
    """hm_comm_unique_id / hm_comm_init / hm_allgather through the C ABI with a one-rank communicator (what a 1-GPU box
    can run): the gather is the identity.  The N-rank case runs in tests/test_gpu_multi.py on boxes with >= 2 GPUs."""
    import ctypes as C
    from hamer_yolo_b200 import dist as hd
    eng = model.engine
    buf = (C.c_uint8 * 128)()
    assert eng.lib.hm_comm_unique_id(buf) == 0
    eng.check(eng.lib.hm_comm_init(eng.h, buf, 1, 0))
    eng.comm_ranks = 1
    try:
        rec = torch.randn(5, hd.RECORD_FLOATS, device=eng.device)
        out = hd.all_gather_records(rec, [5], engine=eng)
        torch.cuda.synchronize()
        assert torch.equal(out, rec)
    finally:
        eng.check(eng.lib.hm_comm_destroy(eng.h))
        eng.comm_ranks = 0
    from hamer_yolo_b200._lib import HmError
    with pytest.raises(HmError):
        eng.check(eng.lib.hm_allgather(eng.h, None, C.c_void_p(rec.data_ptr()), C.c_void_p(out.data_ptr()), 5, eng.stream()))


def test_device_log_map_matches_cv2_rodrigues(model):
    """hm_rotmat_to_axis_angle against cv2.Rodrigues (what matrix_to_axis_angle calls per 3 x 3, infer.py:1082-1096),
    including rotations near 0 and near pi."""
    import cv2
    from oracle.head import rot6d_to_rotmat
    g = torch.Generator().manual_seed(5)
    R = rot6d_to_rotmat(torch.randn(500, 6, generator=g))
    axis = torch.nn.functional.normalize(torch.randn(6, 3, generator=g), dim=1)
    special = []
    for k, ang in enumerate([0.0, 1e-7, 1e-4, np.pi - 1e-7, np.pi - 1e-4, 3.0]):
        special.append(torch.from_numpy(cv2.Rodrigues((axis[k] * ang).numpy().astype(np.float64))[0]).float())
    R = torch.cat([R, torch.stack(special)])
    aa = model.engine.rotmat_to_axis_angle(R.cuda()).cpu().numpy()
    ref = np.stack([cv2.Rodrigues(m.numpy())[0].reshape(3) for m in R])
    # compare as rotations (near pi the axis sign is a convention) and as vectors away from pi
    back = np.stack([cv2.Rodrigues(v.astype(np.float64))[0] for v in aa])
    assert np.abs(back - R.numpy()).max() < 2e-6
    ang = np.linalg.norm(ref, axis=1)
    ok = ang < np.pi - 1e-3
    assert np.abs(aa[ok] - ref[ok]).max() < 5e-6


def test_npz_and_obj_writers(model, tmp_path):
    """process_batch's per-hand npz (infer.py:1003-1025), the device log-map npy records and the OBJ reconstruction from
    the npy (infer.py:1321-1436): the OBJ vertices must equal MANO(axis-angle of the record) + cam_t, left hands mirrored."""
    from hamer_yolo_b200 import hamer_inference
    from hamer_yolo_b200 import inference as hi_mod
    from oracle.mano import mano_forward
    hi = hamer_inference(model=model)
    frame = synthetic_frame(480, 640, 14)
    dets = [['right', [100, 80, 260, 300]], ['left', [330, 120, 520, 420]], ['right', [20, 200, 150, 380]]]
    paths = hi_mod.process_image_npz(hi, frame, dets, str(tmp_path / 'npz'), 'img0')
    assert [p.split('/')[-1] for p in paths] == ['img0_right.npz', 'img0_left.npz', 'img0_right_2.npz']
    z = np.load(paths[1])
    assert z['betas'].shape == (1, 10) and z['global_orient'].shape == (1, 1, 3, 3) and z['hand_pose'].shape == (1, 15, 3, 3)
    assert z['cam_t'].shape == (1, 3) and z['is_right'].shape == (1,) and not z['is_right'][0] and np.load(paths[0])['is_right'][0]
    # npy records: device log map vs the host cv2 path
    out, _ = hi.estimate_from_rgb(frame, dets)
    recs = hi_mod.hand_records_device(model.engine, out, [d[0] for d in dets])
    for i, d in enumerate(dets):
        ref = hi_mod.hand_record(out, i, d[0] == 'right')
        assert set(recs[i]) == set(ref) and recs[i]['is_right'] == ref['is_right']
        assert np.abs(recs[i]['theta'] - ref['theta']).max() < 5e-6 and np.array_equal(recs[i]['betas'], ref['betas'])
    npy_dir = tmp_path / 'npy'
    hi_mod.save_image_results(str(npy_dir), 'img0', {'left': recs[1], 'right': recs[0]})
    objs = hi_mod.reconstruct_and_save_obj_with_wrapper(str(npy_dir), str(tmp_path / 'obj'), hi)
    assert len(objs) == 1
    vs = np.array([[float(x) for x in l.split()[1:4]] for l in open(objs[0]) if l.startswith('v ')])
    fs = np.array([[int(x) for x in l.split()[1:4]] for l in open(objs[0]) if l.startswith('f ')])
    assert vs.shape == (2 * 778, 3) and fs.shape == (2 * 1538, 3) and fs.min() == 1 and fs.max() == 2 * 778
    mano = model.mano.data
    for k, rec in enumerate((recs[0], recs[1])):                 # right first, then left (infer.py:1354)
        aa = torch.from_numpy(np.concatenate([rec['pose_global'], rec['pose_hand']]).reshape(16, 3)).float()
        R = hi_mod.axis_angle_to_rotation_matrix_torch(aa).reshape(1, 16, 3, 3)
        v_ref, _ = mano_forward(R[:, :1], R[:, 1:], torch.from_numpy(rec['betas']).reshape(1, 10), mano)
        v_ref = v_ref[0].numpy().copy()
        if not rec['is_right']:
            v_ref[:, 0] = -v_ref[:, 0]
        v_ref += rec['cam_t']
        assert np.abs(vs[k * 778:(k + 1) * 778] - v_ref).max() < 5e-5
    assert np.array_equal(fs[1538:] - 778, np.asarray(model.mano.faces)[:, [0, 2, 1]] + 1)


@pytest.mark.parametrize('h,w', [(1080, 1920), (480, 640), (240, 320), (640, 360), (333, 517), (384, 640)])
@pytest.mark.parametrize('half', [False, True])
def test_letterbox_matches_the_yolo_loader(model, h, w, half):
    """hm_letterbox against LoadImage.process_img + Detector.detect's tensor conversion (cv2.resize / copyMakeBorder through the
    oracle): bit-identical, down- and up-scaling, host and device frames."""
    from oracle import letterbox as ol
    frame = synthetic_frame(h, w, h + w)
    ref = ol.process_img(frame, 640, 32, half=half)
    img, dev_frame = model.engine.letterbox(frame, 640, 32, half=half)
    assert img.shape == ref.shape and img.dtype == ref.dtype
    assert torch.equal(img.cpu(), ref)
    img2, _ = model.engine.letterbox(torch.from_numpy(frame).cuda(), 640, 32, half=half)
    assert torch.equal(img2.cpu(), ref)
    # the uploaded frame feeds the crop kernel without a second host -> device copy
    from oracle import crop as oc
    dets = [['right', [w * 0.2, h * 0.2, w * 0.5, h * 0.6]], ['left', [w * 0.55, h * 0.3, w * 0.9, h * 0.8]]]
    crops, _, _, _ = model.engine.preprocess(dev_frame, [d[1] for d in dets], [1, 0])
    assert torch.equal(crops.cpu(), torch.from_numpy(oc.prepare_batch_bbox(frame, dets)['img']))


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
def test_scale_boxes_matches_scale_coords(model, dtype):
    """hm_scale_boxes against scale_coords(...).round() + the labelling of Detector.detect, in the detections' own dtype."""
    from oracle import letterbox as ol
    g = torch.Generator().manual_seed(3)
    for (h, w) in [(1080, 1920), (480, 640), (517, 333)]:
        oh = ol.process_img(np.zeros((h, w, 3), np.uint8)).shape
        n = 37
        xy1 = torch.rand(n, 2, generator=g) * torch.tensor([oh[2] * 0.8, oh[1] * 0.8]) - 5.0
        wh = torch.rand(n, 2, generator=g) * 200 + 2
        det = torch.cat([xy1, xy1 + wh, torch.rand(n, 1, generator=g), torch.randint(0, 2, (n, 1), generator=g).float()], 1).to(dtype)
        ref = ol.detections_to_list(det.cuda() if dtype == torch.float16 else det, oh[1:], (h, w, 3))
        got = model.engine.scale_boxes(det.cuda(), oh[1:], (h, w, 3))
        assert got == ref
    assert model.engine.scale_boxes(torch.zeros(0, 6).cuda(), (384, 640), (1080, 1920, 3)) == []
