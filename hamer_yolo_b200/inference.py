"""Drop-in mirror of the reference's pipeline wrapper ``hamer_inference`` (hamer/infer.py:117-528)
and of the per-image ``.npy`` writer (hamer/infer.py:1082-1096, 1223-1318).

  hamer_inference.__init__            infer.py:118-146
  hamer_inference.prepare_batch_bbox  infer.py:154-259   -> hm_preprocess (crop_warp.cuh)
  hamer_inference.estimate_from_rgb   infer.py:354-528   -> hm_forward + camera post-processing
  matrix_to_axis_angle                infer.py:1082-1096 (cv2.Rodrigues, host, kept for exactness)
  process_batch_manopara (npy layout) infer.py:1296-1312 -> hand_record / save_image_results
"""
import os

import numpy as np
import torch

from .hamer import HAMER, load_hamer


def perspective_projection(points, translation, focal_length):
    """hamer/hamer/utils/geometry.py:72-109 with identity rotation and zero camera centre."""
    p = points + translation.unsqueeze(1)
    p = p / p[:, :, -1].unsqueeze(-1)
    return p[:, :, :2] * focal_length.unsqueeze(1)


def custom_cam_crop_to_full(cam_bbox, box_center, box_size, img_size, fx, fy, cx, cy):
    """hamer/hamer/utils/renderer.py:24-74 (depth_refine=None)."""
    b = cam_bbox.shape[0]
    dev = cam_bbox.device

    def vec(v):
        v = torch.as_tensor(v, dtype=torch.float32, device=dev)
        return v.reshape(-1).expand(b) if v.numel() == 1 else v.reshape(b)

    fx, fy, cx, cy = vec(fx), vec(fy), vec(cx), vec(cy)
    bs = box_size * cam_bbox[:, 0] + 1e-9
    tz = 2 * fx / bs
    tx = (2 * (box_center[:, 0] - cx) / bs) + cam_bbox[:, 1]
    ty = (2 * (box_center[:, 1] - cy) / bs) + cam_bbox[:, 2]
    if not torch.allclose(fx, fy):
        ty = ty * (fx / fy)
    return torch.stack([tx, ty, tz], dim=-1)


META_COLS = 16


class _Batch(dict):
    """The dict prepare_batch_bbox returns.  The small per-hand tensors ('box_center', 'box_size', 'img_size',
    'do_flip', 'trans', 'inv_trans') are views of one device tensor that is uploaded once by meta_dev()
    (estimate_from_rgb adds the intrinsics to the same upload before calling it)."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self._meta_host = None
        self._meta_dev = None
        self._dev = None

    def meta_dev(self):
        if self._meta_dev is None:
            m = self._meta_host.to(self._dev, non_blocking=True)
            self._meta_dev = m
            n = m.shape[0]
            dict.__setitem__(self, 'box_center', m[:, 0:2])
            dict.__setitem__(self, 'box_size', m[:, 2])
            dict.__setitem__(self, 'img_size', m[:, 3:5])
            dict.__setitem__(self, 'do_flip', m[:, 5])
            dict.__setitem__(self, 'inv_trans', m[:, 6:12].reshape(n, 2, 3))    # the reference returns the forward
            dict.__setitem__(self, 'trans', m[:, 6:12].reshape(n, 2, 3).clone())  # transform twice
        return self._meta_dev


class hamer_inference:
    """Same constructor contract as the reference: ``cfg`` carries ``ckpt_path`` (and optionally
    ``mano_path``); alternatively pass an already built ``model``."""

    def __init__(self, cfg=None, model=None, device='cuda:0', max_batch=64, mano=None):
        if model is None:
            if getattr(cfg, 'use_onnx', False):
                raise NotImplementedError('the ONNX backend is out of scope (no multi-backend dispatch)')
            model, model_cfg = load_hamer(cfg.ckpt_path, mano=mano, device=device, max_batch=max_batch)
        else:
            model_cfg = model.cfg
        self.use_onnx = False
        self.model = model.eval()
        self.device = model.device
        self.cfg = model_cfg
        self.mean = 255. * np.array(self.cfg.MODEL.IMAGE_MEAN)
        self.std = 255. * np.array(self.cfg.MODEL.IMAGE_STD)

    # -- crop / warp / normalise -----------------------------------------------------------------
    def prepare_batch_bbox(self, img_0, bboxs, _defer_upload=False):
        """infer.py:154-259.  Returns the same dict; 'img' is already on the device (the
        reference moves it there in estimate_from_rgb, infer.py:366-368)."""
        boxes, right = [], []
        for bbox in bboxs:
            if isinstance(bbox, (list, tuple)) and len(bbox) == 2:
                hand_cls, coords = bbox[0], bbox[1]
                if not (isinstance(coords, (list, tuple, np.ndarray)) and len(coords) == 4):
                    raise ValueError(f'Invalid coordinates format: Expected [x1,y1,x2,y2], got {coords}')
            else:
                raise ValueError(f'Invalid bbox format: Expected [class, coords], got {bbox}')
            boxes.append([float(v) for v in coords])
            right.append(1 if hand_cls == 'right' else 0)
        eng = self.model.engine
        n = len(boxes)
        imgs, centers, sizes, trans = [], [], [], []
        for lo in range(0, n, eng.max_batch):
            hi = min(n, lo + eng.max_batch)
            i, c, s, t = eng.preprocess(img_0, boxes[lo:hi], right[lo:hi])
            imgs.append(i); centers.append(c); sizes.append(s); trans.append(t)
        dev = self.device
        H, W = img_0.shape[:2]
        # every per-hand scalar goes to the device in ONE pinned upload; the dict entries are views of it
        # (layout = the `meta` argument of hm_camera_postprocess, include/hamer_b200.h; columns 12..15 are the
        # intrinsics estimate_from_rgb fills in)
        meta = torch.zeros(n, META_COLS, dtype=torch.float32, pin_memory=True)
        mv = meta.numpy()
        mv[:, 0:2] = np.concatenate(centers)
        mv[:, 2] = np.concatenate(sizes)
        mv[:, 3] = W
        mv[:, 4] = H
        mv[:, 5] = [0.0 if r else 1.0 for r in right]
        mv[:, 6:12] = np.concatenate(trans).reshape(n, 6)
        batch = _Batch(img=torch.cat(imgs) if len(imgs) > 1 else imgs[0])
        batch._meta_host, batch._dev = meta, dev
        if not _defer_upload:
            batch.meta_dev()
        return batch

    # -- full per-image estimate -----------------------------------------------------------------
    @torch.no_grad()
    def estimate_from_rgb(self, img_0, detections, k_real=None):
        if not isinstance(detections, list) or len(detections) == 0:
            raise ValueError('Invalid detections format')                     # infer.py:360-361
        batch = self.prepare_batch_bbox(img_0, detections, _defer_upload=True)
        mv = batch._meta_host.numpy()
        W, H = float(mv[0, 3]), float(mv[0, 4])
        if k_real is not None:
            k = (k_real.detach().cpu().numpy() if torch.is_tensor(k_real) else np.asarray(k_real)).astype(np.float32)
            if k.ndim == 2:
                mv[:, 12], mv[:, 13], mv[:, 14], mv[:, 15] = k[0, 0], k[1, 1], k[0, 2], k[1, 2]
            else:
                mv[:, 12], mv[:, 13], mv[:, 14], mv[:, 15] = k[:, 0, 0], k[:, 1, 1], k[:, 0, 2], k[:, 1, 2]
            per_hand_k = k.ndim != 2
        else:
            # scaled_focal_length = FOCAL_LENGTH / IMAGE_SIZE * max(img_size)  (infer.py:489-493), fp32 like torch
            f = np.float32(self.cfg.EXTRA.FOCAL_LENGTH / self.cfg.MODEL.IMAGE_SIZE) * np.float32(max(W, H))
            mv[:, 12] = f
            mv[:, 13] = f
            mv[:, 14] = np.float32(W) / np.float32(2.0)
            mv[:, 15] = np.float32(H) / np.float32(2.0)
            per_hand_k = True
        fxfy_differ = not np.allclose(mv[:, 12], mv[:, 13])              # renderer.py:63 torch.allclose(fx, fy)
        meta = batch.meta_dev()                                           # the one small host -> device copy of the frame
        out, params = self.model(batch)
        # kp3d x-flip (infer.py:392, in place), cam y-flip (:402-409), cam_crop_to_full (renderer.py:24-74) and the
        # full-image projection (:474-486 / geometry.py:72-109) in one kernel, bit-identical to the torch expressions
        cam_t_full, kp2d = self.model.engine.camera_postprocess(out['pred_cam'], out['pred_keypoints_3d'], meta,
                                                                fxfy_differ, k_real is not None)
        do_flip = batch['do_flip']
        scaled_focal_length = meta[:, 12] if per_hand_k else meta[:1, 12]
        out['pred_keypoints_2d_full'] = kp2d
        out['pred_cam_t_full'] = cam_t_full
        out['img'] = batch['img']
        out['focal_length'] = scaled_focal_length
        out['trans'] = batch['trans']
        out['do_flip'] = do_flip
        out['inv_trans'] = batch['inv_trans']
        return out, params


# -- npy output layout (infer.py:1082-1096, 1279-1312) --------------------------------------------
def matrix_to_axis_angle(rot_mats):
    import cv2
    if isinstance(rot_mats, np.ndarray) and rot_mats.ndim == 2:
        rot_mats = [rot_mats]
    return np.concatenate([cv2.Rodrigues(np.asarray(m))[0].flatten() for m in rot_mats])


def hand_record(output, index, is_right):
    """One hand's dict exactly as process_batch_manopara stores it (infer.py:1279-1303)."""
    import cv2
    mp = output['pred_mano_params']
    betas = mp['betas'][index].detach().cpu().numpy().squeeze()
    hand_pose_aa = matrix_to_axis_angle(mp['hand_pose'][index].detach().cpu().numpy())
    go = mp['global_orient'][index].detach().cpu().numpy().squeeze()
    if go.ndim == 3:
        go = go[0]
    global_aa = cv2.Rodrigues(go)[0].flatten()
    cam_t = output['pred_cam_t_full'][index].detach().cpu().numpy().squeeze()
    return {'betas': betas, 'theta': np.concatenate((global_aa, hand_pose_aa)),
            'pose_hand': hand_pose_aa, 'pose_global': global_aa, 'cam_t': cam_t, 'is_right': is_right}


def process_image(hamer, image, detection_list, k_real=None):
    """Inner loop of process_batch_manopara (infer.py:1268-1308) for one image, batched over the
    hands instead of one model call per hand.  Returns {'left': d|None, 'right': d|None}; a later
    hand with the same label overwrites an earlier one, as in the reference."""
    results = {'left': None, 'right': None}
    if not detection_list:
        return results
    output, _ = hamer.estimate_from_rgb(image, list(detection_list), k_real)
    for i, bbox in enumerate(detection_list):
        label = bbox[0]
        results[label] = hand_record(output, i, label == 'right')
    return results


def save_image_results(output_folder, file_name, image_results):
    """np.save(f"{stem}.npy", {'left':..., 'right':...}) (infer.py:1310-1312); read back with
    np.load(..., allow_pickle=True).item() (infer.py:1350)."""
    os.makedirs(output_folder, exist_ok=True)
    path = os.path.join(output_folder, f'{file_name}.npy')
    np.save(path, image_results)
    return path


# -- per-hand npz (process_batch, infer.py:908-1036) ------------------------------------------------
def save_hand_npz(output_folder, file_name, output, index, hand_label, saved_counts):
    """One ``{stem}_{label}[_{k}].npz`` per detected hand exactly as process_batch writes it (infer.py:1003-1025):
    betas (1,10), global_orient (1,1,3,3), hand_pose (1,15,3,3), cam_t (1,3) = pred_cam_t_full, is_right (1,) bool
    (= do_flip == 0, infer.py:968); a second hand with the same label gets the suffix ``_2`` (infer.py:1003-1006)."""
    mp = output['pred_mano_params']
    sl = slice(index, index + 1)
    suffix = f'_{hand_label}'
    if saved_counts[hand_label] > 0:
        suffix += f'_{saved_counts[hand_label] + 1}'
    os.makedirs(output_folder, exist_ok=True)
    path = os.path.join(output_folder, f'{file_name}{suffix}.npz')
    np.savez(path,
             betas=mp['betas'][sl].detach().cpu().numpy(),
             global_orient=mp['global_orient'][sl].detach().cpu().numpy(),
             hand_pose=mp['hand_pose'][sl].detach().cpu().numpy(),
             cam_t=output['pred_cam_t_full'][sl].detach().cpu().numpy(),
             is_right=(output['do_flip'][sl] == 0).detach().cpu().numpy())
    saved_counts[hand_label] += 1
    return path


def process_image_npz(hamer, image, detection_list, output_folder, file_name, k_real=None):
    """Inner loop of process_batch (infer.py:955-1030) for one image, batched over its hands."""
    if not detection_list:
        return []
    output, _ = hamer.estimate_from_rgb(image, list(detection_list), k_real)
    saved = {'left': 0, 'right': 0}
    return [save_hand_npz(output_folder, file_name, output, i, bbox[0], saved) for i, bbox in enumerate(detection_list)]


# -- device log map for the npy records --------------------------------------------------------------
def hand_records_device(engine, output, labels):
    """The dicts of hand_record() for every hand of ``output`` with the axis-angle vectors computed on the device
    (hm_rotmat_to_axis_angle: one launch and one small copy for all hands instead of 16 cv2.Rodrigues calls per hand)."""
    mp = output['pred_mano_params']
    B = mp['betas'].shape[0]
    rot = torch.cat([mp['global_orient'].reshape(B, 1, 3, 3), mp['hand_pose'].reshape(B, 15, 3, 3)], dim=1)
    aa = engine.rotmat_to_axis_angle(rot).reshape(B, 48).cpu().numpy()
    betas = mp['betas'].detach().cpu().numpy()
    cam_t = output['pred_cam_t_full'].detach().cpu().numpy()
    return [{'betas': betas[i], 'theta': aa[i].copy(), 'pose_hand': aa[i, 3:].copy(), 'pose_global': aa[i, :3].copy(),
             'cam_t': cam_t[i], 'is_right': labels[i] == 'right'} for i in range(B)]


# -- OBJ reconstruction from the npy records (reconstruct_and_save_obj_with_wrapper, infer.py:1321-1436) --------------
def axis_angle_to_rotation_matrix_torch(rvec_tensor):
    """infer.py:65-82 (batch Rodrigues with theta = |r| + 1e-8)."""
    theta = torch.norm(rvec_tensor, dim=1, keepdim=True) + 1e-8
    r_hat = rvec_tensor / theta
    cos = torch.cos(theta)
    z = torch.zeros(theta.shape[0], dtype=rvec_tensor.dtype, device=rvec_tensor.device)
    m = torch.stack([z, -r_hat[:, 2], r_hat[:, 1], r_hat[:, 2], z, -r_hat[:, 0], -r_hat[:, 1], r_hat[:, 0], z], dim=1).reshape(-1, 3, 3)
    eye = torch.eye(3, dtype=rvec_tensor.dtype, device=rvec_tensor.device).unsqueeze(0).expand(rvec_tensor.shape[0], -1, -1)
    A = r_hat.unsqueeze(2) * r_hat.unsqueeze(1)
    return cos.unsqueeze(2) * eye + (1 - cos.unsqueeze(2)) * A + torch.sin(theta).unsqueeze(2) * m


RIGHT_COLOR, LEFT_COLOR = (100, 200, 100, 255), (200, 100, 100, 255)     # infer.py:1405, 1416


def hand_meshes_from_record(engine, faces, data):
    """Vertices / faces / colour of the hands of one npy record, as infer.py:1350-1421 builds them: axis-angle -> matrices,
    MANO forward (hm_mano_lbs), + cam_t; left hands get x negated and the face winding flipped.  Hands in the reference's
    order ('right', then 'left')."""
    dev = engine.device
    todo = [(t, data[t]) for t in ('right', 'left') if data.get(t) is not None]
    if not todo:
        return []
    betas = torch.stack([torch.as_tensor(np.asarray(d['betas']).reshape(10), dtype=torch.float32) for _, d in todo]).to(dev)
    aa = torch.stack([torch.cat([torch.as_tensor(np.asarray(d['pose_global']).reshape(1, 3), dtype=torch.float32),
                                 torch.as_tensor(np.asarray(d['pose_hand']).reshape(15, 3), dtype=torch.float32)]) for _, d in todo]).to(dev)
    rot = axis_angle_to_rotation_matrix_torch(aa.reshape(-1, 3)).reshape(len(todo), 16, 3, 3).contiguous()
    verts, _ = engine.mano_lbs(betas, rot)
    verts = verts.cpu().numpy()
    out = []
    for i, (_, d) in enumerate(todo):
        v = verts[i].copy()
        f = np.asarray(faces).astype(np.int32)
        cam_t = np.asarray(d['cam_t'], dtype=v.dtype).reshape(3)
        if d['is_right']:
            v += cam_t
            color = RIGHT_COLOR
        else:
            v[:, 0] = -v[:, 0]
            f = f[:, [0, 2, 1]]
            v += cam_t                                   # (the reference's "Tx flip" is a no-op: real_cam_t[0] = real_cam_t[0])
            color = LEFT_COLOR
        out.append((v, f, color))
    return out


def write_obj(path, meshes):
    """Concatenated Wavefront OBJ of (vertices, faces, rgba) meshes with per-vertex colours (``v x y z r g b``, colours
    in [0, 1]) - the layout trimesh's exporter gives a vertex-coloured mesh (trimesh itself is not a dependency here)."""
    lines, base = ['# hamer_yolo_b200 (layout of trimesh.exchange.obj.export_obj)'], 0
    for v, f, color in meshes:
        r, g, b = (c / 255.0 for c in color[:3])
        lines += [f'v {x:.8f} {y:.8f} {z:.8f} {r:.8f} {g:.8f} {b:.8f}' for x, y, z in v]
    for v, f, color in meshes:
        lines += [f'f {a + 1 + base} {b_ + 1 + base} {c + 1 + base}' for a, b_, c in f]
        base += len(v)
    with open(path, 'w') as fh:
        fh.write('\n'.join(lines) + '\n')
    return path


def reconstruct_and_save_obj_with_wrapper(npy_folder, output_obj_folder, hamer_instance):
    """infer.py:1321-1436: every ``*.npy`` record of ``npy_folder`` becomes ``{stem}.obj`` (both hands in one file)."""
    import glob
    os.makedirs(output_obj_folder, exist_ok=True)
    model = hamer_instance.model if hasattr(hamer_instance, 'model') else hamer_instance
    written = []
    for npy_path in sorted(glob.glob(os.path.join(npy_folder, '*.npy'))):
        stem = os.path.splitext(os.path.basename(npy_path))[0]
        data = np.load(npy_path, allow_pickle=True).item()
        meshes = hand_meshes_from_record(model.engine, model.mano.faces, data)
        if meshes:
            written.append(write_obj(os.path.join(output_obj_folder, f'{stem}.obj'), meshes))
    return written
