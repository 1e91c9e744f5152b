"""Oracle: model assembly + camera post-processing (fp32, CPU).  TEST INFRASTRUCTURE ONLY.

Restates
  hamer/hamer/models/hamer.py:99-156          HAMER.forward_step
  hamer/hamer/utils/geometry.py:72-109        perspective_projection
  hamer/hamer/utils/renderer.py:24-74         custom_cam_crop_to_full
  hamer/infer.py:354-528                      hamer_inference.estimate_from_rgb (post-processing)
"""
import numpy as np
import torch
import torch.nn as nn

from .backbone import OViT
from .head import OManoHead, default_mean_params
from .mano import mano_forward

FOCAL_LENGTH = 5000.0   # cfg.EXTRA.FOCAL_LENGTH (configs_hydra/experiment/default.yaml:12)
IMAGE_SIZE = 256        # cfg.MODEL.IMAGE_SIZE


def perspective_projection(points, translation, focal_length):
    """geometry.py:72-109 with rotation = I and camera centre = 0."""
    p = points + translation.unsqueeze(1)
    p = p / p[:, :, -1].unsqueeze(-1)
    return p[:, :, :2] * focal_length.unsqueeze(1)


class OHamer(nn.Module):
    """Backbone + head + MANO, wired as HAMER.forward_step (hamer.py:99-156)."""

    def __init__(self, mano, mean_params=None):
        super().__init__()
        self.backbone = OViT()
        self.mano_head = OManoHead(mean_params or default_mean_params())
        self.mano_data = mano
        self.eval()

    @torch.no_grad()
    def forward(self, batch):
        x = batch['img']
        B = x.shape[0]
        feats = self.backbone(x[:, :, :, 32:-32])                       # hamer.py:119
        params, pred_cam = self.mano_head(feats)                         # :121
        out = {'pred_cam': pred_cam,
               'pred_mano_params': {k: v.clone() for k, v in params.items()}}
        focal = FOCAL_LENGTH * torch.ones(B, 2, dtype=x.dtype, device=x.device)
        cam_t = torch.stack([pred_cam[:, 1], pred_cam[:, 2],
                             2 * focal[:, 0] / (IMAGE_SIZE * pred_cam[:, 0] + 1e-9)], dim=-1)  # :131-134
        out['pred_cam_t'] = cam_t
        out['focal_length'] = focal
        verts, joints = mano_forward(params['global_orient'], params['hand_pose'], params['betas'],
                                     self.mano_data)                    # :143
        out['pred_keypoints_3d'] = joints.reshape(B, -1, 3)
        out['pred_vertices'] = verts.reshape(B, -1, 3)
        out['pred_keypoints_2d'] = perspective_projection(joints, cam_t, focal / IMAGE_SIZE)  # :150-152
        params = dict(params)
        params['trans'] = cam_t
        out['conditioning_feats'] = feats   # oracle-only extra, for stage-wise checks
        return out, params


def cam_crop_to_full(cam_bbox, box_center, box_size, img_size, fx, fy, cx, cy):
    """renderer.py:24-74 (depth_refine=None branch).  fx..cy: tensors of shape (B,)."""
    bs = box_size * cam_bbox[:, 0] + 1e-9
    tz = 2 * fx / bs
    tx = (2 * (box_center[:, 0] - cx) / bs) + cam_bbox[:, 1]
    ty = (2 * (box_center[:, 1] - cy) / bs) + cam_bbox[:, 2]
    if not torch.allclose(fx, fy):
        ty = ty * (fx / fy)
    return torch.stack([tx, ty, tz], dim=-1)


@torch.no_grad()
def postprocess(out, batch, k_real=None):
    """infer.py:382-528: flip quirk, flip-corrected camera, full-image translation and keypoints.
    Mutates out['pred_keypoints_3d'] in place exactly like the reference (infer.py:392)."""
    pred_cam = out['pred_cam']
    box_center, box_size = batch['box_center'].float(), batch['box_size'].float()
    img_size, do_flip = batch['img_size'].float(), batch['do_flip'].float()
    B = pred_cam.shape[0]
    kp3d = out['pred_keypoints_3d']
    kp3d[:, :, 0] = kp3d[:, :, 0] * do_flip.unsqueeze(1)                 # infer.py:392
    cam = pred_cam.clone()
    cam[:, 1] = cam[:, 1] * (1.0 - 2.0 * do_flip.view(-1))               # :402-409
    if k_real is not None:
        k = torch.as_tensor(np.asarray(k_real), dtype=torch.float32)
        fx, fy, cx, cy = (k[0, 0].repeat(B), k[1, 1].repeat(B), k[0, 2].repeat(B), k[1, 2].repeat(B))
        cam_t_full = cam_crop_to_full(cam, box_center, box_size, img_size, fx, fy, cx, cy)
        kp_cam = kp3d + cam_t_full.unsqueeze(1)
        depth = kp_cam[:, :, 2:3] + 1e-9
        u = kp_cam[:, :, 0:1] / depth * fx.view(-1, 1, 1) + cx.view(-1, 1, 1)
        v = kp_cam[:, :, 1:2] / depth * fy.view(-1, 1, 1) + cy.view(-1, 1, 1)
        kp2d = torch.cat([u, v], dim=-1)
        focal = k[0, 0].unsqueeze(0)
    else:
        focal = FOCAL_LENGTH / IMAGE_SIZE * img_size.max(dim=1)[0]        # :486-491
        cam_t_full = cam_crop_to_full(cam, box_center, box_size, img_size, focal, focal,
                                      img_size[:, 0] / 2.0, img_size[:, 1] / 2.0)
        kp2d = perspective_projection(kp3d, cam_t_full, torch.stack([focal, focal], dim=1))
    res = dict(out)
    res.update({'pred_keypoints_2d_full': kp2d, 'pred_cam_t_full': cam_t_full, 'img': batch['img'],
                'focal_length': focal, 'trans': batch['trans'], 'do_flip': do_flip,
                'inv_trans': batch['inv_trans']})
    return res
