"""CPU emulation of the rounding points of the CUDA backbone (bf16 GEMM operands, fp32 accumulate /
residual / LN / softmax) to predict the parity margin before spending GPU time.
Test infrastructure; run:  python tests/tools/emulate_bf16.py [B]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pipeline import OHamer, perspective_projection  # noqa: E402
from oracle.mano import mano_forward  # noqa: E402
from hamer_yolo_b200.mano_data import synthetic_mano  # noqa: E402


ACT = torch.bfloat16


def bf(t):
    """round to the activation (A-operand) type"""
    return t.to(ACT).float()


def wround(t):
    return t.to(torch.bfloat16).float()


@torch.no_grad()
def emulated_backbone(vit, x):
    x = vit.patch_embed(x)
    x = x + vit.pos_embed[:, 1:] + vit.pos_embed[:, :1]
    for blk in vit.blocks:
        a = blk.attn
        xn = bf(blk.norm1(x))
        B, N, C = xn.shape
        qkv = bf(a.qkv(xn)).reshape(B, N, 3, a.num_heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        s = q @ k.transpose(-2, -1)
        p = torch.exp((s - s.amax(dim=-1, keepdim=True)) * a.scale)
        o = (bf(p) @ v) / p.sum(dim=-1, keepdim=True)
        o = bf(o.transpose(1, 2).reshape(B, N, -1))
        x = x + a.proj(o)
        xn = bf(blk.norm2(x))
        h = bf(blk.mlp.act(blk.mlp.fc1(xn)))
        x = x + blk.mlp.fc2(h)
    return vit.last_norm(x)


def report(tag, out, ref):
    mm = lambda a, b: (a - b).norm(dim=-1).max().item() * 1000
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
    rot = torch.cat([out['pred_mano_params']['global_orient'], out['pred_mano_params']['hand_pose']], 1)
    rot_r = torch.cat([ref['pred_mano_params']['global_orient'], ref['pred_mano_params']['hand_pose']], 1)
    print(f"{tag}: verts {mm(out['pred_vertices'], ref['pred_vertices']):.4f} mm  "
          f"joints {mm(out['pred_keypoints_3d'], ref['pred_keypoints_3d']):.4f} mm  "
          f"kp2d {(out['pred_keypoints_2d'] - ref['pred_keypoints_2d']).abs().max().item() * 256:.4f} px  "
          f"rot rel {rel(rot, rot_r):.2e}  betas rel {rel(out['pred_mano_params']['betas'], ref['pred_mano_params']['betas']):.2e}  "
          f"cam rel {rel(out['pred_cam'], ref['pred_cam']):.2e}")


if __name__ == '__main__':
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    if len(sys.argv) > 2 and sys.argv[2] == 'fp16':
        ACT = torch.float16
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    model = OHamer(synthetic_mano(0))
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.copy_(wround(p))
    g = torch.Generator().manual_seed(1234)
    img = wround(torch.randn(B, 3, 256, 256, generator=g))
    t0 = time.time()
    ref, _ = model({'img': img})
    print('oracle fwd', time.time() - t0, 's')
    feats = emulated_backbone(model.backbone, img[:, :, :, 32:-32])
    print('feats max abs diff', (feats - ref['conditioning_feats']).abs().max().item(),
          'ref absmean', ref['conditioning_feats'].abs().mean().item())
    params, cam = model.mano_head(feats)
    focal = 5000.0 * torch.ones(B, 2)
    cam_t = torch.stack([cam[:, 1], cam[:, 2], 2 * focal[:, 0] / (256 * cam[:, 0] + 1e-9)], dim=-1)
    verts, joints = mano_forward(params['global_orient'], params['hand_pose'], params['betas'], model.mano_data)
    out = {'pred_cam': cam, 'pred_mano_params': params, 'pred_vertices': verts, 'pred_keypoints_3d': joints,
           'pred_keypoints_2d': perspective_projection(joints, cam_t, focal / 256)}
    report('bf16-operand emulation', out, ref)
    rot = torch.cat([params['global_orient'], params['hand_pose']], 1)
    rot_r = torch.cat([ref['pred_mano_params']['global_orient'], ref['pred_mano_params']['hand_pose']], 1)
    l2 = lambda a, b: ((a - b).flatten(1).norm(dim=1) / b.flatten(1).norm(dim=1)).max().item()
    print('per-hand L2-relative: rot', l2(rot, rot_r), 'betas', l2(params['betas'], ref['pred_mano_params']['betas']),
          'cam', l2(cam, ref['pred_cam']))
    print('per-joint rot max abs', (rot - rot_r).abs().amax(dim=(0, 2, 3)))
