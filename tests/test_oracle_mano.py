"""MANO oracle: golden from the reference's in-tree manopth on the real MANO_RIGHT.pkl, plus
model-independent properties on the synthetic model."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE
from oracle.mano import lbs, mano_forward, MANO_TO_OPENPOSE
from hamer_yolo_b200.mano_data import load_mano_pkl, synthetic_mano

PKL = f'{REFERENCE}/rootnet/KeypointFusion/MANO/MANO_RIGHT.pkl'


def rand_rot(n, seed, sigma=0.6):
    g = torch.Generator().manual_seed(seed)
    aa = torch.randn(n, 3, generator=g) * sigma
    th = aa.norm(dim=1, keepdim=True).clamp_min(1e-8)
    k = aa / th
    K = torch.zeros(n, 3, 3)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -k[:, 2], k[:, 1], k[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -k[:, 0], -k[:, 1], k[:, 0]
    s, c = torch.sin(th)[:, :, None], torch.cos(th)[:, :, None]
    return torch.eye(3)[None] + s * K + (1 - c) * (K @ K)


@pytest.mark.skipif(not os.path.exists(PKL), reason='licensed MANO pickle only exists in the authoring container')
def test_lbs_matches_reference_manopth_on_real_mano():
    g = np.load(os.path.join(GOLDEN, 'mano_lbs_real.npz'))
    m = load_mano_pkl(PKL)
    verts, j16 = lbs(torch.tensor(g['betas']), torch.tensor(g['rotmats']), m)
    assert np.abs(verts[:, g['vert_ids']].numpy() - g['verts_sub']).max() < 1e-6      # < 0.001 mm
    inv = {MANO_TO_OPENPOSE[i]: i for i in range(21)}
    j16_ref = np.stack([g['joints21_manopth_tips'][:, inv[k]] for k in range(16)], axis=1)
    assert np.abs(j16.numpy() - j16_ref).max() < 1e-6
    # loader sanity: layout smplx builds
    assert m['posedirs'].shape == (135, 2334) and m['shapedirs'].shape == (778, 3, 10)
    assert (m['lbs_weights'] != 0).sum(1).max() <= 8 and m['parents'][0] == -1


def test_identity_pose_zero_beta_returns_template():
    m = synthetic_mano(0)
    B = 3
    rot = torch.eye(3).expand(B, 16, 3, 3).contiguous()
    verts, joints = mano_forward(rot[:, :1], rot[:, 1:], torch.zeros(B, 10), m)
    assert torch.allclose(verts, torch.tensor(m['v_template']).expand(B, -1, -1), atol=1e-6)
    j_rest = torch.tensor(m['J_regressor']) @ torch.tensor(m['v_template'])
    assert torch.allclose(joints[:, 0], j_rest[0].expand(B, -1), atol=1e-6)


def test_global_rotation_equivariance():
    """Rotating global_orient by Q rotates the whole hand about the root joint."""
    m = synthetic_mano(0)
    B = 4
    rot = rand_rot(B * 16, 1).reshape(B, 16, 3, 3)
    betas = torch.randn(B, 10, generator=torch.Generator().manual_seed(2))
    v0, j0 = mano_forward(rot[:, :1], rot[:, 1:], betas, m)
    Q = rand_rot(B, 3, sigma=1.0)
    rot2 = rot.clone()
    rot2[:, 0] = Q @ rot[:, 0]
    v1, j1 = mano_forward(rot2[:, :1], rot2[:, 1:], betas, m)
    root = j0[:, :1]                                   # OpenPose joint 0 = MANO root
    assert torch.allclose(j1[:, :1], root, atol=1e-6)
    assert torch.allclose(v1 - root, (v0 - root) @ Q.transpose(1, 2), atol=2e-6)
    assert torch.allclose(j1 - root, (j0 - root) @ Q.transpose(1, 2), atol=2e-6)


def test_joint_order_and_tips():
    m = synthetic_mano(0)
    rot = rand_rot(16, 4).reshape(1, 16, 3, 3)
    betas = torch.zeros(1, 10)
    verts, joints = mano_forward(rot[:, :1], rot[:, 1:], betas, m)
    _, j16 = lbs(betas, rot, m)
    # mano_wrapper.py:24,37-39: 16 joints ++ 5 tip vertices, then the OpenPose permutation
    full = torch.cat([j16, verts[:, [744, 320, 443, 554, 671]]], dim=1)
    assert torch.equal(joints, full[:, list(MANO_TO_OPENPOSE)])
    assert torch.equal(joints[:, 4], verts[:, 744]) and torch.equal(joints[:, 20], verts[:, 671])
