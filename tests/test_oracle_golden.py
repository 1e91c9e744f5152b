"""Oracle vs the golden vectors minted from the REFERENCE's own modules
(tests/golden/make_golden.py imported hamer/hamer/models/backbones/vit.py, heads/mano_head.py,
components/pose_transformer.py, utils/geometry.py from /root/reference)."""
import os

import numpy as np
import torch

from conftest import GOLDEN


def test_oracle_reproduces_reference_forward(oracle_model, seeded_images):
    g = np.load(os.path.join(GOLDEN, 'hamer_fwd_seed0.npz'))
    out, params = oracle_model({'img': seeded_images})
    feats = out['conditioning_feats']
    # same weights (seed 0), same input (seed 1234), same maths -> bit-identical up to thread count
    np.testing.assert_allclose(feats[:, ::16, ::16].numpy(), g['feats_sub'], rtol=0, atol=2e-5)
    assert abs(feats.abs().mean().item() - float(g['feats_absmean'])) < 1e-5
    for k in ('global_orient', 'hand_pose', 'betas'):
        np.testing.assert_allclose(out['pred_mano_params'][k].numpy(), g[k], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out['pred_cam'].numpy(), g['pred_cam'], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out['pred_cam_t'].numpy(), g['pred_cam_t'], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(out['pred_vertices'].numpy(), g['pred_vertices'], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out['pred_keypoints_3d'].numpy(), g['pred_keypoints_3d'], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out['pred_keypoints_2d'].numpy(), g['pred_keypoints_2d'], rtol=0, atol=2e-5)
    # dict layout of HAMER.forward_step (hamer.py:124-156)
    assert set(params) == {'global_orient', 'hand_pose', 'betas', 'trans'}
    assert out['pred_vertices'].shape == (2, 778, 3) and out['pred_keypoints_3d'].shape == (2, 21, 3)
    assert out['pred_keypoints_2d'].shape == (2, 21, 2) and out['focal_length'].shape == (2, 2)
    assert torch.all(out['focal_length'] == 5000.0)


def test_rot6d_is_a_rotation():
    from oracle.head import rot6d_to_rotmat
    x = torch.randn(64, 6, generator=torch.Generator().manual_seed(3))
    R = rot6d_to_rotmat(x)
    eye = torch.eye(3).expand(64, 3, 3)
    assert torch.allclose(R.transpose(1, 2) @ R, eye, atol=1e-5)
    assert torch.allclose(torch.linalg.det(R), torch.ones(64), atol=1e-5)
    # identity in the 6D representation used by default_mean_params
    assert torch.allclose(rot6d_to_rotmat(torch.tensor([[1., 0, 0, 0, 1, 0]])), torch.eye(3)[None])
