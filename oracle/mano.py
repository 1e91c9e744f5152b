"""Oracle: MANO layer (fp32, CPU).  TEST INFRASTRUCTURE ONLY.

The reference calls the third-party ``smplx==0.1.28`` (pinned in setup.py:15 / hamer/setup.py:15;
source NOT under /root/reference) through hamer/hamer/models/mano_wrapper.py:12-44:
``smplx.MANOLayer.forward(global_orient, hand_pose, betas, pose2rot=False)`` -> ``smplx.lbs.lbs``.
This file restates that published algorithm (``lbs``, ``blend_shapes``, ``vertices2joints``,
``batch_rigid_transform``, ``transform_mat``) plus the wrapper's fingertip append / OpenPose
re-ordering (mano_wrapper.py:24-39).  Call sites anchoring the semantics: hamer.py:51-52,143;
infer.py:1387-1392.  Cross-checked against the reference's own independent implementation
rootnet/KeypointFusion/manopth/manopth/manolayer.py:172-249 (tests/test_oracle_mano.py).

``mano`` arguments are dicts of numpy/torch arrays:
  v_template (778,3)  shapedirs (778,3,10)  posedirs (135,2334)  J_regressor (16,778)
  lbs_weights (778,16)  parents (16,) int  [tip_ids (5,), joint_map (21,)]
"""
import numpy as np
import torch

# smplx.vertex_ids.vertex_ids['mano'] (thumb, index, middle, ring, pinky) - third-party constant
TIP_IDS = (744, 320, 443, 554, 671)
# mano_wrapper.py:24
MANO_TO_OPENPOSE = (0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20)
# MANO kinematic tree (kintree_table[0] with the root set to -1, as smplx does)
PARENTS = (-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14)


def _t(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).to(dtype)


def lbs(betas, rot_mats, mano):
    """smplx.lbs.lbs with pose2rot=False.  betas (B,10), rot_mats (B,16,3,3) ->
    vertices (B,778,3), posed joints (B,16,3)."""
    dtype, dev = betas.dtype, betas.device
    v_template = _t(mano['v_template'], dtype).to(dev)
    shapedirs = _t(mano['shapedirs'], dtype).to(dev)
    posedirs = _t(mano['posedirs'], dtype).to(dev)
    J_regressor = _t(mano['J_regressor'], dtype).to(dev)
    lbs_weights = _t(mano['lbs_weights'], dtype).to(dev)
    parents = [int(p) for p in mano.get('parents', PARENTS)]
    B = betas.shape[0]
    nj = J_regressor.shape[0]

    # blend_shapes: v_shaped = v_template + einsum('bl,mkl->bmk')
    v_shaped = v_template[None] + torch.einsum('bl,mkl->bmk', betas, shapedirs)
    # vertices2joints: einsum('bik,ji->bjk')
    J = torch.einsum('bik,ji->bjk', v_shaped, J_regressor)
    # pose blend shapes from (R - I) of the non-root joints
    ident = torch.eye(3, dtype=dtype, device=dev)
    pose_feature = (rot_mats[:, 1:] - ident).reshape(B, -1)
    v_posed = v_shaped + torch.matmul(pose_feature, posedirs).view(B, -1, 3)

    # batch_rigid_transform
    joints = J.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] = rel[:, 1:] - joints[:, parents[1:]]
    tm = torch.zeros(B, nj, 4, 4, dtype=dtype, device=dev)
    tm[:, :, :3, :3] = rot_mats
    tm[:, :, :3, 3:] = rel
    tm[:, :, 3, 3] = 1.0
    chain = [tm[:, 0]]
    for i in range(1, nj):
        chain.append(torch.matmul(chain[parents[i]], tm[:, i]))
    transforms = torch.stack(chain, dim=1)
    posed_joints = transforms[:, :, :3, 3]
    jh = torch.cat([joints, torch.zeros(B, nj, 1, 1, dtype=dtype, device=dev)], dim=2)
    shift = torch.matmul(transforms, jh)                     # (B,nj,4,1)
    A = transforms.clone()
    A[:, :, :, 3:] = A[:, :, :, 3:] - shift                  # rel_transforms

    # skinning
    T = torch.matmul(lbs_weights[None].expand(B, -1, -1), A.view(B, nj, 16)).view(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=dtype, device=dev)], dim=2)
    verts = torch.matmul(T, vh.unsqueeze(-1))[:, :, :3, 0]
    return verts, posed_joints


def mano_forward(global_orient, hand_pose, betas, mano):
    """mano_wrapper.py:32-44: MANOLayer LBS, append 5 fingertip vertices, OpenPose order.
    Returns vertices (B,778,3), joints (B,21,3)."""
    B = betas.shape[0]
    rot = torch.cat([global_orient.reshape(B, 1, 3, 3), hand_pose.reshape(B, -1, 3, 3)], dim=1)
    verts, j16 = lbs(betas, rot, mano)
    tips = [int(t) for t in mano.get('tip_ids', TIP_IDS)]
    jmap = [int(j) for j in mano.get('joint_map', MANO_TO_OPENPOSE)]
    joints = torch.cat([j16, verts[:, tips]], dim=1)[:, jmap]
    return verts, joints
