"""Oracle: MANO transformer-decoder head + rot6d (fp32, CPU).  TEST INFRASTRUCTURE ONLY.

Restates
  hamer/hamer/models/heads/mano_head.py:20-115            MANOTransformerDecoderHead
  hamer/hamer/models/components/pose_transformer.py:27-37 PreNorm, :40-52 FeedForward,
      :55-86 Attention, :89-124 CrossAttention, :160-201 TransformerCrossAttn,
      :301-357 TransformerDecoder
  hamer/hamer/models/components/t_cond_mlp.py:48-59       normalization_layer ("layer" -> LayerNorm)
  hamer/hamer/utils/geometry.py:47-70                     rot6d_to_rotmat
with the config of configs_hydra/experiment/hamer_vit_transformer.yaml:31-43
(depth 6, heads 8, dim_head 64, mlp_dim 1024, context_dim 1280, dim 1024, zero input token,
IEF_ITERS 1).  Parameter names / construction order match the reference module tree.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

DIM = 1024
DEPTH = 6
HEADS = 8
DIM_HEAD = 64
MLP_DIM = 1024
CONTEXT_DIM = 1280
NPOSE = 6 * 16


class _PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)  # eps 1e-5
        self.fn = fn

    def forward(self, x, **kw):
        return self.fn(self.norm(x), **kw)


class _FeedForward(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        # indices 0 and 3 as in the reference Sequential (Linear, GELU, Dropout, Linear, Dropout)
        self.net = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Identity(),
                                 nn.Linear(hidden, dim), nn.Identity())

    def forward(self, x):
        return self.net(x)


def _split_heads(t, h):
    b, n, _ = t.shape
    return t.reshape(b, n, h, -1).permute(0, 2, 1, 3)


def _merge_heads(t):
    b, h, n, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(b, n, h * d)


class _SelfAttention(nn.Module):
    def __init__(self, dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Identity())

    def forward(self, x):  # pose_transformer.py:75-86
        q, k, v = (_split_heads(t, self.heads) for t in self.to_qkv(x).chunk(3, dim=-1))
        attn = (torch.matmul(q, k.transpose(-1, -2)) * self.scale).softmax(dim=-1)
        return self.to_out(_merge_heads(torch.matmul(attn, v)))


class _CrossAttention(nn.Module):
    def __init__(self, dim, context_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_kv = nn.Linear(context_dim, inner * 2, bias=False)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Identity())

    def forward(self, x, context=None):  # pose_transformer.py:111-124
        k, v = self.to_kv(context).chunk(2, dim=-1)
        q = self.to_q(x)
        q, k, v = (_split_heads(t, self.heads) for t in (q, k, v))
        attn = (torch.matmul(q, k.transpose(-1, -2)) * self.scale).softmax(dim=-1)
        return self.to_out(_merge_heads(torch.matmul(attn, v)))


class _CrossAttnStack(nn.Module):
    def __init__(self):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(DEPTH):
            sa = _SelfAttention(DIM, HEADS, DIM_HEAD)
            ca = _CrossAttention(DIM, CONTEXT_DIM, HEADS, DIM_HEAD)
            ff = _FeedForward(DIM, MLP_DIM)
            self.layers.append(nn.ModuleList([_PreNorm(DIM, sa), _PreNorm(DIM, ca), _PreNorm(DIM, ff)]))

    def forward(self, x, context):  # pose_transformer.py:191-201
        for sa, ca, ff in self.layers:
            x = sa(x) + x
            x = ca(x, context=context) + x
            x = ff(x) + x
        return x


class _Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.to_token_embedding = nn.Linear(1, DIM)
        self.pos_embedding = nn.Parameter(torch.randn(1, 1, DIM))
        self.transformer = _CrossAttnStack()

    def forward(self, inp, context):  # pose_transformer.py:349-357
        x = self.to_token_embedding(inp)
        x = x + self.pos_embedding[:, : x.shape[1]]
        return self.transformer(x, context)


def rot6d_to_rotmat(x):
    """geometry.py:47-70."""
    x = x.reshape(-1, 2, 3).permute(0, 2, 1).contiguous()
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - torch.einsum('bi,bi->b', b1, a2).unsqueeze(-1) * b1)
    b3 = torch.linalg.cross(b1, b2)
    return torch.stack((b1, b2, b3), dim=-1)


class OManoHead(nn.Module):
    """mano_head.py:20-115.  ``mean_params`` replaces the hard-coded mano_mean_params.npz (:53)."""

    def __init__(self, mean_params):
        super().__init__()
        self.transformer = _Decoder()
        self.decpose = nn.Linear(DIM, NPOSE)
        self.decshape = nn.Linear(DIM, 10)
        self.deccam = nn.Linear(DIM, 3)
        self.register_buffer('init_hand_pose', torch.as_tensor(mean_params['pose'], dtype=torch.float32).reshape(1, NPOSE))
        self.register_buffer('init_betas', torch.as_tensor(mean_params['shape'], dtype=torch.float32).reshape(1, 10))
        self.register_buffer('init_cam', torch.as_tensor(mean_params['cam'], dtype=torch.float32).reshape(1, 3))

    def forward(self, feats):
        B = feats.shape[0]
        token = torch.zeros(B, 1, 1, dtype=feats.dtype, device=feats.device)
        token_out = self.transformer(token, feats).squeeze(1)
        pose6d = self.decpose(token_out) + self.init_hand_pose
        betas = self.decshape(token_out) + self.init_betas
        cam = self.deccam(token_out) + self.init_cam
        rot = rot6d_to_rotmat(pose6d).view(B, 16, 3, 3)
        return {'global_orient': rot[:, [0]], 'hand_pose': rot[:, 1:], 'betas': betas}, cam


def default_mean_params():
    """Stand-in for mano_mean_params.npz (absent from the reference tree):
    identity rotations in 6D, zero betas, cam = [0.9, 0, 0]."""
    import numpy as np
    pose = np.tile(np.array([1, 0, 0, 0, 1, 0], dtype=np.float32), 16)  # a1=e_x, a2=e_y
    return {'pose': pose, 'shape': np.zeros(10, np.float32), 'cam': np.array([0.9, 0.0, 0.0], np.float32)}
