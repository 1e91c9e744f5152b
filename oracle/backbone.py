"""Oracle: ViT-H/16 backbone (fp32, CPU).  TEST INFRASTRUCTURE ONLY.

Restates hamer/hamer/models/backbones/vit.py of the reference:
  vit()            :12-24   (256x192, patch 16, dim 1280, depth 32, heads 16, mlp x4, qkv bias)
  Mlp              :72-87
  Attention        :89-126
  Block            :128-151
  PatchEmbed       :154-176 (Conv2d k16 s16 p2)
  ViT.forward_features :320-339

Parameter names and construction order match the reference module tree, so the same
``torch.manual_seed`` yields the same random-init weights as the reference's own constructor
(checked in tests/golden/make_golden.py).
"""
import math

import torch
import torch.nn as nn

EMBED_DIM = 1280
DEPTH = 32
NUM_HEADS = 16
PATCH = 16
IMG_H, IMG_W = 256, 192
NUM_TOKENS = (IMG_H // PATCH) * (IMG_W // PATCH)  # 192


class OMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()  # exact erf GELU, vit.py:73
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):  # vit.py:82-87 (dropout p=0)
        return self.fc2(self.act(self.fc1(x)))


class OAttention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):  # vit.py:110-126
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        q = q * self.scale
        attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, N, -1)
        return self.proj(x)


class OBlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, eps=1e-6):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)  # eps 1e-6: vit.py:222
        self.attn = OAttention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = OMlp(dim, int(dim * mlp_ratio))

    def forward(self, x):  # vit.py:148-151 (DropPath is identity in eval)
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


class OPatchEmbed(nn.Module):
    def __init__(self, dim):
        super().__init__()
        # padding = 4 + 2*(ratio//2 - 1) with ratio=1 -> 2  (vit.py:168)
        self.proj = nn.Conv2d(3, dim, kernel_size=PATCH, stride=PATCH, padding=2)

    def forward(self, x):  # vit.py:170-176
        x = self.proj(x)
        return x.flatten(2).transpose(1, 2)


class OViT(nn.Module):
    def __init__(self, embed_dim=EMBED_DIM, depth=DEPTH, num_heads=NUM_HEADS):
        super().__init__()
        self.patch_embed = OPatchEmbed(embed_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, NUM_TOKENS + 1, embed_dim))
        self.blocks = nn.ModuleList([OBlock(embed_dim, num_heads) for _ in range(depth)])
        self.last_norm = nn.LayerNorm(embed_dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)  # vit.py:254-255

    def forward(self, x):  # vit.py:320-339
        x = self.patch_embed(x)
        x = x + self.pos_embed[:, 1:] + self.pos_embed[:, :1]
        for blk in self.blocks:
            x = blk(x)
        return self.last_norm(x)
