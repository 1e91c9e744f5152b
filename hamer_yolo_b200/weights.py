"""State-dict layout of the reference checkpoint for the hot path, and a fast random-init generator.

Key layout = the reference module tree:
  backbone.*   hamer/hamer/models/backbones/vit.py:211-262
  mano_head.*  hamer/hamer/models/heads/mano_head.py:24-59, components/pose_transformer.py:301-347

``random_state_dict`` draws from the same distributions as the reference's constructors (PyTorch
default Linear/Conv init = U(+-1/sqrt(fan_in)), LayerNorm (1,0), pos_embed trunc_normal(std .02),
decoder pos_embedding N(0,1); vit() never calls init_weights, vit.py:12-24) but with its own fast
generator - it is for benchmarking, not for parity (parity tests take the oracle's tensors).
"""
import math

import numpy as np
import torch


def state_dict_shapes():
    s = {'backbone.patch_embed.proj.weight': (1280, 3, 16, 16), 'backbone.patch_embed.proj.bias': (1280,),
         'backbone.pos_embed': (1, 193, 1280),
         'backbone.last_norm.weight': (1280,), 'backbone.last_norm.bias': (1280,)}
    for i in range(32):
        p = f'backbone.blocks.{i}.'
        s.update({p + 'norm1.weight': (1280,), p + 'norm1.bias': (1280,),
                  p + 'attn.qkv.weight': (3840, 1280), p + 'attn.qkv.bias': (3840,),
                  p + 'attn.proj.weight': (1280, 1280), p + 'attn.proj.bias': (1280,),
                  p + 'norm2.weight': (1280,), p + 'norm2.bias': (1280,),
                  p + 'mlp.fc1.weight': (5120, 1280), p + 'mlp.fc1.bias': (5120,),
                  p + 'mlp.fc2.weight': (1280, 5120), p + 'mlp.fc2.bias': (1280,)})
    t = 'mano_head.transformer.'
    s.update({t + 'pos_embedding': (1, 1, 1024), t + 'to_token_embedding.weight': (1024, 1),
              t + 'to_token_embedding.bias': (1024,)})
    for l in range(6):
        p = f'{t}transformer.layers.{l}.'
        for sub in range(3):
            s.update({f'{p}{sub}.norm.weight': (1024,), f'{p}{sub}.norm.bias': (1024,)})
        s.update({p + '0.fn.to_qkv.weight': (1536, 1024),
                  p + '0.fn.to_out.0.weight': (1024, 512), p + '0.fn.to_out.0.bias': (1024,),
                  p + '1.fn.to_kv.weight': (1024, 1280), p + '1.fn.to_q.weight': (512, 1024),
                  p + '1.fn.to_out.0.weight': (1024, 512), p + '1.fn.to_out.0.bias': (1024,),
                  p + '2.fn.net.0.weight': (1024, 1024), p + '2.fn.net.0.bias': (1024,),
                  p + '2.fn.net.3.weight': (1024, 1024), p + '2.fn.net.3.bias': (1024,)})
    s.update({'mano_head.decpose.weight': (96, 1024), 'mano_head.decpose.bias': (96,),
              'mano_head.decshape.weight': (10, 1024), 'mano_head.decshape.bias': (10,),
              'mano_head.deccam.weight': (3, 1024), 'mano_head.deccam.bias': (3,),
              'mano_head.init_hand_pose': (1, 96), 'mano_head.init_betas': (1, 10),
              'mano_head.init_cam': (1, 3)})
    return s


def default_mean_params():
    """Stand-in for mano_mean_params.npz (mano_head.py:53; absent from the reference tree):
    identity rotations in the 6D representation, zero betas, cam = [0.9, 0, 0]."""
    return {'pose': np.tile(np.array([1, 0, 0, 0, 1, 0], np.float32), 16),
            'shape': np.zeros(10, np.float32), 'cam': np.array([0.9, 0.0, 0.0], np.float32)}


def random_state_dict(seed=0, device='cpu', bf16_representable=True):
    g = torch.Generator(device=device).manual_seed(seed)
    mean = default_mean_params()
    sd = {}
    shapes = state_dict_shapes()
    for name, shape in shapes.items():
        if name.endswith('norm.weight') or name.endswith('norm1.weight') or name.endswith('norm2.weight'):
            t = torch.ones(shape, device=device)
        elif 'norm' in name and name.endswith('.bias'):
            t = torch.zeros(shape, device=device)
        elif name == 'backbone.pos_embed':
            t = torch.randn(shape, generator=g, device=device).clamp_(-2, 2) * 0.02
        elif name.endswith('pos_embedding'):
            t = torch.randn(shape, generator=g, device=device)
        elif name == 'mano_head.init_hand_pose':
            t = torch.from_numpy(mean['pose']).reshape(shape).to(device)
        elif name == 'mano_head.init_betas':
            t = torch.from_numpy(mean['shape']).reshape(shape).to(device)
        elif name == 'mano_head.init_cam':
            t = torch.from_numpy(mean['cam']).reshape(shape).to(device)
        else:
            wshape = shapes[name[:-4] + 'weight'] if name.endswith('.bias') else shape
            fan_in = int(np.prod(wshape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, device=device) * 2 - 1) * bound
        if bf16_representable and t.dim() >= 2:
            t = t.to(torch.bfloat16).float()
        sd[name] = t
    return sd
