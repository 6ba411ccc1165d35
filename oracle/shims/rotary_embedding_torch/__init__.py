"""Restatement of `rotary-embedding-torch==0.8.4` `RotaryEmbedding` — TEST INFRASTRUCTURE ONLY.

The package is a pinned third-party dependency of the reference (environment.yaml:88) whose source is
NOT under /root/reference. Only `RotaryEmbedding(dim).rotate_queries_or_keys(t)` is used
(video_diffusion_pytorch_conv3d.py:380, :320-321).  Published algorithm ("lang" frequencies):

    freqs_i = theta^(-2i/dim), i = 0..dim/2-1, theta = 10000
    angle[p, 2i] = angle[p, 2i+1] = p * freqs_i          (each frequency repeated for an interleaved pair)
    out = t * cos(angle) + rotate_half(t) * sin(angle)
    rotate_half: (x_{2i}, x_{2i+1}) -> (-x_{2i+1}, x_{2i})

positions p = 0..seq_len-1 along dim -2.  PARITY UNPINNED: there is no copy of the real package to compare
against in this image.
"""
import torch
from torch import nn


def rotate_half(x):
    x = x.reshape(*x.shape[:-1], x.shape[-1] // 2, 2)
    x1, x2 = x.unbind(dim=-1)
    x = torch.stack((-x2, x1), dim=-1)
    return x.reshape(*x.shape[:-2], -1)


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
        # 0.8.x registers `freqs` as a non-trainable parameter; the reference checkpoints carry it.
        self.freqs = nn.Parameter(freqs, requires_grad=False)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        seq_len = t.shape[seq_dim]
        pos = torch.arange(seq_len, device=t.device, dtype=self.freqs.dtype)
        ang = pos[:, None] * self.freqs[None, :]
        ang = ang.repeat_interleave(2, dim=-1)  # [n, dim]
        rot_dim = ang.shape[-1]
        t_rot, t_pass = t[..., :rot_dim], t[..., rot_dim:]
        t_rot = t_rot * ang.cos() + rotate_half(t_rot) * ang.sin()
        return torch.cat((t_rot, t_pass), dim=-1)
