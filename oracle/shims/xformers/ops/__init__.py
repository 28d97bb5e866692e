"""xformers.ops stand-in (xformers 0.0.19 semantics): softmax(q k^T * K^-1/2 + bias) v on (B, M, H, K) tensors."""
import torch
import torch.nn.functional as F

from . import fmha  # noqa: F401
from .fmha import BlockDiagonalMask


def memory_efficient_attention(query, key, value, attn_bias=None, p=0.0, scale=None):
    assert p == 0.0, "shim supports inference only"
    if isinstance(attn_bias, BlockDiagonalMask):
        assert query.shape[0] == 1
        outs, qs, ks = [], 0, 0
        for ql, kl in zip(attn_bias.q_seqlen, attn_bias.kv_seqlen):
            q = query[:, qs:qs + ql].permute(0, 2, 1, 3)
            k = key[:, ks:ks + kl].permute(0, 2, 1, 3)
            v = value[:, ks:ks + kl].permute(0, 2, 1, 3)
            outs.append(F.scaled_dot_product_attention(q, k, v, scale=scale).permute(0, 2, 1, 3))
            qs += ql
            ks += kl
        return torch.cat(outs, dim=1)
    q, k, v = (t.permute(0, 2, 1, 3) for t in (query, key, value))
    mask = None
    if attn_bias is not None:
        B, H = query.shape[0], query.shape[2]
        mask = attn_bias.view(B, H, attn_bias.shape[-2], attn_bias.shape[-1])
    return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, scale=scale).permute(0, 2, 1, 3)
