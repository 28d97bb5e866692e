"""timm.models.layers stand-in (timm 0.6.12): DropPath is the identity at inference (drop_prob 0 / eval)."""
import torch.nn as nn


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not (self.training and self.drop_prob > 0), "shim supports inference only"
        return x


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)
