"""diffusers stand-in: only so that diffusion/model/nets/transformer_controlnet.py imports; never executed."""
import torch.nn as nn


class Transformer2DModel(nn.Module):
    pass
