"""pytorch_lightning stand-in: diffusion/model/swinir.py only subclasses LightningModule; at inference it is an nn.Module."""
import torch.nn as nn


class LightningModule(nn.Module):
    pass
