import torch.nn as nn


class BasicTransformerBlock(nn.Module):
    pass
