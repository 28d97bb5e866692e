"""instarevive_b200: B200-native (sm_100a) one-step restoration forward of InstaRevive behind the reference's operator
surface. Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic is hand-written CUDA in
csrc/ behind the C ABI of include/instarevive_b200.h."""
from . import convert, dpm_solver  # noqa: F401
from .dpm_solver import DPMS  # noqa: F401
from .generate import DDPMSchedulerLite, eps_to_mu, forward_model, generate_sample_1step  # noqa: F401
from .nets import ControlPixArtMSHalf, PixArtMS, PixArtMS_XL_2, PixArtMSBlock  # noqa: F401
from .pipeline import _sliding_windows, process, restore_latents  # noqa: F401
from .swinir import SwinIR  # noqa: F401
from .transformer_controlnet import ControlTransformerHalf, Transformer2DModel, Transformer2DModelOutput  # noqa: F401
from .vae import AutoencoderKL, AutoencoderKLDecoder, DiagonalGaussianDistribution  # noqa: F401

__version__ = "0.1.0"
