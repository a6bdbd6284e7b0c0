"""Denoiser networks of the hot path (same dotted paths as the reference's `models` package)."""
from .ema import EMA
from .unet import UNet
from .unet_categorial_adagn import UNetCategorialAdaGN
