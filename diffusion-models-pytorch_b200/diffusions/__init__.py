"""Diffusion processes of the hot path (same dotted paths as the reference's `diffusions` package)."""
from .schedule import get_beta_schedule, get_respaced_seq

from .ddpm import DDPM, DDPMCFG
from .ddim import DDIM, DDIMCFG
from .euler import EulerSampler
from .heun import HeunSampler
