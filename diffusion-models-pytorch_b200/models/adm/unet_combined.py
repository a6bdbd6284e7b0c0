"""Conditional + unconditional ADM UNets behind one module (reference models/adm/unet_combined.py:6-32): the label
decides which of the two full weight sets runs, which is what classifier-free guidance over the OpenAI
checkpoints needs (they were trained as separate networks)."""
import torch
import torch.nn as nn

from .unet import UNetModel


class UNetCombined(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        assert kwargs.get('num_classes') is not None
        self.unet_cond = UNetModel(*args, **kwargs)
        self.unet_uncond = UNetModel(*args, **dict(kwargs, num_classes=None))
        self.in_channels, self.out_channels = self.unet_cond.in_channels, self.unet_cond.out_channels

    def forward(self, x, timesteps, y=None, out=None):
        if y is None:
            return self.unet_uncond(x, timesteps, None, out=out)
        return self.unet_cond(x, timesteps, y, out=out)

    def make_sampling_runner(self, diffuser):
        from models.runner import SamplingRunner
        import weakref
        cache = self.__dict__.setdefault('_runners', weakref.WeakKeyDictionary())
        r = cache.get(diffuser)
        if r is None:
            r = cache[diffuser] = SamplingRunner(self, diffuser)
        return r

    @property
    def engine(self):
        return _PairEngine(self.unet_cond.engine, self.unet_uncond.engine)

    def combine_weights(self, cond_path, uncond_path, save_path):
        self.unet_cond.load_state_dict(torch.load(cond_path, map_location='cpu'))
        self.unet_uncond.load_state_dict(torch.load(uncond_path, map_location='cpu'))
        torch.save(self.state_dict(), save_path)


class _PairEngine:
    """What models/runner.py needs from an engine (weight-version signature), over both sub-networks."""

    def __init__(self, a, b):
        self.a, self.b = a, b

    def refresh(self):
        self.a.refresh()
        self.b.refresh()

    @property
    def _sig(self):
        return (self.a._sig, self.b._sig)
