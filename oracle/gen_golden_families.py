"""Pins oracle/adm_ref.py against the live reference (ADM UNetModel, UNetCombined dispatch, pesser Model) and freezes
small golden forwards in tests/golden/family_forward.pt.  Build-container only (imports /root/reference):
    python oracle/gen_golden_families.py
Weights are not stored: both sides re-create them from torch.manual_seed(2022) + default initialisers (+ the
N(0, 0.02) re-draw of ADM's zero-initialised tensors, oracle.adm_ref.randomize_zero_params)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.gen_golden import GOLD, import_reference  # noqa: E402

ADM_CFGS = {
    # every structural feature of the 256x256 config at toy width: scale-shift, up/down ResBlocks, head_dim 64,
    # attention at two resolutions, class conditioning, learned-variance output channels
    'adm_tiny_cond': dict(image_size=32, in_channels=3, model_channels=64, out_channels=6, num_res_blocks=1,
                          attention_resolutions=[2, 4], dropout=0.0, channel_mult=[1, 2, 2], conv_resample=True,
                          dims=2, num_classes=10, use_checkpoint=False, use_fp16=False, num_heads=4,
                          num_head_channels=64, num_heads_upsample=-1, use_scale_shift_norm=True,
                          resblock_updown=True, use_new_attention_order=False),
    # the other code paths: additive embedding, strided-conv / nearest+conv resampling, fixed head count,
    # new attention order, unconditional
    'adm_tiny_plain': dict(image_size=32, in_channels=3, model_channels=64, out_channels=3, num_res_blocks=1,
                           attention_resolutions=[2], dropout=0.0, channel_mult=[1, 2], conv_resample=True, dims=2,
                           num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=2, num_head_channels=-1,
                           num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                           use_new_attention_order=True),
}
PESSER_CFGS = {
    'pesser_tiny': dict(resolution=32, in_channels=3, out_ch=3, ch=64, ch_mult=[1, 2, 2], num_res_blocks=1,
                        attn_resolutions=[16], dropout=0.0, resamp_with_conv=True),
}


def main():
    from oracle.adm_ref import FamilyRef, randomize_zero_params
    import_reference()
    from models.adm.unet import UNetModel
    from models.adm.unet_combined import UNetCombined
    from models.pesser.model import Model
    torch.set_grad_enabled(False)
    out = {}
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 3, 32, 32, generator=g)
    t = torch.tensor([37, 911])
    y = torch.tensor([3, 7])
    for name, cfg in ADM_CFGS.items():
        torch.manual_seed(2022)
        ref = UNetModel(**cfg).eval()
        ref.load_state_dict(randomize_zero_params(ref.state_dict()))
        yy = y if cfg['num_classes'] is not None else None
        want = ref(x, t, yy)
        got = FamilyRef('adm', ref.state_dict(), cfg)(x, t, yy)
        err = (got - want).abs().max().item()
        assert err <= 1e-5, (name, err)
        print(f'{name}: oracle vs reference max abs err {err:.2e} (|out| max {want.abs().max():.3f})')
        out[name] = dict(family='adm', cfg=cfg, seed=2022, x=x, t=t, y=yy, out=want.clone(),
                         param_sum=float(sum(p.double().sum() for p in ref.parameters())),
                         keys=[(k, tuple(v.shape)) for k, v in ref.state_dict().items()])
    # UNetCombined: two weight sets, label decides (models/adm/unet_combined.py:23-25)
    cfg = ADM_CFGS['adm_tiny_cond']
    torch.manual_seed(2022)
    comb = UNetCombined(**cfg).eval()
    comb.load_state_dict(randomize_zero_params(comb.state_dict()))
    sd = comb.state_dict()
    for tag, yy, sub, ncls in (('cond', y, 'unet_cond.', 10), ('uncond', None, 'unet_uncond.', None)):
        want = comb(x, t, yy)
        sub_sd = {k[len(sub):]: v for k, v in sd.items() if k.startswith(sub)}
        got = FamilyRef('adm', sub_sd, dict(cfg, num_classes=ncls))(x, t, yy)
        err = (got - want).abs().max().item()
        assert err <= 1e-5, ('combined', tag, err)
        print(f'adm combined [{tag}]: oracle vs reference max abs err {err:.2e}')
        out[f'adm_combined_{tag}'] = dict(family='adm_combined', cfg=cfg, seed=2022, x=x, t=t, y=yy, out=want.clone(),
                                          keys=[(k, tuple(v.shape)) for k, v in sd.items()])
    for name, cfg in PESSER_CFGS.items():
        torch.manual_seed(2022)
        ref = Model(**cfg).eval()
        want = ref(x, t)
        got = FamilyRef('pesser', ref.state_dict(), cfg)(x, t)
        err = (got - want).abs().max().item()
        assert err <= 1e-5, (name, err)
        print(f'{name}: oracle vs reference max abs err {err:.2e} (|out| max {want.abs().max():.3f})')
        out[name] = dict(family='pesser', cfg=cfg, seed=2022, x=x, t=t, y=None, out=want.clone(),
                         param_sum=float(sum(p.double().sum() for p in ref.parameters())),
                         keys=[(k, tuple(v.shape)) for k, v in ref.state_dict().items()])
    torch.save(out, os.path.join(GOLD, 'family_forward.pt'))
    print('written', os.path.join(GOLD, 'family_forward.pt'))


if __name__ == '__main__':
    main()
