"""The `omegaconf` / `accelerate` stand-ins (diffusion-models-pytorch_b200/shims): the call patterns of the
reference's scripts (scripts/sample_uncond.py:115-131, scripts/train_ddpm.py:36-101, utils/misc.py:68-78).  CPU only."""
import importlib
import os
import sys

import pytest
import torch

SHIMS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'diffusion-models-pytorch_b200', 'shims')


@pytest.fixture()
def shims(monkeypatch):
    monkeypatch.syspath_prepend(SHIMS)
    for name in [m for m in sys.modules if m == 'omegaconf' or m.startswith('accelerate')]:
        monkeypatch.delitem(sys.modules, name)
    yield
    for name in [m for m in sys.modules if m == 'omegaconf' or m.startswith('accelerate')]:
        sys.modules.pop(name, None)


def test_omegaconf_script_flow(shims, tmp_path):
    from omegaconf import DictConfig, OmegaConf
    cfg = tmp_path / 'c.yaml'
    cfg.write_text('seed: 2022\nmodel:\n  target: models.unet.UNet\n  params:\n    dim: 128\n    dim_mults: [1, 2, 2, 2]\n'
                   'diffusion:\n  target: diffusions.ddpm.DDPM\n  params:\n    total_steps: 1000\n')
    conf = OmegaConf.load(str(cfg))
    conf = OmegaConf.merge(conf, OmegaConf.from_dotlist(['--diffusion.params.total_steps=200', 'train.batch_size=64',
                                                         'model.params.use_attn=[false,true,false,false]']))
    assert isinstance(conf, DictConfig) and isinstance(conf.model, DictConfig)
    assert conf.diffusion.params.total_steps == 200 and conf.train.batch_size == 64 and conf.seed == 2022
    assert conf.model.params.use_attn == [False, True, False, False] and conf.model.params.dim == 128
    assert conf.get('missing', 7) == 7
    with pytest.raises(AttributeError):
        conf.nope
    plain = OmegaConf.to_container(conf.model)
    assert type(plain) is dict and type(plain['params']) is dict and type(plain['params']['dim_mults']) is list
    assert 'total_steps: 200' in OmegaConf.to_yaml(conf)
    # utils/misc.py:68-78 instantiate_from_config against the product package
    module, cls = plain['target'].rsplit('.', 1)
    klass = getattr(importlib.import_module(module), cls)
    m = klass(**{**plain['params'], 'use_attn': [False, True, False, False]})
    assert sum(p.numel() for p in m.parameters()) == 35746307        # CIFAR-10 UNet parameter count


def test_accelerate_single_process(shims, tmp_path):
    import accelerate
    from accelerate.utils import set_seed
    acc = accelerate.Accelerator(kwargs_handlers=[accelerate.DistributedDataParallelKwargs(find_unused_parameters=True)])
    assert acc.num_processes == 1 and acc.process_index == 0 and acc.is_main_process
    assert acc.distributed_type == accelerate.utils.DistributedType.NO and acc.mixed_precision == 'no'
    lin = torch.nn.Linear(4, 2)
    opt = torch.optim.SGD(lin.parameters(), lr=0.1)
    model, opt2, loader = acc.prepare(lin, opt, [1, 2, 3])
    assert acc.unwrap_model(model) is lin and opt2 is opt and loader == [1, 2, 3]
    x = torch.randn(3, 4)
    with acc.no_sync(model):
        acc.backward(model(x).sum())
    assert float(acc.clip_grad_norm_(model.parameters(), 1.0)) > 0
    assert torch.equal(acc.gather(x), x) and torch.equal(acc.gather_for_metrics(x), x)
    acc.wait_for_everyone()
    acc.save({'a': 1}, str(tmp_path / 'x.pt'))
    assert torch.load(str(tmp_path / 'x.pt'))['a'] == 1
    assert acc.on_main_process(lambda: 5)() == 5
    assert set_seed(10, device_specific=True) == 10 + int(os.environ.get('RANK', '0'))
