"""Minimal `omegaconf` stand-in: the subset the reference's scripts use (scripts/sample_uncond.py:115-116,151,
scripts/train_ddpm.py:50-64, utils/misc.py:8,72-73): OmegaConf.load / create / merge / from_dotlist / to_container /
to_yaml and DictConfig / ListConfig with attribute access and `.get`."""
import copy

import yaml

__all__ = ['OmegaConf', 'DictConfig', 'ListConfig']


def _wrap(v):
    if isinstance(v, dict) and not isinstance(v, DictConfig):
        return DictConfig(v)
    if isinstance(v, (list, tuple)) and not isinstance(v, ListConfig):
        return ListConfig(v)
    return v


class DictConfig(dict):
    def __init__(self, content=None):
        super().__init__()
        for k, v in (content or {}).items():
            dict.__setitem__(self, k, _wrap(v))

    def __getattr__(self, key):
        if key.startswith('__'):
            raise AttributeError(key)
        try:
            return self[key]
        except KeyError:
            raise AttributeError(f"Missing key {key}") from None

    def __setattr__(self, key, value):
        self[key] = value

    def __setitem__(self, key, value):
        dict.__setitem__(self, key, _wrap(value))

    def __deepcopy__(self, memo):
        return DictConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})


class ListConfig(list):
    def __init__(self, content=()):
        super().__init__(_wrap(v) for v in content)


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    return v


def _merge_into(dst: DictConfig, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge_into(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)


class OmegaConf:
    @staticmethod
    def create(obj=None):
        if isinstance(obj, str):
            obj = yaml.safe_load(obj)
        return _wrap(obj if obj is not None else {})

    @staticmethod
    def load(path):
        with open(path) as f:
            return _wrap(yaml.safe_load(f) or {})

    @staticmethod
    def from_dotlist(dotlist):
        """['a.b=1', 'c=[1,2]'] -> nested config; values are parsed as YAML scalars like omegaconf does."""
        root = DictConfig()
        for item in dotlist:
            key, _, val = item.partition('=')
            key = key.lstrip('-')
            node = root
            parts = key.split('.')
            for p in parts[:-1]:
                if not isinstance(node.get(p), dict):
                    node[p] = DictConfig()
                node = node[p]
            node[parts[-1]] = yaml.safe_load(val) if val != '' else None
        return root

    @staticmethod
    def merge(*configs):
        out = DictConfig()
        for c in configs:
            _merge_into(out, c if isinstance(c, dict) else dict(c))
        return out

    @staticmethod
    def to_container(conf, resolve=True, **_):
        return _plain(conf)

    @staticmethod
    def to_yaml(conf, **_):
        return yaml.safe_dump(_plain(conf), sort_keys=False)

    @staticmethod
    def save(config, f):
        text = OmegaConf.to_yaml(config)
        if hasattr(f, 'write'):
            f.write(text)
        else:
            with open(f, 'w') as fh:
                fh.write(text)
