"""Minimal ``mmcv.Config.fromfile``: python config files with ``_base_`` inheritance
(str or list, paths relative to the including file, duplicate keys across bases are an error),
recursive dict merge of the child into the base, and ``_delete_=True`` replacement
(reference configs/mv2d/exp/*.py use all three)."""
import copy
import os
import types


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _to_cfgdict(x):
    if isinstance(x, dict):
        return ConfigDict({k: _to_cfgdict(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_to_cfgdict(v) for v in x]
    if isinstance(x, tuple):
        return tuple(_to_cfgdict(v) for v in x)
    return x


def _merge(child, base):
    out = copy.deepcopy(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get('_delete_', False):
            out[k] = _merge(v, out[k])
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != '_delete_'}
            out[k] = copy.deepcopy(v)
    return out


def _load(path):
    scope = {}
    with open(path) as f:
        exec(compile(f.read(), path, 'exec'), scope)
    cfg = {k: v for k, v in scope.items()
           if not k.startswith('__') and not isinstance(v, types.ModuleType) and not callable(v)}
    bases = cfg.pop('_base_', [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        bcfg = _load(os.path.normpath(os.path.join(os.path.dirname(path), b)))
        dup = set(merged) & set(bcfg)
        if dup:
            raise KeyError(f'duplicate keys in _base_ files of {path}: {sorted(dup)}')
        merged.update(bcfg)
    return _merge(cfg, merged)


class Config:
    @staticmethod
    def fromfile(path):
        return _to_cfgdict(_load(os.path.abspath(path)))
