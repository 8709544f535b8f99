"""mmcv-compatible ``Registry`` / ``build_from_cfg`` (mmcv itself is not a dependency).

Same semantics as the reference relies on: ``cfg['type']`` names a
registered class, the remaining keys are constructor kwargs.  The registry names mirror the
ones ``configs/mv2d/*`` resolve through (mmdet3d_plugin/__init__.py:10-19 side-effect
registration)."""
import copy


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f'{key} is already registered in {self.name}')
            self._modules[key] = cls
            return cls
        return _register(module) if module is not None else _register

    def get(self, key):
        return self._modules.get(key)

    def __contains__(self, key):
        return key in self._modules

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict) or 'type' not in cfg:
        raise TypeError(f'cfg must be a dict with a "type" key, got {cfg!r}')
    args = copy.copy(dict(cfg))
    for k, v in (default_args or {}).items():
        args.setdefault(k, v)
    t = args.pop('type')
    cls = registry.get(t) if isinstance(t, str) else t
    if cls is None:
        raise KeyError(f'{t} is not in the {registry.name} registry')
    return cls(**args)


DETECTORS = Registry('detector')
HEADS = Registry('head')
NECKS = Registry('neck')
TRANSFORMER = Registry('Transformer')
TRANSFORMER_LAYER = Registry('transformerLayer')
TRANSFORMER_LAYER_SEQUENCE = Registry('transformer-layers sequence')
ATTENTION = Registry('attention')
POSITIONAL_ENCODING = Registry('position encoding')
BBOX_CODERS = Registry('bbox_coder')
ROI_EXTRACTORS = Registry('roi_extractor')
LOSSES = Registry('loss')
BBOX_ASSIGNERS = Registry('bbox_assigner')
BBOX_SAMPLERS = Registry('bbox_sampler')
MATCH_COST = Registry('match_cost')
