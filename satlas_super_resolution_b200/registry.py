"""String-keyed class registries -- the reference's plug-in mechanism (basicsr.utils.registry).

If the real `basicsr` package is importable its registries are used (so `ssr/train.py` running on real basicsr finds
our classes); otherwise equivalent local registries are created and, through dropin.install(), exposed under the
`basicsr.utils.registry` name the reference imports.
"""
from copy import deepcopy


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj, force=False):
        if name in self._obj_map and not force:
            raise AssertionError(f"An object named '{name}' was already registered in '{self._name}' registry!")
        self._obj_map[name] = obj

    def register(self, obj=None, force=False):
        if obj is None:
            def deco(func_or_class):
                self._do_register(func_or_class.__name__, func_or_class, force)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj, force)
        return obj

    def get(self, name):
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name):
        return name in self._obj_map

    def keys(self):
        return self._obj_map.keys()


def _real_basicsr_registries():
    try:
        import basicsr.utils.registry as r  # noqa: F401
        if getattr(r, "__ssr_b200_shim__", False):
            return None
        return r
    except Exception:
        return None


_real = _real_basicsr_registries()
if _real is not None:
    DATASET_REGISTRY, ARCH_REGISTRY, MODEL_REGISTRY = _real.DATASET_REGISTRY, _real.ARCH_REGISTRY, _real.MODEL_REGISTRY
    LOSS_REGISTRY, METRIC_REGISTRY = _real.LOSS_REGISTRY, _real.METRIC_REGISTRY
else:
    DATASET_REGISTRY = Registry("dataset")
    ARCH_REGISTRY = Registry("arch")
    MODEL_REGISTRY = Registry("model")
    LOSS_REGISTRY = Registry("loss")
    METRIC_REGISTRY = Registry("metric")


def _register(reg, obj):
    """register, replacing an entry of the same name (the reference's own class may already sit there)"""
    try:
        reg.register(obj)
    except (AssertionError, KeyError):
        reg._obj_map[obj.__name__] = obj
    return obj


def build_network(opt):
    """basicsr.archs.build_network: pop 'type', construct ARCH_REGISTRY[type](**rest)  (ssr_esrgan_model.py:43,53)"""
    opt = deepcopy(opt)
    network_type = opt.pop("type")
    return ARCH_REGISTRY.get(network_type)(**opt)


def build_loss(opt):
    """ssr/losses/__init__.py:21-33"""
    opt = deepcopy(opt)
    loss_type = opt.pop("type")
    return LOSS_REGISTRY.get(loss_type)(**opt)


def build_model(opt):
    """basicsr.models.build_model: MODEL_REGISTRY[opt['model_type']](opt)  (ssr/train.py:62)"""
    opt = deepcopy(opt)
    return MODEL_REGISTRY.get(opt["model_type"])(opt)
