"""Import the UNMODIFIED reference nn.Modules from /root/reference in the build container (test infrastructure).

The reference needs `basicsr` (requirements.txt:1) and `kornia`, neither of which is installed or installable offline.
Only four import-time dependencies stand between us and `ssr.archs.{rrdbnet,discriminator}_arch`; they are stubbed here
with the few symbols the import executes: basicsr.utils.registry (the Registry class), basicsr.utils
(scandir, get_root_logger), basicsr.ops.dcn (names only; DCN is never constructed on this path),
kornia.geometry.transform (Resize name only).  Nothing of the reference's arithmetic is replaced.
"""
import logging
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SSR_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ssr", "archs"))


class _Registry:
    def __init__(self, name):
        self._name, self._obj_map = name, {}

    def register(self, obj=None):
        if obj is None:
            def deco(fn_or_cls):
                self._obj_map[fn_or_cls.__name__] = fn_or_cls
                return fn_or_cls
            return deco
        self._obj_map[obj.__name__] = obj
        return obj

    def get(self, name):
        return self._obj_map[name]

    def keys(self):
        return self._obj_map.keys()


def _scandir(dir_path, suffix=None, recursive=False, full_path=False):
    for entry in sorted(os.listdir(dir_path)):
        if suffix is None or entry.endswith(suffix):
            yield os.path.join(dir_path, entry) if full_path else entry


def install():
    """Insert the stubs (only for modules that are genuinely absent) and put the reference on sys.path."""
    if not available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present (it never is on the GPU box)")

    def need(name):
        try:
            __import__(name)
            return False
        except Exception:
            return True

    if need("basicsr"):
        basicsr = types.ModuleType("basicsr")
        utils = types.ModuleType("basicsr.utils")
        registry = types.ModuleType("basicsr.utils.registry")
        ops = types.ModuleType("basicsr.ops")
        dcn = types.ModuleType("basicsr.ops.dcn")
        registry.Registry = _Registry
        for n in ("DATASET_REGISTRY", "ARCH_REGISTRY", "MODEL_REGISTRY", "LOSS_REGISTRY", "METRIC_REGISTRY"):
            setattr(registry, n, _Registry(n))
        utils.scandir = _scandir
        utils.get_root_logger = lambda *a, **k: logging.getLogger("basicsr")
        utils.registry = registry

        class ModulatedDeformConvPack:  # name only: never instantiated on the ESRGAN path
            def __init__(self, *a, **k):
                raise NotImplementedError("DCN is outside the hot path")

        dcn.ModulatedDeformConvPack = ModulatedDeformConvPack
        dcn.modulated_deform_conv = None
        basicsr.utils, basicsr.ops, ops.dcn = utils, ops, dcn
        sys.modules.update({"basicsr": basicsr, "basicsr.utils": utils, "basicsr.utils.registry": registry,
                            "basicsr.ops": ops, "basicsr.ops.dcn": dcn})
    if need("kornia"):
        kornia = types.ModuleType("kornia")
        geometry = types.ModuleType("kornia.geometry")
        transform = types.ModuleType("kornia.geometry.transform")
        transform.Resize = type("Resize", (), {})
        kornia.geometry, geometry.transform = geometry, transform
        sys.modules.update({"kornia": kornia, "kornia.geometry": geometry, "kornia.geometry.transform": transform})
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def reference_archs():
    install()
    from ssr.archs.discriminator_arch import SSR_UNetDiscriminatorSN
    from ssr.archs.rrdbnet_arch import SSR_RRDBNet
    return SSR_RRDBNet, SSR_UNetDiscriminatorSN
