"""Make the reference's own entry scripts use this engine without editing them.

    import satlas_super_resolution_b200.dropin as dropin; dropin.install()      # before `import ssr...`
or  python -m satlas_super_resolution_b200.run ssr.infer -opt ssr/options/infer_example.yml

install() pre-seeds sys.modules with OUR implementations under the reference's module names

    ssr.archs.rrdbnet_arch        -> SSR_RRDBNet              (archs.py)
    ssr.archs.discriminator_arch  -> SSR_UNetDiscriminatorSN  (archs.py)
    ssr.models.ssr_esrgan_model   -> SSRESRGANModel           (models.py)

so the directory scanners in ssr/archs/__init__.py:7-10 and ssr/models/__init__.py:8-11 -- which call
importlib.import_module on exactly those names -- pick ours up and the `@ARCH_REGISTRY.register()` side effects land our
classes in the registries `build_network` / `build_model` read.  With the real `basicsr` installed nothing else is needed
(`ssr/train.py` keeps its data loaders, loggers, checkpointing).  Without it (this offline image) a minimal stand-in for
the basicsr / kornia / skimage names that are touched AT IMPORT TIME by ssr.archs, ssr.utils and ssr/infer*.py is
installed too; the training control plane (basicsr.data, basicsr.train, loggers) is deliberately NOT re-implemented.
"""
import logging
import os
import random
import sys
import types

_installed = False


def _missing(name):
    try:
        __import__(name)
        return False
    except Exception:
        return True


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_basicsr_shim():
    from . import registry as R

    def scandir(dir_path, suffix=None, recursive=False, full_path=False):
        for entry in sorted(os.listdir(dir_path)):
            if suffix is None or entry.endswith(suffix):
                yield os.path.join(dir_path, entry) if full_path else entry

    def set_random_seed(seed):
        import numpy as np
        import torch
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)

    def get_dist_info():
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def init_dist(launcher, backend="nccl", **kwargs):
        import torch
        import torch.distributed as dist
        if launcher != "pytorch":
            raise NotImplementedError("only the 'pytorch' launcher is provided by the stand-in")
        rank = int(os.environ["RANK"])
        torch.cuda.set_device(rank % torch.cuda.device_count())
        dist.init_process_group(backend=backend, **kwargs)

    def master_only(func):
        import functools

        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            if get_dist_info()[0] == 0:
                return func(*args, **kwargs)
        return wrapper

    class ModulatedDeformConvPack:
        def __init__(self, *a, **k):
            raise NotImplementedError("DCN is outside the built hot path")

    reg = _mod("basicsr.utils.registry", Registry=R.Registry, DATASET_REGISTRY=R.DATASET_REGISTRY, ARCH_REGISTRY=R.ARCH_REGISTRY,
               MODEL_REGISTRY=R.MODEL_REGISTRY, LOSS_REGISTRY=R.LOSS_REGISTRY, METRIC_REGISTRY=R.METRIC_REGISTRY,
               __ssr_b200_shim__=True)
    dist_util = _mod("basicsr.utils.dist_util", get_dist_info=get_dist_info, init_dist=init_dist, master_only=master_only)
    utils = _mod("basicsr.utils", scandir=scandir, get_root_logger=lambda *a, **k: logging.getLogger("basicsr"),
                 set_random_seed=set_random_seed, registry=reg, dist_util=dist_util)
    dcn = _mod("basicsr.ops.dcn", ModulatedDeformConvPack=ModulatedDeformConvPack, modulated_deform_conv=None)
    ops = _mod("basicsr.ops", dcn=dcn)
    arch_m = _mod("basicsr.archs", build_network=R.build_network)
    models_m = _mod("basicsr.models", build_model=R.build_model)
    losses_m = _mod("basicsr.losses", build_loss=R.build_loss)
    _mod("basicsr", utils=utils, ops=ops, archs=arch_m, models=models_m, losses=losses_m, __ssr_b200_shim__=True,
         __path__=[])


def _install_aux_shims():
    if _missing("kornia"):
        tr = _mod("kornia.geometry.transform", Resize=type("Resize", (), {}))
        geo = _mod("kornia.geometry", transform=tr)
        _mod("kornia", geometry=geo, __path__=[])
    if _missing("skimage"):
        def imread(path):
            import cv2
            img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
            if img is None:
                raise FileNotFoundError(path)
            return img[..., ::-1].copy() if img.ndim == 3 and img.shape[2] >= 3 else img

        def imsave(path, arr, check_contrast=False):
            import cv2
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            cv2.imwrite(path, arr[..., ::-1] if arr.ndim == 3 and arr.shape[2] == 3 else arr)

        io = _mod("skimage.io", imread=imread, imsave=imsave)
        _mod("skimage", io=io, __path__=[])


def install(reference_root=None):
    """Idempotent.  `reference_root`: directory that contains the reference's `ssr/` package (added to sys.path)."""
    global _installed
    from . import archs, models, registry  # noqa: F401  (registers everything)
    if registry._real is None and "basicsr" not in sys.modules:
        _install_basicsr_shim()
    _install_aux_shims()
    rrdb = _mod("ssr.archs.rrdbnet_arch", SSR_RRDBNet=archs.SSR_RRDBNet, RRDB=archs._RRDB, ResidualDenseBlock=archs._RDB)
    disc = _mod("ssr.archs.discriminator_arch", SSR_UNetDiscriminatorSN=archs.SSR_UNetDiscriminatorSN)
    esr = _mod("ssr.models.ssr_esrgan_model", SSRESRGANModel=models.SSRESRGANModel)
    for m in (rrdb, disc, esr):
        m.__ssr_b200__ = True
    root = reference_root or os.environ.get("SSR_REFERENCE_ROOT")
    if root and root not in sys.path:
        sys.path.insert(0, root)
    _installed = True
    return True
