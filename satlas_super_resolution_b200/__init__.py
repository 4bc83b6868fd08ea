"""B200-native multi-frame ESRGAN engine behind the reference's registry surface.

Importing the package registers SSR_RRDBNet / SSR_UNetDiscriminatorSN (ARCH_REGISTRY), L1Loss / GANLoss / PerceptualLoss
(LOSS_REGISTRY), SSRESRGANModel (MODEL_REGISTRY) and the packed-shard S2NAIPShardDataset (DATASET_REGISTRY).  All arithmetic runs in libssr_b200.so (include/ssr_b200.h); the
library is loaded on first use and there is no CPU fallback.
"""
from . import registry  # noqa: F401
from . import archs, data, losses, metrics, models  # noqa: F401
from .registry import ARCH_REGISTRY, LOSS_REGISTRY, MODEL_REGISTRY, build_loss, build_model, build_network  # noqa: F401

__all__ = ["ARCH_REGISTRY", "LOSS_REGISTRY", "MODEL_REGISTRY", "build_network", "build_loss", "build_model"]
