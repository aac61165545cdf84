"""mchap_b200 — B200-native (sm_100a) implementation of MCHap's haplotype / genotype inference
hot path behind the reference's Python call surface.

Importing the package does not touch the GPU; every compute entry point goes through the CUDA
library ``mchap_b200/_lib/libmchap_b200.so`` (C ABI in ``include/mchap_b200.h``) and raises if
the library or an sm_100 device is missing — there is no CPU fallback.
"""
from .assemble.mcmc import DenovoMCMC
from .calling.classes import CallingMCMC
from .api import Device, default_device, MchapB200Error

__version__ = "0.1.0"

__all__ = ["DenovoMCMC", "CallingMCMC", "Device", "default_device", "MchapB200Error", "__version__"]
