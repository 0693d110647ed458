"""ott-packager_b200: B200-native (sm_100a) yadif + ABR-ladder scaler + thumbnail path of ott-packager's transvideo.c.

The product is libottgpu.so (C ABI, include/ottgpu.h); this package is only its ctypes binding.
Import with importlib.import_module("ott-packager_b200") (the directory name is not a Python identifier).
"""
from .binding import *  # noqa: F401,F403
from .binding import Library, Channel, Batch, Pool, make_cfg, OttGpuError, EXPORTS, DEFAULT_LIB  # noqa: F401
from .sharding import shard_channels, my_channels  # noqa: F401
