"""csc_b200 -- B200-native libcsc compression hot path (C++/CUDA library + thin bindings).

The product is ``csrc/libcsc_b200.so``: the libcsc C API (CSCEnc_* / CSCDec_*,
include/csc_enc.h, include/csc_dec.h) plus the additive batched entry points of
include/csc_b200.h, all running on sm_100a kernels.  This package only binds
that C ABI with ctypes for tests and bench.py.  There is no CPU fallback: if the
library is missing or no CUDA device is usable, calls raise.
"""
from .api import (CSCProps, available, decode_chunks, decode_stream, encode_bound, encode_chunks, encode_stream,
                  get_stats, lib, props_init, set_device, set_devices, device_count, split_chunks, EXPORTS)

__all__ = ["CSCProps", "available", "decode_chunks", "decode_stream", "encode_bound", "encode_chunks",
           "encode_stream", "get_stats", "lib", "props_init", "set_device", "set_devices", "device_count", "split_chunks",
           "EXPORTS"]
