"""vkjit_b200 — B200-native execution backend for vkjit's trace -> kernel -> launch -> readback path.

`Ir` mirrors `vkjit_core::Ir`; `vkjit_b200.vkjit` mirrors the vkjit-python module
(`Var`, `eval`, `var`, `ir`, `linspace`).  Everything below them is the native
library `libvkjit_b200.so` (C ABI in include/vkjit_b200.h).  There is no CPU path:
without the built library the import of any compute entry point raises.
"""
from ._capi import (VkjitError, VkjitNoDeviceError, VkjitSizeError, VkjitTypeError, product_api)  # noqa: F401
from .ir import (Bop, Ir, Red, Uop, VarType, cache_clear, init, shutdown, stats, stats_reset, stream_ptr,  # noqa: F401
                 sync)

__all__ = ["Ir", "VarType", "Bop", "Uop", "Red", "init", "shutdown", "sync", "stats", "stats_reset",
           "cache_clear", "stream_ptr", "VkjitError", "VkjitTypeError", "VkjitSizeError", "VkjitNoDeviceError"]
