"""
jax.random subset used by the jaxns hot path, on the device, bit-exact with JAX under
jax_threefry_partitionable=True (forced by /root/reference/src/jaxns/internals/mixed_precision.py:11-15).
Keys are uint32[2] (PRNGKey(seed) = (seed >> 32, seed & 0xffffffff)); results are CUDA tensors.
"""
import ctypes

import numpy as np
import torch

from jaxns_b200 import _lib


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def threefry2x32(key, x0: torch.Tensor, x1: torch.Tensor):
    """threefry2x32 primitive over device uint32 counter arrays (stored as int32/uint32 tensors)."""
    _lib.require_cuda()
    n = x0.numel()
    o0 = torch.empty_like(x0)
    o1 = torch.empty_like(x1)
    _lib.check(_lib.lib().nsb200_threefry2x32(_lib.key_arg(key), _lib.ptr(x0), _lib.ptr(x1), ctypes.c_int64(n),
                                               _lib.ptr(o0), _lib.ptr(o1), _lib.stream_arg()))
    return o0, o1


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split(key, num) -> uint32[num, 2] (host array; keys are host-side values in the API)."""
    _lib.require_cuda()
    out = torch.empty((num, 2), dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().nsb200_random_split(_lib.key_arg(key), ctypes.c_int64(num), _lib.ptr(out),
                                               _lib.stream_arg()))
    return out.cpu().numpy().view(np.uint32)


def bits(key, n: int) -> torch.Tensor:
    """jax.random.bits(key, (n,), uint64) as an int64 tensor holding the same bit pattern."""
    _lib.require_cuda()
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().nsb200_random_bits64(_lib.key_arg(key), ctypes.c_int64(n), _lib.ptr(out),
                                                _lib.stream_arg()))
    return out


def uniform(key, n: int, minval: float = 0.0, maxval: float = 1.0) -> torch.Tensor:
    _lib.require_cuda()
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().nsb200_random_uniform(_lib.key_arg(key), ctypes.c_int64(n), ctypes.c_double(minval),
                                                 ctypes.c_double(maxval), _lib.ptr(out), _lib.stream_arg()))
    return out


def normal(key, n: int) -> torch.Tensor:
    _lib.require_cuda()
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().nsb200_random_normal(_lib.key_arg(key), ctypes.c_int64(n), _lib.ptr(out),
                                                _lib.stream_arg()))
    return out
