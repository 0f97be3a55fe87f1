"""count_crossed_edges on the device (/root/reference/src/jaxns/internals/tree_structure.py:12-108)."""
import ctypes
from typing import NamedTuple, Optional

import torch

from jaxns_b200 import _lib


class SampleTreeGraph(NamedTuple):
    sender_node_idx: torch.Tensor  # int64 [N]
    log_L: torch.Tensor  # float64 [N]


class SampleLivePointCounts(NamedTuple):
    samples_indices: torch.Tensor  # int64 [N]
    num_live_points: torch.Tensor  # int32 [N]


def count_crossed_edges(sample_tree: SampleTreeGraph, num_samples: Optional[int] = None) -> SampleLivePointCounts:
    _lib.require_cuda()
    sender = torch.as_tensor(sample_tree.sender_node_idx, device="cuda").to(torch.int64).contiguous()
    log_L = torch.as_tensor(sample_tree.log_L, device="cuda").to(torch.float64).contiguous()
    M = sender.numel()
    idx = torch.empty(M, dtype=torch.int64, device="cuda")
    nlive = torch.empty(M, dtype=torch.int32, device="cuda")
    if M == 0:
        return SampleLivePointCounts(idx, nlive)
    L = _lib.lib()
    ws_bytes = L.nsb200_workspace_bytes(_lib.WS_COUNT_CROSSED_EDGES, M)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    _lib.check(L.nsb200_count_crossed_edges(_lib.ptr(sender), _lib.ptr(log_L), ctypes.c_int64(M),
                                            ctypes.c_int64(-1 if num_samples is None else int(num_samples)),
                                            _lib.ptr(idx), _lib.ptr(nlive), _lib.ptr(ws), ctypes.c_int64(ws_bytes),
                                            _lib.stream_arg()))
    return SampleLivePointCounts(samples_indices=idx, num_live_points=nlive)


def argsort(keys: torch.Tensor) -> torch.Tensor:
    """Stable jnp.argsort of float64 keys (-0 == +0, NaN last)."""
    _lib.require_cuda()
    keys = torch.as_tensor(keys, device="cuda").to(torch.float64).contiguous()
    n = keys.numel()
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    if n == 0:
        return out
    L = _lib.lib()
    ws_bytes = L.nsb200_workspace_bytes(_lib.WS_ARGSORT, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    _lib.check(L.nsb200_argsort_f64(_lib.ptr(keys), ctypes.c_int64(n), _lib.ptr(out), _lib.ptr(ws),
                                    ctypes.c_int64(ws_bytes), _lib.stream_arg()))
    return out
