"""compute_evidence_stats on the device
(/root/reference/src/jaxns/internals/shrinkage_statistics.py:112-157)."""
import ctypes
from typing import Optional, Tuple

import numpy as np
import torch

from jaxns_b200 import _lib
from jaxns_b200.types import EvidenceCalculation


def create_init_evidence_calc() -> EvidenceCalculation:
    ninf = -np.inf
    return EvidenceCalculation(log_L=ninf, log_X_mean=0.0, log_X2_mean=0.0, log_Z_mean=ninf, log_ZX_mean=ninf,
                               log_Z2_mean=ninf, log_dZ_mean=ninf, log_dZ2_mean=ninf)


def compute_evidence_stats(log_L: torch.Tensor, num_live_points: torch.Tensor, num_samples: Optional[int] = None,
                           init: Optional[EvidenceCalculation] = None, per_sample: bool = True
                           ) -> Tuple[EvidenceCalculation, Optional[EvidenceCalculation]]:
    """Returns (final EvidenceCalculation of python floats, per-sample EvidenceCalculation of
    float64 CUDA tensors).  With num_samples the scan stops there (cumulative_op_dynamic) and the
    per-sample tail is filled with the initial value, as in the reference."""
    _lib.require_cuda()
    log_L = torch.as_tensor(log_L, device="cuda").to(torch.float64).contiguous()
    n = torch.as_tensor(num_live_points, device="cuda").to(torch.float64).contiguous()
    M_full = log_L.numel()
    M = M_full if num_samples is None else int(num_samples)
    if init is None:
        init = create_init_evidence_calc()
    cinit = _lib.NsEvidenceCalc(*[float(v) for v in init])
    out_final = torch.empty(8, dtype=torch.float64, device="cuda")
    per = torch.empty((8, M), dtype=torch.float64, device="cuda") if per_sample else None
    L = _lib.lib()
    ws_bytes = L.nsb200_workspace_bytes(_lib.WS_EVIDENCE_STATS, M)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    _lib.check(L.nsb200_evidence_stats(ctypes.byref(cinit), _lib.ptr(log_L), _lib.ptr(n), ctypes.c_int64(M),
                                       _lib.ptr(out_final), _lib.ptr(per), _lib.ptr(ws), ctypes.c_int64(ws_bytes),
                                       _lib.stream_arg()))
    final = EvidenceCalculation(*out_final.cpu().tolist())
    if not per_sample:
        return final, None
    if M < M_full:
        fill = torch.tensor([float(v) for v in init], dtype=torch.float64, device="cuda")[:, None].expand(8, M_full - M)
        per = torch.cat([per, fill], dim=1)
    return final, EvidenceCalculation(*[per[i] for i in range(8)])


def logsumexp(x: torch.Tensor) -> float:
    _lib.require_cuda()
    x = torch.as_tensor(x, device="cuda").to(torch.float64).contiguous()
    out = torch.empty(1, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().nsb200_logsumexp(_lib.ptr(x), ctypes.c_int64(x.numel()), _lib.ptr(out), ctypes.c_void_p(0),
                                            ctypes.c_int64(0), _lib.stream_arg()))
    return float(out.item())
