"""
Post-processing of nested-sampling results, name-for-name with /root/reference/src/jaxns/utils.py (SURVEY §8(f)
row 3): resample / marginalise_* / MAP helpers (:44-262), summary (:284-430), sample_evidence (:433-476) and the
results wire format save_results / load_results (:588-641, internals/namedtuple_utils.py:34-103).

Arrays are torch tensors on the device the results live on; random draws come from the device Threefry
(jaxns_b200.random), so a given key selects the same indices the reference's jax.random call would.
"""
import base64
import io
import json
import math
import warnings
from typing import Callable, Optional, TextIO, Union

import numpy as np
import torch

from jaxns_b200 import random
from jaxns_b200.types import NestedSamplerResults

__all__ = ["resample_indicies", "resample", "marginalise_static", "marginalise_dynamic",
           "maximum_a_posteriori_point", "maximum_a_posteriori_point_U", "evaluate_map_estimate", "summary",
           "sample_evidence", "bruteforce_evidence", "bruteforce_posterior_samples", "save_pytree", "save_results",
           "load_pytree", "load_results"]


def _tree_map(f, tree):
    if isinstance(tree, dict):
        return {k: _tree_map(f, v) for k, v in tree.items()}
    return f(tree)


def cumulative_logsumexp(u: torch.Tensor) -> torch.Tensor:
    """internals/log_semiring.py:51-92 (the reference scans serially; torch.logcumsumexp re-associates, ~1e-15)."""
    return torch.logcumsumexp(u, dim=0)


def resample_indicies(key, log_weights: Optional[torch.Tensor] = None, S: Optional[int] = None, replace: bool = True,
                      num_total: Optional[int] = None) -> torch.Tensor:
    """internals/random.py:34-75."""
    if S is None:
        if log_weights is None:
            raise ValueError("Need log_weights if S is not given.")
        # ESS = (sum w)^2 / sum w^2
        S = int(math.exp(2. * float(torch.logsumexp(log_weights, 0)) - float(torch.logsumexp(2. * log_weights, 0))))
    if replace:
        if log_weights is not None:
            log_p_cuml = cumulative_logsumexp(log_weights)
        else:
            if num_total is None:
                raise ValueError("Need num_total if log_weights is None.")
            log_p_cuml = torch.log(torch.arange(num_total, dtype=torch.float64, device="cuda"))
        log_r = log_p_cuml[-1] + torch.log(1. - random.uniform(key, S).to(log_p_cuml.device))
        return torch.searchsorted(log_p_cuml.contiguous(), log_r.contiguous())  # side='left', like jnp.searchsorted
    n = log_weights.shape[0] if log_weights is not None else num_total
    if n is None:
        raise ValueError("Need num_total if log_weights is None.")
    # random.gumbel = -log(-log(uniform(key, shape, minval=tiny, maxval=1)))
    tiny = float(np.finfo(np.float64).tiny)
    gumbel = -torch.log(-torch.log(random.uniform(key, n, minval=tiny, maxval=1.0)))
    g = -gumbel - (log_weights.to(gumbel.device) if log_weights is not None else 0.0)
    return torch.argsort(g, stable=True)[:S]


def resample(key, samples, log_weights: torch.Tensor, S: int = None, replace: bool = True):
    """utils.py:44-60: weighted samples -> equally weighted samples."""
    idx = resample_indicies(key, log_weights, S=S, replace=replace)
    return _tree_map(lambda s: s[idx.to(s.device), ...], samples)


def marginalise_static(key, samples: dict, log_weights: torch.Tensor, ESS: int, fun: Callable):
    """utils.py:154-171: nanmean of fun(**sample) over ESS resampled points."""
    rs = resample(key, samples, log_weights, S=int(ESS), replace=True)
    n = next(iter(rs.values())).shape[0]
    vals = [fun(**{k: v[i] for k, v in rs.items()}) for i in range(n)]
    return _tree_map(lambda x: torch.nanmean(x, dim=0),
                     _stack(vals))


def _stack(vals):
    first = vals[0]
    if isinstance(first, dict):
        return {k: _stack([v[k] for v in vals]) for k in first}
    return torch.stack([torch.as_tensor(v, dtype=torch.float64) for v in vals])


def marginalise_dynamic(key, samples: dict, log_weights: torch.Tensor, ESS, fun: Callable):
    """utils.py:174-211: one resampled point per step (key, resample_key = split(key)), NaN outputs skipped."""
    total, count = None, None
    for _ in range(int(ESS)):
        ks = random.split(key, 2)
        key, resample_key = ks[0], ks[1]
        one = resample(resample_key, samples, log_weights, S=1)
        y = torch.as_tensor(fun(**{k: v[0] for k, v in one.items()}), dtype=torch.float64)
        if total is None:
            total, count = torch.zeros_like(y), torch.zeros((), dtype=torch.float64, device=y.device)
        if not bool(torch.isnan(y).any()):
            count = count + 1
        total = torch.where(torch.isnan(y), total, total + y)
    return total / count


def maximum_a_posteriori_point(results: NestedSamplerResults):
    """utils.py:214-228."""
    i = int(torch.argmax(results.log_posterior_density))
    return _tree_map(lambda x: x[i], results.samples)


def maximum_a_posteriori_point_U(results: NestedSamplerResults):
    """utils.py:231-245."""
    i = int(torch.argmax(results.log_posterior_density))
    return results.U_samples[i]


def evaluate_map_estimate(results: NestedSamplerResults, fun: Callable):
    """utils.py:248-261."""
    return fun(**maximum_a_posteriori_point(results))


_REASONS = ('Reached max samples', 'Evidence uncertainty low enough', 'Small remaining evidence', 'Reached ESS',
            "Used max num likelihood evaluations", 'Likelihood contour reached', 'Sampler efficiency too low',
            'All live-points are on a single plateau (sign of possible precision error)',
            'relative spread of live points < rtol', 'absolute spread of live points < atol',
            'no seed points left (consider decreasing shell_fraction)', 'XL < max(XL) * peak_XL_frac')


def _sig_round(v, uncert_v):
    v, uncert_v = float(v), float(uncert_v)
    try:
        sig_figs = -int("{:e}".format(uncert_v).split('e')[1]) + 1
        return round(v, sig_figs)
    except Exception:
        return v


def summary(results: NestedSamplerResults, with_parametrised: bool = False,
            f_obj: Optional[Union[str, TextIO]] = None) -> str:
    """utils.py:284-430: same lines, same rounding rules, same fixed resampling key PRNGKey(23426)."""
    samples = dict(results.samples)
    if with_parametrised:
        samples.update(results.parametrised_samples)
    num_samples = int(results.total_num_samples)
    ESS = 100 if (results.ESS != results.ESS) else int(results.ESS)
    samples = {k: v[:num_samples] for k, v in samples.items()}
    log_L_samples = results.log_L_samples[:num_samples]
    log_dp_mean = results.log_dp_mean[:num_samples]
    max_like_idx = int(torch.argmax(log_L_samples))
    max_map_idx = int(torch.argmax(results.log_posterior_density))
    uniform_samples = resample(random.PRNGKey(23426), samples, log_dp_mean, S=max(100, ESS), replace=True)
    lines = []
    out = lines.append
    out("--------")
    out("Termination Conditions:")
    reason = int(results.termination_reason)
    for bit, text in enumerate(_REASONS):
        if (reason >> bit) & 1:
            out(text)
    out("--------")
    out(f"likelihood evals: {int(results.total_num_likelihood_evaluations):d}")
    out(f"samples: {num_samples:d}")
    out(f"phantom samples: {int(results.total_phantom_samples):d}")
    out(f"likelihood evals / sample: {float(results.total_num_likelihood_evaluations / results.total_num_samples):.1f}")
    out(f"phantom fraction (%): {100 * float(results.total_phantom_samples / results.total_num_samples):.1f}%")
    out("--------")
    out(f"logZ={_sig_round(results.log_Z_mean, results.log_Z_uncert)} +- "
        f"{_sig_round(results.log_Z_uncert, results.log_Z_uncert)}")
    out(f"max(logL)={_sig_round(log_L_samples[max_like_idx], results.log_Z_uncert)}")
    out(f"H={_sig_round(results.H_mean, 0.1)}")
    out(f"ESS={int(results.ESS) if results.ESS == results.ESS else results.ESS}")
    for name in uniform_samples.keys():
        s = uniform_samples[name].reshape(uniform_samples[name].shape[0], -1).cpu().numpy()
        ml = samples[name][max_like_idx].reshape(-1).cpu().numpy()
        mp = samples[name][max_map_idx].reshape(-1).cpu().numpy()
        ndims = s.shape[1]
        out("--------")
        var_name = name if ndims == 1 else "{}[#]".format(name)
        out(f"{var_name}: mean +- std.dev. | 10%ile / 50%ile / 90%ile | MAP est. | max(L) est.")
        for dim in range(ndims):
            unc = np.std(s[:, dim])
            sig_figs = -int("{:e}".format(unc).split('e')[1]) + 1

            def rnd(a):
                return round(float(a), sig_figs)

            out("{}: {} +- {} | {} / {} / {} | {} | {}".format(
                name if ndims == 1 else "{}[{}]".format(name, dim), rnd(np.mean(s[:, dim])), rnd(unc),
                *[rnd(a) for a in np.percentile(s[:, dim], np.asarray([10, 50, 90]))], rnd(mp[dim]), rnd(ml[dim])))
    out("--------")
    text = "\n".join(lines)
    if f_obj is None:
        print(text)
    elif isinstance(f_obj, str):
        with open(f_obj, 'w') as f:
            f.write(text)
    elif isinstance(f_obj, io.TextIOBase):
        f_obj.write(text)
    else:
        raise TypeError(f"Invalid f_obj: {type(f_obj)}")
    return text


def sample_evidence(key, num_live_points_per_sample: torch.Tensor, log_L_samples: torch.Tensor, S: int = 100
                    ) -> torch.Tensor:
    """utils.py:433-476: S stochastic simulations of the shrinkage, log T_i = log(u_i) / n_i with
    u_i = uniform(split(split(key, S)[s], M)[i]); returns the S samples of log Z (device kernel
    nsb200_sample_evidence: Threefry + log-space scan per simulation)."""
    import ctypes
    from jaxns_b200 import _lib
    _lib.require_cuda()
    n = num_live_points_per_sample.to(device="cuda", dtype=torch.float64).contiguous()
    logL = log_L_samples.to(device="cuda", dtype=torch.float64).contiguous()
    out = torch.empty(int(S), dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().nsb200_sample_evidence(_lib.key_arg(key), _lib.ptr(n), _lib.ptr(logL),
                                                 ctypes.c_int64(logL.numel()), ctypes.c_int64(int(S)), _lib.ptr(out),
                                                 _lib.stream_arg()))
    return out


# ---- results wire format (utils.py:588-641; internals/namedtuple_utils.py:34-103) ---------------------------
# The JSON schema is the reference's: a namedtuple is {'type': '__namedtuple__', '__class__': <dotted path>,
# '__data__': {...}}, an array {'type': '__jax_ndarray__' | '__ndarray__', '__dtype__', '__data__' (base64 of the raw
# little-endian bytes), '__shape__'}.  Files written here name the REFERENCE's classes, so jaxns.load_results reads
# them; files written by jaxns load here into the same-named types of jaxns_b200.types.
_REF_CLASS = {"NestedSamplerResults": "jaxns.nested_samplers.common.types.NestedSamplerResults",
              "TerminationCondition": "jaxns.nested_samplers.common.types.TerminationCondition",
              "EvidenceCalculation": "jaxns.internals.shrinkage_statistics.EvidenceCalculation"}
_INT_FIELDS = {"total_num_samples", "total_phantom_samples", "total_num_likelihood_evaluations", "termination_reason"}


def _ser_array(a: np.ndarray, kind: str):
    a = np.asarray(a)  # tobytes() is C-ordered whatever the strides; ascontiguousarray would turn 0-d into 1-d
    return {'type': kind, '__dtype__': str(a.dtype), '__data__': base64.b64encode(a.tobytes()).decode('utf-8'),
            '__shape__': list(a.shape)}


def _bruteforce_grid(model, S: int):
    eps = float(np.finfo(np.float64).eps)
    u_vec = torch.linspace(eps, 1.0 - eps, S, dtype=torch.float64, device="cuda")
    du = u_vec[1] - u_vec[0]
    D = model.U_ndims
    if S ** D > 2 ** 28:
        raise ValueError(f"bruteforce grid of {S}^{D} points is too large")
    args = torch.stack([x.flatten() for x in torch.meshgrid(*([u_vec] * D), indexing='ij')], dim=-1).contiguous()
    return args, du


def bruteforce_posterior_samples(model, S: int = 60):
    """utils.py:479-496: posterior over a regular grid in U space -> (samples, log weights)."""
    args, du = _bruteforce_grid(model, S)
    samples = model.transform(args)
    log_L = model.forward(args)
    return samples, log_L + model.U_ndims * torch.log(du)


def bruteforce_evidence(model, S: int = 60) -> float:
    """utils.py:499-516: log of sum_grid L du^D (NaN grid values are skipped, as LogSpace.nansum does)."""
    args, du = _bruteforce_grid(model, S)
    log_L = model.forward(args)
    log_L = log_L[~torch.isnan(log_L)]
    return float(torch.logsumexp(log_L, dim=0).item() + model.U_ndims * math.log(float(du.item())))


def _serialise(obj, field=None):
    if isinstance(obj, tuple) and hasattr(obj, '_asdict') and hasattr(obj, '_fields'):
        name = obj.__class__.__name__
        path = _REF_CLASS.get(name, f"{obj.__class__.__module__}.{name}")
        return {'type': '__namedtuple__', '__class__': path,
                '__data__': {k: _serialise(v, k) for k, v in obj._asdict().items()}}
    if isinstance(obj, torch.Tensor):
        return _ser_array(obj.detach().cpu().numpy(), '__jax_ndarray__')
    if isinstance(obj, np.ndarray):
        return _ser_array(obj, '__ndarray__')
    if isinstance(obj, (bool, np.bool_)):
        return _ser_array(np.asarray(obj, dtype=np.bool_), '__jax_ndarray__')
    if isinstance(obj, (int, np.integer)) and field is not None:
        return _ser_array(np.asarray(obj, dtype=np.int64), '__jax_ndarray__')  # 0-d arrays in the reference
    if isinstance(obj, (float, np.floating)) and field is not None:
        return _ser_array(np.asarray(obj, dtype=np.float64), '__jax_ndarray__')
    if isinstance(obj, (list, tuple)):
        return [_serialise(v) for v in obj]
    if isinstance(obj, dict):
        return {k: _serialise(v) for k, v in obj.items()}
    return obj


def _deserialise(obj, device):
    if isinstance(obj, dict) and obj.get('type') == '__namedtuple__':
        import importlib
        from jaxns_b200 import types as _types
        cls_name = obj['__class__'].rsplit('.', 1)[1]
        cls = getattr(_types, cls_name, None)
        if cls is None:  # a namedtuple of the caller's own module
            module_name = obj['__class__'].rsplit('.', 1)[0]
            cls = getattr(importlib.import_module(module_name), cls_name)
        # the file names the class to build: only NamedTuple classes are ever constructed (a results file is data,
        # it must not be able to call arbitrary callables with arguments of its choosing)
        if not (isinstance(cls, type) and issubclass(cls, tuple) and hasattr(cls, '_fields')):
            raise ValueError(f"{obj['__class__']} is not a NamedTuple class; refusing to deserialise it")
        return cls(**{k: _deserialise(v, device) for k, v in obj['__data__'].items()})
    if isinstance(obj, dict) and obj.get('type') in ('__ndarray__', '__jax_ndarray__'):
        a = np.frombuffer(base64.b64decode(obj['__data__']), dtype=obj['__dtype__']).reshape(obj['__shape__']).copy()
        if obj['type'] == '__ndarray__':
            return a
        if a.ndim == 0:
            return a.item()  # scalars of the results are host values here
        return torch.from_numpy(a).to(device)
    if isinstance(obj, list):
        return [_deserialise(v, device) for v in obj]
    if isinstance(obj, dict):
        return {k: _deserialise(v, device) for k, v in obj.items()}
    return obj


def save_pytree(pytree, save_file: str):
    """utils.py:588-599."""
    if not (isinstance(pytree, tuple) and hasattr(pytree, '_asdict') and hasattr(pytree, '_fields')):
        raise ValueError(f"Expected NamedTuple, got {type(pytree)}")
    with open(save_file, 'w') as fp:
        json.dump(_serialise(pytree), fp, indent=2)


def save_results(results: NestedSamplerResults, save_file: str):
    """utils.py:602-613."""
    if not save_file.lower().endswith('.json'):
        warnings.warn(f"Filename {save_file} does not end with .json. "
                      f"We let this pass, but you should consider using a .json file extension.")
    save_pytree(results, save_file)


def load_pytree(save_file: str, device: Optional[str] = None):
    """utils.py:616-628.  `device`: where arrays are placed (default: cuda if present, else cpu -- loading a file
    is not a compute path)."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    with open(save_file, 'r') as fp:
        data_dict = json.load(fp)
    return _deserialise(data_dict, device)


def load_results(save_file: str, device: Optional[str] = None) -> NestedSamplerResults:
    """utils.py:631-641."""
    return load_pytree(save_file, device)
