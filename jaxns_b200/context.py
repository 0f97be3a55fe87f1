"""
Parameters of parametrised models: `get_parameter` with the reference's signature and first-use initialisation
(/root/reference/src/jaxns/framework/context.py:186-216), `scope` name prefixes (:160-176) and the context's key chain
for random initialisers (`next_rng_key`, `wrap_random`: :100-103, :243-260).

The reference threads parameters through `transform_with_state(f).init / .apply`; here the Model opens a context
around every evaluation of its prior model and likelihood (`bind`): an initialising context records the values the
`init` callables produce, a bound context hands the caller's values (torch tensors, possibly requiring grad -- the
M-step of EvidenceMaximisation differentiates through them) back under the same names.
"""
import warnings
from contextlib import contextmanager
from functools import wraps
from typing import Dict, List, Optional

import numpy as np
import torch

__all__ = ["get_parameter", "scope", "next_rng_key", "wrap_random", "convert_external_params", "bind"]


class _Ctx:
    def __init__(self, params: Optional[Dict[str, torch.Tensor]], rng):
        self.params: Dict[str, torch.Tensor] = dict(params) if params is not None else {}
        self.scopes: List[str] = []
        self.rng = rng

    def full_name(self, name: str) -> str:
        return ".".join(self.scopes + [name]) if self.scopes else name


_stack: List[_Ctx] = []


def _top() -> _Ctx:
    if not _stack:
        raise ValueError("No context available. get_parameter must be called from a Model's prior_model / log_likelihood.")
    return _stack[-1]


@contextmanager
def bind(params: Optional[Dict[str, torch.Tensor]] = None, rng=None):
    """Evaluate model code with `params` bound (None: an initialising context).  Yields the context; after the block
    `ctx.params` holds every parameter the code asked for."""
    ctx = _Ctx(params, rng)
    _stack.append(ctx)
    try:
        yield ctx
    finally:
        _stack.remove(ctx)


@contextmanager
def scope(name: str):
    """context.py:160-176: prefix parameter names with {current_scope}.{name}."""
    ctx = _top()
    ctx.scopes.append(name)
    try:
        yield
    finally:
        ctx.scopes.pop()


def next_rng_key():
    from jaxns_b200 import random
    ctx = _top()
    if ctx.rng is None:
        ctx.rng = random.PRNGKey(0)
    ctx.rng, new = random.split(ctx.rng, 2)
    return new


def wrap_random(f):
    @wraps(f)
    def wrapped(*args, **kwargs):
        return f(next_rng_key(), *args, **kwargs)

    return wrapped


def _device():
    return "cuda" if torch.cuda.is_available() else "cpu"


def _default_init(shape, dtype):
    raise NotImplementedError("No init provided.")


def get_parameter(name: str, shape=None, dtype=None, *, init=_default_init) -> torch.Tensor:
    """context.py:186-216.  `init(shape, dtype)` (or `init()` when neither is given, or a constant) may return numpy
    or torch values; parameters live as float64 device tensors."""
    ctx = _top()
    key = ctx.full_name(name)
    if key not in ctx.params:
        if callable(init):
            value = init() if (shape is None and dtype is None) else init(shape, dtype if dtype is not None else np.float64)
        else:
            warnings.warn("Using a constant initializer for state. This is not recommended as it may induce closure issues.")
            value = init
        if not isinstance(value, torch.Tensor):
            value = torch.as_tensor(np.asarray(value, np.float64))
        ctx.params[key] = value.to(dtype=torch.float64, device=_device())
    return ctx.params[key]


def convert_external_params(external_params: Dict[str, torch.Tensor], prefix: str):
    """context.py:221-240 for a flat mapping name -> value."""
    return {k: get_parameter(f"__{prefix}_{i}", init=v) for i, (k, v) in enumerate(external_params.items())}
