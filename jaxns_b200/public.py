"""
NestedSampler / DefaultNestedSampler façade with the reference's fields and default logic
(/root/reference/src/jaxns/public.py:30-222).
"""
import dataclasses
import logging
from typing import Any, List, Optional, Tuple, Union

import numpy as np

from jaxns_b200.nested_sampler import ShardedStaticNestedSampler
from jaxns_b200.samplers import UniDimSliceSampler
from jaxns_b200.types import NestedSamplerResults, NestedSamplerState, TerminationCondition

logger = logging.getLogger("jaxns")

__all__ = ["NestedSampler", "DefaultNestedSampler"]

_COUNT_MAX = np.iinfo(np.int64).max


@dataclasses.dataclass(eq=False)
class NestedSampler:
    model: Any
    max_samples: Optional[Union[int, float]] = None
    num_live_points: Optional[int] = None
    num_slices: Optional[int] = None
    s: Optional[Union[int, float]] = None
    k: Optional[int] = None
    c: Optional[int] = None
    devices: Optional[List[Any]] = None
    difficult_model: bool = False
    parameter_estimation: bool = False
    shell_fraction: float = 0.5
    gradient_guided: bool = False
    init_efficiency_threshold: float = 0.1
    verbose: bool = False

    def __post_init__(self):
        # number of slices per acceptance (public.py:69-79)
        if self.num_slices is None:
            if self.difficult_model:
                self.s = 10 if self.s is None else float(self.s)
            else:
                self.s = 5 if self.s is None else float(self.s)
            if self.s <= 0:
                raise ValueError(f"Expected s > 0, got s={self.s}")
            self.num_slices = self.model.U_ndims * self.s
        self.num_slices = int(self.num_slices)

        # number of phantom samples (:81-89)
        if self.parameter_estimation:
            if self.s is None:
                raise TypeError("parameter_estimation=True needs s (the reference fails here too when "
                                "num_slices is given without s).")
            max_k = self.s * self.model.U_ndims - 1
            self.k = min(self.model.U_ndims, max_k) if self.k is None else int(self.k)
        else:
            self.k = 0 if self.k is None else int(self.k)
        if not (0 <= self.k < self.num_slices):
            raise ValueError(
                f"Expected 0 <= k < num_slices, got k={self.k}, num_slices={self.num_slices}, "
                f"U_ndims={self.model.U_ndims}")

        # number of parallel Markov chains (:91-101)
        if self.num_live_points is not None:
            self.c = max(1, int(self.num_live_points / (self.k + 1)))
            logger.info(f"Number of Markov-chains set to: {self.c}")
        else:
            if self.difficult_model:
                self.c = 100 * self.model.U_ndims if self.c is None else int(self.c)
            else:
                self.c = 30 * self.model.U_ndims if self.c is None else int(self.c)
            if self.c <= 0:
                raise ValueError(f"Expected c > 0, got c={self.c}")

        # default to 100 shrinkages (:103-107)
        if self.max_samples is None:
            self.max_samples = self.c * (self.k + 1) * 100
        self.max_samples = int(self.max_samples)

        self._nested_sampler = ShardedStaticNestedSampler(
            model=self.model,
            num_live_points=self.c,
            max_samples=self.max_samples,
            sampler=UniDimSliceSampler(
                model=self.model,
                num_slices=self.num_slices,
                num_phantom_save=self.k,
                midpoint_shrink=not self.difficult_model,
                gradient_guided=self.gradient_guided,
                perfect=True
            ),
            init_efficiency_threshold=self.init_efficiency_threshold,
            shell_fraction=self.shell_fraction,
            devices=self.devices,
            verbose=self.verbose,
        )
        self.num_live_points = self._nested_sampler.num_live_points

    @property
    def nested_sampler(self) -> ShardedStaticNestedSampler:
        return self._nested_sampler

    def __call__(self, key, term_cond: Optional[TerminationCondition] = None) -> Tuple[int, NestedSamplerState]:
        """public.py:141-175"""
        if term_cond is None:
            if self.parameter_estimation:
                term_cond = TerminationCondition(peak_XL_frac=0.1, max_samples=_COUNT_MAX)
            else:
                term_cond = TerminationCondition(dlogZ=float(np.log(1. + 1e-3)), max_samples=_COUNT_MAX)
        if isinstance(term_cond, TerminationCondition):
            term_cond = term_cond._replace(
                max_samples=(min(float(term_cond.max_samples), float(self._nested_sampler.max_samples))
                             if term_cond.max_samples is not None else float(self._nested_sampler.max_samples)))
        termination_reason, termination_register, state = self._nested_sampler._run(key=key, term_cond=term_cond)
        return termination_reason, state

    def to_results(self, termination_reason, state: NestedSamplerState, trim: bool = True) -> NestedSamplerResults:
        return self._nested_sampler._to_results(termination_reason=termination_reason, state=state, trim=trim)

    @staticmethod
    def trim_results(results: NestedSamplerResults) -> NestedSamplerResults:
        n = int(results.total_num_samples)

        def trim(x):
            if hasattr(x, "numel") and x.numel() > 1:
                return x[:n]
            if isinstance(x, dict):
                return {k: trim(v) for k, v in x.items()}
            return x

        return NestedSamplerResults(*[trim(x) for x in results])

    def summary(self, results: NestedSamplerResults) -> str:
        from jaxns_b200.utils import summary
        return summary(results)


DefaultNestedSampler = NestedSampler
