"""
Record types of the static nested sampler, field-for-field with
/root/reference/src/jaxns/nested_samplers/common/types.py:12-148.  Arrays are CUDA torch tensors
(float64 measures, int64 indices/counts, int32 num_live_points_per_sample).
"""
from typing import Any, List, NamedTuple, Optional, Union

__all__ = ["TerminationCondition", "NestedSamplerResults", "NestedSamplerState"]


class EvidenceCalculation(NamedTuple):
    log_L: Any
    log_X_mean: Any
    log_X2_mean: Any
    log_Z_mean: Any
    log_ZX_mean: Any
    log_Z2_mean: Any
    log_dZ_mean: Any
    log_dZ2_mean: Any


class TerminationCondition(NamedTuple):
    ess: Optional[Any] = None
    evidence_uncert: Optional[Any] = None
    live_evidence_frac: Optional[Any] = None
    dlogZ: Optional[Any] = None
    max_samples: Optional[Any] = None
    max_num_likelihood_evaluations: Optional[Any] = None
    log_L_contour: Optional[Any] = None
    efficiency_threshold: Optional[Any] = None
    rtol: Optional[Any] = None
    atol: Optional[Any] = None
    peak_XL_frac: Optional[Any] = None

    def __and__(self, other):
        return TerminationConditionConjunction(conds=[self, other])

    def __or__(self, other):
        return TerminationConditionDisjunction(conds=[self, other])


class TerminationConditionConjunction(NamedTuple):
    conds: List[Union["TerminationConditionDisjunction", "TerminationConditionConjunction", TerminationCondition]]


class TerminationConditionDisjunction(NamedTuple):
    conds: List[Union["TerminationConditionDisjunction", TerminationConditionConjunction, TerminationCondition]]


class NestedSamplerResults(NamedTuple):
    log_Z_mean: Any
    log_Z_uncert: Any
    ESS: Any
    H_mean: Any
    samples: Any
    parametrised_samples: Any
    U_samples: Any
    log_L_samples: Any
    log_dp_mean: Any
    log_X_mean: Any
    log_posterior_density: Any
    num_live_points_per_sample: Any
    num_likelihood_evaluations_per_sample: Any
    total_num_samples: Any
    total_phantom_samples: Any
    total_num_likelihood_evaluations: Any
    log_efficiency: Any
    termination_reason: Any


class Sample(NamedTuple):
    U_sample: Any
    log_L_constraint: Any
    log_L: Any
    num_likelihood_evaluations: Any


class LivePointCollection(NamedTuple):
    sender_node_idx: Any
    U_sample: Any
    log_L_constraint: Any
    log_L: Any
    num_likelihood_evaluations: Any


class SampleCollection(NamedTuple):
    sender_node_idx: Any
    log_L: Any
    U_samples: Any
    num_likelihood_evaluations: Any
    phantom: Any


StaticStandardSampleCollection = SampleCollection


class TerminationRegister(NamedTuple):
    num_samples_used: Any
    evidence_calc: EvidenceCalculation
    evidence_calc_with_remaining: EvidenceCalculation
    num_likelihood_evaluations: Any
    log_L_contour: Any
    efficiency: Any
    plateau: Any
    no_seed_points: Any
    relative_spread: Any
    absolute_spread: Any
    peak_log_XL: Any


class NestedSamplerState(NamedTuple):
    key: Any
    next_sample_idx: Any
    num_samples: Any
    sample_collection: StaticStandardSampleCollection


StaticStandardNestedSamplerState = NestedSamplerState
