"""
jaxns_b200: B200-native drop-in for the static nested-sampling hot path of Joshuaalbert/jaxns 2.6.9.
Same user-facing names as `jaxns` (/root/reference/src/jaxns/__init__.py:3-7) for the path in scope.
"""
from jaxns_b200 import distributions, likelihoods, random  # noqa: F401
from jaxns_b200.context import (convert_external_params, get_parameter, next_rng_key, scope,  # noqa: F401
                                wrap_random)
from jaxns_b200.framework import Model, Prior, SingularPrior  # noqa: F401
from jaxns_b200.nested_sampler import ShardedStaticNestedSampler  # noqa: F401
from jaxns_b200.public import DefaultNestedSampler, NestedSampler  # noqa: F401
from jaxns_b200.samplers import UniDimSliceSampler, UniformSampler  # noqa: F401
from jaxns_b200.types import (NestedSamplerResults, NestedSamplerState, TerminationCondition)  # noqa: F401
from jaxns_b200.experimental import (DefaultGlobalOptimisation, EvidenceMaximisation, GlobalOptimisation,  # noqa: F401
                                     GlobalOptimisationResults, GlobalOptimisationTerminationCondition,
                                     SimpleGlobalOptimisation)
from jaxns_b200.likelihoods import jaxify_likelihood  # noqa: F401
from jaxns_b200.utils import (bruteforce_evidence, bruteforce_posterior_samples, evaluate_map_estimate,  # noqa: F401
                              load_pytree, load_results, marginalise_dynamic, marginalise_static,
                              maximum_a_posteriori_point, resample, sample_evidence, save_pytree, save_results, summary)

__version__ = "0.1.0"
