"""Host-side scalar statistics (/root/reference/src/jaxns/internals/stats.py:55-86)."""
import math

_EPS = 2.220446049250313e-16


def linear_to_log_stats(log_f_mean, *, log_f2_mean=None, log_f_var=None):
    if log_f_var is not None:
        a, b = log_f_var, 2.0 * log_f_mean
        log_f2_mean = max(a, b) + math.log1p(math.exp(-abs(a - b)))
    mu = 2.0 * log_f_mean - 0.5 * log_f2_mean
    sigma2 = log_f2_mean - 2.0 * log_f_mean
    return mu, max(sigma2, _EPS)


def effective_sample_size_kish(log_Z_mean, log_dZ2_mean):
    return math.exp(2.0 * log_Z_mean - log_dZ2_mean)
