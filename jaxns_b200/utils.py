"""Small post-processing helpers (text summary; /root/reference/src/jaxns/utils.py:284-430 is the model)."""
import math

from jaxns_b200.types import NestedSamplerResults

_REASONS = ("used maximum allowed number of samples", "evidence uncert below threshold",
            "live points evidence below threshold", "effective sample size big enough",
            "used maximum allowed number of likelihood evaluations", "maximum log-likelihood contour reached",
            "sampler efficiency too low", "entire live-points set is a single plateau",
            "relative spread of live points < rtol", "absolute spread of live points < atol",
            "no seed points left", "XL < max(XL) * peak_XL_frac")


def summary(results: NestedSamplerResults, f_obj=None) -> str:
    lines = ["--------", "Termination Conditions:"]
    reason = int(results.termination_reason)
    for bit, text in enumerate(_REASONS):
        if (reason >> bit) & 1:
            lines.append(text.capitalize())
    lines += ["--------",
              f"likelihood evals: {int(results.total_num_likelihood_evaluations)}",
              f"samples: {int(results.total_num_samples)}",
              f"phantom samples: {int(results.total_phantom_samples)}",
              f"likelihood evals / sample: {math.exp(-results.log_efficiency):.1f}",
              f"phantom fraction (%): {100.0 * results.total_phantom_samples / results.total_num_samples:.1f}%",
              "--------",
              f"logZ={results.log_Z_mean:.2f} +- {results.log_Z_uncert:.2f}",
              f"H={results.H_mean:.2f}",
              f"ESS={results.ESS:.0f}", "--------"]
    text = "\n".join(lines)
    if f_obj is not None:
        f_obj.write(text + "\n")
    else:
        print(text)
    return text
