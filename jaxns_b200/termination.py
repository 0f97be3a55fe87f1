"""
Host mirror of determine_termination
(/root/reference/src/jaxns/nested_samplers/common/termination.py:13-147), used for
TerminationConditionConjunction / Disjunction trees (a plain TerminationCondition is decided on the
device inside the loop) and for converting conditions to the C-ABI struct.
"""
import math
from typing import Tuple

from jaxns_b200 import _lib
from jaxns_b200.internals.stats import linear_to_log_stats
from jaxns_b200.types import (TerminationCondition, TerminationConditionConjunction,
                              TerminationConditionDisjunction, TerminationRegister, EvidenceCalculation)


def _f(x):
    try:
        return float(x)
    except TypeError:
        return float(x.item())


def to_c(term_cond: TerminationCondition) -> _lib.NsTermCond:
    tc = _lib.NsTermCond()
    mask = 0
    for bit, name in enumerate(_lib.TERM_FIELDS):
        v = getattr(term_cond, name)
        if v is not None:
            mask |= 1 << bit
            setattr(tc, name, _f(v))
    tc.mask = mask
    return tc


def register_from_c(r: _lib.NsRegister) -> TerminationRegister:
    def ec(c):
        return EvidenceCalculation(*[getattr(c, n) for n, _ in _lib.NsEvidenceCalc._fields_])

    return TerminationRegister(num_samples_used=r.num_samples_used, evidence_calc=ec(r.evidence_calc),
                               evidence_calc_with_remaining=ec(r.evidence_calc_with_remaining),
                               num_likelihood_evaluations=r.num_likelihood_evaluations,
                               log_L_contour=r.log_L_contour, efficiency=r.efficiency, plateau=bool(r.plateau),
                               no_seed_points=bool(r.no_seed_points), relative_spread=r.relative_spread,
                               absolute_spread=r.absolute_spread, peak_log_XL=r.peak_log_XL)


def _sub(a, b):
    if math.isinf(a) and math.isinf(b) and (a > 0) == (b > 0):
        return math.nan
    return a - b


def determine_termination(term_cond, termination_register: TerminationRegister) -> Tuple[bool, int]:
    if isinstance(term_cond, TerminationConditionConjunction):
        # Deviation from the reference on purpose: termination.py:52-57 starts from done=False and ANDs the
        # children in, so a conjunction could never fire (and _main_ns_thread raises AttributeError on
        # `.max_samples` for it, sharded_static.py:464, so the reference cannot run one at all).  Here `a & b`
        # means what it says: done when every child is done; the reason is the union of the children's bits.
        done, reason = True, 0
        for c in term_cond.conds:
            d, r = determine_termination(c, termination_register)
            done = done and d
            reason = reason | r
        done = done and len(term_cond.conds) > 0
        return done, (reason if done else 0)
    if isinstance(term_cond, TerminationConditionDisjunction):
        done, reason = False, 0
        for c in term_cond.conds:
            d, r = determine_termination(c, termination_register)
            done = done or d
            reason = reason | r
        return done, reason
    reg = termination_register
    ec, ecr = reg.evidence_calc, reg.evidence_calc_with_remaining
    done, reason = False, 0

    def setbit(b, bit):
        nonlocal done, reason
        if b:
            done = True
            reason += 2 ** bit

    tc = term_cond
    if tc.max_samples is not None:
        setbit(reg.num_samples_used >= _f(tc.max_samples), 0)
    if tc.evidence_uncert is not None:
        _, v = linear_to_log_stats(ecr.log_Z_mean, log_f2_mean=ecr.log_Z2_mean)
        setbit(v <= _f(tc.evidence_uncert) ** 2, 1)
    if tc.dlogZ is not None:
        m1 = _sub(2.0 * ecr.log_Z_mean, 0.5 * ecr.log_Z2_mean)
        m0 = _sub(2.0 * ec.log_Z_mean, 0.5 * ec.log_Z2_mean)
        setbit(_sub(m1, m0) < _f(tc.dlogZ), 2)
    if tc.ess is not None:
        d = _sub(2.0 * ecr.log_Z_mean, ecr.log_dZ2_mean)
        setbit((not math.isnan(d)) and math.exp(min(d, 700.0)) >= _f(tc.ess), 3)
    if tc.max_num_likelihood_evaluations is not None:
        setbit(reg.num_likelihood_evaluations >= _f(tc.max_num_likelihood_evaluations), 4)
    if tc.log_L_contour is not None:
        setbit(reg.log_L_contour >= _f(tc.log_L_contour), 5)
    if tc.efficiency_threshold is not None:
        setbit(reg.efficiency < _f(tc.efficiency_threshold), 6)
    setbit(reg.plateau, 7)
    if tc.rtol is not None:
        setbit(reg.relative_spread < _f(tc.rtol), 8)
    if tc.atol is not None:
        setbit(reg.absolute_spread < _f(tc.atol), 9)
    setbit(reg.no_seed_points, 10)
    if tc.peak_XL_frac is not None:
        log_XL = ec.log_X_mean + ec.log_L
        setbit(log_XL < reg.peak_log_XL + math.log(_f(tc.peak_XL_frac)), 11)
    return done, reason
