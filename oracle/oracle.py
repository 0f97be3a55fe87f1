"""
CPU oracle for the jaxns 2.6.9 static nested-sampling hot path (numpy driver over ns_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py.  Nothing under jaxns_b200/ imports this module.

Parity status: "parity unpinned" against a LIVE jaxns run (jax/jaxlib/tfp are not installable in the
build container; see ns_oracle.c header): per-chain trajectories, the word order of 64-bit draws and
XLA's f64 erf_inv rounding have no external vector.  Pinned: Threefry Random123 KATs; the partitionable
split / counter layout, 32-bit draws and the uniform -> normal recipe against the values the JAX
documentation publishes for key(42) (tests/test_oracle_cpu.py::
test_partitionable_stream_matches_published_jax_tutorial); ndtri (scipy Cephes), erf_inv (scipy),
tree-count golden vectors of the reference's own tests, log-space identities.

Reference paths are relative to /root/reference/src/jaxns.
"""
import ctypes
import os
import subprocess
from typing import NamedTuple, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FAM_GAUSS_DENSE, FAM_GAUSS_MIX_DIAG, FAM_EGGBOX, FAM_ROSENBROCK, FAM_SHELLS = range(5)
PRIOR_UNIFORM, PRIOR_NORMAL = 0, 1

_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u32p = ctypes.POINTER(ctypes.c_uint32)


class _CModel(ctypes.Structure):
    _fields_ = [("family", ctypes.c_int32), ("D", ctypes.c_int32), ("prior_kind", ctypes.c_int32),
                ("K", ctypes.c_int32), ("prior_a", _f64p), ("prior_b", _f64p), ("params", _f64p)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libns_oracle.so")
    src = os.path.join(_HERE, "ns_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libns_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.o_erfinv.restype = ctypes.c_double
        _LIB.o_erfinv.argtypes = [ctypes.c_double]
        _LIB.o_ndtri.restype = ctypes.c_double
        _LIB.o_ndtri.argtypes = [ctypes.c_double]
        _LIB.o_logaddexp.restype = ctypes.c_double
        _LIB.o_logaddexp.argtypes = [ctypes.c_double, ctypes.c_double]
        _LIB.o_seed_index_scan.restype = ctypes.c_int64
        _LIB.o_seed_index_table.restype = ctypes.c_int64
        _LIB.o_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


def _key(k):
    k = np.ascontiguousarray(np.asarray(k, dtype=np.uint32).reshape(2))
    return k


# ----------------------------------------------------------------------------------------------
# jax.random restatement (jax/_src/prng.py, random.py; partitionable Threefry)
# ----------------------------------------------------------------------------------------------
def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def threefry2x32(k0, k1, x0, x1):
    o0 = ctypes.c_uint32()
    o1 = ctypes.c_uint32()
    lib().o_threefry2x32(ctypes.c_uint32(int(k0)), ctypes.c_uint32(int(k1)), ctypes.c_uint32(int(x0)),
                         ctypes.c_uint32(int(x1)), ctypes.byref(o0), ctypes.byref(o1))
    return o0.value, o1.value


def split(key, n: int = 2) -> np.ndarray:
    key = _key(key)
    out = np.empty((n, 2), np.uint32)
    for i in range(n):
        out[i] = threefry2x32(key[0], key[1], i >> 32, i & 0xFFFFFFFF)
    return out


def random_bits64(key, n: int) -> np.ndarray:
    key = _key(key)
    out = np.empty(n, np.uint64)
    for i in range(n):
        a, b = threefry2x32(key[0], key[1], i >> 32, i & 0xFFFFFFFF)
        out[i] = (a << 32) | b
    return out


def uniform(key, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    bits = random_bits64(key, n)
    f = ((bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)).view(np.float64) - 1.0
    lo = np.float64(lo)
    hi = np.float64(hi)
    return np.maximum(lo, f * (hi - lo) + lo)


def erfinv(x) -> np.ndarray:
    return np.array([lib().o_erfinv(float(v)) for v in np.atleast_1d(x)])


def ndtri(p) -> np.ndarray:
    return np.array([lib().o_ndtri(float(v)) for v in np.atleast_1d(p)])


def normal(key, n: int) -> np.ndarray:
    lo = np.nextafter(np.float64(-1.0), np.float64(0.0))
    u = uniform(key, n, lo, 1.0)
    return np.float64(np.sqrt(2.0)) * erfinv(u)


def logaddexp(a, b):
    return lib().o_logaddexp(float(a), float(b))


# ----------------------------------------------------------------------------------------------
# Model descriptor (mirrors the packed layout documented in include/nsb200.h)
# ----------------------------------------------------------------------------------------------
class OModel:
    def __init__(self, family: int, D: int, prior_kind: int, prior_a, prior_b, params, K: int = 0):
        self.family = int(family)
        self.D = int(D)
        self.prior_kind = int(prior_kind)
        self.K = int(K)
        self.prior_a = np.ascontiguousarray(np.broadcast_to(np.asarray(prior_a, np.float64), (D,)))
        self.prior_b = np.ascontiguousarray(np.broadcast_to(np.asarray(prior_b, np.float64), (D,)))
        self.params = np.ascontiguousarray(np.asarray(params, np.float64).reshape(-1))
        if self.params.size == 0:
            self.params = np.zeros(1)
        self._c = _CModel(self.family, self.D, self.prior_kind, self.K, _p(self.prior_a, _f64p),
                          _p(self.prior_b, _f64p), _p(self.params, _f64p))

    @property
    def c(self):
        return ctypes.byref(self._c)

    def forward(self, U) -> np.ndarray:
        U = np.ascontiguousarray(np.asarray(U, np.float64).reshape(-1, self.D))
        out = np.empty(U.shape[0])
        lib().o_forward_batch(self.c, _p(U, _f64p), ctypes.c_int64(U.shape[0]), _p(out, _f64p), None)
        return out

    def transform(self, U) -> np.ndarray:
        U = np.ascontiguousarray(np.asarray(U, np.float64).reshape(-1, self.D))
        out = np.empty(U.shape[0])
        X = np.empty_like(U)
        lib().o_forward_batch(self.c, _p(U, _f64p), ctypes.c_int64(U.shape[0]), _p(out, _f64p), _p(X, _f64p))
        return X


def pack_gauss_dense(mu, cov) -> np.ndarray:
    """[c, mu[D], Linv[D*D]] for MultivariateNormalTriL(mu, chol(cov)).log_prob."""
    mu = np.asarray(mu, np.float64)
    D = mu.size
    L = np.linalg.cholesky(np.asarray(cov, np.float64))
    Linv = np.linalg.solve(L, np.eye(D))
    Linv = np.tril(Linv)
    c = -np.sum(np.log(np.diag(L))) - 0.5 * D * np.log(2.0 * np.pi)
    return np.concatenate([[c], mu, Linv.reshape(-1)])


def gauss_model(D: int, data_mu=15.0, rho=0.99) -> "OModel":
    """Paper convention (docs/papers/phantom-powered-nested-sampling/run_experiment.py:23-50):
    prior N(0, I), likelihood N(data_mu * 1, Sigma) with unit diagonal and rho off-diagonal."""
    cov = np.full((D, D), rho) + (1.0 - rho) * np.eye(D)
    return OModel(FAM_GAUSS_DENSE, D, PRIOR_NORMAL, np.zeros(D), np.ones(D),
                  pack_gauss_dense(np.full(D, data_mu), cov))


def gauss_analytic_logZ(D: int, data_mu=15.0, rho=0.99) -> float:
    cov = np.full((D, D), rho) + (1.0 - rho) * np.eye(D) + np.eye(D)
    L = np.linalg.cholesky(cov)
    z = np.linalg.solve(L, np.full(D, data_mu))
    return float(-0.5 * z @ z - np.sum(np.log(np.diag(L))) - 0.5 * D * np.log(2 * np.pi))


def eggbox_model(D: int = 2) -> "OModel":
    return OModel(FAM_EGGBOX, D, PRIOR_UNIFORM, np.zeros(D), np.full(D, 10.0 * np.pi), [])


def rosenbrock_model(D: int = 10) -> "OModel":
    return OModel(FAM_ROSENBROCK, D, PRIOR_UNIFORM, np.full(D, -5.0), np.full(D, 10.0), [])


def shells_model(D: int = 2, w=0.1, r=2.0, offset=3.0, half_width=6.0) -> "OModel":
    c1 = np.zeros(D)
    c2 = np.zeros(D)
    c1[1 if D > 1 else 0] = -offset
    c2[1 if D > 1 else 0] = offset
    params = np.concatenate([[w, r], c1, [w, r], c2])
    return OModel(FAM_SHELLS, D, PRIOR_UNIFORM, np.full(D, -half_width), np.full(D, 2 * half_width), params, K=2)


def mixture_model(D: int = 100) -> "OModel":
    """spike-and-slab generalisation (benchmarks/difficult_problems/main.py:95-125)."""
    m1 = np.zeros(D)
    m2 = np.zeros(D)
    m1[:2] = 6.0
    m2[:2] = 2.5
    comps = []
    for mean, var in ((m1, 0.08), (m2, 0.8)):
        logc = -0.5 * D * np.log(2 * np.pi * var)
        comps.append(np.concatenate([[logc], mean, np.full(D, 1.0 / np.sqrt(var))]))
    return OModel(FAM_GAUSS_MIX_DIAG, D, PRIOR_UNIFORM, np.full(D, -4.0), np.full(D, 12.0),
                  np.concatenate(comps), K=2)


# ----------------------------------------------------------------------------------------------
# Batched samplers
# ----------------------------------------------------------------------------------------------
def seed_table(N: int) -> np.ndarray:
    c = np.empty(N)
    lib().o_seed_table(ctypes.c_int64(N), _p(c, _f64p))
    return c


def seed_index_scan(live_logL, contour, u) -> int:
    a = np.ascontiguousarray(live_logL, np.float64)
    return lib().o_seed_index_scan(_p(a, _f64p), ctypes.c_int64(a.size), ctypes.c_double(contour),
                                   ctypes.c_double(u))


def seed_index_table(live_logL, ctab, contour, u) -> int:
    a = np.ascontiguousarray(live_logL, np.float64)
    return lib().o_seed_index_table(_p(a, _f64p), ctypes.c_int64(a.size), _p(ctab, _f64p),
                                    ctypes.c_double(contour), ctypes.c_double(u))


def init_batch(model: OModel, sample_key, N: int):
    U = np.empty((N, model.D))
    logL = np.empty(N)
    nev = np.empty(N, np.int64)
    k = _key(sample_key)
    lib().o_init_batch(model.c, _p(k, _u32p), ctypes.c_int64(N), _p(U, _f64p), _p(logL, _f64p), _p(nev, _i64p))
    return U, logL, nev


def slice_batch(model: OModel, key, contour, live_U, live_logL, S, k=0, midpoint=True, num_samples=None,
                chain_begin=0, chain_end=None, ctab=None):
    """get_samples (nested_samplers/sharded/sharded_static.py:88-129) with UniDimSliceSampler."""
    live_U = np.ascontiguousarray(live_U, np.float64)
    live_logL = np.ascontiguousarray(live_logL, np.float64)
    N = live_logL.size
    if ctab is None:
        ctab = seed_table(N)
    if num_samples is None:
        num_samples = N // 2
    if chain_end is None:
        chain_end = num_samples
    n = chain_end - chain_begin
    D = model.D
    out_U = np.empty((n, D))
    out_logL = np.empty(n)
    out_nev = np.empty(n, np.int64)
    ph_U = np.empty((n * max(k, 1), D))
    ph_logL = np.empty(n * max(k, 1))
    seed_idx = np.empty(n, np.int64)
    kk = _key(key)
    lib().o_slice_batch(model.c, _p(kk, _u32p), ctypes.c_double(contour), _p(live_U, _f64p), _p(live_logL, _f64p),
                        ctypes.c_int64(N), _p(ctab, _f64p), ctypes.c_int(S), ctypes.c_int(k),
                        ctypes.c_int(int(midpoint)), ctypes.c_int64(chain_begin), ctypes.c_int64(chain_end),
                        _p(out_U, _f64p), _p(out_logL, _f64p), _p(out_nev, _i64p), _p(ph_U, _f64p),
                        _p(ph_logL, _f64p), _p(seed_idx, _i64p))
    return dict(U=out_U, log_L=out_logL, n_evals=out_nev, ph_U=ph_U[:n * k], ph_log_L=ph_logL[:n * k],
                seed_idx=seed_idx)


def uniform_batch(model: OModel, key, contour, num_samples, chain_begin=0, chain_end=None):
    if chain_end is None:
        chain_end = num_samples
    n = chain_end - chain_begin
    out_U = np.empty((n, model.D))
    out_logL = np.empty(n)
    out_nev = np.empty(n, np.int64)
    kk = _key(key)
    lib().o_uniform_batch(model.c, _p(kk, _u32p), ctypes.c_double(contour), ctypes.c_int64(chain_begin),
                          ctypes.c_int64(chain_end), _p(out_U, _f64p), _p(out_logL, _f64p), _p(out_nev, _i64p))
    return dict(U=out_U, log_L=out_logL, n_evals=out_nev)


# ----------------------------------------------------------------------------------------------
# Statistics: tree counts, evidence recurrences
# ----------------------------------------------------------------------------------------------
def stable_argsort(x) -> np.ndarray:
    """jnp.argsort / lax.sort_key_val: stable, -0 == +0, NaN last (jax/_src/lax/lax.py sort keys)."""
    return np.argsort(np.asarray(x, np.float64), kind="stable")


def count_crossed_edges(sender_node_idx, log_L, num_samples: Optional[int] = None):
    """internals/tree_structure.py:33-108."""
    sender = np.asarray(sender_node_idx, np.int64)
    log_L = np.asarray(log_L, np.float64)
    N = sender.size
    nodes = np.concatenate([[-np.inf], log_L])
    sort_idx = stable_argsort(nodes)
    out_degree = np.zeros(N + 1, np.int32)
    np.add.at(out_degree, np.maximum(sender, 0), np.int32(1))
    delta = (out_degree[sort_idx] - np.int32(1)).astype(np.int32)
    crossed = (np.int32(1) + np.cumsum(delta, dtype=np.int32)).astype(np.int32)
    if num_samples is not None:
        fake = np.int32(N - num_samples)
        filled = np.full(N + 1, fake, np.int32)
        filled[:num_samples] = crossed[:num_samples]
        crossed = filled - fake
    return (sort_idx[1:] - 1).astype(np.int64), crossed[:-1].astype(np.int32)


def count_intervals_naive(sender_node_idx, log_L):
    """internals/tree_structure.py:136-154 (O(N^2) definition of the live-point count)."""
    sender = np.asarray(sender_node_idx, np.int64)
    log_L = np.asarray(log_L, np.float64)
    nodes = np.concatenate([[-np.inf], log_L])
    cons = nodes[sender]
    sort_idx = stable_argsort(log_L)
    N = sender.size
    avail = np.ones(N, bool)
    contour = nodes[0]
    out = np.zeros(N, np.int32)
    for i in range(N):
        mask = (cons[sort_idx] <= contour) & (log_L[sort_idx] > contour) & avail[sort_idx]
        out[i] = mask.sum()
        contour = nodes[sort_idx[i] + 1]
        avail[sort_idx[i]] = False
    return sort_idx, out


EV_FIELDS = ("log_L", "log_X_mean", "log_X2_mean", "log_Z_mean", "log_ZX_mean", "log_Z2_mean", "log_dZ_mean",
             "log_dZ2_mean")


def init_evidence_calc() -> np.ndarray:
    """create_init_evidence_calc (internals/shrinkage_statistics.py:112-128)."""
    return np.array([-np.inf, 0.0, 0.0, -np.inf, -np.inf, -np.inf, -np.inf, -np.inf])


def evidence_scan(state, log_L, num_live, per_sample: bool = False):
    """cumulative_op_static(_update_evidence_calc_op) (internals/shrinkage_statistics.py:43-94,131-157)."""
    st = np.array(state, np.float64)
    log_L = np.ascontiguousarray(log_L, np.float64)
    n = np.ascontiguousarray(num_live, np.float64)
    M = log_L.size
    per = np.empty((M, 8)) if per_sample else None
    lib().o_evidence_scan(_p(st, _f64p), _p(log_L, _f64p), _p(n, _f64p), ctypes.c_int64(M),
                          _p(per, _f64p) if per_sample else None)
    return (st, per) if per_sample else st


def sample_evidence(key, num_live, log_L, S: int = 100) -> np.ndarray:
    """utils.py:433-476 (serial scan per simulation, exactly as the reference's cumulative_op_static)."""
    nl = np.ascontiguousarray(num_live, dtype=np.float64)
    ll = np.ascontiguousarray(log_L, dtype=np.float64)
    out = np.empty(S, dtype=np.float64)
    k = _key(key)
    lib().o_sample_evidence(_p(k, _u32p), _p(nl, _f64p), _p(ll, _f64p), ctypes.c_int64(ll.size), ctypes.c_int64(S),
                            _p(out, _f64p))
    return out


def linear_to_log_stats(log_f_mean, log_f2_mean):
    """internals/stats.py:55-74."""
    mu = 2.0 * log_f_mean - 0.5 * log_f2_mean
    sigma2 = log_f2_mean - 2.0 * log_f_mean
    return mu, max(sigma2, np.finfo(np.float64).eps)


def ess_kish(log_Z_mean, log_dZ2_mean):
    """internals/stats.py:77-86."""
    return float(np.exp(2.0 * log_Z_mean - log_dZ2_mean))


# ----------------------------------------------------------------------------------------------
# Termination (nested_samplers/common/termination.py:13-147) and the static loop
# ----------------------------------------------------------------------------------------------
class TermCond(NamedTuple):
    ess: Optional[float] = None
    evidence_uncert: Optional[float] = None
    live_evidence_frac: Optional[float] = None
    dlogZ: Optional[float] = None
    max_samples: Optional[float] = None
    max_num_likelihood_evaluations: Optional[float] = None
    log_L_contour: Optional[float] = None
    efficiency_threshold: Optional[float] = None
    rtol: Optional[float] = None
    atol: Optional[float] = None
    peak_XL_frac: Optional[float] = None


def init_register() -> dict:
    """create_init_termination_register (nested_samplers/common/initialisation.py:87-108)."""
    return dict(num_samples_used=0, evidence_calc=init_evidence_calc(),
                evidence_calc_with_remaining=init_evidence_calc(), num_likelihood_evaluations=0,
                log_L_contour=-np.inf, efficiency=0.0, plateau=False, no_seed_points=False,
                relative_spread=np.inf, absolute_spread=np.inf, peak_log_XL=-np.inf)


def determine_termination(tc: TermCond, reg: dict):
    reason = 0
    done = False
    with np.errstate(all="ignore"):
        ec = reg["evidence_calc"]
        ecr = reg["evidence_calc_with_remaining"]

        def setbit(b, bit):
            nonlocal done, reason
            if b:
                done = True
                reason += 2 ** bit

        if tc.max_samples is not None:
            setbit(reg["num_samples_used"] >= tc.max_samples, 0)
        if tc.evidence_uncert is not None:
            _, v = linear_to_log_stats(ecr[3], ecr[5])
            setbit(v <= tc.evidence_uncert ** 2, 1)
        if tc.dlogZ is not None:
            m1, _ = linear_to_log_stats(ecr[3], ecr[5])
            m0, _ = linear_to_log_stats(ec[3], ec[5])
            setbit(np.float64(m1) - np.float64(m0) < tc.dlogZ, 2)
        if tc.ess is not None:
            setbit(ess_kish(ecr[3], ecr[7]) >= tc.ess, 3)
        if tc.max_num_likelihood_evaluations is not None:
            setbit(reg["num_likelihood_evaluations"] >= tc.max_num_likelihood_evaluations, 4)
        if tc.log_L_contour is not None:
            setbit(reg["log_L_contour"] >= tc.log_L_contour, 5)
        if tc.efficiency_threshold is not None:
            setbit(reg["efficiency"] < tc.efficiency_threshold, 6)
        setbit(reg["plateau"], 7)
        if tc.rtol is not None:
            setbit(reg["relative_spread"] < tc.rtol, 8)
        if tc.atol is not None:
            setbit(reg["absolute_spread"] < tc.atol, 9)
        setbit(reg["no_seed_points"], 10)
        if tc.peak_XL_frac is not None:
            log_XL = ec[1] + ec[0]
            setbit(log_XL < reg["peak_log_XL"] + np.log(tc.peak_XL_frac), 11)
    return done, reason


def round_up_num_live_points(n, shell_frac, num_devices):
    """nested_samplers/sharded/sharded_static.py:577-584."""
    n = int(n)
    while int(n * shell_frac) % num_devices != 0:
        n += 1
    return n


def round_up_max_samples(max_samples, num_discard, num_phantom):
    """nested_samplers/sharded/sharded_static.py:587-594."""
    max_samples = int(max_samples)
    block = num_discard * (1 + num_phantom)
    while max_samples % block != 0:
        max_samples += 1
    return max_samples


class OracleNestedSampler:
    """ShardedStaticNestedSampler._run / _to_results (nested_samplers/sharded/sharded_static.py:597-851)
    with UniDimSliceSampler(perfect=True)."""

    def __init__(self, model: OModel, num_live_points: int, num_slices: int, num_phantom: int = 0,
                 midpoint_shrink: bool = True, max_samples: Optional[int] = None, shell_fraction: float = 0.5,
                 num_devices: int = 1, intended_sender: bool = False):
        self.model = model
        self.shell_fraction = max(shell_fraction, 1.0 / num_live_points)
        self.N = round_up_num_live_points(num_live_points, self.shell_fraction, num_devices)
        self.S = int(num_slices)
        self.k = int(num_phantom)
        self.midpoint = bool(midpoint_shrink)
        if max_samples is None:
            max_samples = self.N * 100
        self.m = int(self.N * self.shell_fraction)
        self.max_samples = round_up_max_samples(max_samples, self.m, self.k)
        self.intended_sender = intended_sender
        self.ctab = seed_table(self.N)
        self.iterations = 0

    # -- state helpers ----------------------------------------------------------------------
    def _append(self, st, sender, U, logL, nev, phantom):
        """_add_samples_to_state (:40-85) incl. the clamped dynamic_update_slice (internals/maps.py:15-25)."""
        n = logL.shape[0]
        cap = self.max_samples
        start = min(max(st["next_sample_idx"], 0), cap - n)
        sl = slice(start, start + n)
        st["sender"][sl] = sender
        st["log_L"][sl] = logL
        st["U"][sl] = U
        st["n_evals"][sl] = nev
        st["phantom"][sl] = phantom
        st["next_sample_idx"] = (st["next_sample_idx"] + n) % cap
        st["num_samples"] += n

    def run(self, key, term_cond: Optional[TermCond] = None, max_iterations: Optional[int] = None):
        model, N, m, D, k = self.model, self.N, self.m, self.model.D, self.k
        cap = self.max_samples
        if term_cond is None:
            term_cond = TermCond(dlogZ=float(np.log(1.0 + 1e-3)), max_samples=float(cap))
        # create_init_state (common/initialisation.py:20-84)
        st = dict(sender=np.zeros(cap, np.int64), log_L=np.full(cap, np.inf), U=np.zeros((cap, D)),
                  n_evals=np.zeros(cap, np.int64), phantom=np.zeros(cap, bool), next_sample_idx=0, num_samples=0)
        key1, sample_key = split(key, 2)
        U, logL, nev = init_batch(model, sample_key, N)
        order = stable_argsort(logL)
        live = dict(sender=np.zeros(N, np.int64), U=U[order], log_L=logL[order],
                    log_L_constraint=np.full(N, -np.inf), n_evals=nev[order])
        st["key"] = key1
        reg = init_register()
        # the uniform phase (:793-815) runs zero iterations: efficiency=0.0 < 0.1 at the first cond
        # _main_ns_thread (:427-574)
        space = m * (1 + k)
        if term_cond.max_samples is not None:
            term_cond = term_cond._replace(max_samples=min(term_cond.max_samples, cap - space))
        reg["no_seed_points"] = bool(live["log_L"][m - 1] >= live["log_L"][-1])
        self.iterations = 0
        while True:
            done, reason = determine_termination(term_cond, reg)
            if done or (max_iterations is not None and self.iterations >= max_iterations):
                break
            st["key"], _eph = split(st["key"], 2)  # :491
            sampler_state = (live["U"].copy(), live["log_L"].copy())  # pre_process: pre-discard live set
            live, reg = self._collect_shell(live, st, reg, sampler_state)
            st["key"], _eph = split(st["key"], 2)  # :510
            self.iterations += 1
        _, reason = determine_termination(term_cond, reg)
        self._append(st, live["sender"], live["U"], live["log_L"], live["n_evals"], False)  # :834-838
        self.register = reg
        self.live = live
        return reason, st

    def _collect_shell(self, live, st, reg, sampler_state):
        """_collect_shell (:210-324)."""
        model, N, m, k = self.model, self.N, self.m, self.k
        disc_logL = live["log_L"][:m].copy()
        self._append(st, live["sender"][:m], live["U"][:m], live["log_L"][:m], live["n_evals"][:m], False)
        st["key"], sample_key = split(st["key"], 2)  # :248
        contour = live["log_L"][m - 1]
        new = slice_batch(model, sample_key, contour, sampler_state[0], sampler_state[1], self.S, k,
                          self.midpoint, num_samples=m, ctab=self.ctab)
        sender = st["next_sample_idx"] - 1 if not self.intended_sender else st["next_sample_idx"]  # :261 (F5)
        live["sender"][:m] = sender
        live["U"][:m] = new["U"]
        live["log_L"][:m] = new["log_L"]
        live["log_L_constraint"][:m] = contour
        live["n_evals"][:m] = new["n_evals"]
        order = stable_argsort(live["log_L"])  # :274
        live = {kk: v[order] for kk, v in live.items()}
        if k > 0:  # add_phantom_samples_to_state (:181-207)
            self._append(st, np.full(m * k, sender, np.int64), new["ph_U"], new["ph_log_L"],
                         np.zeros(m * k, np.int64), True)
        ec = evidence_scan(reg["evidence_calc"], disc_logL, np.full(m, float(N)))  # :284-291
        ecr = evidence_scan(ec, live["log_L"], np.arange(float(N), 0.0, -1.0))  # :292-299
        with np.errstate(all="ignore"):
            absolute_spread = abs(live["log_L"][-1] - live["log_L"][0])
            relative_spread = 2.0 * absolute_spread / abs(live["log_L"][0] + live["log_L"][-1])
        reg = dict(num_samples_used=st["num_samples"], evidence_calc=ec, evidence_calc_with_remaining=ecr,
                   num_likelihood_evaluations=reg["num_likelihood_evaluations"] + int(new["n_evals"].sum()),
                   log_L_contour=contour, efficiency=N / float(live["n_evals"].sum()),
                   plateau=bool(np.all(live["log_L"] == live["log_L"][0])),
                   no_seed_points=bool(live["log_L"][m - 1] >= live["log_L"][-1]),
                   relative_spread=relative_spread, absolute_spread=absolute_spread,
                   peak_log_XL=max(reg["peak_log_XL"], ec[1] + ec[0]))
        return live, reg

    def to_results(self, reason, st) -> dict:
        """_to_results (:652-773), trim=True; scalar/evidence part (transforms are model-specific)."""
        num_samples = min(st["num_samples"], self.max_samples)
        sender = st["sender"][:num_samples]
        log_L = st["log_L"][:num_samples]
        idx, n_live = count_crossed_edges(sender, log_L)
        log_L_s = log_L[idx]
        final, per = evidence_scan(init_evidence_calc(), log_L_s, n_live.astype(np.float64), per_sample=True)
        log_Z_mean, log_Z_var = linear_to_log_stats(final[3], final[5])
        log_Z_uncert = np.sqrt(log_Z_var)
        total_phantom = int(st["phantom"][:num_samples].sum())
        f = total_phantom / num_samples
        keff = f / (1.0 - f)
        log_Z_uncert *= np.sqrt(1.0 + keff)
        ESS = ess_kish(final[3], final[7]) / (1.0 + keff)
        log_dZ = per[:, 6]
        with np.errstate(all="ignore"):
            mx = np.max(log_dZ)
            norm = mx + np.log(np.sum(np.exp(log_dZ - mx)))
            log_dp = log_dZ - norm
            a = log_dp + np.log(np.abs(np.where(np.isneginf(log_dp), 0.0, log_L_s)))
            sgn = np.sign(np.where(np.isneginf(log_dp), 0.0, log_L_s))
            amx = np.max(a)
            H_instable = -((np.sum(sgn * np.exp(a - amx)) * np.exp(amx)) - log_Z_mean)
            b = log_dp + np.log(-per[:, 1])
            bmx = np.max(b[np.isfinite(b)]) if np.any(np.isfinite(b)) else 0.0
            H_stable = -(np.exp(bmx) * np.sum(np.exp(b - bmx)))
        H = H_instable if np.isfinite(H_instable) else H_stable
        nev = st["n_evals"][:num_samples][idx]
        return dict(log_Z_mean=log_Z_mean, log_Z_uncert=log_Z_uncert, ESS=ESS, H_mean=H,
                    total_num_samples=num_samples, total_phantom_samples=total_phantom,
                    log_L_samples=log_L_s, U_samples=st["U"][:num_samples][idx], log_dp_mean=log_dp,
                    log_X_mean=per[:, 1], num_live_points_per_sample=n_live,
                    num_likelihood_evaluations_per_sample=nev, total_num_likelihood_evaluations=int(nev.sum()),
                    log_efficiency=np.log(num_samples) - np.log(float(nev.sum())), termination_reason=reason,
                    samples_indices=idx, per_sample_evidence=per, final_evidence=final)


def num_threads() -> int:
    return lib().o_num_threads()


def set_num_threads(n: int):
    lib().o_set_num_threads(ctypes.c_int(int(n)))


def set_fixed_alpha(a: float):
    """Test knob: constant midpoint-shrink factor instead of linspace(0.5, 1, S) (a < 0 restores the schedule)."""
    lib().o_set_fixed_alpha(ctypes.c_double(float(a)))


def set_gradient_flags(flags: int):
    """Test knob: bit 0 = gradient_slice, bit 1 = gradient_guided chains (uni_slice_sampler.py:202-214, :255-269)
    in slice_batch / OracleNestedSampler; 0 restores the plain sampler."""
    lib().o_set_gradient_flags(ctypes.c_int(int(flags)))


def grad_U(model: "OModel", U) -> np.ndarray:
    """Analytic d log L / dU of the registered families at U [n, D] (ns_oracle.c o_grad_U)."""
    U = np.ascontiguousarray(np.atleast_2d(U), np.float64)
    g = np.empty_like(U)
    for i in range(U.shape[0]):
        ui, gi = np.ascontiguousarray(U[i]), np.empty(U.shape[1])
        lib().o_grad_U(model.c, _p(ui, _f64p), _p(gi, _f64p))
        g[i] = gi
    return g
