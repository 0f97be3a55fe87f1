"""CPU tests of the host-side mirror of the reference interface and of the C-ABI library surface
(no compute calls: there is no GPU here)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(D=4, uniform=False):
    import jaxns_b200 as j
    from jaxns_b200 import distributions as tfpd, likelihoods as lk

    def prior_model():
        if uniform:
            x = yield j.Prior(tfpd.Uniform(low=np.zeros(D), high=np.ones(D)), name="x")
        else:
            x = yield j.Prior(tfpd.Normal(loc=np.zeros(D), scale=np.ones(D)), name="x")
        return x

    return j.Model(prior_model, lk.DenseGaussianLikelihood(np.ones(D), covariance_matrix=np.eye(D)))


def test_library_exports_every_declared_symbol():
    from jaxns_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "nsb200.h")).read()
    declared = set(re.findall(r"\b(nsb200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(L, name) is not None
    assert L.nsb200_abi_version() == 2
    assert L.nsb200_workspace_bytes(_lib.WS_ARGSORT, 1000) > 0
    assert L.nsb200_workspace_bytes(99, 1000) == -1


def test_struct_layouts_match_header():
    from jaxns_b200 import _lib
    assert ctypes.sizeof(_lib.NsModelDesc) == 48
    assert ctypes.sizeof(_lib.NsSliceParams) == 48
    assert ctypes.sizeof(_lib.NsEvidenceCalc) == 64
    assert ctypes.sizeof(_lib.NsTermCond) == 8 + 11 * 8
    assert ctypes.sizeof(_lib.NsRegister) == 8 + 64 + 64 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = _model()
    with pytest.raises(RuntimeError, match="CUDA"):
        m.forward(np.full(4, 0.5))
    import jaxns_b200 as j
    ns = j.NestedSampler(model=m, num_live_points=20)
    with pytest.raises(RuntimeError, match="CUDA"):
        ns(j.random.PRNGKey(0))


def test_round_up_helpers():
    """/root/reference/src/jaxns/nested_samplers/sharded/tests/test_sharded_static.py:4-7"""
    from jaxns_b200.nested_sampler import round_up_max_samples, round_up_num_live_points
    assert round_up_num_live_points(10, 0.5, 1) == 10
    assert round_up_num_live_points(10, 0.5, 2) == 12
    assert round_up_num_live_points(10, 0.5, 3) == 12
    assert round_up_max_samples(1000, 7, 2) == 1008


def test_nested_sampler_defaults():
    """public.py:69-107"""
    import jaxns_b200 as j
    m = _model(D=8)
    ns = j.NestedSampler(model=m)
    assert (ns.num_slices, ns.k, ns.c, ns.max_samples, ns.num_live_points) == (40, 0, 240, 24000, 240)
    ns = j.NestedSampler(model=m, difficult_model=True)
    assert (ns.num_slices, ns.k, ns.c) == (80, 0, 800)
    assert ns.nested_sampler.sampler.midpoint_shrink is False
    ns = j.NestedSampler(model=m, parameter_estimation=True)
    assert (ns.k, ns.c, ns.max_samples) == (8, 240, 240 * 9 * 100)
    ns = j.NestedSampler(model=m, num_live_points=3200, parameter_estimation=True)
    assert ns.c == 355  # int(3200 / 9)
    ns = j.NestedSampler(model=_model(32), num_live_points=3200)
    assert (ns.num_slices, ns.c, ns.max_samples) == (160, 3200, 320000)
    assert j.DefaultNestedSampler is j.NestedSampler
    with pytest.raises(ValueError):
        j.NestedSampler(model=m, k=40)
    with pytest.raises(ValueError):
        j.NestedSampler(model=m, s=0)


def test_slice_sampler_validation():
    """uni_slice_sampler.py:305-325"""
    import jaxns_b200 as j
    m = _model()
    with pytest.raises(ValueError, match="num_slices"):
        j.UniDimSliceSampler(model=m, num_slices=0, num_phantom_save=0, midpoint_shrink=True, perfect=True)
    with pytest.raises(ValueError, match="num_phantom_save"):
        j.UniDimSliceSampler(model=m, num_slices=3, num_phantom_save=3, midpoint_shrink=True, perfect=True)
    with pytest.raises(ValueError, match="perfect"):
        j.UniDimSliceSampler(model=m, num_slices=3, num_phantom_save=0, midpoint_shrink=True, perfect=False)
    with pytest.raises(NotImplementedError):
        j.UniDimSliceSampler(model=m, num_slices=3, num_phantom_save=0, midpoint_shrink=True, perfect=True,
                             adaptive_shrink=True)


def test_model_descriptor_packing():
    from jaxns_b200 import _consts
    m = _model(D=3)
    fam, D, pk, K, a, b, params = m.host_arrays()
    assert (fam, D, pk, K) == (_consts.FAM_GAUSS_DENSE, 3, _consts.PRIOR_NORMAL, 0)
    assert params.size == 1 + 3 + 9
    np.testing.assert_allclose(params[0], -1.5 * np.log(2 * np.pi))
    np.testing.assert_allclose(params[4:].reshape(3, 3), np.eye(3))
    assert m.U_ndims == 3
    mu = _model(D=2, uniform=True)
    assert mu.host_arrays()[2] == _consts.PRIOR_UNIFORM


def test_model_wraps_unregistered_likelihood_as_external():
    """A plain callable becomes an ExternalLikelihood (family EXTERNAL): the split propose / accept path."""
    import jaxns_b200 as j
    from jaxns_b200 import _consts, distributions as tfpd, likelihoods as lk

    def prior_model():
        x = yield j.Prior(tfpd.Uniform(low=0.0, high=1.0), name="x")
        y = yield j.Prior(tfpd.Uniform(low=np.zeros(3), high=np.ones(3)), name="y")
        return y, x

    m = j.Model(prior_model, lambda y, x: -(x ** 2).sum(-1) - (y ** 2).sum(-1))
    assert m.is_external and isinstance(m.log_likelihood, lk.ExternalLikelihood)
    fam, D, pk, K, a, b, params = m.host_arrays()
    assert fam == _consts.FAM_EXTERNAL == 5 and D == 4 and K == 0 and params.size == 0
    assert m._ret_slices == [(1, 4), (0, 1)]  # variables reach the callable in prior_model's return order
    with pytest.raises(TypeError):
        j.Model(prior_model, 3.0)
    # a registered family still needs its variables in order (it consumes their concatenation)
    with pytest.raises(NotImplementedError):
        j.Model(prior_model, lk.EggBoxLikelihood())


def test_termination_condition_algebra_and_host_mirror():
    """types.py:26-74, termination.py:13-147"""
    import jaxns_b200 as j
    from jaxns_b200 import termination as T
    from jaxns_b200.types import (EvidenceCalculation, TerminationConditionConjunction,
                                  TerminationConditionDisjunction, TerminationRegister)
    a = j.TerminationCondition(max_samples=100)
    b = j.TerminationCondition(dlogZ=0.1)
    assert isinstance(a & b, TerminationConditionConjunction) and isinstance(a | b, TerminationConditionDisjunction)
    ninf = -math.inf
    init = EvidenceCalculation(ninf, 0.0, 0.0, ninf, ninf, ninf, ninf, ninf)
    reg = TerminationRegister(0, init, init, 0, ninf, 0.0, False, False, math.inf, math.inf, ninf)
    # initial register: dlogZ compares nan < x -> False; efficiency 0.0 < 0.1 -> bit 6
    assert T.determine_termination(b, reg) == (False, 0)
    assert T.determine_termination(j.TerminationCondition(efficiency_threshold=0.1, dlogZ=0.0, max_samples=10), reg) == (True, 64)
    reg2 = reg._replace(num_samples_used=100, plateau=True)
    assert T.determine_termination(a, reg2) == (True, 1 + 128)
    assert T.determine_termination(a | b, reg2) == (True, 1 + 128)
    # a conjunction fires when every child does (the reference's starts from done=False and could never fire)
    assert T.determine_termination(a & a, reg2) == (True, 1 + 128)
    assert T.determine_termination(a & b, reg2) == (True, 1 + 128)  # plateau fires in every child
    assert T.determine_termination(a & b, reg._replace(num_samples_used=100)) == (False, 0)
    tc = T.to_c(j.TerminationCondition(dlogZ=0.5, max_samples=7, peak_XL_frac=0.1))
    assert tc.mask == (1 << 3) | (1 << 4) | (1 << 10) and tc.dlogZ == 0.5 and tc.max_samples == 7.0


def test_oracle_is_not_imported_by_product():
    """The oracle is test infrastructure: nothing under jaxns_b200/ may reference it."""
    pkg = os.path.join(ROOT, "jaxns_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "__init__.py" and False, f"{f} mentions the oracle"


def _fake_results(n=7, D=3):
    import torch
    from jaxns_b200.types import NestedSamplerResults
    g = torch.Generator().manual_seed(0)
    r = lambda *shape: torch.rand(*shape, generator=g, dtype=torch.float64)  # noqa: E731
    return NestedSamplerResults(
        log_Z_mean=-3.25, log_Z_uncert=0.125, ESS=41.5, H_mean=-2.5, samples={"x": r(n, D)}, parametrised_samples={},
        U_samples=r(n, D), log_L_samples=r(n), log_dp_mean=r(n), log_X_mean=r(n), log_posterior_density=r(n),
        num_live_points_per_sample=torch.arange(n, dtype=torch.int32),
        num_likelihood_evaluations_per_sample=torch.arange(n, dtype=torch.int64), total_num_samples=n,
        total_phantom_samples=0, total_num_likelihood_evaluations=123, log_efficiency=-2.0, termination_reason=4)


def test_results_wire_format_is_the_references(tmp_path):
    """save_results / load_results (utils.py:588-641): the JSON schema of internals/namedtuple_utils.py:34-103 --
    '__namedtuple__' nodes naming the reference's class path, arrays as base64 of the raw bytes with dtype and
    shape, scalars as 0-d arrays -- so a file written here loads in jaxns and vice versa."""
    import base64
    import json
    import torch
    from jaxns_b200 import utils
    res = _fake_results()
    f = str(tmp_path / "results.json")
    utils.save_results(res, f)
    doc = json.load(open(f))
    assert doc["type"] == "__namedtuple__"
    assert doc["__class__"] == "jaxns.nested_samplers.common.types.NestedSamplerResults"
    assert list(doc["__data__"].keys()) == list(res._fields)
    node = doc["__data__"]["log_L_samples"]
    assert node["type"] == "__jax_ndarray__" and node["__dtype__"] == "float64" and node["__shape__"] == [7]
    raw = np.frombuffer(base64.b64decode(node["__data__"]), dtype="float64")
    np.testing.assert_array_equal(raw, res.log_L_samples.numpy())
    assert doc["__data__"]["log_Z_mean"]["__shape__"] == [] and doc["__data__"]["log_Z_mean"]["__dtype__"] == "float64"
    assert doc["__data__"]["termination_reason"]["__dtype__"] == "int64"
    assert doc["__data__"]["num_live_points_per_sample"]["__dtype__"] == "int32"
    assert doc["__data__"]["samples"]["x"]["__shape__"] == [7, 3]
    back = utils.load_results(f, device="cpu")
    assert type(back).__name__ == "NestedSamplerResults"
    for name in res._fields:
        a, b = getattr(res, name), getattr(back, name)
        if isinstance(a, torch.Tensor):
            assert a.dtype == b.dtype and torch.equal(a, b)
        elif isinstance(a, dict):
            assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
        else:
            assert a == b
    with pytest.warns(UserWarning, match="json"):
        utils.save_results(res, str(tmp_path / "results.txt"))
    with pytest.raises(ValueError):
        utils.save_pytree({"a": 1}, f)


def test_load_results_written_by_the_reference_serialiser(tmp_path):
    """A document laid out exactly as jaxns' serialise_namedtuple writes it (namedtuple_utils.py:34-48, :90-96)."""
    import base64
    import json
    from jaxns_b200 import utils

    def arr(a, kind="__jax_ndarray__"):
        a = np.asarray(a)
        return {"type": kind, "__dtype__": str(a.dtype), "__data__": base64.b64encode(a.tobytes()).decode(), "__shape__": a.shape}

    doc = {"type": "__namedtuple__", "__class__": "jaxns.nested_samplers.common.types.TerminationCondition",
           "__data__": {"ess": arr(np.float64(100.0)), "evidence_uncert": None, "live_evidence_frac": None,
                        "dlogZ": arr(np.float64(1e-3)), "max_samples": arr(np.int64(5000)),
                        "max_num_likelihood_evaluations": None, "log_L_contour": None, "efficiency_threshold": None,
                        "rtol": None, "atol": None, "peak_XL_frac": arr(np.arange(3.0), "__ndarray__")}}
    f = str(tmp_path / "tc.json")
    json.dump(doc, open(f, "w"), indent=2)
    tc = utils.load_pytree(f, device="cpu")
    assert type(tc).__name__ == "TerminationCondition" and tc.ess == 100.0 and tc.max_samples == 5000
    assert tc.evidence_uncert is None and isinstance(tc.peak_XL_frac, np.ndarray)


def test_xla_ffi_shim_compiles(tmp_path):
    """csrc/xla_ffi_shim.cc (the jax.ffi custom-call handlers of the north star) compiles against the stand-in of
    XLA's FFI header, which checks every handler's signature against its Bind().Ctx().Arg().Ret().Attr() chain, and
    every nsb200_* function it forwards to is declared in include/nsb200.h with matching arguments."""
    import subprocess
    src = os.path.join(ROOT, "jaxns_b200", "csrc", "xla_ffi_shim.cc")
    obj = os.path.join(tmp_path, "shim.o")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-c", "-fPIC",
                           "-I" + os.path.join(ROOT, "jaxns_b200", "csrc", "ffi_stub"), "-I" + os.path.join(ROOT, "include"),
                           src, "-o", obj])
    syms = subprocess.run(["nm", "-g", "--defined-only", obj], capture_output=True, text=True, check=True).stdout
    for name in ("slice_batch", "init_batch", "forward_batch", "count_crossed_edges", "evidence_stats", "sample_evidence",
                 "split_begin", "split_accept", "split_finish"):
        assert f"nsb200_ffi_{name}" in syms
    # a handler whose signature drifts from its binding must not compile
    bad = os.path.join(tmp_path, "bad.cc")
    text = open(src).read().replace("ffi::ResultBuffer<ffi::F64> out) {", "ffi::ResultBuffer<ffi::S64> out) {")
    assert text != open(src).read()
    open(bad, "w").write(text.replace('"../../include/nsb200.h"', '"nsb200.h"'))
    rc = subprocess.run(["g++", "-std=c++17", "-c", "-fPIC", "-I" + os.path.join(ROOT, "jaxns_b200", "csrc", "ffi_stub"),
                         "-I" + os.path.join(ROOT, "include"), bad, "-o", os.path.join(tmp_path, "bad.o")],
                        capture_output=True, text=True)
    assert rc.returncode != 0 and "does not match its FFI binding" in rc.stderr


# ---------------------------------------------------------------------------------------------------
# parametrised models (framework/context.py, framework/prior.py:146-199) and the pieces of EvidenceMaximisation that
# need no GPU
# ---------------------------------------------------------------------------------------------------
def test_parametrised_model_cpu():
    import warnings
    import torch
    from scipy import stats
    import jaxns_b200 as j
    from jaxns_b200 import distributions as tfpd

    def prior_model():
        x = yield j.Prior(tfpd.Uniform(0., 1.))
        y = yield j.Prior(tfpd.Normal(x, 1.), name='y').parametrised()
        z = yield j.Prior(0., name='z').parametrised()  # a zero-size parameter (the reference's test_basic_zero_size_param)
        sigma = yield j.Prior(tfpd.Exponential(1.))
        with j.scope("lik"):
            shift = j.get_parameter("shift", (1,), init=lambda shape, dtype: np.full(shape, 0.25))
        return y + shift, z, sigma

    def log_likelihood(y, z, sigma):
        return tfpd.Normal(y, sigma).log_prob(0.) + z[:, 0] + j.get_parameter("offset", init=lambda: np.zeros(()))

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = j.Model(prior_model=prior_model, log_likelihood=log_likelihood)
    assert model.is_general and model.U_ndims == 2
    assert set(model.params) == {"y_param", "z_param", "lik.shift", "offset"}
    assert model.num_params == 3 and model.params["z_param"].numel() == 0
    assert "num_params=3" in repr(model)
    U = torch.rand(7, 2, dtype=torch.float64)
    got = model.log_likelihood_torch(U).numpy()
    x, sigma = U[:, 0].numpy(), -np.log1p(-U[:, 1].numpy())
    np.testing.assert_allclose(got, stats.norm(x + 0.25, sigma).logpdf(0.0), rtol=1e-12)  # y = median of N(x, 1) = x
    # new parameter values make a new model; gradients flow to them
    p = {k: v.clone().requires_grad_(True) for k, v in model.params.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m2 = model(params=p)
    total = m2.log_likelihood_torch(U).sum()
    g = torch.autograd.grad(total, [p["y_param"], p["offset"]])
    assert float(g[1]) == 7.0 and torch.isfinite(g[0]).all() and float(g[0].abs().sum()) > 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m3 = model(params={**model.params, "offset": torch.tensor(2.0, dtype=torch.float64)})
    np.testing.assert_allclose(m3.log_likelihood_torch(U).numpy(), got + 2.0, rtol=1e-12)
    with pytest.raises(ValueError):
        j.get_parameter("orphan", (1,), init=np.zeros)  # no context outside a model
    with pytest.raises(ValueError):
        list(iter(lambda: j.Prior(tfpd.Uniform(0., 1.)).parametrised(), None))  # unnamed priors cannot be parametrised


def test_closed_form_quantile_distributions():
    import torch
    from scipy import stats
    from jaxns_b200 import distributions as d
    U = torch.linspace(0.01, 0.99, 9, dtype=torch.float64).reshape(-1, 1)
    checks = [(d.Exponential(2.0), stats.expon(scale=0.5)), (d.HalfNormal(1.5), stats.halfnorm(scale=1.5)),
              (d.Cauchy(1.0, 2.0), stats.cauchy(1.0, 2.0)), (d.HalfCauchy(0.5, 2.0), stats.halfcauchy(0.5, 2.0)),
              (d.Laplace(1.0, 0.7), stats.laplace(1.0, 0.7)), (d.Gumbel(0.3, 1.2), stats.gumbel_r(0.3, 1.2)),
              (d.TruncatedNormal(0.5, 2.0, -1.0, 3.0), stats.truncnorm(-0.75, 1.25, loc=0.5, scale=2.0))]
    for mine, ref in checks:
        x = mine.quantile_torch(U)
        np.testing.assert_allclose(x.numpy()[:, 0], ref.ppf(U.numpy()[:, 0]), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(mine.log_prob_torch(x).numpy(), ref.logpdf(x.numpy()[:, 0]), rtol=1e-10, atol=1e-12)
    k = d.Kumaraswamy(2.0, 3.0)
    x = k.quantile_torch(U)
    np.testing.assert_allclose((1 - (1 - x ** 2) ** 3).numpy(), U.numpy(), rtol=1e-12)


def test_newton_cg_and_m_step_cpu():
    """The optimiser behind GlobalOptimisation's fine-tune and the M-step: Rosenbrock to its minimum; one M-step solve on
    synthetic samples recovers the evidence-maximising parameter exactly where the likelihood does not depend on U."""
    import warnings
    import torch
    import jaxns_b200 as j
    from jaxns_b200 import distributions as tfpd
    from jaxns_b200.experimental import EvidenceMaximisation, MStepData, newton_cg

    def rosen(z):
        return (100.0 * (z[1:] - z[:-1] ** 2) ** 2 + (1.0 - z[:-1]) ** 2).sum()

    z, f, n_cg = newton_cg(rosen, torch.tensor([-1.2, 1.0, -0.5, 0.8], dtype=torch.float64), max_iters=200)
    assert f < 1e-16 and n_cg > 0
    np.testing.assert_allclose(z.numpy(), np.ones(4), atol=1e-7)

    def prior_model():
        x = yield j.Prior(tfpd.Uniform(0., 1.), name="x")
        mu = yield j.Prior(tfpd.Normal(0., 5.), name="mu").parametrised()
        return x, mu

    def log_likelihood(x, mu):
        return tfpd.Normal(mu, 1.0).log_prob(2.0) + 0.0 * x[:, 0]

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = j.Model(prior_model=prior_model, log_likelihood=log_likelihood)
        em = EvidenceMaximisation(model=model)
        n = 64
        data = MStepData(U_samples=torch.rand(n, 1, dtype=torch.float64),
                         log_weights=torch.full((n,), -np.log(n), dtype=torch.float64))
        params, log_Z = em._m_step(None, model.params, data)
        mu = model(params=params).transform_parametrised(torch.full((1,), 0.5, dtype=torch.float64))["mu"]
    assert abs(float(mu) - 2.0) < 1e-6
    assert abs(log_Z + 0.5 * np.log(2 * np.pi)) < 1e-10


def test_bench_clock_sampler_reports_only_the_timed_region():
    """bench.py starts nvidia-smi before the warm-up runs and reports the samples taken after mark()."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    cs = bench.ClockSampler(0)

    class _Proc:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    cs.proc = _Proc()
    row = lambda mhz, cap: ["0", str(mhz), "1965", "700", "x", "Not Active", "Not Active", "Not Active", cap]  # noqa: E731
    cs.rows = [row(300, "Not Active"), row(1200, "Not Active")]  # idle / warm-up samples
    cs.mark()
    cs.rows += [row(1965, "Not Active"), row(1950, "Active"), row(1965, "Not Active")]
    out = cs.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
    empty = bench.ClockSampler(0)
    assert empty.stop()["reasons"] == ["nvidia-smi unavailable"]
