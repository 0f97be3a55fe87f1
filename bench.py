#!/usr/bin/env python
"""
bench.py -- likelihood evals/sec and time-to-logZ of the static nested-sampling hot path.

A "step" is one whole nested-sampling run to termination (default dlogZ = log(1+1e-3)) on the
configuration BASELINE.json's metric is quoted on: the 32-D correlated Gaussian (dense covariance,
analytic log Z = -141.429218), num_live_points = 3200, defaults s=5, k=0 (configs[1]); step i uses
PRNGKey(i).  With --gpus N > 1 (one process per GPU under torchrun) the chains of every iteration
are sharded over the ranks and all-gathered over NCCL; per-GPU work is held fixed (num_live_points =
3200 * N), so scaling is "weak".

  value         whole-job likelihood evals/s, model parameters resident in HBM, device-timed
                (CUDA events on the launching stream around each run, max over ranks)
  e2e           the same metric through the public API (Model -> NestedSampler -> to_results) with
                host buffers: model parameters copied host->device and the posterior samples /
                weights read back device->host inside the timed region (N > 1: every rank builds its
                model and runs its shard of the chains, rank 0 post-processes and reads back the one result)
  roofline      fused slice kernel: algorithmic FP64 flops (evals x (D^2 + 4D)) / its
                CUDA-event time inside the runs, against an FP64-FMA peak measured in the same process
  cpu_baseline  the oracle (CPU restatement of jaxns 2.6.9) on a bounded sample of the same workload

--impl reference times the reference's CPU algorithm (the oracle port; the reference itself needs
JAX/TFP, which cannot be installed offline -- see DESIGN.md) with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D = 32
BASE_LIVE = 3200
ANALYTIC_LOGZ = -141.4292184
FLOPS_PER_EVAL = D * D + 4 * D  # triangular matvec D(D+1) + subtraction/dot 3D


def workload_arrays():
    cov = np.full((D, D), 0.99) + 0.01 * np.eye(D)
    return np.zeros(D), np.ones(D), np.full(D, 15.0), cov


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index
        self.first = 0

    def wait_ready(self, timeout=5.0):
        """Block until nvidia-smi has printed its first sample (its start-up is over)."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        """Start of the timed region: only samples taken from here on are reported.  nvidia-smi itself is started
        BEFORE the warm-up runs: its start-up (NVML initialisation over every GPU of the node, hundreds of ms of driver
        calls) otherwise lands inside the first timed runs and delays their kernel launches -- measured as 88 vs 97-102
        ms per run from one process to the next on the same box."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("NSB200_BENCH_SMI_MS", "100"), "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[self.first:] or self.rows
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on a bounded sample
# --------------------------------------------------------------------------------------------------
def oracle_sample(num_live, iterations, seed=0, threads=None):
    """init + the first `iterations` shells of the same workload; returns (evals, seconds, threads)."""
    from oracle import oracle as o
    o.set_num_threads(threads or os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1
    om = o.gauss_model(D)
    ns = o.OracleNestedSampler(om, num_live, D * 5, 0, True, max_samples=num_live * 100)
    t0 = time.perf_counter()
    reason, st = ns.run(o.PRNGKey(seed), max_iterations=iterations)
    dt = time.perf_counter() - t0
    n = min(st["num_samples"], ns.max_samples)
    evals = int(st["n_evals"][:n].sum())
    return evals, dt, o.num_threads()


def try_real_jaxns(num_live, seed):
    """Plan A: the unmodified reference from baseline/_ref on the JAX CPU backend, if jax
    and tfp happen to exist on this box.  Returns (evals, seconds, cores) or None."""
    try:
        os.environ.setdefault("XLA_FLAGS", f"--xla_force_host_platform_device_count={os.cpu_count()}")
        os.environ.setdefault("JAX_PLATFORMS", "cpu")
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        import jax  # noqa: F401
        import tensorflow_probability.substrates.jax as tfp
        from jax import numpy as jnp, random
        from jaxns import Model, NestedSampler, Prior
    except Exception:
        return None
    tfpd = tfp.distributions
    p_loc, p_scale, mu, cov = workload_arrays()

    def prior_model():
        x = yield Prior(tfpd.MultivariateNormalTriL(loc=jnp.asarray(p_loc), scale_tril=jnp.diag(jnp.asarray(p_scale))))
        return x

    def log_likelihood(x):
        return tfpd.MultivariateNormalTriL(loc=jnp.asarray(mu), scale_tril=jnp.linalg.cholesky(jnp.asarray(cov))).log_prob(x)

    model = Model(prior_model=prior_model, log_likelihood=log_likelihood)
    ns = NestedSampler(model=model, num_live_points=num_live)
    run = jax.jit(lambda key: ns(key)).lower(random.PRNGKey(0)).compile()
    t0 = time.perf_counter()
    reason, state = run(random.PRNGKey(seed))
    reason.block_until_ready()
    dt = time.perf_counter() - t0
    res = ns.to_results(reason, state)
    return int(res.total_num_likelihood_evaluations), dt, os.cpu_count()


def run_reference(args, emit=print):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    num_live = BASE_LIVE * args.gpus
    real = try_real_jaxns(num_live, 0) if os.environ.get("NSB200_TRY_JAXNS", "1") == "1" else None
    if real is not None:
        tot_e, tot_t = 0, 0.0
        for s in range(args.steps):
            e, t, cores = try_real_jaxns(num_live, s)
            tot_e += e
            tot_t += t
        value = tot_e / tot_t
        emit(json.dumps({
            "impl": "reference", "metric": "likelihood_evals_per_sec", "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"32-D correlated Gaussian, num_live_points={num_live}, s=5, k=0 (BASELINE configs[1])"},
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "reference",
                             "sample": "whole runs, jaxns 2.6.9 on the JAX CPU backend"},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    # Bounded sample: whole runs to dlogZ when K of them fit ~3 minutes on this host, otherwise the prior draws + the
    # first `iters` shells of every run (evals/s is flat over a run: the chains do the same work per slice).
    t0 = time.perf_counter()
    e_w, t_w, cores = oracle_sample(num_live, 4)
    rate = e_w / t_w
    for _ in range(max(0, args.warmup - 1)):
        oracle_sample(num_live, 1)
    full_evals = 1.32e8 * args.gpus  # evaluations of one whole run of this workload (bench line of the native arm)
    budget = 180.0
    if args.steps * full_evals / rate <= budget:
        iters, sample = None, "whole runs to dlogZ=log(1+1e-3)"
    else:
        per_shell = full_evals / 136.0
        iters = int(max(4, min(136, budget / args.steps * rate / per_shell)))
        sample = f"prior draws + first {iters} shells of the run per step (whole run: ~136 shells)"
    tot_e, tot_t = 0, 0.0
    for s in range(args.steps):
        e, t, cores = oracle_sample(num_live, iters, seed=s)
        tot_e += e
        tot_t += t
    value = tot_e / tot_t
    line = {
        "impl": "reference", "metric": "likelihood_evals_per_sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"32-D correlated Gaussian (dense cov, rho=0.99, mu=15), num_live_points={num_live}, "
                               "num_slices=160, k=0, run to dlogZ=log(1+1e-3) (BASELINE configs[1])",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"{sample}, {args.steps} steps"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of jaxns 2.6.9 (reference needs jax/tfp: not installable offline)",
    }
    emit(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------------
def run_native(args, emit=print):
    import torch
    import torch.distributed as dist
    import jaxns_b200 as j
    from jaxns_b200 import _lib, distributions as tfpd, likelihoods as lk, random

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    num_live = BASE_LIVE * world
    p_loc, p_scale, mu, cov = workload_arrays()

    def make_model():
        def prior_model():
            x = yield j.Prior(tfpd.MultivariateNormalTriL(loc=p_loc, scale_tril=np.diag(p_scale)), name="x")
            return x

        return j.Model(prior_model, lk.DenseGaussianLikelihood(mu, covariance_matrix=cov))

    model = make_model()
    ns = j.NestedSampler(model=model, num_live_points=num_live)
    assert ns.num_slices == 160 and ns.k == 0
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_run(seed):
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        flush.fill_(float(seed))  # L2 flush between timed iterations (outside the timed region)
        barrier()
        ev0.record()
        reason, state = ns(random.PRNGKey(seed))
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        n = min(state.num_samples, ns.nested_sampler.max_samples)
        evals = int(state.sample_collection.num_likelihood_evaluations[:n].sum().item())
        prof = dict(ns.nested_sampler.last_profile)
        reg = ns.nested_sampler.last_register
        return ms, evals, prof, reason, state, int(reg.num_likelihood_evaluations)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        clocks.wait_ready()
    for w in range(args.warmup):
        one_run(1000 + w)
    clocks.mark()
    tot_ms, tot_evals, slice_ms, slice_evals, launches, iters = 0.0, 0, 0.0, 0, 0, 0
    logZ = []
    barrier()
    for s in range(args.steps):
        ms, evals, prof, reason, state, loop_evals = one_run(s)
        tot_ms += ms
        tot_evals += evals
        slice_ms += prof["slice_ms"]
        slice_evals += loop_evals // world  # this rank's share of the chains
        launches += prof["all_launches"]
        iters += prof["iterations"]
        if rank == 0 and s < 3:
            res = ns.to_results(reason, state)
            logZ.append((res.log_Z_mean, res.log_Z_uncert))
    barrier()
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([tot_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot_ms_max = float(t.item())
    value = tot_evals / (tot_ms_max * 1e-3)

    # ---- strong scaling (N > 1): the headline workload itself (num_live_points = 3200) sharded over the ranks ----
    strong = None
    if world > 1:
        ns_weak = ns
        ns = j.NestedSampler(model=model, num_live_points=BASE_LIVE)
        one_run(2000)
        s_ms, s_evals, s_runs = 0.0, 0, max(1, min(args.steps, 3))
        barrier()
        for s_i in range(s_runs):
            ms, evals, prof, reason, state, loop_evals = one_run(s_i)
            s_ms += ms
            s_evals += evals
        ts = torch.tensor([s_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        strong = {"num_live_points": BASE_LIVE, "runs": s_runs, "time_to_logZ_ms": float(ts.item()) / s_runs,
                  "evals_per_sec": s_evals / (float(ts.item()) * 1e-3),
                  "note": "fixed total work (BASELINE configs[1]) sharded over the ranks: 1600 / n_gpus chains per GPU and body"}
        del state, reason
        ns = ns_weak

    # ---- e2e through the public API with host buffers --------------------------------------------
    e2e_t, e2e_evals, h2d, d2h = 0.0, 0, 0, 0
    n_e2e = max(1, min(args.steps, 3))
    cap = ns.nested_sampler.max_samples if rank == 0 else 1
    pinned = {"log_L": torch.empty(cap, dtype=torch.float64).pin_memory(),
              "log_dp": torch.empty(cap, dtype=torch.float64).pin_memory(),
              "x": torch.empty((cap, D), dtype=torch.float64).pin_memory()}
    for s in range(-1, n_e2e):  # s = -1: untimed warm-up of the public-API path (lazy CUDA module loads, allocator)
        flush.fill_(0.5)
        barrier()
        t0 = time.perf_counter()
        m2 = make_model()  # host numpy -> device copies happen inside (Model.desc)
        ns2 = j.NestedSampler(model=m2, num_live_points=num_live)
        t_build = time.perf_counter()
        reason, state = ns2(random.PRNGKey(max(s, 0)))
        t_run = time.perf_counter()
        # The job has ONE result: rank 0 post-processes it and reads it back (the dead-point store is replicated, the
        # other ranks would only repeat the same work and the same PCIe traffic); the other ranks finish the run.
        if rank == 0:
            res = ns2.to_results(reason, state)
            t_res = time.perf_counter()
            nres = res.total_num_samples  # posterior samples + weights into the user's pinned host buffers
            host = {"log_L": pinned["log_L"][:nres], "log_dp": pinned["log_dp"][:nres], "x": pinned["x"][:nres]}
            host["log_L"].copy_(res.log_L_samples, non_blocking=True)
            host["log_dp"].copy_(res.log_dp_mean, non_blocking=True)
            host["x"].copy_(res.samples["x"], non_blocking=True)
            host["logZ"] = res.log_Z_mean
            e2e_evals_run = res.total_num_likelihood_evaluations
        else:
            t_res = time.perf_counter()
            host = {}
            e2e_evals_run = 0
        torch.cuda.synchronize()
        if os.environ.get("NSB200_BENCH_VERBOSE"):
            print(f"[e2e rank {rank} rep {s}] build {1e3 * (t_build - t0):.1f} ms | run {1e3 * (t_run - t_build):.1f} | "
                  f"to_results {1e3 * (t_res - t_run):.1f} | d2h {1e3 * (time.perf_counter() - t_res):.1f}", file=sys.stderr)
        if s >= 0:
            e2e_t += time.perf_counter() - t0
            e2e_evals += e2e_evals_run
        fam, D_, pk, K, a, b, params = m2.host_arrays()
        h2d = int(a.nbytes + b.nbytes + params.nbytes + 8)
        d2h = int(sum(v.numel() * v.element_size() for v in host.values() if hasattr(v, "numel")) + 8 * 8)
        # release this rep's engine (its device arena is freed with it) OUTSIDE the next timed region: `state` holds
        # zero-copy views of engine memory, so without this the engine would die when `state` is rebound at the end of
        # the next run -- a cudaFree of the arena between NCCL collectives, seen as 100-600 ms stalls at N > 1
        del ns2, m2, reason, state, host
        if rank == 0:
            del res
        import gc
        gc.collect()
        torch.cuda.synchronize()
    te = torch.tensor([e2e_t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = e2e_evals / float(te.item())  # rank 0's count of the job's evaluations / slowest rank's wall time

    if rank == 0:
        # ---- roofline of the dominant kernel ----------------------------------------------------
        tf = _lib.ctypes.c_double()
        _lib.check(_lib.lib().nsb200_bench_fp64_fma(_lib.ctypes.c_int64(1 << 15), _lib.ctypes.byref(tf)))
        achieved = slice_evals * FLOPS_PER_EVAL / (slice_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        roofline = {"bound": "fp64", "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s",
                    "frac": achieved / tf.value if tf.value else None,
                    # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full,
                    # profiles/r2/slice_r2_final.txt: 84.35 MB + 3.16 MB): the pre-generated chain streams (directions,
                    # uniforms, keys), read once; algorithmic bytes ~0.4 MB
                    "traffic": 87.5e6,
                    "kernel": "k_slice_chains<32,1,P> (fused slice chains)",
                    "kernel_share_of_step": slice_ms / tot_ms,
                    "peak_source": "FP64 FMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 "
                                   f"figure; its hbm_gbs={peaks.get('hbm_gbs')} governs only the statistics kernels)",
                    "algorithmic_flops_per_eval": FLOPS_PER_EVAL}
        # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ------------------------------------
        cpu_iters = 40
        if world == 1:
            ce, ct, cores = oracle_sample(num_live, cpu_iters, seed=0)
            cpu_baseline = {"value": ce / ct, "unit": "evals/s", "cores": cores, "kind": "port",
                            "sample": f"oracle (CPU restatement of jaxns 2.6.9): prior draws + first {cpu_iters} "
                                      f"shells of the same run, {ce} evals in {ct:.1f}s"}
        else:
            cpu_baseline = None
        line = {
            "metric": "likelihood_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"32-D correlated Gaussian (dense cov, rho=0.99, mu=15), num_live_points={num_live}, "
                                   "num_slices=160, k=0, run to dlogZ=log(1+1e-3) (BASELINE configs[1])",
                       "l2": "256 MB buffer written between timed runs (L2 flush)",
                       "iterations_per_step": iters / args.steps, "evals_per_step": tot_evals / args.steps,
                       "time_to_logZ_ms": tot_ms_max / args.steps, "strong_scaling": strong,
                       "logZ": [{"mean": m, "uncert": u, "analytic": ANALYTIC_LOGZ} for m, u in logZ]},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "runs": n_e2e, "readback": "posterior samples, weights and log Z on rank 0 (one result per job)"},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout: everything else that libraries print there (NCCL's version banner,
    # the "Running over N devices." notice) is sent to stderr by pointing fd 1 at fd 2 for the duration of the run.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    emit = lambda text: lines.append(text)  # noqa: E731
    try:
        if args.impl == "reference":
            run_reference(args, emit)
        else:
            run_native(args, emit)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for text in lines:
        print(text, flush=True)


if __name__ == "__main__":
    main()
