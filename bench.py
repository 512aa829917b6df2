#!/usr/bin/env python
"""Headline benchmark: bigKRLS() fit wall-seconds at N=20k, P=10 (BASELINE.json configs[2]:
eigtrunc=0.001, all derivatives), 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

One step = one full fit (kernel -> eigen -> lambda -> coefficients + vcov -> marginal effects) of
the same synthetic workload (SURVEY.md 8d generator, seed 1003).

  value  seconds per fit with the standardised inputs already resident in HBM
         (bk_fit_run_device), wall clock bracketed by barrier + synchronize, max over ranks
  e2e    seconds per fit through the public API bigKRLS(y, X) with HOST buffers: standardise,
         H2D of X/y, fit, D2H of every output field incl. the three N x N matrices (pinned
         host buffers from the library's pool)
  roofline  dominant kernel = the FP64 tensor-core (DMMA m8n8k4) GEMM of the dense->band stage of the two-stage
         tridiagonalisation (per panel: Z = A22 (V T) and the symmetric rank-128 update): useful flops
         (4 b m^2 per panel, the update counted on the lower triangle only) / summed CUDA-event duration of
         those launches on the library stream, against the FP64 DMMA peak measured live by the library's
         micro-benchmark (MEASURED_PEAKS.json only carries bf16, which an FP64 path cannot use).  When the
         one-stage reduction is taken instead (all eigenvectors wanted) the dominant kernel is its HBM-bound
         symmetric mat-vec panel kernel and the roofline is bytes against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the compiled literal restatement of the reference (oracle/krls_port.cpp, OpenBLAS,
         all host threads) on a bounded sample, scaled cubically to the metric's size

Nothing under oracle/ is on the measured CUDA path: it is executed only for `cpu_baseline`
and `--impl reference`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_FULL, P_FULL, SEED, EIGTRUNC = 20000, 10, 1003, 0.001
METRIC = "bigKRLS() fit wall-s at N=20k P=10"


def synthetic(N, P, seed):
    """SURVEY.md 8d generator (identical to oracle/krls_oracle.synthetic; duplicated here so the
    CUDA arm never imports oracle/)."""
    rng = np.random.default_rng(seed)
    X0 = rng.standard_normal((N, P))
    eps = rng.standard_normal(N)
    y0 = np.sin(X0[:, 0]) + X0[:, 1] * X0[:, 2] + 0.5 * eps
    return np.asfortranarray(X0), y0


class ClockSampler:
    """SM clocks and throttle reasons DURING the timed region (B200_PROFILING.md), read through NVML
    (pynvml) from a background thread once per second; `nvidia-smi -lms 200` from a subprocess measurably
    slowed the timed region (it perturbs the launch-bound parts of the fit)."""

    def __init__(self, device, period=1.0):
        self.device, self.period, self.rows = device, period, []
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.mx = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            self.thread = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self._stop.set()
        self.thread.join(timeout=3)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = [r[0] for r in self.rows]
        reasons = sorted({k for _, rs in self.rows for k, bit in names.items() if rs & bit})
        busy = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": float(self.mx),
                "reasons": reasons, "samples": len(sm), "source": "nvml, 1 Hz"}


def measured_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_port_sample(n_sample, threads):
    """Times the compiled literal restatement of the reference on n_sample rows of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import krls_oracle as o   # CPU baseline leg only
    import port
    X, y = synthetic(N_FULL, P_FULL, SEED)
    X, y = X[:n_sample], y[:n_sample]
    Xs, ys, *_ = o.standardize(X, y)
    t0 = time.perf_counter()
    r = port.fit(Xs, ys, eigtrunc=EIGTRUNC, threads=threads)
    return time.perf_counter() - t0, r


def fixture_parity(fit, rank, world):
    """Outside the timed region: this run's fit against the committed CPU-oracle fixture of the SAME workload
    (tests/golden/c3_N20000_P10.npz, tools/make_fixtures.py) - relative errors per field, this rank's blocks."""
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "c3_N20000_P10.npz"))
    except Exception:  # noqa: BLE001
        fit.release_device()
        return None
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(np.asarray(b))))
    ev, rev = fit["K.eigenvalues"], z["evals"]
    big = rev >= 1e-3 * rev[0]
    c0, c1 = fit["_col_range"]
    cols = z["cidx"]
    sel = (cols >= c0) & (cols < c1)
    ix = np.ix_(z["ridx"], cols[sel] - c0)
    out = {"rank": rank, "world": world,
           "eigenvalues_rel_retained": float(np.max(np.abs(ev[big] / rev[big] - 1))),
           "eigenvalues_abs_over_lambda1": float(np.max(np.abs(ev - rev)) / rev[0]),
           "lambda_rel": abs(fit["lambda"] / float(z["lambda"]) - 1),
           "lastkeeper_equal": int(fit["lastkeeper"]) == int(z["lastkeeper"]),
           "probes_equal": int(fit["_info"]["n_probes"]) == int(z["nprobe"]),
           "coeffs": rel(fit["coeffs"].reshape(-1), z["coeffs"]), "yfitted": rel(fit["yfitted"], z["yfitted"]),
           "derivatives": rel(fit["derivatives"], z["derivatives"]),
           "var_avgderivatives": rel(fit["var.avgderivatives"].reshape(-1), z["var_avgderivatives"])}
    if sel.any():
        out["vcov_c_block"] = rel(fit["vcov.est.c"][ix], z["Vc_blk"][:, sel])
        out["vcov_fitted_block"] = rel(fit["vcov.est.fitted"][ix], z["Vf_blk"][:, sel])
        out["K_block"] = rel(fit["K"][ix], z["K_blk"][:, sel])
    out["within_tolerance"] = bool(out["eigenvalues_rel_retained"] < 1e-9 and out["lambda_rel"] < 1e-9 and
                                   out["lastkeeper_equal"] and max(out["coeffs"], out["yfitted"], out["derivatives"],
                                                                  out["var_avgderivatives"]) < 1e-8)
    fit.release_device()
    return out


def full_size_shape():
    """lastkeeper and LOO probe count of the full-size workload (from the committed oracle fixture) - the lambda
    and coefficient stages of the reference cost N^2 k per probe, so their scaling law needs both."""
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "c3_N20000_P10.npz"))
        return int(z["lastkeeper"]), int(z["nprobe"]), "tests/golden/c3_N20000_P10.npz (CPU oracle at N=20000)"
    except Exception:  # noqa: BLE001
        return None, None, "the sample's own lastkeeper and probe count (fixture missing; conservative)"


def scale_stages(st, ns, k_s, probes_s, N, P):
    """Scales the per-stage seconds measured on the first ns rows to N rows, EACH STAGE BY ITS OWN LAW (the
    reference's literal structure): kernel N^2 P; eigen (dsyevd) N^3; lambda N^2 k per probe; coefficients N^2 k;
    vcov N^2 k + 4 N^3 (K'(V K)); derivatives 4 N^3 per column."""
    k_f, probes_f, src = full_size_shape()
    if k_f is None or N != N_FULL:
        k_f, probes_f = k_s, probes_s          # conservative: lastkeeper grows (slowly) with N
    r = N / ns
    law = {"kernel": r ** 2, "eigen": r ** 3,
           "lambda": r ** 2 * (k_f / k_s) * (probes_f / max(1, probes_s)),
           "coef": r ** 2 * (k_f / k_s), "vcov": r ** 3, "deriv": r ** 3}
    out = {k: float(st[k]) * law[k] for k in law}
    return out, {k: float(v) for k, v in law.items()}, {"lastkeeper_full": k_f, "probes_full": probes_f, "source": src}


def reference_estimate(ns, threads, steps, warmup):
    """The reference's CPU path on a bounded sample + the per-stage extrapolation to the metric's size."""
    for _ in range(warmup):
        cpu_port_sample(min(ns, 1000), threads)
    ts, st, r = [], None, None
    for _ in range(max(1, steps)):
        t, r = cpu_port_sample(ns, threads)
        ts.append(t)
        st = r["times"] if st is None else {k: st[k] + r["times"][k] for k in st}
    st = {k: v / len(ts) for k, v in st.items()}
    scaled, law, shape = scale_stages(st, ns, r["lastkeeper"], r["probes"], N_FULL, P_FULL)
    # cross-check of the exponents: the same fit on 2/3 of the sample
    ns2 = int(ns * 2 / 3)
    _, r2 = cpu_port_sample(ns2, threads)
    expo = {k: float(np.log(st[k] / r2["times"][k]) / np.log(ns / ns2)) for k in scaled
            if st[k] > 0 and r2["times"][k] > 0}
    return {"sample_seconds": float(np.mean(ts)), "sample_stage_seconds": {k: float(v) for k, v in st.items()},
            "sample_lastkeeper": int(r["lastkeeper"]), "sample_probes": int(r["probes"]),
            "scaled_stage_seconds": scaled, "scale_factor_per_stage": law, "full_size_shape": shape,
            "measured_exponent_between_%d_and_%d_rows" % (ns2, ns): expo,
            "extrapolated_seconds": float(sum(scaled.values()))}


def cpu_baseline_block(est, ns, threads):
    return {"value": est["extrapolated_seconds"], "unit": "s", "cores": threads, "kind": "port",
            "extrapolated": True, "same_config": False,
            "sample": (f"first {ns} rows of the N={N_FULL} P={P_FULL} workload through the literal restatement of the "
                       f"reference (oracle/krls_port.cpp, OpenBLAS dsyevd/dgemm, {threads} threads): measured "
                       f"{est['sample_seconds']:.2f} s per fit; `value` is an EXTRAPOLATION to N={N_FULL}, each stage "
                       f"scaled by its own law (kernel N^2, eigen N^3, lambda/coefficients N^2 k probes, vcov and "
                       f"derivatives N^3) - not a measurement at N={N_FULL}"),
            "measured_sample_seconds": est["sample_seconds"], "sample_n": ns,
            "stage_seconds_on_sample": est["sample_stage_seconds"],
            "stage_seconds_scaled": est["scaled_stage_seconds"],
            "scale_factor_per_stage": est["scale_factor_per_stage"],
            "full_size_shape": est["full_size_shape"],
            "measured_exponents": est[[k for k in est if k.startswith("measured_exponent")][0]],
            "blas": "OpenBLAS (scipy wheel) dsyevd/dgemm/dgemv"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ns = args.sample_n
    est = reference_estimate(ns, threads, args.steps, args.warmup)
    sec = est["extrapolated_seconds"]
    line = {"impl": "reference", "metric": METRIC, "value": sec, "unit": "s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": est["sample_seconds"] * 1e3,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "extrapolated": True, "same_config": False,
            "config": {"workload": f"bigKRLS N={N_FULL} P={P_FULL} eigtrunc={EIGTRUNC} all derivatives (BASELINE.json configs[2])",
                       "note": f"each timed step is the first {ns} rows; value = per-stage extrapolation to N={N_FULL}; "
                               f"the measured same-size pair (GPU and CPU both at N={ns}) is `same_config_pair` of the "
                               f"CUDA arm's line"},
            "cpu_baseline": cpu_baseline_block(est, ns, threads),
            "e2e": {"value": sec, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


OTHER_CONFIGS = {
    2: dict(N=10000, P=10, seed=1002, kw=dict(eigtrunc=0.0), fixture="c2_N10000_P10",
            name="N=10k P=10 full eigendecomposition, all marginal effects (BASELINE.json configs[1])"),
    4: dict(N=60000, P=20, seed=1004, kw=dict(Neig=500, which_derivatives=[1, 3, 5]), fixture=None,
            name="N=60k P=20 Neig=500 truncated eig, which.derivatives=c(1,3,5) (BASELINE.json configs[3])"),
    5: dict(N=20000, P=10, seed=1005, kw=dict(), fixture="c5_cv_N20000_P10",
            name="crossvalidate.bigKRLS Kfolds=5 on N=20k P=10, folds sharded across the GPUs (BASELINE.json configs[4])"),
}


def run_other_config(args):
    """The other BASELINE.json configurations as bench lines (not the headline metric): the public API end to end
    with host buffers, seconds per call, max over ranks."""
    import torch
    from bigkrls_b200 import _lib, bigKRLS, crossvalidate_bigKRLS
    cfg = OTHER_CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        import torch.distributed as dist
        from bigkrls_b200.dist import TorchComm
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = TorchComm(device=f"cuda:{local}", ctx=_lib.default_context(local))
    ctx = _lib.default_context(local)
    N, P = cfg["N"], cfg["P"]
    X, y = synthetic(N, P, cfg["seed"])
    # the three N x N outputs travel to the host only when a rank's share is moderate (config 4 on one GPU would be
    # 3 x 28.8 GB of pinned host memory)
    squares = (8.0 * N * N / world) <= 8e9
    infos, parity = [], None

    def step():
        if args.config == 5:
            folds = np.random.default_rng(cfg["seed"]).permutation(N) % 5 + 1
            return crossvalidate_bigKRLS(y, X, folds=folds, comm=comm, ctx=ctx)
        fit = bigKRLS(y, X, comm=comm, ctx=ctx, pinned=True, return_squares=squares, **cfg["kw"])
        info = dict(fit["_info"])
        info["lambda"], info["lastkeeper"] = fit["lambda"], fit["lastkeeper"]
        fit.release_device()
        fit.release_pinned()
        return info

    def sync():
        torch.cuda.synchronize()
        if comm is not None:
            comm.barrier()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sync()
    launches0 = _lib.load().bk_launch_count(ctx.handle)
    sampler.start()
    t0 = time.perf_counter()
    outs = [step() for _ in range(args.steps)]
    sync()
    sec = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    launches = _lib.load().bk_launch_count(ctx.handle) - launches0
    if comm is not None:
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        sec = float(t.item())
    if rank == 0:
        last = outs[-1]
        line = {"metric": "bigKRLS wall-s, " + cfg["name"], "value": sec, "unit": "s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": cfg["name"], "seed": cfg["seed"], "N": N, "P": P,
                           "l2": "inputs larger than L2", "outputs_to_host": "K, vcov.est.c, vcov.est.fitted + vectors"
                           if squares else "vectors only (N x N fields stay on the device)"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": sec, "unit": "s", "h2d_bytes_per_step": int(8 * N * (P + 1)),
                        "d2h_bytes_per_step": int((3 * 8 * N * N / world if squares and args.config != 5 else 0) + 8 * N * (P + 3))},
                "note": "value == e2e here: these lines time the public API with host buffers only"}
        if args.config == 5:
            line["cv"] = {k: [float(v) for v in last[k]] for k in ("R2_oos", "MSE_oos") if k in last}
            try:
                z = np.load(os.path.join(ROOT, "tests", "golden", cfg["fixture"] + ".npz"))
                line["parity_vs_oracle_fixture"] = {k: float(np.max(np.abs(np.asarray(last[k]) - z[k])) / np.max(np.abs(z[k])))
                                                    for k in ("R2_is", "R2_oos", "MSE_is", "MSE_oos", "R2AME_oos", "MSE_AME_oos")}
            except Exception:  # noqa: BLE001
                pass
        else:
            line["stage_seconds"] = {k: float(last[k]) for k in ("t_kernel", "t_eigen", "t_lambda", "t_coef", "t_vcov",
                                                                  "t_deriv", "t_total")}
            line["fit"] = {"lambda": last["lambda"], "lastkeeper": int(last["lastkeeper"]),
                           "krylov_matvecs": int(last.get("krylov_matvecs", 0)),
                           "krylov_restarts": int(last.get("krylov_restarts", 0))}
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_FULL)
    ap.add_argument("--p", type=int, default=P_FULL)
    ap.add_argument("--sample-n", type=int, default=3000, help="rows of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4, 5],
                    help="BASELINE.json configs[] index + 1 of the workload (3 = the headline metric, the default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != 3:
        return run_other_config(args)

    import torch
    import ctypes as C
    from bigkrls_b200 import _lib, bigKRLS
    from bigkrls_b200._lib import FitInfo, FitOpts, check

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        import torch.distributed as dist
        from bigkrls_b200.dist import TorchComm
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = TorchComm(device=f"cuda:{local}", ctx=_lib.default_context(local))
    N, P = args.n, args.p
    lib = _lib.load()
    ctx = _lib.default_context(local)
    X, y = synthetic(N, P, SEED)
    xsd, ysd = np.std(X, axis=0, ddof=1), float(np.std(y, ddof=1))
    Xs = np.asfortranarray((X - X.mean(axis=0)) / xsd)
    ys = (y - y.mean()) / ysd
    dX = torch.from_numpy(np.ascontiguousarray(Xs.T)).cuda()     # column-major N x P == row-major P x N
    dy = torch.from_numpy(ys).cuda()
    opts = FitOpts()
    lib.bk_fit_default_opts(C.byref(opts), N, P)
    opts.eigtrunc = EIGTRUNC
    opts.y_sd = ysd
    cptr = C.byref(comm.struct) if comm is not None else None

    def sync():
        torch.cuda.synchronize()
        if comm is not None:
            comm.barrier()
            torch.cuda.synchronize()

    def resident_step():
        h = C.c_void_p()
        check(lib.bk_fit_run_device(ctx.handle, C.c_void_p(dX.data_ptr()), C.c_void_p(dy.data_ptr()), N, P,
                                    C.byref(opts), cptr, C.byref(h)))
        info = FitInfo()
        check(lib.bk_fit_get_info(h, C.byref(info)))
        lib.bk_fit_free(h)
        return info.as_dict()

    def e2e_step(pinned=True):
        fit = bigKRLS(y, X, eigtrunc=EIGTRUNC, comm=comm, pinned=pinned, ctx=ctx)
        d2h = sum(fit[k].nbytes for k in ("K", "vcov.est.c", "vcov.est.fitted", "derivatives", "coeffs", "yfitted")
                  if k in fit) + fit["K.eigenvalues"].nbytes
        fit.release_device()
        fit.release_pinned()
        return d2h

    for _ in range(args.warmup):
        resident_step()
    sampler = ClockSampler(local)
    sync()
    launches0 = lib.bk_launch_count(ctx.handle)
    sampler.start()
    t0 = time.perf_counter()
    infos = [resident_step() for _ in range(args.steps)]
    sync()
    t1 = time.perf_counter()
    clocks = sampler.stop()
    launches = lib.bk_launch_count(ctx.handle) - launches0
    sec = (t1 - t0) / args.steps

    e2e_step()                                   # warm the pinned pool
    sync()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        d2h = e2e_step()
    sync()
    e2e_sec = (time.perf_counter() - t0) / args.steps
    # the same call into plain PAGEABLE host memory (what a big.matrix is): the library's bounce-buffer copy engine
    e2e_step(pinned=False)
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step(pinned=False)
    sync()
    e2e_pageable_sec = (time.perf_counter() - t0) / args.steps

    parity = None
    if N == N_FULL and P == P_FULL:
        parity = fixture_parity(bigKRLS(y, X, eigtrunc=EIGTRUNC, comm=comm, ctx=ctx), rank, world)
        if comm is not None:
            parity = comm.gather_objects(parity)      # every rank's own column blocks
    if comm is not None:
        t = torch.tensor([sec, e2e_sec, e2e_pageable_sec], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        sec, e2e_sec, e2e_pageable_sec = t.tolist()
    if rank != 0:
        comm.close()
        torch.distributed.destroy_process_group()
        return
    info = infos[-1]
    if info.get("twostage", 0) > 0:
        r = C.c_double()
        check(lib.bk_microbench(ctx.handle, 1, 0, 0, C.byref(r)))      # FP64 DMMA issue-bound loop, TFLOP/s
        peak = float(r.value)
        ach = info["band_gemm_flops"] / info["band_gemm_seconds"] * 1e-12 if info["band_gemm_seconds"] > 0 else 0.0
        nl = max(1.0, info["band_gemm_launches"])
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_band_gemm_traffic.json")))
            traffic, traffic_src = tr["dram_bytes_per_launch"], tr["source"]
        except Exception:  # noqa: BLE001
            pass
        roofline = {"kernel": "dgemm_kernel<128,64> (FP64 DMMA GEMM: Z = A22 (V T) and A22 -= [V W][W V]' of the "
                              "dense->band stage)", "bound": "tensor", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak,
                    "peak_source": "FP64 mma.sync.m8n8k4 issue-bound loop measured live (bk_microbench kind 1); "
                                   "MEASURED_PEAKS.json holds bf16/HBM only and bf16 is not usable at the 1e-9 "
                                   "tolerance of this path",
                    "launches_per_step": int(info["band_gemm_launches"]),
                    "algorithmic_flops_per_launch": info["band_gemm_flops"] / nl,
                    "avg_launch_seconds": info["band_gemm_seconds"] / nl,
                    "share_of_step": info["band_gemm_seconds"] / info["t_total"],
                    "note": ("in the delayed-update phase of the dense->band stage the products Z = A22 (V T) (high-priority "
                             "stream) and the rank-256 updates (main stream) run at the same time: `achieved` divides their "
                             "flops by the UNION of their event intervals, launches and avg_launch_seconds refer to that busy "
                             "time; one at a time (BK_SY2SB_NOPIPE=1) the same kernels measure 26.9 TF/s = 0.72 of the peak "
                             "and the step is 0.017 s longer (profiles/r02c_bench_N20000.json)") if world == 1 else
                            "multi-GPU run: the dense->band GEMMs are distributed and not "
                            "bracketed by events; the roofline of the dominant kernel is the n_gpus=1 line",
                    "traffic": traffic, "traffic_source": traffic_src}
    else:
        peak, peak_src = measured_peaks()
        ach = info["sytrd_bytes"] / info["sytrd_kernel_seconds"] * 1e-9 if info["sytrd_kernel_seconds"] > 0 else 0.0
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_sytrd_traffic.json")))
            traffic = tr["traffic_over_algorithmic"] * info["sytrd_bytes"] / max(1.0, info["sytrd_launches"])
            traffic_src = ("dram__bytes_read+write of one ncu --set full capture (launch 40, ratio %.3f to its "
                           "algorithmic bytes) scaled to the average launch" % tr["traffic_over_algorithmic"])
        except Exception:  # noqa: BLE001
            pass
        roofline = {"kernel": "sytrd_panel_kernel (symmetric mat-vec of the tridiagonalisation)", "bound": "hbm",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                    "launches_per_step": int(info["sytrd_launches"]),
                    "algorithmic_bytes_per_launch": info["sytrd_bytes"] / max(1.0, info["sytrd_launches"]),
                    "avg_launch_seconds": info["sytrd_kernel_seconds"] / max(1.0, info["sytrd_launches"]),
                    "traffic": traffic, "traffic_source": traffic_src}
    stage = {k: float(np.mean([i[k] for i in infos])) for k in
             ("t_kernel", "t_eigen", "t_tridiag", "t_sy2sb", "t_sb2st", "t_dc", "t_backtransform", "t_q2", "t_q1",
              "t_lambda", "t_coef", "t_vcov", "t_deriv", "t_total")}
    # per-stage floors: algorithmic bytes / measured HBM peak, algorithmic flops / measured DMMA peak
    hbm_peak, _ = measured_peaks()
    rr = C.c_double()
    check(lib.bk_microbench(ctx.handle, 1, 0, 0, C.byref(rr)))
    dmma = float(rr.value) * 1e12
    kk = float(info["lastkeeper"])
    # partitioned stages: the floor of ONE rank's share (work / world).  The eigensolver's floor is left at the
    # single-GPU figure: only its dense->band stage and the back-transformation are distributed, the band->tridiagonal
    # stage and the divide & conquer run on every rank (same bits, no broadcast).
    floors = {
        "t_kernel": 8.0 * N * N / world / (hbm_peak * 1e9),
        "t_eigen": ((4.0 / 3.0) * N ** 3 + 2.0 * N * N * kk) / dmma,
        "t_lambda": info["n_passes"] * 8.0 * N * kk / world / (hbm_peak * 1e9),
        "t_vcov": max(2.0 * N * N * kk / dmma, 2 * 8.0 * N * N / (hbm_peak * 1e9)) / world,
        "t_coef+t_deriv": 8.0 * N * N / world / (hbm_peak * 1e9),   # one pass over K: yhat, derivatives, variance vectors
    }
    stage["t_coef+t_deriv"] = stage["t_coef"] + stage["t_deriv"]
    stage_roofline = {k: {"floor_s": v, "measured_s": stage[k], "frac": (v / stage[k]) if stage[k] > 0 else None}
                      for k, v in floors.items()}
    line = {"metric": METRIC, "value": sec, "unit": "s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"bigKRLS N={N} P={P} eigtrunc={EIGTRUNC} all derivatives (BASELINE.json configs[2])",
                       "seed": SEED, "l2": "inputs larger than L2 (K is %.1f GB)" % (8.0 * N * N * 1e-9),
                       "parallelism": (f"column blocks x{world}: kernel build, LOO, vcov, marginal effects partitioned; "
                                       f"dense->band stage of the eigensolver block-cyclic over the {world} GPUs "
                                       f"(peer stores over NVLink), band->tridiagonal + D&C replicated on every rank, "
                                       f"back-transformation split by columns and all-gathered") if world > 1
                       else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e_sec, "unit": "s", "h2d_bytes_per_step": int(8 * N * (P + 1)),
                    "d2h_bytes_per_step": int(d2h), "host_buffers": "pinned (library pool)"},
            "e2e_pageable": {"value": e2e_pageable_sec, "unit": "s", "h2d_bytes_per_step": int(8 * N * (P + 1)),
                             "d2h_bytes_per_step": int(d2h),
                             "host_buffers": "plain pageable numpy memory (big.matrix stand-in), filled by the "
                                             "library's bounce-buffer copy engine (csrc/hostcopy.cu)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "parity_vs_oracle_fixture": parity,
            "per_step_seconds": [float(i["t_total"]) for i in infos],
            "stage_seconds": stage,
            "stage_roofline": stage_roofline,
            "fit": {"lambda": info["lambda"], "lastkeeper": int(info["lastkeeper"]), "n_probes": info["n_probes"],
                    "n_passes": info["n_passes"], "dc_top_k": int(info["dc_top_k"])}}
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ns = args.sample_n
        est = reference_estimate(ns, threads, 1, 0)
        line["cpu_baseline"] = cpu_baseline_block(est, ns, threads)
        # the measured same-size pair: this repo's public API on the same ns rows, host buffers in and out
        Xn, yn = X[:ns], y[:ns]
        bigKRLS(yn, Xn, eigtrunc=EIGTRUNC, ctx=ctx).release_device()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bigKRLS(yn, Xn, eigtrunc=EIGTRUNC, ctx=ctx).release_device()
        torch.cuda.synchronize()
        tg = time.perf_counter() - t0
        line["same_config_pair"] = {"workload": f"first {ns} rows of the workload, eigtrunc={EIGTRUNC}, all derivatives",
                                    "gpu_e2e_s": tg, "cpu_s": est["sample_seconds"], "cpu_cores": threads,
                                    "ratio": est["sample_seconds"] / tg, "same_config": True}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
