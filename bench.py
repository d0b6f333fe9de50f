#!/usr/bin/env python
"""bench.py -- GP logLik+grad evaluations per second (fp64) on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] "C2": synthetic N=8192, D=8, cmpnd(rbf gamma=1/8, var=1; white
0.01), d=1 -- SURVEY.md 8(d).  One step = one full evaluation as SCG asks of CGp (COptimisable.cpp:309-349):
K build -> Cholesky (jitChol) -> K^-1 -> alpha -> log-likelihood terms -> hyper-parameter gradient.
  value : evaluations/s with X and m resident in HBM (only theta changes per step)
  e2e   : the same through the public call with HOST buffers: every step uploads X and m from pinned host memory
          and reads ll terms + gradient back
  roofline : the dominant kernel -- oz_gemm_kernel, the tcgen05 int8 tensor-core GEMM that carries the fp64 flops of
          potrf + inverse (N^3 per evaluation) as 8 int8 slices: algorithmic fp64 flops of its calls over their summed
          CUDA-event duration, against the fp64-equivalent peak of the int8 pipe (2 x measured bf16 peak / 36)
  cpu_baseline / --impl reference : the unmodified reference (oracle/_ref) on this box's host cores, same inputs
Multi-GPU: one C2 evaluation fits one GPU, so the headline at N GPUs is N independent evaluations (one theta candidate
per rank, what SCG restarts / line searches consume), no data-path collective, "scaling": "weak".  The path that SHARDS
(N beyond one GPU: K, its factor and K^-1 2-D block-cyclic over all ranks, NCCL panel broadcasts, gpc_dist_*) is measured
in the same run on C3 (strong scaling, against this run's own single-GPU time) and C4 (N=65536) and printed under
"sharded"; `--workload c3|c4` makes it the headline (strong scaling) for a 1/2/4/8 table of its own.
"""
import argparse
import ctypes as C
import json
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before any CUDA context exists (see gpc_b200/__init__.py)
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, D, types, natural params)
    "c2": dict(N=8192, D=8, types=["rbf", "white"], desc="C2 synthetic N=8192 D=8 rbf(gamma=1/8,var=1)+white(0.01), d=1"),
    "c3": dict(N=32768, D=16, types=["rbfard", "white"],
               desc="C3 synthetic N=32768 D=16 rbfard(gamma=1/16,var=1,s_k=0.25+0.5k/15)+white(0.01), d=1"),
    "c4": dict(N=65536, D=32, types=["matern52", "white"],
               desc="C4 synthetic N=65536 D=32 matern52(l=sqrt(32),var=1)+white(0.01), d=1"),
}


def make_inputs(name):
    """X ~ N(0,1) from default_rng(20261017); y = sin(X[:,0]) + 0.1 eps, centred."""
    w = WORKLOADS[name]
    N, D = w["N"], w["D"]
    rng = np.random.default_rng(20261017)
    X = np.asfortranarray(rng.standard_normal((N, D)))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
    y = np.asfortranarray(y - y.mean())
    if name == "c2":
        params = np.array([1.0 / D, 1.0, 0.01])
    elif name == "c4":
        params = np.array([np.sqrt(float(D)), 1.0, 0.01])
    else:
        params = np.concatenate([[1.0 / D, 1.0], 0.25 + 0.5 * np.arange(D) / (D - 1), [0.01]])
    return X, y, params


def natural_to_trans(O, w, params):
    """transformed parameters of the workload's compound kernel through the ORACLE's transforms (CTransform.cpp:25-112)"""
    kern, pos = [], 0
    for t in w["types"]:
        n = O.nparams(t, w["D"])
        kern.append((t, np.asarray(params[pos:pos + n], dtype=np.float64)))
        pos += n
    return O.trans_from_kern(kern, w["D"])


def config_of(w, world):
    """the `config` object both arms print (identical dicts: the driver compares them)"""
    N = w["N"]
    return {"workload": w["desc"], "inputs": "X ~ N(0,1) default_rng(20261017), y = sin(x_0) + 0.1 eps, centred",
            "parallelism": "one evaluation per GPU (independent theta candidates), no data-path collective",
            "l2": "inputs larger than L2 (K, L, K^-1 = 3 x %.2f GB per evaluation)" % (8.0 * N * N / 1e9)}


def golden_c3():
    """ll and the 19 gradients of ONE run of the compiled reference at C3's real size (tests/golden/make_golden_c3.py)"""
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "c3_reference.json")))
    except Exception:
        return None


def parity(ll, g, gold):
    g, gg = np.asarray(g, dtype=np.float64), np.asarray(gold["g"], dtype=np.float64)
    return {"ll_rel": abs(ll - gold["ll"]) / max(1.0, abs(gold["ll"])),
            "grad_rel_max": float(np.max(np.abs(g - gg) / np.maximum(1.0, np.abs(gg)))),
            "against": "compiled reference, one run at full size (tests/golden/c3_reference.json)"}


def single_gpu_leg(G, name, device, reps=3):
    """one workload on ONE GPU through gpc_eval (CGp mirror): seconds per evaluation, ll, gradient, phases"""
    w = WORKLOADS[name]
    X, y, p = make_inputs(name)
    k = G.make_kern(w["types"], w["D"])
    k.setParams(p)
    gp = G.CGp(k, X, y, device=device)
    ts = []
    for rep in range(reps):
        gp.KupToDate = False
        t0 = time.time()
        g, ll = gp.logLikelihoodGradient()
        ts.append(time.time() - t0)
    ph = gp.timings()
    out = {"workload": w["desc"], "seconds_per_eval": min(ts[1:]) if len(ts) > 1 else ts[0], "ll": ll, "g": list(map(float, g)),
           "phases_ms": ph, "tflops_equiv": w["N"] ** 3 / (min(ts[1:]) if len(ts) > 1 else ts[0]) / 1e12,
           "potrf_tflops": (w["N"] ** 3 / 3) / (ph["potrf"] * 1e-3) / 1e12,
           "kbuild_gbs_lower_triangle_written": 8.0 * w["N"] ** 2 / 2 / (ph["kbuild"] * 1e-3) / 1e9}
    gp.ctx.close()
    gold = golden_c3() if name == "c3" else None
    if gold:
        out["parity_vs_reference"] = parity(ll, g, gold)
    return out


def concurrent_leg(G, name, device, nctx, steps, warmup=2):
    """THROUGHPUT of `nctx` independent evaluations in flight on ONE GPU (one context + stream + host thread each: theta
    candidates of a line search / restarts / the outputs of a multi-kernel model): the serial chain of one evaluation's
    diagonal blocks runs under the tensor-core products of another.  Not the headline (an SCG run is sequential)."""
    import threading
    import torch
    from gpc_b200._lib import check, lib, ptr
    w = WORKLOADS[name]
    N, D = w["N"], w["D"]
    X, y, params = make_inputs(name)
    workers = []
    for i in range(nctx):
        k = G.make_kern(w["types"], D)
        k.setParams(params)
        ctx = G.DeviceContext(N, D, 1, device=device)
        st = torch.cuda.Stream(device=device)
        ctx.set_stream(st.cuda_stream)
        ctx.set_X(X)
        ctx.set_M(y)
        workers.append((k, ctx, st, k.getTransParams(), np.zeros(3), np.zeros(k.getNumParams())))
    gate = threading.Barrier(nctx + 1)
    errs = []

    def run(i, n0, n):
        k, ctx, st, tp0, out, g = workers[i]
        try:
            gate.wait()
            for s in range(n):
                k.setTransParams(theta_for_step(tp0, n0 + nctx * s + i))
                arr, m, keep = k._kcomps()
                rc = check(lib().gpc_eval(ctx.handle, arr, m, 0, ptr(out), ptr(g), None))
                if rc != 0:
                    raise RuntimeError("not positive definite (info=%d)" % rc)
        except Exception as e:   # a failed worker must not leave the others waiting
            errs.append(str(e))

    def phase(n0, n):
        th = [threading.Thread(target=run, args=(i, n0, n)) for i in range(nctx)]
        for t in th:
            t.start()
        torch.cuda.synchronize()
        gate.wait()
        t0 = time.time()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        return time.time() - t0

    phase(0, warmup)
    dt = phase(1000, steps)
    lls = [float(-0.5 * (wk[4][1] + wk[4][0]) - N * 0.5 * np.log(2 * np.pi)) for wk in workers]
    for wk in workers:
        wk[1].close()
    if errs:
        return {"error": errs[0]}
    return {"contexts": nctx, "evals": nctx * steps, "seconds": dt, "evals_per_sec": nctx * steps / dt,
            "ms_per_eval_amortised": 1e3 * dt / (nctx * steps), "ll_last": lls,
            "timing": "host wall clock around the worker threads, device synchronised on both sides"}


def sharded_leg(G, name, device, world, group, nb, reps=3):
    """one workload SHARDED over all ranks (gpc_dist_*: 2-D block-cyclic one-sweep K -> K^-1); host wall time around the
    collective call, max over ranks.  world == 1: the same algorithm on one GPU (one N^2 matrix resident)."""
    import torch
    import torch.distributed as dist
    from gpc_b200.dist import DistGp
    w = WORKLOADS[name]
    X, y, p = make_inputs(name)
    k = G.make_kern(w["types"], w["D"])
    k.setParams(p)
    if world > 1:
        gp = DistGp(k, X, y, nb=nb, backend="nccl", device=device, group=group)
    else:
        gp = DistGp(k, X, y, nb=nb, backend="local", devices=[device])
    ts = []
    for rep in range(reps):
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize()
        t0 = time.time()
        g, ll = gp.logLikelihoodGradient()   # synchronises internally (returns host scalars)
        t = time.time() - t0
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX, group=group)
            t = float(tt.item())
        ts.append(t)
    info = gp.info()
    gp.close()
    sec = min(ts[1:]) if len(ts) > 1 else ts[0]
    N = w["N"]
    return {"workload": w["desc"], "mode": "sharded over %d GPU(s): %dx%d process grid, %d x %d blocks, back-end %s"
            % (world, info["grid"][0], info["grid"][1], nb, nb, info["backend"]),
            "seconds_per_eval": sec, "all_seconds": ts, "evals_per_sec": 1.0 / sec, "tflops_equiv": N ** 3 / sec / 1e12,
            "ll": ll, "g": list(map(float, g)), "scaling": "strong",
            "comm_nranks": info["ranks"], "steps": info["steps"],
            "per_rank_matrix_bytes": info["local_matrix_bytes"], "per_rank_panel_buffer_bytes": info["panel_buffer_bytes"],
            "bytes_broadcast_per_step": info["bytes_broadcast_per_step"],
            "bytes_broadcast_per_eval": info["bytes_broadcast_per_step"] * info["steps"],
            "phases_ms_rank0": info["phases_ms"]}


def sharded_headline(args, G, name, kern, tp0, X, y, rank, world, local_rank):
    """--workload c3|c4 on N > 1 GPUs: ONE evaluation spread over all ranks per step (strong scaling); the JSON line has
    the same keys as the default one."""
    import torch
    import torch.distributed as dist
    from gpc_b200.dist import DistGp
    w = WORKLOADS[name]
    N, D, P = w["N"], w["D"], kern.getNumParams()
    nb = int(os.environ.get("GPC_DIST_NB", "1024" if N <= 32768 else "2048"))
    gloo = dist.new_group(backend="gloo")
    gp = DistGp(kern, X, y, nb=nb, backend="nccl", device=local_rank, group=gloo)

    def barrier():
        dist.barrier(group=gloo)
        torch.cuda.synchronize()

    def timed(nsteps, upload, base):
        barrier()
        l0 = gp.info()["launches"]
        t0 = time.time()
        for s_ in range(nsteps):
            kern.setTransParams(theta_for_step(tp0, base + s_))
            if upload:
                gp.set_data()
            g, ll = gp.logLikelihoodGradient()   # returns host scalars: the evaluation has completed on every rank
        barrier()
        t = torch.tensor([time.time() - t0], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=gloo)
        return float(t.item()) * 1e3, gp.info()["launches"] - l0, g, ll

    for s_ in range(args.warmup):
        kern.setTransParams(theta_for_step(tp0, s_))
        gp.logLikelihoodGradient()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, launches, g, ll = timed(args.steps, False, 100)
    info = gp.info()
    ms_e2e, _, _, _ = timed(args.steps, True, 200)
    clocks = sampler.stop() if rank == 0 else {}
    gp.close()
    if rank != 0:
        return
    mp = {}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = float(mp.get("bf16_tflops", 1590.0))
    from gpc_b200._lib import check, lib
    S = int(lib().gpc_gemm_engine_slices())
    pk = 2.0 * bf16 / (S * (S + 1) / 2.0)
    pk_src = "fp64-equivalent of 2 x bf16_tflops / 36, per GPU (int8 peak measurement failed)"
    pks = None
    try:   # the int8 pipe measured on this GPU: burst, and sustained over 2 s (the step is long: the power cap applies)
        b_, s_ = C.c_double(0), C.c_double(0)
        check(lib().gpc_bench_imma_peak(local_rank, 4, C.byref(b_)))
        check(lib().gpc_bench_imma_peak_sustained(local_rank, 4, 2.0, C.byref(s_)))
        pk = b_.value / (S * (S + 1) / 2.0)
        pks = s_.value / (S * (S + 1) / 2.0)
        pk_src = ("fp64-equivalent of the int8 tensor pipe measured on this GPU (gpc_bench_imma_peak, burst %.0f TOP/s; "
                  "sustained over 2 s %.0f TOP/s) / %d, per GPU" % (b_.value, s_.value, S * (S + 1) // 2))
    except Exception:
        pass
    sweep_ms = info["phases_ms"]["sweep"]
    ach = float(N) ** 3 / world / (sweep_ms * 1e-3) / 1e12
    line = {
        "metric": "gp_loglik_grad_evals_per_sec", "value": args.steps / (ms_dev * 1e-3), "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"],
                   "parallelism": "one evaluation sharded over %d GPUs: %dx%d process grid, %dx%d blocks 2-D block-cyclic, "
                                  "NCCL panel broadcasts" % (world, info["grid"][0], info["grid"][1], nb, nb),
                   "l2": "inputs larger than L2 (local matrix %.2f GB per rank)" % (info["local_matrix_bytes"] / 1e9)},
        "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": "evals/s",
                "h2d_bytes_per_step": int(world * (8 * N * D + 8 * N)), "d2h_bytes_per_step": int(world * 8 * (3 + P))},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "oz_gemm_kernel in block-cyclic mode (bulk rank-nb update of the local matrix)",
                     "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk, "traffic": None,
                     "peak_source": pk_src, "sustained_peak": pks, "frac_of_sustained_peak": (ach / pks) if pks else None,
                     "note": "achieved = N^3 / ranks / sweep time of rank 0 (panel production and broadcasts included)"},
        "cpu_baseline": None, "ll": ll, "tflops_equiv_total": float(N) ** 3 / (ms_dev / args.steps * 1e-3) / 1e12,
        "sharded": {"comm_nranks": info["ranks"], "per_rank_matrix_bytes": info["local_matrix_bytes"],
                    "per_rank_panel_buffer_bytes": info["panel_buffer_bytes"],
                    "bytes_broadcast_per_step": info["bytes_broadcast_per_step"], "steps_per_eval": info["steps"],
                    "phases_ms_rank0": info["phases_ms"]},
    }
    gold = golden_c3() if name == "c3" else None
    if gold:   # theta of the last step differs slightly from the golden's: parity is measured at the golden's theta
        line["parity_note"] = "see tests/test_gpu_full_size.py and the default run's `sharded.c3.parity_vs_reference`"
    print(json.dumps(line))


def theta_for_step(tp0, step):
    """a slightly different theta every step, as an optimiser would ask (keeps K well conditioned)"""
    tp = tp0.copy()
    tp[0] += 1e-3 * ((step % 7) - 3)
    return tp


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = []
        for line in open(self.f.name):
            t = [x.strip() for x in line.split(",")]
            if len(t) >= 9 and t[0] == str(self.idx):
                rows.append(t)
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        out["samples"] = len(rows)
        if sm:
            # median over the samples taken under load (>= 50% of the max seen)
            hi = [v for v in sm if v >= 0.5 * sm[-1]]
            out["sm_mhz"] = hi[len(hi) // 2]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        out["sm_max_mhz"] = max(mx) if mx else None
        pw = [float(r[3]) for r in rows if r[3].replace(".", "").isdigit()]
        out["power_w_max"] = max(pw) if pw else None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, nm in enumerate(names):
            if any(r[5 + i].lower().startswith("active") for r in rows):
                out["reasons"].append(nm)
        return out


def m2_summary(N, phases, hbm_gbs, bf16_tflops, dmma_peak_tflops, slices, imma_tops=0.0):
    """BASELINE.json's second metric, "potrf + K-build GF/s vs fp64 roofline" (SURVEY.md 8(d) M2), from the phase timings
    of one evaluation: potrf counted as N^3/3 flops against (a) the fp64 DMMA pipe measured on this GPU and (b) the
    fp64-equivalent peak of the int8 tensor pipe the large products run on; K build as 8 N^2/2 bytes written against the
    measured HBM copy bandwidth.  Pure arithmetic on numbers the bench already has; never raises."""
    out = {}
    try:
        pt = float(phases["potrf"]) * 1e-3
        kb = float(phases["kbuild"]) * 1e-3
        tf = (float(N) ** 3 / 3.0) / pt / 1e12 if pt > 0 else None
        gbs = 8.0 * float(N) ** 2 / 2.0 / kb / 1e9 if kb > 0 else None
        int8_equiv = 2.0 * float(bf16_tflops) / (slices * (slices + 1) / 2.0) if slices else None
        if imma_tops and slices:   # the measured pipe peak when available
            int8_equiv = float(imma_tops) / (slices * (slices + 1) / 2.0)
        out = {"potrf_tflops": tf, "potrf_gflops": tf * 1e3 if tf is not None else None,
               "potrf_frac_of_dmma_peak": tf / dmma_peak_tflops if tf and dmma_peak_tflops else None,
               "potrf_frac_of_int8_equiv_peak": tf / int8_equiv if tf and int8_equiv else None,
               "potrf_note": "potrf phase = factor + the triangular inverse built alongside it (1.25 N^3/3 executed flops, "
                             "DESIGN.md 4.3); the fraction counts only the N^3/3 of dpotrf_",
               "kbuild_gbs": gbs, "kbuild_frac_of_hbm": gbs / float(hbm_gbs) if gbs and hbm_gbs else None,
               "peaks": {"dmma_tflops": dmma_peak_tflops, "int8_equiv_fp64_tflops": int8_equiv, "hbm_gbs": hbm_gbs}}
    except Exception as e:  # the headline line must never be lost to this
        out = {"error": str(e)}
    return out


def reference_arm(args, rank, world):
    """--impl reference: the UNMODIFIED reference's CPU path (oracle/_ref = GPc compiled from /root/reference,
    OpenBLAS on all host threads) on the same workload.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import refbind as R
    name = args.workload
    w = WORKLOADS[name]
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgpcref.so missing (build with oracle/build_ref.sh)"}))
        return
    from oracle import gp_oracle as O   # transforms only: the reference arm loads nothing of gpc_b200
    X, y, params = make_inputs(name)
    tp0 = natural_to_trans(O, w, params)
    cores = os.cpu_count() or 1
    R.set_threads(cores)
    budget_s = float(os.environ.get("GPC_REF_BUDGET_S", "420"))
    t_start = time.time()
    # each step = one FULL evaluation of the workload (cold: K dirty); the number of steps is bounded by a wall
    # budget so that the run ends within a few minutes (one C2 evaluation is ~10-20 s of host time)
    times = []
    warm = args.warmup
    for s in range(warm + args.steps):
        r = R.gp_eval(w["types"], theta_for_step(tp0, s), X, y)
        if s >= warm:
            times.append(r["t_eval"])
        if time.time() - t_start > budget_s and times:
            break
    per = float(np.mean(times))
    val = 1.0 / per
    line = {
        "impl": "reference", "metric": "gp_loglik_grad_evals_per_sec", "value": val, "unit": "evals/s", "n_gpus": 0,
        "steps": len(times), "warmup": warm, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(w, 1),
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "reference",
                         "sample": "full %s evaluation x%d (requested %d steps; bounded by a %.0f s budget), "
                                   "GPc -O3 + OpenBLAS %d threads" % (name.upper(), len(times), args.steps, budget_s, cores)},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpc_b200", choices=["gpc_b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the extra C3 (N=32768) timing")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "gpc_b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import gpc_b200 as G
    from gpc_b200._lib import check, lib, ptr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (gpc_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    name = args.workload
    w = WORKLOADS[name]
    N, D = w["N"], w["D"]
    X, y, params = make_inputs(name)
    kern = G.make_kern(w["types"], D)
    kern.setParams(params)
    tp0 = kern.getTransParams()
    P = kern.getNumParams()

    if name != "c2" and world > 1:
        sharded_headline(args, G, name, kern, tp0, X, y, rank, world, local_rank)
        dist.destroy_process_group()
        return

    ctx = G.DeviceContext(N, D, 1, device=local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    # pinned host staging of the inputs (e2e leg uploads from here every step)
    Xp = torch.from_numpy(X.T.copy()).pin_memory()   # memory = column-major N x D
    yp = torch.from_numpy(y.T.copy()).pin_memory()
    ctx.set_X_ptr(Xp.data_ptr(), N, D, N)
    ctx.set_M_ptr(yp.data_ptr(), N, 1, N)
    out = np.zeros(3)
    g = np.zeros(P)

    def one_eval(step, upload):
        kern.setTransParams(theta_for_step(tp0, step))
        if upload:
            ctx.set_X_ptr(Xp.data_ptr(), N, D, N)
            ctx.set_M_ptr(yp.data_ptr(), N, 1, N)
        arr, n, keep = kern._kcomps()
        rc = check(lib().gpc_eval(ctx.handle, arr, n, 0, ptr(out), ptr(g), None))
        if rc != 0:
            raise SystemExit("bench: kernel matrix not positive definite (info=%d)" % rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, upload, base):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = ctx.launch_count()
        e0.record(stream)
        for s in range(nsteps):
            one_eval(base + s, upload)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.launch_count() - l0

    for s in range(args.warmup):
        one_eval(s, True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(args.steps, False, 100)
    phases = ctx.last_timings()
    enqueue_ms = None   # host time the last evaluation spent queueing its launches before its single synchronisation
    try:
        _enq = C.c_double(0.0)
        if lib().gpc_last_enqueue_ms(ctx.handle, C.byref(_enq)) == 0:
            enqueue_ms = float(_enq.value)
    except Exception:
        enqueue_ms = None
    ms_e2e, _ = timed(args.steps, True, 200)
    clocks = sampler.stop() if rank == 0 else {}
    ll = -0.5 * (out[1] + out[0]) - N * 0.5 * np.log(2 * np.pi)

    # ---- roofline of the dominant kernel.  Every GEMM/SYRK call of one evaluation is bracketed by CUDA events on the
    #      context's stream (profiling mode, calls serialised) and attributed to its engine:
    #        ozaki : oz_gemm_kernel (+ its slicing kernels) -- tcgen05.mma.kind::i8, S(S+1)/2 int8 MMAs per fp64 MMA
    #        dmma  : dgemm_kernel -- mma.sync.m8n8k4.f64
    #      The dominant one (by time) is reported: achieved = ALGORITHMIC fp64 flops of its calls / their summed duration.
    check(lib().gpc_ctx_set_profile(ctx.handle, 1))
    one_eval(300, False)
    gms, cnt, gfl = C.c_double(0), C.c_int64(0), C.c_double(0)
    check(lib().gpc_last_gemm_profile(ctx.handle, C.byref(gms), C.byref(cnt), C.byref(gfl)))
    split = np.zeros(8)
    check(lib().gpc_last_gemm_profile_split(ctx.handle, ptr(split)))
    check(lib().gpc_ctx_set_profile(ctx.handle, 0))
    peak = C.c_double(0)
    check(lib().gpc_bench_dmma_peak(local_rank, C.byref(peak)))
    imma = C.c_double(0)   # measured INT8 tensor-pipe peak of this GPU (tcgen05.mma.kind::i8, TMEM/SMEM-resident loop), TOP/s
    imma_sus = C.c_double(0)   # the same loop kept running for 2 s: the rate under the clock the power cap leaves
    try:
        check(lib().gpc_bench_imma_peak(local_rank, 4, C.byref(imma)))
    except Exception:
        imma = C.c_double(0)
    alg_flops = float(N) ** 3  # potrf N^3/3 + inverse 2N^3/3
    S = int(lib().gpc_gemm_engine_slices())
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile))
        except Exception:
            traffic = None
    mp = {}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = float(mp.get("bf16_tflops", 1590.0))
    bf16_src = "MEASURED_PEAKS.json bf16_tflops (burst)" if "bf16_tflops" in mp else "fallback 1.59 PFLOP/s (B200_PROFILING.md)"
    dm_ms, dm_fl, dm_n = split[0], split[1], int(split[2])
    oz_ms, oz_fl, oz_n, oz_ops = split[4], split[5], int(split[6]), split[7]
    engines = {
        "ozaki": {"calls": oz_n, "ms": oz_ms, "fp64_flops": oz_fl, "fp64_equiv_tflops": oz_fl / (oz_ms * 1e-3) / 1e12 if oz_ms else 0.0,
                  "int8_tops": oz_ops / (oz_ms * 1e-3) / 1e12 if oz_ms else 0.0},
        "dmma": {"calls": dm_n, "ms": dm_ms, "fp64_flops": dm_fl, "tflops": dm_fl / (dm_ms * 1e-3) / 1e12 if dm_ms else 0.0,
                 "peak_tflops": peak.value},
    }
    if oz_ms >= dm_ms and oz_ms > 0:
        # fp64-equivalent peak of the int8 pipe: int8 runs at twice the bf16 rate on sm_100a, one fp64 MMA costs
        # S(S+1)/2 = 36 int8 MMAs (S = 8 slices)
        pk_bf16 = 2.0 * bf16 / (S * (S + 1) / 2.0)
        pk = imma.value / (S * (S + 1) / 2.0) if imma.value > 0 else pk_bf16
        ach = oz_fl / (oz_ms * 1e-3) / 1e12
        roofline = {
            "bound": "tensor", "kernel": "oz_gemm_kernel (tcgen05.mma.kind::i8 + TMEM + TMA; fp64 via %d int8 slices) incl. slicing" % S,
            "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk,
            "traffic": ((traffic or {}).get("oz_gemm_kernel") or {}).get("dram_bytes_per_launch"),
            "traffic_note": "ncu --set full, SYRK-shaped call n=8192 k=4096 (profiles/oz_gemm_kernel_ncu_r01.txt); algorithmic bytes of that launch: %s" % ((traffic or {}).get("oz_gemm_kernel") or {}).get("algorithmic_bytes"),
            "peak_source": ("fp64-equivalent of the int8 tensor pipe MEASURED on this GPU in this run (gpc_bench_imma_peak: "
                            "tcgen05.mma.kind::i8 M128 N256 K32 on SMEM-resident operands, %.0f TOP/s) / S(S+1)/2 = %d int8 MMAs "
                            "per fp64 MMA" % (imma.value, S * (S + 1) // 2)) if imma.value > 0 else
                           "fp64-equivalent of 2 x %s (imma peak measurement failed)" % bf16_src,
            "int8_peak_tops_measured": imma.value,
            "peak_note": "burst figure (pseudo-random operands): the C2 step is 13 ms of mixed kernels; the long C3 / C4 legs "
                         "are also given against the sustained figure (2 s of the same loop under the power cap)",
            "peak_from_2x_bf16": pk_bf16, "frac_of_2x_bf16_peak": ach / pk_bf16,
            "int8_mmas_per_fp64_mma": S * (S + 1) // 2,
            "launches_per_eval": oz_n, "kernel_ms_per_eval": oz_ms, "algorithmic_flops": oz_fl,
            "share_of_step": oz_ms / (ms_dev / args.steps) if ms_dev > 0 else None,
            "kernel_only_note": "achieved / kernel_ms_per_eval bracket each engine call INCLUDING its slicing kernels (measured "
                                "live, above); oz_gemm_kernel alone is 6.18 ms of a C2 evaluation in the CUPTI timeline "
                                "(profiles/timeline_c2_r02.txt) = 86 TFLOP/s-equivalent",
        }
    else:
        ach = dm_fl / (dm_ms * 1e-3) / 1e12 if dm_ms else 0.0
        roofline = {
            "bound": "tensor", "kernel": "dgemm_kernel (mma.sync m8n8k4 f64 = DMMA.8x8x4)", "achieved": ach,
            "peak": peak.value, "unit": "TFLOP/s", "frac": ach / peak.value if peak.value else None,
            "traffic": ((traffic or {}).get("dgemm_kernel") or {}).get("dram_bytes_per_launch"),
            "peak_source": "measured on this GPU by gpc_bench_dmma_peak (register-resident DMMA loop, burst); "
                           "MEASURED_PEAKS.json has no fp64 figure",
            "launches_per_eval": dm_n, "kernel_ms_per_eval": dm_ms, "algorithmic_flops": dm_fl,
            "share_of_step": dm_ms / (ms_dev / args.steps) if ms_dev > 0 else None,
        }
    roofline["engines"] = engines
    roofline["algorithmic_flops_per_eval"] = alg_flops
    roofline["gemm_ms_per_eval_serialised"] = gms.value

    # ---- extra legs (never allowed to lose the headline line)
    also = None
    sharded = None
    if name == "c2" and not args.no_also:
        ctx.close()
        if world == 1 and rank == 0:
            # C3 (N=32768, rbfard) on one GPU, the north star's "< 1 s" target, with its parity against the ONE run of the
            # compiled reference committed under tests/golden/ (BASELINE.md section 3); C4 (N=65536) on one GPU through
            # the one-sweep path (one N^2 matrix in HBM instead of four)
            also = {}
            try:
                also["c3"] = single_gpu_leg(G, "c3", local_rank, reps=3)
            except Exception as e:
                also["c3"] = {"error": str(e)}
            try:
                also["c4"] = sharded_leg(G, "c4", local_rank, 1, None, nb=2048, reps=2)
            except Exception as e:
                also["c4"] = {"error": str(e)}
            # throughput with 2 / 3 independent C2 evaluations in flight on the one GPU
            also["c2_concurrent"] = {}
            for nc in (2, 3):
                try:
                    also["c2_concurrent"]["x%d" % nc] = concurrent_leg(G, "c2", local_rank, nc, max(4, args.steps // 2))
                except Exception as e:
                    also["c2_concurrent"]["x%d" % nc] = {"error": str(e)}
            for leg in also.values():
                if "tflops_equiv" in leg and imma.value > 0:
                    leg["roofline_frac"] = leg["tflops_equiv"] / (imma.value / (S * (S + 1) / 2.0))
        if world > 1:
            # the path that SHARDS: K -> K^-1 2-D block-cyclic over all ranks, NCCL panel broadcasts;
            # C3 (strong scaling against this run's own single-GPU evaluation on rank 0) and C4 (BASELINE configs[3])
            sharded = {}
            gloo = dist.new_group(backend="gloo")   # carries the 128-byte NCCL id of the library's own communicator
            for wl, nb in (("c3", 1024), ("c4", 2048)):
                try:
                    r_ = sharded_leg(G, wl, local_rank, world, gloo, nb=nb, reps=3)
                    barrier()
                    if rank == 0:
                        one = single_gpu_leg(G, wl, local_rank, reps=2)
                        r_["single_gpu_seconds_per_eval_same_run"] = one["seconds_per_eval"]
                        r_["speedup_vs_single_gpu"] = one["seconds_per_eval"] / r_["seconds_per_eval"]
                        r_["parity_vs_single_gpu"] = {
                            "ll_rel": abs(r_["ll"] - one["ll"]) / max(1.0, abs(one["ll"])),
                            "grad_rel_max": float(np.max(np.abs(np.array(r_["g"]) - np.array(one["g"])) /
                                                         np.maximum(1.0, np.abs(np.array(one["g"])))))}
                        if "parity_vs_reference" in one:
                            r_["single_gpu_parity_vs_reference"] = one["parity_vs_reference"]
                        gold = golden_c3() if wl == "c3" else None
                        if gold:
                            r_["parity_vs_reference"] = parity(r_["ll"], r_["g"], gold)
                    barrier()
                    if imma.value > 0:   # absolute rate and fraction of the measured int8-pipe roofline, per GPU
                        pk_ = imma.value / (S * (S + 1) / 2.0)
                        r_["tflops_equiv_per_gpu"] = r_["tflops_equiv"] / world
                        r_["roofline_frac_per_gpu"] = r_["tflops_equiv"] / world / pk_
                        r_["roofline_peak_per_gpu"] = pk_
                    sharded[wl] = r_
                except Exception as e:
                    sharded[wl] = {"error": str(e)}

    # ---- the SUSTAINED int8-pipe rate (2 s of the same MMA loop on pseudo-random operands: the clock the 1000 W power cap
    #      leaves), measured last so that it does not heat the GPU before the timed legs: the denominator for the long steps
    if rank == 0 and imma.value > 0:
        try:
            check(lib().gpc_bench_imma_peak_sustained(local_rank, 4, 2.0, C.byref(imma_sus)))
        except Exception:
            imma_sus = C.c_double(0)
        if imma_sus.value > 0:
            pks_ = imma_sus.value / (S * (S + 1) / 2.0)
            roofline["int8_peak_tops_sustained_2s"] = imma_sus.value
            roofline["frac_of_sustained_peak"] = roofline["achieved"] / pks_ if roofline.get("bound") == "tensor" and "int8_mmas_per_fp64_mma" in roofline else None
            for leg in (also or {}).values():
                if isinstance(leg, dict) and "tflops_equiv" in leg:
                    leg["roofline_frac_sustained_peak"] = leg["tflops_equiv"] / pks_
            for leg in (sharded or {}).values():
                if isinstance(leg, dict) and "tflops_equiv" in leg:
                    leg["roofline_frac_per_gpu_sustained_peak"] = leg["tflops_equiv"] / world / pks_
                    leg["roofline_sustained_peak_per_gpu"] = pks_

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import refbind as R
            if R.available():
                cores = os.cpu_count() or 1
                R.set_threads(cores)
                r = R.gp_eval(w["types"], tp0, X, y)
                kern.setTransParams(tp0)   # parity of this run's device result against the reference, same theta
                ctxp = G.DeviceContext(N, D, 1, device=local_rank)
                ctxp.set_X(X)
                ctxp.set_M(y)
                arr, n_, keep = kern._kcomps()
                outp, gp_ = np.zeros(3), np.zeros(P)
                check(lib().gpc_eval(ctxp.handle, arr, n_, 0, ptr(outp), ptr(gp_), None))
                ctxp.close()
                ll = -0.5 * (outp[1] + outp[0]) - N * 0.5 * np.log(2 * np.pi)
                gtr = gp_ * kern._gradfacts()
                cpu = {"value": 1.0 / r["t_eval"], "unit": "evals/s", "cores": cores, "kind": "reference",
                       "sample": "1 full %s evaluation (GPc -O3 + OpenBLAS %d threads): cold logLikelihood %.2f s + "
                                 "logLikelihoodGradient %.2f s" % (name.upper(), cores, r["t_ll"], r["t_grad"]),
                       "parity": {"ll_rel": abs(ll - r["ll"]) / max(1.0, abs(r["ll"])),
                                  "grad_rel_max": float(np.max(np.abs(gtr - r["g"]) / np.maximum(1.0, np.abs(r["g"]))))}}
            else:
                cpu = {"value": None, "unit": "evals/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
        except Exception as e:
            cpu = {"value": None, "unit": "evals/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % e}

    if rank == 0:
        per_ms = ms_dev / args.steps
        line = {
            "metric": "gp_loglik_grad_evals_per_sec", "value": world * args.steps / (ms_dev * 1e-3), "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(w, world),
            "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": "evals/s",
                    "h2d_bytes_per_step": int(8 * N * D + 8 * N), "d2h_bytes_per_step": int(8 * (8 + P) + 4)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "phases_ms": phases, "host_enqueue_ms": enqueue_ms, "ll": ll,
            "potrf_tflops": (N ** 3 / 3) / (phases["potrf"] * 1e-3) / 1e12,
            # the K build writes the LOWER triangle only (8 N^2 / 2 bytes; K is mirrored on download, never on the device):
            # kbuild_gbs counts the bytes actually written; SURVEY 8(d)'s figure for the full symmetric matrix is 2x this
            "kbuild_gbs": 8.0 * N * N / 2 / (phases["kbuild"] * 1e-3) / 1e9,
            "kbuild_bytes_written": int(8 * N * N // 2),
            "kbuild_gbs_full_matrix_equiv": 8.0 * N * N / (phases["kbuild"] * 1e-3) / 1e9,
            "m2": m2_summary(N, phases, mp.get("hbm_gbs"), bf16, peak.value, S, imma.value),
            "also": also,
            "sharded": sharded,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
