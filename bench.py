#!/usr/bin/env python
"""bench.py -- LML+gradient evaluations/sec of a DSMGP on B200 (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg3b|cfg2|cfg4|cfg5] [--impl reference]

A "step" is ONE LML+gradient evaluation of the whole model = optimisers.jl:43-77 of the reference without the
Flux step (setparams! -> fit! -> mll! -> updategradients! -> nabla-mll!) = one `dsmgp_eval`.

* value    device time (CUDA events on the library's stream, max over ranks), inputs resident in HBM.
* e2e      the same evaluation through the public API (`handle.eval(theta)`): host theta in, host (lml, grad) out,
           wall clock including the per-step host->device parameter upload, the device->host read of the per-leaf
           rows and the host tree passes.
* N > 1    the SAME model, leaves sharded over the ranks by LPT on n^3 (strong scaling); one NCCL all-reduce (SUM) of
           the L x (1+H) per-leaf row table per evaluation, then every rank finishes the O(L) tree passes.
* roofline FP64: the dominant kernel's algorithmic flops / its CUDA-event time, against the FP64 DGEMM rate measured
           on this pool's B200 (tools/fp64_peaks.cu; MEASURED_PEAKS.json has no FP64 entry).
* --impl reference   the reference's CPU path: Julia is not installed, so the oracle port in the reference's
           algorithmic shape (oracle/reference_shape.py), all host threads, on a bounded sample of leaves.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)



def fp64_peaks():
    """FP64 roofs measured on this pool's B200 with tools/fp64_peaks.cu (MEASURED_PEAKS.json has HBM / bf16 only): the newest
    profiles/fp64_peaks_r*.json.  DGEMM = cuBLAS 8192^3 (burst == sustained: FP64 is not power capped), DMMA = issue roof."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "fp64_peaks_r*.json")))
    if not files:
        return {"dgemm": 35.9, "dmma": 37.2, "hbm": 6568.0, "source": "fallback constants (profiles/fp64_peaks_r*.json missing)"}
    d = json.load(open(files[-1]))
    return {"dgemm": float(d["dgemm_nt_8192_tflops"]), "dmma": float(d["dmma_tflops_w32"]), "hbm": float(d.get("hbm_copy_gbs", 6568.0)),
            "source": os.path.relpath(files[-1], ROOT)}


def int8_peak():
    """INT8 tcgen05 issue rate measured on this pool (tools/ozaki_proto.cu with OZAKI_RATE=1: 8192 back-to-back kind::i8 MMAs per SM
    from shared memory): the 128 x 256 x 32 shape is the tensor roof, 128 x 128 x 32 the shape the block products use."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "int8_peaks_r02.json")))
        return float(d["int8_m128_n256_tops"]), float(d["int8_m128_n128_tops"]), "profiles/int8_peaks_r02.json"
    except Exception:
        return 4500.0, 4233.0, "fallback: nominal 4.5 POP/s dense INT8 (profiles/int8_peaks_r02.json missing)"


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6458.4, "fallback 6458.4 (MEASURED_PEAKS.json missing)"


def ncu_traffic_gb(kernel, workload, world):
    """dram__bytes_read + dram__bytes_write of one launch from the newest committed `ncu --set full` summary of that kernel
    (profiles/ncu_full_<kernel>_r*.txt, captured on the cfg3 1-GPU bench); None for any other configuration."""
    import glob
    if workload != "cfg3" or world != 1:
        return None, None
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"ncu_full_{kernel}_r*.txt")))
    if not files:
        return None, None
    rd = wr = None
    for ln in open(files[-1]):
        if ln.startswith("dram__bytes_read.sum,Gbyte,"):
            rd = float(ln.strip().split(",")[2])
        if ln.startswith("dram__bytes_write.sum,Gbyte,"):
            wr = float(ln.strip().split(",")[2])
    if rd is None or wr is None:
        return None, None
    return rd + wr, os.path.relpath(files[-1], ROOT)


WORKLOADS = {
    # name: (N, D, kernel, V, K, M, depth, eps, seed)   SURVEY §8(d)
    "cfg1": dict(N=100, D=1, kernel="isose", V=3, K=4, M=10, depth=2, eps=0.5, seed=1),
    "cfg2": dict(N=10_000, D=1, kernel="isose", V=3, K=4, M=100, depth=2, eps=0.5, seed=2),
    "cfg3": dict(N=40_000, D=8, kernel="ardse", V=3, K=4, M=500, depth=2, eps=0.5, seed=3),
    "cfg3iso": dict(N=40_000, D=8, kernel="isose", V=3, K=4, M=500, depth=2, eps=0.5, seed=3),   # HBM-bound Gram build
    "cfg3b": dict(N=40_000, D=8, kernel="ardse", V=3, K=4, M=500, depth=2, eps=0.0, seed=3),
    "cfg4": dict(N=45_730, D=9, kernel="isose+isolinear", V=4, K=4, M=1000, depth=2, eps=0.5, seed=4),
    "cfg5": dict(N=1_000_000, D=8, kernel="ardse", V=3, K=4, M=2000, depth=4, eps=0.1, seed=5),
}


def make_data(w):
    rng = np.random.default_rng(w["seed"])
    x = rng.random((w["N"], w["D"]))
    if w["D"] == 1:
        x = np.sort(x, axis=0)
    wv = rng.standard_normal(w["D"])
    y = np.sin(2 * np.pi * (x @ wv)) + 0.1 * rng.standard_normal(w["N"])
    return x, y


def thetas(nparams_per_kernel, seed):
    rng = np.random.default_rng(1000 + seed)
    base = []
    for npk in nparams_per_kernel:
        base.extend([0.0] * (npk - 2) + [0.0, -1.0])
    base = np.array(base)
    return [base] + [base + 0.3 * rng.standard_normal(base.size) * (np.arange(base.size) >= 0) for _ in range(3)]


def sample_clocks(stop, out):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    dev = os.environ.get("LOCAL_RANK", "0")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", dev, f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            out.append([s.strip() for s in r.stdout.strip().split(",")])
        except Exception:
            pass
        stop.wait(0.2)


def clocks_summary(samples):
    sm = [float(s[0]) for s in samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
    mx = [float(s[1]) for s in samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
    reasons = set()
    for s in samples:
        if len(s) >= 6:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons), "samples": len(sm)}


class _DevPtr:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (ptr, False), "version": 3}


def oracle_kernel(w):
    from oracle import dsm_oracle as orc
    if w["kernel"] == "isose":
        return [orc.IsoSE(0.0, 0.0)]
    if w["kernel"] == "ardse":
        return [orc.ArdSE(np.zeros(w["D"]), 0.0)]
    return [orc.IsoSE(0.0, 0.0), orc.IsoLinear(0.0)]


def build_structure(w, device=False):
    """Region graph: returns x, y, root, kernels (product types).  device=False: the host builder (no GPU; the CPU reference
    arm).  device=True: the same recursion and random draws with the data passes on the GPU (dsmgp_part_*) --
    bit-identical graph (tests/test_gpu_round2.py::test_device_tree_construction_is_bit_identical)."""
    from deepstructuredmixtures_b200 import kernels as kr, structure as st
    x, y = make_data(w)
    if w["kernel"] == "isose":
        kern = kr.IsoSE(0.0, 0.0)
    elif w["kernel"] == "ardse":
        kern = kr.ArdSE(np.zeros(w["D"]), 0.0)
    else:
        kern = [kr.IsoSE(0.0, 0.0), kr.IsoLinear(0.0)]
    cfg = st.DSMGPConfig(None, kern, -1.0, w["M"], w["K"], w["V"], w["depth"], w["eps"], True)
    root = (st.buildTree_device if device else st.buildTree)(x, y, cfg, np.random.default_rng(w["seed"]))
    return x, y, root, kern


def blas_threads(want=None):
    """Pin the BLAS pool to `want` threads (default: every host core) regardless of OMP_NUM_THREADS -- torchrun exports
    OMP_NUM_THREADS=1 to its workers -- and return the number of threads the pool really uses."""
    from threadpoolctl import threadpool_info, threadpool_limits
    want = want or os.cpu_count() or 1
    threadpool_limits(limits=want)
    got = [int(p.get("num_threads", 1)) for p in threadpool_info() if p.get("user_api") == "blas"]
    return max(got) if got else 1


def cpu_baseline(w, x, y, root, budget_s, optimised=False):
    from deepstructuredmixtures_b200 import structure as st
    from oracle import reference_shape as rs
    threads = blas_threads()
    leaves = st.getLeaves(root)
    ok = oracle_kernel(w)
    # mixtures: time the first kernel's leaves and the second's separately through their own kernel objects
    res_total, sample_s, used = 0.0, 0.0, []
    for kid, k in enumerate(ok):
        lv = [lf for lf in leaves if lf.kernelid - 1 == kid]
        r = rs.sample_model_time(x, y, [lf.obs - 1 for lf in lv], [lf.mean for lf in lv], k, -1.0,
                                 budget_s=budget_s / len(ok), optimised=optimised)
        res_total += r["seconds_per_eval"]; sample_s += r["sample_seconds"]; used.extend(r["sample_sizes"])
    sum_n3 = float(sum(float(lf.nobs) ** 3 for lf in leaves))
    what = ("minimal CPU algorithm (1x dpotrf, dpotri, O(n^2) traces" if optimised else
            "oracle port in the reference's algorithmic shape (2x update_cholesky!, potrs(-I)+GEMM traces, gradients twice")
    return {"value": 1.0 / res_total, "unit": "evals/s", "cores": threads, "host_cores": os.cpu_count(),
            "kind": "port", "extrapolated": True, "seconds_per_eval_extrapolated": res_total,
            "sample_seconds": sample_s, "sample_share_of_sum_n3": float(sum(float(n) ** 3 for n in used)) / sum_n3,
            "sample": f"{what}; SciPy/OpenBLAS, {threads} BLAS threads) on leaves of size {used} "
                      f"({sample_s:.1f} s measured), extrapolated by sum n^3 to all {len(leaves)} leaves; "
                      f"seconds/eval={res_total:.1f}"}


def run_reference(args, w):
    """The reference's CPU path (Julia is not installed: the oracle port in the reference's algorithmic shape) on the box's
    host cores.  A step = one bounded sample of leaves (timed for real), whose cost is extrapolated to a whole evaluation by
    sum n^3: `value` is that extrapolated rate, `ms_per_step` the MEASURED wall time of a step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["DSMGP_STRUCTURE_ONLY"] = "1"      # host-side region-graph builder only: libdsmgp.so is not mapped into this arm
    x, y, root, _ = build_structure(w)
    nsteps = max(1, min(args.steps, 3))
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(w, x, y, root, budget_s=3.0)
    vals, walls = [], []
    for _ in range(nsteps):
        t0 = time.perf_counter()
        vals.append(cpu_baseline(w, x, y, root, budget_s=max(10.0, 60.0 / nsteps)))
        walls.append(time.perf_counter() - t0)
    best = max(vals, key=lambda r: r["value"])
    v = float(np.mean([r["value"] for r in vals]))
    line = {"impl": "reference", "metric": "DSMGP LML+gradient evals/sec", "value": v, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": len(vals), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * float(np.mean(walls)),
            "ms_per_eval_extrapolated": 1e3 / v, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, w),
            "cpu_baseline": dict(best, value=v),
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "native_library_loaded": any("libdsmgp" in ln for ln in open("/proc/self/maps")),
            "note": "Julia is not installed in this image: the reference arm is the oracle port timed in the "
                    "reference's algorithmic shape on the host cores; each step times a bounded stratified sample of leaves "
                    "(ms_per_step) and extrapolates it to a whole evaluation by sum n^3 (value, ms_per_eval_extrapolated)"}
    print(json.dumps(line), flush=True)


def workload_config(name, w):
    return {"workload": f"{name}: synthetic {w['N']}x{w['D']} {w['kernel']} DSMGP V={w['V']} K={w['K']} M={w['M']} "
                        f"depth={w['depth']} eps={w['eps']} (SURVEY 8d)",
            "l2": "factor arena (GBs) >> 126 MB L2, no flush needed", "theta": "4 fixed hyper-parameter vectors cycled"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-predict", action="store_true")
    ap.add_argument("--mathematical", action="store_true", help="true gradients instead of as-written")
    ap.add_argument("--no-sub-records", action="store_true", help="skip the `mathematical` and `scale_cfg5` sub-records")
    ap.add_argument("--no-share", action="store_true", help="fit_naive!: do not hand the overlap matrix to the library")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libdsmgp has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import deepstructuredmixtures_b200 as dsm
    from deepstructuredmixtures_b200 import model as mdl

    t_build0 = time.perf_counter()
    x, y, root, kern = build_structure(w, device=True)
    klist = kern if isinstance(kern, list) else [kern]
    t_tree = time.perf_counter() - t_build0
    keep = args.workload != "cfg5"
    t0 = time.perf_counter()
    model = mdl.DSMGP(root, x, y, [k.copy() for k in klist], -1.0, rank=rank, world=world, device=local,
                      keep_factors=keep, as_written_grads=not args.mathematical)
    t_create = time.perf_counter() - t0
    H = model.handle
    ths = thetas([k.nparams for k in klist], w["seed"])
    L = len(model.leaves)
    sizes = np.array([lf.nobs for lf in model.leaves], dtype=np.float64)
    if world > 1:
        # the collective lives INSIDE the library (dsmgp_comm_init): from here on dsmgp_eval is the same call at any world size
        from deepstructuredmixtures_b200.distributed import init_library_comm
        init_library_comm(model)
    sharing = None
    if not args.no_share and world == 1 and L <= 4096:
        # train! calls fit!(spn, D, gpmap) every iteration (optimisers.jl:45): the library gets the overlap matrix and shares
        # what the reference's fit! shares (identical experts once, common leading block rows copied)
        t0 = time.perf_counter()
        H.set_sharing(model.D, 0.05)
        kind, _, blocks = H.get_sharing()
        sharing = {"identical_experts": int((kind == 1).sum()), "prefix_experts": int((kind == 2).sum()),
                   "block_rows_copied": int(blocks.sum()), "overlap_and_plan_s": time.perf_counter() - t0}

    def step_device(i):
        """one LML+gradient evaluation through dsmgp_eval (world > 1: the row table is all-reduced inside the library)"""
        H.eval(ths[i % len(ths)])
        return H.timings()

    for i in range(args.warmup):
        step_device(i)
    samples, stop = [], threading.Event()
    th = threading.Thread(target=sample_clocks, args=(stop, samples), daemon=True)
    th.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    phase = {k: 0.0 for k in ("gram_ms", "potrf_ms", "solve_ms", "inverse_ms", "grad_ms", "total_ms")}
    launches = 0
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        tm = step_device(args.warmup + i)
        for k in phase:
            phase[k] += tm[k]
        launches += int(tm["launches"])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    stop.set(); th.join(timeout=2)
    dev_ms = phase["total_ms"]
    # e2e through the public API on host buffers (world == 1: the same call; world > 1: the wall clock above)
    tt = torch.tensor([dev_ms, t_wall * 1e3, phase["potrf_ms"], phase["inverse_ms"], phase["gram_ms"]], dtype=torch.float64,
                      device=f"cuda:{local}")
    fl = torch.tensor([tm["potrf_flops"], tm["inverse_flops"], tm["gram_bytes"]], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(fl, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms = float(tt[0]), float(tt[1])
    phase["potrf_ms"], phase["inverse_ms"], phase["gram_ms"] = float(tt[2]), float(tt[3]), float(tt[4])
    tm = dict(tm, potrf_flops=float(fl[0]), inverse_flops=float(fl[1]), gram_bytes=float(fl[2]))
    line = None
    if rank == 0:
        value = args.steps / (dev_ms * 1e-3)
        e2e = args.steps / (wall_ms * 1e-3)
        potrf_fl, inv_fl = tm["potrf_flops"], tm["inverse_flops"]
        potrf_tf = potrf_fl * args.steps / (phase["potrf_ms"] * 1e-3) * 1e-12 if phase["potrf_ms"] > 0 else 0.0
        inv_tf = inv_fl * args.steps / (phase["inverse_ms"] * 1e-3) * 1e-12 if phase["inverse_ms"] > 0 else 0.0
        gram_gbs = tm["gram_bytes"] * args.steps / (phase["gram_ms"] * 1e-3) * 1e-9 if phase["gram_ms"] > 0 else 0.0
        # roofline numbers are PER GPU (the aggregate flops of all ranks / world / the slowest rank's phase time)
        potrf_tf /= world; inv_tf /= world; gram_gbs /= world
        dom = "trtri3_kernel" if phase["inverse_ms"] >= phase["potrf_ms"] else "potrf2_kernel"
        ach = inv_tf if dom == "trtri3_kernel" else potrf_tf
        if inv_fl > 0 and phase["inverse_ms"] < 0.02 * phase["potrf_ms"]:
            # small shards run the factorisation and the inverse as ONE persistent launch (csrc/fused2.cuh): one phase, one rate
            dom = "eval2_kernel"
            ach = potrf_tf = inv_tf = (potrf_fl + inv_fl) * args.steps / ((phase["potrf_ms"] + phase["inverse_ms"]) * 1e-3) * 1e-12 / world
        # INT8 split path (csrc/api_ozaki.cu): the GEMM-shaped 3/4 of the factorisation and the inverse run as error-free INT8
        # slice products on the tcgen05 tensor cores; the per-phase FP64 rates above then mix two engines and are reported as
        # FP64-EQUIVALENT rates (they may exceed the FP64 roof).  The dominant kernel is the block-product kernel.
        i8 = H.int8_info()
        int8 = None
        if i8["batches"] > 0 and i8["gemm_ms"] > 0:
            pk8, pk8_shape, pk8_src = int8_peak()
            fp64_part = i8["fp64_tile_flops"]                     # rank 0: 2/3 r^3 per diagonal range of r rows
            int8 = {"slices": i8["slices"], "gemm_ms": i8["gemm_ms"], "slice_ms": i8["slice_ms"], "fp64_tile_ms": i8["fp64_tile_ms"],
                    "int8_tops": i8["int8_ops"] / (i8["gemm_ms"] * 1e-3) * 1e-12,
                    "gemm_fp64_equiv_tflops": i8["fp64_equiv_flops"] / (i8["gemm_ms"] * 1e-3) * 1e-12,
                    "fp64_tile_tflops": fp64_part / (i8["fp64_tile_ms"] * 1e-3) * 1e-12 if i8["fp64_tile_ms"] > 0 else None,
                    "share_of_flops_on_int8": 1.0 - fp64_part * world / (potrf_fl + inv_fl) if potrf_fl + inv_fl > 0 else None,
                    "pool_gb": i8["pool_bytes"] * 1e-9,
                    "what": "rank 0, last timed evaluation, CUDA events around the launches: block products (oz::gemm_kernel, tcgen05 kind::i8, "
                            "TMEM accumulators, TMA loads), slicing (oz::slice_kernel), FP64 tile pipelines (eval2_kernel, DMMA)"}
            if i8["gemm_ms"] >= i8["fp64_tile_ms"]:
                dom = "oz::gemm_kernel"
            else:
                # small shards: the chain-bound FP64 tile launches (factorisation + inverse of the diagonal ranges) take longer
                # than the block products; their rate against the FP64 roof
                dom = "eval2_kernel"
                ach = int8["fp64_tile_tflops"]
        # DRAM bytes per launch from `ncu --set full` (profiles/ncu_full_*_r01e.csv); only known for the profiled config
        traffic, traffic_src = ncu_traffic_gb("oz_gemm_kernel" if dom == "oz::gemm_kernel" else dom, args.workload, world)
        pk = fp64_peaks()
        hbm, hbm_src = hbm_peak()
        ns_local = int(np.sum(H.leaf_owner() == 0))
        line = {
            "metric": "DSMGP LML+gradient evals/sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args.workload, w), leaves=L, n_min=int(sizes.min()),
                           n_median=int(np.median(sizes)), n_max=int(sizes.max()), sum_n3=float(np.sum(sizes ** 3)),
                           grads="mathematical" if args.mathematical else "as-written", parallelism=f"leaf-shard x{world}"),
            "e2e": {"value": e2e, "unit": "evals/s",
                    "h2d_bytes_per_step": int(ns_local * (4 + w["D"]) * 8),
                    "d2h_bytes_per_step": int(L * H.row_width * 8 + ns_local * 48)},
            "gpu_launches": launches,
            "cholesky_gflops": potrf_tf * 1e3 * world,
            "phases_ms_per_step": {k: v / args.steps for k, v in phase.items()},
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": pk["dgemm"], "unit": "TFLOP/s",
                         "frac": ach / pk["dgemm"], "traffic": traffic, "traffic_unit": "GB per launch (ncu dram read+write)",
                         "traffic_source": traffic_src, "frac_of_dmma_issue_roof": ach / pk["dmma"],
                         "algorithmic": "flops per launch = sum over local experts of n^3/3 + n^2/2 + n/6 (SURVEY 8d), one launch per evaluation; per-GPU figures",
                         "peak_source": f"cuBLAS DGEMM 8192^3 measured on this pool (tools/fp64_peaks.cu -> {pk['source']}); "
                                        f"MEASURED_PEAKS.json has no FP64 entry; DMMA issue roof {pk['dmma']:.1f}",
                         "potrf_tflops": potrf_tf, "inverse_tflops": inv_tf, "gram_gbs": gram_gbs,
                         "gram_frac_hbm": gram_gbs / hbm, "hbm_peak_source": hbm_src,
                         "gram_bound": "FP64 ALU (D exp per element)" if "ard" in w["kernel"] else "HBM write"},
            "sharing": sharing,
            "int8_split": int8,
            "dtype_note": ("FP64 in, FP64 out.  75 % of the factorisation / inverse flops are computed as error-free INT8 slice products of the "
                           "FP64 operands (Ozaki scheme, 8 slices of 7 bits, exact int32 accumulation; measured error below cuBLAS DGEMM's), "
                           "the rest on FP64 DMMA; DSMGP_OZAKI=0 runs everything on DMMA") if int8 else None,
            "clocks": clocks_summary(samples),
            "host": {"tree_build_s": t_tree, "create_upload_s": t_create},
        }
        if int8 is not None:
            # split path: two kernels share the evaluation (block products on the INT8 tensor cores, FP64 tile launches on DMMA).  `roofline`
            # is the one that took longer in the timed evaluation, `roofline_other` the second one.
            fp64_roof = dict(line["roofline"])
            t8, t8_src = ncu_traffic_gb("oz_gemm_kernel", args.workload, world)
            t64, t64_src = ncu_traffic_gb("eval2_kernel", args.workload, world)
            r_int8 = {
                "bound": "tensor", "kernel": "oz::gemm_kernel<%d>" % int8["slices"], "achieved": int8["int8_tops"], "peak": pk8,
                "unit": "TFLOP/s", "frac": int8["int8_tops"] / pk8, "ops": "INT8 multiply-adds counted as 2 operations (TOP/s)",
                "ms": int8["gemm_ms"], "frac_of_issue_rate_of_the_shape_used": int8["int8_tops"] / pk8_shape,
                "traffic": t8, "traffic_unit": "GB per launch (ncu dram read+write, one of the four block-product launches)", "traffic_source": t8_src,
                "algorithmic": "operations per evaluation = 2 * 128 * 128 * 32 per tcgen05.mma x S (S + 1) / 2 slice pairs x the k-steps of all "
                               "block products (L21 = A21 X11^T, A22 -= L21 L21^T, T = L21 X11, X21 = -X22 T of every expert with >= 8 block rows); "
                               "achieved = those operations / the CUDA-event time of the four launches; per GPU",
                "peak_source": f"tcgen05 kind::i8 128x256x32 issue rate measured on this pool ({pk8_src}); the 128x128x32 shape used (TMEM holds "
                               f"four 128-column accumulators) issues at {pk8_shape:.0f}",
                "fp64_equivalent_tflops_of_the_block_products": int8["gemm_fp64_equiv_tflops"]}
            r_fp64 = {
                "bound": "tensor", "kernel": "eval2_kernel", "achieved": int8["fp64_tile_tflops"], "peak": pk["dgemm"], "unit": "TFLOP/s",
                "frac": int8["fp64_tile_tflops"] / pk["dgemm"], "ms": int8["fp64_tile_ms"],
                "traffic": t64, "traffic_unit": "GB per launch (ncu dram read+write, one of the two launches)", "traffic_source": t64_src,
                "algorithmic": "flops = 2/3 r^3 per diagonal range of r rows (factorisation + inverse of the two half-size ranges of every split expert, "
                               "all of the unsplit ones); achieved = flops / the CUDA-event time of the two fused launches; per GPU",
                "peak_source": fp64_roof["peak_source"], "frac_of_dmma_issue_roof": int8["fp64_tile_tflops"] / pk["dmma"]}
            first, second = (r_int8, r_fp64) if int8["gemm_ms"] >= int8["fp64_tile_ms"] else (r_fp64, r_int8)
            first["fp64_phases"] = {k: fp64_roof[k] for k in ("potrf_tflops", "inverse_tflops", "gram_gbs", "gram_frac_hbm", "hbm_peak_source", "gram_bound")}
            first["note"] = ("potrf_tflops / inverse_tflops are FP64-equivalent rates of phases that mix the INT8 products with the FP64 (DMMA) tile "
                             "pipelines; with DSMGP_OZAKI=0 every flop runs on DMMA (profiles/bench_r02_cfg3_1gpu_fp64.json: 0.84 of the DGEMM roof)")
            line["roofline"] = first
            line["roofline_other"] = second
        if world == 1 and keep and not args.no_predict:
            # update! + predict (common.jl:323-334, 294-307) on T fresh test points: every point is routed to one leaf
            # per sum-node branch; device time of predict_kernel and wall time of the public call (routing, H2D of the
            # routed points, kernel, D2H, log-space mixing on the host).
            prng = np.random.default_rng(77)
            T = min(w["N"], 40_000)
            xt = prng.random((T, w["D"]))
            mdl.update_(model)
            mdl.predict(model, xt)          # warm-up at full size: the routed-point scratch (GBs) is allocated once
            t0 = time.perf_counter()
            mu, var = mdl.predict(model, xt)
            t_pred = time.perf_counter() - t0
            tp = H.timings()
            line["predict"] = {"T": T, "wall_ms": t_pred * 1e3, "points_per_s": T / t_pred,
                               "kernel_ms": tp["predict_ms"],
                               "kernel_tflops": tp["predict_flops"] / (tp["predict_ms"] * 1e-3) * 1e-12 if tp["predict_ms"] > 0 else None,
                               "kernel_gbs": tp["predict_bytes"] / (tp["predict_ms"] * 1e-3) * 1e-9 if tp["predict_ms"] > 0 else None,
                               "finite": bool(np.isfinite(mu).all() and np.isfinite(var).all()),
                               "algorithmic": "sum over leaves of n^2 T_l + 2 n T_l flop (SURVEY 8d); bytes = L read once per "
                                              "128-point block + inputs + outputs"}
        if world == 1 and args.workload != "cfg5":
            # cold end-to-end: the whole model from HOST arrays every step -- dsmgp_create (upload of x, y and the
            # leaves' index lists, device gather) + one evaluation + read-back + destroy.
            reps = 3
            t0 = 0.0
            for i in range(-1, reps):          # one untimed cold step first: it fills the library's device-buffer cache
                if i == 0:
                    t0 = time.perf_counter()
                m2 = mdl.DSMGP(root, x, y, [k.copy() for k in klist], -1.0, rank=0, world=1, device=local,
                               keep_factors=keep, as_written_grads=not args.mathematical)
                m2.handle.eval(ths[i % len(ths)])
                m2.close()
            t_cold = (time.perf_counter() - t0) / reps
            nidx = int(sizes.sum())
            line["e2e_cold"] = {"value": 1.0 / t_cold, "unit": "evals/s",
                                "h2d_bytes_per_step": int(x.nbytes + nidx * 16 + L * 8 + ns_local * (4 + w["D"]) * 8),
                                "d2h_bytes_per_step": int(L * H.row_width * 8 + ns_local * 48),
                                "what": "model construction from host arrays (dsmgp_create: x, centred y, 1-based leaf rows) "
                                        "+ one dsmgp_eval + dsmgp_destroy per step, after one untimed cold step"}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(w, x, y, root, args.cpu_budget)
            # the algorithmically minimal CPU version (one factorisation, dpotri, O(n^2) traces), so that the GPU/CPU
            # ratio can also be read without the reference's redundant work
            line["cpu_baseline_optimised"] = cpu_baseline(w, x, y, root, max(5.0, args.cpu_budget / 3), optimised=True)
    model.close()
    if not args.no_sub_records:
        if world == 1 and not args.mathematical and args.workload != "cfg5":
            # the benchmarked as-written ArdSE gradient needs no LAUUM pass (its length-scale part is identically 0,
            # kernels.jl:161 / SURVEY App. B Q3): the same evaluation with the TRUE gradients, so that the rate is not only
            # meaningful through that quirk
            m3 = mdl.DSMGP(root, x, y, [k.copy() for k in klist], -1.0, device=local, as_written_grads=False)
            for i in range(2):
                m3.handle.eval(ths[i % len(ths)])
            ms, ph = 0.0, {"gram_ms": 0.0, "potrf_ms": 0.0, "inverse_ms": 0.0, "grad_ms": 0.0}
            nrep = max(3, min(args.steps, 5))
            for i in range(nrep):
                m3.handle.eval(ths[i % len(ths)])
                tmm = m3.handle.timings()
                ms += tmm["total_ms"]
                for k in ph:
                    ph[k] += tmm[k] / nrep
            line["mathematical"] = {"value": nrep / (ms * 1e-3), "unit": "evals/s", "ms_per_step": ms / nrep, "steps": nrep,
                                    "phases_ms_per_step": ph,
                                    "lauum_tflops": tmm["potrf_flops"] / (ph["grad_ms"] * 1e-3) * 1e-12 if ph["grad_ms"] > 0 else None,
                                    "what": "same workload with as_written_grads=0: true d LML / d theta (adds the LAUUM pass with the fused dK traces)"}
            m3.close()
        if args.workload == "cfg3":
            rec = scale_record(args, rank, world, local)
            if rank == 0:
                line["scale_cfg5"] = rec
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def scale_record(args, rank, world, local):
    """north_star's scale run as a bounded sub-record: synthetic 1,000,000 x 8 ArdSE DSMGP (depth 4, 20,736 experts), experts
    sharded over the ranks, factors streamed (factor -> reduce -> discard), ONE all-reduce of the per-leaf rows per evaluation
    inside the library.  1 warm-up + 2 timed evaluations; device time = max over ranks."""
    import torch
    import torch.distributed as dist
    from deepstructuredmixtures_b200 import model as mdl
    w = WORKLOADS["cfg5"]
    t0 = time.perf_counter()
    x, y, root, kern = build_structure(w, device=True)
    t_tree = time.perf_counter() - t0
    t0 = time.perf_counter()
    model = mdl.DSMGP(root, x, y, [kern.copy()], -1.0, rank=rank, world=world, device=local, keep_factors=False)
    if world > 1:
        from deepstructuredmixtures_b200.distributed import init_library_comm
        init_library_comm(model)
    t_create = time.perf_counter() - t0
    H = model.handle
    ths = thetas([kern.nparams], w["seed"])
    H.eval(ths[0])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    nrep, dev_ms, pot_ms, inv_ms = 2, 0.0, 0.0, 0.0
    tw = time.perf_counter()
    for i in range(nrep):
        H.eval(ths[(1 + i) % len(ths)])
        tm = H.timings()
        dev_ms += tm["total_ms"]; pot_ms += tm["potrf_ms"]; inv_ms += tm["inverse_ms"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall_ms = (time.perf_counter() - tw) * 1e3
    tt = torch.tensor([dev_ms, wall_ms, pot_ms, inv_ms], dtype=torch.float64, device=f"cuda:{local}")
    fl = torch.tensor([tm["potrf_flops"], tm["inverse_flops"]], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(fl, op=dist.ReduceOp.SUM)
    L = len(model.leaves)
    model.close()
    dev_ms, wall_ms, pot_ms, inv_ms = (float(v) for v in tt)
    return {"workload": workload_config("cfg5", w)["workload"], "n_gpus": world, "steps": nrep, "warmup": 1,
            "value": nrep / (dev_ms * 1e-3), "unit": "evals/s", "ms_per_step": dev_ms / nrep,
            "e2e": {"value": nrep / (wall_ms * 1e-3), "unit": "evals/s"},
            "potrf_tflops_per_gpu": float(fl[0]) * nrep / (pot_ms * 1e-3) * 1e-12 / world,
            "inverse_tflops_per_gpu": float(fl[1]) * nrep / (inv_ms * 1e-3) * 1e-12 / world,
            "leaves": L, "scaling": "strong (same 1M-point model, experts sharded by LPT on n^3)",
            "host": {"tree_build_s": t_tree, "create_upload_s": t_create}}


if __name__ == "__main__":
    main()
