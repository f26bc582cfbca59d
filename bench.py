#!/usr/bin/env python
"""bench.py -- batched iLQG iterations/sec on the car-parking problem (T=500), BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE config 4 -- B = 262144 car-parking problems (n=4, m=2, T=500, FULL_DDP=0) with
random initial states and controls, strong-scaled: the B problems are split evenly over the N ranks,
one process per GPU, no collective on the solve path (torch.distributed is used for the timing barrier and the
final reduction of counters only).
A "step" is one pass of the iLQG loop (derivative kernel, backward-pass kernel, line-search kernel) over the rank's
whole shard; K steps = a solve with max_iter = K.  W warm-up steps run first on the same inputs (a throw-away
solve with max_iter = W).  `value` = (line searches performed by all problems on all ranks) / (max over ranks of
the device time of the K timed steps incl. the initial rollout), inputs resident in HBM.  `e2e` = the same count
over the time of ONE call of the public C ABI's end-to-end entry (ilqgb_solve_host): upload (pinned host -> HBM) + solve +
download of x, u, cost, iterations (HBM -> pinned host), pipelined per chunk stream.  The iteration count follows SURVEY.md 8d: loop passes that reached line_search.

`--impl reference` times the reference's own C solver (oracle/_ref, the unmodified sources compiled -O3
-ffp-contract=off; falls back to the oracle port if that library is not present) on all host cores, one solver
instance per thread, on a bounded sample of the same batch.
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
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "batched iLQG iterations/sec (car-parking T=500)"
UNIT = "iterations/s"
T_HOR = 500
PROBLEM, FULL_DDP = "car", 0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "sm_mhz_min": float(min(sm)) if sm else None, "power_w_max": float(max(pw)) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
def make_inputs(first, count, pinned):
    """Counter-based synthetic inputs for problems first..first+count-1, written into (pinned) host arrays."""
    from ilqg_b200 import workloads as W

    x0, u0 = pinned
    CH = 16384
    for s in range(0, count, CH):
        n = min(CH, count - s)
        a, b = W.car_batch(n, T=T_HOR, first=first + s)
        x0[s:s + n] = a
        u0[s:s + n] = b


def cpu_reference_run(n_problems, max_iter, threads, first=0):
    """Solve problems first..first+n-1 with the reference's C solver (or the port) on `threads` host threads."""
    import oracle_lib
    from ilqg_b200 import workloads as W

    kind = "reference" if oracle_lib.available("reference", PROBLEM, FULL_DDP, fast=True) else "port"
    if not oracle_lib.available(kind, PROBLEM, FULL_DDP, fast=True):
        return None
    O = oracle_lib.OracleLib(kind, PROBLEM, FULL_DDP, fast=True)
    x0, u0 = W.car_batch(n_problems, T=T_HOR, first=first)
    t0 = time.perf_counter()
    out = O.solve_batch(x0, u0, W.CAR_PARAMS, {"max_iter": float(max_iter)}, threads)
    dt = time.perf_counter() - t0
    its = int(out["n_linesearch"].sum())
    return dict(kind=kind, seconds=dt, iterations=its, value=its / dt, cores=threads, n_problems=n_problems, out=out)


def cpu_single_instance():
    """BASELINE config 1 on the host: one car instance (T=500, max_iter=200) with the reference's single-threaded build
    and with its experimental -DMULTI_THREADED=1 pipeline (north star: both are reported, neither is the target)."""
    import oracle_lib
    from ilqg_b200 import workloads as W

    out = {}
    x0, u0 = W.car_single()
    for key, kind, fast in (("single_thread", "reference", True), ("multi_threaded_2", "reference-mt", False)):
        if not oracle_lib.available(kind, PROBLEM, FULL_DDP, fast):
            continue
        s = oracle_lib.OracleLib(kind, PROBLEM, FULL_DDP, fast).solver(T_HOR)
        s.set_opts({"max_iter": 200})
        s.set_params(W.CAR_PARAMS)
        best = None
        for _ in range(3):
            s.init(x0, u0)
            t0 = time.perf_counter()
            s.solve()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        n = s.scalar("n_linesearch")
        out[key] = {"iterations_per_s": n / best, "ms_per_iteration": 1e3 * best / n, "iterations": int(n), "final_cost": s.scalar("cost")}
        s.close()
    return out


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = host_cores()
    per_core = max(2, int(1250 * 25 / max(args.steps, 1)))       # ~25 s at ~1250 it/s/core
    n = min(args.batch, per_core * cores)
    cpu_reference_run(min(n, 2 * cores), max(args.warmup, 1), cores)          # warm-up
    r = cpu_reference_run(n, args.steps, cores)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref nor the oracle port is built"}))
        return
    sample = f"first {n} problems of the batch, max_iter={args.steps}, one solver instance per thread"
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": cores, "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {"workload": f"car-parking n=4 m=2 T={T_HOR} FULL_DDP=0, batch {args.batch} random initial states "
                        f"(BASELINE config 4), max_iter={args.steps}, default options",
            "batch": args.batch, "horizon": T_HOR, "max_iter": args.steps, "sharding": "contiguous blocks of batch/N problems per GPU",
            "l2": "working set (>=160 KB per problem, tens of GB per GPU) is far larger than the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=262144)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunks", type=int, default=0, help="concurrent streams per GPU (0 = automatic)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    import ilqg_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the solver has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    # ---- shard and inputs (pinned host memory) ---------------------------------------------------------------------
    from ilqg_b200.sharding import shard_range

    first, count = shard_range(args.batch, rank, world)
    nx, nu = 4, 2
    x0_t = torch.empty((count, nx), dtype=torch.float64).pin_memory()
    u0_t = torch.empty((count, T_HOR, nu), dtype=torch.float64).pin_memory()
    make_inputs(first, count, (x0_t.numpy(), u0_t.numpy()))
    x_out = torch.empty((count, T_HOR + 1, nx), dtype=torch.float64).pin_memory()
    u_out = torch.empty((count, T_HOR, nu), dtype=torch.float64).pin_memory()
    cost_out = torch.empty(count, dtype=torch.float64).pin_memory()
    it_out = torch.empty(count, dtype=torch.int32).pin_memory()
    res_out = torch.empty(count, dtype=torch.int32).pin_memory()
    nls_out = torch.empty(count, dtype=torch.int32).pin_memory()

    from ilqg_b200 import workloads as W

    stream = torch.cuda.current_stream().cuda_stream
    S = ilqg_b200.BatchSolver(PROBLEM, FULL_DDP, count, T_HOR, device=local_rank, flags=0, stream=stream, chunks=args.chunks)
    S.set_params(W.CAR_PARAMS)

    # ---- warm-up: W passes on the same inputs ------------------------------------------------------------------------------
    S.set_options({"max_iter": args.warmup})
    S.upload_ptr(x0_t.data_ptr(), u0_t.data_ptr())
    S.run()
    S.sync()

    # ---- timed region 1: inputs resident, K passes ---------------------------------------------------------------------------
    S.set_options({"max_iter": args.steps})
    S.upload_ptr(x0_t.data_ptr(), u0_t.data_ptr())
    launches0 = S.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    S.run()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_dev = e0.elapsed_time(e1)
    launches = S.launch_count() - launches0
    S.download_ptr(None, None, cost_out.data_ptr(), it_out.data_ptr(), res_out.data_ptr(), nls_out.data_ptr())
    n_ls = int(nls_out.numpy().sum())
    cost_resident = cost_out.numpy().copy()

    ms_max = reduce(ms_dev, dist.ReduceOp.MAX if world > 1 else None)
    its_total = reduce(n_ls, dist.ReduceOp.SUM if world > 1 else None)
    value = its_total / (ms_max * 1e-3)

    # ---- timed region 2: end to end through the C ABI with host buffers ------------------------------------------------
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    S.solve_host_ptr(x0_t.data_ptr(), u0_t.data_ptr(), x_out.data_ptr(), u_out.data_ptr(), cost_out.data_ptr(), it_out.data_ptr(),
                     res_out.data_ptr(), nls_out.data_ptr())
    e3.record()
    torch.cuda.synchronize()
    wall_e2e = time.perf_counter() - t0
    ms_e2e = max(e2.elapsed_time(e3), wall_e2e * 1e3)
    barrier()
    n_ls_e2e = int(nls_out.numpy().sum())
    deterministic = bool(np.array_equal(cost_resident, cost_out.numpy()))
    ms_e2e_max = reduce(ms_e2e, dist.ReduceOp.MAX if world > 1 else None)
    its_e2e = reduce(n_ls_e2e, dist.ReduceOp.SUM if world > 1 else None)
    h2d = x0_t.numel() * 8 + u0_t.numel() * 8
    d2h = (x_out.numel() + u_out.numel() + cost_out.numel()) * 8 + (it_out.numel() + res_out.numel() + nls_out.numel()) * 4
    h2d_tot = reduce(h2d, dist.ReduceOp.SUM if world > 1 else None)
    d2h_tot = reduce(d2h, dist.ReduceOp.SUM if world > 1 else None)

    # ---- roofline of the dominant kernel: kernels timed ALONE (one stream, CUDA events around every launch) on the
    #      problems of one chunk of this rank's shard, i.e. at exactly the launch size the timed run uses --------------------
    n_k = max(1, count // S.chunks())
    K = ilqg_b200.BatchSolver(PROBLEM, FULL_DDP, n_k, T_HOR, device=local_rank, flags=ilqg_b200.TIMING, stream=stream, chunks=1)
    K.set_params(W.CAR_PARAMS)
    K.set_options({"max_iter": args.steps})
    K.upload_ptr(x0_t.data_ptr(), u0_t.data_ptr())
    K.sync()
    # the timed runs above drive the GPU into its power cap within ~0.5 s (SM clock sags from 1965 to 1350-1750 MHz at ~990 W,
    # scripts/gpu_probe4.py); the peak this leg is compared with is a burst figure (MEASURED_PEAKS.json: best of 10 copies), so
    # the kernels are timed under the same conditions: after a short idle, clocks sampled alongside
    time.sleep(1.5)
    ksampler = ClockSampler(local_rank)
    ksampler.start()
    K.run()
    K.sync()
    kclocks = ksampler.stop()
    ktime = K.timing(reset=True)
    n_dv = int(K.get_int("n_derivs").sum())
    n_bp = int(K.get_int("n_backpass").sum())
    n_roll = int(K.get_int("n_rollouts").sum())
    n_tail = int(K.get_int("n_tails").sum())
    K.close()
    NV = S.L.deriv_doubles_per_step
    RXU, RLL = 8 * ((nx + nu + 3) // 4) * 4, 8 * ((nu + nu * nx + 3) // 4) * 4      # record sizes in bytes
    alg = {
        "derivs": n_dv * (T_HOR * (RXU + NV * 8) + (RXU + (nx + nx * (nx + 1) // 2) * 8)),
        "backpass": n_bp * (T_HOR * (NV * 8 + nu * 8 + RLL) + (nx + nx * (nx + 1) // 2) * 8),
        # a stored rollout reads the nominal records and writes a candidate; a parallel-alpha tail reads the nominal once
        "linesearch": n_roll * (T_HOR * (RXU + RLL + RXU) + RXU) + n_tail * T_HOR * (RXU + RLL),
    }
    peak, peak_src = measured_peak_hbm()
    kernels = {}
    for k in ("derivs", "backpass", "linesearch"):
        ms, n = ktime[k]
        if n:
            kernels[k] = {"ms_total": ms, "launches": n, "ms_per_launch": ms / n, "algorithmic_bytes": alg[k],
                          "achieved_gbs": alg[k] / (ms * 1e-3) / 1e9 if ms > 0 else None,
                          "share_of_step": ms / max(sum(v[0] for v in ktime.values()), 1e-9)}
    # backward pass: fp64-issue bound.  Algorithmic multiply+add count per (problem, step) from SURVEY.md 8d (dense formula,
    # regType 1, no clamp) against the fp64 pipe peak measured on a B200 by csrc/fp64_peak.cu (profiles/fp64_peak_r01.json;
    # no FMA contraction is allowed on this path, so the comparable peak is the DMUL+DADD instruction rate)
    if "backpass" in kernels:
        n_, m_ = nx, nu
        pairs = 2 * n_**3 + 5 * n_**2 * m_ + 3 * n_ * m_**2 + n_**2 + 4 * n_ * m_ + 2 * m_**2 + m_
        fp = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
        peak64 = json.load(open(fp))["dmul_dadd_instr_per_s"] if os.path.exists(fp) else 1.847e13
        dp = 2.0 * pairs * n_bp * T_HOR / (kernels["backpass"]["ms_total"] * 1e-3)
        kernels["backpass"]["fp64"] = {"algorithmic_dp_instr_per_step": 2 * pairs, "achieved_dp_instr_per_s": dp,
                                       "peak_dp_instr_per_s": peak64, "frac": dp / peak64,
                                       "peak_source": "csrc/fp64_peak.cu on B200 (DMUL+DADD, no FMA)"}
    dom = max(kernels, key=lambda k: kernels[k]["ms_total"]) if kernels else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if dom and os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = None
    if dom:
        a = kernels[dom]["achieved_gbs"]
        roofline = {"kernel": {"derivs": "k_derivs", "backpass": "k_backpass", "linesearch": "k_ls_round"}[dom], "bound": "hbm", "achieved": a, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": a / peak if a else None, "traffic": traffic,
                    "algorithmic_bytes_per_launch": alg[dom] / kernels[dom]["launches"],
                    "timed_on": f"{n_k} problems (one chunk of the timed run), kernels alone on one stream, after 1.5 s idle",
                    "clocks": kclocks,
                    "note": "algorithmic bytes = record/entry bytes of DESIGN.md section 5 x units counted by the kernels"}

    # ---- CPU baseline on the host cores (rank 0, N = 1 only) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        n = min(args.batch, max(cores * 2, int(cores * 1250 * 15 / max(args.steps, 1))))
        r = cpu_reference_run(n, args.steps, cores)
        if r:
            # the sample doubles as a parity check of the timed GPU results (checker only, never the product path)
            same = bool(np.array_equal(r["out"]["cost"], cost_resident[:n])) if first == 0 else None
            cpu = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": r["kind"],
                   "sample": f"first {n} problems of the batch, max_iter={args.steps}, one solver instance per thread, {r['seconds']:.1f} s",
                   "gpu_costs_bit_identical_on_sample": same, "config1_single_instance": cpu_single_instance()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / max(args.steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args), "clocks": clocks,
            "e2e": {"value": its_e2e / (ms_e2e_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_tot / max(args.steps, 1),
                    "d2h_bytes_per_step": d2h_tot / max(args.steps, 1), "h2d_bytes": h2d_tot, "d2h_bytes": d2h_tot,
                    "seconds": ms_e2e_max * 1e-3},
            "gpu_launches": launches, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
            "iterations_total": its_total, "seconds": ms_max * 1e-3, "deterministic_rerun": deterministic,
            "streams_per_gpu": S.chunks(),
        }
        print(json.dumps(line))
    S.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
