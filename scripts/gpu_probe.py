"""Quick per-kernel timing probe (development aid).  usage: [PROBLEM=car|quad] [DDP=0|1] [CHUNKS=n] [ITERS=50] gpu_probe.py B [B ...]"""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200"))
import numpy as np
import ilqg_b200
from ilqg_b200 import workloads as W
PROB = os.environ.get("PROBLEM", "car")
DDP = int(os.environ.get("DDP", "1" if PROB == "quad" else "0"))
CH = int(os.environ.get("CHUNKS", "0"))
ITERS = int(os.environ.get("ITERS", "50"))
for B in [int(a) for a in sys.argv[1:]] or [4096]:
    if PROB == "car":
        T, params = 500, W.CAR_PARAMS
        x0, u0 = W.car_batch(B)
    else:
        T, params = 1000, W.QUAD_PARAMS
        x0, u0 = W.quad_batch(B)
    s = ilqg_b200.BatchSolver(PROB, DDP, B, T, flags=ilqg_b200.TIMING, chunks=CH)
    s.set_params(params); s.set_options({"max_iter": 3}); s.upload(x0, u0); s.run(); s.sync(); s.timing()
    s.set_options({"max_iter": ITERS}); s.upload(x0, u0)
    t = time.perf_counter(); s.run(); s.sync(); dt = time.perf_counter() - t
    out = s.download(False); tm = s.timing()
    nls = out["n_linesearch"].sum()
    print(f"{PROB} ddp{DDP} B={B} chunks={s.chunks()}: {dt:.3f}s its={nls} -> {nls/dt:.0f} it/s ; kernels {tm}; rollouts {s.get_int('n_rollouts').sum()}+{s.get_int('n_tails').sum()}t backpasses {s.get_int('n_backpass').sum()} derivs {s.get_int('n_derivs').sum()} success {np.bincount(out['success']+1)} iters hist {np.bincount(out['iterations'])[-6:]}")
    s.close()
