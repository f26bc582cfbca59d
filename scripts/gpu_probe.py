"""Quick per-kernel timing probe (development aid)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200"))
import numpy as np
import ilqg_b200
from ilqg_b200 import workloads as W
CH = int(os.environ.get("CHUNKS", "0"))
for B in [int(a) for a in sys.argv[1:]] or [4096]:
    x0, u0 = W.car_batch(B)
    s = ilqg_b200.BatchSolver("car", 0, B, 500, flags=ilqg_b200.TIMING, chunks=CH)
    s.set_params(W.CAR_PARAMS); s.set_options({"max_iter": 3}); s.upload(x0, u0); s.run(); s.sync(); s.timing()
    s.set_options({"max_iter": 50}); s.upload(x0, u0)
    t = time.perf_counter(); s.run(); s.sync(); dt = time.perf_counter() - t
    out = s.download(False); tm = s.timing()
    nls = out["n_linesearch"].sum()
    print(f"B={B} chunks={s.chunks()}: {dt:.3f}s its={nls} -> {nls/dt:.0f} it/s ; kernels {tm}; rollouts {s.get_int('n_rollouts').sum()} backpasses {s.get_int('n_backpass').sum()} derivs {s.get_int('n_derivs').sum()} iters hist {np.bincount(out['iterations'])[-5:]}")
    s.close()
