"""Resident solve rate by chunk count at small batches, without the per-kernel timing events (development aid).
usage: gpu_probe_chunks.py B [B ...]   (env ITERS=50, CHUNKSET="1 2 4")"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200"))
import ilqg_b200
from ilqg_b200 import workloads as W
ITERS = int(os.environ.get("ITERS", "50"))
if os.environ.get("ILQG_LIB_DIR"):
    ilqg_b200.LIB_DIR = os.environ["ILQG_LIB_DIR"]
for B in [int(a) for a in sys.argv[1:]] or [4096]:
    x0, u0 = W.car_batch(B, first=int(os.environ.get("FIRST", "0")))
    for ch in [int(c) for c in os.environ.get("CHUNKSET", "1 2 4").split()]:
        s = ilqg_b200.BatchSolver("car", 0, B, 500, chunks=ch)
        s.set_params(W.CAR_PARAMS); s.set_options({"max_iter": 3}); s.upload(x0, u0); s.run(); s.sync()
        s.set_options({"max_iter": ITERS})
        best = 1e9
        for rep in range(3):
            s.upload(x0, u0); s.sync()
            t = time.perf_counter(); s.run(); s.sync(); best = min(best, time.perf_counter() - t)
        nls = s.download(False)["n_linesearch"].sum()
        try:
            split = s.get_int('bp_split')[0]
        except Exception:
            split = "?"
        print(f"B={B} chunks={s.chunks()} split={split}: {best*1e3:.1f} ms, {nls/best/1e6:.3f} M it/s, {best/ITERS*1e3:.3f} ms/pass", flush=True)
        s.close()
