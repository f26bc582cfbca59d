"""ncu target: one short quadrotor solve (development aid). usage: gpu_probe_q.py B max_iter"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200"))
import ilqg_b200
from ilqg_b200 import workloads as W
B, it = int(sys.argv[1]), int(sys.argv[2])
x0, u0 = W.quad_batch(B)
s = ilqg_b200.BatchSolver("quad", 1, B, W.QUAD_T, chunks=int(os.environ.get("CHUNKS", "0")))
s.set_params(W.QUAD_PARAMS); s.set_options({"max_iter": it}); s.upload(x0, u0); s.run(); s.sync()
print("done", s.download(False)["n_linesearch"].sum())
