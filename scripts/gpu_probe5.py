"""Per-pass kernel times over a 50-pass solve (development aid). usage: gpu_probe5.py B"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200"))
import ilqg_b200
from ilqg_b200 import workloads as W
B = int(sys.argv[1])
x0, u0 = W.car_batch(B)
s = ilqg_b200.BatchSolver("car", 0, B, 500, chunks=1, flags=2)
s.set_params(W.CAR_PARAMS); s.set_options({"max_iter": 50}); s.upload(x0, u0); s.start(); s.sync(); s.timing()
prev = None
for g in range(10):
    s.iterate(5); s.sync()
    t = s.timing()
    cur = s.get_int("cur"); st = s.get_int("status"); nd = s.get_int("n_derivs").sum(); nr = s.get_int("n_rollouts").sum()
    print(f"passes {5*g:2d}-{5*g+4:2d}: derivs {t['derivs'][0]/max(t['derivs'][1],1):.3f} ms  backpass {t['backpass'][0]/max(t['backpass'][1],1):.3f} ms  "
          f"ls {t['linesearch'][0]/5:.3f} ms/pass ({t['linesearch'][1]} launches)  running {int((st==1).sum()) if (st==1).any() else int((st==0).sum())}  "
          f"cur=1 frac {cur.mean():.2f}  n_derivs {nd}  n_rollouts {nr}  status hist {np.bincount(st).tolist()}")
