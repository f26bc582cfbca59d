"""compute-sanitizer target (development aid): tiny solves through every backward-pass kernel.
usage: compute-sanitizer --tool racecheck|memcheck python scripts/gpu_sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200"))
import ilqg_b200
from ilqg_b200 import workloads as W

def run(problem, ddp, B, T, params, x0, u0, tuning, it=6):
    s = ilqg_b200.BatchSolver(problem, ddp, B, T)
    s.set_params(params); s.set_options({"max_iter": it})
    for k, v in tuning.items():
        s.set_tuning(k, v)
    out = s.solve(x0, u0)
    print(problem, ddp, tuning, "iterations", int(out["iterations"].sum()), flush=True)
    s.close()

x0, u0 = W.car_batch(19, T=60, seed=4)
run("car", 0, 19, 60, W.CAR_PARAMS, x0, u0, {"bp_split": 4, "bp_ppw": 8})
run("car", 0, 19, 60, W.CAR_PARAMS, x0, u0, {"bp_split": 4, "bp_ppw": 4})
run("car", 0, 19, 60, W.CAR_PARAMS, x0, u0, {"bp_split": 0})
run("car", 1, 19, 60, W.CAR_PARAMS, x0, u0, {})
x0, u0 = W.quad_batch(6, T=40)
run("quad", 1, 6, 40, W.QUAD_PARAMS, x0, u0, {}, it=4)
run("quad", 1, 6, 40, W.QUAD_PARAMS, x0, u0, {"cw_lpp": 8}, it=4)
x0, u0 = W.pend_batch(5)
run("pend", 0, 5, W.PEND_T, W.PEND_PARAMS, x0, u0, {"bp_split": 4}, it=8)
