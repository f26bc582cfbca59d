"""Exploration (development aid): car problems whose value function overflows, GPU vs reference record by record."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_util as PU
from ilqg_b200 import workloads as W
T = 60
x0, u0 = W.car_batch(1, T=T, seed=33)
for scale in (1e300, 1e304, 1e306, 1e307, 1e308):
    for ddp in (0, 1):
        params = dict(W.CAR_PARAMS, cf=[0.1 * scale, 0.1, 1.0, 0.3])
        opts = {"max_iter": 12}
        ora = PU.oracle_record(PU.oracle_kinds("car", ddp)[0], "car", ddp, T, params, x0[0], u0[0], opts)
        gpu = PU.gpu_records("car", ddp, T, params, x0, u0, opts)[0]
        keys = ("result", "iterations", "n_ls", "n_bp", "cost", "lambda")
        diff = [k for k in keys + ("x", "u", "l", "L", "tr_alpha", "tr_lambda") if not np.array_equal(np.asarray(ora[k]), np.asarray(gpu[k]), equal_nan=True)]
        print(scale, ddp, "ORA", {k: ora[k] for k in keys}, "nanL", int(np.isnan(ora["L"]).sum()), "| GPU", {k: gpu[k] for k in keys}, "nanL", int(np.isnan(gpu["L"]).sum()), "| differs:", diff)
