"""Exploration (development aid): car problems whose backward pass overflows (finite * finite = inf, then inf * structural zero),
GPU vs reference record by record."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_util as PU
from ilqg_b200 import workloads as W
T = 30
x0, u0 = W.car_batch(1, T=T, seed=33)
u0 = u0.copy(); u0[:, :, 1] = 0.0
for d, cf2 in ((1e-60, 1e300), (1e-60, 1e250), (1e-40, 1e300), (1e-80, 1e200), (1e-20, 1e305)):
    for ddp in (0, 1):
        params = dict(W.CAR_PARAMS, d=[d], cf=[0.1, 0.1, cf2, 0.3])
        opts = {"max_iter": 6}
        ora = PU.oracle_record(PU.oracle_kinds("car", ddp)[0], "car", ddp, T, params, x0[0], u0[0], opts)
        if not ora["init_ok"]:
            print(d, cf2, ddp, "init failed"); continue
        gpu = PU.gpu_records("car", ddp, T, params, x0, u0, opts)[0]
        keys = ("result", "iterations", "n_ls", "n_bp", "cost", "lambda", "g_norm", "dV0", "dV1")
        diff = [k for k in keys + ("x", "u", "l", "L", "tr_alpha", "tr_lambda", "tr_newcost") if not np.array_equal(np.asarray(ora[k]), np.asarray(gpu[k]), equal_nan=True)]
        print(d, cf2, ddp, "ORA", {k: ora[k] for k in keys}, "nan/inf L", int(np.isnan(ora["L"]).sum()), int(np.isinf(ora["L"]).sum()), "tr_a", ora["tr_alpha"], "tr_l", ora["tr_lambda"],
              "| GPU", {k: float(gpu[k]) for k in keys}, "nan/inf L", int(np.isnan(gpu["L"]).sum()), int(np.isinf(gpu["L"]).sum()), "tr_a", gpu["tr_alpha"], "tr_l", gpu["tr_lambda"], "| differs:", diff)
