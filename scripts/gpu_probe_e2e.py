"""End-to-end overlap probe (development aid): resident vs end-to-end rate for chunk counts / stream priorities / start gates.
usage: gpu_probe_e2e.py [B=262144] [STEPS=20]   (env: CONFIGS="chunks:prio:stagger,..." to override the sweep)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200"))
import numpy as np
import torch

import ilqg_b200
from ilqg_b200 import workloads as W

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
T = 500
IDLE = float(os.environ.get("IDLE", "1.5"))   # seconds of idle before every timed call (bench.py's regime); 0 = back to back
res = {"B": B, "steps": STEPS, "runs": []}

# PCIe copy rates with pinned memory (what the DMA path can do at best)
n = 1 << 27
hp = torch.empty(n, dtype=torch.float64).pin_memory()
dv = torch.empty(n, dtype=torch.float64, device="cuda")
for name, (dst, src) in (("h2d", (dv, hp)), ("d2h", (hp, dv))):
    best = 0.0
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, n * 8 / (time.perf_counter() - t0) / 1e9)
    res["pcie_" + name + "_gbs"] = best
# both directions at once
hp2 = torch.empty(n, dtype=torch.float64).pin_memory()
dv2 = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s1):
    dv.copy_(hp, non_blocking=True)
with torch.cuda.stream(s2):
    hp2.copy_(dv2, non_blocking=True)
torch.cuda.synchronize()
res["pcie_bidir_gbs_each"] = n * 8 / (time.perf_counter() - t0) / 1e9
del hp, dv, hp2, dv2
print(json.dumps({k: v for k, v in res.items() if k.startswith("pcie")}), flush=True)

x0 = torch.empty((B, 4), dtype=torch.float64).pin_memory()
u0 = torch.empty((B, T, 2), dtype=torch.float64).pin_memory()
for s in range(0, B, 16384):
    m = min(16384, B - s)
    a, b = W.car_batch(m, T=T, first=s)
    x0.numpy()[s:s + m] = a
    u0.numpy()[s:s + m] = b
xo = torch.empty((B, T + 1, 4), dtype=torch.float64).pin_memory()
uo = torch.empty((B, T, 2), dtype=torch.float64).pin_memory()
co = torch.empty(B, dtype=torch.float64).pin_memory()
io = torch.empty(B, dtype=torch.int32).pin_memory()
ro = torch.empty(B, dtype=torch.int32).pin_memory()
no = torch.empty(B, dtype=torch.int32).pin_memory()

configs = os.environ.get("CONFIGS")
if configs:
    sweep = [tuple(int(v) for v in c.split(":")) for c in configs.split(",")]
elif B >= 200000:
    sweep = [(4, 0, 0), (4, 1, 0), (8, 1, 0), (8, 0, 0), (4, 1, 4), (8, 1, 2), (6, 1, 0)]
else:
    sweep = [(1, 0, 0), (2, 0, 0), (2, 1, 0), (4, 1, 0), (4, 0, 0), (8, 1, 0), (4, 1, 3)]
ref_cost = None
for chunks, prio, stagger in sweep:
    os.environ["ILQG_E2E_PRIO"] = str(prio)
    os.environ["ILQG_E2E_STAGGER"] = str(stagger)
    S = ilqg_b200.BatchSolver("car", 0, B, T, chunks=chunks)
    S.set_params(W.CAR_PARAMS)
    S.set_options({"max_iter": 3})
    S.upload_ptr(x0.data_ptr(), u0.data_ptr()); S.run(); S.sync()
    S.set_options({"max_iter": STEPS})
    S.upload_ptr(x0.data_ptr(), u0.data_ptr()); S.sync()
    torch.cuda.synchronize()
    time.sleep(IDLE)
    t0 = time.perf_counter(); S.run(); S.sync(); t_res = time.perf_counter() - t0
    S.download_ptr(None, None, co.data_ptr(), io.data_ptr(), ro.data_ptr(), no.data_ptr())
    nls = int(no.numpy().sum())
    cost_res = co.numpy().copy()
    te = []
    S.solve_host_ptr(x0.data_ptr(), u0.data_ptr(), xo.data_ptr(), uo.data_ptr(), co.data_ptr(), io.data_ptr(), ro.data_ptr(), no.data_ptr())   # warm-up (staging buffers)
    for _ in range(2):
        time.sleep(IDLE)
        t0 = time.perf_counter()
        S.solve_host_ptr(x0.data_ptr(), u0.data_ptr(), xo.data_ptr(), uo.data_ptr(), co.data_ptr(), io.data_ptr(), ro.data_ptr(), no.data_ptr())
        te.append(time.perf_counter() - t0)
    same = bool(np.array_equal(cost_res, co.numpy()))
    if ref_cost is None:
        ref_cost = cost_res
    r = {"chunks": S.chunks(), "prio": prio, "stagger": stagger, "t_resident": t_res, "value": nls / t_res, "t_e2e": te, "e2e": nls / min(te),
         "ratio": t_res / min(te), "e2e_equals_resident": same, "equals_first_config": bool(np.array_equal(ref_cost, cost_res))}
    res["runs"].append(r)
    print(json.dumps(r), flush=True)
    S.close()
    del S
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"probe_e2e_B{B}_S{STEPS}.json"), "w"), indent=1)
