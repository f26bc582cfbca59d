"""Why does bench.py's resident leg at 32768 problems take 131 ms with two chunks and a probe 115 ms?  Sequence variants (development aid)."""
import os, sys, time
sys.path.insert(0, "ddp-generator_b200")
import numpy as np, torch, ilqg_b200
from ilqg_b200 import workloads as W
B, T = 32768, 500
x0, u0 = W.car_batch(B)
x0p = torch.from_numpy(x0).pin_memory(); u0p = torch.from_numpy(u0).pin_memory()
xo = torch.empty((B, T + 1, 4), dtype=torch.float64).pin_memory(); uo = torch.empty((B, T, 2), dtype=torch.float64).pin_memory()
co = torch.empty(B, dtype=torch.float64).pin_memory(); io = torch.empty(B, dtype=torch.int32).pin_memory()
ro = torch.empty(B, dtype=torch.int32).pin_memory(); no = torch.empty(B, dtype=torch.int32).pin_memory()
hp = (x0p.data_ptr(), u0p.data_ptr(), xo.data_ptr(), uo.data_ptr(), co.data_ptr(), io.data_ptr(), ro.data_ptr(), no.data_ptr())
for variant in ("e2e_first_pinned", "resident_first", "e2e_first_events"):
    torch.cuda.set_stream(torch.cuda.Stream())
    s = ilqg_b200.BatchSolver("car", 0, B, T, chunks=2, stream=torch.cuda.current_stream().cuda_stream)
    s.set_params(W.CAR_PARAMS); s.set_options({"max_iter": 3})
    if variant.startswith("e2e_first"):
        s.solve_host_ptr(*hp)
    s.upload_ptr(x0p.data_ptr(), u0p.data_ptr()); s.run(); s.sync()
    ts = []
    for rep in range(3):
        s.set_options({"max_iter": 20}); s.upload_ptr(x0p.data_ptr(), u0p.data_ptr()); s.sync()
        torch.cuda.synchronize()
        if variant.endswith("events"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t = time.perf_counter(); s.run()
        if variant.endswith("events"):
            e1.record()
        s.sync(); ts.append((time.perf_counter() - t) * 1e3)
        s.set_options({"max_iter": 3}); s.solve_host_ptr(*hp)
    print(f"{variant}: " + " ".join(f"{t:.1f}" for t in ts) + " ms", flush=True)
    s.close()
