"""Kernel timing sanity check (development aid): the derivative kernel timed by the library's CUDA events, (a) 20 launches back to
back, (b) single launches separated by host idle time, and the per-class times of a normal run.  usage: gpu_probe4.py B"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ddp-generator_b200"))
import ilqg_b200
from ilqg_b200 import workloads as W
B = int(sys.argv[1])
x0, u0 = W.car_batch(B)
s = ilqg_b200.BatchSolver("car", 0, B, 500, chunks=1, flags=2)
s.set_params(W.CAR_PARAMS); s.set_options({"max_iter": 12}); s.upload(x0, u0); s.start(); s.sync()
s.timing()
for rep in range(3):
    for i in range(20):
        s.phase("derivs")
    s.sync()
    ms, n = s.timing()["derivs"]
    print(f"derivs back-to-back x{n}: {ms / n:.4f} ms/launch  -> {B * 500 * 208 / (ms / n) / 1e6:.0f} GB/s algorithmic")
for rep in range(5):
    time.sleep(0.2)
    s.phase("derivs"); s.sync()
    ms, n = s.timing()["derivs"]
    print(f"derivs single after idle: {ms / n:.4f} ms")
s.run(); s.sync()
print({k: (round(v[0] / max(v[1], 1), 4), v[1]) for k, v in s.timing().items()})
# sustained behaviour: the same kernel back to back for seconds (power cap / clock sag shows up as a rising time per launch)
import subprocess
s2 = ilqg_b200.BatchSolver("car", 0, B, 500, chunks=1, flags=2)
s2.set_params(W.CAR_PARAMS); s2.set_options({"max_iter": 12}); s2.upload(x0, u0); s2.start(); s2.sync(); s2.timing()
for which, reps in (("derivs", 12), ("backpass", 8)):
    for rep in range(reps):
        for i in range(200 if which == "derivs" else 60):
            s2.phase(which)
        q = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader"], stdout=subprocess.PIPE, text=True)
        s2.sync()
        ms, n = s2.timing()[which]
        print(f"{which} sustained group {rep}: {ms / n:.4f} ms/launch   smi: {q.communicate()[0].strip()}")
