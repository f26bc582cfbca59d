"""python -m ilqg_gen <problem> <outdir>: write iLQG_problem.h, iLQG_func.c and <problem>_device.cuh."""
import os
import sys

from .emit_c import emit_func_c, emit_problem_h
from .emit_cuda import emit_device
from .lower import lower
from .problems import REGISTRY


def generate(name, outdir):
    prob = REGISTRY[name]()
    m = lower(prob)
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, "iLQG_problem.h"), "w") as f:
        f.write(emit_problem_h(m))
    with open(os.path.join(outdir, "iLQG_func.c"), "w") as f:
        f.write(emit_func_c(m))
    with open(os.path.join(outdir, f"{name}_device.cuh"), "w") as f:
        f.write(emit_device(m, "Prob" + prob.name))
    return m


if __name__ == "__main__":
    names = sys.argv[1:-1] if len(sys.argv) > 2 else list(REGISTRY)
    root = sys.argv[-1] if len(sys.argv) > 1 else "."
    for n in names:
        generate(n, os.path.join(root, n))
        print("generated", n)
