"""python -m ilqg_gen <problem> [...] <outroot>        built-in problem modules
   python -m ilqg_gen --mac file.mac <name> <outroot>  a reference-style Maxima problem file
writes <outroot>/<name>/{iLQG_problem.h, iLQG_func.c, <name>_device.cuh}."""
import os
import sys

from .emit_c import emit_func_c, emit_problem_h
from .emit_cuda import emit_device
from .lower import lower
from .problems import REGISTRY


def generate(name, outdir, mac=None):
    """name: a module of ilqg_gen.problems, or (with mac=path) the name to give a problem read from a .mac file"""
    if mac:
        from .macfile import load_mac

        prob = load_mac(mac, name.capitalize())
    else:
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
    if len(sys.argv) == 5 and sys.argv[1] == "--mac":      # python -m ilqg_gen --mac file.mac <name> <outroot>
        generate(sys.argv[3], os.path.join(sys.argv[4], sys.argv[3]), mac=sys.argv[2])
        print("generated", sys.argv[3], "from", sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:-1] if len(sys.argv) > 2 else list(REGISTRY)
    root = sys.argv[-1] if len(sys.argv) > 1 else "."
    for n in names:
        generate(n, os.path.join(root, n))
        print("generated", n)
