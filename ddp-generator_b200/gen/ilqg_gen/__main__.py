"""python -m ilqg_gen <problem> [...] <outroot>        built-in problem modules
   python -m ilqg_gen --mac file.mac <name> <outroot>  a reference-style Maxima problem file
writes <outroot>/<name>/{iLQG_problem.h, iLQG_func.c, <name>_device.cuh} and, for problems without folded constraints,
iLQG_MMex.c (the single-evaluation gateway of iLQG_MMex.tem)."""
import os
import sys

from .emit_c import emit_func_c, emit_problem_h
from .emit_cuda import emit_device
from .emit_mmex import emit_mmex_c
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
    mmex = os.path.join(outdir, "iLQG_MMex.c")
    if any(m.n_mu.values()):      # no multiplier inputs in the MMex interface: such problems have none (iLQG_MMex.tem)
        if os.path.exists(mmex):
            os.remove(mmex)
    else:
        with open(mmex, "w") as f:
            f.write(emit_mmex_c(m))
    return m


def main(argv=None):
    import argparse

    ap = argparse.ArgumentParser(prog="python -m ilqg_gen", description="Generate iLQG_problem.h, iLQG_func.c and <name>_device.cuh "
                                 "for built-in problem modules or for a reference-style Maxima problem file.")
    ap.add_argument("--mac", metavar="FILE.mac", help="read the problem from a .mac file; exactly one NAME must follow")
    ap.add_argument("names", nargs="*", metavar="NAME", help=f"problems to generate (default: all of {', '.join(REGISTRY)})")
    ap.add_argument("outroot", help="output root; files go to <outroot>/<name>/")
    a = ap.parse_args(argv)
    if a.mac:
        if len(a.names) != 1:
            ap.error("--mac needs exactly one NAME")
        generate(a.names[0], os.path.join(a.outroot, a.names[0]), mac=a.mac)
        print("generated", a.names[0], "from", a.mac)
        return
    unknown = [n for n in a.names if n not in REGISTRY]
    if unknown:
        ap.error(f"unknown problem(s) {unknown}; built-in: {sorted(REGISTRY)}")
    for n in a.names or list(REGISTRY):
        generate(n, os.path.join(a.outroot, n))
        print("generated", n)


if __name__ == "__main__":
    main()
