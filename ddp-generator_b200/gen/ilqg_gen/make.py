"""One-step build of a problem, the counterpart of the reference's `make_iLQG('optDefCar', '-DFULL_DDP=0')` (make_iLQG.m:1-86):
problem description -> generated code -> compiled libraries.

    python -m ilqg_gen.make <problem.mac | built-in name> [--name NAME] [--out DIR] [--flags "..."]

* a `.mac` file in the reference's format (README.md:25-38) or the name of a module in ilqg_gen.problems;
* generated code goes to DIR/problems/<name>/ (iLQG_problem.h, iLQG_func.c, <name>_device.cuh);
* libraries go to DIR/lib/: libilqg_b200_<name>_ddp{0,1}.so (batched C ABI, include/ilqg_b200.h) and
  libilqg_dropin_<name>_ddp{0,1}.so (the reference's single-problem entry points); both FULL_DDP settings are built,
  where the reference selects one with -DFULL_DDP (make_iLQG.m:2);
* DIR defaults to the package directory (ddp-generator_b200/), which is where ilqg_b200.Library looks;
* --flags are passed to nvcc (the reference's second argument: extra compiler switches);
* --mex additionally builds the MATLAB / Octave gateway mex/iLQG_mex_b200.c against the FULL_DDP=<d> library with `mkoctfile --mex`
  (or `mex`), as make_iLQG.m:65-86 does for the reference; it needs one of the two on PATH.
Needs nvcc (sm_100a cross-compiles without a GPU) and gcc; there is no Maxima / gentran step.
"""
from __future__ import annotations

import argparse
import os
import re
import shutil
import subprocess
import sys

from . import __main__ as gen
from .problems import REGISTRY

PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build(source, name=None, out=None, flags="", jobs=8, quiet=True):
    """Returns {"name", "struct", "problem_dir", "libs": [...]}."""
    out = os.path.abspath(out or PKG)
    is_mac = source.endswith(".mac")
    if not is_mac and source not in REGISTRY:
        raise SystemExit(f"{source!r} is neither a .mac file nor one of the built-in problems {sorted(REGISTRY)}")
    if name is None:
        name = re.sub(r"^optdef", "", os.path.basename(source)[:-4].lower()) if is_mac else source
    if not re.fullmatch(r"[a-z][a-z0-9_]*", name):
        raise SystemExit(f"problem name {name!r} must be a lower-case C identifier")
    pdir = os.path.join(out, "problems", name)
    gen.generate(name, pdir, mac=source if is_mac else None)
    struct = re.search(r"struct (Prob\w+)", open(os.path.join(pdir, f"{name}_device.cuh")).read()).group(1)
    libdir, builddir = os.path.join(out, "lib"), os.path.join(out, "build")
    targets = [os.path.join(libdir, f"lib{kind}_{name}_ddp{d}.so") for kind in ("ilqg_b200", "ilqg_dropin") for d in (0, 1)]
    cmd = ["make", "-C", PKG, f"-j{jobs}", f"PROBLEMS={name}", f"STRUCT_{name}={struct}", f"PROBDIR={os.path.join(out, 'problems')}",
           f"LIBDIR={libdir}", f"BUILDDIR={builddir}", f"XFLAGS={flags}", *targets]
    if quiet:
        cmd.insert(1, "-s")
    subprocess.run(cmd, check=True)
    return {"name": name, "struct": struct, "problem_dir": pdir, "libs": targets}


def build_mex(name, full_ddp, out=None):
    """The gateway as iLQG<Name>.mex / .mexa64 in DIR/lib (naming of make_iLQG.m:84)."""
    out = os.path.abspath(out or PKG)
    libdir = os.path.join(out, "lib")
    tool = shutil.which("mkoctfile") or shutil.which("mex")
    if not tool:
        raise SystemExit("--mex needs mkoctfile (Octave) or mex (MATLAB) on PATH")
    src = os.path.join(PKG, "mex", "iLQG_mex_b200.c")
    target = os.path.join(libdir, f"iLQG{name.capitalize()}")
    args = [tool] + (["--mex"] if tool.endswith("mkoctfile") else []) + [src, "-I" + os.path.join(os.path.dirname(PKG), "include"), "-L" + libdir,
            f"-lilqg_b200_{name}_ddp{int(full_ddp)}", f"-Wl,-rpath,{libdir}", "-o" if tool.endswith("mkoctfile") else "-output", target]
    subprocess.run(args, check=True)
    return target


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m ilqg_gen.make", description=__doc__.split("\n\n")[0])
    ap.add_argument("source", help="problem.mac or a built-in problem name")
    ap.add_argument("--name", help="problem name (default: from the file name, optDefCar.mac -> car)")
    ap.add_argument("--out", help="output root (default: the package directory)")
    ap.add_argument("--flags", default="", help="extra nvcc flags")
    ap.add_argument("--mex", type=int, choices=[0, 1], metavar="FULL_DDP", help="also build the mex gateway against the FULL_DDP=0|1 library")
    a = ap.parse_args(argv)
    r = build(a.source, a.name, a.out, a.flags, quiet=False)
    print(f"problem {r['name']} ({r['struct']}): code in {r['problem_dir']}")
    for l in r["libs"]:
        print("  ", l)
    if a.mex is not None:
        print("  ", build_mex(r["name"], a.mex, a.out))


if __name__ == "__main__":
    main(sys.argv[1:])
