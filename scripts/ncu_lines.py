"""Attribute the per-instruction counters of an ncu report to source lines (development aid).

ncu's source page needs the sources at the path they had on the GPU box; this goes the other way: the SASS of the kernel
is disassembled HERE with line info (nvdisasm --print-line-info on the sm_100a cubin of the in-tree library, built with
-lineinfo) and matched by instruction order with `ncu --page source --csv` of a report captured with `--set full`.  The
library must be the build the report was captured with.

usage: python scripts/ncu_lines.py <report.ncu-rep> <kernel regex> [problem=car] [full_ddp=0] [top=40]
  e.g. python scripts/ncu_lines.py gpurun_out/prof65k_r01.ncu-rep 'k_backpass<ProbCar' car 0
Prints, per source line: share of executed warp instructions, share of stall samples, and for selected lines the
warp-level and thread-level execution counts (divergence shows up as a low ratio of the two)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_with_lines(problem, ddp, mangled_part):
    lib = os.path.join(ROOT, "ddp-generator_b200", "lib", f"libilqg_b200_{problem}_ddp{ddp}.so")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if "sm_100a" in f][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    starts = [i for i, l in enumerate(dis) if l.startswith(".text.") and mangled_part in l]
    if not starts:
        raise SystemExit(f"no function matching {mangled_part!r} in {cubin}")
    cur, seq = None, []
    for l in dis[starts[0] + 1:]:
        if l.startswith("\t.section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            seq.append((m.group(2), cur))
    return seq


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    problem = sys.argv[3] if len(sys.argv) > 3 else "car"
    ddp = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre.split("<")[0]],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    name = rows[0][1]
    hdr, data = rows[1], rows[2:]
    # stop at the next kernel block, if the regex matched several launches
    for i, r in enumerate(data):
        if r and r[0] == "Kernel Name":
            data = data[:i]
            break
    ix, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    # mangled name fragment: kernel identifier + first template argument
    m = re.search(r"(k_\w+)<([^>]*)>", name)
    frag = f"{len(m.group(1))}{m.group(1)}I"
    for a in [t.strip() for t in m.group(2).split(",")]:
        mb = re.fullmatch(r"\((bool|int)\)(-?\d+)", a)
        frag += f"L{'b' if mb.group(1) == 'bool' else 'i'}{mb.group(2)}E" if mb else f"{len(a)}{a}"
    frag += "E"
    seq = sass_with_lines(problem, ddp, frag)
    if len(seq) != len(data):
        print(f"warning: {len(seq)} SASS instructions here, {len(data)} in the report (different build or template instance?)")
    inst, samp, warp, thr = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    for (txt, li), r in zip(seq, data):
        n, s = int(r[ix]), int(r[isamp])
        inst[li] += n
        samp[li] += s
        if n > warp[li]:
            warp[li], thr[li] = n, int(r[it])
    tot, tots = sum(inst.values()), max(sum(samp.values()), 1)
    print(name)
    print(f"{tot} warp instructions, {tots} stall samples")
    src = {}
    for li, n in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
        text = ""
        if li:
            if li[0] not in src:
                cands = [os.path.join(dp, li[0]) for dp, _, fs in os.walk(os.path.join(ROOT, "ddp-generator_b200")) if li[0] in fs]
                src[li[0]] = open(cands[0]).read().split("\n") if cands else []
            if 0 < li[1] <= len(src[li[0]]):
                text = src[li[0]][li[1] - 1].strip()[:80]
        lanes = thr[li] / warp[li] if warp[li] else 0
        print(f"{100 * n / tot:5.1f}% inst {100 * samp[li] / tots:5.1f}% samp  max-exec {warp[li]:>10d} avg lanes {lanes:4.1f}  {li}  {text}")


if __name__ == "__main__":
    main()
