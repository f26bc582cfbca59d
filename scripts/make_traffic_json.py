"""profiles/traffic.json from an ncu launch list with dram__bytes_{read,write}.sum (scripts/gpu_profile.sh step 2).
Per kernel class: DRAM bytes per pass and per bench-timed launch (bench.py times each k_ls_round on its own and
k_ls_tail + k_ls_commit as one launch).  usage: make_traffic_json.py traffic.csv out.json "<how it was captured>" """
import csv, json, sys
from collections import OrderedDict
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
L = OrderedDict()
for r in rows:
    L.setdefault(r["ID"], {"k": r["Kernel Name"]})[r["Metric Name"]] = float(r["Metric Value"])
launches = list(L.values())
cls = lambda k: "derivs" if "k_derivs" in k else "backpass" if "k_backpass" in k else "linesearch"
# whole passes only: from the first k_derivs to the last launch before the final k_derivs
idx = [i for i, l in enumerate(launches) if "k_derivs" in l["k"]]
sel = launches[idx[0]:idx[-1]]
n_pass = len(idx) - 1
out = {"source": sys.argv[3], "passes_averaged": n_pass}
for c in ("derivs", "backpass", "linesearch"):
    ls = [l for l in sel if cls(l["k"]) == c]
    byt = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in ls)
    n_timed = sum(1 for l in ls if "k_ls_commit" not in l["k"])   # tail + commit_seg + commit are one timed launch
    out[c] = {"dram_bytes_per_pass": byt / n_pass, "kernels_per_pass": len(ls) / n_pass, "bench_launches_per_pass": n_timed / n_pass,
              "dram_bytes_per_launch": byt / max(n_timed, 1), "ms_per_pass_under_ncu": sum(l["gpu__time_duration.sum"] for l in ls) / n_pass / 1e6}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
