run() { python scripts/gpu_probe.py $1 2>&1 | tail -1 | python -c "
import sys,re
l=sys.stdin.read()
m=re.search(r'chunks=(\d+): ([\d.]+)s its=(\d+) -> (\d+) it/s', l); bp=re.search(r\"'backpass': \(([\d.]+), (\d+)\)\", l); ls=re.search(r\"'linesearch': \(([\d.]+), (\d+)\)\", l); dv=re.search(r\"'derivs': \(([\d.]+), (\d+)\)\", l)
print('chunks',m.group(1),m.group(2),'s',m.group(4),'it/s | bp %.3f ms x%s | ls %.3f ms x%s | dv %.3f'%(float(bp.group(1))/int(bp.group(2)),bp.group(2),float(ls.group(1))/int(ls.group(2)),ls.group(2),float(dv.group(1))/int(dv.group(2))) if m else l[-300:])"; }
echo "== armijo A/B (B=4096 lane32 / split ppw4; B=65536)"
for L in lib lib_adiv; do
  echo -n "$L B=4096 split0: "; ILQG_LIB_DIR=$PWD/ddp-generator_b200/$L ILQG_BP_SPLIT=0 CHUNKS=1 ITERS=50 run 4096
  echo -n "$L B=4096 split4 ppw4: "; ILQG_LIB_DIR=$PWD/ddp-generator_b200/$L ILQG_BP_SPLIT=4 ILQG_BP_PPW=4 CHUNKS=1 ITERS=50 run 4096
  echo -n "$L B=65536: "; ILQG_LIB_DIR=$PWD/ddp-generator_b200/$L CHUNKS=1 ITERS=30 run 65536
done
echo "== phase stagger"
for ST in 0 1; do
  for CFG in 32768:2 32768:4 32768:1 262144:4; do
    B=${CFG%%:*}; CH=${CFG##*:}
    echo -n "stagger=$ST B=$B: "; ILQG_PHASE_STAGGER=$ST CHUNKS=$CH ITERS=20 run $B
  done
done
