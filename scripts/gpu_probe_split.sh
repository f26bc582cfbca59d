# backward pass: lanes per problem (1 = k_backpass, 4 = k_backpass_split) x problems per warp, by batch size (development aid)
for B in ${BATCHES:-4096 16384 32768}; do
  for CFG in ${CFGS:-0:32 0:16 0:8 0:4 0:2 4:8 4:4 4:2 4:1}; do
    SP=${CFG%%:*}; PPW=${CFG##*:}
    echo -n "B=$B split=$SP ppw=$PPW: "; ILQG_BP_SPLIT=$SP ILQG_BP_PPW=$PPW CHUNKS=${CHUNKS:-1} ITERS=${ITERS:-50} python scripts/gpu_probe.py $B 2>&1 | tail -1 | python -c "
import sys,re
l=sys.stdin.read()
m=re.search(r'-> (\d+) it/s', l); bp=re.search(r\"'backpass': \(([\d.]+), (\d+)\)\", l); ls=re.search(r\"'linesearch': \(([\d.]+), (\d+)\)\", l)
print(m.group(1) if m else l[:200], 'it/s  bp ms/launch %.3f' % (float(bp.group(1))/int(bp.group(2))) if bp else '', ' ls ms total %.1f'%float(ls.group(1)) if ls else '')"
  done
done
