# line-search mode sweep (development aid): parallel commit on/off x tail start, per batch size
for B in ${BATCHES:-4096 32768 65536}; do
  for PAR in 0 1; do
    for FROM in ${FROMS:--1 0 1 2 3}; do
      if [ "$FROM" = "-1" ]; then unset ILQG_LS_TAIL_FROM; else export ILQG_LS_TAIL_FROM=$FROM; fi
      echo "B=$B PAR=$PAR FROM=$FROM"
      ILQG_LS_COMMIT_PAR=$PAR CHUNKS=${CHUNKS:-1} ITERS=${ITERS:-30} python scripts/gpu_probe.py $B 2>&1 | tail -1 | cut -c1-330
    done
  done
done
