# A/B of two builds of the car library (development aid): default lib vs $ALT (a LIBDIR under ddp-generator_b200/)
ALT=${ALT:-lib_noshare}
for B in ${BATCHES:-4096 32768 65536}; do
  for IT in ${ITERSET:-20 50}; do
    echo "B=$B ITERS=$IT default"; CHUNKS=1 ITERS=$IT python scripts/gpu_probe.py $B 2>&1 | tail -1 | cut -c1-230
    echo "B=$B ITERS=$IT $ALT"; ILQG_LIB_DIR=$PWD/ddp-generator_b200/$ALT CHUNKS=1 ITERS=$IT python scripts/gpu_probe.py $B 2>&1 | tail -1 | cut -c1-230
  done
done
