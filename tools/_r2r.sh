mkdir -p gpurun_out/r2r; O=gpurun_out/r2r
build/dmma_chain_probe > $O/dmma_chain_probe.jsonl 2>&1
cat $O/dmma_chain_probe.jsonl
for lib in slabsep ablate; do
export FLATDD_B200_LIB=build/variants/$lib.so
for t in "3,7,12,20;5,9,14,22" "3,7,12,20" "3,7,12;5,9,14"; do
for s in 0 2; do
  FLATDD_B200_BLOCK_SKIP=$s python tools/block_ablate.py 26 "$t" >> $O/ablate.txt 2>&1
done; done; done
cat $O/ablate.txt
