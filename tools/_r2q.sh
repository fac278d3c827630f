mkdir -p gpurun_out/r2q; O=gpurun_out/r2q
export FLATDD_B200_LIB=build/variants/ablate.so
for t in "3,7,12,20;5,9,14,22" "3,7,12,20"; do
for s in 0 1 2 6 10 14; do
  FLATDD_B200_BLOCK_SKIP=$s python tools/block_ablate.py 26 "$t" >> $O/ablate.txt 2>&1
done; done
cat $O/ablate.txt
unset FLATDD_B200_LIB
ncu --set full --import-source on --clock-control none -k regex:dmavm_block_ws -s 2 -c 1 -o $O/two_block python tools/block_ablate.py 26 "3,7,12,20;5,9,14,22" > $O/ncu.log 2>&1
tail -2 $O/ncu.log
