mkdir -p gpurun_out/r2t; O=gpurun_out/r2t
for lib in default w12; do
if [ $lib = default ]; then unset FLATDD_B200_LIB; else export FLATDD_B200_LIB=build/variants/$lib.so; fi
for t in "3,7,12,20;5,9,14,22" "3,7,12,20"; do
  python tools/block_ablate.py 26 "$t" >> $O/ablate.txt 2>&1
  FLATDD_OPTS=block_buffers=2 python tools/block_ablate.py 26 "$t" >> $O/ablate.txt 2>&1
  FLATDD_B200_CARVEOUT=60 FLATDD_OPTS=block_buffers=2 python tools/block_ablate.py 26 "$t" >> $O/ablate.txt 2>&1
  FLATDD_B200_CARVEOUT=75 FLATDD_OPTS=block_buffers=2 python tools/block_ablate.py 26 "$t" >> $O/ablate.txt 2>&1
  FLATDD_B200_CARVEOUT=100 FLATDD_OPTS=block_buffers=2 python tools/block_ablate.py 26 "$t" >> $O/ablate.txt 2>&1
done; done
cat $O/ablate.txt
