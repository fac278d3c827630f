mkdir -p gpurun_out/r2s; O=gpurun_out/r2s
export FLATDD_B200_LIB=build/variants/w12.so
python -m pytest tests/test_gpu_block.py tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_w12.log 2>&1; tail -3 $O/pytest_w12.log
for lib in w12 default; do
if [ $lib = default ]; then unset FLATDD_B200_LIB; else export FLATDD_B200_LIB=build/variants/$lib.so; fi
for t in "3,7,12,20;5,9,14,22" "3,7,12,20" "3,7,12;5,9,14" "3,7,12,20;5,9,14,22;6,10,15,21" "0,1,2,3;5,9,14,22"; do
for s in 0 2; do
  FLATDD_B200_BLOCK_SKIP=$s python tools/block_ablate.py 26 "$t" >> $O/ablate.txt 2>&1
done; done; done
cat $O/ablate.txt
export FLATDD_B200_LIB=build/variants/w12.so
python bench.py > $O/bench_w12.json 2> $O/bench_w12.err; python -c "
import json;d=json.loads(open('$O/bench_w12.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['roofline']['frac'],d['check'])"
