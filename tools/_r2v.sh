mkdir -p gpurun_out/r2v; O=gpurun_out/r2v
for lib in default w12; do
if [ $lib = default ]; then unset FLATDD_B200_LIB; else export FLATDD_B200_LIB=build/variants/$lib.so; fi
python -m pytest tests/test_gpu_block.py tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_$lib.log 2>&1; tail -1 $O/pytest_$lib.log
for t in "3,7,12,20;5,9,14,22" "3,7,12;5,9,14"; do
  python tools/block_ablate.py 26 "$t" >> $O/ablate.txt 2>&1
  FLATDD_OPTS=block_tables_shared=0 python tools/block_ablate.py 26 "$t" >> $O/ablate.txt 2>&1
done
python bench.py > $O/bench_$lib.json 2> $O/bench_$lib.err; python -c "
import json;d=json.loads(open('$O/bench_$lib.json').read().strip().splitlines()[-1]);print('$lib',d['ms_per_step'],d['roofline']['frac'],d['check']['max_amp_err_vs_reference'])"
python tools/pass_log.py > $O/passlog_$lib.txt 2>&1
done
cat $O/ablate.txt
