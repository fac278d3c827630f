mkdir -p gpurun_out/r3n; O=gpurun_out/r3n
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
python bench.py > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.json
SANITIZE_ONLY_NEW=2 timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_run.py > $O/sanitizer_racecheck.log 2>&1; tail -1 $O/sanitizer_racecheck.log
SANITIZE_ONLY_NEW=2 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/sanitizer_memcheck.log 2>&1; tail -1 $O/sanitizer_memcheck.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/b_under_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:dmavm_block_ws -s 3 -c 2 -o $O/bench_passes python bench.py --steps 1 --warmup 1 --no-cpu > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log
ncu --set full --clock-control none -k regex:convert_kernel -s 1 -c 1 -o $O/convert python bench.py --steps 1 --warmup 1 --no-cpu > $O/ncu_conv.log 2>&1; tail -1 $O/ncu_conv.log
