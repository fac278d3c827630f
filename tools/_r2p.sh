mkdir -p gpurun_out/r2p; O=gpurun_out/r2p
for lib in order0 default; do
  if [ $lib = default ]; then unset FLATDD_B200_LIB; else export FLATDD_B200_LIB=build/variants/$lib.so; fi
  for t in "3,7,12,20" "0,2,9,15" "6,7,8,9" "3,7,12,20;5,9,14,22" "3,7,12;5,9,14" "3,7"; do
    python tools/block_ablate.py 26 "$t" >> $O/ablate.txt 2>&1
  done
  python bench.py > $O/bench_$lib.json 2> $O/bench_$lib.err
done
cat $O/ablate.txt
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
