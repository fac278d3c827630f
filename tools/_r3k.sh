mkdir -p gpurun_out/r3k; O=gpurun_out/r3k
timeout 500 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > $O/pytest_sharded.log 2>&1; tail -5 $O/pytest_sharded.log
