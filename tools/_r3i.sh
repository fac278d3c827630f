mkdir -p gpurun_out/r3i; O=gpurun_out/r3i
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; tail -c 600 $O/bench_n8.json
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > $O/pytest_sharded.log 2>&1; tail -2 $O/pytest_sharded.log
