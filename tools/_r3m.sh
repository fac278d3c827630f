mkdir -p gpurun_out/r3m; O=gpurun_out/r3m
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r3m/bench_n8.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['check']['max_amp_err_vs_reference'], {k:d['exchange'][k] for k in ('per_step','fused_into_the_preceding_pass','own_kernel','ms_mean','frac_of_nvlink_nominal_900')})
for w in d.get('extra_workloads',[]):
    print(w['workload'], w['ms_per_step'], w['check'].get('max_amp_err_vs_reference'), w['check'].get('norm2_device'), {k:w['exchange'][k] for k in ('per_step','fused_into_the_preceding_pass','own_kernel','ms_mean','frac_of_nvlink_nominal_900')})
P
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "cli or fused" 2>&1 | tail -2
