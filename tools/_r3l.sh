mkdir -p gpurun_out/r3l; O=gpurun_out/r3l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 300 $O/bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r3l/bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['check']['max_amp_err_vs_reference'], d.get('exchange'), d['e2e'])
for w in d.get('extra_workloads',[]):
    print(w['workload'], w['ms_per_step'], w['check']['max_amp_err_vs_reference'], w.get('exchange'))
P
