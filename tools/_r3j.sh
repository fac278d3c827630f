mkdir -p gpurun_out/r3j; O=gpurun_out/r3j
python -m pytest tests/test_gpu_standalone.py -m gpu -x -q > $O/pytest_standalone.log 2>&1; tail -2 $O/pytest_standalone.log
python tools/run_all_circuits.py 3 standalone > $O/all_circuits_standalone_fuse3.jsonl 2> $O/all.err
python - <<'P'
import json
for l in open("gpurun_out/r3j/all_circuits_standalone_fuse3.jsonl"):
    try: d=json.loads(l)
    except Exception: continue
    print({k:d.get(k) for k in ("circuit","array_phase_time","gate_merging_time","simulation_time","array_phase_launches","gpu_kernel_launches") if k in d})
P
