python -m pytest tests/test_dist_gpu.py -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 20 --no-closed-loop > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -c 300 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','n_gpus','updates_per_s','learner_ms_per_update','learner_weak_efficiency','exchange_us','learner_checks']})
PY
python tools/timeline_probe.py prioritydqn 2>&1 | sed -n 3,28p | cut -c1-100
