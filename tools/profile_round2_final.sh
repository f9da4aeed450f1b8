# final measurements of the round on one GPU: the GPU test suite, smoke, the default bench line, the update's timeline,
# and the ncu --set full capture of the env kernel as shipped (exported to CSV on the box)
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_tests.log 2>&1; tail -3 gpurun_out/r02_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -2 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/r02_bench_default_1gpu.json 2> gpurun_out/r02_bench_default_1gpu.err; tail -c 300 gpurun_out/r02_bench_default_1gpu.json
python tools/timeline_probe.py > gpurun_out/r02_timeline_update_b256.txt 2>&1
python tools/env_phase_probe.py > gpurun_out/r02_env_phase_probe.json 2>/dev/null
python bench.py --steps 20 --warmup 5 --no-learner --no-cpu-baseline > gpurun_out/r02_plain_env.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 8 -c 2 -f -o gpurun_out/r02_env_step python bench.py --steps 20 --warmup 5 --no-learner --no-cpu-baseline > gpurun_out/r02_ncu_env.log 2>&1
ncu -i gpurun_out/r02_env_step.ncu-rep --page raw --csv > gpurun_out/r02_env_step_raw.csv 2>/dev/null
ls -la gpurun_out/r02_env_step*
# every tensor-core / reduction kernel of one update, full set (report exported to CSV here: it is too large to travel back)
python tools/learner_probe.py --updates 4 --warmup 3 --precision fp16 > gpurun_out/r02_plain_learner.log 2>&1 && \
ncu --set full --clock-control none -k regex:'tc_|fc1_head|finalize|adam_wf1|colsum|pack_x2|gather|sample_uniform' -s 60 -c 26 -o /tmp/r02_learner_step python tools/learner_probe.py --updates 4 --warmup 3 --precision fp16 > gpurun_out/r02_ncu_learner.log 2>&1
ncu -i /tmp/r02_learner_step.ncu-rep --page raw --csv > gpurun_out/r02_learner_step_raw.csv 2>/dev/null
ls -la gpurun_out/r02_learner_step_raw.csv
