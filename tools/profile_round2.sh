set -x
python tools/timeline_probe.py > gpurun_out/r02_timeline_update_b256.txt 2>&1
python tools/write_bw_probe.py > gpurun_out/r02_write_bw_probe.json 2>/dev/null
python tools/env_phase_probe.py > gpurun_out/r02_env_phase_probe.json 2>/dev/null
# env kernel: full capture of the final kernel
python bench.py --steps 20 --warmup 5 --no-learner --no-cpu-baseline > gpurun_out/r02_plain_env.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 30 -c 2 -o gpurun_out/r02_env_step python bench.py --steps 20 --warmup 5 --no-learner --no-cpu-baseline > gpurun_out/r02_ncu_env.log 2>&1
# launch list of the default bench command (short)
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-learner-variants --no-closed-loop > gpurun_out/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-learner-variants --no-closed-loop > gpurun_out/r02_ncu_bench.log 2>&1
# the learner step: every kernel of two updates, full set
python tools/learner_probe.py --updates 4 --warmup 3 > gpurun_out/r02_plain_learner.log 2>&1 && \
ncu --set full --clock-control none -k regex:'tc_|fc1_head|finalize|adam_wf1|colsum|pack_x2|gather|sample_uniform' -s 200 -c 26 -o /tmp/r02_learner_step python tools/learner_probe.py --updates 4 --warmup 3 > gpurun_out/r02_ncu_learner.log 2>&1
# the report itself is too large to travel back (gpurun_out/ is capped at 64 MiB): export what the summaries need here
ncu -i /tmp/r02_learner_step.ncu-rep --page raw --csv > gpurun_out/r02_learner_step_raw.csv 2>/dev/null
python tools/timeline_big_probe.py > gpurun_out/r02_timeline_update_b4096.txt 2>&1
ls -la gpurun_out/r02_* /tmp/r02_learner_step.ncu-rep; du -sh gpurun_out
