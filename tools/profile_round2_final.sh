# final measurements of the round on one GPU: the GPU test suite, the default bench line, the update's timeline
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_tests.log 2>&1; tail -3 gpurun_out/r02_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -2 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/r02_bench_default_1gpu.json 2> gpurun_out/r02_bench_default_1gpu.err; tail -c 600 gpurun_out/r02_bench_default_1gpu.json
python tools/timeline_probe.py > gpurun_out/r02_timeline_update_b256.txt 2>&1
