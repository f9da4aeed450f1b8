python -m pytest tests/test_qnet_tc_gpu.py -x -q 2>&1 | tail -3
for c in 1 2; do python tools/learner_probe.py --updates 600 --warmup 20 --precision fp16 2>&1 | tail -1 | cut -c1-110; done
python tools/timeline_probe.py 2>&1 | sed -n 13,26p | cut -c1-100
