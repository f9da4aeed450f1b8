for m in 1024 2048 4096 8192 16384; do
  echo "act batch $m: $(python tools/learner_probe.py --envs 16384 --max-act-batch $m 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_act"], "%.3e" % d["act_envs_per_s"])')"
done
