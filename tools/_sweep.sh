python tools/timeline_probe.py 2>&1 | sed -n 3,26p | cut -c1-100
