python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tests/dist_peer_exchange_worker.py > /tmp/w8.log 2>&1; echo rc=$?
grep -n "Error\|assert\|Traceback" /tmp/w8.log | head -20
grep -n -B12 "AssertionError" /tmp/w8.log | head -60
