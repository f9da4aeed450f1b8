#!/usr/bin/env python
"""cycles per tcgen05.mma (M=128, K=16, bf16) vs N, operand layout and number of independent accumulators"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dqnflappybird_b200 import _lib
L = _lib.lib()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
iters = 2048
print("N  major  accs  same_ops  cycles/MMA(done)  cycles/MMA(issue)   ideal(N/2)")
for mn in (0, 1):
    for n in (32, 64, 128, 256):
        for naccs in (1, 2, 4):
            if naccs * n > 512:
                continue
            for same in (0, 1):
                for _ in range(2):
                    _lib.check(L.fb_debug_tc_mma_rate(n, mn, naccs, iters, same, out.data_ptr(), st), "rate")
                    torch.cuda.synchronize()
                print(f"{n:4d} {'MN' if mn else 'K ':>4s} {naccs:5d} {same:8d} {out[0].item() / iters:10.1f} {out[1].item() / iters:10.1f} {n / 2:10.1f}", flush=True)
