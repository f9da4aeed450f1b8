#!/usr/bin/env python
"""Write-only HBM bandwidth of this B200: what a pure store stream (the env kernel writes 6,400 of its 6,474 bytes per frame) can
reach, next to the read+write copy figure MEASURED_PEAKS.json uses as the roofline denominator."""
import json
import sys

import torch

def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

n = 131072 * 25600           # the bench ring: 3.36 GB
x = torch.empty(n, dtype=torch.uint8, device="cuda")
y = torch.empty(n, dtype=torch.uint8, device="cuda")
xi = x.view(torch.int32)
out = {}
ms = t(lambda: x.fill_(255)); out["fill_u8_gbs"] = n / ms / 1e6
ms = t(lambda: xi.fill_(-1)); out["fill_i32_gbs"] = n / ms / 1e6
ms = t(lambda: torch.cuda.current_stream().synchronize() or x.zero_()); out["zero_gbs"] = n / ms / 1e6
ms = t(lambda: y.copy_(x)); out["copy_gbs_read_plus_write"] = 2 * n / ms / 1e6
# strided like the ring: one 6,400-byte frame of every 25,600
xr = x.view(131072, 4, 6400)
ms = t(lambda: xr[:, 1].fill_(255)); out["fill_one_slot_of_four_gbs"] = 131072 * 6400 / ms / 1e6
# the env kernel's own pattern without its computation (fb_debug_write_probe): 131,072 frames of 6,400 bytes
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dqnflappybird_b200 import _lib
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
for name, stride in (("ring_env_major_stride_25600", 25600), ("contiguous_stride_6400", 6400)):
    for mode, mname in ((0, "stg128"), (1, "bulk")):
        for ctas in (296, 592, 1184):
            ms = t(lambda: _lib.check(L.fb_debug_write_probe(x.data_ptr(), 131072, stride, mode, ctas, st), "probe"))
            out[f"{name}_{mname}_{ctas}ctas_gbs"] = 131072 * 6400 / ms / 1e6
print(json.dumps(out))
