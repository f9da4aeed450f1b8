#!/usr/bin/env python
"""Does a tcgen05 K-major SWIZZLE_128B descriptor started `shift` rows into a TMA-written slab read rows shift..shift+127?"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dqnflappybird_b200 import _lib
L = _lib.lib()
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn((256, 64), device="cuda", generator=g).to(torch.bfloat16)
b = torch.randn((64, 64), device="cuda", generator=g).to(torch.bfloat16)
for shift in (0, 1, 2, 5, 7, 8, 9, 21, 22, 64, 100, 128):
    ref = a[shift:shift + 128].double() @ b.double().T
    out = []
    for bo in sorted({0, shift & 7}):
        d = torch.full((128, 64), float("nan"), device="cuda")
        _lib.check(L.fb_debug_tc_slab(shift, bo, a.data_ptr(), b.data_ptr(), d.data_ptr(), torch.cuda.current_stream().cuda_stream), "slab")
        torch.cuda.synchronize()
        out.append((bo, float((d.double() - ref).abs().max())))
    print("shift", shift, " (base_offset, max err):", out, flush=True)
