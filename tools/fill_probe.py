import torch
x = torch.empty(131072*25600, dtype=torch.uint8, device="cuda")
y = x[:131072*6400*1]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, t in (("fill 839MB (torch fill_)", y), ("fill 3.36GB", x)):
    for _ in range(3): t.fill_(7)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): t.fill_(7)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/20
    print(name, f"{ms*1e3:.1f} us  {t.numel()/ms/1e6:.1f} GB/s")
# strided fill like the env kernel: every env writes one 6400-B frame out of a 25600-B ring row
z = x.view(131072, 4, 6400)
for _ in range(3): z[:, 1].fill_(3)
torch.cuda.synchronize(); e0.record()
for k in range(20): z[:, k % 4].fill_(3)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/20
print("strided frame fill", f"{ms*1e3:.1f} us  {131072*6400/ms/1e6:.1f} GB/s")
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemsetAsync.argtypes=[ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
for _ in range(3): rt.cudaMemsetAsync(x.data_ptr(), 1, 131072*6400, None)
torch.cuda.synchronize(); e0.record()
for _ in range(20): rt.cudaMemsetAsync(x.data_ptr(), 1, 131072*6400, None)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/20
print("cudaMemsetAsync 839MB", f"{ms*1e3:.1f} us  {131072*6400/ms/1e6:.1f} GB/s")
