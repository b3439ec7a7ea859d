"""What one cudaMalloc / cudaFree costs on the box (host wall clock, primary context already up, device idle): the reason the streaming mapper
reserves its buffers at creation.  Run under gpurun: python profiles/tools/alloc_cost.py"""
import ctypes as C
import glob
import os
import time

import torch

torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = C.CDLL(cands[0])
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]; rt.cudaFree.argtypes = [C.c_void_p]
print("runtime:", cands[0])
for mb in (1, 8, 64, 512, 4096):
    tm, tf = [], []
    for _ in range(6):
        p = C.c_void_p()
        t0 = time.perf_counter(); e = rt.cudaMalloc(C.byref(p), mb << 20); t1 = time.perf_counter()
        assert e == 0, e
        t2 = time.perf_counter(); rt.cudaFree(p); t3 = time.perf_counter()
        tm.append(1e3 * (t1 - t0)); tf.append(1e3 * (t3 - t2))
    print(f"{mb:5d} MB: cudaMalloc ms {' '.join(f'{x:8.2f}' for x in tm)} | cudaFree ms {' '.join(f'{x:8.2f}' for x in tf)}")
