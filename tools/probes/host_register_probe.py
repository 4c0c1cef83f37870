"""How fast can pageable memory be pinned in place?  cudaHostRegister / cudaHostUnregister of 64 MB chunks of a touched
numpy array, one thread and four threads, GB/s.  (Probe behind DESIGN.md section 8, item 3.)"""
import json
import threading
import time

import numpy as np
import torch

rt = torch.cuda.cudart()
torch.cuda.init()
x = np.ones(1 << 28, dtype=np.float32)          # 1 GiB, touched
chunk = 64 << 20
base = x.ctypes.data
n = x.nbytes // chunk


def reg(i):
    r = rt.cudaHostRegister(base + i * chunk, chunk, 0)
    assert int(r) == 0, r


def unreg(i):
    r = rt.cudaHostUnregister(base + i * chunk)
    assert int(r) == 0, r


out = {}
for threads in (1, 4):
    for name, fn in (("register", reg), ("unregister", unreg)):
        t0 = time.perf_counter()
        if threads == 1:
            for i in range(n):
                fn(i)
        else:
            ths = [threading.Thread(target=lambda k=k: [fn(i) for i in range(k, n, threads)]) for k in range(threads)]
            [t.start() for t in ths]
            [t.join() for t in ths]
        dt = time.perf_counter() - t0
        out[f"{name}_{threads}thread_GBs"] = round(x.nbytes / dt / 1e9, 2)
print(json.dumps(out))
