import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from wembed_b200 import cabi
from helpers import make_problem, lr_exponential
n, d = 1000000, 8
edges, w, x0 = make_problem(n, d)
rp, col = cabi.csr_from_edges(n, edges)
def fresh():
    dev = cabi.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
    dev.set_weights(w); dev.set_coordinates(x0)
    for it in range(1, 21): dev.step(lr_exponential(it))
    return dev
dev = fresh()
t = time.perf_counter(); x = dev.coordinates(); t_get = time.perf_counter() - t
t = time.perf_counter(); dev.set_coordinates(x); t_set = time.perf_counter() - t
t = time.perf_counter()
for it in range(21, 121): dev.step(lr_exponential(it))
t_block = time.perf_counter() - t
dev.close()
dev = fresh()
t = time.perf_counter()
for it in range(21, 121): dev.step_async(lr_exponential(it)) if it - 21 < 60 else None
t_enq = time.perf_counter() - t
for it in range(21, 81): dev.step_collect()
t_async = time.perf_counter() - t
print(f"get {t_get*1e3:.1f} ms set {t_set*1e3:.1f} ms | 100 blocking steps {t_block*1e3:.1f} ms | 60 async: enqueue {t_enq*1e3:.1f} ms total {t_async*1e3:.1f} ms")
