import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle
from wembed_b200 import cabi
from wembed_b200.datasets import degree_weights, heavy_tailed_graph, initial_coordinates
from helpers import lr_exponential
n, d = 20000, 8
edges, _ = heavy_tailed_graph(n, 20, seed=3)
w, x0 = degree_weights(n, edges, d), initial_coordinates(n, d, seed=5)
rp, col = cabi.csr_from_edges(n, edges)
deg = np.diff(rp)
cpu = oracle.CpuEmbedder("port", edges, n=n, embeddingDimension=d, init_state=False)
dev = cabi.DeviceEmbedder(rp, col, embedding_dimension=d, keep_forces=1, seed=1234)
for e in (cpu, dev):
    e.set_weights(w); e.set_coordinates(x0)
for it in range(1, 4):
    flagged = cpu.near_threshold(1e-5)
    cpu.step(); st = dev.step(lr_exponential(it))
    fr, fd, xr, xd = cpu.forces(), dev.forces(), cpu.coordinates(), dev.coordinates()
    xerr = np.abs(xr - xd); ferr = np.abs(fr - fd)
    worst = np.argsort(-xerr.max(1))[:8]
    print(f"it {it} lr {lr_exponential(it):.3f} flagged {flagged.sum()} fscale {np.abs(fr).max():.3g}")
    for v in worst:
        k = xerr[v].argmax()
        print(f"   v={v} deg={deg[v]} w={w[v]:.3f} flagged={flagged[v]} xerr={xerr[v,k]:.3g} f_ref={fr[v,k]:.6g} f_dev={fd[v,k]:.6g} ferr={ferr[v,k]:.3g} |f_ref|max={np.abs(fr[v]).max():.3g}")
    cpu.set_coordinates(xd)
