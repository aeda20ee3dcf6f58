"""Diagnostics: which vertices of a local group's step differ from the single-handle step (forces after step 1)?
usage: python scripts/local_group_probe.py [n] [d] [world] [geometric|heavy]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
os.environ.setdefault("WB_PAIR_CAP", str(400 * n))
from helpers import lr_exponential, make_problem  # noqa: E402
from wembed_b200 import cabi  # noqa: E402

family = sys.argv[4] if len(sys.argv) > 4 else "geometric"
if family == "heavy":
    from wembed_b200.datasets import degree_weights, heavy_tailed_graph, initial_coordinates
    edges, _ = heavy_tailed_graph(n, 20, seed=3)
    w, x0 = degree_weights(n, edges, d), initial_coordinates(n, d, seed=5)
else:
    edges, w, x0 = make_problem(n, d)
rp, col = cabi.csr_from_edges(n, edges)


def fresh():
    dev = cabi.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234, keep_forces=1)
    dev.set_weights(w)
    dev.set_coordinates(x0)
    return dev


single = fresh()
st_ref = single.step(lr_exponential(1))
f_ref = single.forces()
x_ref = single.coordinates()
single.close()
devs = [fresh() for _ in range(world)]
cabi.comm_init_local(devs)
parts = [dv.partition() for dv in devs]
st = cabi.step_group(devs, lr_exponential(1))[0]
f = np.zeros_like(f_ref)
for dv, (b, e) in zip(devs, parts):
    f[b:e] = dv.forces()[b:e]
print("lib", os.environ.get("WB_LIB", "default"), "pairs", st["num_repulsion_pairs"], "ref", st_ref["num_repulsion_pairs"], "listed", st["num_listed_pairs"], st_ref["num_listed_pairs"])
bad = np.nonzero((f != f_ref).any(1))[0]
print("vertices with different force rows:", len(bad), "partition", parts)
if len(bad):
    rows = parts[0][1]
    iw = w ** (-1.0 / d)
    nbr = [set(col[rp[v]:rp[v + 1]].tolist()) for v in bad[:40]]
    for v, nb in zip(bad[:40], nbr):
        dist = np.sqrt(((x0 - x0[v]) ** 2).sum(1)) * iw * iw[v]
        cand = [int(u) for u in np.nonzero(dist <= 1.0)[0] if u != v and int(u) not in nb]
        print(f"v={v} owner={v // rows} deg={rp[v + 1] - rp[v]} partners={len(cand)} partner owners={[u // rows for u in cand]} partners in bad={[u for u in cand if u in set(bad.tolist())]}")
xs = [dv.coordinates() for dv in devs]
deg = np.diff(rp)
print("stats group", st["sum_displacement"], st["sum_radius_sq"], st["centroid"][:d])
print("stats ref  ", st_ref["sum_displacement"], st_ref["sum_radius_sq"], st_ref["centroid"][:d])
for r, xr in enumerate(xs):
    badx = np.nonzero((xr != x_ref).any(1))[0]
    print(f"replica {r}: {len(badx)} rows differ from the single handle; owners {sorted(set((badx // parts[0][1]).tolist()))}; first {badx[:12].tolist()}")
xo = np.zeros_like(x_ref)
for xr, (b, e) in zip(xs, parts):
    xo[b:e] = xr[b:e]
badx = np.nonzero((xo != x_ref).any(1))[0]
print("owners' own rows that differ:", len(badx))
for v in badx[:25]:
    print(f"  v={v} owner={v // parts[0][1]} deg={deg[v]} w/mean={w[v] / w.mean():.1f} |x_ref|={np.linalg.norm(x_ref[v]):.4f} |x_group|={np.linalg.norm(xo[v]):.4f} |dx|={np.linalg.norm(xo[v] - x_ref[v]):.3e} force row equal={bool((f[v] == f_ref[v]).all())}")
for dv in devs:
    dv.close()
