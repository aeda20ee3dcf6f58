"""First-light GPU check: one step from identical state, device vs CPU port (prints diagnostics)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle
from wembed_b200 import cabi
from helpers import make_problem, lr_exponential, near_threshold_vertices, ring_graph

oracle.build("port")
print(cabi.lib().wb_build_info(), "devices", cabi.lib().wb_device_count(), flush=True)

def compare(n, d, steps=3):
    edges, w, x0 = make_problem(n, d)
    rp, col = cabi.csr_from_edges(n, edges)
    cpu = oracle.CpuEmbedder("port", edges, n=n, embeddingDimension=d, init_state=False)
    assert np.array_equal(cpu.csr()[0], rp) and np.array_equal(cpu.csr()[1], col)
    cpu.set_weights(w); cpu.set_coordinates(x0)
    dev = cabi.DeviceEmbedder(rp, col, embedding_dimension=d, keep_forces=1, seed=1234)
    dev.set_weights(w); dev.set_coordinates(x0)
    dev.enable_timing(True)
    for it in range(1, steps + 1):
        xin = cpu.coordinates()
        cpu.step()
        st = dev.step(lr_exponential(it))
        cs = cpu.stats()
        fr, fd = cpu.forces(), dev.forces()
        xr, xd = cpu.coordinates(), dev.coordinates()
        flagged = near_threshold_vertices(xin, w, rp, col)
        ok = ~flagged
        ferr = np.abs(fr - fd).max(1)
        fscale = np.abs(fr).max()
        xerr = np.abs(xr - xd).max(1)
        print(f"n={n} d={d} it={it}: lossA {cs['loss_attract']:.9g} vs {st['loss_attract']:.9g} | lossR {cs['loss_repel']:.9g} vs {st['loss_repel']:.9g}"
              f" | pairs cpu {cs['num_rep_pairs']:.0f} dev {st['num_repulsion_pairs']:.0f} tests/v {st['num_candidates']/n:.0f}"
              f" | flagged {flagged.sum()} | force err max(all) {ferr.max()/fscale:.3g} max(unflagged) {ferr[ok].max()/fscale:.3g}"
              f" | coord err max(unflagged) {xerr[ok].max():.3g} p99.9 {np.quantile(xerr[ok],0.999):.3g} | reldisp {cs['rel_displacement']:.9g} vs {st['rel_displacement']:.9g}"
              f" | phases {dev.phase_times()}", flush=True)
        # keep both on the same trajectory
        cpu.set_coordinates(xd)
    return dev

compare(2000, 4)
if os.environ.get("WB_QUICK"): sys.exit(0)
compare(20000, 4)
compare(20000, 8)
compare(5000, 2)
compare(3000, 3)
compare(3000, 16)

# timing at scale
for n, d in ((100000, 4), (1000000, 8)):
    edges, w, x0 = make_problem(n, d)
    rp, col = cabi.csr_from_edges(n, edges)
    dev = cabi.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
    dev.set_weights(w); dev.set_coordinates(x0); dev.enable_timing(True)
    for it in range(1, 61):
        t = time.time(); st = dev.step(lr_exponential(it)); dt = time.time() - t
        if it <= 6 or it % 10 == 0:
            print(f"n={n} d={d} it={it} wall {dt*1e3:.2f} ms pairs/v {st['num_repulsion_pairs']/n:.2f} tests/v {st['num_candidates']/n:.0f} lossA {st['loss_attract']:.4g} lossR {st['loss_repel']:.4g} {dev.phase_times()}", flush=True)
