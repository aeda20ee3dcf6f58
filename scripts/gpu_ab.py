"""A/B timing of library builds along a bench workload's trajectory: average phase times over steps w+1..w+k.
usage: gpu_ab.py WORKLOAD WARMUP STEPS lib1.so [lib2.so ...]   (each build runs in its own process: WB_LIB is read at import)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("WB_AB_CHILD"):
    import numpy as np
    import bench
    from wembed_b200 import cabi
    name, warm, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    cache = f"/tmp/ab_{name}.npz"
    if os.path.exists(cache):
        z = np.load(cache); wl = {k: z[k] for k in z.files}; wl["d"] = int(wl["d"])
    else:
        wl = bench.make_workload(name)
        np.savez(cache, row_ptr=wl["row_ptr"], col=wl["col"], weights=wl["weights"], x0=wl["x0"], d=wl["d"])
    dev = cabi.DeviceEmbedder(wl["row_ptr"], wl["col"], embedding_dimension=wl["d"], seed=1234)
    dev.set_weights(wl["weights"]); dev.set_coordinates(wl["x0"]); dev.enable_timing(True)
    acc = {}
    for it in range(1, warm + steps + 1):
        st = dev.step(bench.lr_schedule(it))
        if it > warm:
            for k, v in dev.phase_times().items(): acc[k] = acc.get(k, 0.0) + v / steps
    print(os.path.basename(os.environ.get("WB_LIB", "default")), name, " ".join(f"{k} {v:.4f}" for k, v in acc.items()),
          f"| pairs {st['num_repulsion_pairs']:.0f} pt/v {st['num_candidates']/len(wl['weights']):.0f} box/v {st['num_box_tests']/len(wl['weights']):.0f} lossA {st['loss_attract']:.8g}", flush=True)
else:
    for lib in sys.argv[4:]:
        env = dict(os.environ, WB_AB_CHILD="1", WB_LIB=os.path.join(ROOT, lib))
        subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:4], env=env, check=False)
