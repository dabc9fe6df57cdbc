#!/usr/bin/env python3
"""Debug aid: per-phase clock64() timeline of CTA 0 of the row-split kernel.
Build the library with -DTE_TIMELINE first:  NVCC_EXTRA=-DTE_TIMELINE python -c "import __graft_entry__ as g; g.build(force=True)"
Marks: 0 loop top | 1 after store-drain wait + next load issue | 2 tile arrived | 3 before barrier 1 | 4 after |
       5 before barrier 2 | 6 after | 7 before final barrier."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import target_estimation_b200 as te

model = sys.argv[1] if len(sys.argv) > 1 else "angular_rates"
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 1 << 20
mtype, _, Q, R, P0 = te.load_model(model)
pool = te.TargetPool(mtype)
pool.set_variant(variant)
pool.register_class(Q, R, P0)
rng = np.random.default_rng(0)
p0 = np.zeros((n, 7)); p0[:, :3] = rng.uniform(-5, 5, (n, 3)); p0[:, 6] = 1
pool.add(np.arange(n, dtype=np.uint32), p0)
meas = torch.from_numpy(p0).cuda()
for _ in range(3):
    pool.step_dense(0.004, meas, 7, None, te.ACT_UPDATE)
pool.sync()
out = np.zeros(2 * 64 * 12, dtype=np.int64)
assert te.lib.te_debug_timeline(out.ctypes.data_as(C.c_void_p)) == 0
T = out.reshape(2, 64, 12)
print("tile period (cycles): median %.0f" % np.median(np.diff(T[0, 4:60, 0])))
for wi, name in ((0, "warp 0 (producer, angle conversion, v)"), (1, "warp 3 (W columns)")):
    t = T[wi]
    print(name)
    marks = ["top", "after STAGES==1 issue", "tile arrived", "before barrier 1", "after barrier 1", "after drain+issue",
             "after conversion | factor", "after bar.sync 2 | (unused)", "before barrier 2", "after barrier 2", "before final barrier"]
    base = t[4:60, 0:1]
    rel = np.median(t[4:60, :11] - base, axis=0)
    for k, nm in enumerate(marks):
        print("  %-32s t = %7.0f" % (nm, rel[k]))
print("raw, relative to warp 0's top of tile 20:")
b = T[0, 20, 0]
for it in (20, 21, 22):
    for wi in (0, 1):
        print("  tile %d warp %d: %s" % (it, 0 if wi == 0 else 3, " ".join("%6d" % (T[wi, it, k] - b) if k != 7 or wi == 0 else "     -" for k in range(11))))
