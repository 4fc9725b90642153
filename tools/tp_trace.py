#!/usr/bin/env python
"""Dev tool: clock64 timeline of one CTA of the thread-per-tile kernel (harness built with -DMGP_TP_TRACE)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muygpys_b200.neighbors import NN_Wrapper
rng = np.random.default_rng(7)
n, b, k = 1_000_000, 100_000, 50
x = torch.as_tensor(rng.uniform(size=(n, 2))).cuda(); y = torch.as_tensor(rng.normal(size=n)).cuda()
q = torch.as_tensor(rng.uniform(size=(b, 2))).cuda()
nn, _ = NN_Wrapper(x, k).get_nns(q)
lib = C.CDLL(os.path.join(os.path.dirname(__file__), "..", "muygpys_b200", "csrc", "build", sys.argv[1]))
lib.one_run.argtypes = [C.c_void_p] * 4 + [C.c_longlong, C.c_longlong, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
mean = torch.empty(b, dtype=torch.float64, device="cuda"); var = torch.empty_like(mean)
for _ in range(3):
    assert lib.one_run(x.data_ptr(), q.data_ptr(), nn.data_ptr(), y.data_ptr(), n, b, k, 0.1, 1e-3, mean.data_ptr(), var.data_ptr(), None) == 0
torch.cuda.synchronize()
tr = np.zeros((32, 64), dtype=np.int64)
lib.one_trace.argtypes = [C.c_void_p]
assert lib.one_trace(tr.ctypes.data) == 0
U = int(sys.argv[2]) if len(sys.argv) > 2 else 15
t0 = tr[:U, 3].min()
names = {3: "iter start", 0: "compact done", 1: "col0 built", 2: "prev last col synced"}
for J in range(6):
    names[4 * J + 4] = f"col{J+1} built"; names[4 * J + 5] = f"sync(2) col{J}"; names[4 * J + 6] = f"finish col{J}"
order = [3, 0, 1, 2] + [4 * J + o for J in range(6) for o in (4, 5, 6)]
print("event               " + " ".join(f"w{w:02d}" for w in range(U)))
for e in order:
    print(f"{names[e]:20s}" + " ".join(f"{(tr[w, e] - t0):5d}" if tr[w, e] else "    -" for w in range(U)))
print("hand-off of col J+1 (arrive):")
for J in range(6):
    print(f"col{J+1:1d} handed off     " + " ".join(f"{(tr[w, 32 + J] - t0):5d}" for w in range(U)))
print("factor warp (start, end) per column:")
for J in range(7):
    print(J, tr[U, 2 * J] - t0, tr[U, 2 * J + 1] - t0, "L =", tr[U, 2 * J + 1] - tr[U, 2 * J])

print("\nsummary (cycles from iteration start; u = update warps):")
fe = [tr[U, 2 * J + 1] - t0 for J in range(7)]
fs = [tr[U, 2 * J] - t0 for J in range(7)]
ho = {0: (tr[:U, 2] - t0)}  # approx: hand-off of col 0 right after 'prev last col synced'
for J in range(6):
    ho[J + 1] = tr[:U, 32 + J] - t0
for J in range(7):
    built = (tr[:U, 4 * J + 4] - t0) if J < 6 else None
    print(f"col{J}: hand-off u[min {ho[J].min():6d} max {ho[J].max():6d}]  factor at barrier {fs[J]:6d} end {fe[J]:6d}"
          f"  (L <= {fe[J] - max(fs[J], ho[J].max()):5d})"
          + (f"  u reach sync(2) [min {built.min():6d} max {built.max():6d}]" if built is not None else ""))
print("period:", (tr[:U, 3].max() - t0), "(start skew);  next iteration start = finish col5 ~", (tr[:U, 26] - t0).max())
