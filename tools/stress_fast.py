import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.cases import by_name, make_data
from muygpys_b200 import ops
from muygpys_b200.gp.tensors import fast_nn_update
case = by_name("c5_m05_2d"); data = make_data(case)
x = torch.as_tensor(data["train_x"]).cuda(); y = torch.as_tensor(data["train_y"][:,0]).cuda(); q = torch.as_tensor(data["test_x"]).cuda()
grid = ops.KnnGrid(x)
ref_nn = None; ref_c = None; ref_f = None
for it in range(300):
    nn, _ = grid.query(x, case.k)
    if ref_nn is None: ref_nn = nn.clone()
    if not torch.equal(nn, ref_nn): print(it, "KNN differs", (nn != ref_nn).sum().item())
    nf = fast_nn_update(nn)
    out = ops.fused_posterior(x, x, None, nf, y, kernel_id=1, metric_id=0, length_scale=0.1, noise=1e-3, want_mean=False, want_var=False, want_coeffs=True, want_status=True)
    c = out["coeffs"]
    if ref_c is None: ref_c = c.clone()
    bad = torch.isnan(c).any(dim=(1,2)).nonzero().flatten()
    if len(bad) or not torch.equal(c, ref_c):
        print(it, "coeffs NaN rows", bad[:10].tolist(), "status sum", int(out["status"].sum()), "differs", int((c != ref_c).any(dim=(1,2)).sum()))
    tn, _ = grid.query(q, case.k); closest = tn[:, 0].contiguous()
    f = ops.fast_mean(x, q, None, nf[closest], closest, c, kernel_id=1, metric_id=0, length_scale=0.1)
    if ref_f is None: ref_f = f.clone()
    if torch.isnan(f).any() or not torch.equal(f, ref_f): print(it, "fast mean NaN/diff", int(torch.isnan(f).sum()))
print("done")
