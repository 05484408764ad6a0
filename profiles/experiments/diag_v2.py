import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pleas_merging_b200 import ops
from oracle import ref_oracle as O
g = torch.Generator().manual_seed(3)
for shape in [(3,200,9,10), (3,200,9,10), (2,1300,12,12), (1,200,16,17), (1,200,16,33)]:
    x = torch.relu(torch.randn(*shape, generator=g)); y = torch.relu(torch.randn(*shape, generator=g))
    X, Y = O._rows(x.numpy(),1).astype(np.float64), O._rows(y.numpy(),1).astype(np.float64)
    ref = X@Y.T
    for impl in ("tcgen05", "tcgen05_v1"):
        ops.set_gemm_impl(impl)
        G = ops.cross_statistic(x.cuda(), y.cuda(), 1, ops.MODE_INNER).cpu().numpy()
        err = np.abs(G-ref)
        bad = np.argwhere(err > 1e-5*np.abs(ref).max())
        K = X.shape[1]
        print(shape, impl, "K", K, "kb", (K+15)//16, "maxerr/max %.2e"%(err.max()/np.abs(ref).max()), "nbad", len(bad),
              "rows", (bad[:,0].min(), bad[:,0].max()) if len(bad) else None, "cols", (bad[:,1].min(), bad[:,1].max()) if len(bad) else None)
        if len(bad):
            i,j = bad[0]; print("   e.g.", i, j, G[i,j], ref[i,j], "ratio", G[i,j]/ref[i,j])
