"""Experiment: 3xTF32 tcgen05 GEMM error vs. accumulation-chain length (k-blocks per CTA).
Run on the GPU box: python profiles/experiments/exp_chain_length.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pleas_merging_b200 import ops

def run(name, x, y):
    C, K = x.shape
    X, Y = x.double().numpy(), y.double().numpy()
    ref = X @ Y.T
    kb = (K + 15) // 16
    xd, yd = x.cuda(), y.cuda()
    pa, pb = ops.Planes(C, kb, xd.device), ops.Planes(C, kb, xd.device)
    ops.pack_split(xd, 0, pa); ops.pack_split(yd, 0, pb)
    for chain in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 100000):
        splits = max(1, -(-kb // chain))
        for impl in (0, 1):
            plan = ops.GemmPlan(pa, pb, C, C, kb, splits=splits)
            plan.run(impl)
            out = torch.empty(C, C, device=xd.device); plan.finalize(out)
            o = out.cpu().double().numpy()
            err = np.abs(o - ref)
            big = np.abs(ref) > 0.1 * np.abs(ref).max()
            print(f"{name} impl={'tc' if impl==0 else 'simt'} chain<={chain:6d} splits={splits:5d} "
                  f"maxerr/max={err.max()/np.abs(ref).max():.2e} max-rel(big entries)={(err[big]/np.abs(ref[big])).max():.2e} "
                  f"mean-signed-rel={((o-ref)[big]/ref[big]).mean():+.2e}")

g = torch.Generator().manual_seed(0)
run("randn  96x8192 ", torch.randn(96, 8192, generator=g), torch.randn(96, 8192, generator=g))
run("relu  256x8192 ", torch.relu(torch.randn(256, 8192, generator=g)), torch.relu(torch.randn(256, 8192, generator=g)))
