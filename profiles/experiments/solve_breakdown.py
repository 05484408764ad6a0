"""Where the PLeaS solve phase spends its time (ResNet-50 pair, full merge): per-layer wall time of
_LayerLS.solve split into host mask work, right-hand-side preparation, Cholesky and the rest."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torchvision
import pleas_merging_b200 as P
import importlib
PM = importlib.import_module("pleas_merging_b200.methods.pleas_merging")
from pleas_merging_b200 import ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
torch.manual_seed(0); m1 = torchvision.models.resnet50().eval().cuda()
torch.manual_seed(1); m2 = torchvision.models.resnet50().eval().cuda()
spec = P.get_permutation_spec(m1, ((1, 3, 224, 224),))
perm = P.make_identity_perm(spec)
costs = {k: torch.eye(pg.size, device="cuda") for k, pg in spec.items()}
m3 = P.partial_merge(spec, m1, m2, perm, costs, 0.0)
loader = [(torch.randn(32, 3, 224, 224), 0) for _ in range(6)]

chol_t = [0.0]
orig_chol = ops.chol_solve_
def timed_chol(G, B, r):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = orig_chol(G, B, r)
    torch.cuda.synchronize(); chol_t[0] += time.perf_counter() - t0
    return out
ops.chol_solve_ = timed_chol
per_layer = []
orig_solve = PM._LayerLS.solve
def timed_solve(self, layer, ridge):
    torch.cuda.synchronize(); t0 = time.perf_counter(); c0 = chol_t[0]
    out = orig_solve(self, layer, ridge)
    torch.cuda.synchronize()
    per_layer.append((self.name, self.K, self.cout, time.perf_counter() - t0, chol_t[0] - c0))
    return out
PM._LayerLS.solve = timed_solve
stats = {}
pr = cProfile.Profile()
for rep in range(2):
    per_layer.clear(); chol_t[0] = 0.0
    if rep == 1: pr.enable()
    P.train(loader, m1, m2, m3, spec, perm, costs, 0.0, False, 5, None, stats=stats)
    if rep == 1: pr.disable()
print("timing", stats["_timing"])
tot = sum(t for *_, t, _ in per_layer); totc = sum(c for *_, c in per_layer)
print("solve total %.3f s, of which chol %.3f s" % (tot, totc))
for name, K, co, t, c in sorted(per_layer, key=lambda r: -r[3])[:12]:
    print("  %-28s K=%5d Co=%5d  %.1f ms (chol %.1f ms)" % (name, K, co, t * 1e3, c * 1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
