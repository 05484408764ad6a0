"""Runs a few PLeaS normal-equation steps on the ResNet-50 pair un-captured inside an NVTX range
(for the ncu launch list) and prints the eager / graph-replay step times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
import pleas_merging_b200 as P
import importlib
PM = importlib.import_module("pleas_merging_b200.methods.pleas_merging")
from pleas_merging_b200.methods.partial_matching import get_blocks
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
torch.manual_seed(0); m1 = torchvision.models.resnet50().eval().cuda()
torch.manual_seed(1); m2 = torchvision.models.resnet50().eval().cuda()
spec = P.get_permutation_spec(m1, ((1, 3, 224, 224),))
perm = P.make_identity_perm(spec)
costs = {k: torch.eye(pg.size, device="cuda") for k, pg in spec.items()}
m3 = P.partial_merge(spec, m1, m2, perm, costs, 0.0)
blocks = get_blocks(spec, perm, costs, 0.0)
pb = dict(blocks)
for axis, pg in spec.items():
    for ax in pg.state:
        pb[ax] = pb[axis]
runner = PM.LstsqRunner(m1, m2, m3, pb, 1000, False, "rn50", use_cuda_graph=True)
xs = [torch.randn(32, 3, 224, 224, device="cuda") for _ in range(4)]
with torch.no_grad():
    for i in range(3):
        runner.run(xs[i])          # eager, capture+replay, replay
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(10):
        runner.run(xs[i % 4])
    torch.cuda.synchronize()
    t_graph = (time.perf_counter() - t0) / 10
    torch.cuda.nvtx.range_push("plb_pleas")
    runner._eager(xs[0])
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("pleas step (graph replay): %.2f ms" % (t_graph * 1e3))
