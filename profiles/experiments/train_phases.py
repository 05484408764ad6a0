"""Times the phases of pleas_merging.train's closed form (first eager batch, capture, replays, solve)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torchvision, importlib
import pleas_merging_b200 as P
PM = importlib.import_module("pleas_merging_b200.methods.pleas_merging")
from pleas_merging_b200.parallel import BatchSharder, device_prefetch
torch.backends.cudnn.allow_tf32 = False; torch.backends.cudnn.benchmark = True
torch.manual_seed(0); m1 = torchvision.models.resnet50().eval().cuda()
torch.manual_seed(1); m2 = torchvision.models.resnet50().eval().cuda()
spec = P.get_permutation_spec(m1, ((1, 3, 224, 224),))
g = torch.Generator().manual_seed(1)
host = [(torch.randn(32, 3, 224, 224, generator=g).pin_memory(), 0) for _ in range(8)]
loader = [host[i % 8] for i in range(41)]
perm, costs = P.activation_matching(spec, m1, m2, loader[:3], 3, output_costs=True, accumulate="sum")
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    model3 = P.partial_merge(spec, m1, m2, perm, costs, 0.0)
    blocks = P.get_blocks(spec, perm, costs, 0.0)
    pb = dict(blocks)
    for axis, pg in spec.items():
        for ax in pg.state: pb[ax] = pb[axis]
    t0 = sync()
    runner = PM.LstsqRunner(m1, m2, model3, pb, 1000, False, "rn50", True)
    t1 = sync(); marks = []
    with torch.no_grad():
        for i, x in device_prefetch(BatchSharder(loader, 41, 0, 1), runner.device):
            runner.run(x)
            if i < 4 or i == 40: marks.append((i, sync()))
    t2 = sync()
    for name, acc in runner.accs.items():
        acc.solve(runner.layers3[name], 1e-4)
    t3 = sync()
    runner.close(); t4 = sync()
    print(f"rep{rep}: build {t1-t0:.3f}s batches " + " ".join(f"[{i}]@{t-t1:.3f}" for i, t in marks) + f" loop {t2-t1:.3f}s solve {t3-t2:.3f}s close {t4-t3:.3f}s")
    t0 = sync(); stats = {}
    P.train(loader, m1, m2, P.partial_merge(spec, m1, m2, perm, costs, 0.0), spec, perm, costs, 0.0, False, 40, None, stats=stats)
    print(f"  train() total {sync()-t0:.3f}s timing {stats['_timing']}")
