"""Per-batch wall time of the PLeaS accumulation loop (LstsqRunner) on a ResNet-50 pair: PLB_CONV=0/1."""
import sys, time
import torch
sys.path.insert(0, "/root/repo")
import bench
import pleas_merging_b200 as P
from pleas_merging_b200.methods import pleas_merging as PM
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
m1, m2 = bench.make_models("resnet50", dev)
spec = P.get_permutation_spec(m1, ((1, 3, 224, 224),))
xs = [torch.randn(32, 3, 224, 224, device=dev) for _ in range(4)]
perm, costs = P.activation_matching(spec, m1, m2, [(x, 0) for x in xs[:2]], 2, output_costs=True, accumulate="sum")
blocks = P.get_blocks(spec, perm, costs, 0.0)
pb = dict(blocks)
for axis, pg in spec.items():
    for ax in pg.state:
        pb[ax] = pb[axis]
m3 = P.partial_merge(spec, m1, m2, perm, costs, 0.0)
torch.cuda.synchronize()
t0 = time.perf_counter()
lr = PM.LstsqRunner(m1, m2, m3, pb, 1000, False, "rn50", use_cuda_graph=True)
torch.cuda.synchronize()
print(f"runner construction {time.perf_counter() - t0:.3f} s")
with torch.no_grad():
    for i in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lr.run(xs[i % 4])
        torch.cuda.synchronize()
        print(f"batch {i}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
lr.close()
