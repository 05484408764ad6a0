"""LAP on real ResNet-50 activation-matching cost matrices: SciPy-order cold start vs a start from the column
reduction v_j = min_i c_ij (plb_lap_solve_batched_warm): time, assignments, objectives."""
import sys, time
import torch
sys.path.insert(0, "/root/repo")
import bench
import pleas_merging_b200 as P
from pleas_merging_b200 import ops
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda", 0)
m1, m2 = bench.make_models("resnet50", dev)
spec = P.get_permutation_spec(m1, ((1, 3, 224, 224),))
loader = [(torch.randn(32, 3, 224, 224, device=dev), 0) for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4)]
perm, costs = P.activation_matching(spec, m1, m2, loader, len(loader), output_costs=True, accumulate="sum")
mats = list(costs.values())


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, time.perf_counter() - t0


for rep in range(2):
    (o_cold, obj_cold, st), t_cold = timed(lambda: ops.lap_solve_batched(mats, True))
    vin = [(-m.max(dim=0).values).double().contiguous() for m in mats]
    (o_warm, obj_warm, st2), t_warm = timed(lambda: ops.lap_solve_batched(mats, True, v_init=vin))
    diff = sum(int((a != b).sum()) for a, b in zip(o_cold, o_warm))
    print(f"cold {t_cold*1e3:.1f} ms   column-reduced start {t_warm*1e3:.1f} ms   differing assignments {diff} of "
          f"{sum(m.shape[0] for m in mats)}   max |objective diff| {float((obj_cold - obj_warm).abs().max()):.3e}", flush=True)
big = [m for m in mats if m.shape[0] >= 1024]
for m in big[:3] + big[-1:]:
    (_, _, _), tc = timed(lambda: ops.lap_solve_batched([m], True))
    v = [(-m.max(dim=0).values).double().contiguous()]
    (_, _, _), tw = timed(lambda: ops.lap_solve_batched([m], True, v_init=v))
    print(f"n={m.shape[0]}: cold {tc*1e3:.1f} ms, column-reduced {tw*1e3:.1f} ms")
