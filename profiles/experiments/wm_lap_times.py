"""Per-sweep time of the batched LAP launches inside weight_matching (ResNet-50 pair, seed 0), warm vs cold."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torchvision
import pleas_merging_b200 as P
from pleas_merging_b200 import ops
torch.manual_seed(0); m1 = torchvision.models.resnet50().eval().cuda()
torch.manual_seed(1); m2 = torchvision.models.resnet50().eval().cuda()
spec = P.get_permutation_spec(m1, ((1, 3, 224, 224),))
orig = ops.lap_solve_batched
log = []
def timed(costs, maximize=True, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = orig(costs, maximize, **kw); e1.record()
    log.append((e0, e1, max(c.shape[0] for c in costs), len(costs), kw.get("v_init") is not None and any(v is not None for v in kw["v_init"])))
    return out
ops.lap_solve_batched = timed
for rep in range(2):
    log.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    P.weight_matching(spec, m1.state_dict(), m2.state_dict(), verbose=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ms = [(a.elapsed_time(b), n, k, w) for a, b, n, k, w in log]
    print(f"rep{rep}: total {dt:.3f}s, {len(ms)} LAP launches, LAP time {sum(m[0] for m in ms)/1e3:.3f}s")
    big = [m for m in ms if m[1] == 2048]
    print("  launches containing the n=2048 group (ms):", " ".join(f"{m[0]:.1f}{'w' if m[3] else 'c'}" for m in big))
