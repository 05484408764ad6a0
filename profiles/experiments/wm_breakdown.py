"""Where does ResNet-50 weight matching spend its time?  Host wall vs GPU time per phase."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torchvision, importlib
import pleas_merging_b200 as P
from pleas_merging_b200 import ops
WM = importlib.import_module("pleas_merging_b200.methods.weight_matching")
torch.manual_seed(0); m1 = torchvision.models.resnet50().eval().cuda()
torch.manual_seed(1); m2 = torchvision.models.resnet50().eval().cuda()
spec = P.get_permutation_spec(m1, ((1, 3, 64, 64),))
sd1, sd2 = m1.state_dict(), m2.state_dict()
P.weight_matching(spec, sd1, sd2, max_iter=1, verbose=False)
acc = collections.defaultdict(lambda: [0.0, 0.0, 0])
def wrap(mod, name, label):
    fn = getattr(mod, name)
    def inner(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(*a, **k); e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        acc[label][0] += t1 - t0; acc[label][1] += e0.elapsed_time(e1) / 1e3; acc[label][2] += 1
        return out
    setattr(mod, name, inner)
torch.cuda.synchronize(); t0 = time.perf_counter()
P.weight_matching(spec, sd1, sd2, max_iter=100, seed=0, verbose=False)
torch.cuda.synchronize(); base = time.perf_counter() - t0
wrap(WM._GroupPlan, "build_cost", "cost build (pack+gemm+finalize)")
wrap(ops, "lap_solve_batched", "lap")
wrap(ops, "wm_progress", "progress")
wrap(ops, "compose_perm", "compose")
wrap(WM, "apply_perm", "apply_perm (gathers)")
torch.cuda.synchronize(); t0 = time.perf_counter()
P.weight_matching(spec, sd1, sd2, max_iter=100, seed=0, verbose=False)
torch.cuda.synchronize(); total = time.perf_counter() - t0
print(f"uninstrumented {base:.3f}s; instrumented (synchronising) {total:.3f}s")
for k, (host, gpu, n) in acc.items():
    print(f"  {k:34s} calls {n:5d}  host-issue {host:.3f}s  gpu {gpu:.3f}s")
