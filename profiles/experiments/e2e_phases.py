"""Times the phases of the public activation_matching call (trace, first eager batch, capture, replays, LAP)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torchvision, importlib
import pleas_merging_b200 as P
from pleas_merging_b200 import ops
from pleas_merging_b200.parallel import BatchSharder, device_prefetch
from pleas_merging_b200.core.solvers import solve_lsa_batched
AM = importlib.import_module("pleas_merging_b200.methods.activation_matching")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cudnn.benchmark = True
torch.manual_seed(0); m1 = torchvision.models.resnet50().eval().cuda()
torch.manual_seed(1); m2 = torchvision.models.resnet50().eval().cuda()
spec = P.get_permutation_spec(m1, ((1, 3, 224, 224),))
g = torch.Generator().manual_seed(1)
host = [(torch.randn(32, 3, 224, 224, generator=g).pin_memory(), 0) for _ in range(8)]
loader = [host[i % 8] for i in range(30)]
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = sync()
    runner = AM.CalibrationRunner(spec, m1, m2, ops.MODE_NEG_CDIST, "sum", True)
    t1 = sync()
    marks = []
    with torch.inference_mode():
        for i, x in device_prefetch(BatchSharder(loader, 30, 0, 1), runner.device):
            runner.run(x)
            if i < 4 or i == 29: marks.append((i, sync()))
    t2 = sync()
    costs = {k: c.clone() for k, c in zip(runner.acc.keys, runner.acc.costs)}
    perms = solve_lsa_batched(costs.values()); t3 = sync()
    runner.close(); del runner; import gc; gc.collect(); t4 = sync()
    print(f"rep{rep}: build {t1-t0:.3f}s  batches " + " ".join(f"[{i}]@{t-t1:.3f}" for i, t in marks) +
          f"  loop {t2-t1:.3f}s  lap {t3-t2:.3f}s  close {t4-t3:.3f}s  mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB reserved {torch.cuda.memory_reserved()/2**30:.1f} GiB")
