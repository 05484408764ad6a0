"""Per-box timeline of the TMA-fed Gram kernel's pipeline (CTA 0): when the stage became free and its
loads were issued, when the raw tile had landed, when it was converted, when its MMAs were issued.
    python profiles/experiments/tma_trace.py C H [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pleas_merging_b200 import _native as N  # noqa: E402
from pleas_merging_b200 import ops  # noqa: E402

C, H = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dev = torch.device("cuda")
x = torch.relu(torch.randn(B, C, H, H, device=dev))
y = torch.relu(torch.randn(B, C, H, H, device=dev))
plan = ops.TmaGramPlan(C, B, H * H, dev)
buf = torch.zeros(1024, dtype=torch.int64, device=dev)
for _ in range(2):
    plan.run(x, y, 1)
torch.cuda.synchronize()
N.lib().plb_debug_set_trace(buf.data_ptr())
plan.run(x, y, 1)
torch.cuda.synchronize()
N.lib().plb_debug_set_trace(None)
full = buf.cpu()
t = full[:1024].view(256, 4)[:254]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); plan.run(x, y, 1); e1.record(); torch.cuda.synchronize()
n = int((t[:, 0] > 0).sum())
t0 = int(t[0, 0])
print(f"C={C} HW={H*H} splits={plan.splits} cg={plan.cta_group}: {n} boxes traced (clocks relative to the first issue)")
k0 = int(full[1016])
print(f"kernel (events) {e0.elapsed_time(e1)*1e3:.1f} us; CTA 0 clocks from entry: setup done {int(full[1017])-k0}, first TMA issue "
      f"{t0-k0}, last chain drained {int(full[1018])-k0}, partial stored {int(full[1019])-k0}, exit {int(full[1020])-k0}; "
      f"globaltimer entry->exit {(int(full[1023])-int(full[1022]))/1e3:.1f} us")
print(" box   issue  landed  convd  mma_start | load  conv  wait_mma | d_issue")
prev = None
for i in list(range(min(n, 6))) + list(range(max(6, n - 6), n)):
    a, b, c, d = (int(v) - t0 for v in t[i])
    print(f"{i:4d} {a:7d} {b:7d} {c:7d} {d:7d} | {b-a:5d} {c-b:5d} {d-c:5d} | {'' if prev is None else a-prev}")
    prev = a
