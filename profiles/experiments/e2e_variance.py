"""Why does the e2e activation_matching wall time vary between runs?  H2D bandwidth of the pinned batches and
repeated timed calls of the public API on the same process."""
import sys, time
import torch
sys.path.insert(0, "/root/repo")
import bench
import pleas_merging_b200 as P
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
m1, m2 = bench.make_models("resnet50", dev)
spec = P.get_permutation_spec(m1, ((1, 3, 224, 224),))
host = [(torch.randn(32, 3, 224, 224).pin_memory(), 0) for _ in range(16)]
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ys = [h[0].to(dev, non_blocking=True) for h in host]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D 16 x 19.3 MB pinned: {dt*1e3:.1f} ms = {16*19.27/dt/1e3:.1f} GB/s", flush=True)
loader = [host[i % 16] for i in range(20)]
P.activation_matching(spec, m1, m2, loader[:2], 2, accumulate="sum")
for rep in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    P.activation_matching(spec, m1, m2, loader, 20, accumulate="sum")
    torch.cuda.synchronize()
    print(f"activation_matching 20 batches: {time.perf_counter() - t0:.3f} s", flush=True)
