"""Times a few convolution classes of the library kernel (PLB_CONV_DEBUG experiments read the env in-process)."""
import sys
import torch
sys.path.insert(0, "/root/repo")
from pleas_merging_b200 import conv
SHAPES = [(256, 56, 128, 1, 1, 0), (64, 56, 256, 1, 1, 0), (256, 14, 256, 3, 1, 1), (128, 56, 128, 3, 2, 1), (3, 224, 64, 7, 2, 3)]
only = int(sys.argv[1]) if len(sys.argv) > 1 else -1
with torch.no_grad():
    for idx, (cin, h, cout, k, s, p) in enumerate(SHAPES):
        if only >= 0 and idx != only:
            continue
        ma = torch.nn.Conv2d(cin, cout, k, s, p, bias=False).cuda()
        mb = torch.nn.Conv2d(cin, cout, k, s, p, bias=False).cuda()
        nbuf = 1 if only >= 0 else max(2, int(300e6 // (32 * cin * h * h * 8)) + 1)
        xs = [(torch.randn(32, cin, h, h, device="cuda"), torch.randn(32, cin, h, h, device="cuda")) for _ in range(nbuf)]
        pair = conv.ConvPair(ma, mb)
        reps = 1 if only >= 0 else 20
        for r in range(3 if only < 0 else 1):
            pair(*xs[r % nbuf])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            pair(*xs[r % nbuf])
        e1.record()
        torch.cuda.synchronize()
        print(f"cin={cin} hw={h} cout={cout} k={k} s={s}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us", flush=True)
