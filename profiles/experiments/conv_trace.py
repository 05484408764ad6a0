"""Per-box clock64 stamps of CTA 0 of the convolution kernel (plb_conv_debug_set_trace)."""
import sys
import torch
sys.path.insert(0, "/root/repo")
from pleas_merging_b200 import conv, _native as N
SHAPES = [(256, 56, 128, 1, 1, 0), (64, 56, 256, 1, 1, 0), (256, 14, 256, 3, 1, 1), (64, 56, 64, 3, 1, 1), (256, 14, 1024, 1, 1, 0)]
idx = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cin, h, cout, k, s, p = SHAPES[idx]
with torch.no_grad():
    ma = torch.nn.Conv2d(cin, cout, k, s, p, bias=False).cuda()
    mb = torch.nn.Conv2d(cin, cout, k, s, p, bias=False).cuda()
    xa, xb = torch.randn(32, cin, h, h, device="cuda"), torch.randn(32, cin, h, h, device="cuda")
    pair = conv.ConvPair(ma, mb)
    pair(xa, xb)
    torch.cuda.synchronize()
    buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
    N.lib().plb_conv_debug_set_trace(buf.data_ptr())
    pair(xa, xb)
    torch.cuda.synchronize()
    N.lib().plb_conv_debug_set_trace(None)
t = buf.view(512, 8).cpu()
n = int((t[:, 6] > 0).sum())
t0 = int(t[0, 0])
print(f"shape {SHAPES[idx]}: {n} boxes traced; clocks relative to the first TMA issue")
print(" box | tma_issue | mma_top mma_acc_ok loader_stored | mma_sees issued | epi_sees epi_done (chain = box when PLB_CONV_CHAIN=1) | d_issued")
prev = None
rows = list(range(min(n, 20))) + list(range(max(20, n - 12), n))
for i in rows:
    flags = (0, 0)
    v = [int(x) - t0 for x in t[i]]
    print(f"{i:4d} | {v[0]:8d} | {v[1]:8d} {v[2]:8d} {v[3]:8d} | {v[4]:8d} {v[6]:8d} | {v[5]:8d} {v[7]:8d} | {'' if prev is None else v[6]-prev} acc_ok={flags[0]} ready={flags[1]}")
    prev = v[6]
