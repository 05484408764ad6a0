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
    buf = torch.zeros(8192, dtype=torch.int64, device="cuda")
    N.lib().plb_conv_debug_set_trace(buf.data_ptr())
    pair(xa, xb)
    torch.cuda.synchronize()
    N.lib().plb_conv_debug_set_trace(None)
t = buf.view(512, 16).cpu()
n = int((t[:, 6] > 0).sum())
t0 = int(t[0, 0])
print(f"shape {SHAPES[idx]}: {n} boxes traced; clocks relative to the first TMA issue")
print(" box | top acc_ok full(sees) fenced peeked mmas_out commits_out synced | epi_sees epi_done | next_ready next_acc")
rows = list(range(min(n, 12))) + list(range(max(12, n - 24), n))
for i in rows:
    r = [int(x) for x in t[i]]
    nr, na = r[11] & 1, r[13] & 1
    r[11] //= 2
    r[13] //= 2
    v = [x - t0 for x in r]
    print(f"{i:4d} | {v[8]:7d} {v[9]:7d} {v[4]:7d} {v[10]:7d} {v[11]:7d} {v[12]:7d} {v[13]:7d} {v[6]:7d} | {v[5]:7d} {v[7]:7d} | {nr} {na}")
