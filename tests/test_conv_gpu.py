"""The 3xTF32 tcgen05 convolution (csrc/conv.cu, plb_conv2d_forward) against the fp64 convolution of the same
inputs: every geometry class of the ResNet family the configs use (1x1, 3x3 pad 1, strided, the 3-channel 7x7
stem, 7x7 / 14x14 maps whose rows are not 16-byte multiples, ragged position / channel tiles, bias), through
the C ABI.  Tolerance: 3e-6 of the output's largest magnitude (fp32 cuDNN itself sits at ~1e-6)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    # (NB, Cin, H, W, Cout, k, stride, pad, bias)
    (2, 64, 56, 56, 64, 1, 1, 0, False),
    (2, 64, 56, 56, 256, 1, 1, 0, False),
    (3, 64, 28, 28, 64, 3, 1, 1, False),
    (2, 128, 56, 56, 128, 3, 2, 1, False),
    (2, 256, 56, 56, 512, 1, 2, 0, False),
    (5, 256, 14, 14, 256, 3, 1, 1, True),
    (32, 512, 7, 7, 512, 3, 1, 1, False),
    (32, 2048, 7, 7, 512, 1, 1, 0, False),
    (4, 3, 224, 224, 64, 7, 2, 3, False),
    (2, 96, 17, 13, 72, 3, 1, 1, True),      # ragged everything, Cout not a tile multiple
    (1, 32, 5, 5, 8, 5, 1, 2, True),
    (2, 40, 9, 9, 24, 3, 1, 0, False),       # Cin % 32 != 0 -> flat form
    (2, 64, 12, 12, 130, 1, 1, 0, True),
]


def _mods(cin, cout, k, stride, pad, bias, seed):
    torch.manual_seed(seed)
    m = torch.nn.Conv2d(cin, cout, k, stride, pad, bias=bias).cuda()
    return m


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_conv_pair_matches_fp64(case):
    from pleas_merging_b200 import conv

    nb, cin, h, w, cout, k, stride, pad, bias = case
    ma, mb = _mods(cin, cout, k, stride, pad, bias, 1), _mods(cin, cout, k, stride, pad, bias, 2)
    g = torch.Generator(device="cuda").manual_seed(3)
    xa = torch.randn(nb, cin, h, w, device="cuda", generator=g)
    xb = torch.relu(torch.randn(nb, cin, h, w, device="cuda", generator=g)) * 3.0
    assert conv.eligible(ma)
    pair = conv.ConvPair(ma, mb)
    with torch.no_grad():
        ya, yb = pair(xa, xb)
        for y, m, x in ((ya, ma, xa), (yb, mb, xb)):
            ref = torch.nn.functional.conv2d(x.double(), m.weight.double(), None if m.bias is None else m.bias.double(),
                                             stride, pad)
            assert y.shape == ref.shape
            err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
            assert err <= 3e-6, err
        # single-problem launch
        y1 = conv.conv2d(xa, ma)
        assert torch.equal(y1, ya)


def test_conv_weight_update_is_seen():
    from pleas_merging_b200 import conv

    m = _mods(64, 64, 3, 1, 1, False, 0)
    x = torch.randn(2, 64, 8, 8, device="cuda")
    pk = conv.PackedConv(m)
    with torch.no_grad():
        y0 = conv.conv2d(x, m, pk)
        m.weight.mul_(2.0)
        y1 = conv.conv2d(x, m, pk)
    assert torch.allclose(y1, 2 * y0, rtol=1e-5, atol=1e-6)


def test_conv_ineligible_falls_back_to_module():
    from pleas_merging_b200 import conv

    m = torch.nn.Conv2d(64, 64, 3, 1, 1, groups=2).cuda()
    assert not conv.eligible(m)
    x = torch.randn(1, 64, 8, 8, device="cuda")
    with torch.no_grad():
        assert torch.equal(conv.conv2d(x, m), m(x))


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("case", [(2, 64, 28, 28, 256, 1, 1, 0), (3, 64, 14, 14, 72, 3, 1, 1), (2, 3, 32, 32, 24, 7, 2, 3)],
                         ids=["1x1", "3x3-ragged", "stem"])
def test_conv_with_fused_batchnorm_relu_output(case, relu):
    """plb_conv2d_affine_forward: the eval-mode BatchNorm (+ ReLU) behind a convolution as a second output of the same
    launch, against the fp64 modules; the first output stays the plain convolution."""
    from pleas_merging_b200 import conv

    nb, cin, h, w, cout, k, stride, pad = case
    mods, bns = [], []
    for seed in (1, 2):
        torch.manual_seed(seed)
        mods.append(torch.nn.Conv2d(cin, cout, k, stride, pad, bias=(seed == 2 and k == 3)).cuda())
        bn = torch.nn.BatchNorm2d(cout).cuda().eval()
        with torch.no_grad():
            bn.weight.copy_(torch.randn(cout).cuda())        # negative scales included
            bn.bias.copy_(0.3 * torch.randn(cout).cuda())
            bn.running_mean.copy_(0.2 * torch.randn(cout).cuda())
            bn.running_var.copy_(0.5 + torch.rand(cout).cuda())
        bns.append(bn)
    same_bias = (mods[0].bias is None) == (mods[1].bias is None)
    if not same_bias:  # a pair needs the same geometry
        mods[0] = torch.nn.Conv2d(cin, cout, k, stride, pad, bias=True).cuda()
    xa, xb = torch.randn(nb, cin, h, w, device="cuda"), torch.randn(nb, cin, h, w, device="cuda")
    pair = conv.ConvPair(mods[0], mods[1], bns=tuple(bns), relu=relu)
    with torch.no_grad():
        ya, yb, za, zb = pair(xa, xb)
        for y, z, m, bn, x in ((ya, za, mods[0], bns[0], xa), (yb, zb, mods[1], bns[1], xb)):
            ref = torch.nn.functional.conv2d(x.double(), m.weight.double(), None if m.bias is None else m.bias.double(),
                                             stride, pad)
            ref2 = torch.nn.functional.batch_norm(ref, bn.running_mean.double(), bn.running_var.double(),
                                                  bn.weight.double(), bn.bias.double(), False, 0.0, bn.eps)
            if relu:
                ref2 = torch.relu(ref2)
            assert (y.double() - ref).abs().max().item() <= 3e-6 * ref.abs().max().item()
            assert (z.double() - ref2).abs().max().item() <= 4e-6 * max(ref2.abs().max().item(), 1.0)
        # a BatchNorm update is seen by the next call
        bns[0].running_mean.add_(1.0)
        _, _, za2, _ = pair(xa, xb)
        assert not torch.equal(za2, za)
