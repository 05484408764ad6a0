"""Forward convolutions of the source models on the library's 3xTF32 tcgen05 implicit-GEMM kernel
(csrc/conv.cu, plb_conv2d_forward) instead of ATen's cuDNN fp32 kernels.

The reference's loops (pleas/methods/activation_matching.py:123, pleas_merging.py:262-281) spend most of
their time in the two source models' convolutions; the weights of those models are constants of the merge,
so they are split into tf32 hi / lo planes once (``PackedConv``) and every batch runs the same layer of
BOTH models in one launch (``conv2d_pair``).

``eligible(module)`` names what the kernel covers: plain ``nn.Conv2d`` (groups 1, dilation 1, zero padding,
equal strides, fp32).  Anything else keeps the module's own forward.  ``PLB_CONV=0`` switches the path off.
"""
import ctypes
import os

import torch

from . import _native as N

ENABLED = os.environ.get("PLB_CONV", "1") == "1"
MAX_FLAT_K = 512
CONV_TIMER = None  # bench: list of (start event, end event, algorithmic flops, bytes)


def eligible(mod) -> bool:
    if not ENABLED or type(mod) is not torch.nn.Conv2d:
        return False
    if mod.groups != 1 or tuple(mod.dilation) != (1, 1) or mod.padding_mode != "zeros":
        return False
    if isinstance(mod.padding, str) or mod.stride[0] != mod.stride[1]:
        return False
    w = mod.weight
    if w.dtype != torch.float32 or not w.is_cuda:
        return False
    cin, kh, kw = w.shape[1], w.shape[2], w.shape[3]
    kwp = 1 << max(kw - 1, 0).bit_length()  # flat form: kernel rows padded to a power of two of taps
    if cin % 32 != 0 and -(-cin * kh * kwp // 32) * 32 > MAX_FLAT_K:
        return False
    return kh < 256 and kw < 256 and (cin % 32 == 0 or kw <= 16)


class PackedConv:
    """tf32 hi / lo planes of one Conv2d weight in the kernel's order; repacked when the weight tensor
    changes (address or in-place version)."""

    def __init__(self, mod):
        self.mod = mod
        self.packed = None
        self.key = None
        self.refresh()

    def refresh(self):
        w = self.mod.weight
        try:
            version = w._version
        except RuntimeError:  # inference tensors carry no version counter
            version = -1
        key = (w.data_ptr(), version, tuple(w.shape))
        if key == self.key:
            return
        cout, cin, kh, kw = w.shape
        n = N.lib().plb_conv_packed_floats(cout, cin, kh, kw)
        if self.packed is None or self.packed.numel() != n:
            self.packed = torch.empty(n, dtype=torch.float32, device=w.device)
        wc = w.detach().contiguous()
        N.call("plb_conv_pack_weights", w.device, wc.data_ptr(), cout, cin, kh, kw, self.packed.data_ptr())
        self.key = key


def _ptr_array(ts):
    return (ctypes.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])


def bn_foldable(bn) -> bool:
    """Eval-mode BatchNorm2d with running statistics: a per-channel affine map the convolution's epilogue can apply."""
    return type(bn) is torch.nn.BatchNorm2d and not bn.training and bn.track_running_stats and \
        bn.running_mean is not None and bn.running_var is not None and bn.running_mean.dtype == torch.float32 and \
        bn.running_mean.is_cuda


class FoldedBN:
    """Interleaved fp32 (scale, shift) of an eval-mode BatchNorm: bn(x) = scale x + shift per channel (formed in
    float64); refreshed when one of the module's tensors changes."""

    def __init__(self, bn):
        self.bn, self.vec, self.key = bn, None, None
        self.refresh()

    def refresh(self):
        bn = self.bn
        ts = (bn.running_mean, bn.running_var, bn.weight, bn.bias)
        key = tuple(None if t is None else (t.data_ptr(), t._version) for t in ts)
        if key == self.key:
            return
        s = (bn.running_var.detach().double() + bn.eps).rsqrt()
        if bn.weight is not None:
            s = s * bn.weight.detach().double()
        t = -bn.running_mean.detach().double() * s
        if bn.bias is not None:
            t = t + bn.bias.detach().double()
        vec = torch.stack([s, t], dim=1).float().contiguous()
        if self.vec is None:
            self.vec = vec
        else:
            self.vec.copy_(vec)
        self.key = key


def conv2d_forward(xs, packs, folded=None, relu=False):
    """Runs ``packs[i].mod`` on ``xs[i]`` (1 or 2 problems of identical geometry) in one launch; returns
    the list of outputs.  Inputs must be fp32 CUDA NCHW-contiguous (the caller checks / falls back).
    ``folded`` (one FoldedBN per problem): the launch also writes relu?(bn(conv(x))) and the result is
    ``(outs, outs2)``."""
    mod = packs[0].mod
    w = mod.weight
    cout, cin, kh, kw = w.shape
    nb, _, ih, iw = xs[0].shape
    stride, (ph, pw) = mod.stride[0], mod.padding
    oh, ow = (ih + 2 * ph - kh) // stride + 1, (iw + 2 * pw - kw) // stride + 1
    outs = [torch.empty((nb, cout, oh, ow), dtype=torch.float32, device=x.device) for x in xs]
    outs2 = [torch.empty_like(o) for o in outs] if folded is not None else None
    biases = [pk.mod.bias.detach() if pk.mod.bias is not None else None for pk in packs]
    if CONV_TIMER is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    N.call("plb_conv2d_affine_forward", xs[0].device, _ptr_array(xs), _ptr_array([pk.packed for pk in packs]),
           _ptr_array(biases), _ptr_array(outs), None if folded is None else _ptr_array([f.vec for f in folded]),
           None if folded is None else _ptr_array(outs2), int(bool(relu)), len(xs), nb, cin, ih, iw, cout, kh, kw,
           stride, ph, pw)
    if CONV_TIMER is not None:
        e1.record()
        flops = 2.0 * len(xs) * nb * oh * ow * cout * cin * kh * kw
        nbytes = 4.0 * len(xs) * (nb * cin * ih * iw + (1 if folded is None else 2) * nb * cout * oh * ow)
        CONV_TIMER.append((e0, e1, flops, nbytes, (cin, cout, kh, stride, ih)))
    return outs if folded is None else (outs, outs2)


def input_ok(x, mod):
    return x.dim() == 4 and x.dtype == torch.float32 and x.is_cuda and x.is_contiguous() and \
        x.shape[1] == mod.weight.shape[1] and x.numel() > 0 and x.shape[1] * x.shape[2] * x.shape[3] < 2 ** 31 and \
        not (torch.is_grad_enabled() and (x.requires_grad or mod.weight.requires_grad))


class ConvPair:
    """The same convolution layer of the two source models: fx ``call_function`` target state.  With ``bns``
    (the eval-mode BatchNorm modules behind the convolutions) the call returns four tensors: the convolution
    outputs and relu?(bn(.)) of them, written by the same launch."""

    def __init__(self, mod_a, mod_b, bns=None, relu=False):
        self.mods = (mod_a, mod_b)
        self.bns, self.relu = bns, relu
        self.packs = None
        self.folded = None

    def same_geometry(self):
        a, b = self.mods
        return a.weight.shape == b.weight.shape and a.stride == b.stride and a.padding == b.padding and \
            (a.bias is None) == (b.bias is None)

    def refresh(self):
        if self.packs is not None:
            self.packs[0].refresh()
            self.packs[1].refresh()
        if self.folded is not None:
            self.folded[0].refresh()
            self.folded[1].refresh()

    def _fallback(self, xa, xb):
        a, b = self.mods
        ya, yb = a(xa), b(xb)
        if self.bns is None:
            return ya, yb
        za, zb = self.bns[0](ya), self.bns[1](yb)
        if self.relu:
            za, zb = torch.relu(za), torch.relu(zb)
        return ya, yb, za, zb

    def __call__(self, xa, xb):
        a, b = self.mods
        if not (input_ok(xa, a) and input_ok(xb, b) and xa.shape == xb.shape):
            return self._fallback(xa, xb)
        if self.packs is None:
            self.packs = (PackedConv(a), PackedConv(b))
            if self.bns is not None:
                self.folded = (FoldedBN(self.bns[0]), FoldedBN(self.bns[1]))
        else:
            self.refresh()
        if self.bns is None:
            ya, yb = conv2d_forward([xa, xb], self.packs)
            return ya, yb
        (ya, yb), (za, zb) = conv2d_forward([xa, xb], self.packs, self.folded, self.relu)
        return ya, yb, za, zb


def conv2d(x, mod, pack=None):
    """Single convolution through the library kernel (falls back to the module when not covered)."""
    if not (eligible(mod) and input_ok(x, mod)):
        return mod(x)
    pack = pack or PackedConv(mod)
    pack.refresh()
    return conv2d_forward([x], [pack])[0]


class patched_convs:
    """Context manager / handle: while active, every eligible ``nn.Conv2d`` of ``models`` runs through the library
    kernel (module hooks keep firing: only ``forward`` is replaced, per instance).  Used by the PLeaS loops, which
    call the source models as plain modules (pleas_merging.py:262-281 in the reference)."""

    def __init__(self, *models):
        self.models, self.patched = models, []

    def __enter__(self):
        for model in self.models:
            for mod in model.modules():
                if eligible(mod) and "forward" not in mod.__dict__:
                    pack = PackedConv(mod)

                    def forward(x, mod=mod, pack=pack):
                        if not input_ok(x, mod):
                            return torch.nn.Conv2d.forward(mod, x)
                        pack.refresh()
                        return conv2d_forward([x], [pack])[0]

                    mod.forward = forward
                    self.patched.append(mod)
        return self

    def __exit__(self, *exc):
        for mod in self.patched:
            mod.__dict__.pop("forward", None)
        self.patched = []
        return False

    open, close = __enter__, __exit__
