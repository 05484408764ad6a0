"""PLeaS per-layer least squares on B200 (drop-in for pleas/methods/pleas_merging.py:11-405).

The reference fits every Conv2d/Linear of the merged model to the (permuted, averaged)
activations of the two source models — ``min sum_l mean((layer_l(X-bar_l) - Y-bar_l)^2)`` —
with 401 Adam steps (:357-375).  The layers are independent (inputs and targets come from the
frozen source models), so the optimum is a linear least-squares problem per layer.  The
default ``solver="lstsq"`` therefore streams the calibration batches ONCE and

  * builds the normal equations ``G_l = U^T U``, ``R_l = U^T Y-bar`` (U = im2col of X-bar) with
    the fused gather-average-im2col pack kernel and the 3xTF32 tcgen05 GEMM, accumulated in
    fp64 on device;
  * solves ``(G_l + ridge I) dW^T = R_l - G_l W0^T`` per layer with the library's blocked fp64
    Cholesky (csrc/chol.cu), i.e. the minimum-ridge UPDATE of the partial_merge init ``W0``:
    directions the data never excites keep their init value, exactly like a gradient method
    started at ``W0``; entries whose gradient mask is zero stay at ``W0``.

The alternative targets of the reference (``merging="reg_mean" | "perm_separatels" |
"perm_mixedls"``, :125-144) stack two sample groups along the batch axis; here each group is one
more pack + GEMM pass with its own channel maps into the same accumulators.

``solver="adam"`` replays the reference's Adam trajectory (same hooks, targets, masks,
optimizer and schedule) for weight-level parity checks.  Gradient masks reproduce the
reference's index order ``mask[si1, so2]``.
"""
import os
import time
from copy import copy, deepcopy

import torch

from .. import conv, ops
from ..core.utils import Axis, get_attr
from ..graphs import GraphedStep
from ..parallel import (BatchSharder, allreduce_sum_, assign_owners, device_prefetch, reduce_to_owners_,
                        world)
import importlib

_AM = importlib.import_module(".activation_matching", __package__)  # the package re-exports a function of this name
from .partial_matching import get_blocks


# ------------------------------------------------------------------ reference helpers

def get_gradient_mask(perm_blocks, model_weights):
    """0/1 masks per trained parameter (reference :11-60, including its index order)."""
    masks = []
    for layer_name, params in model_weights.items():
        bi = perm_blocks.get(Axis(f"{layer_name}.weight", 1), (torch.arange(3), torch.arange(3), [], []))
        bo = perm_blocks.get(Axis(f"{layer_name}.weight", 0), (torch.arange(1000), torch.arange(1000), [], []))
        ni, mi, no, mo = len(bi[0]), len(bi[2]), len(bo[0]), len(bo[2])
        si1, si2 = slice(ni, ni + mi), slice(ni + mi, ni + 2 * mi)
        so1, so2 = slice(no, no + mo), slice(no + mo, no + 2 * mo)
        for v in params.parameters():
            mask = torch.ones_like(v)
            if v.dim() >= 2:
                mask[si1, so2] = 0.0
                mask[si2, so1] = 0.0
            masks.append(mask)
    return masks


def get_model_dict_and_params(model3):
    """Detached fp32 CUDA copies of every Conv2d/Linear of model3 (reference :152-177)."""
    model3_dict = {}
    for name, v in model3.named_modules():
        if isinstance(v, (torch.nn.Conv2d, torch.nn.Linear)):
            model3_dict[name] = deepcopy(v).float().cuda()
            for p in model3_dict[name].parameters():
                p.requires_grad = True
    m3_params = [p for m in model3_dict.values() for p in m.parameters()]
    return model3, model3_dict, m3_params


def capture_inputs(model, act_dict):
    """Forward hooks storing (input, output) of every Conv2d/Linear (reference :216-231 stores
    the input and recomputes the output with a second forward of the layer, :113-114)."""
    handles = []
    for name, module in model.named_modules():
        if isinstance(module, (torch.nn.Conv2d, torch.nn.Linear)):
            def hook(mod, inp, out, name=name):
                act_dict[name] = (inp[0] if isinstance(inp, tuple) else inp, out)
            handles.append(module.register_forward_hook(hook))
    return handles


def _layer_blocks(perm_blocks, name, cin, num_classes, separate_classifier, model_type):
    """Block lookup with the reference's fallbacks (:91-111)."""
    e = torch.zeros(0, dtype=torch.int64)
    bo = perm_blocks.get(Axis(f"{name}.weight", 0))
    if bo is None:
        if separate_classifier:
            feat = {"rn50": 2048, "rn101": 2048, "rn20": 1024, "rn18": 512}.get(model_type)
            if feat is None:
                raise ValueError(f"Unknown model type: {model_type}")
            bo = (torch.arange(feat), torch.arange(feat), e, e)
        else:
            bo = (torch.arange(num_classes), torch.arange(num_classes), e, e)
    bi = perm_blocks.get(Axis(f"{name}.weight", 1))
    if bi is None:
        bi = (torch.arange(cin), torch.arange(cin), e, e)
    return bi, bo


def get_model_orig_activations(acts1, acts2, bi, bo, merging="perm_gradmask"):
    """(X-bar, Y-bar) of one layer with torch ops (reference :116-147) — used by the Adam
    replay; the least-squares path fuses this into its pack kernel."""
    (ip1, op1), (ip2, op2) = acts1, acts2
    dev = ip1.device
    sel = lambda t, idx: t.index_select(1, idx.to(dev).int())
    i11, i22, i1c, i2c = sel(ip1, bi[0]), sel(ip2, bi[1]), sel(ip1, bi[2]), sel(ip2, bi[3])
    o11, o22, o1c, o2c = sel(op1, bo[0]), sel(op2, bo[1]), sel(op1, bo[2]), sel(op2, bo[3])
    if merging == "reg_mean":
        return torch.cat([ip1, ip2], 0), torch.cat([op1, op2], 0)
    if "perm_separatels" in merging:
        X = torch.cat([torch.cat([i11, i1c, torch.zeros_like(i2c)], 1),
                       torch.cat([i22, torch.zeros_like(i1c), i2c], 1)], 0)
        Y = torch.cat([torch.cat([o11, o1c, torch.zeros_like(o2c)], 1),
                       torch.cat([o22, torch.zeros_like(o1c), o2c], 1)], 0)
        return X, Y
    if "perm_mixedls" in merging:
        X = torch.cat([torch.cat([(i11 + i22) / 2, i1c, torch.zeros_like(i2c)], 1),
                       torch.cat([(i11 + i22) / 2, torch.zeros_like(i1c), i2c], 1)], 0)
        Y = torch.cat([torch.cat([o11, o1c, torch.zeros_like(o2c)], 1),
                       torch.cat([o22, torch.zeros_like(o1c), o2c], 1)], 0)
        return X, Y
    return torch.cat([(i11 + i22) / 2, i1c, i2c], 1), torch.cat([(o11 + o22) / 2, o1c, o2c], 1)


# ------------------------------------------------------------------ least-squares path

def _maps(idx1, idx2, w1, w2, device):
    """(chan1, chan2, scale1, scale2) device tensors from per-channel source indices (-1 = none)."""
    c1 = torch.as_tensor(idx1, dtype=torch.int32)
    c2 = torch.as_tensor(idx2, dtype=torch.int32)
    s1 = torch.as_tensor(w1, dtype=torch.float32)
    s2 = torch.as_tensor(w2, dtype=torch.float32)
    return tuple(t.to(device) for t in (c1, c2, s1, s2))


def _target_passes(b, merging, device, src_channels, is_input):
    """Sample groups of one side (inputs or targets) of the least-squares problem as channel maps
    for the pack kernel, one entry per group stacked along the batch axis by the reference
    (pleas_merging.py:125-147).  Returns (list of maps, merged channel count)."""
    b1, b2, b1c, b2c = (t.to("cpu", torch.int64).tolist() for t in b)
    nm, ms = len(b1), len(b1c)
    none, zeros, ones, half = [-1] * ms, [0.0] * ms, [1.0] * ms, [0.5] * nm
    if merging == "reg_mean":  # no permutation: the two models' samples are stacked as they are
        ident = list(range(src_channels))
        one = [1.0] * src_channels
        off = [-1] * src_channels
        zero = [0.0] * src_channels
        return [_maps(ident, off, one, zero, device), _maps(off, ident, zero, one, device)], src_channels
    if "perm_separatels" in merging or ("perm_mixedls" in merging and not is_input):
        # group 1: model 1 only [x1[b1], x1[b1c], 0]; group 2: model 2 only [x2[b2], 0, x2[b2c]]
        g1 = _maps(b1 + b1c + none, [-1] * (nm + 2 * ms), [1.0] * nm + ones + zeros, [0.0] * (nm + 2 * ms), device)
        g2 = _maps([-1] * (nm + 2 * ms), b2 + none + b2c, [0.0] * (nm + 2 * ms), [1.0] * nm + zeros + ones, device)
        return [g1, g2], nm + 2 * ms
    if "perm_mixedls" in merging:  # inputs: averaged merged units, each model's own separate units
        g1 = _maps(b1 + b1c + none, b2 + none + none, half + ones + zeros, half + zeros + zeros, device)
        g2 = _maps(b1 + none + none, b2 + none + b2c, half + zeros + zeros, half + zeros + ones, device)
        return [g1, g2], nm + 2 * ms
    # default 'perm_gradmask': one group, cat[(x1[b1] + x2[b2]) / 2, x1[b1c], x2[b2c]]
    return [_maps(b1 + b1c + none, b2 + none + b2c, half + ones + zeros, half + zeros + ones, device)], nm + 2 * ms


class _Workspace:
    """Staging shared by all layers (they run back to back on one stream): packed U / Y-bar planes
    and the K-split partial tiles.  ``version`` changes whenever a buffer is reallocated."""

    def __init__(self, device):
        self.device, self.cap, self.buf, self.version = device, {}, {}, 0
        self._retired = []  # outgrown buffers stay alive: captured CUDA graphs may point at them
        self.tables = ops.TableArena(device)  # GEMM problem entries through pinned staging (no synchronising copies)

    def ensure(self, name, floats):
        if self.cap.get(name, 0) < floats:
            self._retired.append(self.buf.get(name))
            self.cap[name] = int(floats * 1.1) + 1024
            self.buf[name] = torch.empty(self.cap[name], dtype=torch.float32, device=self.device)
            self.version += 1
        return self.buf[name]


class _LayerLS:
    """Normal-equation accumulator of one trained layer: G = U^T U, R = U^T Y-bar in fp64."""

    def __init__(self, name, layer, bi, bo, ip_shape, device, merging="perm_gradmask", cout_src=None):
        self.name = name
        self.is_conv = isinstance(layer, torch.nn.Conv2d)
        if self.is_conv:
            if layer.groups != 1:
                raise NotImplementedError(f"{name}: grouped convolutions are not supported by the closed form")
            if isinstance(layer.padding, str):
                raise NotImplementedError(f"{name}: string padding is not supported")
            self.kernel, self.stride = tuple(layer.kernel_size), tuple(layer.stride)
            self.padding, self.dilation = tuple(layer.padding), tuple(layer.dilation)
        else:
            if len(ip_shape) != 2:
                raise NotImplementedError(f"{name}: Linear inputs must be [batch, features]")
            self.kernel = self.stride = self.dilation = (1, 1)
            self.padding = (0, 0)
        self.has_bias = layer.bias is not None
        self.ipasses, self.cin = _target_passes(bi, merging, device, ip_shape[1], True)
        self.opasses, self.cout = _target_passes(bo, merging, device, cout_src, False)
        assert len(self.ipasses) == len(self.opasses)
        self.K = self.cin * self.kernel[0] * self.kernel[1] + int(self.has_bias)
        self.bi, self.bo = bi, bo
        self.G = self.R = None  # views into the runner's flat fp64 accumulator
        self.count = 0
        self.bound = {}  # (L, workspace version) -> (pu, py, plan_g, plan_r)

    def _geometry(self, n_cols, symmetric, kb):
        """(bn, m_tiles, n_tiles, splits) of the GEMM U^T [U | Y] for a batch with kb k-blocks."""
        bn = ops.choose_bn(n_cols)
        m_tiles, n_tiles = (self.K + 127) // 128, (n_cols + bn - 1) // bn
        active = m_tiles * n_tiles if not symmetric else sum(
            1 for mt in range(m_tiles) for nt in range(n_tiles) if 128 * mt + 127 >= bn * nt)
        return bn, m_tiles, n_tiles, ops.choose_splits(active, kb, 128, bn)

    def requirements(self, L):
        """(U-plane floats, Y-plane floats, partial floats) for a batch with L output positions."""
        kb = (L + 15) // 16
        rgu, rgy = ops.row_groups_of(self.K), ops.row_groups_of(self.cout)
        need = 0
        for n_cols, sym in ((self.K, True), (self.cout, False)):
            bn, m_tiles, n_tiles, splits = self._geometry(n_cols, sym, kb)
            need = max(need, splits * m_tiles * 128 * n_tiles * bn)
        return rgu * kb * 128, rgy * kb * 128, need

    def bind(self, L, ws):
        from .activation_matching import _View

        kb = (L + 15) // 16
        rgu, rgy = ops.row_groups_of(self.K), ops.row_groups_of(self.cout)
        pu = _View(ws.buf["u_hi"], ws.buf["u_lo"], self.K, rgu, kb)
        py = _View(ws.buf["y_hi"], ws.buf["y_lo"], self.cout, rgy, kb)
        plan_g = ops.GemmPlan(pu, pu, self.K, self.K, kb, symmetric=True, partial=ws.buf["partial"],
                              splits=self._geometry(self.K, True, kb)[3], tables=ws.tables)
        plan_r = ops.GemmPlan(pu, py, self.K, self.cout, kb, partial=ws.buf["partial"],
                              splits=self._geometry(self.cout, False, kb)[3], tables=ws.tables)
        plan_g.alg_flops = 2.0 * self.K * self.K * L
        plan_r.alg_flops = 2.0 * self.K * self.cout * L
        self.bound[L, ws.version] = (pu, py, plan_g, plan_r)

    def accumulate(self, acts1, acts2, ws):
        (ip1, op1), (ip2, op2) = acts1, acts2
        if not self.is_conv:
            ip1, ip2 = ip1[:, :, None, None], ip2[:, :, None, None]
            op1, op2 = op1[:, :, None, None], op2[:, :, None, None]
        B, _, Ho, Wo = op1.shape
        L = B * Ho * Wo
        pu, py, plan_g, plan_r = self.bound[L, ws.version]
        ip1, ip2, op1, op2 = ip1.float(), ip2.float(), op1.float(), op2.float()
        for imap, omap in zip(self.ipasses, self.opasses):  # one pass per stacked sample group
            ops.pack_im2col(ip1, ip2, *imap, self.cin, self.kernel, self.stride, self.padding, self.dilation,
                            (Ho, Wo), self.has_bias, pu)
            ops.pack_im2col(op1, op2, *omap, self.cout, (1, 1), (1, 1), (0, 0), (1, 1), (Ho, Wo), False, py)
            plan_g.run()
            plan_g.finalize(self.G, ops.MODE_INNER, accumulate=True)
            plan_r.run()
            plan_r.finalize(self.R, ops.MODE_INNER, accumulate=True)

    def out_positions(self, acts1):
        op = acts1[1]
        return op.shape[0] * (op.shape[2] * op.shape[3] if op.dim() == 4 else 1)

    def mask2d(self, wshape):
        """Gradient mask over the flattened [Co, K] weight (+ bias column)."""
        ni, mi, no, mo = len(self.bi[0]), len(self.bi[2]), len(self.bo[0]), len(self.bo[2])
        mask = torch.ones(wshape)
        if len(wshape) >= 2:
            mask[slice(ni, ni + mi), slice(no + mo, no + 2 * mo)] = 0.0
            mask[slice(ni + mi, ni + 2 * mi), slice(no, no + mo)] = 0.0
        mask = mask.reshape(wshape[0], -1)
        if self.has_bias:
            mask = torch.cat([mask, torch.ones(wshape[0], 1)], 1)
        return mask > 0

    def mask_classes(self, wshape):
        """The distinct rows of ``mask2d`` without materialising it: (patterns [P, K] bool, class id
        per output row [Co]).  The reference's mask (:52-59, index order kept) zeroes two
        row ranges against two ranges of dim 1, so there are at most three kinds of rows."""
        ni, mi, no, mo = len(self.bi[0]), len(self.bi[2]), len(self.bo[0]), len(self.bo[2])
        co = wshape[0]
        row_class = torch.zeros(co, dtype=torch.int64)
        if len(wshape) >= 2:
            for cls, (lo, hi) in enumerate(((ni, ni + mi), (ni + mi, ni + 2 * mi)), start=1):
                lo, hi = min(lo, co), min(hi, co)
                if hi > lo:
                    row_class[lo:hi] = cls
        pats, remap = [], {}
        per = 1
        for d in wshape[2:]:
            per *= d
        for cls in sorted(set(row_class.tolist())):
            m = torch.ones(wshape[1] if len(wshape) >= 2 else 1, per)
            if cls == 1:
                m[no + mo:no + 2 * mo] = 0.0
            elif cls == 2:
                m[no:no + mo] = 0.0
            m = m.reshape(-1)
            if self.has_bias:
                m = torch.cat([m, torch.ones(1)])
            m = m > 0
            same = [i for i, q in enumerate(pats) if torch.equal(q, m)]
            remap[cls] = same[0] if same else len(pats)
            if not same:
                pats.append(m)
        inverse = torch.tensor([remap[c] for c in row_class.tolist()], dtype=torch.int64)
        return torch.stack(pats), inverse

    def solve(self, layer, ridge_rel):
        """Returns the fitted flattened weight [Co, K] (fp64, CUDA) and the init it started from."""
        dev = self.G.device
        # the symmetric GEMM / finalize only maintain the tiles touching the lower triangle
        low = torch.tril(self.G)
        self.G.copy_(low + torch.tril(self.G, -1).T)
        del low
        W0 = layer.weight.detach().to(dev, torch.float64).reshape(self.cout, -1)
        if self.has_bias:
            W0 = torch.cat([W0, layer.bias.detach().to(dev, torch.float64)[:, None]], 1)
        assert W0.shape == (self.cout, self.K), f"{self.name}: merged layer shape {tuple(W0.shape)} != {(self.cout, self.K)}"
        grad = self.R - self.G @ W0.T  # [K, Co] residual of the normal equations at the init
        ridge = ridge_rel * float(self.G.diagonal().mean())
        W = W0.clone()
        patterns, inverse = self.mask_classes(tuple(layer.weight.shape))
        for pi in range(patterns.shape[0]):
            free = patterns[pi].to(dev)
            rows = (inverse == pi).nonzero().flatten().to(dev)
            if not bool(free.any()):
                continue
            idx = None if bool(free.all()) else free.nonzero().flatten()
            # The Gram matrix carries fp32-level noise (3xTF32, ~1e-6 of its entries); when the data
            # do not span all features (few samples, dead channels) the ridge must dominate that
            # noise, so a failed pivot escalates it tenfold instead of aborting.
            r = ridge
            for attempt in range(6):
                if idx is None:
                    Gf, rhs = self.G.clone(), grad[:, rows].contiguous()
                else:
                    Gf, rhs = self.G[idx][:, idx].contiguous(), grad[idx][:, rows].contiguous()
                info = int(ops.chol_solve_(Gf, rhs, r).item())
                if info == 0:
                    break
                r *= 10.0
            else:
                raise RuntimeError(f"{self.name}: normal equations not positive definite at pivot {info - 1} "
                                   f"even with ridge {r:.3e}")
            self.ridge_used = max(getattr(self, "ridge_used", 0.0), r)
            if idx is None:
                W[rows] = W0[rows] + rhs.T
            else:
                upd = torch.zeros(len(rows), self.K, dtype=torch.float64, device=dev)
                upd[:, idx] = rhs.T
                W[rows] = W0[rows] + upd
        return W, W0


def _record_layer(sink_id: int, name: str, in_a, in_b, out_a, out_b):
    """fx ``call_function`` target: hands a trained layer's inputs / outputs of both models to the runner."""
    _AM._SINKS[sink_id].record(name, in_a, in_b, out_a, out_b)
    return None


class LstsqRunner:
    """Streams calibration batches through both source models (forward hooks capture every
    trained layer's input and output) and accumulates all layers' normal equations; the
    per-batch pipeline is replayed as a CUDA graph."""

    def __init__(self, model1, model2, model3, perm_blocks, num_classes, separate_classifier, model_type,
                 use_cuda_graph=True, merging="perm_gradmask", world_size=1):
        self.merging, self.world_size = merging, world_size
        self.device = next(iter(model1.parameters())).device
        if self.device.type != "cuda":
            raise RuntimeError("pleas_merging_b200 runs on a CUDA device: move the models to cuda first")
        self.model1, self.model2 = model1, model2
        self.perm_blocks, self.num_classes = perm_blocks, num_classes
        self.separate_classifier, self.model_type = separate_classifier, model_type
        self.acts1, self.acts2 = {}, {}
        # Both source models run side by side through the activation-matching dual graph: the same layer of the two
        # models is one launch of the library's convolution kernel, with the eval-mode BatchNorm (+ ReLU) behind it
        # folded in, and every trained layer's (input, output) is recorded on the way.  Models fx cannot trace, or
        # that trace to different graphs, run as plain modules with forward hooks (convolutions patched one by one).
        self.gm, self.hooks, self.conv_patch, self.conv_pairs = None, [], None, []
        self.sink_id = next(_AM._sink_ids)
        if conv.ENABLED:
            try:
                _AM._SINKS[self.sink_id] = self
                self.gm = _AM._dual_graph(model1, model2, (), None, conv_pairs=self.conv_pairs, on_layer=self._on_layer)
            except Exception:  # not traceable: plain module calls
                self.gm = None
                self._release_graph()
        if self.gm is None:
            self.hooks = capture_inputs(model1, self.acts1) + capture_inputs(model2, self.acts2)
            self.conv_patch = conv.patched_convs(model1, model2).open()
        self.layers3 = {n: m for n, m in model3.named_modules() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))}
        self.accs, self.flat, self.ws = None, None, _Workspace(self.device)
        self.step = GraphedStep(self._eager, self._rebind, use_cuda_graph)
        self._last_L = {}
        self._L_by_shape = {}  # input shape -> {layer: output positions}; rows are counted per run(), not per launch

    def close(self):
        self.step.clear()
        self._release_graph()
        if self.conv_patch is not None:
            self.conv_patch.close()
        for h in self.hooks:
            h.remove()
        self.acts1.clear()
        self.acts2.clear()

    def _release_graph(self):
        _AM._SINKS.pop(self.sink_id, None)
        for pid in self.conv_pairs:
            _AM._CONV_PAIRS.pop(pid, None)
        self.conv_pairs = []

    def _on_layer(self, g, name, in_a, in_b, out_a, out_b):
        return g.call_function(_record_layer, (self.sink_id, name, in_a, in_b, out_a, out_b))

    def record(self, name, in_a, in_b, out_a, out_b):
        self.acts1[name] = (in_a, out_a)
        self.acts2[name] = (in_b, out_b)

    def _create(self):
        self.accs = {}
        for name, layer in self.layers3.items():
            if name not in self.acts1 or name not in self.acts2:
                print(f"Key error on {name}")  # same message as the reference (pleas_merging.py:274)
                continue
            bi, bo = _layer_blocks(self.perm_blocks, name, self.acts1[name][0].shape[1], self.num_classes,
                                   self.separate_classifier, self.model_type)
            self.accs[name] = _LayerLS(name, layer, bi, bo, tuple(self.acts1[name][0].shape), self.device,
                                       self.merging, self.acts1[name][1].shape[1])
        # one flat fp64 accumulator for every layer's [G | R].  Multi-GPU runs solve layer-parallel:
        # each layer has an owner rank (balanced by solve cost) and the buffer is laid out owner by
        # owner, so summing the normal equations is one reduce per owner instead of an all-reduce.
        accs = list(self.accs.values())
        owners = assign_owners([a.K ** 3 / 3.0 + float(a.K) ** 2 * a.cout for a in accs], self.world_size)
        total = sum(a.K * (a.K + a.cout) for a in accs)
        self.flat = torch.zeros(total, dtype=torch.float64, device=self.device)
        off = 0
        self.segments, self.segment_owners = [], []
        for r in range(self.world_size):
            for a, o in zip(accs, owners):
                if o != r:
                    continue
                a.owner, start = r, off
                a.G = self.flat[off:off + a.K * a.K].view(a.K, a.K)
                off += a.K * a.K
                a.R = self.flat[off:off + a.K * a.cout].view(a.K, a.cout)
                off += a.K * a.cout
                self.segments.append((start, off - start))
                self.segment_owners.append(r)

    def _bind_all(self):
        """Sizes the shared workspace for this batch shape and binds every layer's plans to it."""
        Ls = {n: a.out_positions(self.acts1[n]) for n, a in self.accs.items()}
        req = [a.requirements(Ls[n]) for n, a in self.accs.items()]
        for nm, floats in (("u_hi", max(r[0] for r in req)), ("u_lo", max(r[0] for r in req)),
                           ("y_hi", max(r[1] for r in req)), ("y_lo", max(r[1] for r in req)),
                           ("partial", max(r[2] for r in req))):
            self.ws.ensure(nm, floats)
        for n, a in self.accs.items():
            if (Ls[n], self.ws.version) not in a.bound:
                a.bind(Ls[n], self.ws)
        self._last_L = Ls

    def _rebind(self):
        pass  # _bind_all already bound everything against the final workspace of this shape

    def _eager(self, x):
        self.acts1.clear()
        self.acts2.clear()
        if self.gm is not None:
            self.gm(x)
        else:
            self.model1(x)
            self.model2(x)
        if self.accs is None:
            self._create()
        Ls = {n: a.out_positions(self.acts1[n]) for n, a in self.accs.items()}
        if Ls != self._last_L or any((Ls[n], self.ws.version) not in a.bound for n, a in self.accs.items()):
            self._bind_all()
        for n, a in self.accs.items():
            a.accumulate(self.acts1[n], self.acts2[n], self.ws)
        self._L_by_shape[tuple(x.shape)] = Ls

    def run(self, x):
        self.step(x)
        # sample rows seen by every layer (host bookkeeping: a CUDA-graph replay runs no Python)
        for n, L in self._L_by_shape[tuple(x.shape)].items():
            self.accs[n].count += L * len(self.accs[n].ipasses)


def layer_objective(W, G, R, yy_plus=0.0):
    """sum_o (w_o G w_o - 2 w_o r_o): the layer's squared error up to the constant ||Y||^2."""
    return float(((W @ G) * W).sum() - 2 * (W * R.T).sum()) + yy_plus


def _write_back(layer, acc, W):
    """Stores a fitted flattened [Co, K] weight (bias in the last column) into the layer."""
    kw = acc.K - int(acc.has_bias)
    layer.weight.data.copy_(W[:, :kw].reshape(layer.weight.shape).to(layer.weight.dtype))
    if acc.has_bias:
        layer.bias.data.copy_(W[:, kw].to(layer.bias.dtype))


def _train_lstsq(dataloader, model1, model2, model3, perm_blocks, MAX_STEPS, separate_classifier, num_classes,
                 model_type, ridge, verbose, stats, distributed=False, use_cuda_graph=True,
                 merging="perm_gradmask"):
    model1.eval()
    model2.eval()
    rank, wsize = world() if distributed else (0, 1)
    dbg = os.environ.get("PLB_DEBUG_TIMING") == "1"
    if dbg:
        torch.cuda.synchronize()
        t_dbg = time.perf_counter()
    runner = LstsqRunner(model1, model2, model3, perm_blocks, num_classes, separate_classifier, model_type,
                         use_cuda_graph, merging, wsize)
    if dbg:
        torch.cuda.synchronize()
        print(f"[plb timing] LstsqRunner construction {time.perf_counter() - t_dbg:.3f}s", flush=True)
    t_start = time.perf_counter()
    try:
        # the reference's loop breaks when idx > MAX_STEPS, i.e. it consumes MAX_STEPS + 1 batches
        sharder = BatchSharder(((b[0], 0) for b in dataloader), MAX_STEPS + 1, rank, wsize)
        with torch.no_grad():
            for bi_, x in device_prefetch(sharder, runner.device):
                runner.run(x)
                if dbg and bi_ < 4:
                    torch.cuda.synchronize()
                    print(f"[plb timing] PLeaS batch {bi_} done at +{time.perf_counter() - t_start:.3f}s", flush=True)
            if dbg:
                torch.cuda.synchronize()
                print(f"[plb timing] PLeaS loop end +{time.perf_counter() - t_start:.3f}s", flush=True)
            if wsize > 1:  # collective decision: a rank without a batch has no layer shapes to reduce into
                have = torch.tensor([int(runner.accs is not None)], device=runner.device)
                torch.distributed.all_reduce(have, op=torch.distributed.ReduceOp.MIN)
                if int(have.item()) == 0:
                    raise RuntimeError(f"distributed PLeaS needs at least one batch per rank "
                                       f"({sharder.total} batches for {wsize} ranks)")
            if runner.accs is None:
                return model3
            if wsize > 1:  # every owner receives the sum of its layers' [G | R]
                reduce_to_owners_(runner.flat, runner.segments, runner.segment_owners)
            if stats is not None:
                torch.cuda.synchronize()
                stats["_timing"] = {"accumulate_s": time.perf_counter() - t_start}
                t_solve = time.perf_counter()
            # fitted weights travel in one flat fp64 buffer ([Co, K] per layer): a rank fills the layers
            # it owns, the rest stays zero, and one all-reduce hands every rank every layer
            sizes = {n: a.cout * a.K for n, a in runner.accs.items()}
            wflat = torch.zeros(sum(sizes.values()), dtype=torch.float64, device=runner.device) if wsize > 1 else None
            woff, off = {}, 0
            for n in runner.accs:
                woff[n], off = off, off + sizes[n]
            local_stats = {}
            for name, acc in runner.accs.items():
                if acc.owner != rank:
                    continue
                layer = runner.layers3[name]
                W, W0 = acc.solve(layer, ridge)
                if stats is not None:
                    local_stats[name] = {"objective_init": layer_objective(W0, acc.G, acc.R),
                                         "objective_fit": layer_objective(W, acc.G, acc.R),
                                         "rows": acc.count, "cout": acc.cout, "K": acc.K, "owner": rank,
                                         "ridge_rel": getattr(acc, "ridge_used", 0.0)
                                         / max(float(acc.G.diagonal().mean()), 1e-300)}
                if verbose:
                    print(f"{name}: K={acc.K} Co={acc.cout} rows={acc.count}")
                if wflat is not None:
                    wflat[woff[name]:woff[name] + sizes[name]].copy_(W.reshape(-1))
                else:
                    _write_back(layer, acc, W)
            if wflat is not None:
                allreduce_sum_(wflat)
                for name, acc in runner.accs.items():
                    _write_back(runner.layers3[name], acc,
                                wflat[woff[name]:woff[name] + sizes[name]].view(acc.cout, acc.K))
            if stats is not None:
                if wsize > 1:
                    gathered = [None] * wsize
                    torch.distributed.all_gather_object(gathered, local_stats)
                    for part in gathered:
                        local_stats.update(part)
                    rows = torch.tensor([a.count for a in runner.accs.values()], dtype=torch.int64,
                                        device=runner.device)
                    allreduce_sum_(rows)
                    for n, r in zip(runner.accs, rows.tolist()):
                        local_stats[n]["rows"] = r
                for n in runner.accs:  # layer order, like the single-GPU run
                    stats[n] = local_stats[n]
                torch.cuda.synchronize()
                stats["_timing"]["solve_s"] = time.perf_counter() - t_solve
    finally:
        if dbg:
            torch.cuda.synchronize()
            t_dbg = time.perf_counter()
        runner.close()
        if dbg:
            torch.cuda.synchronize()
            print(f"[plb timing] runner.close {time.perf_counter() - t_dbg:.3f}s", flush=True)
    return model3


def _train_adam(dataloader, model1, model2, model3, perm_blocks, MAX_STEPS, separate_classifier, merging,
                num_classes, lr, verbose, model_type, WANDB, wandb_run):
    """The reference's optimisation loop (:234-302, 357-397) on the same hooks/targets/masks."""
    acts1, acts2 = {}, {}
    hooks = capture_inputs(model1, acts1) + capture_inputs(model2, acts2)
    model1.eval()
    model2.eval()
    model3, model3_dict, m3_params = get_model_dict_and_params(model3)
    optimizer = torch.optim.Adam(m3_params, lr=lr)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, MAX_STEPS)
    grad_masks = get_gradient_mask(perm_blocks, model3_dict)
    assert len(grad_masks) == len(m3_params)
    device = m3_params[0].device
    all_layer_loss = {k: 0 for k in model3_dict}
    tracking = 0.0
    try:
        for idx, batch in enumerate(dataloader):
            if idx > MAX_STEPS:
                break
            with torch.no_grad():
                x = batch[0].to(device)
                acts1.clear()
                acts2.clear()
                model1(x)
                model2(x)
            optimizer.zero_grad()
            total = 0.0
            for name, layer in model3_dict.items():
                if name not in acts1:
                    print(f"Key error on {name}")
                    continue
                bi, bo = _layer_blocks(perm_blocks, name, acts1[name][0].shape[1], num_classes,
                                       separate_classifier, model_type)
                X, Y = get_model_orig_activations(acts1[name], acts2[name], bi, bo, merging)
                loss = ((layer(X) - Y) ** 2).mean()
                total = total + loss
                all_layer_loss[name] += loss.detach()
            total.backward()
            for p, m in zip(m3_params, grad_masks):
                p.grad *= m
            optimizer.step()
            sched.step()
            tracking += float(total.detach())
            if idx % 20 == 0 and idx:
                print(f"Loss: {tracking / 20:.3f}")
                if WANDB:
                    metrics = {f"loss_{k}": v / 20 for k, v in all_layer_loss.items()}
                    metrics.update(step=idx, loss=tracking / 20)
                    wandb_run.log(metrics)
                all_layer_loss = {k: 0 for k in model3_dict}
                tracking = 0.0
        sd = model3.state_dict()
        for k, v in model3_dict.items():
            for k2, t in v.state_dict().items():
                sd[f"{k}.{k2}"] = t
        model3.load_state_dict(sd)
    finally:
        for h in hooks:
            h.remove()
    return model3


def train(dataloader, model1, model2, model3, spec, perm, costs, budget_ratios, WANDB, MAX_STEPS, wandb_run,
          separate_classifier=False, merging="perm_gradmask", num_classes=1000, lr=5e-4, verbose=False,
          model_type="rn50", *, solver="lstsq", ridge=1e-4, stats=None, distributed=False, use_cuda_graph=True):
    """Fit the merged model's layers to the source models' activations (reference :305-405).

    Same positional signature as the reference.  ``solver="lstsq"`` (default) is the closed
    form over the first ``MAX_STEPS + 1`` batches (the reference's loop consumes that many);
    ``solver="adam"`` replays the reference optimiser.  ``ridge`` is relative to the mean
    diagonal of each layer's Gram matrix (default 1e-4: layers with barely more sample rows than unknowns
    — a classifier fitted on a few hundred images — are otherwise dominated by directions the calibration
    data hardly excites; tests/test_configs_gpu.py measures the sensitivity) and is escalated tenfold when a
    pivot fails.  ``stats`` (dict) receives per-layer objectives.
    ``distributed=True`` (initialised torch.distributed job, same loader on every rank) deals the
    batches round-robin, reduces every layer's normal equations onto the rank that owns the layer
    (balanced by solve cost), solves layer-parallel and all-reduces the fitted weights."""
    blocks = get_blocks(spec, perm, costs, budget_ratios, False)
    perm_blocks = copy(blocks)
    for axis, pg in spec.items():
        for ax in pg.state:
            perm_blocks[ax] = perm_blocks[axis]
    if solver == "adam":
        return _train_adam(dataloader, model1, model2, model3, perm_blocks, MAX_STEPS, separate_classifier, merging,
                           num_classes, lr, verbose, model_type, WANDB, wandb_run)
    if solver != "lstsq":
        raise ValueError("solver must be 'lstsq' or 'adam'")
    return _train_lstsq(dataloader, model1, model2, model3, perm_blocks, MAX_STEPS, separate_classifier, num_classes,
                        model_type, ridge, verbose, stats, distributed, use_cuda_graph, merging)
