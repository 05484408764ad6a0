"""Permutation-spec compiler (host side, runs once per model).

Same contract as the reference's ``get_permutation_spec`` (pleas/core/compiler.py:786-797 and
the ``PermutationProp`` interpreter :28-783): the model is fx-traced and executed once; every
op contributes "these tensor axes are tied" relations; tied axes are merged in a disjoint-set
forest; a class that touches neither a model input nor the model output and contains at least
two state-dict axes is a permutation group, keyed by its earliest tensor in ``state_dict()``
order (ties on the same tensor towards the higher axis, :737-739).  Node names are fx's, so
the taps (``PermutationGroup.node``) are interchangeable with the reference's.

The implementation is table driven: a rule maps one executed op to a list of ties between
*endpoints* — ``("arg", path, axis)``, ``("self", axis)`` (method receiver), ``("out", axis)``,
``("state", name, axis)`` — which the tracer resolves to global axis names.
"""
import operator
import warnings
from typing import Any, Callable, Dict, List, Tuple

import torch
import torch.fx
import torch.nn.functional as F
from torch import nn

from .utils import Axis, InputsOrShapes, PermutationGroup, PermutationSpec, apply_perm

try:  # optional, only to recognise torchvision's stochastic depth as elementwise
    import torchvision.ops as _tvops
except Exception:  # pragma: no cover
    _tvops = None

Tie = Tuple[tuple, ...]


def arg(i, axis, *path):
    return ("arg", (i, *path), axis)


def kwarg(name, axis):
    return ("kwarg", name, axis)


def recv(axis):
    return ("self", axis)


def out(axis):
    return ("out", axis)


def state(name, axis):
    return ("state", name, axis)


class DisjointSets:
    """Union-find over hashable items (path halving + union by size)."""

    def __init__(self):
        self.parent: Dict[Any, Any] = {}
        self.size: Dict[Any, int] = {}

    def find(self, x):
        if x not in self.parent:
            self.parent[x] = x
            self.size[x] = 1
            return x
        while self.parent[x] != x:
            self.parent[x] = self.parent[self.parent[x]]
            x = self.parent[x]
        return x

    def union(self, a, b):
        ra, rb = self.find(a), self.find(b)
        if ra == rb:
            return
        if self.size[ra] < self.size[rb]:
            ra, rb = rb, ra
        self.parent[rb] = ra
        self.size[ra] += self.size[rb]

    def classes(self) -> List[set]:
        groups: Dict[Any, set] = {}
        for x in list(self.parent):
            groups.setdefault(self.find(x), set()).add(x)
        return list(groups.values())


# --------------------------------------------------------------------------------- rules

def _all_axes(src, nd):
    return [(src(i), out(i)) for i in range(nd)]


def rule_elementwise_arg0(_, args, kwargs, result):
    assert args[0].shape == result.shape
    return _all_axes(lambda i: arg(0, i), result.dim())


def rule_elementwise_self(obj, args, kwargs, result):
    assert obj.shape == result.shape
    return _all_axes(recv, result.dim())


def rule_conv2d(mod: nn.Conv2d, args, kwargs, result):
    (x,) = args
    assert not kwargs and x.dim() == 4
    ties = [(arg(0, 0), out(0)), (out(1), state("weight", 0))]
    if mod.groups == 1:
        ties.append((arg(0, 1), state("weight", 1)))
    else:  # depthwise: channels pass straight through
        assert mod.groups == mod.in_channels == x.shape[1] and x.shape[:2] == result.shape[:2]
        ties.append((arg(0, 1), out(1)))
    if mod.bias is not None:
        ties.append((out(1), state("bias", 0)))
    return ties


def rule_batchnorm2d(mod: nn.BatchNorm2d, args, kwargs, result):
    (x,) = args
    assert not kwargs and x.dim() == 4 and x.shape == result.shape
    ties = _all_axes(lambda i: arg(0, i), 4)
    present = set(dict(mod.named_parameters(recurse=False))) | set(dict(mod.named_buffers(recurse=False)))
    for name in ("weight", "bias", "running_mean", "running_var"):
        if name in present and getattr(mod, name) is not None:
            ties.append((arg(0, 1), state(name, 0)))
    return ties


def rule_pool(_, args, kwargs, result):
    return [(arg(0, 0), out(0)), (arg(0, 1), out(1))]


def rule_linear(mod: nn.Linear, args, kwargs, result):
    (x,) = args
    assert not kwargs and x.shape[:1] == result.shape[:1]
    last = x.dim() - 1
    ties = [(arg(0, i), out(i)) for i in range(last)]
    ties += [(arg(0, last), state("weight", 1)), (out(last), state("weight", 0))]
    if mod.bias is not None:
        ties.append((out(last), state("bias", 0)))
    return ties


def _broadcast_ties(x, y, ex, ey, result):
    """Ties for an elementwise binary op under numpy broadcasting; ex/ey build endpoints."""
    xs = tuple(x.shape) if torch.is_tensor(x) else ()
    ys = tuple(y.shape) if torch.is_tensor(y) else ()
    r = max(len(xs), len(ys))
    ties = []
    for i in range(r):
        xa, ya = i - (r - len(xs)), i - (r - len(ys))
        hx, hy = xa >= 0, ya >= 0
        if hx and hy and xs[xa] == ys[ya]:
            ties += [(ex(xa), ey(ya)), (ex(xa), out(i))]
        elif hx and (not hy or ys[ya] == 1):
            ties.append((ex(xa), out(i)))
        elif hy and (not hx or xs[xa] == 1):
            ties.append((ey(ya), out(i)))
        else:
            raise AssertionError(f"shapes {xs} and {ys} do not broadcast")
    return ties


def rule_binop(_, args, kwargs, result):
    assert len(args) == 2 and not kwargs
    return _broadcast_ties(args[0], args[1], lambda a: arg(0, a), lambda a: arg(1, a), result)


def rule_binop_method(obj, args, kwargs, result):
    assert len(args) == 1 and not kwargs
    return _broadcast_ties(obj, args[0], recv, lambda a: arg(0, a), result)


def rule_matmul(_, args, kwargs, result):
    x, y = args
    assert not kwargs and x.dim() == y.dim() == 2
    return [(arg(0, 0), out(0)), (arg(0, 1), arg(1, 0)), (arg(1, 1), out(1))]


def _flatten_ties(x, start, end):
    nd = x.dim()
    start, end = start % nd, end % nd
    live = [i for i in range(start, end + 1) if x.shape[i] != 1]
    assert len(live) == 1, "flatten may only squeeze unit axes"
    ties = [(arg(0, i), out(i)) for i in range(start)]
    ties.append((arg(0, live[0]), out(start)))
    ties += [(arg(0, i), out(i - (end - start))) for i in range(end + 1, nd)]
    return ties


def rule_flatten_fn(_, args, kwargs, result):
    assert 1 <= len(args) <= 3 and not kwargs
    return _flatten_ties(args[0], args[1] if len(args) > 1 else 0, args[2] if len(args) > 2 else -1)


def rule_flatten_mod(mod: nn.Flatten, args, kwargs, result):
    return _flatten_ties(args[0], mod.start_dim, mod.end_dim)


def rule_none(*_):
    return []


def rule_getitem(_, args, kwargs, result):
    if not torch.is_tensor(result):
        return []
    x, index = args
    if isinstance(x, (tuple, list)):
        assert isinstance(index, int)
        return [(arg(0, i, index), out(i)) for i in range(result.dim())]
    index = index if isinstance(index, tuple) else (index,)
    ties, src, dst = [], 0, 0
    for it in index:
        if it is None:
            dst += 1
            continue
        assert isinstance(it, (slice, int))
        if isinstance(it, slice):
            assert it == slice(None)
            ties.append((arg(0, src), out(dst)))
            dst += 1
        src += 1
    while src < x.dim():
        ties.append((arg(0, src), out(dst)))
        src, dst = src + 1, dst + 1
    return ties


def rule_reshape(obj, args, kwargs, result):
    """Axes survive a reshape when they are preserved as whole factors at matching offsets."""
    before, after = list(obj.shape), list(result.shape)
    ties, i, j, pi, pj = [], 0, 0, 1, 1
    while i < len(before) and j < len(after):
        if pi == pj and before[i] == after[j]:
            ties.append((recv(i), out(j)))
            pi = pj = 1
            i, j = i + 1, j + 1
        elif pj <= pi:
            pj *= after[j]
            j += 1
        else:
            pi *= before[i]
            i += 1
    return ties


def rule_expand(obj, args, kwargs, result):
    sizes = args[0] if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)) else args
    nb, na = obj.dim(), result.dim()
    return [(recv(nb - i), out(na - i)) for i in range(1, nb + 1) if sizes[-i] == -1]


def rule_permute_fn(_, args, kwargs, result):
    x, dims = args
    return [(arg(0, dims[i]), out(i)) for i in range(x.dim())]


def rule_permute_method(obj, args, kwargs, result):
    dims = args[0] if len(args) == 1 and isinstance(args[0], (tuple, list)) else args
    return [(recv(dims[i]), out(i)) for i in range(obj.dim())]


def rule_tile(obj, args, kwargs, result):
    n, m = len(args), obj.dim()
    assert not kwargs and set(args[max(0, n - m):]) == {1}
    return [(recv(i), out(i + max(n - m, 0))) for i in range(m)]


def rule_reduce(obj, args, kwargs, result):
    if not args and "dim" not in kwargs:
        return []
    dims = args[0] if args else kwargs["dim"]
    dims = [dims] if not isinstance(dims, (list, tuple)) else list(dims)
    dims = [d % obj.dim() for d in dims]
    keep = bool(kwargs.get("keepdim") or (len(args) >= 2 and args[1]))
    ties, o = [], 0
    for i in range(obj.dim()):
        if i in dims:
            o += 1 if keep else 0
        else:
            ties.append((recv(i), out(o)))
            o += 1
    return ties


def rule_cat(_, args, kwargs, result):
    tensors = args[0]
    dim = kwargs["dim"] if "dim" in kwargs else (args[1] if len(args) > 1 else 0)
    dim %= result.dim()
    return [(arg(0, j, i), out(j)) for i in range(len(tensors)) for j in range(result.dim()) if j != dim]


def _layernorm_ties(x, weight_ep, bias_ep, weight, bias):
    n = x.dim()
    ties = []
    for t, ep in ((weight, weight_ep), (bias, bias_ep)):
        if t is not None:
            m = t.dim()
            ties += [(ep(i), arg(0, n - m + i)) for i in range(m)]
    return ties + _all_axes(lambda i: arg(0, i), n)


def rule_layernorm_fn(_, args, kwargs, result):
    x = args[0]
    w = args[2] if len(args) >= 3 else kwargs.get("weight")
    b = args[3] if len(args) >= 4 else kwargs.get("bias")
    wep = (lambda i: arg(2, i)) if len(args) >= 3 else (lambda i: kwarg("weight", i))
    bep = (lambda i: arg(3, i)) if len(args) >= 4 else (lambda i: kwarg("bias", i))
    return _layernorm_ties(x, wep, bep, w, b)


def rule_layernorm_mod(mod: nn.LayerNorm, args, kwargs, result):
    w = mod.weight if mod.elementwise_affine else None
    b = mod.bias if mod.elementwise_affine else None
    return _layernorm_ties(args[0], lambda i: state("weight", i), lambda i: state("bias", i), w, b)


def rule_mha(mod, args, kwargs, result):
    warnings.warn("MultiheadAttention is not supported by the permutation compiler; its axes are left untied")
    return []


MODULE_RULES: Dict[type, Callable] = {
    nn.Conv2d: rule_conv2d, nn.BatchNorm2d: rule_batchnorm2d, nn.Linear: rule_linear,
    nn.MaxPool2d: rule_pool, nn.AdaptiveAvgPool2d: rule_pool, nn.AvgPool2d: rule_pool,
    nn.Flatten: rule_flatten_mod, nn.LayerNorm: rule_layernorm_mod, nn.MultiheadAttention: rule_mha,
    **{m: rule_elementwise_arg0 for m in (nn.ReLU, nn.GELU, nn.Identity, nn.Dropout, nn.Sigmoid, nn.SiLU)},
}
FUNCTION_RULES: Dict[Callable, Callable] = {
    **{f: rule_elementwise_arg0 for f in (torch.sigmoid, F.gelu, F.relu, torch.relu, torch.sqrt)},
    **{f: rule_binop for f in (operator.add, operator.sub, operator.mul, operator.truediv)},
    F.adaptive_avg_pool2d: rule_pool, F.avg_pool2d: rule_pool, operator.matmul: rule_matmul,
    torch.flatten: rule_flatten_fn, getattr: rule_none, operator.getitem: rule_getitem,
    torch.permute: rule_permute_fn, torch.cat: rule_cat, F.layer_norm: rule_layernorm_fn,
}
if _tvops is not None:
    FUNCTION_RULES[_tvops.stochastic_depth] = rule_elementwise_arg0
METHOD_RULES: Dict[str, Callable] = {
    **{m: rule_binop_method for m in ("mul", "add", "sub")},
    **{m: rule_reshape for m in ("reshape", "view")},
    **{m: rule_elementwise_self for m in ("to", "type", "pow", "sqrt", "contiguous", "float", "clone", "detach")},
    **{m: rule_reduce for m in ("sum", "mean")},
    "expand": rule_expand, "permute": rule_permute_method, "tile": rule_tile,
    "size": rule_none, "dim": rule_none,
}


# --------------------------------------------------------------------------------- tracer

class AxisTracer(torch.fx.Interpreter):
    """Executes the traced module once and unions tied axes."""

    def __init__(self, gm: torch.fx.GraphModule, verbose=False):
        super().__init__(gm)
        self.sets = DisjointSets()
        self.inputs: List[torch.Tensor] = []
        self.result = None
        self.verbose = verbose

    # -- endpoint resolution
    def _resolve(self, n: torch.fx.Node, ep) -> str:
        kind = ep[0]
        if kind == "out":
            return f"node.{n.name}:{ep[1]}"
        if kind == "state":
            return f"state.{n.target}.{ep[1]}:{ep[2]}"
        if kind == "self":
            return f"node.{n.args[0].name}:{ep[1]}"
        if kind == "kwarg":
            return f"node.{n.kwargs[ep[1]].name}:{ep[2]}"
        path, axis = ep[1], ep[2]
        args = n.args[1:] if n.op == "call_method" else n.args
        sub, tail = args[path[0]], list(path[1:])
        while not isinstance(sub, torch.fx.Node):  # descend into list/tuple literals
            sub = sub[tail.pop(0)]
        suffix = "".join(f".{t}" for t in tail)  # index into a node that returns a tuple
        return f"node.{sub.name}{suffix}:{axis}"

    def _apply(self, n, ties):
        for tie in ties:
            names = [self._resolve(n, ep) for ep in tie]
            for other in names[1:]:
                self.sets.union(names[0], other)
            if len(names) == 1:
                self.sets.find(names[0])

    def _log(self, kind, target, args):
        if self.verbose:
            print(kind, target, [tuple(a.shape) if torch.is_tensor(a) else a for a in args])

    # -- node kinds
    def run_node(self, n):
        self._node = n
        return super().run_node(n)

    def placeholder(self, target, args, kwargs):
        value = super().placeholder(target, args, kwargs)
        self.inputs.append(value)
        for i in range(value.dim()):
            self.sets.union(f"placeholders.{len(self.inputs) - 1}:{i}", f"node.{self._node.name}:{i}")
        return value

    def get_attr(self, target, args, kwargs):
        value = super().get_attr(target, args, kwargs)
        assert target in self.module.state_dict()
        for i in range(value.dim()):
            self.sets.union(f"state.{target}:{i}", f"node.{self._node.name}:{i}")
        return value

    def output(self, target, args, kwargs):
        value = super().output(target, args, kwargs)
        assert torch.is_tensor(value), "the model must return a single tensor"
        self.result = value
        for i in range(value.dim()):
            self.sets.union(f"result:{i}", f"node.{self._node.args[0].name}:{i}")
        return value

    def call_module(self, target, args, kwargs):
        mod = self.fetch_attr(target)
        self._log("call_module", mod, args)
        value = mod(*args, **kwargs)
        rule = MODULE_RULES.get(type(mod))
        if rule is None:
            raise NotImplementedError(f"No handler for module {mod}")
        self._apply(self._node, rule(mod, args, kwargs, value))
        return value

    def call_function(self, target, args, kwargs):
        self._log("call_function", target, args)
        value = target(*args, **kwargs)
        if not any(torch.is_tensor(a) for a in (*args, *kwargs.values())) and not any(
                isinstance(a, (list, tuple)) and any(torch.is_tensor(t) for t in a) for a in args):
            return value
        rule = FUNCTION_RULES.get(target)
        if rule is None:
            raise NotImplementedError(f"No handler for function {target}")
        self._apply(self._node, rule(target, args, kwargs, value))
        return value

    def call_method(self, target, args, kwargs):
        obj, *rest = args
        self._log("call_method", target, args)
        value = getattr(obj, target)(*rest, **kwargs)
        rule = METHOD_RULES.get(target)
        if rule is None:
            raise NotImplementedError(f"No handler for method call {type(obj)}.{target}")
        self._apply(self._node, rule(obj, rest, kwargs, value))
        return value

    # -- spec extraction
    def permutation_spec(self) -> PermutationSpec:
        io = {f"placeholders.{i}:{j}" for i, t in enumerate(self.inputs) for j in range(t.dim())}
        io |= {f"result:{i}" for i in range(self.result.dim())}
        sd = self.module.state_dict()
        order = {k: i for i, k in enumerate(sd.keys())}

        def parse(name, prefix):
            key, axis = name[len(prefix):].rsplit(":", 1)
            return Axis(key, int(axis))

        spec: PermutationSpec = {}
        for cls in self.sets.classes():
            if cls & io:
                continue
            st = {parse(x, "state.") for x in cls if x.startswith("state.")}
            if len(st) <= 1:
                continue
            nodes = {parse(x, "node.") for x in cls if x.startswith("node.")}
            sizes = {sd[a.key].shape[a.axis] for a in st}
            assert len(sizes) == 1, f"inconsistent sizes {sizes} in one permutation group"
            size = sizes.pop()
            if size == 1:
                continue
            key = min(st, key=lambda a: (order[a.key], -a.axis))
            spec[key] = PermutationGroup(size, st, nodes)
        return spec


def get_permutation_spec(model: nn.Module, inputs_or_shapes: InputsOrShapes, verbose=False) -> PermutationSpec:
    """Drop-in for pleas.core.compiler.get_permutation_spec (compiler.py:786-797)."""
    device = next(iter(model.parameters())).device
    inputs = [torch.randn(*ios).to(device) if isinstance(ios, tuple) else ios.to(device) for ios in inputs_or_shapes]
    gm = torch.fx.symbolic_trace(model)
    tracer = AxisTracer(gm, verbose=verbose)
    with torch.no_grad():
        tracer.run(*inputs)
    return tracer.permutation_spec()


def check_permutation_spec(model: nn.Module, spec: PermutationSpec, x, rtol=1e-2, atol=1e-3, generator=None):
    """Functional-invariance self-check (the reference's test_permutation_spec,
    compiler.py:754-783): permuting any single group must leave the model's function unchanged.
    Returns the set of failing group keys (empty = pass)."""
    from copy import deepcopy

    saved = deepcopy(model.state_dict())
    failed = set()
    with torch.no_grad():
        ref = model(x)
        for key, pg in spec.items():
            P = torch.randperm(pg.size, generator=generator)
            apply_perm({key: P}, spec, model, inplace=True)
            if not torch.allclose(model(x), ref, rtol=rtol, atol=atol):
                failed.add(key)
            model.load_state_dict(saved)
    return failed
