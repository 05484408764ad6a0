"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so
nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this at run
time; it is used by ``oracle/make_golden.py`` to create the committed fixtures
under ``tests/golden/`` and by the ``not gpu`` tests that pin the numpy
restatement (``oracle/ref_oracle.py``) when the reference tree is present.

Two third-party modules the reference imports at module top are absent here
(``gurobipy`` — pleas/methods/partial_matching.py:4, ``torchmetrics`` —
pleas/methods/pleas_merging.py:2); they are stubbed.  The reference hard-codes
``.cuda()`` (activation_matching.py:121, partial_matching.py:86,
pleas_merging.py:169,266,277); on a CPU-only host those calls are neutralised
inside this process only.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PLEAS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pleas"))


def load_reference():
    """Returns a namespace with the reference's hot-path callables."""
    import torch

    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    if "gurobipy" not in sys.modules:
        g = types.ModuleType("gurobipy")
        g.GRB = types.SimpleNamespace(CONTINUOUS=0, MAXIMIZE=1)

        def _no_gurobi(*a, **k):
            raise RuntimeError("gurobipy is not installed")

        g.Model = _no_gurobi
        sys.modules["gurobipy"] = g
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")
        tm.Accuracy = object
        sys.modules["torchmetrics"] = tm

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    import importlib

    # pleas/methods/__init__.py:12-35 re-exports functions named like the submodules, so the
    # submodules must be fetched by their dotted names.
    mod = importlib.import_module
    return types.SimpleNamespace(
        compiler=mod("pleas.core.compiler"), solvers=mod("pleas.core.solvers"),
        utils=mod("pleas.core.utils"), am=mod("pleas.methods.activation_matching"),
        pm=mod("pleas.methods.partial_matching"), pl=mod("pleas.methods.pleas_merging"),
        wm=mod("pleas.methods.weight_matching"),
    )


def accumulate_fixed_costs(ref, spec, gm_cross, dataloader, num_batches):
    """The reference's compute_matching_costs (activation_matching.py:103-136) with the
    one-token F1 fix (membership tested with ``Axis(*ka)``), i.e. the paper-intended sum
    over batches.  Implemented here, not by editing /root/reference."""
    import torch

    Axis = ref.utils.Axis
    cross_sum = {}
    with torch.inference_mode():
        for (x, _), _ in zip(dataloader, range(num_batches)):
            x = x.cuda()
            _, cross = gm_cross(x)
            for ka, v in cross.items():
                k = Axis(*ka)
                if k not in cross_sum:
                    cross_sum[k] = v.clone()
                else:
                    cross_sum[k].add_(v)
    return {
        next(kax for kax in spec.keys() if kax in pg.state): sum(
            cross_sum[nax] for nax in pg.node if nax in cross_sum
        )
        for pg in spec.values()
    }
