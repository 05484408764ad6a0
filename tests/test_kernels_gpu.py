"""Kernel-level parity tests: every CUDA kernel, called through the C ABI, against the CPU
oracle / numpy on seeded inputs.  Bit-exact for integer and index work, stated tolerances for
floating point."""
import numpy as np
import pytest
import torch

from conftest import key_str, load_spec_json
from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from pleas_merging_b200 import ops

    return ops


def unpack_planes(plane, rows, K, row_groups, kb_offset=0):
    """numpy view of one packed plane back to [rows, K] (include/pleas_b200.h layout)."""
    p = plane.cpu().numpy()
    r = np.arange(rows)[:, None]
    k = np.arange(K)[None, :]
    kb, j, e = k // 16 + kb_offset, (k % 16) // 4, k % 4
    off = ((kb * row_groups + r // 8) * 4 + j) * 32 + (r % 8) * 4 + e
    return p[off]


@pytest.mark.parametrize("shape,axis", [((3, 20, 5, 4), 1), ((2, 12, 7, 7), 1), ((5, 24), 1), ((24, 12, 3, 3), 0),
                                        ((24, 12, 3, 3), 1), ((2, 130, 4, 4), 1)])
def test_pack_split_layout_and_stats(shape, axis):
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(*shape, generator=g)
    xd = x.cuda()
    outer, rows, inner = ops.as_rows_view(xd, axis)
    K = outer * inner
    planes = ops.Planes(rows, (K + 15) // 16, xd.device)
    planes.hi.fill_(float("nan"))
    planes.lo.fill_(float("nan"))
    q = torch.zeros(2, rows, dtype=torch.float64, device=xd.device)
    ops.pack_split(xd, axis, planes, sumsq=q[0], rowsum=q[1])
    ref = O._rows(x.numpy(), axis)
    hi = unpack_planes(planes.hi, rows, K, planes.row_groups)
    lo = unpack_planes(planes.lo, rows, K, planes.row_groups)
    assert (hi.view(np.uint32) & 0x1FFF == 0).all() and (lo.view(np.uint32) & 0x1FFF == 0).all()
    assert np.abs(hi - ref).max() <= 2.0 ** -11 * np.abs(ref).max()
    assert np.abs((hi.astype(np.float64) + lo) - ref).max() <= 2.0 ** -21 * np.abs(ref).max()
    # zero padding in k for real rows
    kpad = planes.k_blocks * 16
    if kpad > K:
        full = unpack_planes(planes.hi, rows, kpad, planes.row_groups)
        assert (full[:, K:] == 0).all()
    np.testing.assert_allclose(q[0].cpu().numpy(), (ref.astype(np.float64) ** 2).sum(1), rtol=1e-6)
    np.testing.assert_allclose(q[1].cpu().numpy(), ref.astype(np.float64).sum(1), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape,axis", [((3, 20, 5, 4), 1), ((2, 130, 7, 7), 1), ((5, 24), 1), ((32, 64, 16, 16), 1)])
def test_pack_split_pair_equals_two_single_packs(shape, axis):
    """One launch for both operands of a tap == two plb_pack_split launches, bit for bit."""
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    xa, xb = torch.randn(*shape, generator=g).cuda(), torch.randn(*shape, generator=g).cuda()
    outer, rows, inner = ops.as_rows_view(xa, axis)
    kb = (outer * inner + 15) // 16
    single = [ops.Planes(rows, kb, xa.device) for _ in range(2)]
    pair = [ops.Planes(rows, kb, xa.device) for _ in range(2)]
    for p in single + pair:
        p.hi.zero_()
        p.lo.zero_()
    q1 = torch.zeros(2, rows, dtype=torch.float64, device=xa.device)
    q2 = torch.zeros(2, rows, dtype=torch.float64, device=xa.device)
    ops.pack_split(xa, axis, single[0], sumsq=q1[0])
    ops.pack_split(xb, axis, single[1], sumsq=q1[1])
    ops.pack_split_pair(xa, xb, axis, pair[0], pair[1], q2[0], q2[1])
    for a, b in zip(single, pair):
        assert torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo)
    torch.testing.assert_close(q1, q2, rtol=1e-12, atol=0)


def test_pack_split_row_gather():
    ops = _ops()
    x = torch.randn(4, 10, 6, generator=torch.Generator().manual_seed(1))
    P = torch.randperm(10, generator=torch.Generator().manual_seed(2))
    xd = x.cuda()
    planes = ops.Planes(10, 2, xd.device)
    ops.pack_split(xd, 1, planes, row_index=P.cuda())
    hi = unpack_planes(planes.hi, 10, 24, planes.row_groups)
    lo = unpack_planes(planes.lo, 10, 24, planes.row_groups)
    ref = O._rows(x.numpy(), 1)[P.numpy()]
    assert np.abs(hi.astype(np.float64) + lo - ref).max() <= 2.0 ** -21 * np.abs(ref).max()


GEMM_CASES = [((2, 12, 5, 5), 1), ((4, 64, 16, 16), 1), ((3, 200, 9, 10), 1), ((2, 300, 7, 7), 1),
              ((32, 24), 1), ((1, 128, 64, 64), 1), ((2, 520, 4, 4), 1)]


@pytest.mark.parametrize("impl", ["simt", "tcgen05_v1", "tcgen05"])
@pytest.mark.parametrize("shape,axis", GEMM_CASES)
def test_cross_statistic_inner_and_cdist(shape, axis, impl, monkeypatch):
    ops = _ops()
    monkeypatch.setattr(ops, "DIRECT_MAX_ROWS", 0)  # the packed-plane path (pack + GEMM); fused path below
    g = torch.Generator().manual_seed(3)
    x = torch.relu(torch.randn(*shape, generator=g)) + 0.1 * torch.randn(*shape, generator=g)
    y = torch.relu(torch.randn(*shape, generator=g))
    ops.set_gemm_impl(impl)
    try:
        G = ops.cross_statistic(x.cuda(), y.cuda(), axis, ops.MODE_INNER).cpu().numpy()
        D = ops.cross_statistic(x.cuda(), y.cuda(), axis, ops.MODE_NEG_CDIST).cpu().numpy()
    finally:
        ops.set_gemm_impl("tcgen05")
    X, Y = O._rows(x.numpy(), axis).astype(np.float64), O._rows(y.numpy(), axis).astype(np.float64)
    Gref = X @ Y.T
    # north-star bar is rel-err <= 1e-4; 3xTF32 holds ~1e-6
    assert np.abs(G - Gref).max() <= 2.5e-6 * np.abs(Gref).max()
    d2 = (X * X).sum(1)[:, None] + (Y * Y).sum(1)[None] - 2 * Gref
    Dref = -np.sqrt(np.maximum(d2, 0))
    assert np.abs(D - Dref).max() <= 1e-4 * np.abs(Dref).max()
    # and against the fp32 oracle restatement of the reference operator
    assert np.abs(D - O.cross_neg_cdist(x.numpy(), y.numpy(), axis)).max() <= 1e-4 * np.abs(Dref).max()


DIRECT_CASES = [((4, 64, 16, 16), 1), ((1, 128, 64, 64), 1), ((6, 24, 4, 16), 1), ((2, 72, 16, 16), 1),
                ((32, 64, 56, 56), 1), ((3, 8, 16, 16), 1), ((20, 128, 4, 4), 1), ((7, 40, 48), 1)]


@pytest.mark.parametrize("shape,axis", DIRECT_CASES)
def test_fused_narrow_tap_kernel(shape, axis, monkeypatch):
    """plb_gram_direct (round-1 narrow-tap kernel, cp.async loaders; kept for A/B runs behind PLB_TMA_GRAM=0)
    == fp64 reference at 3xTF32 accuracy, == the packed path to fp32 rounding, for both statistics; rows
    that are not a multiple of the tile and K ranges that split unevenly included."""
    ops = _ops()
    monkeypatch.setattr(ops, "TMA_GRAM", False)
    g = torch.Generator().manual_seed(5)
    x = (torch.relu(torch.randn(*shape, generator=g)) + 0.1 * torch.randn(*shape, generator=g)).cuda()
    y = torch.relu(torch.randn(*shape, generator=g)).cuda()
    assert ops.direct_gram_eligible(x, y, axis)
    before = ops.N.LAUNCH_COUNTS["plb_gram_direct"]
    G = ops.cross_statistic(x, y, axis, ops.MODE_INNER)
    D = ops.cross_statistic(x, y, axis, ops.MODE_NEG_CDIST)
    assert ops.N.LAUNCH_COUNTS["plb_gram_direct"] == before + 2
    old, ops.DIRECT_MAX_ROWS = ops.DIRECT_MAX_ROWS, 0
    try:
        Gp = ops.cross_statistic(x, y, axis, ops.MODE_INNER)
        Dp = ops.cross_statistic(x, y, axis, ops.MODE_NEG_CDIST)
    finally:
        ops.DIRECT_MAX_ROWS = old
    X = O._rows(x.cpu().numpy(), axis).astype(np.float64)
    Y = O._rows(y.cpu().numpy(), axis).astype(np.float64)
    Gref = X @ Y.T
    assert np.abs(G.cpu().numpy() - Gref).max() <= 2.5e-6 * np.abs(Gref).max()
    Dref = -np.sqrt(np.maximum((X * X).sum(1)[:, None] + (Y * Y).sum(1)[None] - 2 * Gref, 0))
    assert np.abs(D.cpu().numpy() - Dref).max() <= 1e-4 * np.abs(Dref).max()
    assert (G - Gp).abs().max() <= 2e-6 * Gp.abs().max()
    assert (D - Dp).abs().max() <= 1e-4 * Dp.abs().max()


TMA_CASES = [((4, 64, 16, 16), 1), ((1, 128, 64, 64), 1), ((2, 72, 16, 16), 1), ((32, 64, 56, 56), 1),
             ((3, 8, 16, 16), 1), ((20, 128, 4, 4), 1), ((7, 40, 48), 1), ((2, 256, 14, 14), 1),
             ((3, 512, 28, 28), 1), ((4, 320, 8, 8), 1), ((4, 1024, 14, 14), 1), ((2, 256, 56, 56), 1),
             ((8, 136, 6, 6), 1)]


@pytest.mark.parametrize("shape,axis", TMA_CASES)
def test_tma_fused_gram_kernel(shape, axis):
    """plb_gram_tma (TMA boxes straight from the activations, in-kernel hi/lo split; cta_group::2 tile
    pairs above 128 channels) == fp64 reference at 3xTF32 accuracy and == the packed path to fp32
    rounding, for both statistics.  Covers channel counts that are not a tile multiple (TMA zero-fills
    the rows), spatial extents that are not a multiple of the 32-wide box (14x14 = 196, 6x6 = 36, 48)
    and several output tiles with K splits."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    x = (torch.relu(torch.randn(*shape, generator=g)) + 0.1 * torch.randn(*shape, generator=g)).cuda()
    y = torch.relu(torch.randn(*shape, generator=g)).cuda()
    assert ops.tma_gram_eligible(x, y, axis)
    before = ops.N.LAUNCH_COUNTS["plb_gram_tma"]
    G = ops.cross_statistic(x, y, axis, ops.MODE_INNER)
    D = ops.cross_statistic(x, y, axis, ops.MODE_NEG_CDIST)
    assert ops.N.LAUNCH_COUNTS["plb_gram_tma"] == before + 2
    old, oldd = ops.TMA_GRAM, ops.DIRECT_MAX_ROWS
    ops.TMA_GRAM, ops.DIRECT_MAX_ROWS = False, 0
    try:
        Gp = ops.cross_statistic(x, y, axis, ops.MODE_INNER)
        Dp = ops.cross_statistic(x, y, axis, ops.MODE_NEG_CDIST)
    finally:
        ops.TMA_GRAM, ops.DIRECT_MAX_ROWS = old, oldd
    X = O._rows(x.cpu().numpy(), axis).astype(np.float64)
    Y = O._rows(y.cpu().numpy(), axis).astype(np.float64)
    Gref = X @ Y.T
    assert np.abs(G.cpu().numpy() - Gref).max() <= 2.5e-6 * np.abs(Gref).max()
    Dref = -np.sqrt(np.maximum((X * X).sum(1)[:, None] + (Y * Y).sum(1)[None] - 2 * Gref, 0))
    assert np.abs(D.cpu().numpy() - Dref).max() <= 1e-4 * np.abs(Dref).max()
    assert (G - Gp).abs().max() <= 2e-6 * Gp.abs().max()
    assert (D - Dp).abs().max() <= 1e-4 * Dp.abs().max()


@pytest.mark.parametrize("splits", [1, 3, 7])
def test_tma_fused_gram_any_split_and_row_norms(splits):
    """Explicit K splits of plb_gram_tma give the same Gram (fp32 rounding apart) and exact row norms."""
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(6, 256, 10, 12, generator=g).cuda()
    y = torch.randn(6, 256, 10, 12, generator=g).cuda()
    plan = ops.TmaGramPlan(256, 6, 120, x.device, splits=splits)
    q = torch.zeros(2, 256, dtype=torch.float64, device=x.device)
    plan.run(x, y, 1, q[0], q[1])
    out = torch.empty(256, 256, device=x.device)
    plan.finalize(out, ops.MODE_INNER)
    X = O._rows(x.cpu().numpy(), 1).astype(np.float64)
    Y = O._rows(y.cpu().numpy(), 1).astype(np.float64)
    Gref = X @ Y.T
    assert np.abs(out.cpu().numpy() - Gref).max() <= 2.5e-6 * np.abs(Gref).max()
    np.testing.assert_allclose(q[0].cpu().numpy(), (X * X).sum(1), rtol=1e-6)
    np.testing.assert_allclose(q[1].cpu().numpy(), (Y * Y).sum(1), rtol=1e-6)


@pytest.mark.parametrize("C", [64, 256])
def test_raw_fp32_word_as_tf32_operand_is_truncated(C):
    """The fused kernels feed the RAW fp32 word as the tf32 "hi" operand and take lo against
    trunc_tf32(x): correct only if kind::tf32 ignores (truncates) the 13 low mantissa bits.  Every value
    here has low bits 0x1800 set, where round-to-nearest tf32 differs from truncation by a full tf32
    ulp: a tensor core that rounded internally would be off by ~2^-11 relative, 500x the bound."""
    ops = _ops()
    g = torch.Generator().manual_seed(13)
    shape = (4, C, 16, 16)
    bits = torch.randint(0, 1 << 23, shape, generator=g, dtype=torch.int32)
    bits = (bits & ~0x1FFF) | 0x1800 | (127 << 23)  # 1.xxx with the bits just below the tf32 mantissa set
    x = bits.view(torch.float32).cuda().contiguous()
    y = torch.flip(x, dims=[1]).contiguous()
    for eligible, fn in ((ops.tma_gram_eligible(x, y, 1), None),):
        assert eligible
    G = ops.cross_statistic(x, y, 1, ops.MODE_INNER).cpu().numpy()
    X = O._rows(x.cpu().numpy(), 1).astype(np.float64)
    Y = O._rows(y.cpu().numpy(), 1).astype(np.float64)
    Gref = X @ Y.T
    assert np.abs(G - Gref).max() <= 2.5e-6 * np.abs(Gref).max()


CORR_CASES = [((4, 64, 16, 16), 1), ((3, 200, 9, 10), 1), ((2, 256, 14, 14), 1), ((6, 24), 1), ((2, 300, 7, 7), 1),
              ((3, 512, 12, 12), 1)]


@pytest.mark.parametrize("shape,axis", CORR_CASES)
def test_cross_statistic_correlation(shape, axis):
    """PLB_MODE_CORR (cross-Gram + fused row sums / sums of squares + fp64 epilogue) == numpy.corrcoef in
    float64 to 1e-4 absolute (north-star bar; measured ~1e-6), through the TMA-fed kernel where the tap is
    eligible and the packed path otherwise; a dead unit (all zeros) correlates with nothing."""
    ops = _ops()
    g = torch.Generator().manual_seed(17)
    x = torch.relu(torch.randn(*shape, generator=g) + 0.5)  # post-ReLU: large means against the spread
    y = torch.relu(torch.randn(*shape, generator=g)) + 0.3 * x
    x.select(axis, 1).zero_()  # a dead channel
    C = ops.cross_statistic(x.cuda(), y.cuda(), axis, ops.MODE_CORR).cpu().numpy()
    ref = O.cross_corr(x.numpy(), y.numpy(), axis)
    assert np.isfinite(C).all() and (C[1] == 0).all()
    assert np.abs(C - ref).max() <= 1e-4
    assert np.abs(C - ref).max() <= 2e-5  # what 3xTF32 + fp64 moments actually deliver


@pytest.mark.parametrize("impl", ["tcgen05", "tcgen05_v1"])
def test_gemm_long_k_accuracy_and_splits(impl):
    """Long contractions stay at fp32-level accuracy for any K split: the persistent kernel
    promotes every MAX_CHAIN_KB k-blocks into registers (the tensor core's accumulator truncates:
    -1e-7 relative per chained k-block); the v1 kernel needs its default chain-bounding split."""
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    x, y = torch.randn(8, 96, 32, 32, generator=g).cuda(), torch.randn(8, 96, 32, 32, generator=g).cuda()
    kb = 8 * 1024 // 16
    pa, pb = ops.Planes(96, kb, x.device), ops.Planes(96, kb, x.device)
    ops.pack_split(x, 1, pa)
    ops.pack_split(y, 1, pb)
    X = O._rows(x.cpu().numpy(), 1).astype(np.float64)
    Y = O._rows(y.cpu().numpy(), 1).astype(np.float64)
    ref = X @ Y.T
    errs = {}
    for splits in (None, 1, 3, 32, 512):
        plan = ops.GemmPlan(pa, pb, 96, 96, kb, splits=splits)
        plan.run(impl)
        out = torch.empty(96, 96, device=x.device)
        plan.finalize(out)
        errs[splits] = np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max()
    ok = (None, 1, 3, 32, 512) if impl == "tcgen05" else (32, 512)
    for splits in ok:
        assert errs[splits] <= 2.5e-6, (impl, errs)
    if impl == "tcgen05_v1":
        assert 2.5e-6 < errs[1] <= 1e-4, errs  # one 512-block chain: inside the matrix bar, not the perm bar


def test_gemm_persistent_many_items():
    """More work items than SMs: every CTA of the persistent kernel walks several tiles."""
    ops = _ops()
    g = torch.Generator().manual_seed(8)
    x = torch.relu(torch.randn(2, 1300, 12, 12, generator=g)).cuda()
    y = torch.relu(torch.randn(2, 1300, 12, 12, generator=g)).cuda()
    G = ops.cross_statistic(x, y, 1, ops.MODE_INNER).cpu().numpy()
    X = O._rows(x.cpu().numpy(), 1).astype(np.float64)
    Y = O._rows(y.cpu().numpy(), 1).astype(np.float64)
    ref = X @ Y.T
    assert np.abs(G - ref).max() <= 2.5e-6 * np.abs(ref).max()


def test_finalize_accumulates_fp32_and_fp64():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(2, 40, 6, 6, generator=g).cuda(), torch.randn(2, 40, 6, 6, generator=g).cuda()
    kb = (72 + 15) // 16
    pa, pb = ops.Planes(40, kb, x.device), ops.Planes(40, kb, x.device)
    ops.pack_split(x, 1, pa)
    ops.pack_split(y, 1, pb)
    plan = ops.GemmPlan(pa, pb, 40, 40, kb)
    plan.run()
    acc32 = torch.ones(40, 40, device=x.device)
    acc64 = torch.ones(40, 40, device=x.device, dtype=torch.float64)
    plan.finalize(acc32, accumulate=True)
    plan.finalize(acc64, accumulate=True)
    ref = 1.0 + O.cross_inner(x.cpu().numpy(), y.cpu().numpy(), 1)
    np.testing.assert_allclose(acc32.cpu().numpy(), ref, rtol=0, atol=2e-5 * np.abs(ref).max())
    np.testing.assert_allclose(acc64.cpu().numpy(), ref, rtol=0, atol=2e-5 * np.abs(ref).max())


def test_lap_golden_instances(lap_golden):
    ops = _ops()
    for maximize in (True, False):
        names = [n for n, g in lap_golden.items() if bool(g["maximize"]) == maximize]
        costs = [torch.from_numpy(lap_golden[n]["A"]).cuda() for n in names]
        outs, obj, status = ops.lap_solve_batched(costs, maximize)
        assert (status.cpu() == 0).all()
        for n, o, ob in zip(names, outs, obj.cpu().tolist()):
            assert (o.cpu().numpy() == lap_golden[n]["col"]).all(), n  # identical to SciPy, ties included
            assert ob == pytest.approx(float(lap_golden[n]["obj"]), rel=1e-12, abs=1e-12), n


@pytest.mark.parametrize("n", [1, 2, 31, 257, 1000, 2048])
def test_lap_random_vs_oracle(n):
    ops = _ops()
    rng = np.random.default_rng(n)
    mats = [rng.standard_normal((n, n)).astype(np.float32), rng.integers(0, 5, (n, n)).astype(np.float32)]
    if n >= 31:
        X = rng.standard_normal((n, 64)).astype(np.float32)
        Y = X[rng.permutation(n)] + 0.3 * rng.standard_normal((n, 64)).astype(np.float32)
        mats.append(O.cross_neg_cdist(X, Y, 0))
    outs, obj, status = ops.lap_solve_batched([torch.from_numpy(m).cuda() for m in mats], True)
    assert (status.cpu() == 0).all()
    for m, o in zip(mats, outs):
        col, _ = O.solve_lsa(m, True)
        assert (o.cpu().numpy() == col).all()


def test_lap_groups_wider_than_2048_and_mixed_batch():
    """Groups of 2049..4096 units take the four-columns-per-thread instantiation of lap_kernel_v3; a small problem in
    the same launch keeps one warp of the 1024-thread block.  Integer costs: ties everywhere."""
    ops = _ops()
    rng = np.random.default_rng(2500)
    mats = [rng.standard_normal((2500, 2500)).astype(np.float32), rng.integers(0, 7, (40, 40)).astype(np.float32),
            rng.integers(0, 50, (2100, 2100)).astype(np.float32)]
    for maximize in (True, False):
        outs, obj, status = ops.lap_solve_batched([torch.from_numpy(m).cuda() for m in mats], maximize)
        assert (status.cpu() == 0).all()
        for m, o in zip(mats, outs):
            col, _ = O.solve_lsa(m, maximize)
            assert (o.cpu().numpy() == col).all(), (m.shape, maximize)


def test_lap_strided_and_invalid():
    ops = _ops()
    big = torch.randn(40, 64, generator=torch.Generator().manual_seed(6)).cuda()
    view = big[:40, :40]  # leading dimension 64
    outs, _, status = ops.lap_solve_batched([view], True)
    assert (outs[0].cpu().numpy() == O.solve_lsa(view.cpu().numpy(), True)[0]).all()
    bad = torch.ones(4, 4).cuda()
    bad[1, 2] = float("nan")
    _, _, status = ops.lap_solve_batched([bad], True)
    assert status.cpu().tolist() == [2]
    with pytest.raises(ValueError):
        ops.raise_on_lap_status(status)


@pytest.mark.parametrize("n,ratio", [(12, 0.0), (12, 0.5), (24, 0.3), (24, 1.0), (100, 0.7), (1000, 0.25), (4096, 0.5)])
def test_get_blocks_vs_oracle(n, ratio):
    ops = _ops()
    rng = np.random.default_rng(n)
    C = rng.standard_normal((n, n)).astype(np.float32)
    C[3 % n] = C[5 % n]  # duplicate row
    P = rng.permutation(n)
    spec = [{"key": ("w", 0), "size": n, "state": [("w", 0)], "node": []}]
    ref = O.get_blocks(spec, {("w", 0): P}, {("w", 0): C}, ratio)[("w", 0)]
    identity = abs(ratio - 1.0) < 1e-3
    got = ops.get_blocks_group(torch.from_numpy(C).cuda(), torch.from_numpy(P).cuda(), ratio, identity)
    for a, b in zip(got, ref):
        assert (a.cpu().numpy() == b).all()


@pytest.mark.parametrize("name", ["r0", "r05", "r1", "mixed"])
def test_block_merge_vs_reference_state(tiny_golden, name):
    """Merged tensors equal the reference's partial_merge state dict bit for bit."""
    from oracle import tinynet

    ops = _ops()
    spec = load_spec_json("tiny")
    m1, m2 = tinynet.make_pair(12, 10)
    blocks = {g["key"]: [t.cuda() for t in tiny_golden[f"pm/{name}/blocks"][key_str(g["key"])]] for g in spec}
    per_axis = O.blocks_by_state_axis(spec, blocks)
    by_tensor = {}
    for (tname, ax), b in per_axis.items():
        by_tensor.setdefault(tname, {})[ax] = b
    sd1, sd2 = m1.state_dict(), m2.state_dict()
    gold = tiny_golden[f"pm/{name}/state"]
    for tname, bl in by_tensor.items():
        out = ops.block_merge(sd1[tname].cuda(), sd2[tname].cuda(), bl)
        assert tuple(out.shape) == tuple(gold[tname].shape), tname
        assert torch.equal(out.cpu(), gold[tname]), tname


def test_gather_compose_progress():
    ops = _ops()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(5, 9, 3, 3, generator=g)
    P = torch.randperm(9, generator=g)
    for axis in (0, 1):
        xx = x if axis == 1 else x.transpose(0, 1).contiguous()
        assert torch.equal(ops.gather_axis(xx.cuda(), axis, P.cuda()).cpu(), torch.index_select(xx, axis, P))
    a, b = torch.randperm(50, generator=g), torch.randperm(50, generator=g)
    assert torch.equal(ops.compose_perm(a.cuda(), b.cuda()).cpu(), a[b])
    A = torch.randn(50, 50, generator=g)
    flag = torch.zeros(1, dtype=torch.int32).cuda()
    gain = torch.zeros(1, dtype=torch.float64).cuda()
    ops.wm_progress(A.cuda(), torch.arange(50).cuda(), flag, gain)
    assert flag.item() == 0 and gain.item() == 0.0
    col, _ = O.solve_lsa(A.numpy(), True)
    ops.wm_progress(A.cuda(), torch.from_numpy(col).cuda(), flag, gain)
    assert flag.item() == 1
    assert gain.item() == pytest.approx(float(A.double()[torch.arange(50), col].sum() - A.double().diag().sum()))


@pytest.mark.parametrize("n,nrhs", [(1, 1), (5, 3), (33, 7), (100, 130), (128, 1), (129, 33), (257, 2), (300, 64),
                                    (1000, 257)])
def test_chol_solve_vs_numpy(n, nrhs):
    ops = _ops()
    rng = np.random.default_rng(n)
    U = rng.standard_normal((2 * n + 3, n))
    G = U.T @ U
    B = rng.standard_normal((n, nrhs))
    ridge = 1e-3
    Gd, Bd = torch.from_numpy(G).cuda(), torch.from_numpy(B.copy()).cuda()
    info = ops.chol_solve_(Gd, Bd, ridge)
    assert info.item() == 0
    ref = np.linalg.solve(G + ridge * np.eye(n), B)
    np.testing.assert_allclose(Bd.cpu().numpy(), ref, rtol=1e-8, atol=1e-10 * np.abs(ref).max())
    L = np.tril(Gd.cpu().numpy())
    np.testing.assert_allclose(L @ L.T, G + ridge * np.eye(n), rtol=1e-10, atol=1e-10 * np.abs(G).max())


def test_chol_reports_non_spd():
    ops = _ops()
    G = torch.eye(40, dtype=torch.float64)
    G[17, 17] = -1.0
    info = ops.chol_solve_(G.cuda(), torch.ones(40, 2, dtype=torch.float64).cuda(), 0.0)
    assert info.item() == 18


@pytest.mark.parametrize("impl", ["simt", "tcgen05_v1", "tcgen05"])
@pytest.mark.parametrize("rows,K", [(40, 100), (300, 500), (700, 2000)])
def test_symmetric_gram_lower_tiles_mirrored(rows, K, impl):
    """U U^T with symmetric=True computes only the tiles touching the lower triangle; the
    finalize kernel mirrors the rest — result equals the full product (fp64 accumulator)."""
    ops = _ops()
    g = torch.Generator().manual_seed(rows)
    u = torch.randn(rows, K, generator=g)
    kb = (K + 15) // 16
    pu = ops.Planes(rows, kb, "cuda")
    ops.pack_split(u.cuda(), 0, pu)
    # the v1 kernel does not promote: bound its accumulation chain through the split count
    plan = ops.GemmPlan(pu, pu, rows, rows, kb, symmetric=True, splits=max(1, kb // 4) if impl == "tcgen05_v1" else None)
    plan.run(impl)
    out = torch.full((rows, rows), 1.0, dtype=torch.float64, device="cuda")
    plan.finalize(out, accumulate=True)
    ref = 1.0 + u.double().numpy() @ u.double().numpy().T
    got = out.cpu().numpy()
    low = np.tril(np.ones_like(ref, dtype=bool))
    assert np.abs(got - ref)[low].max() <= 2.5e-6 * np.abs(ref).max()  # lower triangle incl. diagonal
    bn = plan.bn  # tiles that do not touch the lower triangle are left untouched (still 1.0)
    i, j = np.indices(ref.shape)
    untouched = 128 * (i // 128) + 127 < bn * (j // bn)
    assert (got[untouched] == 1.0).all()


def test_pack_im2col_vs_unfold():
    import torch.nn.functional as F

    ops = _ops()
    g = torch.Generator().manual_seed(11)
    x1, x2 = torch.randn(3, 10, 9, 8, generator=g), torch.randn(3, 10, 9, 8, generator=g)
    b1, b2 = torch.randperm(10, generator=g)[:6], torch.randperm(10, generator=g)[:6]
    b1c = torch.tensor([i for i in range(10) if i not in b1.tolist()])
    b2c = torch.tensor([i for i in range(10) if i not in b2.tolist()])
    xbar = torch.cat([(x1[:, b1] + x2[:, b2]) / 2, x1[:, b1c], x2[:, b2c]], 1)  # pleas_merging.py:146
    # the fourth and fifth cases take the kernel's 1x1 / stride-1 fast path (H * W = 72 is a multiple of 4:
    # 128-bit loads), with and without the bias row
    for kernel, stride, pad, dil, bias in (((3, 3), (1, 1), (1, 1), (1, 1), False), ((1, 1), (2, 2), (0, 0), (1, 1), True),
                                           ((3, 2), (2, 1), (0, 1), (1, 2), True), ((1, 1), (1, 1), (0, 0), (1, 1), True),
                                           ((1, 1), (1, 1), (0, 0), (1, 1), False)):
        U = F.unfold(xbar, kernel, dil, pad, stride)  # [B, K, L']
        Ho = (9 + 2 * pad[0] - dil[0] * (kernel[0] - 1) - 1) // stride[0] + 1
        Wo = (8 + 2 * pad[1] - dil[1] * (kernel[1] - 1) - 1) // stride[1] + 1
        ref = U.permute(1, 0, 2).reshape(U.shape[1], -1)  # rows (c,dy,dx), k = (n, ho, wo)
        if bias:
            ref = torch.cat([ref, torch.ones(1, ref.shape[1])], 0)
        neg = lambda n: torch.full((n,), -1, dtype=torch.int64)
        c1 = torch.cat([b1, b1c, neg(4)]).int().cuda()
        c2 = torch.cat([b2, neg(4), b2c]).int().cuda()
        s1 = torch.cat([torch.full((6,), 0.5), torch.ones(4), torch.zeros(4)]).cuda()
        s2 = torch.cat([torch.full((6,), 0.5), torch.zeros(4), torch.ones(4)]).cuda()
        rows, Kc = ref.shape
        planes = ops.Planes(rows, (Kc + 15) // 16, "cuda")
        ops.pack_im2col(x1.cuda(), x2.cuda(), c1, c2, s1, s2, 14, kernel, stride, pad, dil, (Ho, Wo), bias, planes)
        hi = unpack_planes(planes.hi, rows, Kc, planes.row_groups)
        lo = unpack_planes(planes.lo, rows, Kc, planes.row_groups)
        got = hi.astype(np.float64) + lo
        assert np.abs(got - ref.numpy()).max() <= 2.0 ** -21 * max(np.abs(ref.numpy()).max(), 1.0)


@pytest.mark.parametrize("n", [5, 64, 257, 1024])
def test_lap_warm_start_reaches_the_same_optimum(n):
    """plb_lap_solve_batched_warm: started from the column duals of a related problem (or from arbitrary duals)
    the solver returns the optimal assignment — the cold start's / SciPy's whenever the optimum is unique."""
    from scipy.optimize import linear_sum_assignment

    ops = _ops()
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)).astype(np.float32)
    Ad = torch.from_numpy(A).cuda()
    cold, obj0, st0, v = ops.lap_solve_batched([Ad], True, return_duals=True)
    assert int(st0.item()) == 0
    assert (cold[0].cpu().numpy() == linear_sum_assignment(A, maximize=True)[1]).all()
    # (1) the same problem from its own duals: nothing left to augment
    again, obj1, st1 = ops.lap_solve_batched([Ad], True, v_init=[v[0]])
    assert torch.equal(again[0], cold[0]) and float(obj1.item()) == pytest.approx(float(obj0.item()), rel=1e-12)
    # (2) a perturbed, column-permuted problem from the permuted duals (what weight matching does between sweeps)
    P = torch.from_numpy(rng.permutation(n)).cuda()
    B = (Ad[:, P] + 0.05 * torch.from_numpy(rng.standard_normal((n, n)).astype(np.float32)).cuda()).contiguous()
    ref = linear_sum_assignment(B.cpu().numpy(), maximize=True)[1]
    warm, objw, stw = ops.lap_solve_batched([B], True, v_init=[v[0][P].contiguous()])
    assert int(stw.item()) == 0 and (warm[0].cpu().numpy() == ref).all()
    # (3) arbitrary duals and a scale: still the optimum
    junk = torch.from_numpy(rng.standard_normal(n)).cuda()
    w2, _, _ = ops.lap_solve_batched([B], True, v_init=[junk], v_scale=3.0)
    assert (w2[0].cpu().numpy() == ref).all()
    # (4) mixed batch: one warm, one cold problem in the same launch
    w3, _, _ = ops.lap_solve_batched([B, Ad], True, v_init=[v[0][P].contiguous(), None])
    assert (w3[0].cpu().numpy() == ref).all() and torch.equal(w3[1], cold[0])
