"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch on the
B200 box, gloo in CPU tests).

The merge path shards only where the work splits naturally: calibration
batches are dealt round-robin at WHOLE-batch granularity (the -cdist statistic takes a sqrt per
batch, so a batch is never split) and the additive accumulators — group cost matrices, PLeaS
normal equations — are summed with ONE all-reduce at the end.  There is no collective inside
the data path.  The per-layer least-squares solves are independent, so in a multi-GPU job they
are dealt to owner ranks (``assign_owners``): every owner receives the sum of its layers'
normal equations (``reduce_to_owners_``), solves them, and the fitted weights are exchanged with
one more all-reduce of a flat weight buffer in which each rank filled only its own layers.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class BatchSharder:
    """Iterates this rank's batches and remembers how many batches existed in total."""

    def __init__(self, loader, num_batches, rank=None, world_size=None):
        r, w = world()
        self.rank = r if rank is None else rank
        self.world = w if world_size is None else world_size
        self.loader, self.num_batches, self.total = loader, num_batches, 0

    def __iter__(self):
        self.total = 0
        for idx, (batch, _) in enumerate(zip(self.loader, range(self.num_batches))):
            self.total = idx + 1
            if idx % self.world == self.rank:
                yield idx, batch

    def owns_last(self):
        """True on the rank that processed the globally last batch (reference accumulate mode)."""
        return self.total > 0 and (self.total - 1) % self.world == self.rank


def allreduce_sum_(tensor):
    """In-place sum over ranks (no-op for a single process)."""
    if world()[1] > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def assign_owners(costs, world_size):
    """Deals independent work items (per-layer solves, weight ~ K^3/3 + K^2 Co) to ranks:
    longest-processing-time-first greedy — heaviest item to the least-loaded rank, ties to the
    lowest rank — so every rank computes the same assignment without communicating."""
    load = [0.0] * world_size
    owner = [0] * len(costs)
    for i in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
        r = min(range(world_size), key=lambda r: (load[r], r))
        owner[i] = r
        load[r] += costs[i]
    return owner


def reduce_to_owners_(flat, segments, owners):
    """Sums each ``(offset, length)`` segment of the flat accumulator onto its owner rank
    (runs of consecutive segments with one owner travel as one reduce).  On the other ranks the
    segment's content is unspecified afterwards."""
    if world()[1] == 1:
        return flat
    i = 0
    while i < len(segments):
        j = i
        while j + 1 < len(segments) and owners[j + 1] == owners[i] and \
                segments[j + 1][0] == segments[j][0] + segments[j][1]:
            j += 1
        lo, hi = segments[i][0], segments[j][0] + segments[j][1]
        dist.reduce(flat[lo:hi], dst=owners[i], op=dist.ReduceOp.SUM)
        i = j + 1
    return flat


def combine_costs_(flat, sharder, accumulate):
    """Combines per-rank cost accumulators: a sum in "sum" mode; in "reference" mode (only the
    last processed batch counts) the owner of the global last batch wins."""
    if sharder.world == 1:
        return flat
    if accumulate == "reference" and not sharder.owns_last():
        flat.zero_()
    return allreduce_sum_(flat)


def device_prefetch(indexed_batches, device):
    """Yields (index, x_on_device) one batch ahead: the host-to-device copy of batch i+1 is issued
    on a side stream while batch i computes, so PCIe time is hidden behind the kernels.

    The device copies live in a ring of three buffers allocated on the CONSUMER's stream: a fresh side stream has
    no cached blocks, so allocating there meant a (device-synchronising) cudaMalloc per batch until the allocator
    warmed up — 0.1-0.3 s of jitter on a 20-batch call (profiles/experiments/e2e_variance.py).  The yielded tensor
    is valid until two more batches have been requested."""
    device = torch.device(device)
    main = torch.cuda.current_stream(device)
    side = torch.cuda.Stream(device)
    ring, free_ev, count = [None] * 3, [None] * 3, [0]

    def stage(item):
        idx, (x, _) = item
        if x.device == device:
            return idx, x, None, None
        slot = count[0] % 3
        count[0] += 1
        buf = ring[slot]
        if buf is None or buf.shape != x.shape or buf.dtype != x.dtype:
            buf = ring[slot] = torch.empty(x.shape, dtype=x.dtype, device=device)  # consumer-stream pool
            side.wait_stream(main)  # the block may have just been released by work still in flight on `main`
        elif free_ev[slot] is not None:
            side.wait_event(free_ev[slot])  # the consumer of the slot's previous batch has finished
        with torch.cuda.stream(side):
            buf.copy_(x, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(side)
        return idx, buf, ev, slot

    it = iter(indexed_batches)
    try:
        cur = stage(next(it))
    except StopIteration:
        return
    while cur is not None:
        try:
            nxt = stage(next(it))
        except StopIteration:
            nxt = None
        idx, x, ev, slot = cur
        if ev is not None:
            main.wait_event(ev)
        yield idx, x
        if slot is not None:  # the consumer has enqueued its work on `main`
            free_ev[slot] = torch.cuda.Event()
            free_ev[slot].record(main)
        cur = nxt
    if any(b is not None for b in ring):
        for b in ring:
            if b is not None:
                b.record_stream(side)
