"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in CPU tests).

The hot path shards trivially: every query, ray and training point is independent and the
40 MB model is replicated.  So:
  * queries / rays  -> contiguous ranges per rank, NO data-path collective; an optional final gather of results;
    rays are x-major (ray = ix*H + iy), so a contiguous ray range is a vertical strip of the image;
  * training        -> each rank takes its slice of the point batch and runs the fused step with the loss scaled by
    the GLOBAL batch; then (lib/trainer.py: FusedTrainer) reduce-scatter of the flat fp32 gradient buffer (40.6 MB),
    Adam on this rank's 1/N of the parameters, all-gather of the parameters -- or, `sharded=False`, ONE all-reduce(sum)
    of the flat gradient and replicated Adam.
The reference has no distributed code at all, so nothing here replaces reference lines.
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (no-op for world size 1)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, **kw)
    return rank, world, local


def bind_to_gpu_numa_node(local_rank):
    """Best effort: pin this process to the CPU cores of the NUMA node its GPU hangs off, BEFORE it page-locks host
    buffers (first touch places the pages).  With one process per GPU the eight ranks of a box otherwise allocate
    their pinned frame buffers wherever the launcher left them, and every D2H copy crosses the socket interconnect.
    Returns the node id, or None when the topology cannot be read (then nothing is changed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:                       # 00000000:1b:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:  # noqa: BLE001
        pass
    return None


def shard_range(n, rank, world, align=1):
    """[start, end) of the contiguous shard `rank` of `n` items; boundaries are multiples of `align`
    (use align=H to cut an x-major image on column boundaries).  Shards differ by at most one `align` unit."""
    units = (n + align - 1) // align
    base, rem = divmod(units, world)
    u0 = rank * base + min(rank, rem)
    u1 = u0 + base + (1 if rank < rem else 0)
    return min(u0 * align, n), min(u1 * align, n)


def interleaved_strips(n_cols, rank, world, strips_per_rank=4):
    """Column ranges for load-balanced rendering: the image is cut into world*strips_per_rank vertical strips
    dealt round-robin, so a rank gets both silhouette-heavy and empty regions.  Returns [(c0, c1), ...]."""
    total = min(n_cols, world * strips_per_rank)
    out = []
    for s in range(rank, total, world):
        c0, c1 = shard_range(n_cols, s, total)
        if c1 > c0:
            out.append((c0, c1))
    return out


def allreduce_sum_(flat):
    """In-place sum over ranks of one flat buffer (the whole gradient in a single bucket)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def gather_shards(local, n_total, rank, world, dst=0, align=1):
    """Gather contiguous shards (as cut by shard_range) to rank `dst`; returns the full tensor there, else None.
    Shards may differ in length, so they are padded to the longest and trimmed after the gather."""
    if world == 1 or not dist.is_initialized():
        return local
    sizes = [shard_range(n_total, r, world, align) for r in range(world)]
    longest = max(e - s for s, e in sizes)
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[: e - s] for b, (s, e) in zip(bufs, sizes)], dim=0)


def gather_strips(local, strips_by_rank, rows_per_col, rank, world, dst=0):
    """Final gather of a frame rendered in interleaved column strips (interleaved_strips): `local` holds this rank's
    strips back to back ([cols_local * rows_per_col, C], x-major like the rays); rank `dst` gets the whole frame
    [total_cols * rows_per_col, C] in image order, the others None.  ONE NCCL gather; when every strip has the same
    width (the usual case: the image width divides by world * strips_per_rank) the receive buffer is a single
    [world, strips_per_rank, ...] tensor and the frame is one transposed copy of it, otherwise shards are padded to
    the longest and copied strip by strip."""
    if world == 1 or not dist.is_initialized():
        return local
    widths = {c1 - c0 for s in strips_by_rank for c0, c1 in s}
    counts = {len(s) for s in strips_by_rank}
    tail = tuple(local.shape[1:])
    if len(widths) == 1 and len(counts) == 1:
        # strip s of the image belongs to rank s % world and is that rank's strip s // world
        big = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device) if rank == dst else None
        dist.gather(local.contiguous(), list(big.unbind(0)) if rank == dst else None, dst=dst)
        if rank != dst:
            return None
        spr, rows = counts.pop(), widths.pop() * rows_per_col
        return big.view((world, spr, rows) + tail).transpose(0, 1).reshape((world * spr * rows,) + tail)
    cols = [sum(c1 - c0 for c0, c1 in s) for s in strips_by_rank]
    longest = max(cols) * rows_per_col
    pad = torch.zeros((longest,) + tail, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    total = sum(cols)
    full = torch.empty((total * rows_per_col,) + tail, dtype=local.dtype, device=local.device)
    for r in range(world):
        pos = 0
        for c0, c1 in strips_by_rank[r]:
            nrow = (c1 - c0) * rows_per_col
            full[c0 * rows_per_col:c1 * rows_per_col] = bufs[r][pos:pos + nrow]
            pos += nrow
    return full


def max_over_ranks(value, device):
    """Max of a python float over ranks (used for device-timed durations)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
