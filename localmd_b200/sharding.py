"""Frame sharding of one movie over the GPUs of a node.

One process per GPU, `torch.distributed` process group (NCCL on the GPUs, gloo in the CPU tests).  The movie is
cut into contiguous frame ranges aligned to the 1024-frame chunks of the stats pass (pmd_loader.py:245), so the
two full-movie passes (mean/noise, projection) need no data-path communication at all.  What is exchanged:

  * all-reduce of the per-pixel mean / noise partial sums (2 x d floats) after the stats pass,
  * the background sample: raw frames held by their owner ranks -> every rank (ragged all-gather by per-rank broadcasts,
    moved as bytes so every movie dtype works),
  * the init window, PATCH PARTITIONED: a rank fits a contiguous range of blocks and receives from the window's owner
    rank(s) only the image rows those blocks cover (its rows plus half a block of halo) -- point-to-point sends of one
    slab per (owner, receiver) pair instead of N broadcasts of the whole window,
  * the per-block results (kept spatial components + their temporal traces): blocks are partitioned over the
    ranks, results all-gathered (ragged) so that every rank holds the same sparse U,
  * all-reduce of the k x k Gram of the projected movie before the final eigendecomposition,
  * all-gather of the Vt column shards into the result object.

Everything here is device agnostic (it only moves tensors), which is what the gloo tests exercise."""
import math
from typing import List, Optional, Sequence, Tuple

import torch


def dist_info(group) -> Tuple[int, int]:
    """(rank, world size) within `group`; (0, 1) when group is None."""
    if group is None:
        return 0, 1
    import torch.distributed as dist

    return dist.get_rank(group), dist.get_world_size(group)


def shard_bounds(n_frames: int, world: int, align: int = 1024) -> List[Tuple[int, int]]:
    """Contiguous frame ranges [lo, hi) per rank, boundaries on multiples of `align` (chunks as equal as possible);
    trailing ranks may be empty when the movie has fewer chunks than ranks."""
    chunks = math.ceil(n_frames / align)
    out = []
    for r in range(world):
        c0, c1 = (chunks * r) // world, (chunks * (r + 1)) // world
        out.append((min(c0 * align, n_frames), min(c1 * align, n_frames)))
    return out


def block_partition(n_blocks: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous block index ranges [b0, b1) per rank."""
    return [((n_blocks * r) // world, (n_blocks * (r + 1)) // world) for r in range(world)]


def owners_of(frames: Sequence[int], bounds: Sequence[Tuple[int, int]]) -> List[int]:
    """Owner rank of every requested global frame id."""
    out = []
    for f in frames:
        for r, (lo, hi) in enumerate(bounds):
            if lo <= f < hi:
                out.append(r)
                break
        else:
            raise IndexError("frame %d lies outside the movie" % f)
    return out


def ragged_all_gather(local: torch.Tensor, counts: Sequence[int], group) -> torch.Tensor:
    """Concatenate along dim 0 the per-rank tensors `local` (rank r contributes counts[r] rows; trailing dims and dtype
    equal on all ranks) and return the result on every rank.  Implemented as one broadcast per non-empty rank over a
    byte view, so any dtype (uint16 included) and any backend works."""
    import torch.distributed as dist

    rank, world = dist_info(group)
    assert local.shape[0] == counts[rank], (local.shape, counts, rank)
    total = int(sum(counts))
    out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    off = 0
    for r in range(world):
        n = int(counts[r])
        if n:
            seg = out[off : off + n]
            if r == rank:
                seg.copy_(local)
            dist.broadcast(seg.view(torch.uint8) if seg.dtype not in (torch.float32, torch.float64, torch.int32, torch.int64)
                           else seg, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
        off += n
    return out


def gather_frames(movie, frame_ids: Sequence[int], group, bounds: Optional[Sequence[Tuple[int, int]]] = None) -> torch.Tensor:
    """(n, d) tensor (native movie dtype) of arbitrary global frames on EVERY rank: each rank reads the frames it owns
    from its shard, the rows are exchanged, and the requested order is restored."""
    ids = [int(f) for f in frame_ids]
    if group is None:
        return movie.gather(ids)
    rank, world = dist_info(group)
    bounds = bounds if bounds is not None else shard_bounds(movie.T_total, world)
    own = owners_of(ids, bounds)
    order = sorted(range(len(ids)), key=lambda i: (own[i], i))  # grouped by owner, original order inside
    counts = [sum(1 for o in own if o == r) for r in range(world)]
    mine = [ids[i] for i in order if own[i] == rank]
    if mine:
        local = movie.gather(mine)
    else:
        local = torch.empty((0, movie.d), dtype=movie.torch_dtype, device=movie.device)
    gathered = ragged_all_gather(local.contiguous(), counts, group)
    inv = torch.empty(len(ids), dtype=torch.int64)
    inv[torch.tensor(order, dtype=torch.int64)] = torch.arange(len(ids), dtype=torch.int64)
    if len(ids) and order != list(range(len(ids))):
        gathered = gathered.index_select(0, inv.to(gathered.device)) if gathered.dtype != torch.uint16 else \
            gathered.view(torch.int16).index_select(0, inv.to(gathered.device)).view(torch.uint16)
    return gathered


def all_reduce_sum(t: torch.Tensor, group) -> torch.Tensor:
    if group is not None:
        import torch.distributed as dist

        dist.all_reduce(t, group=group)
    return t


def block_row_ranges(starts, bh: int, parts: Sequence[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """Image-row range [r_lo, r_hi) covered by the blocks starts[b0:b1] of every rank (blocks in row-major order of their
    starts, as the reference enumerates them); (0, 0) for a rank without blocks."""
    out = []
    for b0, b1 in parts:
        if b1 <= b0:
            out.append((0, 0))
        else:
            out.append((int(starts[b0][0]), int(starts[b1 - 1][0]) + int(bh)))
    return out


def owned_row_ranges(ranges: Sequence[Tuple[int, int]], d1: int) -> List[Tuple[int, int]]:
    """A disjoint cover of the image rows [0, d1) by the ranks: rank k owns the rows from the start of its range up to
    the start of the next non-empty rank's range (sums over the field of view, e.g. the background traces, count every
    pixel exactly once)."""
    los = [lo for lo, hi in ranges if hi > lo]
    out, k = [], 0
    for lo, hi in ranges:
        if hi <= lo:
            out.append((0, 0))
            continue
        nxt = los[k + 1] if k + 1 < len(los) else d1
        out.append((0 if k == 0 else lo, max(nxt, lo) if k + 1 < len(los) else d1))
        k += 1
    return out


def exchange_frame_rows(movie, frame_ids: Sequence[int], row_ranges: Sequence[Tuple[int, int]], d2: int, group,
                        bounds: Optional[Sequence[Tuple[int, int]]] = None) -> torch.Tensor:
    """Patch-partitioned gather: rank k gets, for ALL requested global frames (in the requested order), the pixels of the
    image rows row_ranges[k] = [r_lo, r_hi) -> (n_frames, (r_hi - r_lo) * d2) in the movie's dtype.  Every owner rank
    reads the requested frames of its shard once and sends one contiguous slab to every receiver (point to point)."""
    import torch.distributed as dist

    ids = [int(f) for f in frame_ids]
    rank, world = dist_info(group)
    bounds = bounds if bounds is not None else shard_bounds(movie.T_total, world)
    own = owners_of(ids, bounds)
    order = sorted(range(len(ids)), key=lambda i: (own[i], i))
    counts = [sum(1 for o in own if o == r) for r in range(world)]
    mine = [ids[i] for i in order if own[i] == rank]
    lo, hi = row_ranges[rank]
    n_pix = (hi - lo) * d2
    out = torch.empty((len(ids), n_pix), dtype=movie.torch_dtype, device=movie.device)
    as_bytes = movie.torch_dtype not in (torch.float32, torch.float64, torch.int32, torch.int64)
    view = (lambda t: t.view(torch.uint8)) if as_bytes else (lambda t: t)
    ops, keep = [], []
    local = None                                                     # (my frames, d)
    if mine:
        src = None
        if mine == list(range(mine[0], mine[0] + len(mine))) and hasattr(movie, "frame_source"):
            src, idx = movie.frame_source(mine)                      # resident shard: a view, no copy of the frames
            i0 = int(idx[0].item()) if idx.numel() else 0
            local = src[i0 : i0 + len(mine)] if src.shape[0] >= i0 + len(mine) and bool((idx == torch.arange(
                i0, i0 + len(mine), device=idx.device)).all()) else None
        if local is None:
            local = movie.gather(mine)
    offs = [sum(counts[:r]) for r in range(world)]
    for k in range(world):                                           # my frames -> every receiver's rows
        klo, khi = row_ranges[k]
        if local is None or khi <= klo:
            continue
        slab = local[:, klo * d2 : khi * d2]
        if k == rank:
            out[offs[rank] : offs[rank] + counts[rank]].copy_(slab)
        else:
            slab = slab.contiguous()
            keep.append(slab)
            ops.append(dist.P2POp(dist.isend, view(slab), dist.get_global_rank(group, k) if group is not None else k, group))
    if n_pix > 0:
        for o in range(world):                                       # every owner's frames of my rows
            if o != rank and counts[o]:
                seg = out[offs[o] : offs[o] + counts[o]]
                ops.append(dist.P2POp(dist.irecv, view(seg), dist.get_global_rank(group, o) if group is not None else o, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if len(ids) and order != list(range(len(ids))):                  # rows are grouped by owner: restore the requested order
        inv = torch.empty(len(ids), dtype=torch.int64)
        inv[torch.tensor(order, dtype=torch.int64)] = torch.arange(len(ids), dtype=torch.int64)
        inv = inv.to(out.device)
        out = out.index_select(0, inv) if out.dtype != torch.uint16 else out.view(torch.int16).index_select(0, inv).view(torch.uint16)
    return out
