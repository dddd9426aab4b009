"""Multi-GPU all-pairs Chamfer matrix: 2-D tiling of the [N_sample, N_ref] pair grid over the ranks.

SURVEY.md section 8(e): the reference has no multi-GPU path for evaluation; the pair grid is embarrassingly
parallel, so each rank (one process per GPU, torch.distributed) computes one [rows x cols] tile with the
single-GPU kernel and only the per-pair scalars are exchanged: one all_gather of 4*N_s*N_r/P bytes per rank
(0.5 MB at 1000x1000 on 8 GPUs) over NCCL/NVLink.  No collective touches the point data: both cloud sets are
already resident on every rank (the caller passes the same tensors everywhere, as compute_all_metrics does).
"""
import torch
import torch.distributed as dist_

from . import ops


def is_distributed():
    return dist_.is_available() and dist_.is_initialized() and dist_.get_world_size() > 1


def rank_grid(world):
    """(pr, pc) with pr*pc == world and pr <= pc as square as possible: 1->1x1, 2->1x2, 4->2x2, 8->2x4."""
    pr = int(world ** 0.5)
    while world % pr:
        pr -= 1
    return pr, world // pr


def split(n, parts, i):
    """[lo, hi) of part i when n items are cut into `parts` near-equal contiguous ranges."""
    base, rem = divmod(n, parts)
    lo = i * base + min(i, rem)
    return lo, lo + base + (1 if i < rem else 0)


def tile_of(rank, world, n_rows, n_cols):
    """Row/column ranges owned by `rank`: ((r0, r1), (c0, c1))."""
    pr, pc = rank_grid(world)
    return split(n_rows, pr, rank // pc), split(n_cols, pc, rank % pc)


def assemble(tiles, world, n_rows, n_cols, like=None):
    """Inverse of tile_of: the per-rank padded tiles ([world, max_r, max_c], or a list of [max_r, max_c]) -> full matrix.
    One strided copy when the grid divides the matrix evenly; otherwise two index_selects drop the padding."""
    if isinstance(tiles, (list, tuple)):
        tiles = torch.stack(list(tiles))
    pr, pc = rank_grid(world)
    mr, mc = tiles.shape[1], tiles.shape[2]
    full = tiles.view(pr, pc, mr, mc).permute(0, 2, 1, 3).reshape(pr * mr, pc * mc)
    if pr * mr != n_rows:
        keep = torch.cat([torch.arange(*split(n_rows, pr, i)) - split(n_rows, pr, i)[0] + i * mr for i in range(pr)])
        full = full.index_select(0, keep.to(full.device))
    if pc * mc != n_cols:
        keep = torch.cat([torch.arange(*split(n_cols, pc, i)) - split(n_cols, pc, i)[0] + i * mc for i in range(pc)])
        full = full.index_select(1, keep.to(full.device))
    return full


def max_tile(world, n_rows, n_cols):
    pr, pc = rank_grid(world)
    return -(-n_rows // pr), -(-n_cols // pc)


def sym_plan(world, n, blocks_per_rank=4):
    """Same-set matrix (rr / ss of compute_all_metrics): CD(a, b) == CD(b, a), so only the block pairs (i <= j) of a
    T x T block grid are computed, T = blocks_per_rank * world.  Returns (bounds, owner) with bounds[i] = (lo, hi) of block i
    and owner = list over ranks of their [(i, j), ...] tiles, assigned largest-first to the least loaded rank (a diagonal
    tile costs half: the kernel walks its upper triangle only).  Deterministic: every rank derives the same plan."""
    t = max(1, min(n, blocks_per_rank * world))
    bounds = [split(n, t, i) for i in range(t)]
    size = [hi - lo for lo, hi in bounds]
    tiles = [(i, j) for i in range(t) for j in range(i, t)]
    cost = {(i, j): size[i] * size[j] * (0.5 if i == j else 1.0) for i, j in tiles}
    tiles.sort(key=lambda ij: (-cost[ij], ij))
    load = [0.0] * world
    owner = [[] for _ in range(world)]
    for ij in tiles:
        r = min(range(world), key=lambda q: (load[q], q))
        owner[r].append(ij)
        load[r] += cost[ij]
    return bounds, owner


_SYM_MAPS = {}


def _sym_maps(world, n, device):
    """(bounds, owner, longest, src, dst, dst_mirror): flat gather/scatter indices that paste the all_gathered tile buffers of
    sym_plan into the full [n, n] matrix (and its mirror image) with two index_puts instead of one copy per tile."""
    key = (world, n, str(device))
    if key not in _SYM_MAPS:
        import numpy as np
        bounds, owner = sym_plan(world, n)
        area = lambda ij: (bounds[ij[0]][1] - bounds[ij[0]][0]) * (bounds[ij[1]][1] - bounds[ij[1]][0])
        longest = max(1, max(sum(area(ij) for ij in tiles) for tiles in owner))
        src, dst, msrc, mdst = [], [], [], []
        for r in range(world):
            at = r * longest
            for ij in owner[r]:
                (r0, r1), (c0, c1) = bounds[ij[0]], bounds[ij[1]]
                rr, cc = np.meshgrid(np.arange(r0, r1), np.arange(c0, c1), indexing="ij")
                pos = at + np.arange(rr.size)
                src.append(pos); dst.append((rr * n + cc).reshape(-1))
                if ij[0] != ij[1]:
                    msrc.append(pos); mdst.append((cc * n + rr).reshape(-1))
                at += rr.size
        cat = lambda parts: torch.from_numpy(np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)).to(device)
        _SYM_MAPS[key] = (bounds, owner, longest, cat(src), cat(dst), cat(msrc), cat(mdst))
    return _SYM_MAPS[key]


def pairwise_cd_symmetric(pcs, group=None, compute_tile=None):
    """Full [N, N] Chamfer matrix of a cloud set against itself on every rank, computing every unordered pair once."""
    world = dist_.get_world_size(group)
    rank = dist_.get_rank(group)
    n = pcs.shape[0]
    if compute_tile is None:
        compute_tile = ops.cd_allpairs
    bounds, owner, longest, src, dst, msrc, mdst = _sym_maps(world, n, pcs.device)
    area = lambda ij: (bounds[ij[0]][1] - bounds[ij[0]][0]) * (bounds[ij[1]][1] - bounds[ij[1]][0])
    mine = pcs.new_zeros((longest,))
    at = 0
    for ij in owner[rank]:
        rows, cols = bounds[ij[0]], bounds[ij[1]]
        if rows[1] > rows[0] and cols[1] > cols[0]:
            mine[at: at + area(ij)] = compute_tile(pcs, pcs, rows, cols).reshape(-1)
        at += area(ij)
    gathered = mine.new_empty((world * longest,))
    dist_.all_gather_into_tensor(gathered, mine, group=group)
    full = mine.new_empty((n * n,))
    full[dst] = gathered[src]
    full[mdst] = gathered[msrc]
    return full.view(n, n)


def _same_set(a, b):
    return a.shape == b.shape and a.data_ptr() == b.data_ptr() and a.stride() == b.stride()


def _device_of(t):
    return t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())


def pairwise_emd(sample_pcs, ref_pcs, group=None):
    """Full [N_sample, N_ref] approximate-EMD matrix on every rank (same tiling and collective as pairwise_cd)."""
    return pairwise_cd(sample_pcs, ref_pcs, group=group, compute_tile=ops.emd_allpairs)


def pairwise_cd(sample_pcs, ref_pcs, group=None, compute_tile=None):
    """Full [N_sample, N_ref] Chamfer matrix on every rank.  `compute_tile(sample, ref, rows, cols)` defaults to
    the CUDA kernel; tests substitute a CPU function to exercise the tiling + collective with gloo.

    HOST inputs (the end-to-end path: compute_all_metrics is handed CPU tensors by a loader) are not replicated: each
    rank copies only the rows and columns of its own tile to its GPU (750 of 2000 clouds at 8 ranks)."""
    world = dist_.get_world_size(group)
    rank = dist_.get_rank(group)
    n_rows, n_cols = sample_pcs.shape[0], ref_pcs.shape[0]
    host = compute_tile is None and not sample_pcs.is_cuda
    kernel = ops.cd_allpairs if compute_tile is None else compute_tile
    on_gpu = compute_tile is None or compute_tile is ops.emd_allpairs
    if compute_tile is None and _same_set(sample_pcs, ref_pcs) and n_rows >= 128:
        # rr / ss matrices: every unordered pair once (CD is symmetric); below ~128 clouds the extra launches cost more
        dev_set = sample_pcs.to(_device_of(sample_pcs), non_blocking=True) if host else sample_pcs
        return pairwise_cd_symmetric(dev_set, group=group)
    rows, cols = tile_of(rank, world, n_rows, n_cols)
    mr, mc = max_tile(world, n_rows, n_cols)
    nr, nc = rows[1] - rows[0], cols[1] - cols[0]
    if on_gpu and not sample_pcs.is_cuda:
        dev = _device_of(sample_pcs)
        a = sample_pcs[rows[0]:rows[1]].to(dev, non_blocking=True)
        b = ref_pcs[cols[0]:cols[1]].to(dev, non_blocking=True)
        mine = a.new_zeros((mr, mc))
        if nr > 0 and nc > 0:
            kernel(a, b, (0, nr), (0, nc), out=mine[:nr, :nc])
    else:
        mine = sample_pcs.new_zeros((mr, mc))
        if nr > 0 and nc > 0:
            if on_gpu:
                kernel(sample_pcs, ref_pcs, rows, cols, out=mine[:nr, :nc])     # written in place, row stride mc
            else:
                mine[:nr, :nc] = kernel(sample_pcs, ref_pcs, rows, cols)
    gathered = mine.new_empty((world * mr, mc))                 # ranks concatenated along dim 0 (the layout gloo accepts too)
    dist_.all_gather_into_tensor(gathered, mine, group=group)
    return assemble(gathered.view(world, mr, mc), world, n_rows, n_cols)
