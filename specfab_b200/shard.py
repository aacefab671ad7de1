"""Multi-GPU sharding of a node field: contiguous node ranges, no collective while stepping.

Every node's ODE, moments and Eij depend only on that node's state and forcing (reference:
src/specfabpy.f90:483-485 loops nodes independently), so rank r of G owns the
contiguous range [lo, hi) of the node-contiguous arrays and steps it with no data-path exchange.
The only collective is the optional final gather of a (small) per-node output.
"""


def node_range(N, rank, world):
    """contiguous range [lo, hi) of rank `rank` out of `world` for N nodes (sizes differ by <= 1)"""
    if world < 1 or not (0 <= rank < world) or N < 0:
        raise ValueError("bad partition arguments")
    base, rem = divmod(N, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def gather_rows(local, N, group=None):
    """All-gather a per-node output.  local: tensor (..., n_local) whose LAST dim is this rank's node
    range (library layout: node contiguous).  Returns the (..., N) tensor on every rank.
    Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [node_range(N, r, world)[1] - node_range(N, r, world)[0] for r in range(world)]
    if local.shape[-1] != sizes[rank]:
        raise ValueError("local node count %d does not match the partition (%d)" % (local.shape[-1], sizes[rank]))
    nmax = max(sizes)
    lead = tuple(local.shape[:-1])
    pad = torch.zeros(lead + (nmax,), dtype=local.dtype, device=local.device)
    pad[..., : sizes[rank]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[..., :s] for b, s in zip(bufs, sizes)], dim=-1)
