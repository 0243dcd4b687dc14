"""Multi-GPU: registration pairs are independent (no cross-sample op in eval mode), so the batch is
sharded contiguously over ranks, weights are replicated, and the ONLY exchange is one all_gather of the poses
([B_local, 12] fp32: R row-major + t) over NCCL / NVLink (gloo on CPU for tests)."""
import torch
import torch.distributed as dist


def shard_range(n_pairs: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of rank `rank`; the first n_pairs % world ranks get one extra pair."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_pose(R, t):
    return torch.cat([R.reshape(R.shape[0], 9), t.reshape(t.shape[0], 3)], dim=1).contiguous()


def unpack_pose(p):
    return p[:, :9].reshape(-1, 3, 3), p[:, 9:12]


def gather_poses(R, t, n_pairs=None, group=None):
    """all_gather of the local poses -> (R [B_total,3,3], t [B_total,3]) in pair order on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return R, t
    world = dist.get_world_size(group)
    # the pose kernel of the model path leaves the packed rows next to R (engine.weighted_kabsch(packed=True))
    local = getattr(R, "hrn_pose12", None)
    if local is None:
        local = pack_pose(R, t)
    if n_pairs is None or n_pairs % world == 0:
        out = torch.empty(world * local.shape[0], 12, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return unpack_pose(out)
    sizes = [shard_range(n_pairs, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros(mx, 12, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    out = torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
    return unpack_pose(out)
