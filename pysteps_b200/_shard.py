"""Multi-GPU plumbing of the advection path: one process per GPU (torch.distributed),
independent fields sharded round-robin over ranks, ONE broadcast of the motion field
(ensemble members of nowcasts.steps each have their own precipitation
field and displacement; nothing else is exchanged)."""
import torch
import torch.distributed as dist


def member_indices(n_members, world_size, rank):
    """Members owned by `rank`: i with i % world_size == rank (3 per GPU for 24 members on 8)."""
    return list(range(rank, n_members, world_size))


def row_band(m, world_size, rank):
    """Rows [r0, r1) of an m-row composite owned by `rank` (even split, remainder to the first
    ranks).  Every output pixel depends on the inputs only, so bands need no halo exchange."""
    base, rem = divmod(m, world_size)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def broadcast_field(field, src=0, shape=None, dtype=torch.float64, device=None):
    """Broadcast the (2,m,n) motion field computed on `src` to every rank.  Ranks other than
    `src` may pass None (a buffer of `shape` is allocated).  No-op without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return field
    if field is None:
        if shape is None:
            raise ValueError("shape is required on ranks that do not hold the field")
        field = torch.empty(shape, dtype=dtype, device=device)
    dist.broadcast(field, src=src)
    return field


def max_over_ranks(value, device=None):
    """Max of a Python float over ranks (device timing is reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_row_bands(band, m, world_size, rank):
    """Full (c, m, n) field from every rank's (c, rows, n) band of it (bands as `row_band`
    assigns them).  The one exchange step of the tile-partitioned path: the trajectory kernel
    samples the motion field anywhere, so every rank needs all of it."""
    if world_size == 1:
        return band
    c, rows, n = band.shape
    full = torch.empty((c, m, n), dtype=band.dtype, device=band.device)
    bands = [row_band(m, world_size, r) for r in range(world_size)]
    assert bands[rank][1] - bands[rank][0] == rows
    even = all(b[1] - b[0] == rows for b in bands)
    for ch in range(c):
        views = [full[ch, b[0]:b[1]] for b in bands]
        if even and dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(full[ch], band[ch].contiguous())  # bands are consecutive
        elif even:
            dist.all_gather(views, band[ch].contiguous())
        else:
            for r, v in enumerate(views):
                if r == rank:
                    v.copy_(band[ch])
                dist.broadcast(v, src=r)
    return full
