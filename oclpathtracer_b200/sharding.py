"""Image sharding across ranks (BUILD-DEFINED; the reference is single-device).

Pixels are independent and the RNG seed depends only on (global pixel id, frame)
(GenerateColors.cl:305-308), so the image is cut into blocks of `block`
consecutive gids dealt round-robin to the ranks: rank r renders the gids with
(gid // block) % world == r and stores them compacted (ptb_render_params.shard_*).
Results are bit-identical to a single-device render.  One exchange at the end:
all_gather of the equal-sized local frames, then `assemble` un-interleaves.
"""
import torch


def local_pixels(n_pixels, rank, world, block):
    if world <= 1:
        return n_pixels
    c = 0
    b = rank
    while b * block < n_pixels:
        c += min(n_pixels, (b + 1) * block) - b * block
        b += world
    return c


def local_to_gid(n_local, rank, world, block, device=None):
    li = torch.arange(n_local, dtype=torch.int64, device=device)
    if world <= 1:
        return li
    return ((li // block) * world + rank) * block + li % block


def assemble(parts, n_pixels, world, block):
    """parts: list (len world) of [n_local_r, C] tensors -> [n_pixels, C] image in gid order."""
    if world <= 1:
        return parts[0]
    c = parts[0].shape[1]
    if n_pixels % (block * world) == 0 and all(p.shape[0] == parts[0].shape[0] for p in parts):
        g = torch.stack(list(parts), 0)  # [world, n_local, C]
        return g.view(world, -1, block, c).permute(1, 0, 2, 3).reshape(n_pixels, c)
    out = torch.empty((n_pixels, c), dtype=parts[0].dtype, device=parts[0].device)
    for r, p in enumerate(parts):
        out[local_to_gid(p.shape[0], r, world, block, device=p.device)] = p
    return out


def gather_image(local, n_pixels, rank, world, block, group=None):
    """all_gather the local frames of every rank and return the assembled image (on every rank)."""
    import torch.distributed as dist

    if world <= 1:
        return local
    n_max = local_pixels(n_pixels, 0, world, block)  # rank 0 owns the most
    if local.shape[0] < n_max:
        pad = torch.zeros((n_max - local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    bufs = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(bufs, local.contiguous(), group=group)
    parts = [bufs[r][: local_pixels(n_pixels, r, world, block)] for r in range(world)]
    return assemble(parts, n_pixels, world, block)
