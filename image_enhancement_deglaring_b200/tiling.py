"""High-resolution images as independent 512x512 tiles (BASELINE.json configs[2], SURVEY 8e "definition A").

The reference has no whole-image path for a 4096x4096 input: its model is trained and served at 512x512
(api/app.py:149 resizes, src/preprocess.py crops), so what a user of the repo can do today is run the network on every
512x512 tile.  GroupNorm statistics are per sample, so tiles are independent units: they are one batch, sharded over
ranks with no data-path collective (parallel.shard_range); only the finished tiles are gathered, and only if asked.
This is NOT the whole-image function (GroupNorm + receptive field couple tiles): that is SURVEY 8e definition B, whole_image.py.
"""
import torch

from .parallel import shard_range


def split_tiles(image, tile=512):
    """image [H, W] or [C, H, W] (H, W multiples of `tile`) -> tiles [T, C, tile, tile], row-major over the tile grid,
    plus the grid (rows, cols).  A view-free reshape: no padding, no overlap."""
    if image.dim() == 2:
        image = image[None]
    C, H, W = image.shape
    if H % tile or W % tile:
        raise RuntimeError(f"image {H}x{W} is not a whole number of {tile}x{tile} tiles")
    gh, gw = H // tile, W // tile
    t = image.reshape(C, gh, tile, gw, tile).permute(1, 3, 0, 2, 4).reshape(gh * gw, C, tile, tile)
    return t.contiguous(), (gh, gw)


def merge_tiles(tiles, grid):
    """Inverse of split_tiles: [T, C, tile, tile] -> [C, H, W]."""
    gh, gw = grid
    T, C, th, tw = tiles.shape
    if T != gh * gw:
        raise RuntimeError(f"{T} tiles for a {gh}x{gw} grid")
    return tiles.reshape(gh, gw, C, th, tw).permute(2, 0, 3, 1, 4).reshape(C, gh * th, gw * tw).contiguous()


def infer_tiled(forward, image, tile=512, rank=0, world=1, group=None, gather=True, batch=64, out_channels=None, out_dtype=None):
    """Run `forward` (a callable [n, C, tile, tile] -> [n, C', tile, tile]: the drop-in module, `forward_u8`, ...) over the
    tiles of `image` that belong to `rank` (contiguous shard of the row-major tile list), `batch` tiles per call.

    gather=True: every rank returns the full [C', H, W] result (one all_gather of the finished tiles over `group`).
    gather=False: returns (tiles of this rank, (first, last) tile index, grid) and never communicates.
    out_channels / out_dtype: shape of `forward`'s output per tile if it differs from the input's (C' != C, a float callable fed a
    uint8 image, ...); only needed by a rank that owns NO tile (more ranks than tiles) -- every rank must hand all_gather the same
    shape and dtype.  When omitted such a rank runs `forward` on one zero tile to learn them."""
    tiles, grid = split_tiles(image, tile)
    lo, hi = shard_range(tiles.shape[0], rank, world)
    outs = [forward(tiles[i:min(i + batch, hi)]) for i in range(lo, hi, batch)]
    if outs:
        mine = torch.cat(outs, 0)
    else:  # more ranks than tiles: an empty shard with the OUTPUT's channel count and dtype (not the input's)
        if out_channels is None or out_dtype is None:
            probe = forward(torch.zeros_like(tiles[:1]))
            out_channels, out_dtype = probe.shape[1], probe.dtype
        mine = torch.zeros((0, out_channels) + tuple(tiles.shape[2:]), dtype=out_dtype, device=tiles.device)
    if not gather:
        return mine, (lo, hi), grid
    if world > 1:
        import torch.distributed as dist
        per = [shard_range(tiles.shape[0], r, world) for r in range(world)]
        width = max(b - a for a, b in per)
        pad = mine.new_zeros((width,) + tuple(mine.shape[1:]))
        pad[:mine.shape[0]] = mine
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        mine = torch.cat([p[:b - a] for p, (a, b) in zip(parts, per)], 0)
    return merge_tiles(mine, grid)
