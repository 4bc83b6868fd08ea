"""Batched tile inference -- the fast path for what `ssr/infer.py:45-67` and `ssr/infer_grid.py:46-85` do one 32x32 chunk at
a time (batch 1, autograd graph recorded, PNG round trip per chunk): all chunks of a tile go through the generator in a few
large batches and are clamped / converted / stitched into the 2048^2 canvas on the GPU.
"""
import random

import numpy as np
import torch

from . import _lib as L
from .ops import cur_stream, lib


def format_s2naip_data(s2_data, n_s2_images, device=None, rng=None):
    """ssr/utils/infer_utils.py:6-39: [T*32, 32, 3] uint8 -> ([1, n*3, 32, 32] float in [0,1], first frame).  Frames that
    contain a pure-black pixel are used only when there are not enough clean ones; the choice is a `random.sample` (pass
    `rng=random.Random(seed)` for reproducibility -- the reference uses the unseeded global one)."""
    rng = rng or random
    chunks = np.reshape(s2_data, (-1, 32, 32, 3))
    first = chunks[0]
    goods, bads = [], []
    for i, ts in enumerate(chunks):
        (bads if [0, 0, 0] in ts else goods).append(i)      # same (quirky) membership test as the reference
    if len(goods) >= n_s2_images:
        idx = rng.sample(goods, n_s2_images)
    else:
        idx = goods + rng.sample(bads, n_s2_images - len(goods))
    sel = np.array([chunks[i] for i in idx])
    t = torch.cat([torch.as_tensor(img).permute(2, 0, 1) for img in sel]).unsqueeze(0)
    if device is not None:
        t = t.to(device)
    return t.float() / 255, first


@torch.no_grad()
def super_resolve(net_g, lr_u8, batch=256, canvas=None, grid_cols=None):
    """lr_u8: uint8 [N, T*C, h, w] (host or device).  Returns uint8 [N, H, W, 3] chunks, or -- with `canvas` -- pastes chunk i at
    tile (i // grid_cols, i % grid_cols) of the uint8 [rows, cols, 3] canvas (the `stitch` layout)."""
    eng = net_g._get_engine() if hasattr(net_g, "_get_engine") else net_g
    if hasattr(net_g, "_weights_dirty") and net_g._weights_dirty():
        eng.repack()
    dev = eng.device
    N, Cc, h, w = lr_u8.shape
    s = cur_stream()
    H, W = h * eng.scale, w * eng.scale
    out = None
    if canvas is None:
        out = torch.empty((N, H, W, 3), dtype=torch.uint8, device=dev)
    lb = lib()
    for i0 in range(0, N, batch):
        nb = min(batch, N - i0)
        x8 = lr_u8[i0:i0 + nb].to(dev, non_blocking=True).contiguous()
        ws = eng.workspace(nb, h, w, False)
        L.check(lb.ssr_ingest_nchw(x8.data_ptr(), 0, ws.in0.ptr(), ws.in0.stride, nb, Cc, h, w, eng.cin_pad, 1.0 / 255.0, None,
                                   None, s))
        ws.fwd.run(s)
        if canvas is None:
            L.check(lb.ssr_f32_nchw_to_u8_canvas(ws.out.data_ptr(), out.data_ptr() + i0 * H * W * 3, nb, 3, H, W, W, 1, 0, s))
        else:
            L.check(lb.ssr_f32_nchw_to_u8_canvas(ws.out.data_ptr(), canvas.data_ptr(), nb, 3, H, W, canvas.shape[1], grid_cols, i0, s))
    return out if canvas is None else canvas


@torch.no_grad()
def infer_grid(net_g, lr_u8, grid_size=16, batch=256):
    """ssr/infer_grid.py for one tile: grid_size^2 chunks (row-major i_j order) -> uint8 [grid*H, grid*W, 3] stitched image."""
    N, _, h, w = lr_u8.shape
    assert N == grid_size * grid_size
    eng = net_g._get_engine() if hasattr(net_g, "_get_engine") else net_g
    H, W = h * eng.scale, w * eng.scale
    canvas = torch.empty((grid_size * H, grid_size * W, 3), dtype=torch.uint8, device=eng.device)
    return super_resolve(net_g, lr_u8, batch=batch, canvas=canvas, grid_cols=grid_size)


class TilePipeline:
    """Tiles of grid_size^2 chunks streamed through one GPU with the host work off the critical path: the uint8 chunks of tile
    t+1 are copied from pinned host memory (copy stream) while tile t runs, and the stitched uint8 canvas of tile t-1 drains
    to pinned host memory -- what `ssr/infer_grid.py:46-85` does with one PNG round trip per 32x32 chunk.  PNG decode / encode
    stay on host threads of the caller (out of scope: SURVEY.md section 2 rows 8, 12)."""

    def __init__(self, net_g, grid_size=16, chunk_hw=32, in_ch=None, batch=256, depth=2):
        eng = net_g._get_engine() if hasattr(net_g, "_get_engine") else net_g
        self.net_g, self.eng, self.grid, self.batch = net_g, eng, grid_size, batch
        n = grid_size * grid_size
        cin = in_ch if in_ch is not None else eng.cin
        side = grid_size * chunk_hw * eng.scale
        dev = eng.device
        self.host_in = [torch.empty((n, cin, chunk_hw, chunk_hw), dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.dev_in = [torch.empty((n, cin, chunk_hw, chunk_hw), dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.dev_out = [torch.empty((side, side, 3), dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.host_out = [torch.empty((side, side, 3), dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.copy_in, self.copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.depth = depth

    @torch.no_grad()
    def run(self, tiles, consume):
        """tiles: iterable of uint8 [grid^2, C, h, w] host tensors; consume(index, uint8 [S, S, 3] pinned host canvas) is called
        once the canvas has arrived (the buffer is reused `depth` tiles later)."""
        main = torch.cuda.current_stream()
        pending = []
        for i, t in enumerate(tiles):
            k = i % self.depth
            if i >= self.depth:                       # slot k: its previous canvas must have been consumed
                self.ev_out[k].synchronize()
                j = pending.pop(0)
                consume(j, self.host_out[j % self.depth])
            self.host_in[k].copy_(t)
            with torch.cuda.stream(self.copy_in):
                self.dev_in[k].copy_(self.host_in[k], non_blocking=True)
                self.ev_in[k].record()
            main.wait_event(self.ev_in[k])
            super_resolve(self.net_g, self.dev_in[k], batch=self.batch, canvas=self.dev_out[k], grid_cols=self.grid)
            self.ev_done[k].record(main)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(self.ev_done[k])
                self.host_out[k].copy_(self.dev_out[k], non_blocking=True)
                self.ev_out[k].record()
            pending.append(i)
        for j in pending:
            self.ev_out[j % self.depth].synchronize()
            consume(j, self.host_out[j % self.depth])


def infer_tiles_sharded(net_g, tiles, rank=0, world=1, consume=None, **kw):
    """Grid inference over many tiles on `world` GPUs: tiles are independent (ssr/infer_grid.py stitches each directory on its
    own), so rank r simply takes the contiguous slice ops.rank_slice(len(tiles), r, world) -- replicas, no collective.
    Returns {tile index: uint8 [S, S, 3] host tensor} for this rank's tiles unless `consume` takes them."""
    from .ops import rank_slice
    mine = list(rank_slice(len(tiles), rank, world))
    out = {}

    def sink(j, canvas):
        if consume is not None:
            consume(mine[j], canvas)
        else:
            out[mine[j]] = canvas.clone()

    if mine:
        t0 = tiles[mine[0]]
        TilePipeline(net_g, grid_size=int(round(t0.shape[0] ** 0.5)), chunk_hw=t0.shape[-1], in_ch=t0.shape[1], **kw).run(
            (tiles[i] for i in mine), sink)
    return out


# ------------------------------------------------------------------------------------------------ directory workflow
def _read_png_rgb(path):
    import cv2
    im = cv2.imread(path, cv2.IMREAD_COLOR)
    if im is None:
        raise FileNotFoundError(f"cannot decode {path}")
    return cv2.cvtColor(im, cv2.COLOR_BGR2RGB)


def _write_png_rgb(path, rgb):
    import cv2
    if not cv2.imwrite(path, cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR)):
        raise OSError(f"cannot write {path}")


def load_tile_dir(tile_dir, n_s2_images, grid_size=16, pool=None, rng=None, out=None):
    """One tile directory of `ssr/infer_grid.py` ({tile}/{i}_{j}.png, each [T*32, 32, 3] uint8) -> (uint8 [grid^2, n*3, 32, 32]
    chunk stack in i_j row-major order, uint8 [grid*32, grid*32, 3] stitched first frames = the reference's `stitched_s2.png`).
    Frame choice per chunk = infer_utils.format_s2naip_data (clean frames first, `random.sample`); decode runs on `pool` threads
    (cv2 releases the GIL); `out` may be a pinned buffer to decode into."""
    rng = rng or random
    names = [f"{i}_{j}.png" for i in range(grid_size) for j in range(grid_size)]
    paths = [f"{tile_dir}/{n}" for n in names]
    ims = list(pool.map(_read_png_rgb, paths)) if pool is not None else [_read_png_rgb(p) for p in paths]
    n = grid_size * grid_size
    stack = out if out is not None else torch.empty((n, n_s2_images * 3, 32, 32), dtype=torch.uint8)
    s2 = np.zeros((grid_size * 32, grid_size * 32, 3), dtype=np.uint8)
    dst = stack.numpy()
    for k, im in enumerate(ims):
        chunks = np.reshape(im, (-1, 32, 32, 3))
        goods, bads = [], []
        for t, ts in enumerate(chunks):
            (bads if [0, 0, 0] in ts else goods).append(t)          # same membership test as the reference (infer_utils.py:16-20)
        idx = rng.sample(goods, n_s2_images) if len(goods) >= n_s2_images else goods + rng.sample(bads, n_s2_images - len(goods))
        dst[k] = np.concatenate([chunks[t].transpose(2, 0, 1) for t in idx], 0)   # channel index = frame * 3 + rgb
        i, j = divmod(k, grid_size)
        s2[i * 32:(i + 1) * 32, j * 32:(j + 1) * 32] = chunks[0]
    return stack, s2


def infer_grid_dir(net_g, data_dir, save_path, n_s2_images=8, grid_size=16, threads=8, batch=256, rank=0, world=1, rng=None,
                   write_chunks=False):
    """`ssr/infer_grid.py:46-85` for a directory tree {data_dir}/{tile}/{i}_{j}.png: every complete tile (grid_size^2 chunks) is
    decoded on a thread pool into pinned memory, super-resolved as one batched pass with the stitching done on the GPU
    (TilePipeline: H2D of tile t+1 and D2H of tile t-1 overlap tile t), and written as {save_path}/{tile}/stitched_sr.png next to
    stitched_s2.png -- the two files the reference produces -- by the same pool (PNG encode of tile t-1 overlaps tile t).  With
    write_chunks the grid_size^2 per-chunk PNGs {i}_{j}.png are written too (the reference always writes them; they are only the
    input of its own stitch step).  Tiles are independent: rank r of `world` takes ops.rank_slice of the sorted tile list."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    from .ops import rank_slice
    need = grid_size * grid_size
    tiles = sorted(t for t in os.listdir(data_dir) if os.path.isdir(os.path.join(data_dir, t)))
    complete = [t for t in tiles if len([f for f in os.listdir(os.path.join(data_dir, t)) if f.endswith(".png")]) >= need]
    skipped = [t for t in tiles if t not in complete]       # "contains less than 256 chunks, cannot stitch" (infer_grid.py:72-74)
    mine = [complete[i] for i in rank_slice(len(complete), rank, world)]
    if not mine:
        return dict(tiles=[], skipped=skipped)
    eng = net_g._get_engine() if hasattr(net_g, "_get_engine") else net_g
    side = 32 * eng.scale
    with ThreadPoolExecutor(max_workers=threads) as pool:
        pending = []

        def source():
            for t in mine:
                stack, s2 = load_tile_dir(os.path.join(data_dir, t), n_s2_images, grid_size, pool, rng)
                os.makedirs(os.path.join(save_path, t), exist_ok=True)
                pending.append(pool.submit(_write_png_rgb, os.path.join(save_path, t, "stitched_s2.png"), s2))
                yield stack

        def sink(j, canvas):
            img = canvas.numpy().copy()                       # the pinned buffer is reused two tiles later
            out_dir = os.path.join(save_path, mine[j])
            pending.append(pool.submit(_write_png_rgb, os.path.join(out_dir, "stitched_sr.png"), img))
            if write_chunks:
                for k in range(need):
                    i, jj = divmod(k, grid_size)
                    pending.append(pool.submit(_write_png_rgb, os.path.join(out_dir, f"{i}_{jj}.png"),
                                               img[i * side:(i + 1) * side, jj * side:(jj + 1) * side]))

        TilePipeline(net_g, grid_size=grid_size, chunk_hw=32, in_ch=n_s2_images * 3, batch=batch).run(source(), sink)
        for f in pending:
            f.result()
    return dict(tiles=mine, skipped=skipped)
