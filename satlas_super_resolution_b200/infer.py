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
