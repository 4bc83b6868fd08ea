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
