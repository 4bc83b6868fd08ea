"""GPU parity of the batched inference path (super_resolve / infer_grid) against the oracle forward + the reference's
clamp / *255 / astype(uint8) / stitch arithmetic (ssr/infer.py:61-64, ssr/utils/infer_utils.py:41-60)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_super_resolve_and_stitch():
    from oracle import nets
    from satlas_super_resolution_b200.archs import SSR_RRDBNet
    from satlas_super_resolution_b200.infer import infer_grid, super_resolve
    nb, grid = 2, 4
    sd = nets.rrdbnet_init(24, 3, num_block=nb, seed=3)
    net = SSR_RRDBNet(24, 3, num_block=nb)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    g = torch.Generator().manual_seed(4)
    lr = torch.randint(0, 256, (grid * grid, 24, 32, 32), generator=g, dtype=torch.uint8)
    with torch.no_grad():
        ref = nets.rrdbnet_forward(sd, lr.float() / 255, num_block=nb)
    ref_u8 = (torch.clamp(ref, 0, 1) * 255).permute(0, 2, 3, 1).numpy().astype(np.uint8)     # [N,128,128,3]
    got = super_resolve(net, lr, batch=5).cpu().numpy()          # ragged batches: 5,5,5,1
    assert got.shape == ref_u8.shape
    diff = np.abs(got.astype(np.int32) - ref_u8.astype(np.int32))
    assert diff.max() <= 6 and (diff > 2).mean() < 0.01, (diff.max(), (diff > 2).mean())
    canvas = infer_grid(net, lr, grid_size=grid, batch=16).cpu().numpy()
    assert canvas.shape == (grid * 128, grid * 128, 3)
    for i in range(grid):
        for j in range(grid):
            tile = canvas[i * 128:(i + 1) * 128, j * 128:(j + 1) * 128]
            assert np.array_equal(tile, got[i * grid + j]), (i, j)


def test_tile_pipeline_and_sharding_equal_direct_inference():
    """TilePipeline (pinned double-buffered H2D / D2H around the batched forward) and infer_tiles_sharded (rank slices, no
    collective) return exactly the canvases infer_grid produces tile by tile"""
    from oracle import nets
    from satlas_super_resolution_b200.archs import SSR_RRDBNet
    from satlas_super_resolution_b200.infer import infer_grid, infer_tiles_sharded
    nb, grid, n_tiles = 1, 4, 5
    net = SSR_RRDBNet(24, 3, num_block=nb)
    net.load_state_dict(nets.rrdbnet_init(24, 3, num_block=nb, seed=8))
    net = net.cuda().eval()
    g = torch.Generator().manual_seed(9)
    tiles = [torch.randint(0, 256, (grid * grid, 24, 32, 32), generator=g, dtype=torch.uint8) for _ in range(n_tiles)]
    want = [infer_grid(net, t, grid_size=grid, batch=16).cpu() for t in tiles]
    got = {}
    for rank in range(2):                                   # two "ranks" on one device: 3 + 2 tiles
        part = infer_tiles_sharded(net, tiles, rank=rank, world=2, batch=16)
        assert set(part) == set(range(3)) if rank == 0 else set(part) == {3, 4}
        got.update(part)
    assert sorted(got) == list(range(n_tiles))
    for i in range(n_tiles):
        assert torch.equal(got[i], want[i]), i


def test_infer_grid_dir_writes_the_reference_files(tmp_path):
    """infer_grid_dir: the directory workflow of ssr/infer_grid.py:46-85 -- PNG chunks in, stitched_sr.png / stitched_s2.png out
    (threaded decode / encode around the batched GPU pass); incomplete tiles are skipped as the reference does"""
    import random
    import cv2
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from satlas_super_resolution_b200.archs import SSR_RRDBNet
    from satlas_super_resolution_b200.infer import infer_grid, infer_grid_dir, load_tile_dir
    torch.manual_seed(0)
    net = SSR_RRDBNet(num_in_ch=6, num_out_ch=3, num_block=1).cuda().eval()
    rs = np.random.RandomState(1)
    grid, T, n = 2, 3, 2
    data = tmp_path / "in"
    for tile, count in (("t0", grid * grid), ("t1", grid * grid), ("short", 1)):
        (data / tile).mkdir(parents=True)
        for k in range(count):
            i, j = divmod(k, grid)
            im = rs.randint(1, 256, (T * 32, 32, 3)).astype(np.uint8)
            cv2.imwrite(str(data / tile / f"{i}_{j}.png"), cv2.cvtColor(im, cv2.COLOR_RGB2BGR))
    out = tmp_path / "out"
    res = infer_grid_dir(net, str(data), str(out), n_s2_images=n, grid_size=grid, threads=4, batch=4, rng=random.Random(3), write_chunks=True)
    assert res["tiles"] == ["t0", "t1"] and res["skipped"] == ["short"]
    rng = random.Random(3)
    with ThreadPoolExecutor(2) as pool:
        for tile in ("t0", "t1"):
            stack, s2 = load_tile_dir(str(data / tile), n, grid, pool, rng)
            want = infer_grid(net, stack.cuda(), grid_size=grid, batch=4).cpu().numpy()
            got = cv2.cvtColor(cv2.imread(str(out / tile / "stitched_sr.png")), cv2.COLOR_BGR2RGB)
            assert np.array_equal(got, want)
            assert np.array_equal(cv2.cvtColor(cv2.imread(str(out / tile / "stitched_s2.png")), cv2.COLOR_BGR2RGB), s2)
            chunk = cv2.cvtColor(cv2.imread(str(out / tile / "1_0.png")), cv2.COLOR_BGR2RGB)
            assert np.array_equal(chunk, want[128:256, 0:128])
