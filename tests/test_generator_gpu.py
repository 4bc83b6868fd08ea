"""GPU parity: RRDBNetEngine forward (C ABI, tcgen05 convs) vs the fp32 CPU oracle (oracle/nets.py).

Tolerance: the engine rounds activations and weights to bf16 between layers (fp32 accumulate); the stated
bound is relative L2 error <= 1e-2 and max-abs error <= 3e-2 of the output range on the full 23-block net.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(num_block, B, cin, seed=0, hw=32):
    from oracle import nets
    from satlas_super_resolution_b200.generator import RRDBNetEngine
    p = nets.rrdbnet_init(cin, 3, num_block=num_block, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(B, cin, hw, hw, generator=g)
    with torch.no_grad():
        ref = nets.rrdbnet_forward(p, x, num_block=num_block)
    pc = {k: v.cuda() for k, v in p.items()}
    eng = RRDBNetEngine(pc, cin, 3, num_block=num_block, want_grad=False)
    eng.repack()
    outs = []
    for train in (False, True):
        out = eng.forward(x.cuda().contiguous(), train=train).clone()
        torch.cuda.synchronize()
        outs.append(out.cpu())
    return ref, outs


def _errs(got, ref):
    rel_l2 = ((got - ref).norm() / ref.norm()).item()
    max_abs = (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
    return rel_l2, max_abs


@pytest.mark.parametrize("num_block,B,cin", [(1, 1, 24), (2, 2, 24), (23, 2, 24), (23, 1, 3), (2, 1, 96)])
def test_rrdbnet_forward_parity(num_block, B, cin):
    ref, outs = _run(num_block, B, cin)
    for got in outs:
        assert got.shape == ref.shape
        rel_l2, max_abs = _errs(got, ref)
        print(f"blocks={num_block} B={B} cin={cin}: rel_l2={rel_l2:.3e} max_abs/range={max_abs:.3e}")
        assert rel_l2 < 1e-2 and max_abs < 3e-2
    # eval (rotating buffers) and train (all buffers kept) workspaces must agree bit for bit
    assert torch.equal(outs[0], outs[1])


def test_rrdbnet_other_tile_sizes():
    ref, outs = _run(2, 1, 24, seed=3, hw=48)
    rel_l2, max_abs = _errs(outs[0], ref)
    assert rel_l2 < 1e-2 and max_abs < 3e-2
