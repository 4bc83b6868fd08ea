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


@pytest.mark.parametrize("num_block,B", [(3, 2), (23, 2)])
def test_forward_matches_the_bf16_rounding_model(num_block, B):
    """Kernel error vs operand-precision noise, separated: oracle/nets.rrdbnet_forward_bf16_model applies the engine's bf16
    roundings (stored activations, weights) in an fp32 CPU evaluation.  Against it the engine must agree to f32-summation-order
    level -- two orders of magnitude below its distance to the plain fp32 oracle, which is then pure bf16 operand noise."""
    from oracle import nets
    from satlas_super_resolution_b200.generator import RRDBNetEngine
    p = nets.rrdbnet_init(24, 3, num_block=num_block, seed=40)
    x = torch.rand(B, 24, 32, 32, generator=torch.Generator().manual_seed(41))
    with torch.no_grad():
        plain = nets.rrdbnet_forward(p, x, num_block=num_block)
        model = nets.rrdbnet_forward_bf16_model(p, x, num_block=num_block)
    eng = RRDBNetEngine({k: v.cuda() for k, v in p.items()}, 24, 3, num_block=num_block, want_grad=False)
    eng.repack()
    out = eng.forward(x.cuda().contiguous(), train=False).clone().cpu()
    e_model, _ = _errs(out, model)
    e_plain, _ = _errs(out, plain)
    m_plain, _ = _errs(model, plain)
    print(f"blocks={num_block}: engine vs bf16 model {e_model:.3e}   engine vs fp32 oracle {e_plain:.3e}   bf16 model vs fp32 oracle {m_plain:.3e}")
    # measured on the B200: 3 blocks 3e-4 (vs 3.3e-3 to the fp32 oracle); 23 blocks 2.2e-3 (vs 7.4e-3): the flips of bf16 neighbours
    # that different f32 summation orders cause compound over 69 dense blocks, the remaining 7.4e-3 is operand rounding itself
    assert e_model < (1e-3 if num_block <= 3 else 4e-3) and e_model < 0.5 * e_plain


@pytest.mark.parametrize("num_block,B,cin,scale", [(3, 2, 24, 4), (23, 2, 24, 4), (2, 1, 96, 4), (1, 1, 3, 2)])
def test_split_bf16_forward_matches_fp32_oracle(num_block, B, cin, scale):
    """The tight-parity mode (tight.SplitBf16RRDBNet): every conv operand as a (hi, lo) pair of bf16 values, three launches of the SAME
    tcgen05 conv kernel per convolution -- ~2^-16 per layer, tighter than the TF32 the reference's own GPU path uses.  Against the
    plain fp32 oracle this must sit at fp32-noise level, two to three orders of magnitude below the bf16 production forward:
    what remains between the production forward and the oracle is operand rounding, not kernel error."""
    import json
    import os
    from oracle import nets
    from satlas_super_resolution_b200.generator import RRDBNetEngine
    from satlas_super_resolution_b200.tight import SplitBf16RRDBNet
    p = nets.rrdbnet_init(cin, 3, num_block=num_block, scale=scale, seed=50)
    x = torch.rand(B, cin, 32, 32, generator=torch.Generator().manual_seed(51))
    with torch.no_grad():
        ref = nets.rrdbnet_forward(p, x, scale=scale, num_block=num_block)
    pc = {k: v.cuda() for k, v in p.items()}
    tight = SplitBf16RRDBNet(pc, cin, 3, scale=scale, num_block=num_block)
    out_t = tight.forward(x.cuda()).cpu()
    eng = RRDBNetEngine(pc, cin, 3, scale=scale, num_block=num_block, want_grad=False)
    eng.repack()
    out_b = eng.forward(x.cuda().contiguous(), train=False).clone().cpu()
    e_t, m_t = _errs(out_t, ref)
    e_b, m_b = _errs(out_b, ref)
    print(f"blocks={num_block} cin={cin} scale={scale}: split-bf16 rel_l2={e_t:.3e} max/range={m_t:.3e}   bf16 rel_l2={e_b:.3e}")
    os.makedirs("gpurun_out/parity", exist_ok=True)
    with open(f"gpurun_out/parity/split_bf16_nb{num_block}_c{cin}_s{scale}.json", "w") as fh:
        json.dump(dict(num_block=num_block, B=B, cin=cin, scale=scale, split_bf16_rel_l2=e_t, split_bf16_max_over_range=m_t,
                       bf16_rel_l2=e_b, bf16_max_over_range=m_b), fh)
    assert out_t.shape == ref.shape
    assert e_t < 1e-4 and m_t < 3e-4          # the review's bar for a TF32-class mode was 1e-4; expected ~1e-5
    assert e_t < 0.05 * e_b


@pytest.mark.parametrize("num_block,B,cin", [(1, 2, 24), (3, 2, 24)])
def test_split_bf16_backward_matches_fp32_autograd(num_block, B, cin):
    """The tight-parity BACKWARD: the generator's parameter gradients through the same tcgen05 input-gradient (ssr_conv_tc over the
    mirrored operand) and weight-gradient (ssr_wgrad_tc) kernels with (hi, lo) bf16 operand pairs, against torch autograd of the fp32
    oracle.  Two comparisons, every tensor listed in gpurun_out/parity/:
      * the oracle evaluated with THIS forward's LeakyReLU pattern (oracle.nets.masked_lrelu): the kernels' arithmetic -- 1e-4;
      * the PLAIN oracle (its own pattern): a forward that agrees to 1e-5 still flips the sign of the ~1e-5 of all pre-activations
        that lie that close to zero, and a flipped derivative (1 <-> 0.2) is a discrete 0.8 |g| change: a gradient error of
        ~0.8 sqrt(fraction flipped) ~ 2e-3 per masked layer that NO operand precision removes (fp32 on another device has it too).
        Bounded at 2e-2 and reported with the count of flipped activations."""
    import json
    import os
    from oracle import nets
    from satlas_super_resolution_b200.tight import SplitBf16RRDBNet
    p = nets.rrdbnet_init(cin, 3, num_block=num_block, seed=60)
    g = torch.Generator().manual_seed(61)
    x = torch.rand(B, cin, 32, 32, generator=g)
    d_out = torch.randn(B, 3, 128, 128, generator=g) / (B * 3 * 128 * 128)
    net = SplitBf16RRDBNet({k: v.cuda() for k, v in p.items()}, cin, 3, num_block=num_block, want_grad=True)
    out = net.forward(x.cuda(), keep=True)
    grads = {k: v.cpu() for k, v in net.backward(d_out.cuda()).items()}
    torch.cuda.synchronize()
    # the activation pattern of the tight forward: sign of the stored (post-LeakyReLU) values
    sv = net._saved
    nchw = lambda t: (t.float() > 0).permute(0, 3, 1, 2).cpu()
    masks = {}
    for i in range(3 * num_block):
        blk, j = divmod(i, 3)
        for k in range(1, 5):
            lo = 64 + 32 * (k - 1)
            masks[f"body.{blk}.rdb{j + 1}.conv{k}"] = nchw(sv["bufs"][i].hi.t[..., lo:lo + 32])
    masks["conv_up1"], masks["conv_up2"], masks["conv_hr"] = nchw(sv["up_out"][0].hi.t), nchw(sv["up_out"][1].hi.t), nchw(sv["hr"].hi.t)
    flips = [0, 0]

    def counting(name, t):     # plain LeakyReLU that also counts where its pattern differs from the tight forward's
        flips[0] += int(((t > 0) != masks[name]).sum())
        flips[1] += t.numel()
        return torch.nn.functional.leaky_relu(t, 0.2)

    res = {}
    for tag, act in (("pattern", nets.masked_lrelu(masks)), ("plain", counting)):
        pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        ref = nets.rrdbnet_forward(pl, x, num_block=num_block, act=act)
        (ref * d_out).sum().backward()
        rows = {k: ((grads[k] - v.grad).norm() / (v.grad.norm() + 1e-30)).item() for k, v in pl.items()}
        res[tag] = dict(forward_rel_l2=_errs(out.cpu(), ref.detach())[0], grad_rel_l2=rows)
        worst = max(rows, key=rows.get)
        print(f"blocks={num_block} vs {tag} oracle: forward {res[tag]['forward_rel_l2']:.2e}; gradients: worst {rows[worst]:.2e} ({worst}), "
              f"median {sorted(rows.values())[len(rows) // 2]:.2e}")
    print(f"  activations whose LeakyReLU branch differs from the plain oracle's: {flips[0]} of {flips[1]} ({flips[0] / flips[1]:.1e})")
    os.makedirs("gpurun_out/parity", exist_ok=True)
    with open(f"gpurun_out/parity/split_bf16_grads_nb{num_block}.json", "w") as fh:
        json.dump(dict(num_block=num_block, B=B, cin=cin, flipped_activations=flips[0], activations=flips[1], **res), fh, indent=0)
    assert res["plain"]["forward_rel_l2"] < 1e-4
    assert max(res["pattern"]["grad_rel_l2"].values()) < 1e-4
    assert max(res["plain"]["grad_rel_l2"].values()) < 2e-2
