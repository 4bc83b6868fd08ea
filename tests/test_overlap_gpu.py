"""SSR_OVERLAP / cfg['overlap']: the step with the side lane (dense-block weight gradients beside the next input-gradient launch,
ground-truth VGG features and the discriminator's weight preparation beside the generator forward -- DESIGN.md section 4) computes
what the single-stream step computes.  Only the launch ORDER changes, so the two may differ by the f32-atomic run-to-run spread of
the weight gradients and by nothing else; a missing dependency (a buffer rewritten while a side launch still reads it) shows up as
an O(1) difference."""
import pytest
import torch

pytestmark = pytest.mark.gpu

NB = 9          # 27 dense blocks: three resident input-gradient launches (12 + 12 + 3) -> fork, join + fork, join + fork, final join


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def _batch(B=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(1, 256, (B, 24, 32, 32), generator=g, dtype=torch.uint8),
            torch.randint(1, 256, (B, 3, 128, 128), generator=g, dtype=torch.uint8))


def _run(overlap, graph, iters, B=2, phase_graphs=False):
    from oracle import losses, nets
    from satlas_super_resolution_b200.trainer import ESRGANTrainer
    gp, dp, vp = nets.rrdbnet_init(24, 3, num_block=NB, seed=1), nets.unet_disc_init(27, seed=2), losses.vgg19_init(seed=3)
    tr = ESRGANTrainer(gp, dp, vp, dict(ema_decay=0.999, lr=1e-4, network_g=dict(num_in_ch=24, num_block=NB), cuda_graph=graph,
                                        overlap=overlap, phase_graphs=phase_graphs))
    lr, hr = _batch(B)
    snaps = []
    for it in range(1, iters + 1):
        tr.feed_data(lr, hr)
        tr.optimize_parameters(it)
        torch.cuda.synchronize()
        snaps.append(({k: v.clone() for k, v in tr.g_grads().items()}, {k: v.clone() for k, v in tr.d_grads().items()},
                      dict(tr.get_current_log())))
    ws = tr.G.workspace(B, 32, 32, True)
    return tr, ws, snaps


def test_side_lane_first_step_equals_single_stream():
    _, ws0, base = _run(False, False, 1)
    _, ws1, over = _run(True, False, 1)
    _, _, base2 = _run(False, False, 1)
    assert not ws0.overlap_bwd and not ws0.bwd.has_side
    assert ws1.overlap_bwd and ws1.bwd.has_side, "the overlapped plan was not built: nothing was tested"
    worst = 0.0
    for which, floor in ((0, 1e-2), (1, 2e-2)):          # generator / discriminator gradients: several times the measured run-to-run floors
        for k, v in base[0][which].items():
            spread = rel_l2(base2[0][which][k], v)
            dev = rel_l2(over[0][which][k], v)
            worst = max(worst, dev)
            assert dev < 3 * spread + floor, (k, dev, spread)
    print(f"  worst overlapped-vs-single-stream gradient difference {worst:.3e}")
    for k, v in base[0][2].items():
        assert abs(over[0][2][k] - v) < 2e-3 * abs(v) + 2e-4, k     # same forward (up to the atomics of the spectral-norm sums)


def test_side_lane_inside_the_cuda_graph():
    """iteration 1 eager, 2 = capture + replay, 3 = replay: the captured step has the side lane as parallel graph branches"""
    _, _, eager = _run(False, False, 3)
    tr, ws, graph = _run(True, True, 3)
    assert tr._last_mode == "graph" and ws.overlap_bwd and tr._cap_stream is not None
    _, _, eager2 = _run(False, False, 3)
    for k in ("conv_first.weight", "body.4.rdb2.conv3.weight", "body.0.rdb1.conv1.weight", "conv_last.bias"):
        spread = rel_l2(eager2[1][0][k], eager[1][0][k])
        dev = rel_l2(graph[1][0][k], eager[1][0][k])
        print(f"  step-2 grad {k}: eager-vs-eager {spread:.3e}  overlapped-graph-vs-eager {dev:.3e}")
        assert dev < 3 * spread + 5e-2, k        # same bound as test_cuda_graph_replay_equals_eager (heavy-tailed spread, O(1) if wrong)
    for k, v in eager[2][2].items():
        assert abs(graph[2][2][k] - v) < 5e-3 * abs(v) + 5e-4, k


def test_side_lane_inside_per_phase_graphs():
    """world > 1 replays one graph per phase around its asynchronous all-reduces (captured on the same high-priority stream, sharing
    one memory pool); `phase_graphs` takes that capture path on one GPU: the side lane forks and joins inside the phase-1 graph."""
    _, _, eager = _run(False, False, 3)
    tr, ws, graph = _run(True, True, 3, phase_graphs=True)
    assert tr._last_mode == "graph" and ws.overlap_bwd and len(next(iter(tr._graphs.values()))) == 4
    _, _, eager2 = _run(False, False, 3)
    for k in ("conv_first.weight", "body.4.rdb2.conv3.weight", "conv_last.bias"):
        spread = rel_l2(eager2[1][0][k], eager[1][0][k])
        dev = rel_l2(graph[1][0][k], eager[1][0][k])
        print(f"  step-2 grad {k}: eager-vs-eager {spread:.3e}  per-phase-graphs-vs-eager {dev:.3e}")
        assert dev < 3 * spread + 5e-2, k
    for k, v in eager[2][2].items():
        assert abs(graph[2][2][k] - v) < 5e-3 * abs(v) + 5e-4, k


@pytest.mark.parametrize("split", ["pool3", "all", "pool1"])
def test_perceptual_pass_split_over_the_side_lane_equals_the_2b_batch(split):
    """forward_gt (ground-truth half of the layers up to `split` on the side lane) + loss_and_grad(gt_lane=...) (generated half, join,
    deeper layers as one 2B batch) == the plain 2B pass: same launches over the same buffers, only batched differently"""
    from oracle import losses
    from satlas_super_resolution_b200.ops import SideLane, cur_stream
    from satlas_super_resolution_b200.vgg import PerceptualEngine
    B, H = 3, 128
    vp = {k: v.cuda() for k, v in losses.vgg19_init(seed=3).items()}
    g = torch.Generator().manual_seed(4)
    x = torch.rand(B, 3, H, H, generator=g).cuda()
    gt = torch.rand(B, 3, H, H, generator=g).cuda()
    res = []
    for lane_split in (None, split):
        eng = PerceptualEngine(vp, losses.DEFAULT_LAYER_WEIGHTS, split_upto=lane_split)
        loss = torch.zeros(1, device="cuda")
        dx = torch.zeros(B, 3, H, H, device="cuda")
        if lane_split is None:
            eng.loss_and_grad(x, gt, loss, dx)
        else:
            lane = SideLane.get()
            lane.fork(cur_stream())
            eng.forward_gt(gt, lane)
            eng.loss_and_grad(x, gt, loss, dx, gt_lane=lane)
            plans = eng.workspace(B, H, H).half_plans(split)
            n_deep = {"pool3": 9, "all": 0, "pool1": 17}[split]      # 16 convs + 4 pools: 11 / 20 / 3 of them per half
            assert len(plans[2]) == n_deep and len(plans[0]) == len(plans[1]) == 20 - n_deep
        torch.cuda.synchronize()
        feats = [item[3].t.float().clone() for item in eng.workspace(B, H, H).order if item[0] == "conv"]
        res.append((loss.item(), dx.clone(), feats))
    # the same tiles through the same kernel: expected bit-identical; asserted to 1e-5 so that a batch-dependent tiling choice inside
    # the conv launcher (another accumulation order) could not fail it, while a wrong half / offset is O(1)
    worst = max(rel_l2(a, b) for a, b in zip(res[1][2], res[0][2]))
    print(f"  split {split}: worst feature difference {worst:.2e}, loss {res[1][0]:.7f} vs {res[0][0]:.7f}")
    assert worst < 1e-5
    assert abs(res[1][0] - res[0][0]) <= 1e-5 * abs(res[0][0])
    assert rel_l2(res[1][1], res[0][1]) < 1e-4
