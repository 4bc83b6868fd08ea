"""Batched GPU validation metrics (satlas_super_resolution_b200/metrics.py) against the numpy oracle (oracle/metrics.py, itself pinned
to the reference's ssr/metrics/cpsnr.py through tests/golden/metrics_cpsnr.json) -- SURVEY.md 8f row 3."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pairs(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(B, C, H, W, generator=g)
    # a shifted, biased, noisy version with values outside [0, 1] (tensor2img clamps) and exact .5 / 255 ties (rounds half to even)
    sr = torch.roll(gt, (1, -2), (2, 3)) * 0.9 + 0.07 + 0.05 * torch.randn(B, C, H, W, generator=g)
    sr[:, :, :4, :4] = (torch.arange(16).view(4, 4).float() + 0.5) / 255
    return sr, gt


@pytest.mark.parametrize("C,crop", [(3, 0), (3, 4), (1, 2)])
def test_psnr_ssim_cpsnr_match_the_oracle(C, crop):
    from oracle import metrics as om
    from satlas_super_resolution_b200 import metrics as M
    B, H, W = 5, 128, 128
    sr, gt = _pairs(B, C, H, W, 7 + C)
    sr8, gt8 = M.to_uint8_images(sr.cuda()), M.to_uint8_images(gt.cuda())
    torch.cuda.synchronize()
    want_sr = np.stack([om.tensor2img(sr[i]) for i in range(B)])
    want_gt = np.stack([om.tensor2img(gt[i]) for i in range(B)])
    assert np.array_equal(sr8.cpu().numpy().reshape(want_sr.shape), want_sr)      # bit-exact tensor2img (clamp, half-to-even, BGR)
    assert np.array_equal(gt8.cpu().numpy().reshape(want_gt.shape), want_gt)
    psnr = M.calculate_psnr(sr8, gt8, crop_border=crop)
    ssim = M.calculate_ssim(sr8, gt8, crop_border=crop)
    cpsnr = M.calculate_cpsnr(sr8, gt8, crop_border=crop)
    for i in range(B):
        a, b = want_sr[i].reshape(H, W, C), want_gt[i].reshape(H, W, C)
        assert abs(psnr[i] - om.calculate_psnr(a, b, crop)) < 1e-9
        assert abs(cpsnr[i] - om.calculate_cpsnr(a, b, crop)) < 1e-9
        assert abs(ssim[i] - om.calculate_ssim(a, b, crop)) < 1e-9
    assert all(c >= p - 1e-9 for c, p in zip(cpsnr, psnr))          # the (8, 8)... offset search contains the aligned comparison at bias 0 or better
    # identical images: infinite PSNR like the reference
    assert M.calculate_psnr(gt8, gt8, crop_border=0)[0] == float("inf") and M.calculate_cpsnr(gt8, gt8, crop_border=0)[0] == float("inf")


def test_validation_loop_reports_batched_metrics(tmp_path):
    """SSRESRGANModel.validation (ssr_esrgan_model.py:269-352): EMA generator over the loader, PSNR / SSIM / cPSNR averaged over
    images, best-result bookkeeping -- against the same quantities computed one image at a time with the oracle"""
    from oracle import metrics as om
    from satlas_super_resolution_b200.registry import build_model
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("_modules_helpers", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_modules_gpu.py"))
    helpers = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(helpers)
    _batch, _opt = helpers._batch, helpers._opt
    opt = _opt(tmp_path, nb=1)
    opt["val"] = {"metrics": {"psnr": dict(type="calculate_psnr", crop_border=4, test_y_channel=False),
                              "ssim": dict(type="calculate_ssim", crop_border=4, test_y_channel=False),
                              "cpsnr": dict(type="calculate_cpsnr", crop_border=4, test_y_channel=False, better="higher")}}
    model = build_model(opt)
    loader = [_batch(B=2, seed=s) for s in (11, 12)]
    n = model.validation(loader, current_iter=7, tb_logger=None, save_img=False)
    assert n == 4
    want = {"psnr": [], "ssim": [], "cpsnr": []}
    for data in loader:
        model.feed_data(data)
        model.test()
        for i in range(2):
            a, b = om.tensor2img(model.output[i]), om.tensor2img(model.gt[i])
            want["psnr"].append(om.calculate_psnr(a, b, 4))
            want["ssim"].append(om.calculate_ssim(a, b, 4))
            want["cpsnr"].append(om.calculate_cpsnr(a, b, 4))
    for k, v in want.items():
        assert abs(model.metric_results[k] - float(np.mean(v))) < 1e-8, k
        assert model.best_metric_results["val"][k]["iter"] == 7
    opt["val"]["metrics"]["lpips"] = dict(type="calculate_lpips")
    with pytest.raises(NotImplementedError):
        build_model(opt).validation(loader, 1, None)
