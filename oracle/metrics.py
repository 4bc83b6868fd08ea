"""numpy restatement of the validation metrics (test infrastructure only, never imported by the product package):
ssr/metrics/cpsnr.py:7-59 (follows the reference loop literally), basicsr.metrics.psnr_ssim.calculate_psnr / calculate_ssim and
basicsr.utils.img_util.tensor2img (basicsr 1.4.2 is not vendored: restated from its published algorithm, cross-checked against
cv2.GaussianBlur-free direct evaluation in tests/test_oracle.py; cpsnr is pinned against the imported reference function)."""
import numpy as np


def tensor2img(t, rgb2bgr=True):
    """one image [C, H, W] float in any range -> uint8 HWC, clamp to [0, 1], * 255, np.round (half to even)"""
    a = np.clip(t.detach().float().cpu().numpy(), 0.0, 1.0).transpose(1, 2, 0)
    if a.shape[2] == 3 and rgb2bgr:
        a = a[:, :, ::-1]
    return (a * 255.0).round().astype(np.uint8)


def _crop(img, crop_border):
    return img[crop_border:-crop_border, crop_border:-crop_border, ...] if crop_border else img


def calculate_psnr(img, img2, crop_border):
    a, b = _crop(img, crop_border).astype(np.float64), _crop(img2, crop_border).astype(np.float64)
    mse = np.mean((a - b) ** 2)
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse)


def _ssim(img, img2):
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    x = np.arange(11, dtype=np.float64) - 5
    k = np.exp(-(x * x) / (2 * 1.5 * 1.5))
    k /= k.sum()
    win = np.outer(k, k)

    def filt(a):   # cv2.filter2D(a, -1, window)[5:-5, 5:-5] == 'valid' cross-correlation (the window is symmetric)
        h, w = a.shape
        out = np.zeros((h - 10, w - 10))
        for dy in range(11):
            for dx in range(11):
                out += win[dy, dx] * a[dy:dy + h - 10, dx:dx + w - 10]
        return out
    mu1, mu2 = filt(img), filt(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    s1, s2, s12 = filt(img ** 2) - mu1_sq, filt(img2 ** 2) - mu2_sq, filt(img * img2) - mu1_mu2
    return (((2 * mu1_mu2 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2))).mean()


def calculate_ssim(img, img2, crop_border):
    a, b = _crop(img, crop_border).astype(np.float64), _crop(img2, crop_border).astype(np.float64)
    return float(np.mean([_ssim(a[..., i], b[..., i]) for i in range(a.shape[2])]))


def calculate_cpsnr(img, img2, crop_border):
    """ssr/metrics/cpsnr.py:7-59, statement by statement"""
    img1, img2 = _crop(img, crop_border).astype(np.float64), _crop(img2, crop_border).astype(np.float64)
    max_offset = 8
    height, width = img1.shape[0], img1.shape[1]
    crop_height, crop_width = height - max_offset, width - max_offset
    best_mse = None
    for row_offset in range(max_offset + 1):
        for col_offset in range(max_offset + 1):
            cur1 = img1[row_offset:, col_offset:][0:crop_height, 0:crop_width].copy()
            cur2 = img2[(max_offset - row_offset):, (max_offset - col_offset):][0:crop_height, 0:crop_width].copy()
            for c in range(img1.shape[2]):
                cur2[:, :, c] += np.mean(cur1[:, :, c] - cur2[:, :, c])
            mse = np.mean(np.square(cur1 - cur2))
            if best_mse is None or mse < best_mse:
                best_mse = mse
    return float("inf") if best_mse == 0 else 10.0 * np.log10(255.0 * 255.0 / best_mse)
