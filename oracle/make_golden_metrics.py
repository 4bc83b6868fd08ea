"""Golden values for the validation metrics from the UNMODIFIED reference function ssr/metrics/cpsnr.py (imported by file path in
the build container; the GPU box has no /root/reference).  Writes tests/golden/metrics_cpsnr.json.

    python oracle/make_golden_metrics.py
"""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402


def image_pair(seed, h=48, w=40, c=3):
    """a noisy, shifted, brightness-biased copy: the case cPSNR exists for"""
    rng = np.random.RandomState(seed)
    a = rng.randint(0, 256, (h, w, c)).astype(np.uint8)
    b = np.clip(np.roll(a.astype(int), (seed % 5 - 2, 1 - seed % 3), (0, 1)) + rng.randint(-6, 7, a.shape) + 3 * (seed % 7) - 9, 0, 255)
    return a, b.astype(np.uint8)


def main():
    ref_shim.install()
    spec = importlib.util.spec_from_file_location("_ref_cpsnr", os.path.join(ref_shim.REFERENCE_ROOT, "ssr", "metrics", "cpsnr.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = []
    for seed in range(6):
        for cb in (0, 4):
            a, b = image_pair(seed, c=3 if seed % 2 == 0 else 1)
            cases.append(dict(seed=seed, crop_border=cb, channels=a.shape[2], cpsnr=float(mod.calculate_cpsnr(a, b, crop_border=cb))))
    a, _ = image_pair(9)
    cases.append(dict(seed=9, crop_border=0, channels=3, identical=True, cpsnr=float(mod.calculate_cpsnr(a, a.copy(), crop_border=0))))
    out = os.path.join(ROOT, "tests", "golden", "metrics_cpsnr.json")
    with open(out, "w") as fh:
        json.dump(dict(source="ssr/metrics/cpsnr.py:7-59 (reference @ 3da1525), image_pair() of this script", cases=cases), fh, indent=1)
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
