"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (through oracle/ref_shim.py) in the build
container.  Run from the repo root:  python -m oracle.make_golden

Each fixture stores the config, the seeds that regenerate weights (oracle.nets.*_init) and input, a checksum of the
weights, and the reference OUTPUT tensors.  tests/test_oracle.py replays them against oracle/nets.py on any machine.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import nets, ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def main():
    os.makedirs(OUT, exist_ok=True)
    RRDB, UNetD = ref_shim.reference_archs()
    torch.manual_seed(0)
    torch.set_num_threads(8)
    # ---- generator: small (2 blocks), full (23 blocks, 8-frame RGB), plumbing config (1 frame), 12-band
    for tag, cin, nb, B, seed in (("g_small", 24, 2, 2, 11), ("g_full_rgb8", 24, 23, 1, 12), ("g_cfg1_1frame", 3, 23, 1, 13),
                                  ("g_12band", 96, 2, 1, 14)):
        sd = nets.rrdbnet_init(cin, 3, num_block=nb, seed=seed)
        m = RRDB(num_in_ch=cin, num_out_ch=3, num_block=nb)
        missing = m.load_state_dict(sd, strict=True)          # strict: proves the key schema of nets.py == reference
        m.eval()
        x = torch.rand(B, cin, 32, 32, generator=torch.Generator().manual_seed(seed + 100))
        with torch.no_grad():
            y = m(x)
        torch.save({"kind": "rrdbnet", "num_in_ch": cin, "num_block": nb, "batch": B, "seed": seed, "x_seed": seed + 100,
                    "param_checksum": checksum(sd), "n_params": sum(v.numel() for v in sd.values()), "y": y.clone(),
                    "load_result": str(missing)}, os.path.join(OUT, f"{tag}.pt"))
        print(tag, tuple(y.shape), "params", sum(v.numel() for v in sd.values()))
    # ---- discriminator: two training-mode forwards (power iteration advances), then eval
    for tag, cin, seed in (("d_rgb8", 27, 21), ("d_plain", 3, 22)):
        sd = nets.unet_disc_init(cin, seed=seed)
        m = UNetD(num_in_ch=cin)
        m.load_state_dict(sd, strict=True)
        m.train()
        x = torch.rand(1, cin, 64, 64, generator=torch.Generator().manual_seed(seed + 100))
        with torch.no_grad():
            y1 = m(x).clone()
            y2 = m(x).clone()
        after = {k: v.clone() for k, v in m.state_dict().items() if k.endswith(("weight_u", "weight_v"))}
        m.eval()
        with torch.no_grad():
            y3 = m(x).clone()
        # gradients of sum(logits * r) w.r.t. conv3.weight_orig and conv0.weight (spectral-norm backward included)
        m.train()
        m.load_state_dict(sd, strict=True)
        r = torch.randn(1, 1, 64, 64, generator=torch.Generator().manual_seed(seed + 200))
        (m(x) * r).sum().backward()
        torch.save({"kind": "unet_disc", "num_in_ch": cin, "seed": seed, "x_seed": seed + 100, "r_seed": seed + 200,
                    "param_checksum": checksum(sd), "y_train1": y1, "y_train2": y2, "y_eval": y3, "uv_after_2": after,
                    "grad_conv3_head": m.conv3.weight_orig.grad[:4].clone(),
                    "grad_conv3_abs_sum": float(m.conv3.weight_orig.grad.double().abs().sum()),
                    "grad_conv0": m.conv0.weight.grad.clone()},
                   os.path.join(OUT, f"{tag}.pt"))
        print(tag, tuple(y1.shape))
    # ---- arch_util.pixel_unshuffle (scale 1 / 2 front-end)
    from ssr.archs.arch_util import pixel_unshuffle
    x = torch.arange(2 * 3 * 8 * 8, dtype=torch.float32).view(2, 3, 8, 8)
    torch.save({"kind": "pixel_unshuffle", "x": x, "y2": pixel_unshuffle(x, 2), "y4": pixel_unshuffle(x, 4)},
               os.path.join(OUT, "pixel_unshuffle.pt"))
    # ---- infer_utils.format_s2naip_data is covered in tests/test_infer_utils.py (needs seeding of `random`)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
