"""Diagnostics: clock64 timeline of one chained dense-block launch (forward and input-gradient form) at B=32, 32x32.

SSR_CHAIN_TIMELINE=1 python scripts/chain_timeline.py
"""
import ctypes as C
import os
import sys

os.environ.setdefault("SSR_CHAIN_TIMELINE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from satlas_super_resolution_b200 import _lib as L  # noqa: E402

lib = L.load()
B, H, W, nf, g = 32, 32, 32, 64, 32
cw = nf + 4 * g
EVENTS = ["in-ready", "last-issue", "stage0-full", "stageN-full", "acc0-full", "accN-full", "stores-out", "arrived"]


def pack(w, k_pad):
    cout, cin, r, _ = w.shape
    n_pad = C.c_int32(0)
    nbytes = lib.ssr_packed_weight_bytes(k_pad, cout, r, C.byref(n_pad))
    packed = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    wd = w.cuda().contiguous()
    L.check(lib.ssr_pack_conv_weight(wd.data_ptr(), cout, cin, r, L.PACK_FWD, None, packed.data_ptr(), k_pad, n_pad.value, None))
    return packed, n_pad.value


def fwd_chain():
    buf = (torch.randn(B, H, W, cw) * 0.5).cuda().to(torch.bfloat16)
    nxt = torch.zeros(B, H, W, nf, dtype=torch.bfloat16, device="cuda")
    t32 = torch.zeros(B, H, W, nf, device="cuda")
    trunk = torch.randn(B, H, W, nf).cuda()
    arr = (L.ConvTcArgs * 5)()
    keep = [buf, nxt, t32, trunk]
    for k in range(5):
        cin, cout = nf + k * g, (g if k < 4 else nf)
        packed, n_pad = pack(torch.randn(cout, cin, 3, 3) * 0.02, (cin + 63) // 64 * 64)
        bias = torch.zeros(cout, device="cuda")
        keep += [packed, bias]
        a = arr[k]
        a.x, a.n_img, a.h, a.w, a.x_pix_stride, a.cin = buf.data_ptr(), B, H, W, cw, cin
        a.w_packed, a.r, a.cout, a.n_pad, a.bias = packed.data_ptr(), 3, cout, n_pad, bias.data_ptr()
        if k < 4:
            a.act, a.s0, a.out_bf16, a.out_pix_stride = 1, 1.0, buf.data_ptr() + 2 * cin, cw
        else:
            a.s0, a.res1, a.res1_kind, a.res1_pix_stride, a.s1 = 0.2, trunk.data_ptr(), L.SSR_F32_PLANAR4, nf, 1.0
            a.out_bf16, a.out_pix_stride = nxt.data_ptr(), nf
            a.out_f32, a.out32_mode, a.out32_pix_stride = t32.data_ptr(), L.OUT32_PLANAR4, nf
    return arr, keep


def dgrad_chain():
    xin = (torch.randn(B, H, W, nf) * 0.5).cuda().to(torch.bfloat16)
    cur = torch.randn(B, H, W, cw).cuda().to(torch.bfloat16)
    dg = torch.zeros(B, H, W, cw, dtype=torch.bfloat16, device="cuda")
    G32 = torch.zeros(B, H, W, cw, device="cuda")
    gout = torch.zeros(B, H, W, nf, dtype=torch.bfloat16, device="cuda")
    arr = (L.ConvTcArgs * 5)()
    bg = torch.zeros(5, 64, device="cuda")
    keep = [xin, cur, dg, G32, gout, bg]
    for i, k in enumerate(range(5, 0, -1)):
        nk = nf + (k - 1) * g
        cin = nf if k == 5 else g
        packed, n_pad = pack(torch.randn(nk, cin, 3, 3) * 0.02, 64)
        keep.append(packed)
        a = arr[i]
        a.x = xin.data_ptr() if k == 5 else dg.data_ptr() + 2 * nk
        a.x_pix_stride = nf if k == 5 else cw
        a.n_img, a.h, a.w, a.cin = B, H, W, cin
        a.w_packed, a.r, a.cout, a.n_pad = packed.data_ptr(), 3, nk, n_pad
        a.s0 = 0.2 if k == 5 else 1.0
        a.res1, a.res1_kind, a.res1_pix_stride, a.s1 = G32.data_ptr(), L.SSR_F32_PLANAR4, cw, 1.0
        if k == 5:
            a.res1_cmax = nf
        a.out_f32, a.out32_mode, a.out32_pix_stride = G32.data_ptr(), L.OUT32_PLANAR4, cw
        a.bias_grad, a.bias_grad_scale = bg.data_ptr() + 256 * i, 1.0
        if k > 1:
            a.mask, a.mask_pix_stride, a.mask_lo, a.out_lo = cur.data_ptr(), cw, nk - g, nk - g
            a.out_bf16, a.out_pix_stride = dg.data_ptr(), cw
        else:
            a.out_bf16, a.out_pix_stride = gout.data_ptr(), nf
    return arr, keep


def dgrad_chain_acc():
    """the input-gradient chain as the generator issues it: running sum in tensor memory, dY slots resident in shared memory"""
    xin = (torch.randn(B, H, W, nf) * 0.5).cuda().to(torch.bfloat16)
    cur = torch.randn(B, H, W, cw).cuda().to(torch.bfloat16)
    dg = torch.zeros(B, H, W, cw, dtype=torch.bfloat16, device="cuda")
    P = B * H * W
    incoming = (torch.randn(nf // 4, P, 4) * 0.3).cuda()
    G32 = torch.zeros(nf // 4, P, 4, device="cuda")
    gout = torch.zeros(B, H, W, nf, dtype=torch.bfloat16, device="cuda")
    arr = (L.ConvTcArgs * 5)()
    bg = torch.zeros(5, 64, device="cuda")
    keep = [xin, cur, dg, G32, gout, bg, incoming]
    for i, k in enumerate(range(5, 0, -1)):
        nk = nf + (k - 1) * g
        cin = nf if k == 5 else g
        packed, n_pad = pack(torch.randn(nk, cin, 3, 3) * 0.02, 64)
        keep.append(packed)
        a = arr[i]
        a.x = xin.data_ptr() if k == 5 else dg.data_ptr() + 2 * nk
        a.x_pix_stride = nf if k == 5 else cw
        a.n_img, a.h, a.w, a.cin = B, H, W, cin
        a.w_packed, a.r, a.cout, a.n_pad, a.s0 = packed.data_ptr(), 3, nk, n_pad, 1.0
        a.bias_grad, a.bias_grad_scale = bg.data_ptr() + 256 * i, 1.0
        if k > 1:
            a.mask, a.mask_pix_stride, a.mask_lo, a.out_lo = cur.data_ptr(), cw, nk - g, nk - g
            a.out_bf16, a.out_pix_stride = dg.data_ptr(), cw
        else:
            a.out_bf16, a.out_pix_stride = gout.data_ptr(), nf
            a.res1, a.res1_kind, a.s1 = incoming.data_ptr(), L.SSR_F32_PLANAR4, 1.0
            a.out_f32, a.out32_mode = G32.data_ptr(), L.OUT32_PLANAR4
    return arr, keep


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    s = torch.cuda.current_stream().cuda_stream
    mhz = 1965.0
    for name, (arr, keep) in (("forward", fwd_chain()), ("input-gradient (f32 sum in global memory)", dgrad_chain()),
                              ("input-gradient (sum in tensor memory)", dgrad_chain_acc())):
        acc = "tensor memory" in name

        def chain():
            L.check((lib.ssr_conv_tc_chain_acc if acc else lib.ssr_conv_tc_chain)(arr, 5, s))

        def plain():
            for k in range(5):
                L.check(lib.ssr_conv_tc(C.byref(arr[k]), s))
        r0 = lib.ssr_debug_resident_launches()
        t_chain = timeit(chain)
        resident = lib.ssr_debug_resident_launches() > r0
        print(f"== {name}: chained {t_chain:.1f} us ({'shared-memory-resident kernel' if resident else 'chain over global memory'})"
              + ("" if acc else f"   five launches {timeit(plain):.1f} us"))
        if resident:
            # rdb_resident_kernel stamps: 0 inputs complete (MMA warp), 2 first / 3 last weights landed, 5 accumulators complete,
            # 7 arrived on the cluster barrier, 6 global stores issued
            chain()
            torch.cuda.synchronize()
            n_ctas = 128
            tl = (C.c_longlong * (n_ctas * 5 * 8))()
            L.check(lib.ssr_debug_chain_timeline(tl, n_ctas))
            t = torch.tensor(list(tl), dtype=torch.float64).view(n_ctas, 5, 8)
            names = {0: "in-ready", 2: "w0-landed", 3: "wN-landed", 5: "acc-full", 1: "tile-written", 4: "fenced", 7: "arrived", 6: "stores-out"}
            for cta in (0, 1, 64, 127):
                t0 = min(t[cta, 0, 0].item(), t[cta, 0, 2].item())
                print(f"  CTA {cta}: us since the first stamp")
                for l in range(5):
                    print(f"    layer {l}: " + "  ".join(f"{names[e]} {(t[cta, l, e].item() - t0) / mhz:6.2f}" for e in (2, 0, 3, 5, 1, 4, 7, 6)))
            d = lambda a, b: ((t[:, :, a] - t[:, :, b]) / mhz).mean(0)
            print("  mean over CTAs per layer [us]:")
            print("    in-ready -> wN-landed (MMAs that waited for the previous layer):", [f"{v:.2f}" for v in d(3, 0).tolist()])
            print("    wN-landed -> acc-full  :", [f"{v:.2f}" for v in d(5, 3).tolist()])
            print("    acc-full  -> tile-written:", [f"{v:.2f}" for v in d(1, 5)[:4].tolist()])
            print("    tile-written -> fenced :", [f"{v:.2f}" for v in d(4, 1)[:4].tolist()])
            print("    fenced    -> arrived   :", [f"{v:.2f}" for v in d(7, 4)[:4].tolist()])
            print("    acc-full  -> arrived   :", [f"{v:.2f}" for v in d(7, 5)[:4].tolist()])
            print("    arrived (layer l) -> in-ready (layer l+1) at the MMA warp:", [f"{v:.2f}" for v in ((t[:, 1:, 0] - t[:, :-1, 7]) / mhz).mean(0).tolist()])
            print("    acc-full  -> stores-out:", [f"{v:.2f}" for v in d(6, 5).tolist()])
            print("    layer total (acc-full -> next acc-full):", [f"{v:.2f}" for v in ((t[:, 1:, 5] - t[:, :-1, 5]) / mhz).mean(0).tolist()])
            print(f"    first stamp -> last stores-out: {((t[:, 4, 6] - torch.minimum(t[:, 0, 0], t[:, 0, 2])) / mhz).mean().item():.2f}")
            continue
        chain()
        torch.cuda.synchronize()
        n_ctas = 128
        tl = (C.c_longlong * (n_ctas * 5 * 8))()
        L.check(lib.ssr_debug_chain_timeline(tl, n_ctas))
        t = torch.tensor(list(tl), dtype=torch.float64).view(n_ctas, 5, 8)
        for cta in (0, 1, 2, 3, 64, 127):
            t0 = t[cta, 0, 0].item()
            print(f"  CTA {cta}: us since the CTA's first TMA issue")
            for l in range(5):
                row = "  ".join(f"{EVENTS[e]} {(t[cta, l, e].item() - t0) / mhz:6.2f}" for e in range(8))
                print(f"    layer {l}: {row}")
        # averages over CTAs of the per-layer phases
        d = lambda a, b: ((t[:, :, a] - t[:, :, b]) / mhz).mean(0)
        print("  mean over CTAs per layer [us]:")
        print("    in-ready -> stage0-full :", [f"{v:.2f}" for v in d(2, 0).tolist()])
        print("    stage0   -> stageN-full :", [f"{v:.2f}" for v in d(3, 2).tolist()])
        print("    stageN   -> accN-full   :", [f"{v:.2f}" for v in d(5, 3).tolist()])
        print("    accN     -> stores-out  :", [f"{v:.2f}" for v in d(6, 5).tolist()])
        print("    stores   -> arrived     :", [f"{v:.2f}" for v in d(7, 6).tolist()])
        nxt = ((t[:, 1:, 0] - t[:, :-1, 7]) / mhz).mean(0)
        print("    arrived  -> next in-ready:", [f"{v:.2f}" for v in nxt.tolist()])
        print("    layer total (in-ready -> next in-ready):", [f"{v:.2f}" for v in ((t[:, 1:, 0] - t[:, :-1, 0]) / mhz).mean(0).tolist()])


if __name__ == "__main__":
    main()
