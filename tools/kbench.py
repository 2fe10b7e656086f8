#!/usr/bin/env python
"""Kernel microbenchmarks on one B200 (CUDA events, >=3 warm-ups, inputs larger than L2 or rotated buffers).

    python tools/kbench.py [gemm] [attn] [verify] [sample] [ln] [conv] [hbm] [--json out.json]

`verify` is BASELINE.json configs[4]: synthetic draft+target logits B x 680 tokens x V=4096, batch sweep, segment table =
the 256 px pyramid.  Numbers are ALGORITHMIC bytes (or FLOPs) / event time, against MEASURED_PEAKS.json."""
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdvar_b200 import _cabi  # noqa: E402

DEV = "cuda"
P256 = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)
LS = [p * p for p in P256]
SEG = [0] + [int(v) for v in np.cumsum(LS)]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = json.load(open(p)) if os.path.exists(p) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    return d


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def bench_gemm(out):
    pk = peaks()
    rows = []
    for C, name in ((1920, "d30"), (1024, "d16")):
        for M in (128, 1152, 4608, 12800, 32768, 54400):
            for kind, N, K, epi in (("qkv", 3 * C, C, "bf16"), ("proj", C, C, "resid"), ("fc1", 4 * C, C, "gelu"), ("fc2", C, 4 * C, "resid"),
                                    ("head", 4096, C, "f32")):
                A = torch.randn(M, K, device=DEV).bfloat16()
                W = torch.randn(N, K, device=DEV).bfloat16()
                bias = torch.zeros(N, device=DEV)
                E = _cabi.GemmEpilogue
                if epi == "f32":
                    o = torch.empty(M, N, device=DEV)
                    e = E(epilogue=_cabi.EPI_F32, bias=bias.data_ptr(), out_f32=o.data_ptr(), ldo=N)
                elif epi == "resid":
                    o = torch.zeros(M, N, device=DEV)
                    gate = torch.ones(1, N, device=DEV)
                    e = E(epilogue=_cabi.EPI_RESID_F32, bias=bias.data_ptr(), out_f32=o.data_ptr(), ldo=N, gate=gate.data_ptr(), ld_gate=N, tokens_per_img=M)
                else:
                    o = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
                    e = E(epilogue=_cabi.EPI_GELU_BF16 if epi == "gelu" else _cabi.EPI_BF16, bias=bias.data_ptr(), out_bf16=o.data_ptr(), ldo=N)
                ms = timeit(lambda: _cabi.gemm_bf16(A, K, W, K, M, N, K, e))
                ms_cublas = timeit(lambda: torch.matmul(A, W.t()))
                tf = 2.0 * M * N * K / ms / 1e9
                rows.append(dict(model=name, op=kind, M=M, N=N, K=K, ms=ms, tflops=tf, frac=tf / pk["bf16_tflops"], cublas_tflops=2.0 * M * N * K / ms_cublas / 1e9))
                print(f"gemm {name} {kind:5s} M={M:6d} N={N:5d} K={K:5d}  {ms:8.3f} ms  {tf:7.1f} TF/s ({tf / pk['bf16_tflops'] * 100:5.1f}% of burst)  cuBLAS {rows[-1]['cublas_tflops']:7.1f}")
                del A, W, o
    out["gemm"] = rows


def bench_attn(out):
    rows = []
    for H, name in ((30, "d30"), (16, "d16")):
        for B in (64,):
            imgs = 2 * B
            for stages in ([9], [8], [6], [3], [8, 9]):
                ls = [LS[s] for s in stages]
                Lq, kv_off = sum(ls), SEG[stages[0]]
                seg = [0] + [int(v) for v in np.cumsum(ls)]
                Lmax, Lp = 680, 680
                q = torch.randn(imgs, H, Lq, 64, device=DEV).bfloat16()
                kc = torch.nn.functional.normalize(torch.randn(imgs, H, Lmax, 64, device=DEV), dim=-1).bfloat16()
                vc = torch.randn(imgs, H, 64, Lp, device=DEV).bfloat16()
                o = torch.empty(imgs * Lq, H * 64, device=DEV, dtype=torch.bfloat16)
                sm = torch.full((H,), math.log(4.0), device=DEV)
                ms = timeit(lambda: _cabi.attention(q, kc, vc, imgs, H, Lq, Lmax, Lp, kv_off, [int(v) for v in seg], 1.0, o, logit_bound_log=sm))
                vis = sum(l * (kv_off + e) for l, e in zip(ls, seg[1:]))
                fl = 4.0 * 64 * vis * imgs * H
                rows.append(dict(model=name, stages=stages, ms=ms, tflops=fl / ms / 1e9))
                print(f"attn {name} B={B} stages={stages} Lq={Lq} kv={kv_off + Lq}  {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TF/s")
    out["attention"] = rows


def _verify_inputs(B, scale, same_support):
    L, V = 680, 4096
    g = torch.Generator(device=DEV).manual_seed(0)
    xt = torch.randn(B, L, V, device=DEV, generator=g) * scale
    xd = torch.randn(B, L, V, device=DEV, generator=g) * scale if not same_support else xt + torch.randn(B, L, V, device=DEV, generator=g) * scale * 0.3
    d = torch.multinomial(xd.view(-1, V).softmax(-1), 1, generator=torch.Generator(device=DEV).manual_seed(0)).view(B, L)
    u = torch.rand(B, L, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    noise = torch.empty(B * L, V, device=DEV).exponential_(generator=torch.Generator(device=DEV).manual_seed(2))
    return xt, xd, d, u, noise


def bench_verify(out):
    pk = peaks()
    rows = []
    L, V, S = 680, 4096, 10
    for scale in (0.05, 3.0):
        for B in (1, 4, 16, 64, 256, 1024):
            if B * L * V * 4 * 3 > 120e9:
                continue
            xt, xd, d, u, noise = _verify_inputs(B, scale, False)
            o = torch.empty(B, L, dtype=torch.int64, device=DEV); a = torch.empty(B, L, dtype=torch.uint8, device=DEV)
            fr = torch.empty(B, S, dtype=torch.int32, device=DEV); na = torch.empty(B, S, dtype=torch.int32, device=DEV)
            st = torch.empty(B, dtype=torch.int32, device=DEV); sm = torch.empty(4, dtype=torch.int32, device=DEV)
            ws = torch.zeros(_cabi.verify_workspace_ints(B, S), dtype=torch.int32, device=DEV)
            f = lambda: _cabi.verify_accept_resample(xt, xd, d, u, noise, B, L, V, SEG, o, a, None, None, fr, na, st, sm, ws)
            ms = timeit(f, iters=5 if B >= 256 else 20)
            rej = int(sm[2])
            by = B * L * (2 * V * 4 + 17.0) + rej * V * 4.0
            gbs = by / ms / 1e6
            rows.append(dict(scale=scale, B=B, ms=ms, reject_rate=rej / (B * L), gbs=gbs, frac=gbs / pk["hbm_gbs"]))
            print(f"verify s={scale} B={B:5d}  {ms:8.3f} ms  reject={rej / (B * L):.3f}  {gbs:7.0f} GB/s ({gbs / pk['hbm_gbs'] * 100:5.1f}% of measured HBM)")
            del xt, xd, noise
    out["verify"] = rows


def bench_sample(out):
    pk = peaks()
    rows = []
    V = 4096
    for (tk, tp) in ((0, 0.0), (900, 0.96)):
        for B, l in ((64, 256), (64, 64), (256, 256)):
            lg = torch.randn(2 * B, l, V, device=DEV) * 1.0
            noise = torch.empty(B * l, V, device=DEV).exponential_()
            idx = torch.empty(B, l, dtype=torch.int64, device=DEV)
            mixed = torch.empty(B, l, V, device=DEV)
            thr = float(np.float32(1 - tp)) if tp > 0 else -1.0
            for want_mixed in (False, True):
                f = lambda: _cabi.sample_cfg_topk_topp(lg, B, l, V, [0, l], [1.5], [0.5], tk, thr, noise, idx, mixed if want_mixed else None, None)
                ms = timeit(f)
                by = B * l * (3 * V * 4 + 8.0 + (V * 4 if want_mixed else 0))
                gbs = by / ms / 1e6
                rows.append(dict(top_k=tk, top_p=tp, B=B, l=l, mixed=want_mixed, ms=ms, gbs=gbs, frac=gbs / pk["hbm_gbs"]))
                print(f"sample k={tk} p={tp} B={B} l={l} mixed_out={want_mixed}  {ms:8.3f} ms  {gbs:7.0f} GB/s ({gbs / pk['hbm_gbs'] * 100:5.1f}%)")
    out["sample"] = rows


def bench_ln(out):
    pk = peaks()
    rows = []
    for C in (1920, 1024):
        for M in (4608, 32768, 54400):
            x = torch.randn(M, C, device=DEV)
            mod = torch.randn(128, 6 * C, device=DEV)
            o = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
            tpi = M // 128
            ms = timeit(lambda: _cabi.ln_modulate(x, M, C, tpi, mod.data_ptr() + 8 * C, mod.data_ptr() + 16 * C, 6 * C, 1e-6, o))
            gbs = M * C * 6.0 / ms / 1e6
            rows.append(dict(M=M, C=C, ms=ms, gbs=gbs))
            print(f"ln_modulate M={M} C={C}  {ms:8.3f} ms  {gbs:7.0f} GB/s ({gbs / pk['hbm_gbs'] * 100:5.1f}%)")
    out["ln_modulate"] = rows


def bench_hbm(out):
    """what a READ-ONLY stream reaches on this box (torch.sum over 4 GiB of fp32, and of bf16) next to the read+write copy
    figure MEASURED_PEAKS.json quotes: the roof the read-only row kernels (K4 verify, K3 without mixed output) can see"""
    pk = peaks()
    rows = []
    for dt, name in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        x = torch.ones(4 * 1024 ** 3 // torch.finfo(dt).bits * 8, device=DEV, dtype=dt)
        ms = timeit(lambda: x.sum(), iters=5)
        gbs = x.numel() * x.element_size() / ms / 1e6
        rows.append(dict(kind=f"read-only sum {name}", gbs=gbs, frac_of_copy_peak=gbs / pk["hbm_gbs"]))
        print(f"hbm read-only torch.sum {name}  {ms:8.3f} ms  {gbs:7.0f} GB/s ({gbs / pk['hbm_gbs'] * 100:5.1f}% of the measured copy peak)")
        del x
    a = torch.empty(1024 ** 3, device=DEV, dtype=torch.float32); b = torch.empty_like(a)
    ms = timeit(lambda: b.copy_(a), iters=5)
    gbs = 2 * a.numel() * 4 / ms / 1e6
    rows.append(dict(kind="copy fp32 (read+write)", gbs=gbs, frac_of_copy_peak=gbs / pk["hbm_gbs"]))
    print(f"hbm copy (read+write)          {ms:8.3f} ms  {gbs:7.0f} GB/s ({gbs / pk['hbm_gbs'] * 100:5.1f}%)")
    out["hbm"] = rows


def bench_conv(out):
    """decoder layer shapes at B=64 (ch=160, 256 px): tcgen05 implicit GEMM vs cuDNN on the same channels-last bf16 tensors"""
    pk = peaks()
    rows = []
    B = 64
    for (cin, cout, hw, taps) in ((160, 160, 256, 9), (160, 160, 128, 9), (320, 160, 128, 9), (320, 320, 128, 9), (320, 320, 64, 9),
                                  (640, 320, 32, 9), (320, 320, 32, 9), (640, 640, 32, 9), (640, 640, 16, 9), (32, 640, 16, 9),
                                  (640, 1920, 16, 1), (640, 320, 32, 1), (320, 160, 128, 1)):
        k = 3 if taps == 9 else 1
        x = torch.randn(B, cin, hw, hw, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
        w = (torch.randn(cout, cin, k, k, device=DEV) / math.sqrt(cin * taps)).bfloat16().contiguous(memory_format=torch.channels_last)
        wp = w.permute(2, 3, 0, 1).reshape(taps, cout, cin).contiguous()
        bias = torch.zeros(cout, device=DEV)
        y = torch.empty(B, cout, hw, hw, device=DEV, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ms = timeit(lambda: _cabi.conv_nhwc(x, B, hw, hw, cin, wp, taps, cout, bias, None, y=y), iters=5)
        ms_lib = timeit(lambda: torch.nn.functional.conv2d(x, w, None, padding=k // 2), iters=5)
        fl = 2.0 * B * hw * hw * cin * cout * taps
        rows.append(dict(cin=cin, cout=cout, hw=hw, taps=taps, ms=ms, tflops=fl / ms / 1e9, cudnn_ms=ms_lib, cudnn_tflops=fl / ms_lib / 1e9))
        print(f"conv {cin:4d}->{cout:4d} {hw:3d}x{hw:<3d} taps={taps}  {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TF/s ({fl / ms / 1e9 / pk['bf16_tflops'] * 100:5.1f}% of burst)"
              f"  cuDNN {ms_lib:8.3f} ms {fl / ms_lib / 1e9:7.1f}")
        del x, y
    out["conv"] = rows


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--") and not a.endswith(".json")] or ["gemm", "attn", "verify", "sample", "ln"]
    out = {"gpu": torch.cuda.get_device_name(0), "peaks": peaks()}
    for w in which:
        {"gemm": bench_gemm, "attn": bench_attn, "verify": bench_verify, "sample": bench_sample, "ln": bench_ln, "conv": bench_conv, "hbm": bench_hbm}[w](out)
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
