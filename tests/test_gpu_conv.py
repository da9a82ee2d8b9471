"""GPU parity tests of the tcgen05 implicit-GEMM conv and the stem, through the C ABI
(yb_conv2d_fwd / yb_stem_conv_fwd).  Reference: plain PyTorch fp32 conv2d on the same
fp16/bf16-rounded operands.  Tolerance: fp32-accumulation-order noise + one output
rounding: |err| <= 2^-9 * max(1, |ref|) for fp16 storage (2^-6 for bf16)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["1cta", "2cta", "mc", "1cta-reg", "2cta-reg", "1cta-eg1", "2cta-eg1", "1cta-eg2", "2cta-eg2"], autouse=True)
def conv_mode(request):
    """Every conv test runs against the three tcgen05 kernels: cta_group::1 (128-row tiles), cta_group::2 (a CTA
    pair per 256-row tile) and the cluster-multicast pair kernel (2x2 / 2x1 pairs; only 256-wide tiles take it,
    the other shapes fall back to the plain pair kernel), and — for the first two — against both epilogues: the
    TMA-store / TMA-residual one (default) and the register-store one of round 1 ("-reg"), and with one ("-eg1") or —
    wherever a kernel exists — two ("-eg2") groups of epilogue warps (the default picks two for 1x1 and Cin <= 64
    layers).  The library's option
    table (yb_set_option) overrides the heuristics; it is restored after each test."""
    L = _lib()
    mode, _, epi = request.param.partition("-")
    L.set_option("YB_CONV_MODE", "2cta" if mode == "mc" else mode)
    L.set_option("YB_CONV_MC", "1" if mode == "mc" else "0")
    L.set_option("YB_CONV_EPI", "reg" if epi == "reg" else None)
    L.set_option("YB_CONV_EG", epi[2:] if epi.startswith("eg") else None)
    yield request.param
    for k in ("YB_CONV_MODE", "YB_CONV_MC", "YB_CONV_EPI", "YB_CONV_EG"):
        L.set_option(k, None)


def _lib():
    from yolov3_tensorflow_b200 import _lib
    return _lib


def _run_conv(n, h, w, cin, cout, k, s, dtype=torch.float16, in_extra=0, out_extra=0, residual=False, upsample=False,
              out_fp32=False, leaky=True, stats=False, seed=0, halo=False):
    L = _lib()
    lib, check, ptr, st = L.lib, L.check, L.ptr, L.stream_handle
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(seed)
    in_ld = cin + in_extra
    xfull = (torch.randn((n, h, w, in_ld), generator=g) * 1.0).to(dtype).to(dev)
    in_off = in_extra // 2 // 8 * 8
    x = xfull[..., in_off:in_off + cin]
    wt = (torch.randn((cout, k, k, cin), generator=g) / (k * (cin ** 0.5))).to(dev)      # OHWI fp32
    scale = (torch.rand(cout, generator=g) + 0.5).to(dev)
    shift = (torch.randn(cout, generator=g) * 0.1).to(dev)
    cout_pad = lib.yb_conv_cout_pad(cout)
    wp = torch.zeros((cout_pad, k, k, cin), dtype=dtype, device=dev)
    code = L.YB_F16 if dtype == torch.float16 else L.YB_BF16
    check(lib.yb_pack_conv_weights(ptr(wt), L.YB_W_OHWI, cout, cin, k, cout_pad, code, ptr(wp), st()), "pack")
    sc = torch.ones(cout_pad, device=dev); sc[:cout] = scale
    sh = torch.zeros(cout_pad, device=dev); sh[:cout] = shift
    ho, wo = h // s, w // s
    up = 2 if upsample else 1
    out_ld = cout + out_extra
    out_off = out_extra // 2 // 8 * 8 if not out_fp32 else 0
    odt = torch.float32 if out_fp32 else dtype
    outfull = torch.full((n, ho * up, wo * up, out_ld), -7.0, dtype=odt, device=dev)
    res = None
    if residual:
        res = (torch.randn((n, ho, wo, cout), generator=g)).to(dtype).to(dev)
    d = L.ConvDesc(n=n, h=h, w=w, cin=cin, cout=cout, ksize=k, stride=s, in_ld=in_ld, out_ld=out_ld,
                   res_ld=cout, dtype=code, out_fp32=int(out_fp32), leaky=int(leaky), upsample2x=int(upsample))
    esz = 4 if out_fp32 else 2
    ssum = torch.zeros(cout_pad, device=dev) if stats else None
    ssq = torch.zeros(cout_pad, device=dev) if stats else None
    xp = C.c_void_p(xfull.data_ptr() + in_off * 2)
    op = C.c_void_p(outfull.data_ptr() + out_off * esz)
    if halo:
        assert lib.yb_conv3x3_halo_supported(C.byref(d)) == 1
        check(lib.yb_conv3x3_halo_fwd(C.byref(d), xp, ptr(wp), ptr(sc), ptr(sh), ptr(res), op, st()), "conv_halo")
    else:
        check(lib.yb_conv2d_fwd(C.byref(d), xp, ptr(wp), ptr(sc), ptr(sh), ptr(res), op, ptr(ssum), ptr(ssq), st()), "conv")
    torch.cuda.synchronize()
    # ---- reference (fp32 math on the rounded operands) ----
    xr = x.float().permute(0, 3, 1, 2)
    wr = wp[:cout].float().permute(0, 3, 1, 2)
    raw = F.conv2d(xr, wr, None, stride=s, padding=k // 2)
    y = raw * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if leaky:
        y = torch.where(y > 0, y, 0.1 * y)
    if residual:
        y = y + res.float().permute(0, 3, 1, 2)
    if upsample:
        y = y.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    ref = y.permute(0, 2, 3, 1).contiguous()
    got = outfull[..., out_off:out_off + cout].float()
    eps = 2.0 ** -9 if dtype == torch.float16 else 2.0 ** -6
    if out_fp32:
        eps = 1e-4
    err = (got - ref).abs()
    tol = eps * torch.clamp(ref.abs(), min=1.0)
    bad = err > tol
    if bad.any():
        idx = bad.nonzero()[0].tolist()
        raise AssertionError(
            f"conv mismatch n={n} h={h} w={w} cin={cin} cout={cout} k={k} s={s}: {int(bad.sum())}/{bad.numel()} bad, "
            f"max err {float(err.max()):.4g}, first bad at {idx}: got {float(got[tuple(idx)]):.5g} ref {float(ref[tuple(idx)]):.5g}; "
            f"got[0,0,0,:4]={got[0,0,0,:4].tolist()} ref[0,0,0,:4]={ref[0,0,0,:4].tolist()}")
    # untouched padding channels of a wider output buffer
    if out_extra and not out_fp32:
        mask = torch.ones(out_ld, dtype=torch.bool, device=dev); mask[out_off:out_off + cout] = False
        assert torch.all(outfull[..., mask] == -7.0), "conv wrote outside its channel slice"
    if stats:
        s_ref = raw.sum(dim=(0, 2, 3)); q_ref = (raw * raw).sum(dim=(0, 2, 3))
        torch.testing.assert_close(ssum[:cout], s_ref, rtol=2e-3, atol=2e-2 * float(raw.abs().max()))
        torch.testing.assert_close(ssq[:cout], q_ref, rtol=2e-3, atol=1e-2)
    return float(err.max())


def test_conv1x1_gemm_basic():
    _run_conv(2, 16, 16, 64, 128, 1, 1)          # M = 512: 4 full tiles, BN=128, BK=64


def test_conv1x1_tail_rows_and_bn64():
    _run_conv(3, 13, 13, 128, 64, 1, 1)          # M = 507: tail tile, BN=64


def test_conv1x1_bk32():
    _run_conv(2, 13, 13, 32, 64, 1, 1)           # cin = 32 -> 64B swizzle path


def test_conv1x1_long_k_multi_ntile():
    _run_conv(2, 13, 13, 1024, 512, 1, 1)        # 16 k-blocks (> pipeline depth), 4 n-tiles


def test_conv1x1_slices_residual():
    _run_conv(2, 26, 26, 256, 128, 1, 1, in_extra=128, out_extra=64, residual=True)


def test_conv1x1_head_fp32_255():
    _run_conv(2, 13, 13, 256, 255, 1, 1, out_fp32=True, leaky=False)


def test_conv1x1_upsample_into_concat():
    _run_conv(2, 13, 13, 128, 64, 1, 1, upsample=True, out_extra=128)


def test_conv3x3_s1_im2col():
    _run_conv(2, 16, 16, 64, 128, 3, 1)


def test_conv3x3_s1_odd_size_crossing_images():
    _run_conv(3, 13, 13, 64, 128, 3, 1, residual=True)   # 169 px/img: tiles straddle rows and images


def test_conv3x3_s1_bk32():
    _run_conv(2, 20, 12, 32, 64, 3, 1)            # non-square, cin=32


def test_conv3x3_s2():
    _run_conv(2, 52, 52, 64, 128, 3, 2)           # darknet pad-1 + stride 2 -> 26x26


def test_conv3x3_s2_nonsquare_bk32():
    _run_conv(1, 64, 96, 32, 64, 3, 2)


def test_conv3x3_input_slice_of_concat():
    _run_conv(2, 26, 26, 128, 256, 3, 1, in_extra=256)


def test_conv3x3_bf16_and_stats():
    _run_conv(2, 26, 26, 128, 128, 3, 1, dtype=torch.bfloat16, stats=True)


def test_conv3x3_tiny_tensor_under_128k():
    _run_conv(1, 4, 4, 64, 64, 3, 1)              # exercises the small-tensor im2col descriptor workaround


def test_conv3x3_wide_odd_mtiles_residual():
    _run_conv(3, 16, 16, 128, 512, 3, 1, residual=True)      # M=768: 3 m-tiles (odd: a pair of the 2x2 cluster idles), 2 n-tiles


def test_conv3x3_s2_wide():
    _run_conv(2, 52, 52, 128, 256, 3, 2)                     # 1 n-tile -> 2x1 cluster (B multicast only)


def test_conv1x1_wide_1024_bf16_stats():
    _run_conv(4, 13, 13, 512, 1024, 1, 1, dtype=torch.bfloat16, stats=True)   # 4 n-tiles, statistics epilogue


def test_conv1x1_residual_tail_rows_many_chunks():
    _run_conv(3, 13, 13, 128, 256, 1, 1, residual=True, out_extra=64)   # 507 rows: tail tile; 8 chunks/tile + TMA residual


def test_conv3x3_many_tiles_per_cta_residual():
    _run_conv(8, 52, 52, 64, 128, 3, 1, residual=True)                  # 169 m-tiles on <= 148 CTAs: residual prefetch across tiles


def test_conv1x1_cout32_single_chunk():
    _run_conv(2, 26, 26, 64, 32, 1, 1, residual=True)                   # cout 32: one valid chunk of a 64-wide tile


def test_conv_rejects_bad_arguments():
    L = _lib()
    d = L.ConvDesc(n=1, h=8, w=8, cin=24, cout=64, ksize=3, stride=1, in_ld=24, out_ld=64, res_ld=0, dtype=0,
                   out_fp32=0, leaky=1, upsample2x=0)
    t = torch.zeros(1 << 16, device="cuda")
    rc = L.lib.yb_conv2d_fwd(C.byref(d), L.ptr(t), L.ptr(t), L.ptr(t), L.ptr(t), None, L.ptr(t), None, None, L.stream_handle())
    assert rc == -1 and b"cin" in L.lib.yb_last_error_string()
    with pytest.raises(ValueError):
        L.check(rc, "conv")


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_stem_conv(dtype):
    L = _lib()
    lib, check, ptr, st = L.lib, L.check, L.ptr, L.stream_handle
    g = torch.Generator().manual_seed(3)
    n, h, w = 2, 40, 56
    x = torch.rand((n, h, w, 3), generator=g).cuda()
    wt = (torch.randn((32, 3, 3, 3), generator=g) * 0.2).cuda()    # OHWI
    sc = (torch.rand(32, generator=g) + 0.5).cuda()
    sh = (torch.randn(32, generator=g) * 0.1).cuda()
    out = torch.empty((n, h, w, 32), dtype=dtype, device="cuda")
    code = L.YB_F16 if dtype == torch.float16 else L.YB_BF16
    check(lib.yb_stem_conv_fwd(ptr(x), ptr(wt), ptr(sc), ptr(sh), n, h, w, 32, code, 1, ptr(out), st()), "stem")
    ref = F.conv2d(x.permute(0, 3, 1, 2), wt.permute(0, 3, 1, 2), None, padding=1)
    ref = ref * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    ref = torch.where(ref > 0, ref, 0.1 * ref).permute(0, 2, 3, 1)
    eps = 2.0 ** -9 if dtype == torch.float16 else 2.0 ** -6
    err = (out.float() - ref).abs()
    assert torch.all(err <= eps * torch.clamp(ref.abs(), min=1.0)), float(err.max())


def test_pack_weights_layouts():
    L = _lib()
    g = torch.Generator().manual_seed(1)
    cout, cin, k = 48, 32, 3
    ohwi = torch.randn((cout, k, k, cin), generator=g).cuda()
    hwio = ohwi.permute(1, 2, 3, 0).contiguous()
    oihw = ohwi.permute(0, 3, 1, 2).contiguous()
    for src, lay in ((hwio, L.YB_W_HWIO), (oihw, L.YB_W_OIHW), (ohwi, L.YB_W_OHWI)):
        dst = torch.full((64, k, k, cin), 9.0, dtype=torch.float16, device="cuda")
        L.check(L.lib.yb_pack_conv_weights(L.ptr(src), lay, cout, cin, k, 64, L.YB_F16, L.ptr(dst), L.stream_handle()), "pack")
        assert torch.equal(dst[:cout], ohwi.half())
        assert torch.all(dst[cout:] == 0)


# ------------------------------------------------------------------------- thin-layer kernels (csrc/conv_thin.cu)
@pytest.mark.parametrize("n,h,w,cout,s,res,dtype", [
    (2, 32, 48, 64, 1, True, torch.float16),      # darknet53_body/Conv_3 shape family (+ residual)
    (2, 64, 96, 64, 2, False, torch.float16),     # Conv_1: stride 2
    (1, 40, 24, 64, 1, False, torch.bfloat16),    # partial tiles (40 % 8 == 0, 24 % 16 != 0)
    (3, 26, 26, 32, 2, False, torch.float16),     # odd sizes, cout 32
])
def test_conv3x3_thin(n, h, w, cout, s, res, dtype, conv_mode):
    if conv_mode != "1cta":
        pytest.skip("independent of the tcgen05 kernel mode")
    L = _lib()
    g = torch.Generator().manual_seed(11)
    cin = 32
    x = torch.randn((n, h, w, cin), generator=g).to(dtype).cuda()
    wt = (torch.randn((cout, 3, 3, cin), generator=g) / (3 * cin ** 0.5)).cuda()
    cp = L.lib.yb_conv_cout_pad(cout)
    wp = torch.zeros((cp, 3, 3, cin), dtype=dtype, device="cuda")
    code = L.YB_F16 if dtype == torch.float16 else L.YB_BF16
    L.check(L.lib.yb_pack_conv_weights(L.ptr(wt), L.YB_W_OHWI, cout, cin, 3, cp, code, L.ptr(wp), L.stream_handle()), "pack")
    sc = (torch.rand(cp, generator=g) + 0.5).cuda(); sh = (torch.randn(cp, generator=g) * 0.1).cuda()
    ho, wo = h // s, w // s
    r = torch.randn((n, ho, wo, cout), generator=g).to(dtype).cuda() if res else None
    out = torch.full((n, ho, wo, cout), -7.0, dtype=dtype, device="cuda")
    d = L.ConvDesc(n=n, h=h, w=w, cin=cin, cout=cout, ksize=3, stride=s, in_ld=cin, out_ld=cout, res_ld=cout, dtype=code,
                   out_fp32=0, leaky=1, upsample2x=0)
    L.check(L.lib.yb_conv3x3_thin_fwd(C.byref(d), L.ptr(x), L.ptr(wp), L.ptr(sc), L.ptr(sh), L.ptr(r), L.ptr(out), L.stream_handle()), "thin")
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wp[:cout].float().permute(0, 3, 1, 2), None, stride=s, padding=1)
    y = y * sc[:cout].view(1, -1, 1, 1) + sh[:cout].view(1, -1, 1, 1)
    y = torch.where(y > 0, y, 0.1 * y)
    if res:
        y = y + r.float().permute(0, 3, 1, 2)
    ref = y.permute(0, 2, 3, 1)
    eps = 2.0 ** -9 if dtype == torch.float16 else 2.0 ** -6
    err = (out.float() - ref).abs()
    assert torch.all(err <= eps * torch.clamp(ref.abs(), min=1.0)), float(err.max())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_stem_conv_tensor_core(dtype, conv_mode):
    if conv_mode != "1cta":
        pytest.skip("independent of the tcgen05 kernel mode")
    L = _lib()
    g = torch.Generator().manual_seed(3)
    n, h, w = 2, 40, 56
    x = torch.rand((n, h, w, 3), generator=g).cuda()
    wt = (torch.randn((32, 3, 3, 3), generator=g) * 0.2).cuda()
    sc = (torch.rand(32, generator=g) + 0.5).cuda(); sh = (torch.randn(32, generator=g) * 0.1).cuda()
    out = torch.empty((n, h, w, 32), dtype=dtype, device="cuda")
    code = L.YB_F16 if dtype == torch.float16 else L.YB_BF16
    L.check(L.lib.yb_stem_conv_fwd_tc(L.ptr(x), L.ptr(wt), L.ptr(sc), L.ptr(sh), n, h, w, code, 1, L.ptr(out), L.stream_handle()), "stem_tc")
    # operands are rounded to the 16-bit storage type before the tensor-core product (the oracle's storage model)
    xr = x.to(dtype).float(); wr = wt.to(dtype).float()
    ref = F.conv2d(xr.permute(0, 3, 1, 2), wr.permute(0, 3, 1, 2), None, padding=1)
    ref = ref * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    ref = torch.where(ref > 0, ref, 0.1 * ref).permute(0, 2, 3, 1)
    eps = 2.0 ** -9 if dtype == torch.float16 else 2.0 ** -6
    err = (out.float() - ref).abs()
    assert torch.all(err <= eps * torch.clamp(ref.abs(), min=1.0)), float(err.max())
    # the training-forward form: same outputs, plus the batch statistics of the STORED values accumulated on top of what the
    # buffers hold (h = 40, w = 56: partial tiles in both directions, their out-of-image pixels must not be counted)
    out2 = torch.empty_like(out)
    ssum = torch.full((32,), 1.0, device="cuda"); ssq = torch.full((32,), 2.0, device="cuda")
    L.check(L.lib.yb_stem_conv_fwd_tc_stats(L.ptr(x), L.ptr(wt), L.ptr(sc), L.ptr(sh), n, h, w, code, 1, L.ptr(out2), L.ptr(ssum),
                                            L.ptr(ssq), L.stream_handle()), "stem_tc_stats")
    # this form multiplies split-precision operands (x = hi + lo, w = hi + lo): it matches the float32 convolution of the
    # UNROUNDED image and weights to the rounding of its 16-bit output (YB_STEM_SPLIT=0 gives the plain kernel's bits)
    ref32 = F.conv2d(x.permute(0, 3, 1, 2), wt.permute(0, 3, 1, 2), None, padding=1)
    ref32 = ref32 * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    ref32 = torch.where(ref32 > 0, ref32, 0.1 * ref32).permute(0, 2, 3, 1)
    ulp = 2.0 ** -11 if dtype == torch.float16 else 2.0 ** -8
    err2 = (out2.float() - ref32).abs()
    assert torch.all(err2 <= 1.02 * ulp * torch.clamp(ref32.abs(), min=1.0) + 1e-5), float(err2.max())
    L.set_option("YB_STEM_SPLIT", "0")
    try:
        out3 = torch.empty_like(out); s3 = torch.zeros(32, device="cuda"); q3 = torch.zeros(32, device="cuda")
        L.check(L.lib.yb_stem_conv_fwd_tc_stats(L.ptr(x), L.ptr(wt), L.ptr(sc), L.ptr(sh), n, h, w, code, 1, L.ptr(out3), L.ptr(s3),
                                                L.ptr(q3), L.stream_handle()), "stem_tc_stats")
        assert torch.equal(out3, out)
    finally:
        L.set_option("YB_STEM_SPLIT", None)
    of = out2.double()
    torch.testing.assert_close(ssum.double() - 1.0, of.sum(dim=(0, 1, 2)), rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(ssq.double() - 2.0, (of * of).sum(dim=(0, 1, 2)), rtol=1e-4, atol=1e-2)


# ------------------------------------------------------------------------- halo-tile tcgen05 conv (csrc/conv_halo.cu)
@pytest.mark.parametrize("n,h,w,cin,cout,s,res,dtype", [
    (2, 32, 16, 64, 128, 1, False, torch.float16),      # exact tiles
    (2, 32, 16, 32, 64, 1, True, torch.float16),
    (3, 40, 24, 64, 128, 1, True, torch.float16),       # partial bottom tile (40 = 2.5 x 16), several tiles per CTA row
    (2, 104, 104, 64, 128, 1, True, torch.bfloat16),    # the 416-input layer 6 / 8 geometry (6.5 tile rows)
    (1, 208, 208, 32, 64, 1, True, torch.float16),      # layer 3 geometry
    (2, 64, 32, 32, 64, 2, False, torch.float16),       # stride 2: four parity planes
    (3, 80, 48, 32, 64, 2, False, torch.bfloat16),      # stride 2, partial bottom tile (40 rows out)
    (2, 64, 32, 64, 64, 2, False, torch.float16),
    (1, 416, 416, 32, 64, 2, False, torch.float16),     # layer 1 geometry
    (2, 48, 40, 32, 128, 1, False, torch.float16),
    (2, 32, 16, 64, 64, 1, True, torch.bfloat16),
])
def test_conv3x3_halo(n, h, w, cin, cout, s, res, dtype, conv_mode):
    if conv_mode != "1cta":
        pytest.skip("independent of the igemm kernel selection")
    _run_conv(n, h, w, cin, cout, 3, s, dtype=dtype, residual=res, halo=True)
    _run_conv(n, h, w, cin, cout, 3, s, dtype=dtype, residual=res, halo=True, in_extra=16, out_extra=32, seed=3)


def test_conv3x3_halo_rejects_unsupported():
    L = _lib()
    d = L.ConvDesc(n=1, h=26, w=26, cin=256, cout=512, ksize=3, stride=1, in_ld=256, out_ld=512, res_ld=0, dtype=0,
                   out_fp32=0, leaky=1, upsample2x=0)
    assert L.lib.yb_conv3x3_halo_supported(C.byref(d)) == 0
    d2 = L.ConvDesc(n=1, h=20, w=20, cin=64, cout=128, ksize=3, stride=1, in_ld=64, out_ld=128, res_ld=0, dtype=0,
                    out_fp32=0, leaky=1, upsample2x=0)
    assert L.lib.yb_conv3x3_halo_supported(C.byref(d2)) == 0          # 20 % 8 != 0


# ------------------------------------------------------------------------- stem fused into Conv_1 (csrc/conv_halo.cu, STEMW)
@pytest.mark.parametrize("n,h,w,dtype", [(2, 64, 32, torch.float16), (3, 80, 48, torch.bfloat16), (1, 416, 416, torch.float16),
                                         (2, 96, 160, torch.float16)])
def test_stem_conv1_fused(n, h, w, dtype, conv_mode):
    """yb_stem_conv1_fused_fwd (the stem computed on the fly as the producer of Conv_1's halo planes) against the two
    separate launches (stem kernel, then Conv_1 on its 16-bit output) and against fp32 convs on the same rounded
    operands: image borders (the stem's SAME padding AND Conv_1's pad-1), partial bottom tiles, both storage types."""
    if conv_mode != "1cta":
        pytest.skip("independent of the igemm kernel selection")
    L = _lib()
    lib, check, ptr, st = L.lib, L.check, L.ptr, L.stream_handle
    g = torch.Generator(device="cpu").manual_seed(5 + h)
    dev = "cuda"
    code = L.YB_F16 if dtype == torch.float16 else L.YB_BF16
    x = torch.rand((n, h, w, 3), generator=g).to(dev)
    w0 = (torch.randn((32, 3, 3, 3), generator=g) / 5.0).to(dev)                       # OHWI fp32
    s0 = (torch.rand(32, generator=g) + 0.5).to(dev); b0 = (torch.randn(32, generator=g) * 0.1).to(dev)
    w1 = (torch.randn((64, 3, 3, 32), generator=g) / (3 * 32 ** 0.5)).to(dev)
    s1 = (torch.rand(64, generator=g) + 0.5).to(dev); b1 = (torch.randn(64, generator=g) * 0.1).to(dev)
    w1p = torch.zeros((64, 3, 3, 32), dtype=dtype, device=dev)
    check(lib.yb_pack_conv_weights(ptr(w1), L.YB_W_OHWI, 64, 32, 3, 64, code, ptr(w1p), st()), "pack")
    # ---- two launches
    a0 = torch.empty((n, h, w, 32), dtype=dtype, device=dev)
    check(lib.yb_stem_conv_fwd_tc(ptr(x), ptr(w0), ptr(s0), ptr(b0), n, h, w, code, 1, ptr(a0), st()), "stem")
    d = L.ConvDesc(n=n, h=h, w=w, cin=32, cout=64, ksize=3, stride=2, in_ld=32, out_ld=64, res_ld=0, dtype=code, out_fp32=0,
                   leaky=1, upsample2x=0)
    two = torch.empty((n, h // 2, w // 2, 64), dtype=dtype, device=dev)
    check(lib.yb_conv2d_fwd(C.byref(d), ptr(a0), ptr(w1p), ptr(s1), ptr(b1), None, ptr(two), None, None, st()), "conv1")
    # ---- fused
    one = torch.full((n, h // 2, w // 2, 64), -7.0, dtype=dtype, device=dev)
    check(lib.yb_stem_conv1_fused_fwd(C.byref(d), ptr(x), ptr(w0), ptr(s0), ptr(b0), ptr(w1p), ptr(s1), ptr(b1), ptr(one), st()), "fused")
    torch.cuda.synchronize()
    # ---- fp32 reference on the rounded operands (image and stem weights rounded to the storage type, like both kernels)
    xr = x.to(dtype).float().permute(0, 3, 1, 2)
    y0 = F.conv2d(xr, w0.to(dtype).float().permute(0, 3, 1, 2), None, stride=1, padding=1) * s0.view(1, -1, 1, 1) + b0.view(1, -1, 1, 1)
    y0 = torch.where(y0 > 0, y0, 0.1 * y0).to(dtype).float()
    y1 = F.conv2d(y0, w1p.float().permute(0, 3, 1, 2), None, stride=2, padding=1) * s1.view(1, -1, 1, 1) + b1.view(1, -1, 1, 1)
    ref = torch.where(y1 > 0, y1, 0.1 * y1).permute(0, 2, 3, 1)
    eps = 2.0 ** -8 if dtype == torch.float16 else 2.0 ** -5      # two rounding stages (stem output, Conv_1 output)
    for name, got in (("two launches", two), ("fused", one)):
        err = (got.float() - ref).abs()
        tol = eps * torch.clamp(ref.abs(), min=1.0)
        assert not (err > tol).any(), f"{name}: {int((err > tol).sum())} bad, max err {float(err.max()):.4g}"
    # the fused kernel sees bit-identical stem activations; only Conv_1's accumulation order may differ
    assert float((one.float() - two.float()).abs().max()) <= eps * max(1.0, float(ref.abs().max()))
