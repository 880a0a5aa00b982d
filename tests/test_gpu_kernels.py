"""GPU parity of the non-GEMM kernels against the CPU oracle (through the C ABI / ops wrappers)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

import oracle
from oracle import model as omodel
from imageenhancement_mp_b200 import synth

pytestmark = pytest.mark.gpu


def to_raster(x):
    from imageenhancement_mp_b200 import ops
    n, h, w, c = x.shape
    data = F.pad(x, (0, 0, 0, 1, 1, 0)).reshape(-1, c).to(torch.bfloat16).contiguous()   # zero row above, zero pixel right
    return ops.Raster(data, n, h, w)


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def border_is_zero(r):
    full = r.data.float().view(r.n, r.h + 1, r.w + 1, -1)
    return bool((full[:, 0] == 0).all() and (full[:, :, -1] == 0).all())


# ------------------------------------------------------------------ layout glue
def test_im2col_first_layer(cuda):
    """im2col pack + 1x1 GEMM == Conv2D(64, 3, 'same') on the raw input (model_library.py:323, 376)."""
    from imageenhancement_mp_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = bf16_round(torch.rand(2, 16, 24, 5, generator=g))
    w = bf16_round(torch.randn(3, 3, 5, 64, generator=g) * 0.2)
    b = torch.randn(64, generator=g) * 0.1
    ref = omodel.conv2d_relu(x, (w, b), "same")
    src = ops.pack_input_im2col3x3(x.to(cuda))
    wp = ops.pack_conv_weights(w.to(cuda), ktot_pad=64)
    dst = ops.new_raster(2, 16, 24, 64, cuda)
    ops.conv2d(src.slice(), wp, b.to(cuda), dst.slice(), k=1)
    got = ops.raster_to_nhwc(dst.slice()).cpu()
    assert torch.allclose(got, bf16_round(ref), atol=2e-2, rtol=2 ** -7)
    assert border_is_zero(dst)


@pytest.mark.parametrize("c,n,hs,ws,h,w", [(5, 2, 16, 24, 16, 24), (3, 3, 13, 21, 16, 24), (5, 7, 100, 100, 104, 104),
                                           # Basis_kpn with T = 8 + dualparams: 10 channels, K = 90 -> two K blocks
                                           (10, 2, 16, 24, 16, 24), (10, 5, 61, 64, 64, 64),
                                           # staged source rows: tiny images (tiles span many rows and images), rows wider
                                           # than a tile, a source smaller than the raster, one row short of it
                                           (5, 9, 8, 8, 8, 8), (5, 1, 40, 300, 40, 300), (3, 4, 12, 20, 16, 24),
                                           (5, 4, 31, 32, 32, 32), (5, 3, 1, 1, 8, 8)])
@pytest.mark.parametrize("gather", [False, True])
def test_first_layer_fused_im2col(cuda, c, n, hs, ws, h, w, gather):
    """conv_first_staged_kernel (source rows staged with bulk copies, im2col rows built in shared memory; the default
    when the float count is a multiple of 4) and conv_first_kernel (per-thread gathers: `gather`, or the fallback)
    == im2col raster + 1x1 GEMM, bit for bit, and == Conv2D(64, 3, 'same', relu) of the (zero-padded) input
    (model_library.py:323, 376)."""
    from imageenhancement_mp_b200 import _lib, ops
    lib = _lib.load()
    lib.ie_conv_set_mode(-1, (1 << 15) if gather else 0)
    try:
        _first_layer_case(cuda, c, n, hs, ws, h, w)
    finally:
        lib.ie_conv_set_mode(-1, 0)


def _first_layer_case(cuda, c, n, hs, ws, h, w):
    from imageenhancement_mp_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = bf16_round(torch.rand(n, hs, ws, c, generator=g))
    wt = bf16_round(torch.randn(3, 3, c, 64, generator=g) * 0.2)
    b = torch.randn(64, generator=g) * 0.1
    kw = ops.im2col_width(c)
    wp = ops.pack_conv_weights(wt.to(cuda), ktot_pad=kw)
    fused = ops.new_raster(n, h, w, 128, cuda)
    fused.data.fill_(float("nan"))
    ops.conv_first_layer(x.to(cuda), wp, b.to(cuda), fused.slice(64, 64))
    src = ops.pack_input_im2col3x3(x.to(cuda), ops.new_raster(n, h, w, kw, cuda))
    two = ops.new_raster(n, h, w, 64, cuda)
    ops.conv2d(src.slice(), wp, b.to(cuda), two.slice(), k=1)
    torch.cuda.synchronize()
    assert torch.equal(fused.data[:, 64:], two.data)
    assert bool(torch.isnan(fused.data[:, :64]).all())          # wrote only its slice
    ref = omodel.conv2d_relu(F.pad(x, (0, 0, 0, w - ws, 0, h - hs)), (wt, b), "same")
    got = ops.raster_to_nhwc(fused.slice(64, 64)).cpu()
    assert torch.allclose(got, bf16_round(ref), atol=2e-2, rtol=2 ** -7)


@pytest.mark.parametrize("c", [5, 2, 3, 9])
def test_im2col_implicit_stride_padding(cuda, c):
    """A source smaller than the raster is zero-padded at the bottom/right (100x100 patches -> 104x104)."""
    from imageenhancement_mp_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 13, 21, c, generator=g)
    xp = F.pad(x, (0, 0, 0, 3, 0, 3))                                # -> 16 x 24
    kw = ops.im2col_width(c)
    a = ops.pack_input_im2col3x3(x.to(cuda), ops.new_raster(2, 16, 24, kw, cuda))
    b = ops.pack_input_im2col3x3(xp.to(cuda))
    assert torch.equal(a.data, b.data)
    full = b.data.float().view(2, 17, 25, kw)
    assert bool((full[..., 9 * c:] == 0).all()) and border_is_zero(b)
    # every tap of an interior pixel is the corresponding (zero-padded) neighbour
    ref = F.pad(xp, (0, 0, 1, 1, 1, 1))
    for tap in range(9):
        i, j = divmod(tap, 3)
        assert torch.equal(full[:, 1:17, 0:24, tap * c:(tap + 1) * c].cpu(), bf16_round(ref[:, i:i + 16, j:j + 24]))


def test_maxpool(cuda):
    from imageenhancement_mp_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = bf16_round(torch.randn(3, 12, 20, 192, generator=g))
    src = to_raster(x.to(cuda))
    dst = ops.new_raster(3, 6, 10, 128, cuda)
    dst.data.fill_(5.0)
    assert ops.maxpool2(src.slice(64, 64), dst.slice(64, 64)) is None
    got = ops.raster_to_nhwc(dst.slice(64, 64)).cpu()
    assert torch.equal(got, omodel.maxpool2(x[..., 64:128]))
    assert torch.all(dst.data[:, :64] == 5.0)
    dst.data[:, :64] = 0
    assert border_is_zero(dst)
    # fused Poolskip statistics: per-image channel means of the INPUT slice (GlobalAveragePooling2D, model_library.py:110)
    for coff, c in ((0, 192), (64, 128), (128, 64)):
        d2 = ops.new_raster(3, 6, 10, c, cuda)
        mean = ops.maxpool2(src.slice(coff, c), d2.slice(), want_mean=True)
        assert torch.equal(ops.raster_to_nhwc(d2.slice()).cpu(), omodel.maxpool2(x[..., coff:coff + c]))
        assert torch.allclose(mean.cpu(), x[..., coff:coff + c].mean(dim=(1, 2)), atol=1e-5, rtol=1e-5)


def to_dense(x):
    from imageenhancement_mp_b200 import ops
    n, h, w, c = x.shape
    return ops.Raster(x.reshape(-1, c).to(torch.bfloat16).contiguous(), n, h, w, 0)


@pytest.mark.parametrize("din,dout", [(False, True), (True, True), (True, False)])
def test_layout_glue_between_rasters_and_dense_tensors(cuda, din, dout):
    """max-pool, bilinear up-sample (x2 and general), channel mean, broadcast with dense NHWC tensors on either side:
    same numbers as the raster-to-raster kernels, nothing written outside the channel slice, borders (where the output
    has one) zeroed."""
    from imageenhancement_mp_b200 import ops
    g = torch.Generator().manual_seed(12)
    x = bf16_round(torch.randn(3, 12, 20, 192, generator=g))
    src = (to_dense if din else to_raster)(x.to(cuda))
    # max-pool (+ fused channel means)
    dst = ops.new_raster(3, 6, 10, 192, cuda, dense=dout)
    dst.data.fill_(5.0)
    mean = ops.maxpool2(src.slice(64, 128), dst.slice(64, 128), want_mean=True)
    assert torch.equal(ops.raster_to_nhwc(dst.slice(64, 128)).cpu(), omodel.maxpool2(x[..., 64:192]))
    assert torch.allclose(mean.cpu(), x[..., 64:192].mean(dim=(1, 2)), atol=1e-5, rtol=1e-5)
    assert torch.all(dst.data[:, :64] == 5.0)
    if not dout:
        dst.data[:, :64] = 0
        assert border_is_zero(dst)
    # channel mean of a dense tensor
    m = ops.channel_mean(src.slice(128, 64))
    assert torch.allclose(m.cpu(), x[..., 128:].mean(dim=(1, 2)), atol=1e-5, rtol=1e-5)
    # up-sampling
    for scale in (2, 4):
        up = ops.new_raster(3, 12 * scale, 20 * scale, 128, cuda, dense=dout)
        up.data.fill_(float("nan"))
        ops.upsample_bilinear(src.slice(0, 64), up.slice(64, 64), scale)
        got = ops.raster_to_nhwc(up.slice(64, 64)).cpu()
        assert torch.allclose(got, omodel.upsample_bilinear(x[..., :64], scale), atol=2e-2, rtol=2 ** -7)
        assert bool(torch.isnan(up.data[:, :64]).all())
        if not dout:
            assert border_is_zero(ops.Raster(up.data[:, 64:].contiguous(), up.n, up.h, up.w))
        else:
            assert bool(torch.isfinite(up.data[:, 64:]).all())
    # broadcast into a dense / raster slice
    bc = ops.new_raster(3, 2, 2, 128, cuda, dense=dout)
    bc.data.zero_()
    ops.broadcast_hw(m, bc.slice(64, 64))
    assert torch.equal(ops.raster_to_nhwc(bc.slice(64, 64)).cpu(), bf16_round(m.cpu())[:, None, None, :].expand(3, 2, 2, 64))
    assert torch.all(bc.data[:, :64] == 0)


@pytest.mark.parametrize("scale,h,w", [(2, 5, 7), (8, 2, 2), (2, 1, 1), (2, 13, 13)])
def test_upsample_bilinear(cuda, scale, h, w):
    from imageenhancement_mp_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = bf16_round(torch.randn(2, h, w, 64, generator=g))
    src = to_raster(x.to(cuda))
    dst = ops.new_raster(2, h * scale, w * scale, 128, cuda)
    dst.data.fill_(float("nan"))
    ops.upsample_bilinear(src.slice(), dst.slice(0, 64), scale)
    got = ops.raster_to_nhwc(dst.slice(0, 64)).cpu()
    ref = omodel.upsample_bilinear(x, scale)
    assert torch.allclose(got, ref, atol=2e-2, rtol=2 ** -7)
    assert border_is_zero(ops.Raster(dst.data[:, :64].contiguous(), dst.n, dst.h, dst.w))   # every border entry written
    assert bool(torch.isnan(dst.data[:, 64:]).all())                                          # nothing outside the slice


def test_channel_mean_and_broadcast(cuda):
    from imageenhancement_mp_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = bf16_round(torch.randn(3, 13, 26, 256, generator=g) + 0.5)
    src = to_raster(x.to(cuda))
    m = ops.channel_mean(src.slice(128, 128))
    ref = x[..., 128:].mean(dim=(1, 2))
    assert torch.allclose(m.cpu(), ref, atol=1e-5, rtol=1e-5)
    dst = ops.new_raster(3, 16, 16, 192, cuda)
    dst.data.zero_()
    ops.broadcast_hw(m, dst.slice(64, 128))
    got = ops.raster_to_nhwc(dst.slice(64, 128)).cpu()
    assert torch.equal(got, bf16_round(ref)[:, None, None, :].expand(3, 16, 16, 128))
    assert border_is_zero(dst)


@pytest.mark.parametrize("n,T,B", [(3, 4, 10), (2, 8, 90), (1, 8, 50), (2, 1, 1), (1, 2, 256), (1, 1, 300), (5, 3, 7)])
def test_softmax_taps(cuda, n, T, B):
    """Row-tiled kernel (B <= 256; 1024 threads when the basis is large) and the strided one (B > 256)."""
    from imageenhancement_mp_b200 import ops
    g = torch.Generator().manual_seed(5)
    ob = torch.relu(torch.randn(n, 15, 15, T * B, generator=g) * 2)
    got = ops.softmax_taps(ob.to(cuda), T, B).cpu()
    ref = omodel.basis_softmax(ob.double(), 15, T, B)
    assert got.shape == (n, 15, 15, T, B)
    assert torch.allclose(got.double(), ref, atol=1e-7, rtol=2e-5)
    assert torch.allclose(got.sum(dim=(1, 2, 3)), torch.ones(n, B), atol=1e-5)


# ------------------------------------------------------------------ per-pixel filter
def _kpn_inputs(n, h, w, T, B, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, h, w, T + 1, generator=g)
    coef = torch.softmax(torch.randn(n, h, w, B, generator=g) * 2, -1)
    bas = omodel.basis_softmax(torch.randn(n, 15, 15, T * B, generator=g) * 3, 15, T, B)
    return x, coef, bas


@pytest.mark.parametrize("n,h,w,T,B", [(2, 16, 24, 4, 10), (1, 40, 72, 2, 10), (1, 24, 24, 4, 7), (1, 19, 70, 3, 20)])
def test_kpn_apply_vs_literal(cuda, n, h, w, T, B):
    """Fused filter == the reference's tile/multiply/reduce_sum + Convolve (+perlayer), fp64 oracle."""
    from imageenhancement_mp_b200 import ops
    x, coef, bas = _kpn_inputs(n, h, w, T, B, 7)
    ref = oracle.kpn_apply_literal(x[..., :T].double(), coef.double(), bas.double())
    got = ops.kpn_apply(x.to(cuda), T, coef.to(cuda), bas.to(cuda)).cpu()
    assert torch.allclose(got.double(), ref, atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("n,h,w,T,B", [(2, 16, 24, 4, 10), (1, 40, 72, 2, 10), (3, 104, 104, 4, 10), (1, 21, 150, 3, 16),
                                       (1, 200, 300, 1, 7), (2, 64, 64, 8, 10), (1, 33, 47, 6, 12),
                                       # more than 16 bases: chunks of 16 accumulated into the output (Basis_kpn, remote/)
                                       (2, 24, 24, 4, 17), (1, 20, 30, 2, 32), (1, 33, 47, 3, 20), (1, 40, 72, 8, 90)])
def test_kpn_apply_tf32_tensor_core_variant(cuda, n, h, w, T, B):
    """The mma.sync TF32 variant: same contract; operands rounded to a 10-bit mantissa, fp32 accumulation.  Stated
    bound: the output is a convex combination of burst pixels, so |err| <= 2^-10 * max|burst| (observed ~1e-4)."""
    from imageenhancement_mp_b200 import ops
    x, coef, bas = _kpn_inputs(n, h, w, T, B, 11)
    x = torch.cat([x, torch.rand(n, h, w, 1)], -1) if T > 4 else x      # T+2 channels: the pitch is not T+1
    ref = oracle.kpn_apply_algebraic(x[..., :T].double(), coef.double(), bas.double())
    got = ops.kpn_apply(x.to(cuda), T, coef.to(cuda), bas.to(cuda), precision="tf32").cpu()
    # channel 0 is a convex combination of burst pixels: |err| <= 2^-10 max|burst| (two operands rounded to 2^-11
    # each); the per-frame channels carry the reference's factor T (model_library.py:164) on a partial sum
    bound = 2.0 ** -10 * float(x.abs().max()) * 1.05
    assert float((got[..., 0].double() - ref[..., 0]).abs().max()) <= bound
    assert float((got[..., 1:].double() - ref[..., 1:]).abs().max()) <= T * bound
    f32 = ops.kpn_apply(x.to(cuda), T, coef.to(cuda), bas.to(cuda)).cpu()
    assert float((got - f32).abs().max()) <= T * bound
    # partition of unity survives the rounding of the basis to within its TF32 resolution
    ones = torch.ones_like(x).to(cuda)
    got1 = ops.kpn_apply(ones, T, coef.to(cuda), bas.to(cuda), precision="tf32").cpu()
    if h > 14 and w > 14:
        assert float((got1[:, 7:-7, 7:-7, 0] - 1).abs().max()) <= 1e-3


@pytest.mark.parametrize("n,h,w,T,B,pad", [(2, 16, 24, 4, 10, 0), (1, 40, 72, 8, 10, 0), (3, 104, 104, 4, 10, 0),
                                           (1, 21, 150, 4, 16, 0), (1, 33, 47, 4, 32, 0), (2, 20, 28, 4, 10, 4),
                                           # more than 32 bases: blocks of 32 added into the output (Basis_kpn, remote/)
                                           (1, 24, 40, 4, 33, 0), (2, 40, 24, 8, 50, 0), (1, 64, 64, 8, 90, 0),
                                           # more images than SMs x tiles per CTA: ranges cross image boundaries
                                           (300, 8, 16, 4, 10, 0), (5, 50, 70, 4, 10, 0)])
def test_kpn_apply_tcgen05_variant(cuda, n, h, w, T, B, pad):
    """The tcgen05 filter-synthesis kernel: same contract and the same TF32 bound as the mma.sync variant, incl. a coef
    tensor at the stride-padded extent and a burst pitch that is not T + 1."""
    from imageenhancement_mp_b200 import ops
    x, coef, bas = _kpn_inputs(n, h, w, T, B, 13)
    x = torch.cat([x, torch.rand(n, h, w, 1)], -1) if T > 4 else x
    ref = oracle.kpn_apply_algebraic(x[..., :T].double(), coef.double(), bas.double())
    big = torch.rand(n, h + pad, w + pad, B)
    big[:, :h, :w] = coef
    got = ops.kpn_apply(x.to(cuda), T, big.to(cuda), bas.to(cuda), precision="tcgen05").cpu()
    bound = 2.0 ** -10 * float(x.abs().max()) * 1.05
    assert float((got[..., 0].double() - ref[..., 0]).abs().max()) <= bound
    assert float((got[..., 1:].double() - ref[..., 1:]).abs().max()) <= T * bound
    tf32 = ops.kpn_apply(x.to(cuda), T, big.to(cuda), bas.to(cuda), precision="tf32").cpu()
    assert float((got - tf32).abs().max()) <= T * bound
    # a second call into the same output buffer overwrites, never accumulates across calls
    out = torch.full_like(got, 7.0).to(cuda)
    again = ops.kpn_apply(x.to(cuda), T, big.to(cuda), bas.to(cuda), out=out, precision="tcgen05").cpu()
    assert torch.equal(again, got)
    with pytest.raises(Exception):
        ops.kpn_apply(x[..., :3].contiguous().to(cuda), 3, coef.to(cuda), bas[:, :, :, :3].contiguous().to(cuda),
                      precision="tcgen05")                      # T = 3 is outside its scope


def test_kpn_apply_coef_at_padded_size(cuda):
    """coef may be larger than the image (network runs at the stride-padded size): only the top-left is read."""
    from imageenhancement_mp_b200 import ops
    x, coef, bas = _kpn_inputs(2, 20, 28, 4, 10, 9)
    ref = ops.kpn_apply(x.to(cuda), 4, coef.to(cuda), bas.to(cuda))
    big = torch.rand(2, 24, 32, 10)
    big[:, :20, :28] = coef
    got = ops.kpn_apply(x.to(cuda), 4, big.to(cuda), bas.to(cuda))
    assert torch.equal(got, ref)


def test_kpn_apply_large_vs_algebraic(cuda):
    from imageenhancement_mp_b200 import ops
    x, coef, bas = _kpn_inputs(2, 104, 104, 4, 10, 8)
    ref = oracle.kpn_apply_algebraic(x[..., :4].double(), coef.double(), bas.double())
    got = ops.kpn_apply(x.to(cuda), 4, coef.to(cuda), bas.to(cuda)).cpu()
    assert torch.allclose(got.double(), ref, atol=2e-6, rtol=1e-5)
    # partition of unity (two softmaxes): a constant burst stays constant >= 7 px from the border
    ones = torch.ones_like(x).to(cuda)
    got1 = ops.kpn_apply(ones, 4, coef.to(cuda), bas.to(cuda)).cpu()
    assert torch.allclose(got1[:, 7:-7, 7:-7, 0], torch.ones(2, 90, 90), atol=1e-5)


@pytest.mark.parametrize("n,h,w,T,K", [(2, 12, 20, 4, 15), (1, 9, 7, 3, 5), (1, 17, 33, 8, 15)])
def test_convolve_layers_with_materialised_filters(cuda, n, h, w, T, K):
    """Convolve / cus_convolve / Convolve_perlayer (model_library.py:114-168) against the oracle's literal restatement."""
    from imageenhancement_mp_b200 import model_library as ml
    g = torch.Generator().manual_seed(5)
    imgs = torch.rand(n, h, w, T, generator=g)
    filts = torch.randn(n, h, w, K, K, T, generator=g) / (K * K * T)
    ref = omodel.convolve(imgs.double(), filts.double(), K)
    ref_pl = omodel.convolve_perlayer(imgs.double(), filts.double(), K)
    got = ml.Convolve(K)(imgs.to(cuda), filts.to(cuda)).cpu()
    got_pl = ml.Convolve_perlayer(K, T)(imgs.to(cuda), filts.to(cuda)).cpu()
    assert got.shape == (n, h, w) and got_pl.shape == (n, h, w, T)
    assert torch.allclose(got.double(), ref, atol=2e-6, rtol=1e-5)
    assert torch.allclose(got_pl.double(), ref_pl, atol=2e-6, rtol=1e-5)
    assert torch.equal(ml.cus_convolve(imgs.to(cuda), filts.to(cuda), K).cpu(), got)


# ------------------------------------------------------------------ metrics
def _metric_inputs(n, h, w, T, seed):
    x, truth = synth.make_batch(n, h, w, {"BURST_LENGTH": T}, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    recon = (truth[..., :1] + 0.03 * torch.randn(n, h, w, T + 1, generator=g)).contiguous()
    recon[0, 3, 3, 0] = 1.7 * truth[0, 3, 3, 1]          # exercise the x > 1 branch of the sRGB curve
    return recon, x, truth


def test_reference_api_metrics(cuda):
    """Function-by-function parity of the data_utils API (fp64 oracle, <=1e-5 relative)."""
    from imageenhancement_mp_b200 import data_utils as du
    T = 4
    recon, x, truth = _metric_inputs(3, 40, 56, T, 21)
    rc, xc, tc = recon.to(cuda), x.to(cuda), truth.to(cuda)
    wn_ref = truth[..., 1:2].double().mean(dim=1, keepdim=True).mean(dim=2, keepdim=True)
    wn = du.white_level_of(tc)
    assert torch.allclose(wn.cpu().double(), wn_ref, rtol=1e-6)
    vals = torch.tensor([0., .001, .0031308, .5, 1., 2.], device=cuda)
    assert torch.allclose(du.sRGBforward(vals).cpu().double(), oracle.sRGBforward(vals.cpu().double()), atol=1e-6)
    igt_ref = oracle.invert_preproc(truth[..., 0].double(), wn_ref)
    igt = du.invert_preproc(tc[..., 0], wn)
    assert igt.shape == (3, 24, 40)
    assert torch.allclose(igt.cpu().double(), igt_ref, atol=2e-6, rtol=1e-5)
    ide_ref = oracle.invert_preproc(recon[..., 0].double(), wn_ref)
    ide = du.invert_preproc(rc[..., 0], wn)
    rel = lambda a, b: abs(float(a) - float(b)) / max(abs(float(b)), 1e-12)
    assert rel(du.deblur_loss(ide, igt), oracle.deblur_loss(ide_ref, igt_ref)) < 1e-5
    assert rel(du.gradient_loss(ide, igt), oracle.gradient_loss(ide_ref, igt_ref)) < 1e-5
    assert rel(du.deblur_layer_loss(rc, igt, wn), oracle.deblur_layer_loss(recon.double(), igt_ref, wn_ref)) < 1e-5
    assert rel(du.psnr_deblur(ide, igt), oracle.psnr_deblur(ide_ref, igt_ref)) < 1e-5
    per, per_ref = du.psnr_each_layer(igt, wn, rc), oracle.psnr_each_layer(igt_ref, wn_ref, recon.double())
    for k in per_ref:
        assert rel(per[k], per_ref[k]) < 1e-5
    burst, burst_ref = xc[..., :T], x[..., :T].double()
    assert rel(du.psnr_burst0(igt, wn, burst), oracle.psnr_burst0(igt_ref, wn_ref, burst_ref)) < 1e-5
    assert rel(du.psnr_average_f(igt, wn, burst), oracle.psnr_average_f(igt_ref, wn_ref, burst_ref)) < 1e-5
    lay = du.invert_deblur_layer(rc, wn)
    assert lay.shape == (3, 24, 40 * T)
    assert torch.allclose(lay.cpu().double(), oracle.invert_deblur_layer(recon.double(), wn_ref), atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("n,h,w", [(3, 37, 52), (2, 19, 23), (1, 16, 4), (2, 65, 516)])
def test_basic_img_loss_shapes(cuda, n, h, w):
    """basic_img_loss / gradient_loss (data_utils.py:37-51): vectorised (w % 4 == 0) and scalar kernels, edge sizes."""
    from imageenhancement_mp_b200 import data_utils as du
    g = torch.Generator().manual_seed(h * w)
    a, b = torch.rand(n, h, w, generator=g), torch.rand(n, h, w, generator=g)
    rel = lambda p, q: abs(float(p) - float(q)) / max(abs(float(q)), 1e-12)
    assert rel(du.basic_img_loss(a.to(cuda), b.to(cuda)), oracle.basic_img_loss(a.double(), b.double())) < 1e-5
    assert rel(du.gradient_loss(a.to(cuda), b.to(cuda)), oracle.gradient_loss(a.double(), b.double())) < 1e-5


@pytest.mark.parametrize("n,h,w,T", [(3, 40, 56, 4), (2, 104, 104, 4), (1, 33, 49, 2), (2, 100, 100, 8)])
def test_fused_eval_metrics(cuda, n, h, w, T):
    """One fused pass == the reference's per-batch numbers (eval.py:144-182), fp64 oracle."""
    from imageenhancement_mp_b200 import data_utils as du
    recon, x, truth = _metric_inputs(n, h, w, T, 31)
    ref = oracle.eval_step(recon.double(), x.double(), truth.double(), T)
    got = du.eval_metrics(recon.to(cuda), x.to(cuda), truth.to(cuda), T)
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-12)
    assert rel(got["loss1"], ref["loss1"]) < 1e-5
    assert rel(got["perlayer_loss"], ref["perlayer_loss"]) < 1e-5
    assert rel(got["psnr"], ref["psnr"]) < 1e-5
    for t in range(T):
        assert rel(got["psnr_perlayer"][t], ref["psnr_perlayer"][t]) < 1e-5
    assert rel(got["psnr_noise0"], ref["psnr_noise0"]) < 1e-5
    assert rel(got["psnr_average"], ref["psnr_average"]) < 1e-5


@pytest.mark.parametrize("n,h,w,T,pitch", [(2, 40, 56, 4, 5), (1, 33, 49, 2, 2), (2, 50, 300, 3, 4), (1, 48, 160, 6, 8)])
def test_fused_eval_metrics_kernel_variants(cuda, n, h, w, T, pitch):
    """Row-streaming kernel (16-byte aligned tensors, every rows-per-batch / block width) and the tile kernel that
    serves misaligned views give the same per-image sums: even channel counts (bank conflicts, not errors), widths
    beyond one block, burst pitch > T, pointers offset by one float."""
    from imageenhancement_mp_b200 import _lib, data_utils as du
    g = torch.Generator().manual_seed(n * h + w)
    truth = torch.rand(n, h, w, 2, generator=g) * 0.5 + 0.25
    truth[..., 1] = torch.rand(n, 1, 1, generator=g) * 0.5 + 0.5
    recon = (truth[..., :1] + 0.03 * torch.randn(n, h, w, T + 1, generator=g)).contiguous()
    burst = (truth[..., :1] + 0.05 * torch.randn(n, h, w, pitch, generator=g)).contiguous()
    ref = oracle.eval_step(recon.double(), burst.double(), truth.double(), T)
    lib = _lib.load()
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-12)

    def check(got):
        for key in ("loss1", "perlayer_loss", "psnr", "psnr_noise0", "psnr_average"):
            assert rel(got[key], ref[key]) < 1e-5, key
        for t in range(T):
            assert rel(got["psnr_perlayer"][t], ref["psnr_perlayer"][t]) < 1e-5

    rc, bc, tc = recon.to(cuda), burst.to(cuda), truth.to(cuda)
    try:
        for rb, warps in ((0, 0), (1, 2), (3, 3), (4, 4)):
            lib.ie_eval_metrics_tune(rb, warps, 0)
            check(du.eval_metrics(rc, bc, tc, T))
        lib.ie_eval_metrics_tune(0, 0, 1)
        check(du.eval_metrics(rc, bc, tc, T))
    finally:
        lib.ie_eval_metrics_tune(0, 0, 0)

    def shifted(t):                                             # same values at a data pointer that is 4 (mod 16)
        buf = torch.empty(t.numel() + 1, device=cuda)
        v = buf[1:].view(t.shape)
        v.copy_(t)
        assert v.data_ptr() % 16 == 4 and v.is_contiguous()
        return v
    check(du.eval_metrics(shifted(rc), shifted(bc), shifted(tc), T))


def test_psnr_known_answer(cuda):
    from imageenhancement_mp_b200 import data_utils as du
    a = torch.rand(4, 50, 60, device=cuda)
    assert abs(float(du.psnr_tf_batch(a + 0.1, a)) - 20.0) < 1e-4      # constant error 0.1 -> 20 dB


@pytest.mark.parametrize("n,h,w", [(2, 11, 11), (2, 64, 80), (1, 100, 137), (1, 210, 330), (3, 12, 12), (1, 150, 270)])
def test_ssim(cuda, n, h, w):
    """SSIM extension vs the fp64 tf.image.ssim restatement."""
    from imageenhancement_mp_b200 import data_utils as du
    g = torch.Generator().manual_seed(41)
    a = torch.rand(n, h, w, generator=g)
    b = (a + 0.1 * torch.randn(n, h, w, generator=g)).clamp(0, 1)
    from imageenhancement_mp_b200 import _lib
    lib = _lib.load()
    ref = oracle.ssim(a, b)
    try:
        for legacy in (0, 1):          # two-column / four-moment kernel (even w) and the one-column kernel
            lib.ie_ssim_tune(legacy)
            got = du.ssim(a.to(cuda), b.to(cuda)).cpu().double()
            assert torch.allclose(got, ref, atol=2e-5, rtol=1e-4), legacy
            assert torch.allclose(du.ssim(a.to(cuda), a.to(cuda)).cpu(), torch.ones(n), atol=1e-5)
    finally:
        lib.ie_ssim_tune(0)


# ------------------------------------------------------------------ preprocessing
@pytest.mark.parametrize("layer_type,color,T", [("singlestd", False, 3), ("dualparams", True, 3), ("empty", False, 3),
                                                ("singlestd", False, 4), ("dualparams", False, 8), ("empty", False, 4)])
def test_preprocess(cuda, layer_type, color, T):
    """T = 3 / colour: one-pixel-per-thread kernel; grey T in {4, 8}: the 4-pixels-per-thread kernel."""
    from oracle import preprocess as opre
    from imageenhancement_mp_b200 import data_utils as du
    params = dict(synth.DEFAULT_PARAMS, height=24, width=32, BURST_LENGTH=T, layer_type=layer_type)
    C = 3 if color else 1
    g = torch.Generator().manual_seed(51)
    N, up, jit, sj = 2, 4, 16, 2
    hs, ws = 24 * up + 2 * jit * up + 7, 32 * up + 2 * jit * up + 5
    xs, ts, orgs, wl, sr, ss, nr, ns = [], [], [], [], [], [], [], []
    src = torch.randint(0, 256, (N, hs, ws, C), generator=g, dtype=torch.uint8)
    for n in range(N):
        draws = {
            "crop0": (int(torch.randint(0, 8, (1,), generator=g)), int(torch.randint(0, 6, (1,), generator=g))),
            "use_big": [bool(torch.rand(1, generator=g) < 0.5) for _ in range(T - 1)],
            "white_level": float(10 ** (torch.rand(1, generator=g) - 1)),
            "sig_read": float(10 ** (torch.rand(1, generator=g) * 1.5 - 3)),
            "sig_shot": float(10 ** (torch.rand(1, generator=g) - 2)),
            "n_read": torch.randn(24, 32, T, generator=g),
            "n_shot": torch.randn(24, 32, T, generator=g),
        }
        draws["frame_off"] = []
        for k in range(T - 1):
            lim = 2 * jit * up if draws["use_big"][k] else 2 * sj * up
            draws["frame_off"].append((int(torch.randint(0, lim + 1, (1,), generator=g)),
                                       int(torch.randint(0, lim + 1, (1,), generator=g))))
        x, t = opre.preprocess_image(src[n], params, draws, dtype=torch.float64)
        xs.append(x); ts.append(t)
        orgs.append(opre.frame_origins(params, draws))
        wl.append(draws["white_level"]); sr.append(draws["sig_read"]); ss.append(draws["sig_shot"])
        nr.append(draws["n_read"]); ns.append(draws["n_shot"])
    f = lambda v: torch.tensor(v, dtype=torch.float32, device=cuda)
    gx, gt = du.preprocess_image(src.to(cuda), torch.tensor(orgs, dtype=torch.int32, device=cuda), params,
                                 f(wl), f(sr), f(ss), torch.stack(nr).to(cuda), torch.stack(ns).to(cuda))
    assert torch.allclose(gx.cpu().double(), torch.stack(xs), atol=2e-6, rtol=2e-5)
    assert torch.allclose(gt.cpu().double(), torch.stack(ts), atol=2e-6, rtol=2e-5)


@pytest.mark.parametrize("T,layer_type,h,w", [(4, "singlestd", 20, 28), (8, "dualparams", 9, 12), (4, "empty", 33, 260)])
def test_preprocess_quad_kernel_vs_general_kernel(cuda, T, layer_type, h, w):
    """The 4-px-per-thread kernel against the one-pixel-per-thread kernel (itself checked against the oracle above)
    on crops that LEAVE the source on every side (zero padding, data_utils.py:436-438) and on odd source pitches:
    truth within the rounding of a re-ordered 16-term sum, noisy frames within the MUFU.SQRT rounding."""
    from imageenhancement_mp_b200 import _lib, data_utils as du
    params = dict(synth.DEFAULT_PARAMS, height=h, width=w, BURST_LENGTH=T, layer_type=layer_type)
    N, up = 3, 4
    g = torch.Generator().manual_seed(h * w + T)
    hs, ws = h * up + 13, w * up + 7
    src = torch.randint(0, 256, (N, hs, ws, 1), generator=g, dtype=torch.uint8).to(cuda)
    org = torch.randint(-9, 21, (N, T, 2), generator=g, dtype=torch.int32)
    org[0, 0] = torch.tensor([0, 0]); org[1, 1] = torch.tensor([13, 7]); org[2, 0] = torch.tensor([-3, 18])
    org = org.to(cuda)
    wl = torch.tensor([0.3, 0.7, 1.0], device=cuda)
    sr = torch.tensor([0.01, 0.003, 0.02], device=cuda)
    ss = torch.tensor([0.05, 0.02, 0.09], device=cuda)
    nr = torch.randn(N, h, w, T, generator=g).to(cuda)
    ns = torch.randn(N, h, w, T, generator=g).to(cuda)
    lib = _lib.load()
    try:
        lib.ie_preprocess_tune(1)
        x_ref, t_ref = du.preprocess_image(src, org, params, wl, sr, ss, nr, ns)
        xr_ref, _ = du.preprocess_image(src, org, params, wl, sr, ss, seed=77)
    finally:
        lib.ie_preprocess_tune(0)
    x_new, t_new = du.preprocess_image(src, org, params, wl, sr, ss, nr, ns)
    xr_new, tr_new = du.preprocess_image(src, org, params, wl, sr, ss, seed=77)
    assert torch.equal(t_new, tr_new)
    assert torch.allclose(t_new, t_ref, atol=1e-7, rtol=2e-6)       # clipped windows are summed in a different order
    assert torch.allclose(x_new, x_ref, atol=1e-6, rtol=1e-5)
    assert torch.allclose(xr_new, xr_ref, atol=1e-6, rtol=1e-5)


def test_preprocess_device_noise_matches_numpy_philox(cuda):
    """ie_preprocess_u8_rng: the normals drawn on the device (Philox4x32-10 + Box-Muller) are the ones the numpy
    restatement produces, element by element - so the noisy burst equals the explicit-noise kernel fed with them."""
    from oracle import preprocess as opre
    from imageenhancement_mp_b200 import data_utils as du
    params = dict(synth.DEFAULT_PARAMS, height=20, width=28, BURST_LENGTH=4, layer_type="singlestd")
    N, T, up = 3, 4, 4
    g = torch.Generator().manual_seed(9)
    hs, ws = 20 * up + 40, 28 * up + 24
    src = torch.randint(0, 256, (N, hs, ws, 1), generator=g, dtype=torch.uint8).to(cuda)
    org = torch.randint(0, 20, (N, T, 2), generator=g, dtype=torch.int32).to(cuda)
    wl = torch.tensor([0.3, 0.7, 1.0], device=cuda)
    sr = torch.tensor([0.01, 0.003, 0.02], device=cuda)
    ss = torch.tensor([0.05, 0.02, 0.09], device=cuda)
    seed = 0x1234_5678_9abc
    x_rng, t_rng = du.preprocess_image(src, org, params, wl, sr, ss, seed=seed)
    zs, zr = opre.philox_normals(seed, N * 20 * 28 * T)
    n_shot = torch.from_numpy(zs).view(N, 20, 28, T).to(cuda)
    n_read = torch.from_numpy(zr).view(N, 20, 28, T).to(cuda)
    x_exp, t_exp = du.preprocess_image(src, org, params, wl, sr, ss, n_read=n_read, n_shot=n_shot)
    assert torch.equal(t_rng, t_exp)
    assert torch.allclose(x_rng, x_exp, atol=2e-6, rtol=1e-5)
    clean, _ = du.preprocess_image(src, org, params, wl, sr, ss)
    assert float((x_rng[..., :T] - clean[..., :T]).abs().max()) > 1e-3          # noise really was added


@pytest.mark.parametrize("n,h,w,T", [(3, 40, 56, 4), (2, 100, 100, 4), (1, 33, 47, 8), (2, 32, 32, 2)])
def test_fused_metrics_emit_the_ssim_crops(cuda, n, h, w, T):
    """ie_eval_metrics_crops_f32: same sums as ie_eval_metrics_f32, and the two by-product crops equal
    invert_preproc(deblurred) / invert_preproc(gt) (eval.py:146-149) - so the SSIM extension needs no extra passes."""
    from imageenhancement_mp_b200 import data_utils as du
    g = torch.Generator().manual_seed(21)
    recon = torch.rand(n, h, w, T + 1, generator=g).to(cuda)
    burst = torch.rand(n, h, w, T + 1, generator=g).to(cuda)
    truth = torch.rand(n, h, w, 2, generator=g)
    truth[..., 1] = torch.rand(n, 1, 1, generator=g) * 0.9 + 0.1
    truth = truth.to(cuda)
    wl = du.white_level_of(truth)
    plain = du.eval_metric_sums(recon, burst, truth, T, white_noise=wl)
    sums, db, gt = du.eval_metric_sums(recon, burst, truth, T, white_noise=wl, want_crops=True)
    assert torch.allclose(sums, plain, rtol=1e-12, atol=0)
    ref_db, ref_gt = du.invert_preproc(recon[..., 0], wl), du.invert_preproc(truth[..., 0], wl)
    assert torch.allclose(db, ref_db, atol=2e-6) and torch.allclose(gt, ref_gt, atol=2e-6)
    s2, ss = du.eval_metric_sums_with_ssim(recon, burst, truth, T, white_noise=wl)
    assert torch.allclose(ss, du.ssim_deblur_sums(recon, truth, white_noise=wl), rtol=1e-5)
