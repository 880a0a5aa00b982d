"""GPU parity of the tcgen05 implicit-GEMM convolution (csrc/conv_tcgen05.cu) through the C ABI.

Reference = torch fp32 conv on the CPU over the SAME bf16-rounded inputs and weights (the
kernel accumulates in fp32; only the bf16 output rounding differs), plus the naive CUDA-core
validation kernel at sizes where the CPU is slow.  Tolerance: bf16 output rounding
(2^-8 relative) + fp32 accumulation-order noise.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def to_raster(x):
    """fp32 NHWC (cuda) -> ops.Raster with a zero border."""
    from imageenhancement_mp_b200 import ops
    n, h, w, c = x.shape
    data = F.pad(x, (0, 0, 0, 1, 1, 0)).reshape(-1, c).to(torch.bfloat16).contiguous()   # zero row above, zero pixel right
    return ops.Raster(data, n, h, w)


def to_dense(x):
    """fp32 NHWC (cuda) -> ops.Raster in the dense NHWC layout (no border)."""
    from imageenhancement_mp_b200 import ops
    n, h, w, c = x.shape
    return ops.Raster(x.reshape(-1, c).to(torch.bfloat16).contiguous(), n, h, w, 0)


def ref_conv(x, w, b, k, relu=True):
    """x NHWC fp32 (already bf16-representable), w HWIO; 3x3 same / 2x2 valid / 1x1."""
    pad = 1 if k == 3 else 0
    y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=pad)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def make_case(n, h, w, cin, cout, k, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = bf16_round(torch.randn(n, h, w, cin, generator=g))
    wt = bf16_round(torch.randn(k, k, cin, cout, generator=g) * scale / (k * k * cin) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    return x, wt, b


def assert_close_bf16(got, ref, what=""):
    err = (got - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-3 * ref.abs().max().clamp(min=1e-3)
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} / {bad.numel()} elements off, max err {float(err.max()):.4g} " \
                          f"(ref max {float(ref.abs().max()):.4g})"


CASES = [
    # n, h, w, cin, cout, k
    (2, 16, 16, 64, 64, 3),        # one K block per tap, N tile 64
    (1, 8, 8, 128, 128, 3),        # two K blocks per tap, N tile 128
    (1, 8, 8, 256, 512, 3),        # N tile 256, two N tiles
    (2, 2, 2, 2048, 512, 3),       # basis-branch shape: 32 raster rows (< one 128-row TMA box), K = 18432
    (1, 16, 16, 128, 128, 2),      # the 2x2 'valid' conv (model_library.py:364)
    (2, 24, 40, 64, 64, 1),        # 1x1 (first layer over im2col rows), non-square
    (3, 26, 26, 1024, 1024, 3),    # the dominant quarter-resolution layer shape
    (5, 100, 100, 64, 128, 3),     # cout = 128 with >= 2 x SMs M tiles: the transposed kernel (51005 rows: ragged last item)
    (4, 104, 104, 128, 128, 3),    # same with two K blocks per tap (345 tiles)
]


@pytest.mark.parametrize("n,h,w,cin,cout,k", CASES)
def test_conv_bf16_vs_cpu(cuda, n, h, w, cin, cout, k):
    from imageenhancement_mp_b200 import ops
    x, wt, b = make_case(n, h, w, cin, cout, k)
    ref = bf16_round(ref_conv(x, wt, b, k))
    src = to_raster(x.to(cuda))
    dst = ops.new_raster(n, h, w, cout, cuda)
    dst.data.fill_(float("nan"))
    wp = ops.pack_conv_weights(wt.to(cuda))
    valid = (h - 1, w - 1) if k == 2 else None
    ops.conv2d(src.slice(), wp, b.to(cuda), dst.slice(), k=k, valid=valid)
    torch.cuda.synchronize()
    got = ops.raster_to_nhwc(dst.slice()).cpu()
    if k == 2:
        got = got[:, :h - 1, :w - 1]
    assert_close_bf16(got, ref, f"conv {k}x{k} {cin}->{cout}")
    # the zero border (and everything outside the valid extent) must be exactly zero
    full = dst.data.float().view(n, h + 1, w + 1, cout)
    hv, wv = (h - 1, w - 1) if k == 2 else (h, w)
    mask = torch.ones(h + 1, w + 1, dtype=torch.bool, device=cuda)
    mask[1:hv + 1, 0:wv] = False
    assert torch.all(full[:, mask] == 0), "border rows were not zeroed"


DENSE_CASES = [
    # n, h, w, cin, cout, k      dense NHWC tensors, taps fetched by TMA im2col (csrc/conv_tcgen05.cu, IM2COL)
    (3, 26, 26, 1024, 1024, 3),    # the dominant quarter-resolution layer: 2028 pixels = 15.8 tiles (ragged last tile)
    (2, 13, 13, 128, 256, 3),      # 338 pixels: tiles cross image rows AND images
    (5, 2, 2, 2048, 512, 3),       # basis branch: 20 pixels, every tap of every pixel touches the padding
    (7, 1, 1, 1024, 512, 3),       # 1x1 images: only the centre tap is inside
    (2, 4, 8, 64, 64, 3),          # narrow layer through the streaming kernel
    (1, 24, 40, 192, 128, 1),      # 1x1 convolution (pad 0)
    (2, 9, 31, 64, 128, 3),        # odd sizes, 558 pixels
    (60, 26, 26, 128, 128, 3),     # cout = 128, 40560 pixels: the transposed kernel through 256-pixel im2col boxes (ragged last item)
]


@pytest.mark.parametrize("n,h,w,cin,cout,k", DENSE_CASES)
def test_conv_dense_im2col_vs_cpu(cuda, n, h, w, cin, cout, k):
    """Dense NHWC in / out: the zero padding comes from the im2col tensor map's pixel box, tiles are 128 consecutive
    pixels of the [n][h][w] index space, no border row is computed.  Also read from / written to channel slices."""
    from imageenhancement_mp_b200 import ops
    x, wt, b = make_case(n, h, w, cin, cout, k)
    ref = bf16_round(ref_conv(x, wt, b, k))
    # the input is channels [64 : 64 + cin] of a wider tensor, the output channels [64 : 64 + cout] of another
    wide = torch.cat([torch.full((n, h, w, 64), 9.0), x, torch.full((n, h, w, 64), -9.0)], dim=-1)
    src = to_dense(wide.to(cuda))
    dst = ops.new_raster(n, h, w, cout + 128, cuda, dense=True)
    assert dst.rows == n * h * w and dst.dense
    dst.data.fill_(7.0)
    wp = ops.pack_conv_weights(wt.to(cuda))
    ops.conv2d(src.slice(64, cin), wp, b.to(cuda), dst.slice(64, cout), k=k)
    torch.cuda.synchronize()
    got = ops.raster_to_nhwc(dst.slice(64, cout)).cpu()
    assert_close_bf16(got, ref, f"dense conv {k}x{k} {cin}->{cout}")
    assert torch.all(dst.data[:, :64] == 7.0) and torch.all(dst.data[:, 64 + cout:] == 7.0)
    # same answer as the naive validation kernel on the same dense tensors, and as the raster path
    d2 = ops.new_raster(n, h, w, cout, cuda, dense=True)
    ops.conv2d(src.slice(64, cin), wp, b.to(cuda), d2.slice(), k=k, fn="ie_debug_conv2d_naive")
    assert_close_bf16(got, ops.raster_to_nhwc(d2.slice()).cpu(), "dense conv vs naive")
    d3 = ops.new_raster(n, h, w, cout, cuda)
    ops.conv2d(to_raster(x.to(cuda)).slice(), wp, b.to(cuda), d3.slice(), k=k)
    # (bit-identical when the raster path also takes the streaming kernel; the wide-N / resident kernels add the taps
    # in a different order)
    assert_close_bf16(ops.raster_to_nhwc(d3.slice()).cpu(), got, "dense vs raster layout")


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 8, 8, 1024, 1024), (1, 4, 4, 2048, 512), (3, 2, 2, 2048, 512),
                                            (1, 16, 16, 128, 128), (2, 8, 8, 512, 64),
                                            # a persistent grid with a partial last wave: only its tiles are split
                                            (30, 26, 26, 256, 512),      # 159 M tiles x 2 N tiles = 318 = 2 waves + 22
                                            (64, 13, 13, 128, 256)])     # 85 (dense) / 98 (raster) M tiles
def test_conv_split_k_for_tiny_m(cuda, dense, n, h, w, cin, cout):
    """Tiny M (eval.py's default call: one 32 x 32 patch): the K loop of a tile is split over many CTAs, the fp32
    partial sums are reduced in a fixed order by a second kernel.  Same numbers as the unsplit kernel (to fp32
    summation order), deterministic from run to run, channel slices and borders intact."""
    from imageenhancement_mp_b200 import _lib, ops
    lib = _lib.load()
    x, wt, b = make_case(n, h, w, cin, cout, 3)
    ref = bf16_round(ref_conv(x, wt, b, 3))
    src = (to_dense if dense else to_raster)(x.to(cuda))
    wp = ops.pack_conv_weights(wt.to(cuda))
    outs = []
    try:
        for flags in (0, 0, 1 << 9):                          # split-K (auto) twice, then forced off
            lib.ie_conv_set_mode(0, flags)                    # mode 0: the streaming kernel
            dst = ops.new_raster(n, h, w, cout + 64, cuda, dense=dense)
            dst.data.fill_(3.0)
            ops.conv2d(src.slice(), wp, b.to(cuda), dst.slice(64, cout), workspace=ops.conv_workspace(cuda))
            outs.append((ops.raster_to_nhwc(dst.slice(64, cout)).cpu(), dst))
    finally:
        lib.ie_conv_set_mode(-1, 0)
    assert torch.equal(outs[0][0], outs[1][0]), "split-K must be deterministic"
    assert_close_bf16(outs[0][0], ref, "split-K conv vs CPU")
    assert_close_bf16(outs[0][0], outs[2][0], "split-K vs unsplit")
    for _, dst in outs:
        assert torch.all(dst.data[:, :64] == 3.0)
        if not dense:
            full = dst.data[:, 64:].float().view(n, h + 1, w + 1, cout)
            assert torch.all(full[:, 0] == 0) and torch.all(full[:, :, -1] == 0)


def test_conv_dense_rejects_what_it_does_not_support(cuda):
    from imageenhancement_mp_b200 import ops, ImgEnhError
    x, wt, b = make_case(1, 8, 8, 64, 64, 3)
    wp = ops.pack_conv_weights(wt.to(cuda))
    with pytest.raises(ImgEnhError):                       # mixed layouts
        ops.conv2d(to_dense(x.to(cuda)).slice(), wp, b.to(cuda), ops.new_raster(1, 8, 8, 64, cuda).slice())
    x2, wt2, b2 = make_case(1, 8, 8, 64, 64, 2)
    with pytest.raises(ImgEnhError):                       # 2x2 'valid' is a raster kernel
        ops.conv2d(to_dense(x2.to(cuda)).slice(), ops.pack_conv_weights(wt2.to(cuda)), b2.to(cuda),
                   ops.new_raster(1, 8, 8, 64, cuda, dense=True).slice(), k=2, valid=(7, 7))


@pytest.mark.parametrize("mode", [0, 1, 2], ids=["stream", "resident", "wideN"])
@pytest.mark.parametrize("n,h,w,cin", [(2, 16, 16, 64), (2, 24, 40, 128), (3, 40, 44, 64), (1, 20, 20, 640),
                                       (5, 8, 8, 64)])
def test_conv_narrow_flavours(cuda, mode, n, h, w, cin):
    """Every main-loop flavour that can run a 3x3 cout=64 layer gives the same answer: streaming, resident
    weights with row-shifted descriptors, and the wide-N kernel (horizontal taps as accumulator column groups,
    shift applied in the epilogue; tiles overlap by two rows, so multi-tile and multi-image cases matter)."""
    from imageenhancement_mp_b200 import ops, _lib, ImgEnhError
    lib = _lib.load()
    x, wt, b = make_case(n, h, w, cin, 64, 3, seed=21)
    ref = bf16_round(ref_conv(x, wt, b, 3))
    src = to_raster(x.to(cuda))
    wp = ops.pack_conv_weights(wt.to(cuda))
    dst = ops.new_raster(n, h, w, 64, cuda)
    dst.data.fill_(float("nan"))
    lib.ie_conv_set_mode(mode, 0)
    try:
        try:
            ops.conv2d(src.slice(), wp, b.to(cuda), dst.slice())
        except ImgEnhError:
            if mode == 1 and cin > 128:
                pytest.skip("weights do not fit in shared memory for the resident flavour")
            raise
        torch.cuda.synchronize()
    finally:
        lib.ie_conv_set_mode(-1, 0)
    assert_close_bf16(ops.raster_to_nhwc(dst.slice()).cpu(), ref, f"flavour {mode} {cin}->64")
    full = dst.data.float().view(n, h + 1, w + 1, 64)
    mask = torch.ones(h + 1, w + 1, dtype=torch.bool, device=cuda)
    mask[1:h + 1, 0:w] = False
    assert torch.all(full[:, mask] == 0), "border rows were not zeroed"


def test_conv_slices_and_concat(cuda):
    """Input read from a channel window, output written into a channel window (concat by slices)."""
    from imageenhancement_mp_b200 import ops
    n, h, w = 2, 12, 20
    x, wt, b = make_case(n, h, w, 128, 64, 3, seed=3)
    g = torch.Generator().manual_seed(9)
    junk_in = bf16_round(torch.randn(n, h, w, 64, generator=g))
    wide = torch.cat([junk_in, x], dim=-1)                       # x lives at channels [64,192)
    src = to_raster(wide.to(cuda))
    dst = ops.new_raster(n, h, w, 256, cuda)
    dst.data.fill_(7.0)
    wp = ops.pack_conv_weights(wt.to(cuda))
    ops.conv2d(src.slice(64, 128), wp, b.to(cuda), dst.slice(128, 64))
    torch.cuda.synchronize()
    got = ops.raster_to_nhwc(dst.slice(128, 64)).cpu()
    assert_close_bf16(got, bf16_round(ref_conv(x, wt, b, 3)), "sliced conv")
    other = torch.cat([dst.data[:, :128], dst.data[:, 192:]], dim=1)
    assert torch.all(other == 7.0), "conv wrote outside its output slice"


def test_conv_many_tiles_vs_naive(cuda):
    """More tiles than SMs: exercises the persistent loop, both TMEM buffers and all phase bits."""
    from imageenhancement_mp_b200 import ops
    n, h, w = 6, 104, 104
    x, wt, b = make_case(n, h, w, 64, 64, 3, seed=5)
    src = to_raster(x.to(cuda))
    wp = ops.pack_conv_weights(wt.to(cuda))
    bias = b.to(cuda)
    d1 = ops.new_raster(n, h, w, 64, cuda)
    d2 = ops.new_raster(n, h, w, 64, cuda)
    ops.conv2d(src.slice(), wp, bias, d1.slice())
    ops.conv2d(src.slice(), wp, bias, d2.slice(), fn="ie_debug_conv2d_naive")
    torch.cuda.synchronize()
    a, r = d1.data.float(), d2.data.float()
    assert_close_bf16(a.cpu(), r.cpu(), "tcgen05 vs naive, 64->64 @104x104")


def test_conv_deep_k_vs_naive(cuda):
    from imageenhancement_mp_b200 import ops
    n, h, w = 8, 26, 26
    x, wt, b = make_case(n, h, w, 2048, 512, 3, seed=6)
    src = to_raster(x.to(cuda))
    wp = ops.pack_conv_weights(wt.to(cuda))
    bias = b.to(cuda)
    d1 = ops.new_raster(n, h, w, 512, cuda)
    d2 = ops.new_raster(n, h, w, 512, cuda)
    ops.conv2d(src.slice(), wp, bias, d1.slice())
    ops.conv2d(src.slice(), wp, bias, d2.slice(), fn="ie_debug_conv2d_naive")
    torch.cuda.synchronize()
    assert_close_bf16(d1.data.float().cpu(), d2.data.float().cpu(), "tcgen05 vs naive, 2048->512 @26x26")


@pytest.mark.parametrize("cout,softmax", [(10, True), (40, False), (16, True), (64, False), (80, False), (90, True),
                                          (12, False), (7, True),
                                          # Basis_kpn's layer3_3 with the remote/ settings: T*B > 256 channels run as
                                          # chunks of 256 writing channel slices of one NHWC tensor
                                          (256, False), (360, False), (400, False), (720, False)])
def test_conv_f32_heads(cuda, cout, softmax):
    """The two small fp32 epilogues: coef conv + softmax (model_library.py:405-406) and layer3_3."""
    from imageenhancement_mp_b200 import ops
    from imageenhancement_mp_b200._lib import IE_EPI_F32_NHWC, IE_EPI_F32_SOFTMAX
    n, h, w, cin = (2, 16, 16, 64 if softmax else 128) if cout != 7 else (3, 40, 44, 64)   # 7: many overlapping tiles
    x, wt, b = make_case(n, h, w, cin, cout, 3, seed=11, scale=3.0)
    ref = ref_conv(x, wt, b, 3)
    src = to_raster(x.to(cuda))
    wp = ops.pack_conv_weights(wt.to(cuda), IE_EPI_F32_SOFTMAX if softmax else IE_EPI_F32_NHWC)
    if softmax:
        y, logits = ops.conv2d_f32(src.slice(), wp, b.to(cuda), cout, softmax=True, want_logits=True)
        torch.cuda.synchronize()
        assert torch.allclose(logits.cpu(), ref, atol=2e-4, rtol=1e-4)
        assert torch.allclose(y.cpu(), torch.softmax(ref, -1), atol=1e-5, rtol=1e-4)
    else:
        y = ops.conv2d_f32(src.slice(), wp, b.to(cuda), cout, valid=(15, 15))
        torch.cuda.synchronize()
        assert torch.allclose(y.cpu(), ref[:, :15, :15], atol=2e-4, rtol=1e-4)
        if cout > 256:       # the naive validation kernel takes the same chunked route
            y2 = ops.conv2d_f32(src.slice(), wp, b.to(cuda), cout, valid=(15, 15), fn="ie_debug_conv2d_naive")
            assert torch.allclose(y2, y, atol=2e-4, rtol=1e-4)
            with pytest.raises(Exception):
                ops.conv2d_f32(src.slice(), wp, b.to(cuda), cout, softmax=True)


def test_conv_rejects_bad_descriptors(cuda):
    from imageenhancement_mp_b200 import ops, ImgEnhError
    src = ops.new_raster(1, 8, 8, 96, cuda)          # 96 is not a multiple of 64
    dst = ops.new_raster(1, 8, 8, 64, cuda)
    wp = torch.zeros(64, 9 * 96, dtype=torch.bfloat16, device=cuda)
    with pytest.raises(ImgEnhError):
        ops.conv2d(src.slice(), wp, None, dst.slice())


def test_conv_random_shape_sweep_vs_naive(cuda):
    """Randomised layer shapes through every dispatch path (streaming at three N tiles, resident, wide-N resident /
    streaming, adaptive N tile on small rasters) against the CUDA-core validation kernel."""
    import random
    from imageenhancement_mp_b200 import ops
    rnd = random.Random(7)
    for trial in range(24):
        cin = rnd.choice([64, 64, 128, 192, 320, 640])
        cout = rnd.choice([64, 64, 128, 192, 256, 512])
        k = rnd.choice([3, 3, 3, 1, 2])
        n = rnd.choice([1, 2, 5])
        h, w = rnd.randint(2, 40), rnd.randint(2, 44)
        if k == 2 and (h < 2 or w < 2):
            continue
        x, wt, b = make_case(n, h, w, cin, cout, k, seed=100 + trial)
        src = to_raster(x.to(cuda))
        wp = ops.pack_conv_weights(wt.to(cuda))
        bias = b.to(cuda)
        valid = (h - 1, w - 1) if k == 2 else None
        d1 = ops.new_raster(n, h, w, cout, cuda)
        d2 = ops.new_raster(n, h, w, cout, cuda)
        d1.data.fill_(float("nan"))
        ops.conv2d(src.slice(), wp, bias, d1.slice(), k=k, valid=valid)
        ops.conv2d(src.slice(), wp, bias, d2.slice(), k=k, valid=valid, fn="ie_debug_conv2d_naive")
        torch.cuda.synchronize()
        assert not torch.isnan(d1.data).any(), f"trial {trial}: unwritten output ({n},{h},{w},{cin}->{cout},k{k})"
        assert_close_bf16(d1.data.float().cpu(), d2.data.float().cpu(), f"trial {trial}: ({n},{h},{w},{cin}->{cout},k{k})")


@pytest.mark.parametrize("dense", [False, True])
def test_cout128_flavours_are_bit_identical(cuda, dense):
    """The three ways a cout = 128 layer with a large grid can run - conv_streamT_kernel (D^T = W X^T: pixels as the
    N = 256 dimension, transposing epilogue; the default), conv_stream_kernel<.., 2> (two 128-row M tiles per work item
    sharing each weight tile; flag 1 << 17) and one tile per item (flag 1 << 16) - accumulate every output in the same
    order: same bits, borders and masked rows included."""
    from imageenhancement_mp_b200 import _lib, ops
    lib = _lib.load()
    n, h, w, cin, cout = (60, 26, 26, 192, 128) if dense else (5, 100, 100, 128, 128)
    x, wt, b = make_case(n, h, w, cin, cout, 3)
    src = to_dense(x.to(cuda)) if dense else to_raster(x.to(cuda))
    wp = ops.pack_conv_weights(wt.to(cuda))
    outs = []
    try:
        for flags in (0, 1 << 17, 1 << 16):
            lib.ie_conv_set_mode(-1, flags)
            dst = ops.new_raster(n, h, w, 2 * cout, cuda, dense=dense)          # written as the upper channel slice
            dst.data.fill_(float("nan"))
            ops.conv2d(src.slice(), wp, b.to(cuda), dst.slice(cout, cout), k=3)
            torch.cuda.synchronize()
            assert bool(torch.isnan(dst.data[:, :cout].float()).all())             # wrote only its slice
            outs.append(dst.data[:, cout:].clone())
    finally:
        lib.ie_conv_set_mode(-1, 0)
    assert not torch.isnan(outs[0].float()).any()
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(outs[0], outs[2])
