"""Python-level wrappers of the C ABI, operating on torch CUDA tensors.

A ``Raster`` is the library's activation container: bf16 ``[rows, pitch]`` in one of two layouts
(include/imgenh_b200.h).  ``b = 1``: ``rows = n*(h+1)*(w+1)``, every image preceded by one zero row and every image
row followed by one zero pixel (a shared one-pixel border: a k x k tap is a constant row shift).  ``b = 0``
(dense NHWC): ``rows = n*h*w``; the convolution fetches its taps with TMA im2col tensor maps - the layout of the
1/4-resolution and smaller tensors, where border rows would be 8 - 125 % of the GEMM.  ``Slice`` is a channel window
of a raster - how the reference's ``layers.concatenate([up, skip])``
(/root/reference/model_library.py:96) is expressed without a copy.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import (ConvDesc, IE_EPI_BF16_RASTER, IE_EPI_F32_NHWC, IE_EPI_F32_SOFTMAX, IE_LAYOUT_X_DENSE, IE_LAYOUT_Y_DENSE,
                   call, ptr, stream)


# bench.py sets this to a list to collect (start, end) CUDA events around every convolution launch
CONV_EVENTS = None


def _timed_conv(fn, *args):
    if CONV_EVENTS is None:
        call(fn, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(fn, *args)
    e1.record()
    CONV_EVENTS.append((e0, e1))


@dataclass
class Raster:
    data: torch.Tensor      # bf16 [rows, pitch]
    n: int
    h: int
    w: int
    b: int = 1              # border: 1 = shared-border raster, 0 = dense NHWC

    @property
    def pitch(self):
        return self.data.shape[1]

    @property
    def rows(self):
        return self.n * (self.h + self.b) * (self.w + self.b)

    @property
    def dense(self):
        return self.b == 0

    def slice(self, coff=0, c=None):
        return Slice(self, coff, self.pitch - coff if c is None else c)


@dataclass
class Slice:
    r: Raster
    coff: int
    c: int


def new_raster(n, h, w, c, device, dense=False):
    # torch.empty: every row (borders included) is written by the producing kernel
    b = 0 if dense else 1
    return Raster(torch.empty(n * (h + b) * (w + b), c, dtype=torch.bfloat16, device=device), n, h, w, b)


def _layout(src=None, dst=None):
    return (IE_LAYOUT_X_DENSE if (src is not None and src.r.dense) else 0) | \
           (IE_LAYOUT_Y_DENSE if (dst is not None and dst.r.dense) else 0)


def conv_n_tile(cout, epilogue):
    """Mirror of choose_n_tile() in csrc/conv_tcgen05.cu: rows the packed weights must be padded to."""
    if epilogue != IE_EPI_BF16_RASTER:
        return ((cout + 15) // 16) * 16
    return 256 if cout >= 256 else (128 if cout > 64 else 64)


def pack_conv_weights(kernel_hwio, epilogue=IE_EPI_BF16_RASTER, ktot_pad=None):
    """HWIO fp32 (CUDA) -> bf16 [cout_pad, ktot_pad], K index (i*kw+j)*cin + c."""
    _lib.require_cuda(kernel_hwio)
    kh, kw, cin, cout = kernel_hwio.shape
    ktot = kh * kw * cin
    ktot_pad = ktot if ktot_pad is None else ktot_pad
    nt = conv_n_tile(cout, epilogue)
    cout_pad = -(-cout // nt) * nt
    out = torch.zeros(cout_pad, ktot_pad, dtype=torch.bfloat16, device=kernel_hwio.device)
    src = kernel_hwio.contiguous().float()
    call("ie_pack_conv_weights", ptr(src), kh, kw, cin, cout, ktot_pad, ptr(out), stream())
    return out


def im2col_width(c):
    """Channels of the im2col raster of a c-channel input: 9*c rounded up to a multiple of 64."""
    return -(-9 * c // 64) * 64


def pack_input_im2col3x3(x, out=None):
    """fp32 NHWC [n,hs,ws,c] -> Raster(n,h,w) with im2col_width(c) channels holding the 3x3xc neighbourhoods.

    ``out`` may be larger than the source (h >= hs, w >= ws): the source is zero-padded at the bottom/right."""
    _lib.require_cuda(x)
    n, hs, ws, c = x.shape
    x = x.contiguous()
    if out is None:
        out = new_raster(n, hs, ws, im2col_width(c), x.device)
    assert out.pitch == im2col_width(c) and out.n == n and out.h >= hs and out.w >= ws
    call("ie_pack_input_im2col3x3", ptr(x), n, hs, ws, c, out.h, out.w, ptr(out.data), stream())
    return out


FIRST_LAYER_FUSED_CHANNELS = (3, 5, 10)
F32_HEAD_MAX_COUT = 256        # output channels one launch of an fp32-epilogue convolution handles (one N tile)


def conv_first_layer(x, w_packed, bias, dst: Slice, relu=True):
    """Conv2D(64, 3, 'same') on the raw fp32 NHWC input with the im2col fused into the kernel (c in 3, 5, 10)."""
    _lib.require_cuda(x)
    n, hs, ws, c = x.shape
    r = dst.r
    assert x.is_contiguous() and r.n == n and r.h >= hs and r.w >= ws and w_packed.shape == (64, im2col_width(c))
    e0 = e1 = None
    if CONV_EVENTS is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    call("ie_conv_first_layer_f32", ptr(x), n, hs, ws, c, r.h, r.w, ptr(w_packed), ptr(bias), dst.c, int(relu),
         ptr(r.data), r.pitch, dst.coff, stream())
    if e0 is not None:
        e1.record()
        CONV_EVENTS.append((e0, e1))


def _desc(src: Slice, kh, kw, cout, relu, epilogue, dst: Slice | None, valid=None):
    r = src.r
    hv, wv = (r.h, r.w) if valid is None else valid
    d = ConvDesc()
    d.n_img, d.h, d.w, d.hv, d.wv = r.n, r.h, r.w, hv, wv
    d.kh, d.kw = kh, kw
    d.cin, d.x_pitch, d.x_coff = src.c, r.pitch, src.coff
    d.cout = cout
    d.y_pitch, d.y_coff = (dst.r.pitch, dst.coff) if dst is not None else (0, 0)
    d.relu, d.epilogue = int(relu), epilogue
    d.dense = int(r.dense)
    if dst is not None and dst.r.dense != r.dense:
        raise _lib.ImgEnhError("conv: input and output must use the same layout (both dense or both shared-border rasters)")
    return d


_WORKSPACES = {}
SPLITK_WORKSPACE_BYTES = 32 << 20


def conv_workspace(device):
    """Per-(device, stream) split-K scratch for convolutions with very few output tiles (see ie_conv2d_nhwc_bf16).  One
    buffer per stream: the two kernels that use it run back to back on that stream."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None:
        ws = _WORKSPACES[key] = torch.empty(SPLITK_WORKSPACE_BYTES, dtype=torch.uint8, device=device)
    return ws


def conv2d(src: Slice, w_packed, bias, dst: Slice, k=3, relu=True, valid=None, fn="ie_conv2d_nhwc_bf16", workspace=None):
    """Conv2D(k, relu) from a raster slice into a raster slice (bf16 epilogue).  ``workspace``: uint8 CUDA tensor for
    split-K (tiny inputs); None = never split."""
    if dst.r.dense != src.r.dense:
        raise _lib.ImgEnhError("conv: input and output must use the same layout (both dense or both shared-border rasters)")
    assert dst.r.rows == src.r.rows, "conv output must share the input raster geometry"
    d = _desc(src, k, k, dst.c, relu, IE_EPI_BF16_RASTER, dst, valid)
    _timed_conv(fn, C.byref(d), ptr(src.r.data), ptr(w_packed), ptr(bias), ptr(dst.r.data), None, None,
                ptr(workspace), workspace.numel() if workspace is not None else 0, stream())


def conv2d_f32(src: Slice, w_packed, bias, cout, k=3, relu=True, valid=None, softmax=False, want_logits=False,
               fn="ie_conv2d_nhwc_bf16"):
    """Conv2D(k, relu) with an fp32 NHWC output [n,hv,wv,cout] (optionally softmax over channels)."""
    r = src.r
    hv, wv = (r.h, r.w) if valid is None else valid
    epi = IE_EPI_F32_SOFTMAX if softmax else IE_EPI_F32_NHWC
    y = torch.empty(r.n, hv, wv, cout, dtype=torch.float32, device=r.data.device)
    if cout > F32_HEAD_MAX_COUT:
        # Basis_kpn's layer3_3 with the remote/ settings (T*B = 8*50 .. 8*90 output channels): chunks of 256 output
        # channels, each launch writing its channel slice of the same NHWC tensor
        if softmax:
            raise _lib.ImgEnhError(f"softmax heads support at most {F32_HEAD_MAX_COUT} channels (got {cout})")
        # equal-ish chunks that are multiples of 16: no remainder of <= 16 channels (the fp32 channel-slice epilogue
        # is built for cout > 16; e.g. T*B = 264 runs as 144 + 120, not 256 + 8)
        nchunks = -(-cout // F32_HEAD_MAX_COUT)
        step = -(-(-(-cout // nchunks)) // 16) * 16
        bounds = [min(i * step, cout) for i in range(nchunks + 1)]
        if cout - bounds[-2] <= 16 and nchunks > 1:          # last chunk still tiny: borrow 16 channels from its neighbour
            bounds[-2] -= 16
        for c0, c1 in zip(bounds[:-1], bounds[1:]):
            cc = c1 - c0
            d = _desc(src, k, k, cc, relu, epi, None, valid)
            d.y_pitch, d.y_coff = cout, c0
            _timed_conv(fn, C.byref(d), ptr(r.data), ptr(w_packed[c0:]), ptr(bias[c0:]) if bias is not None else None, None,
                        ptr(y), None, None, 0, stream())
        return y
    d = _desc(src, k, k, cout, relu, epi, None, valid)
    aux = torch.empty_like(y) if (softmax and want_logits) else None
    _timed_conv(fn, C.byref(d), ptr(r.data), ptr(w_packed), ptr(bias), None, ptr(y), ptr(aux), None, 0, stream())
    return (y, aux) if softmax else y


def maxpool2(src: Slice, dst: Slice, want_mean=False, rows=None, count=0):
    """MaxPooling2D(2,2); with ``want_mean`` also returns the per-image channel means [n, c] of the INPUT slice
    (over input rows ``rows`` = (y0, y1) and divided by ``count`` pixels when given: spatial shards)."""
    r = src.r
    assert (dst.r.n, dst.r.h, dst.r.w) == (r.n, r.h // 2, r.w // 2) and dst.c == src.c
    mean = scratch = None
    nbytes = 0
    if want_mean:
        # the statistics are reduced across blocks in a fixed order through a scratch buffer (bit-reproducible means)
        mean = torch.empty(r.n, src.c, dtype=torch.float32, device=r.data.device)
        nbytes = _lib.load().ie_maxpool2_stat_scratch_bytes(r.n, r.h, r.w, src.c, _layout(src, dst))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=r.data.device)
    y0, y1 = rows if rows is not None else (0, 0)
    call("ie_maxpool2_nhwc_bf16", ptr(r.data), r.n, r.h, r.w, src.c, r.pitch, src.coff,
         ptr(dst.r.data), dst.r.pitch, dst.coff, ptr(mean), ptr(scratch), nbytes, y0, y1, int(count), _layout(src, dst),
         stream())
    return mean


def upsample_bilinear(src: Slice, dst: Slice, scale):
    r = src.r
    assert (dst.r.n, dst.r.h, dst.r.w) == (r.n, r.h * scale, r.w * scale) and dst.c == src.c
    call("ie_upsample_bilinear_nhwc_bf16", ptr(r.data), r.n, r.h, r.w, src.c, r.pitch, src.coff, scale,
         ptr(dst.r.data), dst.r.pitch, dst.coff, _layout(src, dst), stream())


def channel_mean(src: Slice, out=None, rows=None, count=0):
    r = src.r
    if out is None:
        out = torch.empty(r.n, src.c, dtype=torch.float32, device=r.data.device)
    y0, y1 = rows if rows is not None else (0, 0)
    nbytes = _lib.load().ie_channel_mean_scratch_bytes(r.n, r.h, r.w, src.c, y0, y1)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=r.data.device)
    call("ie_channel_mean_nhwc_bf16", ptr(r.data), r.n, r.h, r.w, src.c, r.pitch, src.coff, ptr(out), ptr(scratch), nbytes,
         y0, y1, int(count), _layout(src), stream())
    return out


def broadcast_hw(vec, dst: Slice):
    assert vec.shape == (dst.r.n, dst.c) and vec.dtype == torch.float32
    call("ie_broadcast_hw_bf16", ptr(vec), dst.r.n, dst.r.h, dst.r.w, dst.c, ptr(dst.r.data), dst.r.pitch,
         dst.coff, _layout(None, dst), stream())


def raster_to_nhwc(src: Slice):
    r = src.r
    y = torch.empty(r.n, r.h, r.w, src.c, dtype=torch.float32, device=r.data.device)
    call("ie_raster_to_nhwc_f32", ptr(r.data), r.n, r.h, r.w, src.c, r.pitch, src.coff, ptr(y), _layout(src), stream())
    return y


def softmax_taps(originbasis, T, B):
    """[n,K,K,T*B] fp32 -> Bas [n,K,K,T,B]: softmax over the K*K*T taps per basis (model_library.py:436)."""
    n, K = originbasis.shape[0], originbasis.shape[1]
    ob = originbasis.contiguous()
    out = torch.empty_like(ob)
    call("ie_softmax_taps_f32", ptr(ob), n, K * K * T, B, ptr(out), stream())
    return out.view(n, K, K, T, B)


def kpn_tf32_supported(T, K, B):
    return K == 15 and B <= 128 and T <= 8


def kpn_tcgen05_supported(T, K, B):
    """ie_kpn_apply_tc (csrc/kpn_tcgen05.cu): filter synthesis as a tcgen05 GEMM + apply in the epilogue; frames in
    passes of four, bases in blocks of 64 (each further pass / block adds into the output)."""
    return K == 15 and T % 4 == 0 and T >= 4 and B >= 1


def kpn_apply(x, T, coef, bas, out=None, precision="fp32"):
    """Per-pixel filter (model_library.py:439-451).  x: fp32 NHWC whose first T channels are the burst.

    precision "fp32": CUDA-core kernel, 1e-5 of the fp64 oracle.  "tf32": tensor-core kernel (burst and basis rounded
    to TF32, fp32 accumulation; K = 15, B <= 128 in chunks of 16, T <= 8), < 1e-3 absolute on [0,1] pixels, ~2.5x faster.
    "tcgen05" (the models' default where it applies, csrc/kpn_tcgen05.cu): the filter synthesised by a tcgen05 GEMM
    (coefficients and basis - softmax outputs in [0, 1] - as fp16: the same 10-bit mantissa as TF32; fp32 accumulation,
    the burst is not rounded) and applied in its epilogue; K = 15, T % 4 == 0, any B (blocks of 64)."""
    _lib.require_cuda(x, coef, bas)
    n, h, w, pitch = x.shape
    K, B = bas.shape[1], bas.shape[-1]
    hc, wc = coef.shape[1], coef.shape[2]          # >= (h, w): the network runs at the stride-padded size
    assert x.is_contiguous() and coef.is_contiguous() and bas.is_contiguous()
    assert coef.shape[0] == n and coef.shape[3] == B and hc >= h and wc >= w and bas.shape == (n, K, K, T, B)
    assert precision in ("fp32", "tf32", "tcgen05")
    if out is None:
        out = torch.empty(n, h, w, T + 1, dtype=torch.float32, device=x.device)
    fn = {"fp32": "ie_kpn_apply_f32", "tf32": "ie_kpn_apply_tf32", "tcgen05": "ie_kpn_apply_tc"}[precision]
    call(fn, ptr(x), pitch, ptr(coef), hc, wc, ptr(bas), ptr(out), n, h, w, T, K, B, stream())
    return out


def convolve_filts(img_stack, filts, K):
    """Materialised-filter path (model_library.py:114-168): img_stack [n,h,w,T], filts [n,h,w,K,K,T] (or flattened
    [n,h,w,K*K*T]) -> [n,h,w,T+1]: channel 0 = Convolve, channels 1.. = Convolve_perlayer."""
    _lib.require_cuda(img_stack, filts)
    n, h, w, T = img_stack.shape
    x = img_stack.contiguous().float()
    f = filts.contiguous().float()
    assert f.numel() == n * h * w * K * K * T, "filts must hold K*K*T taps per pixel"
    out = torch.empty(n, h, w, T + 1, dtype=torch.float32, device=x.device)
    call("ie_convolve_filts_f32", ptr(x), T, ptr(f), ptr(out), n, h, w, T, K, stream())
    return out
