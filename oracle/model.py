"""Torch-CPU restatement of the reference networks (test infrastructure, parity unpinned).

Follows /root/reference/model_library.py:
  Downblock :65-79, Upblock :81-101, Poolskip :102-113, Convolve :114-135,
  cus_convolve :136-152, Convolve_perlayer :153-168, Basis_kpn :179-295,
  Simplemodel :306-452.

Tensors are NHWC like the reference.  Weights come as a flat dict keyed by the Keras
attribute path (``layer0``, ``down1.conv2d1`` ...), each entry ``(kernel HWIO, bias)``.
TensorFlow semantics assumed: Conv2D = cross-correlation, 'same' = symmetric zero pad for
odd kernels; MaxPooling2D(2,2,'valid') floors; UpSampling2D(interpolation='bilinear') is the
TF2 half-pixel-centre resize (== F.interpolate(align_corners=False)); tf.nn.softmax.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- layers
def conv2d_relu(x, wb, padding):
    """layers.Conv2D(c, k, padding, activation='relu')  (model_library.py:72,89,323...)."""
    w, b = wb
    kh = w.shape[0]
    pad = (kh // 2) if padding == "same" else 0
    y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1).to(x.dtype), b.to(x.dtype), padding=pad)
    return torch.relu(y).permute(0, 2, 3, 1).contiguous()


def maxpool2(x):
    """MaxPooling2D(pool_size=2, strides=2, padding='valid')  (model_library.py:74)."""
    return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()


def upsample_bilinear(x, s, legacy=False):
    """UpSampling2D(size=(s,s), interpolation='bilinear')  (model_library.py:92).

    legacy=True gives the TF1 (align_corners=False, no half-pixel) kernel for comparison.
    """
    xc = x.permute(0, 3, 1, 2)
    if not legacy:
        y = F.interpolate(xc, scale_factor=s, mode="bilinear", align_corners=False)
    else:
        n, c, h, w = xc.shape
        ys = torch.arange(h * s, dtype=x.dtype) / s
        xs = torch.arange(w * s, dtype=x.dtype) / s
        y0 = ys.floor().long().clamp(max=h - 1)
        y1 = (y0 + 1).clamp(max=h - 1)
        x0 = xs.floor().long().clamp(max=w - 1)
        x1 = (x0 + 1).clamp(max=w - 1)
        fy = (ys - y0).view(1, 1, -1, 1)
        fx = (xs - x0).view(1, 1, 1, -1)
        top = xc[:, :, y0][:, :, :, x0] * (1 - fx) + xc[:, :, y0][:, :, :, x1] * fx
        bot = xc[:, :, y1][:, :, :, x0] * (1 - fx) + xc[:, :, y1][:, :, :, x1] * fx
        y = top * (1 - fy) + bot * fy
    return y.permute(0, 2, 3, 1).contiguous()


def downblock(x, W, name, taps=None):
    """Downblock.call  (model_library.py:75-79): conv, conv -> (pre-pool, pooled)."""
    x = conv2d_relu(x, W[name + ".conv2d1"], "same")
    o1 = conv2d_relu(x, W[name + ".conv2d2"], "same")
    if taps is not None:
        taps[name + ".conv2d1"] = x
        taps[name + ".conv2d2"] = o1
    return o1, maxpool2(o1)


def upblock(x, skip, W, name, s=2, legacy_resize=False, taps=None):
    """Upblock.call  (model_library.py:93-101): upsample, concat [up, skip], 3x conv."""
    x = upsample_bilinear(x, s, legacy_resize)
    o1 = torch.cat([x, skip], dim=-1)
    o2 = conv2d_relu(o1, W[name + ".conv2d1"], "same")
    o3 = conv2d_relu(o2, W[name + ".conv2d2"], "same")
    out = conv2d_relu(o3, W[name + ".conv2d3"], "same")
    if taps is not None:
        taps[name + ".conv2d1"] = o2
        taps[name + ".conv2d2"] = o3
        taps[name + ".conv2d3"] = out
    return out


def poolskip(x, k):
    """Poolskip.call  (model_library.py:109-113): global average pool, tile k x k."""
    m = x.mean(dim=(1, 2), keepdim=True)
    return m.repeat(1, k, k, 1)


# ------------------------------------------------------------------ per-pixel filtering
def convolve(img_stack, filts, K):
    """Convolve.call / cus_convolve  (model_library.py:120-135, 136-152), literal."""
    N, H, W, T = img_stack.shape
    filts = filts.reshape(N, H, W, K * K * T)
    kpad = K // 2
    imgs = F.pad(img_stack, (0, 0, kpad, kpad, kpad, kpad))
    stack = []
    for i in range(K):
        for j in range(K):
            stack.append(imgs[:, i:i + H, j:j + W, :])
    stack = torch.stack(stack, dim=-2).reshape(N, H, W, K * K * T)
    return (stack * filts).sum(-1)


def convolve_perlayer(conv_stack, filts, K):
    """Convolve_perlayer.call  (model_library.py:160-168): per frame, scaled by T."""
    T = conv_stack.shape[-1]
    outs = []
    for i in range(T):
        one = convolve(conv_stack[..., i:i + 1], filts[..., i:i + 1], K) * T
        outs.append(one.unsqueeze(-1))
    return torch.cat(outs, dim=-1)


def kpn_apply_literal(burst, coef, bas):
    """model_library.py:439-451 exactly as written (tile / multiply / reduce_sum).

    burst [N,H,W,T], coef [N,H,W,B] (post-softmax), bas [N,K,K,T,B] -> [N,H,W,T+1].
    Memory is O(N*H*W*K*K*T*B): small inputs only.
    """
    N, H, W, T = burst.shape
    K = bas.shape[1]
    Coefficients = coef[:, :, :, None, None, None, :].repeat(1, 1, 1, K, K, T, 1)
    Basis = bas[:, None, None].repeat(1, H, W, 1, 1, 1, 1)
    filts = (Basis * Coefficients).sum(-1)                       # :444
    deblur = convolve(burst, filts, K).unsqueeze(-1)             # :447
    per = convolve_perlayer(burst, filts, K)                     # :449
    return torch.cat([deblur, per], dim=-1)                      # :451


def kpn_apply_algebraic(burst, coef, bas):
    """Same result, basis-first: out_t = T * sum_b coef_b * (burst_t (*) bas[:,:,t,b]).

    Used for inputs too large for the literal form; asserted equal in the tests.
    """
    N, H, W, T = burst.shape
    K, B = bas.shape[1], bas.shape[-1]
    x = burst.permute(0, 3, 1, 2).reshape(1, N * T, H, W)
    # weight [N*T*B, 1, K, K], grouped so that (n,t) only sees its own B kernels
    w = bas.permute(0, 3, 4, 1, 2).reshape(N * T * B, 1, K, K)
    g = F.conv2d(x, w, padding=K // 2, groups=N * T).reshape(N, T, B, H, W)
    per = T * (g * coef.permute(0, 3, 1, 2)[:, None]).sum(2)     # [N,T,H,W]
    per = per.permute(0, 2, 3, 1)
    deblur = per.mean(-1, keepdim=True)
    return torch.cat([deblur, per], dim=-1)


def basis_softmax(originbasis, K, T, B):
    """model_library.py:436-438: softmax over the K*K*T taps (axis=1) for each basis b."""
    N = originbasis.shape[0]
    s = torch.softmax(originbasis.reshape(N, K * K * T, B), dim=1)
    return s.reshape(N, K, K, T, B)


# ------------------------------------------------------------------------------ models
_ADD = {"singlestd": 1, "dualparams": 2, "empty": 0}


def simplemodel_forward(W, params, inputs, legacy_resize=False, literal_filter=None, taps=None):
    """Simplemodel.call  (model_library.py:372-452).

    Returns (output [N,H,W,T+1], Bas [N,K,K,T,B], originbasis [N,K,K,T*B]).
    If ``taps`` is a dict it is filled with intermediates (Coef, coef_logits, every conv output).
    """
    T, K, B = params["BURST_LENGTH"], params["Kernel_size"], params["Basis_num"]
    N, H, Wd, C = inputs.shape
    assert C == T + _ADD[params["layer_type"]]
    assert H % 8 == 0 and Wd % 8 == 0, "Simplemodel needs H, W multiples of 8 (reference crashes at :96)"
    burst = inputs[..., 0:T]                                           # :373
    x0 = conv2d_relu(inputs, W["layer0"], "same")                      # :376
    skip1, o1 = downblock(x0, W, "down1", taps)                              # :378
    skip2, o2 = downblock(o1, W, "down2", taps)                              # :379
    skip5, o5 = downblock(o2, W, "down5", taps)                              # :384
    o6 = conv2d_relu(o5, W["layer1_1"], "same")                        # :386
    up5 = upblock(o6, skip5, W, "Coef_up1", 2, legacy_resize, taps)    # :391
    up2 = upblock(up5, skip2, W, "Coef_up4", 2, legacy_resize, taps)   # :398
    up1 = upblock(up2, skip1, W, "Coef_up5", 2, legacy_resize, taps)   # :399
    o7 = conv2d_relu(up1, W["layer2_1"], "same")                       # :401
    logits = conv2d_relu(o7, W["coef"], "same")                        # :405 (ReLU before softmax)
    coef = torch.softmax(logits, dim=-1)                               # :406
    gavg = o6.mean(dim=1, keepdim=True).mean(dim=2, keepdim=True)      # :409-410
    ps5 = poolskip(skip5, 2)                                           # :411
    ub5 = upblock(gavg, ps5, W, "Basis_up1", 2, legacy_resize, taps)   # :414
    ps2 = poolskip(skip2, 16)                                          # :421
    ub2 = upblock(ub5, ps2, W, "Basis_up4", 8, legacy_resize, taps)    # :422 (s=8 at :362)
    o8 = conv2d_relu(ub2, W["layer3_1"], "valid")                      # :424 (2x2 valid -> 15x15)
    originbasis = conv2d_relu(o8, W["layer3_3"], "same")               # :428
    assert originbasis.shape[1] == K and originbasis.shape[2] == K, "Kernel_size must be 15"
    bas = basis_softmax(originbasis, K, T, B)                          # :436-438
    if literal_filter is None:
        literal_filter = N * H * Wd * K * K * T * B <= 40_000_000
    fn = kpn_apply_literal if literal_filter else kpn_apply_algebraic
    output = fn(burst, coef, bas)                                      # :439-451
    if taps is not None:
        taps.update({"layer0": x0, "layer1_1": o6, "layer2_1": o7, "coef_logits": logits,
                     "Coef": coef, "layer3_1": o8})
    return output, bas, originbasis


def basis_kpn_forward(W, params, inputs, legacy_resize=False, literal_filter=None, taps=None):
    """Basis_kpn.call  (model_library.py:231-295).  Needs H, W multiples of 32.

    Returns (output [N,H,W,T+1], Bas [N,K,K,T,B]).
    """
    T, K, B = params["BURST_LENGTH"], params["Kernel_size"], params["Basis_num"]
    N, H, Wd, C = inputs.shape
    assert C == T + _ADD[params["layer_type"]]
    assert H % 32 == 0 and Wd % 32 == 0
    burst = inputs[..., 0:T]                                           # :234
    x0 = conv2d_relu(inputs, W["layer0"], "same")                      # :235
    skip1, o1 = downblock(x0, W, "down1", taps)                              # :237
    skip2, o2 = downblock(o1, W, "down2", taps)
    skip3, o3 = downblock(o2, W, "down3", taps)
    skip4, o4 = downblock(o3, W, "down4", taps)
    skip5, o5 = downblock(o4, W, "down5", taps)                              # :241
    o6_1 = conv2d_relu(o5, W["layer1_1"], "same")                      # :243
    o6 = conv2d_relu(o6_1, W["layer1_2"], "same")                      # :244
    up5 = upblock(o6, skip5, W, "Coef_up1", 2, legacy_resize, taps)    # :246
    up4 = upblock(up5, skip4, W, "Coef_up2", 2, legacy_resize, taps)
    up3 = upblock(up4, skip3, W, "Coef_up3", 2, legacy_resize, taps)
    up2 = upblock(up3, skip2, W, "Coef_up4", 2, legacy_resize, taps)
    up1 = upblock(up2, skip1, W, "Coef_up5", 2, legacy_resize, taps)   # :250
    o7_1 = conv2d_relu(up1, W["layer2_1"], "same")                     # :252
    o7 = conv2d_relu(o7_1, W["layer2_2"], "same")                      # :253
    logits = conv2d_relu(o7, W["coef"], "same")                        # :254
    coef = torch.softmax(logits, dim=-1)                               # :255
    gavg = o6.mean(dim=1, keepdim=True).mean(dim=2, keepdim=True)      # :258-259
    ub5 = upblock(gavg, poolskip(skip5, 2), W, "Basis_up1", 2, legacy_resize, taps)   # :260-263
    ub4 = upblock(ub5, poolskip(skip4, 4), W, "Basis_up2", 2, legacy_resize, taps)    # :264-265
    ub3 = upblock(ub4, poolskip(skip3, 8), W, "Basis_up3", 2, legacy_resize, taps)    # :266-267
    ub2 = upblock(ub3, poolskip(skip2, 16), W, "Basis_up4", 2, legacy_resize, taps)   # :268-269
    o8_1 = conv2d_relu(ub2, W["layer3_1"], "valid")                    # :271
    o8_2 = conv2d_relu(o8_1, W["layer3_2"], "same")                    # :272
    originbasis = conv2d_relu(o8_2, W["layer3_3"], "same")             # :273
    bas = basis_softmax(originbasis, K, T, B)                          # :279-281
    if literal_filter is None:
        literal_filter = N * H * Wd * K * K * T * B <= 40_000_000
    fn = kpn_apply_literal if literal_filter else kpn_apply_algebraic
    output = fn(burst, coef, bas)                                      # :282-294
    if taps is not None:
        taps.update({"layer0": x0, "layer1_1": o6_1, "layer1_2": o6, "layer2_1": o7_1, "layer2_2": o7,
                     "coef_logits": logits, "Coef": coef, "originbasis": originbasis,
                     "layer3_1": o8_1, "layer3_2": o8_2})
    return output, bas
