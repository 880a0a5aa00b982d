"""Forward-pass engine for the reference's two networks on one B200.

Table-driven restatement of the call graphs of ``Simplemodel.call``
(/root/reference/model_library.py:372-452) and ``Basis_kpn.call`` (:231-295) as a
sequence of C-ABI kernel launches on the current CUDA stream:

    im2col pack -> [tcgen05 conv]* with max-pool / bilinear-upsample / channel-mean glue
    -> coef conv with fused softmax -> basis branch (tiny rasters) -> basis softmax
    -> fused per-pixel filter (kpn_apply).

Concatenations are channel slices of shared rasters: the skip half is written by the
encoder conv's epilogue, the up-sampled half by the upsample kernel.  No activation is
copied, nothing is synchronised, and activation buffers are cached per input shape, so a
forward is capturable in a CUDA graph.
"""
from __future__ import annotations

import os

import torch

from . import _lib, ops
from ._lib import IE_EPI_BF16_RASTER, IE_EPI_F32_NHWC, IE_EPI_F32_SOFTMAX, ImgEnhError
from .weights import ADD_LENGTHS, basis_kpn_layers, simplemodel_layers

# (down blocks, bottleneck convs, coef up blocks (name, cout, skip), head convs,
#  basis up blocks (name, cout, skip, k, scale), tail convs) - model_library.py:323-368 / 196-227
ARCH = {
    "simple": dict(
        downs=[("down1", 64), ("down2", 128), ("down5", 1024)],
        bottleneck=["layer1_1"],
        coef_ups=[("Coef_up1", 512, "down5"), ("Coef_up4", 64, "down2"), ("Coef_up5", 64, "down1")],
        head=["layer2_1"],
        basis_ups=[("Basis_up1", 512, "down5", 2, 2), ("Basis_up4", 128, "down2", 16, 8)],
        tail=["layer3_1", "layer3_3"],
        layers=simplemodel_layers,
    ),
    "basis_kpn": dict(
        downs=[("down1", 64), ("down2", 128), ("down3", 256), ("down4", 512), ("down5", 1024)],
        bottleneck=["layer1_1", "layer1_2"],
        coef_ups=[("Coef_up1", 512, "down5"), ("Coef_up2", 256, "down4"), ("Coef_up3", 128, "down3"),
                  ("Coef_up4", 64, "down2"), ("Coef_up5", 64, "down1")],
        head=["layer2_1", "layer2_2"],
        basis_ups=[("Basis_up1", 512, "down5", 2, 2), ("Basis_up2", 256, "down4", 4, 2),
                   ("Basis_up3", 256, "down3", 8, 2), ("Basis_up4", 128, "down2", 16, 2)],
        tail=["layer3_1", "layer3_2", "layer3_3"],
        layers=basis_kpn_layers,
    ),
}


class Engine:
    def __init__(self, arch, params, weights, device="cuda"):
        if not torch.cuda.is_available():
            raise ImgEnhError("a CUDA device is required (the hot path has no CPU fallback)")
        self.arch = ARCH[arch]
        self.params = params
        self.T = params["BURST_LENGTH"]
        self.K = params["Kernel_size"]
        self.B = params["Basis_num"]
        self.cin = self.T + ADD_LENGTHS[params["layer_type"]]
        self.k0 = ops.im2col_width(self.cin)                 # K of the first layer as a 1x1 GEMM over im2col rows
        if self.K != 15:
            raise ImgEnhError("Kernel_size must be 15: the basis branch emits 15x15 kernels (model_library.py:364)")
        self.device = torch.device(device)
        # per-pixel filter (model_library.py:439-451).  "auto" (default): the tcgen05 filter-synthesis kernel where it
        # applies (K = 15, T % 4 == 0), else the mma.sync TF32 kernel (K = 15, B <= 128, T <= 8), else fp32.  Both
        # tensor-core kernels round their MMA operands to a 10-bit mantissa (tcgen05: coefficients and basis as fp16,
        # the burst stays fp32; mma.sync: burst and basis as TF32) and accumulate in fp32: |err| <= 2^-10 of the pixel
        # range.  "fp32" = CUDA-core kernel (1e-5 of the fp64 oracle); "tf32" / "tcgen05"
        # pin one tensor-core kernel (falling back down the same chain where it does not apply).
        self.filter_precision = params.get("filter_precision", "auto")
        if self.filter_precision not in ("auto", "tf32", "fp32", "tcgen05"):
            raise ImgEnhError("filter_precision must be 'auto', 'tcgen05', 'tf32' or 'fp32'")
        self.stride = 2 ** len(self.arch["downs"])          # 8 for Simplemodel, 32 for Basis_kpn
        # tensors at resolution 1/2^level and below are dense NHWC (TMA im2col); 99 = shared-border rasters everywhere
        self.dense_from_level = int(params.get("dense_from_level", 2))
        if self.device.type != "cuda":
            raise ImgEnhError("device must be a CUDA device (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.wp, self.bias = {}, {}
        self._plans = {}
        self._graphs = {}
        self._ws = None
        self._side = None
        # run the per-image basis branch on a second stream beside the coefficient decoder (see _forward)
        self.overlap_branches = bool(params.get("overlap_branches", os.environ.get("IE_OVERLAP", "1") != "0"))
        with torch.cuda.device(self.device):
            self.load_weights(weights)
        # small batches are launch-bound (44 kernels through ctypes, ~1.0 ms of host time per forward against ~0.25 ms
        # of GPU time for eval.py's default batch of one 32x32 patch, and still against 1.5 ms at 32 patches of 100x100
        # once a loop around the forward adds its own host work): replay a captured CUDA graph below this many pixels.
        # On the GPU eager and replay cost the same from 16 x 100 x 100 up (profiles/r02_latency_small_final.txt); the
        # threshold is where the host stops mattering (64 x 100 x 100: 2.8 ms of GPU time per forward).
        self.graph_max_pixels = int(params.get("graph_max_pixels", 1 << 19))

    # ------------------------------------------------------------------ weights
    def load_weights(self, weights):
        """weights: {name: (kernel HWIO, bias)} (CPU or CUDA).  Packs to bf16 [cout_pad, K] on the device."""
        f32_heads = {"coef": IE_EPI_F32_SOFTMAX, "layer3_3": IE_EPI_F32_NHWC}
        # captured CUDA graphs bake in the packed-weight / bias pointers (and tensor maps built from them): a graph
        # captured before this call would replay against freed memory
        self._graphs = {}
        for name, k, cin, cout, _ in self.arch["layers"](self.params):
            w, b = weights[name]
            assert tuple(w.shape) == (k, k, cin, cout), (name, tuple(w.shape), (k, k, cin, cout))
            w = w.to(self.device, torch.float32)
            epi = f32_heads.get(name, IE_EPI_BF16_RASTER)
            ktot_pad = self.k0 if name == "layer0" else None
            self.wp[name] = ops.pack_conv_weights(w, epi, ktot_pad)
            self.bias[name] = b.to(self.device, torch.float32).contiguous()

    # ------------------------------------------------------------------ buffers
    def _workspace(self, side=False):
        """Split-K scratch of this engine (its kernels are stream-ordered; a captured graph keeps using the buffer it was
        captured with).  ``side``: the second buffer, for the basis branch when it runs on its own stream."""
        if self._ws is None:
            self._ws = [torch.empty(ops.SPLITK_WORKSPACE_BYTES, dtype=torch.uint8, device=self.device) for _ in range(2)]
        return self._ws[1 if side else 0]

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        return self._side

    def _plan(self, n, h, w):
        key = (n, h, w)
        if key in self._plans:
            return self._plans[key]
        A = self.arch
        dev = self.device
        # Layout per tensor: full and 1/2 resolution (the cout <= 64 wide-N kernels live there, and a border costs
        # 2 - 4 % of the rows) are shared-border rasters; from 1/4 resolution down, and the small per-image tensors of
        # the basis branch, are dense NHWC read through TMA im2col maps (a border would be 8 % of the GEMM rows at
        # 26 x 26, 16 % at 13 x 13, 125 % at 2 x 2).  16 x 16 basis tensors stay rasters: the 2 x 2 'valid' conv and
        # the fp32 head that follow them are raster kernels.
        dense_levels = self.dense_from_level
        lvl_dense = lambda l: l >= dense_levels
        R = lambda hh, ww, c, dense=False: ops.new_raster(n, hh, ww, c, dev, dense=dense)
        p = {}
        chans = dict(A["downs"])
        # encoder rasters per level l (resolution h >> l)
        p["x0"] = R(h, w, 64)
        up_in = {}                                   # channels entering each coef up block
        prev = 1024
        for name, cout, skip in A["coef_ups"]:
            up_in[skip] = prev
            prev = cout
        for l, (dname, c) in enumerate(A["downs"]):
            hh, ww = h >> l, w >> l
            p[dname + ".c1"] = R(hh, ww, c, lvl_dense(l))
            p["cat." + dname] = R(hh, ww, up_in[dname] + c, lvl_dense(l))   # [0:up_in]=upsampled, [up_in:]=skip
            p[dname + ".pool"] = R(hh >> 1, ww >> 1, c, lvl_dense(l + 1))
        L = len(A["downs"])
        for i, bname in enumerate(A["bottleneck"]):
            p[bname] = R(h >> L, w >> L, 1024, lvl_dense(L))
        for name, cout, skip in A["coef_ups"]:
            l = [d for d, _ in A["downs"]].index(skip)
            for j in (1, 2, 3):
                p[f"{name}.c{j}"] = R(h >> l, w >> l, cout, lvl_dense(l))
        for hname in A["head"]:
            p[hname] = R(h, w, 64)
        # basis branch
        small = lambda k: dense_levels < 99 and k <= 8
        prev = 1024
        for name, cout, skip, k, s in A["basis_ups"]:
            p["bcat." + name] = R(k, k, prev + chans[skip], small(k))
            for j in (1, 2, 3):
                p[f"{name}.c{j}"] = R(k, k, cout, small(k))
            prev = cout
        p["seed"] = R(1, 1, 1024, small(1))
        for tname in A["tail"][:-1]:
            p[tname] = R(16, 16, 128)
        self._plans[key] = p
        return p

    # ------------------------------------------------------------------ CUDA-graph replay for small inputs
    def forward_auto(self, x):
        """forward(), through a captured CUDA graph when the input is small enough to be launch-bound."""
        n, hs, ws, _ = x.shape
        if x.is_cuda and x.device != self.device:
            raise ImgEnhError(f"input is on {x.device}, the model on {self.device}")
        if n * hs * ws > self.graph_max_pixels or torch.cuda.is_current_stream_capturing():
            return self.forward(x)
        with torch.cuda.device(self.device):
            return self._forward_graph(x)

    def _forward_graph(self, x):
        key = (tuple(x.shape), x.device.index)
        ent = self._graphs.get(key)
        if ent is None:
            static_x = torch.empty_like(x, dtype=torch.float32).contiguous()
            static_x.copy_(x)
            side = torch.cuda.Stream(x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):                   # warm-up outside the capture: buffers, function attributes
                self.forward(static_x)
            torch.cuda.current_stream(x.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            before = sum(_lib.LAUNCHES.values())
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                outs = self.forward(static_x)
            calls = sum(_lib.LAUNCHES.values()) - before          # C-ABI calls recorded into the graph
            ent = self._graphs[key] = (g, static_x, outs, calls)
        g, static_x, outs, calls = ent
        static_x.copy_(x)
        g.replay()
        _lib.LAUNCHES["(graph replay)"] += calls                 # the launch accounting (bench.py gpu_launches) stays honest
        return tuple(o.clone() for o in outs)               # the graph's output buffers are overwritten by the next replay

    # ------------------------------------------------------------------ forward
    def forward(self, x, taps=None, conv_fn="ie_conv2d_nhwc_bf16", shard=None):
        """x: fp32 NHWC [n,hs,ws,T+add] on CUDA.  The network runs at (h, w) = (hs, ws) rounded up to the
        network stride; the zero padding at the bottom/right (what the boundary does for the 100x100 patches
        of BASELINE.json) is implicit in the first and last kernels - no padded copy is made.

        Returns (output [n,hs,ws,T+1], Bas [n,K,K,T,B], originbasis [n,K,K,T*B]), all fp32.
        ``taps``: optional dict filled with fp32 copies of intermediates (parity tests).
        ``conv_fn``: tests may route every convolution through the naive validation kernel.
        ``shard``: spatial sharding of ONE large image over several ranks (dist.spatial_shards).  ``x`` is then this
        rank's slab (its rows plus a halo): dict(own=(r0, r1) rows of the slab this rank owns, total_rows=H of the
        whole image, reduce=callable).  The per-image pooled statistics that feed the basis branch
        (model_library.py:409-411, 421) are taken over the owned rows only, scaled by the whole image's pixel count and
        handed to ``reduce`` (an all-reduce SUM over the ranks) before the basis branch runs - the one exchange step
        spatial sharding needs.  The output covers the whole slab; rows outside ``own`` are halo (discard them).
        """
        if not x.is_cuda:
            raise ImgEnhError("input must be a CUDA tensor (no CPU fallback)")
        if x.device != self.device:
            raise ImgEnhError(f"input is on {x.device}, the model on {self.device}")
        with torch.cuda.device(self.device):      # kernels go to the current stream of the MODEL's device
            return self._forward(x, taps, conv_fn, shard)

    def _forward(self, x, taps, conv_fn, shard):
        n, hs, ws, c = x.shape
        if c != self.cin:
            raise ImgEnhError(f"expected {self.cin} input channels, got {c}")
        st = self.stride
        h, w = -(-hs // st) * st, -(-ws // st) * st
        x = x.contiguous().float()
        A, W, Bv = self.arch, self.wp, self.bias
        p = self._plan(n, h, w)
        if shard is not None:
            r0, r1 = shard["own"]
            if r0 % st or r1 % st or not (0 <= r0 < r1 <= h) or shard["total_rows"] % st:
                raise ImgEnhError(f"spatial shard rows {r0}:{r1} / {shard['total_rows']} must be multiples of {st}")

        def stat_rows(level):
            """(rows, count) of the pooled statistics of a tensor at resolution 1/2^level."""
            if shard is None:
                return None, 0
            r0, r1 = shard["own"]
            return (r0 >> level, r1 >> level), (shard["total_rows"] >> level) * (w >> level)
        chans = dict(A["downs"])
        dnames = [d for d, _ in A["downs"]]

        # split-K scratch: the library only uses it for layers with too few output tiles to fill the SMs (small inputs,
        # and the per-image basis branch of small batches)
        wksp = [self._workspace()]

        def conv(name, src, dst, k=3, valid=None):
            ops.conv2d(src, W[name], Bv[name], dst, k=k, relu=True, valid=valid, fn=conv_fn, workspace=wksp[0])
            if taps is not None:
                hv, wv = (dst.r.h, dst.r.w) if valid is None else valid
                taps[name] = ops.raster_to_nhwc(dst)[:, :hv, :wv]

        # ---- encoder (model_library.py:376-386 / 235-244)
        if c in ops.FIRST_LAYER_FUSED_CHANNELS and conv_fn == "ie_conv2d_nhwc_bf16":
            ops.conv_first_layer(x, W["layer0"], Bv["layer0"], p["x0"].slice())        # im2col fused in-kernel
            if taps is not None:
                taps["layer0"] = ops.raster_to_nhwc(p["x0"].slice())
        else:
            if "in0" not in p:
                p["in0"] = ops.new_raster(n, h, w, self.k0, self.device)
            ops.pack_input_im2col3x3(x, p["in0"])
            conv("layer0", p["in0"].slice(), p["x0"].slice(), k=1)
        cur = p["x0"].slice()
        up_in = {}
        prev = 1024
        for name, cout, skip in A["coef_ups"]:
            up_in[skip] = prev
            prev = cout
        basis_skips = {skip for _, _, skip, _, _ in A["basis_ups"]}
        skip_means = {}
        for level, (dname, cch) in enumerate(A["downs"]):
            conv(dname + ".conv2d1", cur, p[dname + ".c1"].slice())
            skip = p["cat." + dname].slice(up_in[dname], cch)
            conv(dname + ".conv2d2", p[dname + ".c1"].slice(), skip)
            # the pool also yields the channel means of the skip it reads, if the basis branch wants them (Poolskip)
            rows, count = stat_rows(level)
            skip_means[dname] = ops.maxpool2(skip, p[dname + ".pool"].slice(), want_mean=dname in basis_skips,
                                             rows=rows, count=count)
            cur = p[dname + ".pool"].slice()
        for bname in A["bottleneck"]:
            conv(bname, cur, p[bname].slice())
            cur = p[bname].slice()
        bott = cur
        # ---- coefficient decoder (model_library.py:391-406 / 246-255)
        def coef_decoder():
            cur = bott
            for name, cout, skip in A["coef_ups"]:
                cat = p["cat." + skip]
                ops.upsample_bilinear(cur, cat.slice(0, up_in[skip]), 2)
                conv(name + ".conv2d1", cat.slice(), p[name + ".c1"].slice())
                conv(name + ".conv2d2", p[name + ".c1"].slice(), p[name + ".c2"].slice())
                conv(name + ".conv2d3", p[name + ".c2"].slice(), p[name + ".c3"].slice())
                cur = p[name + ".c3"].slice()
            for hname in A["head"]:
                conv(hname, cur, p[hname].slice())
                cur = p[hname].slice()
            return ops.conv2d_f32(cur, W["coef"], Bv["coef"], self.B, softmax=True, want_logits=taps is not None, fn=conv_fn)

        # ---- basis branch (model_library.py:409-438 / 258-281)
        def basis_branch():
            rows, count = stat_rows(len(A["downs"]))
            gavg = ops.channel_mean(bott, rows=rows, count=count)            # :409-410
            if shard is not None:
                # the exchange step of spatial sharding: partial means (already divided by the global pixel count) of
                # the bottleneck and of every pooled skip, summed over the ranks
                names = [d for d in dnames if skip_means.get(d) is not None]
                packed = torch.cat([gavg] + [skip_means[d] for d in names], dim=1)
                packed = shard["reduce"](packed)
                gavg = packed[:, :gavg.shape[1]].contiguous()
                off = gavg.shape[1]
                for d in names:
                    cd = skip_means[d].shape[1]
                    skip_means[d] = packed[:, off:off + cd].contiguous()
                    off += cd
            ops.broadcast_hw(gavg, p["seed"].slice())
            cur = p["seed"].slice()
            for name, cout, skip, k, s in A["basis_ups"]:
                cat = p["bcat." + name]
                cin_up = cur.c
                ops.upsample_bilinear(cur, cat.slice(0, cin_up), s)          # Upblock.upsampling :94
                skip_mean = skip_means[skip]                                  # Poolskip :110, from the max-pool kernel
                ops.broadcast_hw(skip_mean, cat.slice(cin_up, chans[skip]))  # tile :112, concat :96
                conv(name + ".conv2d1", cat.slice(), p[name + ".c1"].slice())
                conv(name + ".conv2d2", p[name + ".c1"].slice(), p[name + ".c2"].slice())
                conv(name + ".conv2d3", p[name + ".c2"].slice(), p[name + ".c3"].slice())
                cur = p[name + ".c3"].slice()
            tail = A["tail"]
            conv(tail[0], cur, p[tail[0]].slice(), k=2, valid=(15, 15))      # 2x2 'valid' :364/424
            cur = p[tail[0]].slice()
            for tname in tail[1:-1]:
                conv(tname, cur, p[tname].slice(), valid=(15, 15))
                cur = p[tname].slice()
            originbasis = ops.conv2d_f32(cur, W[tail[-1]], Bv[tail[-1]], self.T * self.B, valid=(15, 15), fn=conv_fn)
            return ops.softmax_taps(originbasis, self.T, self.B), originbasis              # :436-438

        # The basis branch reads the bottleneck and the pooled skip means only; the coefficient decoder never reads
        # anything the branch writes.  Its ~15 launches work on 2x2 ... 16x16 pixels per image (tens of CTAs each,
        # latency-bound), so it runs on a second stream beside the decoder and joins before the filter: at small
        # batches the two really overlap (32 patches: 1.55 -> 1.4 ms per step; one 32x32 patch: 0.30 -> 0.2x ms),
        # at 256 patches the small kernels fill the gaps between the decoder's waves.  Fork / join are events, so the
        # same code is captured into the CUDA graph of the small-batch path.  Tensors the branch allocates come from
        # the side stream's pool; they are only re-used by a later forward's branch, which starts behind that forward's
        # fork event and therefore behind every main-stream reader enqueued before it.
        if self.overlap_branches and shard is None and taps is None:
            main, side = torch.cuda.current_stream(self.device), self._side_stream()
            fork, join = torch.cuda.Event(), torch.cuda.Event()
            fork.record(main)
            with torch.cuda.stream(side):
                side.wait_event(fork)
                wksp[0] = self._workspace(side=True)
                bas, originbasis = basis_branch()
                join.record(side)
            wksp[0] = self._workspace()
            coef, logits = coef_decoder()
            main.wait_event(join)
        else:
            coef, logits = coef_decoder()
            bas, originbasis = basis_branch()
        # ---- per-pixel filter (model_library.py:439-451)
        want = self.filter_precision
        if want in ("auto", "tcgen05") and ops.kpn_tcgen05_supported(self.T, self.K, self.B):
            prec = "tcgen05"
        elif want != "fp32" and ops.kpn_tf32_supported(self.T, self.K, self.B):
            prec = "tf32"
        else:
            prec = "fp32"
        out = ops.kpn_apply(x, self.T, coef, bas, precision=prec)
        if taps is not None:
            taps["Coef"], taps["coef_logits"] = coef, logits
        return out, bas, originbasis
