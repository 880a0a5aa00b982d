/* imgenh_b200 - C ABI of the B200-native hot path of hanxuel/ImageEnhancement_MP.
 *
 * The reference (/root/reference) is pure Python on TensorFlow/Keras and has no FFI layer of
 * its own; its hot path is `Simplemodel.call` / `Basis_kpn.call` (model_library.py:372-452,
 * 231-295) plus the metric functions of data_utils.py:24-164 that eval.py:139-195 drives.
 * Each entry point below names the reference op(s) it replaces.  The Python shim
 * (imageenhancement_mp_b200/) binds these with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless it says "host".
 *   - the caller owns and allocates every buffer; the library keeps no state between calls
 *     (besides a per-thread error string and the cached SM count).
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises.
 *   - return 0 on success, <0 on error; ie_last_error() gives the message for this thread.
 *
 * Activation layout ("raster"): bf16, NHWC, with a SHARED one-pixel zero border: every image is preceded by
 *   one zero row and every image row is followed by one zero pixel,
 *   rows = n_img * (h+1) * (w+1), pixel (n, y, x) at row r = (n*(h+1) + y + 1)*(w+1) + x, `pitch` channels per row.
 *   The zero pixel ending row y is also the left neighbour of row y+1, the zero row of image n+1 is also the
 *   bottom border of image n, and rows before the first / after the last image do not exist (TMA fills
 *   out-of-range rows with zeros).  A k x k 'same' convolution tap is then a constant row shift of this 2-D
 *   [rows][pitch] matrix, so TMA fetches every tap with a plain 2-D box and the border supplies the padding -
 *   at (h+1)(w+1)/(hw) - 1 extra rows (8 % at 26x26) instead of 16 % for a full border.
 *   Channel slices (coff, c) of a wider raster are how concatenation is expressed.
 */
#ifndef IMGENH_B200_H
#define IMGENH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IE_VERSION 100

#define IE_OK 0
#define IE_ERR_INVALID (-1)
#define IE_ERR_CUDA (-2)

/* epilogue kinds for ie_conv2d_nhwc_bf16 */
#define IE_EPI_BF16_RASTER 0 /* bias(+ReLU) -> bf16 raster slice (border rows zeroed)                    */
#define IE_EPI_F32_NHWC 1    /* bias(+ReLU) -> fp32 [n][hv][wv][cout] (interior only), cout <= 256         */
#define IE_EPI_F32_SOFTMAX 2 /* as 1, then softmax over cout; y_aux (nullable) receives the logits        */

/* Activation layouts.  RASTER (default): bf16 NHWC with a shared one-pixel zero border - rows = n*(h+1)*(w+1),
 * pixel (n,y,x) at row (n*(h+1)+y+1)*(w+1)+x; a k x k tap is a constant row shift, fetched with plain 2-D TMA boxes.
 * DENSE: plain bf16 NHWC, rows = n*h*w, pixel (n,y,x) at row (n*h+y)*w+x; the convolution fetches its taps with TMA
 * im2col-mode tensor maps (the zero padding comes from the tensor map's pixel box), so no border row is ever
 * multiplied - the layout of the 1/4-resolution and smaller tensors (a border costs 7.8 % of the rows at 26 x 26,
 * 16 % at 13 x 13, 125 % at 2 x 2).  `layout` arguments of the glue kernels are a bit set of these flags.          */
#define IE_LAYOUT_X_DENSE 1 /* the input tensor is dense NHWC  */
#define IE_LAYOUT_Y_DENSE 2 /* the output tensor is dense NHWC */

typedef struct ie_conv_desc {
  int32_t n_img, h, w; /* interior size of the INPUT raster (rows = n_img*(h+1)*(w+1); dense: n_img*h*w)  */
  int32_t hv, wv;      /* valid OUTPUT extent inside the same raster geometry: (h,w) for 'same',
                          (h-1,w-1) for the 2x2 'valid' conv; everything outside is written as zero      */
  int32_t kh, kw;      /* 3x3 ('same', centred), 2x2 ('valid', taps at +0/+1) or 1x1                     */
  int32_t cin;         /* input channels consumed, multiple of 64                                        */
  int32_t x_pitch;     /* channels per row of x                                                          */
  int32_t x_coff;      /* first channel of x to read, multiple of 64                                     */
  int32_t cout;        /* output channels                                                                */
  int32_t y_pitch;     /* channels per row of y (IE_EPI_BF16_RASTER); IE_EPI_F32_NHWC: 0 = dense [..][cout], else the
                          floats per pixel of a wider NHWC tensor of which this launch writes a channel slice    */
  int32_t y_coff;      /* first channel of y to write (bf16 raster: multiple of 64)                      */
  int32_t relu;        /* 1: max(0, .) after the bias                                                    */
  int32_t epilogue;    /* IE_EPI_*                                                                       */
  int32_t dense;       /* 0: x and y are shared-border rasters; 1: both are dense NHWC (3x3 'same' / 1x1,
                          IE_EPI_BF16_RASTER only; taps fetched by TMA im2col)                            */
} ie_conv_desc;

int ie_version(void);
const char* ie_last_error(void);
int ie_sm_count(void);

/* ---- convolutions: layers.Conv2D(c, 3, 'same', relu) / Conv2D(128, 2, 'valid', relu)
 *      (model_library.py:72-73, 89-91, 323-368) as tcgen05 implicit GEMM ----------------------------- */

/* HWIO fp32 [kh][kw][cin][cout] -> bf16 [cout][ktot_pad], k = (i*kw+j)*cin + c, zero padded.       */
int ie_pack_conv_weights(const float* hwio, int kh, int kw, int cin, int cout, int ktot_pad, void* packed_bf16,
                         void* stream);

/* fp32 NHWC [n][hs][ws][c] -> bf16 raster [n*(h+1)*(w+1)][kpad], kpad = 9*c rounded up to a multiple of 64,
 * holding each pixel's zero-padded 3x3xc neighbourhood, k = (i*3+j)*c + ch (rest zero): turns the first
 * conv (model_library.py:323/376, 196/235) into a 1x1 GEMM with K = kpad.  hs <= h, ws <= w: the source
 * is implicitly zero-padded at the bottom/right to the raster size (the network-stride padding).       */
int ie_pack_input_im2col3x3(const float* x, int n, int hs, int ws, int c, int h, int w, void* raster_bf16,
                            void* stream);

/* The first layer with the im2col fused into the kernel (c = 3, 5 or 10, cout = 64): x fp32 NHWC [n][hs][ws][c]
 * (implicitly zero-padded to h x w) -> bias(+ReLU) -> bf16 raster slice.  w_packed = ie_pack_conv_weights of the
 * 3x3xcx64 kernel with ktot_pad = 9*c rounded up to a multiple of 64 (64; 128 for c = 10: Basis_kpn with T = 8 +
 * dualparams).  Same result as ie_pack_input_im2col3x3 + a 1x1 ie_conv2d_nhwc_bf16.                             */
int ie_conv_first_layer_f32(const float* x, int n, int hs, int ws, int c, int h, int w, const void* w_packed,
                            const float* bias, int cout, int relu, void* y_bf16, int y_pitch, int y_coff, void* stream);

/* y = epilogue(conv(x, w) + bias).  w_packed from ie_pack_conv_weights with ktot_pad = kh*kw*cin.
 * y_bf16 is used by IE_EPI_BF16_RASTER; y_f32 (and optional y_aux) by the fp32 epilogues.
 * workspace (nullable, 16-byte aligned device memory of workspace_bytes, caller-owned like every other buffer): scratch
 * for split-K.  When a layer has so few output tiles that most SMs would idle (eval.py's default call is ONE 32 x 32
 * patch: 1024-channel layers with K up to 18432 on a single M tile), or when the last wave of its persistent grid is
 * less than half full, the K loop of those tiles is cut into parts computed by different CTAs (128 * n_tile * 4 bytes
 * of workspace per part); the fp32 partial sums go through the workspace and a second kernel reduces them in a fixed
 * order (deterministic).  NULL: never split.
 * The workspace is only used between this call's two kernels; calls on the same stream may share it.            */
int ie_conv2d_nhwc_bf16(const ie_conv_desc* d, const void* x, const void* w_packed, const float* bias, void* y_bf16,
                        float* y_f32, float* y_aux, void* workspace, long long workspace_bytes, void* stream);

/* Tuning / test hook: force the main-loop flavour of ie_conv2d_nhwc_bf16 (-1 auto, 0 streaming, 1 resident
 * weights, 2 wide-N).  flags: bit 1 one filter row per stage in the resident kernel, bit 2 wide-N streams its
 * weights, bit 3 flip the number of epilogue warp sets, bit 4 one filter row per stage in wide-N, bit 9 never split
 * the K loop (tiny-M layers and the last partial wave of larger grids otherwise run split-K), bit 10 plain
 * stream-ordered launches instead of programmatic dependent launch, bit 11 exchange epilogue for the one-block wide-N
 * layers, bit 12 no L2 prefetch in wide-N, bit 13 split-K only for tiny M, bit 14 narrow N tiles
 * instead of split-K on small grids, bit 15 first layer with per-thread gathers instead of staged source rows,
 * bit 16 N = 128 streaming layers without M-tile pairs, bit 17 cout = 128 layers through M-tile pairs instead of the
 * transposed kernel.  Process-wide.                        */
int ie_conv_set_mode(int mode, int flags);

/* Slow CUDA-core convolution with the same contract; TESTS ONLY (cross-checks the tcgen05 kernel at
 * sizes the CPU oracle cannot reach).  Never called by the product path.                             */
int ie_debug_conv2d_naive(const ie_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                          void* y_bf16, float* y_f32, float* y_aux, void* workspace, long long workspace_bytes,
                          void* stream);

/* ---- layout glue of the U-Net (bandwidth kernels) ------------------------------------------------- */

/* MaxPooling2D(2,2) (model_library.py:74): raster (h,w) slice -> raster (h/2,w/2) slice, border zeroed.
 * chan_mean (nullable, [n][c] fp32): also receives the per-image channel means of the INPUT slice - the
 * GlobalAveragePooling2D of Poolskip (model_library.py:110) on the tensor the pool reads anyway.  For a spatial
 * shard of a larger image the statistics cover input rows [stat_y0, stat_y1) only (even numbers; stat_y1 <= 0: all
 * rows) and are divided by stat_count pixels (<= 0: h*w), so that a SUM over the shards is the global mean.
 * The statistics are bit-reproducible: every block stores its partial sums in `stat_scratch` and the last block of an
 * image to finish adds them in a fixed order (no floating-point atomics).  stat_scratch (device memory, only read and
 * written by this call; required when chan_mean is given) must hold ie_maxpool2_stat_scratch_bytes(...) bytes.  */
long long ie_maxpool2_stat_scratch_bytes(int n, int h, int w, int c, int layout);
int ie_maxpool2_nhwc_bf16(const void* x, int n, int h, int w, int c, int x_pitch, int x_coff, void* y, int y_pitch,
                          int y_coff, float* chan_mean, void* stat_scratch, long long stat_scratch_bytes, int stat_y0,
                          int stat_y1, long long stat_count, int layout, void* stream);

/* UpSampling2D(scale, 'bilinear') half-pixel centres (model_library.py:92) written straight into a
 * channel slice of the consumer's concat raster (model_library.py:96); border zeroed.               */
int ie_upsample_bilinear_nhwc_bf16(const void* x, int n, int h, int w, int c, int x_pitch, int x_coff, int scale,
                                   void* y, int y_pitch, int y_coff, int layout, void* stream);

/* Per-image channel means over the interior (reduce_mean x2 :409-410, GlobalAveragePooling2D :110):
 * mean[n][c] fp32.  Rows [y0, y1) only when y1 > 0, divided by `count` pixels when count > 0 (spatial shards).
 * Bit-reproducible like the pooled statistics above: `scratch` must hold ie_channel_mean_scratch_bytes(...) bytes. */
long long ie_channel_mean_scratch_bytes(int n, int h, int w, int c, int y0, int y1);
int ie_channel_mean_nhwc_bf16(const void* x, int n, int h, int w, int c, int x_pitch, int x_coff, float* mean,
                              void* scratch, long long scratch_bytes, int y0, int y1, long long count, int layout,
                              void* stream);

/* tf.tile of a [n][c] vector to a k_h x k_w raster slice (Poolskip :111-112), border zeroed.          */
int ie_broadcast_hw_bf16(const float* vec, int n, int kh, int kw, int c, void* y, int y_pitch, int y_coff, int layout,
                         void* stream);

/* bf16 raster slice -> fp32 NHWC [n][h][w][c] (interior).  Debug / parity taps.                       */
int ie_raster_to_nhwc_f32(const void* x, int n, int h, int w, int c, int x_pitch, int x_coff, float* y, int layout,
                          void* stream);

/* ---- basis softmax + per-pixel filter -------------------------------------------------------------- */

/* softmax over the taps axis of [n][taps][b] (model_library.py:436-437).                              */
int ie_softmax_taps_f32(const float* originbasis, int n, int taps, int b, float* bas, void* stream);

/* cost_volume(Basis) (data_utils.py:97-113; printed by eval.py:159-162,189-191 when params["ps"]): per image
 *   -mean_{tap,t}( var_b bas[tap][t][b] ) + 0.1 * mean_{t,b}( (max(sum_tap bas[tap][t][b], 0.75) - 0.75)^2 ),
 * then the mean over the batch.  bas [n][taps = K*K][tb = T*B] fp32 with b = B; per_image [n] fp64 (scratch, holds
 * the per-image values afterwards); out [2] fp64 = {batch mean (the reference's scalar), sum over images (the
 * additive form that is all-reduced)}.                                                                  */
int ie_cost_volume_f32(const float* bas, int n, int taps, int tb, int b, double* per_image, double* out,
                       void* stream);

/* Filter synthesis + Convolve + Convolve_perlayer (model_library.py:439-451, 114-168), fused:
 *   out[n,y,x,1+t] = T * sum_b coef[n,y,x,b] * sum_{i,j} bas[n,i,j,t,b] * pad0(burst)[n,y+i-K/2,x+j-K/2,t]
 *   out[n,y,x,0]   = mean_t out[n,y,x,1+t]
 * burst: fp32 NHWC [n][h][w] with `burst_pitch` channels per pixel (first T used); coef [n][hc][wc][B]
 * with hc >= h, wc >= w (the network runs at the stride-padded size; only the top-left h x w is read);
 * bas [n][K][K][T][B]; out [n][h][w][T+1].  The per-pixel kernels are never materialised.            */
int ie_kpn_apply_f32(const float* burst, int burst_pitch, const float* coef, int hc, int wc, const float* bas,
                     float* out, int n, int h, int w, int T, int K, int B, void* stream);

/* Same contract on the tensor cores (warp-level TF32 MMA, fp32 accumulate) for K = 15, B <= 128, T <= 8: burst and
 * basis are rounded to TF32 (10-bit mantissa), so the result differs from ie_kpn_apply_f32 by < 2^-10 of the
 * pixel range - inside the path's tolerance (max-abs 1e-2 on [0,1] pixels), ~2.5x faster.  One launch handles 4
 * frames x 16 bases; longer bursts / more bases (Basis_kpn, remote/record.txt: B up to 90) run as further launches
 * that add into `out` (the output is a sum over the bases).                                               */
int ie_kpn_apply_tf32(const float* burst, int burst_pitch, const float* coef, int hc, int wc, const float* bas,
                      float* out, int n, int h, int w, int T, int K, int B, void* stream);

/* Same contract on tcgen05 in the reference's own order of operations (model_library.py:439-451): the per-pixel
 * filter is synthesised as a GEMM coef x Bas into TMEM and applied to the burst window in the epilogue
 * (csrc/kpn_tcgen05.cu).  MMA operands: coef and basis (softmax outputs in [0, 1]) as fp16 - the 10-bit mantissa of
 * TF32, 64 bases per 128-byte row; fp32 accumulation; the burst is not rounded: |err| <= 2^-10 of the pixel range.
 * K = 15, T a multiple of 4 (passes of four frames), any B (blocks of 64 bases; later passes / blocks add into `out`,
 * the first overwrites it).  The models' default filter where it applies (validated on B200 through this entry
 * point: tests/test_gpu_kernels.py::test_kpn_apply_tcgen05_variant).                                              */
int ie_kpn_apply_tc(const float* burst, int burst_pitch, const float* coef, int hc, int wc, const float* bas,
                    float* out, int n, int h, int w, int T, int K, int B, void* stream);

/* Convolve / cus_convolve / Convolve_perlayer (model_library.py:114-168) with MATERIALISED filters, for callers
 * that build `filts` [n][h][w][K][K][T] themselves (the models never do - they use ie_kpn_apply_f32):
 *   out[n,y,x,0]   = sum_{i,j,t} pad0(burst)[n,y+i-K/2,x+j-K/2,t] * filts[n,y,x,i,j,t]        (Convolve.call)
 *   out[n,y,x,1+t] = T * sum_{i,j} pad0(burst)[..,t] * filts[n,y,x,i,j,t]                     (Convolve_perlayer.call)
 * burst: fp32 NHWC with `burst_pitch` channels per pixel (first T used); out [n][h][w][T+1].             */
int ie_convolve_filts_f32(const float* burst, int burst_pitch, const float* filts, float* out, int n, int h, int w,
                          int T, int K, void* stream);

/* ---- metrics (data_utils.py:24-164 as eval.py:139-182 calls them) --------------------------------- */

/* mean over H,W of channel `coff` of an fp32 NHWC tensor: white level of eval.py:144-145. out [n].    */
int ie_mean_hw_f32(const float* x, int n, int h, int w, int pitch, int coff, float* out, void* stream);

/* invert_preproc (data_utils.py:42-45): sRGBforward(img / wl[n]) cropped by `crop` px per side.
 * img: mean of channels [coff, coff+nch) of fp32 NHWC with `pitch` channels (nch = 1: one channel;
 * nch = T: the burst average of psnr_average_f :162); out [n][h-2crop][w-2crop].                      */
int ie_invert_preproc_f32(const float* img, int pitch, int coff, int nch, const float* wl, int n, int h, int w,
                          int crop, float* out, void* stream);

/* One pass over recon [n][h][w][T+1], burst (pitch channels, first T), truth [n][h][w][2] producing,
 * per image, fp64 sums over the cropped sRGB'd images:
 *   sums[n][k]        k in [0,T+3): sum (e_k - gt)^2, e = {deblur, frame_0..T-1, burst0, burst mean}
 *   sums[n][T+3+k]    k in [0,T+1): sum |grad e_k - grad gt| (both components), e = {deblur, frames}
 * from which psnr_deblur / psnr_each_layer / psnr_burst0 / psnr_average_f (data_utils.py:121-164)
 * and deblur_loss / deblur_layer_loss (:52-96) follow.  `sums` must be zeroed by the caller.          */
int ie_eval_metrics_f32(const float* recon, const float* burst, int burst_pitch, const float* truth,
                        const float* wl, int n, int h, int w, int T, int crop, double* sums, void* stream);

/* Same, and as a by-product the two images the SSIM extension needs - invert_preproc(recon[...,0]) and
 * invert_preproc(truth[...,0]) (eval.py:146-149), which this kernel forms anyway - written as dense fp32
 * [n][h-2*crop][w-2*crop] crops: saves the two ie_invert_preproc_f32 passes in front of ie_ssim_f32.  Needs 16-byte
 * aligned recon / burst / truth (the row-streaming kernel); returns an error otherwise.                           */
int ie_eval_metrics_crops_f32(const float* recon, const float* burst, int burst_pitch, const float* truth,
                              const float* wl, int n, int h, int w, int T, int crop, double* sums, float* crop_deblur,
                              float* crop_gt, void* stream);

/* Tuning / A-B knob of ie_eval_metrics_f32 (tools/metric_sweep.py): rows per bulk-copy batch and warps per block of
 * the row-streaming kernel (0 = default), legacy != 0 selects the 32x32-tile kernel that also serves pointers that
 * are not 16-byte aligned.  Process-wide; not part of the reference-facing surface.                     */
int ie_eval_metrics_tune(int rows_per_batch, int warps, int legacy);

/* Per-image sums of ie_eval_metrics_f32 -> the additive totals one eval step contributes (fp64, T+6 values):
 *   [ sum_n psnr_deblur, sum_n psnr_frame_0..T-1, sum_n psnr_burst0, sum_n psnr_average,
 *     sum_n loss_deblur_n, sum_n loss_perlayer_n, n ],  psnr = -10 log10(mse) (data_utils.py:118-119),
 *   loss = mse + mean|grad diff| (:46-51).  This is the vector that is all-reduced across ranks.          */
int ie_metric_totals_f64(const double* sums, int n, int h, int w, int T, int crop, double* totals, void* stream);

/* Same with one more total for the SSIM EXTENSION (not in the reference): ssim_sums[n] = sum of image n's SSIM map
 * (ie_ssim_f32 on the cropped sRGB'd deblurred / ground-truth images); totals (T+7 values) =
 *   [ ...the T+5 values above..., sum_n mean-SSIM_n, n ].                                                  */
int ie_metric_totals_ssim_f64(const double* sums, const double* ssim_sums, int n, int h, int w, int T, int crop,
                              double* totals, void* stream);

/* psnr_tf_batch's inner reduction (data_utils.py:118-119): sums[n] += sum (a-b)^2 over `count` px.    */
int ie_sqdiff_sum_f32(const float* a, const float* b, int n, long long count, double* sums, void* stream);

/* basic_img_loss pieces (data_utils.py:37-51) on [n][h][w] pairs: sums[0] += sum (a-b)^2,
 * sums[1] += sum |grad a - grad b|.                                                                    */
int ie_img_loss_sums_f32(const float* a, const float* b, int n, int h, int w, double* sums, void* stream);

/* SSIM, tf.image.ssim semantics (11x11 Gaussian sigma 1.5, VALID, K1=.01, K2=.03, max_val 1).
 * EXTENSION: not in the reference.  sums[n] += sum of the SSIM map of image n ((h-10)*(w-10) values). */
int ie_ssim_f32(const float* a, const float* b, int n, int h, int w, double* sums, void* stream);
/* A-B knob: legacy != 0 forces the one-column-per-thread kernel that also serves odd widths and images that are
 * not 16-byte aligned.  Process-wide.                                                                   */
int ie_ssim_tune(int legacy);

/* ---- preprocessing (data_utils.py:198-265 arithmetic with explicit random draws) ------------------ */

/* src u8 [n][hs][ws][c]; per (image, frame) crop origin org[n][T][2] (y,x) in source pixels (may be
 * negative / outside: zero padded like make_first_truth :436-438); each output pixel is the `up` x `up`
 * box mean (AREA resize :459) of (src/255)^degamma averaged over c (:220), times wl[n] (:230);
 * noisy = truth + sqrt(truth)*sig_shot[n]*n_shot + sig_read[n]*n_read (:462-466; noise nullable);
 * x [n][h][w][T+add] = noisy ++ noise-level channel(s) (:256-264; layer_type 0 empty, 1 singlestd,
 * 2 dualparams); truth [n][h][w][2] = clean frame 0 ++ white level (:265).                            */
int ie_preprocess_u8(const uint8_t* src, int n, int hs, int ws, int c, const int32_t* org, int up, float degamma,
                     const float* wl, const float* sig_read, const float* sig_shot, const float* n_read,
                     const float* n_shot, int layer_type, int h, int w, int T, float* x, float* truth,
                     void* stream);

/* Same, with the read / shot noise normals of add_read_shot_tf (data_utils.py:462-466) drawn ON THE DEVICE:
 * Philox4x32-10 keyed by `seed`, counter = element index ((n*h + y)*w + x)*T + t, Box-Muller; element-wise
 * reproducible (numpy restatement: oracle/preprocess.py::philox_normals).  The reference draws from TF's RNG
 * stream, which is not reproducible outside TF: the distribution is the contract, not the bits.               */
int ie_preprocess_u8_rng(const uint8_t* src, int n, int hs, int ws, int c, const int32_t* org, int up, float degamma,
                         const float* wl, const float* sig_read, const float* sig_shot, unsigned long long seed,
                         int layer_type, int h, int w, int T, float* x, float* truth, void* stream);

/* A-B knob: legacy != 0 forces the one-pixel-per-thread kernel (any up / c / T / alignment); by default upscale 4,
 * grey source, w % 4 == 0, T in {4, 8} run the 4-pixels-per-thread kernel.  Process-wide.                   */
int ie_preprocess_tune(int legacy);

#ifdef __cplusplus
}
#endif
#endif /* IMGENH_B200_H */
