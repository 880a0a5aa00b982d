// Library-wide pieces of the C ABI: version, error string, device info, tensor-map encoding.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "ie_common.cuh"

namespace ie {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_pdl_env = -1;          // IE_PDL=0 in the environment disables the attribute
static int g_pdl_off = 0;           // ie_conv_set_mode flag bit 10 (tests / A-B timing)
bool pdl_enabled() {
  if (g_pdl_env < 0) {
    const char* e = getenv("IE_PDL");
    g_pdl_env = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl_env != 0 && !g_pdl_off;
}
void pdl_set(bool on) { g_pdl_off = on ? 0 : 1; }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Encoded tensor maps are cached by (kind, base pointer, shape, box): the descriptor depends on nothing else, an eager
// forward re-creates the same ~100 maps on every call (3 - 5 driver calls per convolution), and activation buffers
// are cached per input shape by the caller - so after the first forward every lookup hits.
struct TmapKey {
  uint64_t v[7];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint64_t x : k.v) h = (h ^ x) * 0xBF58476D1CE4E5B9ull + (h >> 29);
    return static_cast<size_t>(h);
  }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static bool tmap_lookup(const TmapKey& k, CUtensorMap* out) {
  std::lock_guard<std::mutex> lock(g_tmap_mu);
  auto it = g_tmap_cache.find(k);
  if (it == g_tmap_cache.end()) return false;
  *out = it->second;
  return true;
}
static void tmap_store(const TmapKey& k, const CUtensorMap& m) {
  std::lock_guard<std::mutex> lock(g_tmap_mu);
  if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
  g_tmap_cache[k] = m;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_elems,
                      uint32_t box_cols, uint32_t box_rows) {
  const TmapKey key{{1, reinterpret_cast<uint64_t>(base), cols, rows, pitch_elems, box_cols, box_rows}};
  if (tmap_lookup(key, out)) return IE_OK;
  EncodeTiledFn fn = encode_fn();
  IE_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  IE_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base %p is not 16-byte aligned", base);
  IE_REQUIRE((pitch_elems * 2) % 16 == 0, "tensor map: row pitch must be a multiple of 16 bytes");
  IE_REQUIRE(box_cols * 2 == 128 && box_rows >= 1 && box_rows <= 256, "tensor map: bad box %u x %u", box_cols, box_rows);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (cols %llu rows %llu pitch %llu box %ux%u)",
             (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch_elems, box_cols, box_rows);
  tmap_store(key, *out);
  return IE_OK;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static EncodeIm2colFn encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  });
  return fn;
}

// Rank-4 bf16 NHWC tensor [n][h][w][pitch_elems] in TMA im2col mode for a (2*pad+1)^2 'same' convolution: the pixel
// box runs from -pad to (size - 1 - pad) in W and H (lower corner -pad, upper corner -pad), so a load of
// `pixels` consecutive output pixels starting at base coordinates (x0 - pad, y0 - pad, n0) with offsets (j, i) fetches
// tap (i, j) of each of them, wrapping rows and images and zero-filling everything outside the image (measured on
// B200: tools/micro/tma_im2col.cu).  64 channels per pixel (one 128-byte swizzle row).
int make_tmap_im2col_bf16(CUtensorMap* out, const void* base, int n, int h, int w, uint64_t pitch_elems, int pad,
                          uint32_t pixels) {
  const TmapKey key{{2, reinterpret_cast<uint64_t>(base), (uint64_t)n, ((uint64_t)h << 32) | (uint32_t)w, pitch_elems,
                     (uint64_t)pad, pixels}};
  if (tmap_lookup(key, out)) return IE_OK;
  EncodeIm2colFn fn = encode_im2col_fn();
  IE_REQUIRE(fn != nullptr, "cuTensorMapEncodeIm2col is not available from this driver");
  IE_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base %p is not 16-byte aligned", base);
  IE_REQUIRE((pitch_elems * 2) % 16 == 0 && pitch_elems >= 64, "tensor map: bad pixel pitch %llu", (unsigned long long)pitch_elems);
  IE_REQUIRE(pixels >= 1 && pixels <= 1024 && pad >= 0 && pad <= 64, "tensor map: bad im2col box (%u pixels, pad %d)", pixels, pad);
  cuuint64_t gdim[4] = {pitch_elems, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t gstride[3] = {pitch_elems * 2, (cuuint64_t)w * pitch_elems * 2, (cuuint64_t)h * w * pitch_elems * 2};
  int lo[2] = {-pad, -pad}, hi[2] = {-pad, -pad};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, lo, hi, 64, pixels, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed with CUresult %d (n %d h %d w %d pitch %llu pad %d)", (int)r, n,
             h, w, (unsigned long long)pitch_elems, pad);
  tmap_store(key, *out);
  return IE_OK;
}

}  // namespace ie

extern "C" int ie_version(void) { return IE_VERSION; }
extern "C" const char* ie_last_error(void) { return ie::g_err; }
extern "C" int ie_sm_count(void) { return ie::sm_count(); }
