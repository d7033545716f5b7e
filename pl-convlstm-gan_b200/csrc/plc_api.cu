// C ABI of libplc.so (declared in include/plc.h): descriptor validation, weight packing, tensor-map
// construction and kernel launches.  No torch types, no persistent device allocations.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/plc.h"
#include "conv_igemm_tc.cuh"
#include "conv_simt.cuh"
#include "frame_conv.cuh"
#include "frame_io.cuh"
#include "frontend_tc.cuh"
#include "loss.cuh"
#include "wgrad_tc.cuh"

namespace {

thread_local std::string g_err;
unsigned long long* g_prof_buf = nullptr;   // debug cycle counters (plc_debug_set_prof)

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define PLC_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) return fail(PLC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

constexpr int kMaxDevices = 64;

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

// SM count of the CURRENT device (one process per GPU is the normal deployment, but nothing here assumes device 0)
int sm_count() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device) instead of once per launch.
// `Tag` makes the flag array unique per call site / template instantiation.
template <int V> struct IntTag {};
struct TagW1 {};
struct TagW2 {};
template <typename Tag, typename Fn>
cudaError_t set_smem_once(Fn fn, int bytes) {
  static std::atomic<int> done[kMaxDevices];
  const int dev = current_device();
  if (done[dev].load(std::memory_order_acquire) >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done[dev].store(bytes, std::memory_order_release);
  return e;
}

// ------------------------------------------------------------------ per-launch timing (plc_timing_*)
// When enabled, every kernel launch of the library is bracketed by a pair of CUDA events on the launching stream and
// tagged with its kind; bench.py reads the list back to attribute the step time to kernels and to compute the
// roofline of the dominant one from live per-launch durations.  Off by default: zero cost beyond one relaxed load.
std::atomic<int> g_timing_on{0};
std::mutex g_timing_mu;
struct TimedLaunch { cudaEvent_t e0, e1; int kind; double flops; };
std::vector<TimedLaunch> g_timed;

struct LaunchTimer {
  cudaStream_t st;
  int kind;
  double flops;            // algorithmic FLOPs of the launch (2*M*N*K over the TRUE channel counts; 0 = not a contraction)
  cudaEvent_t e0 = nullptr;
  LaunchTimer(int kind_, cudaStream_t st_, double flops_ = 0.0) : st(st_), kind(kind_), flops(flops_) {
    if (g_timing_on.load(std::memory_order_relaxed)) {
      if (cudaEventCreate(&e0) != cudaSuccess || cudaEventRecord(e0, st) != cudaSuccess) { cudaGetLastError(); e0 = nullptr; }
    }
  }
  ~LaunchTimer() {
    if (!e0) return;
    cudaEvent_t e1 = nullptr;
    if (cudaEventCreate(&e1) != cudaSuccess || cudaEventRecord(e1, st) != cudaSuccess) {
      cudaGetLastError();
      cudaEventDestroy(e0);
      return;
    }
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_timed.push_back({e0, e1, kind, flops});
  }
};

// ------------------------------------------------------------------ tensor-map cache
// cuTensorMapEncodeTiled is a pure host-side encoder of (pointer, dims, strides, box, swizzle): the result can be
// reused whenever the same buffer is passed with the same geometry -- which is every call of a rollout after the first
// (state rings, workspaces and packed weights are preallocated).  Thread-local, direct-mapped, no eviction policy.
struct TmapKey {
  const void* ptr;
  long long a, b;       // leading dims (B / rows, cols)
  int v[8];
};
struct TmapEntry { TmapKey key; CUtensorMap tm; bool valid; };
constexpr int kTmapCacheSize = 1024;
thread_local TmapEntry g_tmap_cache[kTmapCacheSize];

inline unsigned tmap_hash(const TmapKey& k) {
  unsigned long long h = reinterpret_cast<unsigned long long>(k.ptr) * 0x9E3779B97F4A7C15ull;
  h ^= static_cast<unsigned long long>(k.a) * 0xC2B2AE3D27D4EB4Full + static_cast<unsigned long long>(k.b) * 0x165667B19E3779F9ull;
  for (int i = 0; i < 8; ++i) h = (h ^ static_cast<unsigned>(k.v[i])) * 0x100000001B3ull;
  return static_cast<unsigned>(h >> 40) & (kTmapCacheSize - 1);
}
inline bool tmap_lookup(const TmapKey& k, CUtensorMap* tm) {
  const TmapEntry& e = g_tmap_cache[tmap_hash(k)];
  if (e.valid && memcmp(&e.key, &k, sizeof(TmapKey)) == 0) { *tm = e.tm; return true; }
  return false;
}
inline void tmap_store(const TmapKey& k, const CUtensorMap& tm) {
  TmapEntry& e = g_tmap_cache[tmap_hash(k)];
  e.key = k; e.tm = tm; e.valid = true;
}

// NHWC activation [B,(T,)H,W,C] as a 4-D (T == 0) or 5-D tensor map [C, W, H, (T,) B]; box = [box_c, tw, th, 1(, 1)]
// pixels.  `estride` > 1 makes the box pick every estride-th pixel along W and H (strided convolutions: the box of a
// filter tap covers tw x th OUTPUT pixels, i.e. an input extent of tw*estride x th*estride).
int make_tmap_act(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int tw, int th, int esize = 2,
                  int box_c = 64, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int estride = 1, int T = 0) {
  TmapKey key{};
  key.ptr = ptr; key.a = B; key.b = T;
  key.v[0] = H; key.v[1] = W; key.v[2] = C; key.v[3] = tw; key.v[4] = th; key.v[5] = esize | (estride << 8);
  key.v[6] = box_c; key.v[7] = static_cast<int>(swz) | 0x100;
  if (tmap_lookup(key, tm)) return PLC_OK;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(PLC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t es = (cuuint64_t)esize;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 1};
  cuuint64_t strides[4] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es, 0};
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)(tw * estride), (cuuint32_t)(th * estride), 1, 1};
  cuuint32_t estr[5] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1, 1};
  int rank = 4;
  if (T > 0) {
    rank = 5;
    dims[3] = (cuuint64_t)T; dims[4] = (cuuint64_t)B;
    strides[3] = (cuuint64_t)T * H * W * C * es;
  }
  if (box[1] > 256 || box[2] > 256) return fail(PLC_ERR_UNSUPPORTED, "TMA box %ux%u exceeds 256", box[1], box[2]);
  CUresult r = enc(tm, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                   const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(PLC_ERR_CUDA, "cuTensorMapEncodeTiled(act B=%d T=%d H=%d W=%d C=%d box=%dx%d stride=%d) -> %d", B, T, H,
                W, C, tw, th, estride, (int)r);
  tmap_store(key, *tm);
  return PLC_OK;
}

// row-major bf16 matrix [rows, cols] (cols contiguous); box = [box_cols, box_rows], SWIZZLE_128B.
int make_tmap_mat(CUtensorMap* tm, const void* ptr, long rows, long cols, int box_cols, int box_rows) {
  TmapKey key{};
  key.ptr = ptr; key.a = rows; key.b = cols;
  key.v[0] = box_cols; key.v[1] = box_rows; key.v[7] = 0x200;
  if (tmap_lookup(key, tm)) return PLC_OK;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(PLC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(PLC_ERR_CUDA, "cuTensorMapEncodeTiled(mat %ldx%ld box=%dx%d) -> %d", rows, cols, box_rows, box_cols,
                (int)r);
  tmap_store(key, *tm);
  return PLC_OK;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// algorithmic FLOPs of a conv-shaped contraction over B*H*W output pixels: 2 * M * Cout * (Cin * k^2)
inline double conv_flops(int B, int H, int W, int cin, int cout, int k) {
  return 2.0 * B * H * W * static_cast<double>(cin) * cout * k * k;
}

int check_desc(const PlcCellDesc* d) {
  if (!d) return fail(PLC_ERR_BAD_DESC, "null descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Ch <= 0 || d->Cin < 0)
    return fail(PLC_ERR_BAD_DESC, "bad sizes B=%d H=%d W=%d Cin=%d Ch=%d", d->B, d->H, d->W, d->Cin, d->Ch);
  if (d->k <= 0 || (d->k % 2) == 0)
    return fail(PLC_ERR_BAD_DESC, "kernel size %d must be odd (reference padding k//2 breaks even k, convlstm.py:12)",
                d->k);
  if (d->mode != PLC_MODE_BF16_TC && d->mode != PLC_MODE_FP32) return fail(PLC_ERR_BAD_DESC, "bad mode %d", d->mode);
  if ((long long)d->B * d->H * d->W >= (1ll << 31)) return fail(PLC_ERR_UNSUPPORTED, "B*H*W must be < 2^31");
  if (d->mode == PLC_MODE_BF16_TC) {
    if (d->Cin % 8) return fail(PLC_ERR_ALIGNMENT, "bf16 mode needs Cin %% 8 == 0 (got %d): pad channels", d->Cin);
    if (d->Ch % 16) return fail(PLC_ERR_ALIGNMENT, "bf16 mode needs Ch %% 16 == 0 (got %d)", d->Ch);
    if (d->Ch > 256) return fail(PLC_ERR_UNSUPPORTED, "bf16 mode supports Ch <= 256 (got %d)", d->Ch);
    if (d->k > 7) return fail(PLC_ERR_UNSUPPORTED, "bf16 mode supports k <= 7 (got %d)", d->k);
  }
  return PLC_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------ tensor-core tiling choices
struct TcGeom {
  int ch_tile, n_tile;       // LSTM: N_TILE = 4*CH_TILE
  int tw, th, tw_log2, tiles_x, tiles_y;
};

int pick_ch_tile(int Ch) {
  const int cands[4] = {64, 48, 32, 16};
  for (int c : cands)
    if (Ch % c == 0) return c;
  return 16;
}

void pick_spatial_tile(int H, int W, TcGeom* g, int pixels_log2 = 7) {
  long best = -1;
  for (int l2 = pixels_log2; l2 >= 0; --l2) {
    const int tw = 1 << l2, th = (1 << pixels_log2) >> l2;
    const long area = (long)cdiv(W, tw) * tw * cdiv(H, th) * th;
    if (best < 0 || area < best) {
      best = area;
      g->tw = tw; g->th = th; g->tw_log2 = l2;
    }
  }
  g->tiles_x = cdiv(W, g->tw);
  g->tiles_y = cdiv(H, g->th);
}

int pick_plain_n_tile(int n_total) {
  if (n_total <= 64) return 64;
  if (n_total <= 128) return 128;
  if (n_total <= 192) return 192;
  if (n_total <= 256) return 256;
  // several N tiles: prefer the candidate with the least padding
  int best = 256, best_pad = cdiv(n_total, 256) * 256 - n_total;
  const int cands[3] = {192, 128, 64};
  for (int c : cands) {
    const int pad = cdiv(n_total, c) * c - n_total;
    if (pad < best_pad) { best = c; best_pad = pad; }
  }
  return best;
}

// K-side geometry: sources with C0 and C1 channels are read as TMA boxes of kc channels (64, 32 or 16); G = 64/kc
// boxes (consecutive (source, tap, chunk) triples) fill one 64-element K stage.  Pick the kc with the fewest stages.
struct KGeom { int kc, chunks0, chunks1, num_boxes, num_kb; };
KGeom kgeom_taps(int C0, int C1, int taps);
KGeom kgeom(int C0, int C1, int k) { return kgeom_taps(C0, C1, k * k); }
KGeom kgeom_taps(int C0, int C1, int taps) {
  KGeom best{};
  const int cands[3] = {64, 32, 16};
  for (int kc : cands) {
    KGeom g;
    g.kc = kc;
    g.chunks0 = cdiv(C0, kc);
    g.chunks1 = cdiv(C1, kc);
    g.num_boxes = taps * (g.chunks0 + g.chunks1);
    g.num_kb = cdiv(g.num_boxes, 64 / kc);
    if (best.kc == 0 || g.num_kb < best.num_kb) best = g;
  }
  return best;
}
CUtensorMapSwizzle swizzle_for_kc(int kc) {
  return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}
// packed K index kp -> input channel of cat(src0, src1) and tap; returns -1 for padding
__device__ __forceinline__ int decode_packed_k(int kp, int kc, int kk, int chunks0, int chunks1, int C0, int C1,
                                               int& tap) {
  const int G = 64 / kc;
  const int box = (kp >> 6) * G + (kp & 63) / kc, jj = (kp & 63) % kc;
  tap = 0;
  if (box >= kk * (chunks0 + chunks1)) return -1;
  if (box < kk * chunks0) {
    tap = box / chunks0;
    const int c = (box % chunks0) * kc + jj;
    return c < C0 ? c : -1;
  }
  const int b1 = box - kk * chunks0;
  tap = b1 / chunks1;
  const int c = (b1 % chunks1) * kc + jj;
  return c < C1 ? C0 + c : -1;
}

// ------------------------------------------------------------------ bf16 packing kernels
// forward image Wp[n'][k'] : n' = slice*N_TILE + gate*CH_TILE + j ; k' = kb*64 + jj, kb = (src, tap, chunk)
__global__ void pack_w_tc_fwd_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cin, int Ch,
                                     int ksize, int ch_tile, KGeom kg) {
  const int kk = ksize * ksize, ctot = Cin + Ch;
  const int ktot = kg.num_kb * 64, ntot = 4 * Ch, n_tile = 4 * ch_tile;
  const size_t total = static_cast<size_t>(ntot) * ktot;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int kp = idx % ktot, np = idx / ktot;
    const int slice = np / n_tile, r = np % n_tile, gate = r / ch_tile, j = r % ch_tile;
    const int o = gate * Ch + slice * ch_tile + j;
    int tap;
    const int ic = decode_packed_k(kp, kg.kc, kk, kg.chunks0, kg.chunks1, Cin, Ch, tap);
    float v = 0.f;
    if (ic >= 0) v = w[(static_cast<size_t>(o) * ctot + ic) * kk + tap];
    out[idx] = __float2bfloat16(v);
  }
}
// generic conv forward image Wp[n'][k'] : k' = (tap, 64-chunk)*64 + jj ; packed row n' = sub*(Cout/4) + c for the
// PixelShuffle(2) store (natural channel n = c*4 + sub), else n' = n.  bias_p[n'] = bias[n].
__global__ void pack_w_conv_fwd_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                       __nv_bfloat16* __restrict__ out, float* __restrict__ bias_p, int Cin, int Cout,
                                       int kk /* taps */, KGeom kg, int shuffle) {
  const int ktot = kg.num_kb * 64, cps = Cout >> 2;
  const size_t total = static_cast<size_t>(Cout) * ktot;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int kp = idx % ktot, np = idx / ktot;
    const int n = shuffle ? (np % cps) * 4 + np / cps : np;
    int tap;
    const int c = decode_packed_k(kp, kg.kc, kk, kg.chunks0, 0, Cin, 0, tap);
    float v = 0.f;
    if (c >= 0) v = w[(static_cast<size_t>(n) * Cin + c) * kk + tap];
    out[idx] = __float2bfloat16(v);
    if (kp == 0 && bias_p) bias_p[np] = bias ? bias[n] : 0.f;
  }
}
// generic dgrad image Wd[c][k'] : rows c in [0, ctot) ; k' = (tap', chunk of dZ channel n)*64 + jj, flipped taps
__global__ void pack_w_conv_dgrad_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int ctot,
                                         int nout, int kk /* taps: k*k or kt*k*k */, KGeom kg) {
  const int ktot = kg.num_kb * 64;
  const size_t total = static_cast<size_t>(ctot) * ktot;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int kp = idx % ktot, c = idx / ktot;
    int tap;
    const int n = decode_packed_k(kp, kg.kc, kk, kg.chunks0, 0, nout, 0, tap);
    float v = 0.f;
    if (n >= 0) {
      // every tap coordinate flipped == the row-major tap index reversed
      v = w[(static_cast<size_t>(n) * ctot + c) * kk + (kk - 1 - tap)];
    }
    out[idx] = __float2bfloat16(v);
  }
}
// dZ (natural conv-output channel order) = dY * (Y > 0), optionally undoing PixelShuffle(2):
//   shuffle: y, dy are [B, 2H, 2W, Cout/4]; dz[b,y,x, c*4 + sub] = dy[b, 2y+sub/2, 2x+sub%2, c] * (y[...] > 0)
__global__ void conv_grad_mask_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dy,
                                      __nv_bfloat16* __restrict__ dz, size_t npix_in, int H, int W, int Cout, int relu,
                                      int shuffle) {
  const size_t total = npix_in * Cout;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t src = idx;
    if (shuffle) {
      const int n = idx % Cout;
      const size_t pix = idx / Cout;
      const int x = pix % W;
      const size_t t = pix / W;
      const int yy = t % H;
      const size_t b = t / H;
      const int c = n >> 2, sub = n & 3;
      src = ((b * (2 * H) + 2 * yy + (sub >> 1)) * (2 * W) + 2 * x + (sub & 1)) * (Cout >> 2) + c;
    }
    const float g = __bfloat162float(dy[src]);
    const bool on = !relu || __bfloat162float(y[src]) > 0.f;
    dz[idx] = __float2bfloat16(on ? g : 0.f);
  }
}

// no-shuffle form, 8 channels (16 bytes) per thread: dz = dy * (y > 0); the scalar kernel above ran at 1.5 TB/s on the
// front-end's 671 MB tensors (1.38 ms per training step at cfg3)
__global__ void conv_grad_mask_vec_kernel(const uint4* __restrict__ y, const uint4* __restrict__ dy, uint4* __restrict__ dz,
                                          size_t total) {
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 g = dy[idx], yv = y[idx];
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w};
    uint32_t ow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // bf16 "value > 0" on the raw halves: 0x0001 .. 0x7f80 (positive, not zero, not NaN)
      const uint32_t lo = yw[i] & 0xffffu, hi = yw[i] >> 16;
      const uint32_t keep = ((lo - 1u) < 0x7f80u ? 0x0000ffffu : 0u) | ((hi - 1u) < 0x7f80u ? 0xffff0000u : 0u);
      ow[i] = gw[i] & keep;
    }
    dz[idx] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// im2col of a NARROW conv input for the weight gradient: x [B,H,W,Cp] bf16 (Cp = 8: the true channels c < C, zero
// padded) -> col [B,H,W,32] bf16 with col[pix][tap * C + c] = x[pix + tap offset][c] (zero outside the image, zero
// beyond k*k*C).  One thread per pixel: k*k 16-byte loads (neighbouring threads share them through L1), 64 bytes out.
__global__ void im2col_narrow_kernel(const __nv_bfloat16* __restrict__ x, uint4* __restrict__ col, int B, int H, int W,
                                     int C, int k) {
  // per-column source (row offset, column offset, element offset); columns >= k*k*C read nothing
  __shared__ int s_dy[32], s_dx[32], s_off[32];
  const int pad = k / 2, ncol = k * k * C;
  if (threadIdx.x < 32) {
    const int i = threadIdx.x;
    const int tap = i < ncol ? i / C : 0, c = i < ncol ? i - tap * C : 0;
    const int ky = tap / k, kx = tap - ky * k;
    s_dy[i] = i < ncol ? ky - pad : (1 << 20);       // far outside any image: the bounds test rejects it
    s_dx[i] = kx - pad;
    s_off[i] = ((ky - pad) * W + (kx - pad)) * 8 + c;
  }
  __syncthreads();
  const size_t npix = static_cast<size_t>(B) * H * W;
  const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < npix;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int xx = pix % W;
    const int yy = (pix / W) % H;
    const unsigned short* base = xs + pix * 8;
    uint32_t out[16];
#pragma unroll
    for (int i = 0; i < 32; ++i) {          // static register indices
      const unsigned sy = static_cast<unsigned>(yy + s_dy[i]), sx = static_cast<unsigned>(xx + s_dx[i]);
      const unsigned short v = (sy < static_cast<unsigned>(H) && sx < static_cast<unsigned>(W)) ? __ldg(base + s_off[i]) : 0;
      if (i & 1) out[i >> 1] |= static_cast<uint32_t>(v) << 16; else out[i >> 1] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) col[pix * 4 + i] = make_uint4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
  }
}

// ------------------------------------------------------------------ layout kernels
// src [B, Cs, H*W] fp32  ->  dst [B, H*W, Cd] bf16 (channels >= Cs zero-filled); 32x32 smem transpose
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Cs,
                                             int Cd, int HW) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, pp = p0 + threadIdx.x;
    t[i][threadIdx.x] = (c < Cs && pp < HW) ? src[(static_cast<size_t>(b) * Cs + c) * HW + pp] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pp = p0 + i, c = c0 + threadIdx.x;
    if (pp < HW && c < Cd) dst[(static_cast<size_t>(b) * HW + pp) * Cd + c] = __float2bfloat16(t[threadIdx.x][i]);
  }
}
__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int C,
                                             int HW) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pp = p0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (pp < HW && c < C) ? __bfloat162float(src[(static_cast<size_t>(b) * HW + pp) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, pp = p0 + threadIdx.x;
    if (c < C && pp < HW) dst[(static_cast<size_t>(b) * C + c) * HW + pp] = t[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------ tensor-core launches
// cta_group choice: a CTA pair (cta_group::2) halves the per-SM weight-tile traffic and wins once there are
// enough 128-pixel tiles to keep all 74 pairs busy; tiny problems keep 148 independent CTAs.
// PLC_CTA_GROUP=1|2 overrides (used by the benchmarks to A/B the two paths).
// Programmatic dependent launch for the persistent tensor-core kernels (PLC_PDL=0 turns it off for A/B runs): consecutive
// cell-step kernels of a rollout form a dependent chain, so without it every launch pays grid drain + launch latency +
// prologue back to back (~19 us fixed per launch measured at the 8-GPU shard of cfg3, where a cell kernel is ~55 us).
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("PLC_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
std::atomic<int> g_cta_override{0};   // plc_debug_set_cta_group (process-wide on purpose: backward runs on autograd threads)
int pick_cta_group(int num_m_tiles) {
  static const int forced = [] {
    const char* e = getenv("PLC_CTA_GROUP");
    return (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
  }();
  const int ov = g_cta_override.load(std::memory_order_relaxed);
  if (ov) return ov;
  if (forced) return forced;
  return num_m_tiles >= 2 * sm_count() ? 2 : 1;
}


template <int NT, int EPI, int CTA>
int launch_conv_tc_inst(const plc::ConvTcParams& p_in, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                        const CUtensorMap& o0, const CUtensorMap& o1, cudaStream_t st, int kind, double flops,
                        const plc::GateMaps* gm) {
  using Cfg = plc::ConvTcCfg<NT, CTA, EPI>;
  plc::ConvTcParams p = p_in;
  p.prof = g_prof_buf;
  auto kfn = plc::conv_igemm_tc_kernel<NT, EPI, CTA>;
  {  // fast exact division constants for the tile decode (valid while every dividend stays below 2^20)
    auto magic = [](int d) -> unsigned long long { return ((1ull << 40) + d - 1) / static_cast<unsigned long long>(d); };
    const bool small = static_cast<long long>(p.num_m_tiles + 1) * p.num_n_tiles < (1 << 20);
    p.div_n_tiles = small ? magic(p.num_n_tiles) : 0;
    p.div_tiles_x = small ? magic(p.tiles_x) : 0;
    p.div_tiles_y = small ? magic(p.tiles_y) : 0;
  }
  if (p.patch) {   // carve the pipeline region into patch slots + weight stages
    const int region = Cfg::kPipeBytes;
    if (p.kc < 64) p.patch_slots = 4;    // narrow single-unit tiles are short: several tiles of look-ahead
    else p.patch_slots = (region - 3 * p.patch_slot_bytes) / Cfg::kBBytes >= 4 ? 3 : 2;
    int nb = (region - p.patch_slots * p.patch_slot_bytes) / Cfg::kBBytes;
    if (nb < 2) return fail(PLC_ERR_UNSUPPORTED, "patch mode does not fit (N_TILE=%d cta=%d)", NT, CTA);
    p.b_stages = nb > 8 ? 8 : nb;
  }
  PLC_CUDA(set_smem_once<Cfg>(kfn, Cfg::kSmemBytes));
  const int tiles = CTA == 2 ? ((p.num_m_tiles + 1) / 2) * p.num_n_tiles : p.num_tiles;
  const int slots = sm_count() / CTA;
  const int grid = (tiles < slots ? tiles : slots) * CTA;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(plc::conv_tc_threads<EPI>());
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  LaunchTimer timer(kind, st, flops);
  static const plc::GateMaps no_gate_maps{};     // only the gate-gradient epilogue reads them
  PLC_CUDA(cudaLaunchKernelEx(&cfg, kfn, p, a0, a1, b, o0, o1, gm ? *gm : no_gate_maps));
  return PLC_OK;
}

// o0 / o1: output tensor maps of the TMA-store epilogue (forward, N_TILE = 256); ignored by the other instantiations
template <int EPI>
int launch_conv_tc(int n_tile, int cta, const plc::ConvTcParams& p, const CUtensorMap& a0, const CUtensorMap& a1,
                   const CUtensorMap& b, const CUtensorMap& o0, const CUtensorMap& o1, cudaStream_t st, int kind,
                   double flops, const plc::GateMaps* gm = nullptr) {
#define PLC_LAUNCH_TC(NT)                                                                         \
  case NT:                                                                                        \
    return cta == 2 ? launch_conv_tc_inst<NT, EPI, 2>(p, a0, a1, b, o0, o1, st, kind, flops, gm)  \
                    : launch_conv_tc_inst<NT, EPI, 1>(p, a0, a1, b, o0, o1, st, kind, flops, gm);
  switch (n_tile) {
    PLC_LAUNCH_TC(64)
    PLC_LAUNCH_TC(128)
    PLC_LAUNCH_TC(192)
    PLC_LAUNCH_TC(256)
    default:
      return fail(PLC_ERR_UNSUPPORTED, "no tensor-core kernel for N_TILE=%d", n_tile);
  }
#undef PLC_LAUNCH_TC
}

void fill_geom(const PlcCellDesc* d, const TcGeom& g, plc::ConvTcParams* p) {
  memset(p, 0, sizeof(*p));
  p->B = d->B; p->H = d->H; p->W = d->W;
  p->ksize = d->k; p->pad = d->k / 2;
  p->stride = 1; p->kt = 1; p->stride_t = 1; p->T_out = 1;      // plain 2-D stride-1 conv unless the caller says otherwise
  p->ky_n = p->kx_n = d->k; p->tap_x0 = p->tap_y0 = -(d->k / 2); p->tap_t0 = 0; p->tap_dir = 1;
  p->tw = g.tw; p->th = g.th; p->tw_log2 = g.tw_log2;
  p->tiles_x = g.tiles_x; p->tiles_y = g.tiles_y;
  p->num_m_tiles = d->B * g.tiles_x * g.tiles_y;
  p->Ch = d->Ch; p->Cin = d->Cin;
}

void set_kgeom(plc::ConvTcParams* p, const KGeom& kg) {
  p->kc = kg.kc; p->chunks0 = kg.chunks0; p->chunks1 = kg.chunks1; p->num_boxes = kg.num_boxes; p->num_kb = kg.num_kb;
}

// Haloed-patch pipeline (conv_igemm_tc.cuh, "patch mode"): applies to 3x3 / 5x5 kernels whose sources are read as
// 64-channel boxes (kc == 64), or as ONE narrower box (a single source of <= 32 channels, e.g. the frame front-end); it needs 16 x 8-pixel tiles, so it is skipped when that tiling wastes > 5 % more
// pixels than the default one.  PLC_PATCH=0|1 / plc_debug_set_patch override (A/B runs, parity tests of both paths).
std::atomic<int> g_patch_override{-1};   // plc_debug_set_patch
// bytes of the operand pipeline region of one instantiation (what patch mode re-carves into patch slots + weight stages)
template <int EPI, int CTA>
int pipeline_region_nt(int nt, int* b_bytes) {
#define PLC_REGION(NT)                                       \
  case NT: {                                                 \
    using Cfg = plc::ConvTcCfg<NT, CTA, EPI>;                \
    *b_bytes = Cfg::kBBytes;                                 \
    return Cfg::kPipeBytes;                                  \
  }
  switch (nt) { PLC_REGION(64) PLC_REGION(128) PLC_REGION(192) PLC_REGION(256) }
#undef PLC_REGION
  *b_bytes = 1;
  return 0;
}
int pipeline_region(int epi, int nt, int cta, int* b_bytes) {
  if (epi == plc::EPI_LSTM_FWD)
    return cta == 2 ? pipeline_region_nt<plc::EPI_LSTM_FWD, 2>(nt, b_bytes) : pipeline_region_nt<plc::EPI_LSTM_FWD, 1>(nt, b_bytes);
  if (epi == plc::EPI_LSTM_BWD_GATES)
    return cta == 2 ? pipeline_region_nt<plc::EPI_LSTM_BWD_GATES, 2>(nt, b_bytes)
                    : pipeline_region_nt<plc::EPI_LSTM_BWD_GATES, 1>(nt, b_bytes);
  return cta == 2 ? pipeline_region_nt<plc::EPI_PLAIN, 2>(nt, b_bytes) : pipeline_region_nt<plc::EPI_PLAIN, 1>(nt, b_bytes);
}
void maybe_patch(TcGeom* g, plc::ConvTcParams* p, int epi, int n_tile) {
  static const int env_mode = [] {
    const char* e = getenv("PLC_PATCH");
    return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1;
  }();
  const int ov = g_patch_override.load(std::memory_order_relaxed);
  const int mode = ov >= 0 ? ov : env_mode;
  p->patch = 0;
  if (mode == 0 || (p->ksize != 3 && p->ksize != 5)) return;
  // 64-channel boxes: any number of (source, chunk) units; narrower boxes: a single unit (one narrow source)
  if (p->kc != 64 && p->chunks0 + p->chunks1 != 1) return;
  const long area_def = static_cast<long>(g->tiles_x) * g->tw * g->tiles_y * g->th;
  const long area_patch = static_cast<long>(cdiv(p->W, 8)) * 8 * cdiv(p->H, 16) * 16;
  if (mode != 1 && area_patch * 100 > area_def * 105) return;
  const int slot_bytes = ((16 + 2 * p->pad) * (8 + 2 * p->pad) * 128 + 1023) / 1024 * 1024;
  if (mode != 1) {   // needs room for 2 patch slots + at least 3 weight stages (single-CTA N_TILE = 256 tiles do not)
    int b_bytes = 1;
    const int region = pipeline_region(epi, n_tile, pick_cta_group(p->B * cdiv(p->W, 8) * cdiv(p->H, 16)), &b_bytes);
    if ((region - 2 * slot_bytes) / b_bytes < 3) return;
  }
  g->tw = 8; g->th = 16; g->tw_log2 = 3;
  g->tiles_x = cdiv(p->W, 8); g->tiles_y = cdiv(p->H, 16);
  p->tw = 8; p->th = 16; p->tw_log2 = 3;
  p->tiles_x = g->tiles_x; p->tiles_y = g->tiles_y;
  p->num_m_tiles = p->B * g->tiles_x * g->tiles_y;
  p->num_tiles = p->num_m_tiles * p->num_n_tiles;
  p->patch = 1;
  p->patch_slot_bytes = slot_bytes;
}
// activation tensor map of an A source: the plain [kc ch, tw, th] tile box, or the haloed patch box in patch mode
int make_tmap_src(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, const plc::ConvTcParams& p) {
  // patch boxes are always 64 channels wide (128-byte SWIZZLE_128B rows): narrow sources (kc < 64) are zero-filled
  // beyond their channels -- the 32/64-byte swizzle modes cost ~4x the tensor-core issue time per UMMA
  if (p.patch) return make_tmap_act(tm, ptr, B, H, W, C, p.tw + 2 * p.pad, p.th + 2 * p.pad, 2, 64);
  return make_tmap_act(tm, ptr, B, H, W, C, p.tw, p.th, 2, p.kc, swizzle_for_kc(p.kc));
}

// EPI_PLAIN outputs through TMA tensor stores: needs 64-channel-aligned destinations (the staging boxes are
// [128 px][64 ch]) and no PixelShuffle.  q.out0 / q.out1 / q.Cin / q.n_total / tile geometry must be final.
int setup_plain_stores(plc::ConvTcParams* q, int B, int H, int W, CUtensorMap* o0, CUtensorMap* o1) {
  q->plain_tma = 0;
  const int c0 = q->Cin < q->n_total ? q->Cin : q->n_total, c1 = q->n_total - c0;
  if (q->plain_shuffle || (c0 % 64) || (c1 % 64)) return PLC_OK;
  if ((c0 > 0 && !q->out0) && (c1 == 0 || !q->out1)) return PLC_OK;   // nothing to store
  int rc;
  if (c0 > 0 && q->out0 && (rc = make_tmap_act(o0, q->out0, B, H, W, c0, q->tw, q->th, 2, 64))) return rc;
  if (c1 > 0 && q->out1 && (rc = make_tmap_act(o1, q->out1, B, H, W, c1, q->tw, q->th, 2, 64))) return rc;
  q->plain_tma = 1;
  return PLC_OK;
}

int lstm_tc_setup(const PlcCellDesc* d, TcGeom* g, plc::ConvTcParams* p, int epi) {
  g->ch_tile = pick_ch_tile(d->Ch);
  g->n_tile = 4 * g->ch_tile;
  pick_spatial_tile(d->H, d->W, g);
  fill_geom(d, *g, p);
  p->num_n_tiles = d->Ch / g->ch_tile;
  p->num_tiles = p->num_m_tiles * p->num_n_tiles;
  set_kgeom(p, kgeom(d->Cin, d->Ch, d->k));
  maybe_patch(g, p, epi, g->n_tile);
  return PLC_OK;
}

// wgrad on the tensor cores + bias-gradient column sum (bf16 mode)
// sources x [.,Cin] and h [.,Ch]; N = dZ channels.  B, H, W = the OUTPUT grid (dZ).  Strided / 3-D convs: the source
// lives on its own grid [Bs, Ts, Hs, Ws] and is read through strided (and, with kt > 1 or stride_t > 1, 5-D) maps.
struct WgradShape {
  int B, H, W, k, Cin, Ch, N;
  int stride = 1, kt = 1, stride_t = 1, T_out = 1;
  int Bs = 0, Ts = 0, Hs = 0, Ws = 0;      // source grid (0 = same as the output grid)
};

int launch_wgrad_tc_shape(const WgradShape* d, const void* x, const void* h_prev, const void* dz, float* dW, float* db,
                          cudaStream_t st, int kind);

int launch_wgrad_tc(const PlcCellDesc* d, const void* x, const void* h_prev, const void* dz, float* dW, float* db,
                    cudaStream_t st) {
  WgradShape w;
  w.B = d->B; w.H = d->H; w.W = d->W; w.k = d->k; w.Cin = d->Cin; w.Ch = d->Ch; w.N = 4 * d->Ch;
  return launch_wgrad_tc_shape(&w, x, h_prev, dz, dW, db, st, PLC_K_BWD_WGRAD);
}

int launch_wgrad_tc_shape(const WgradShape* d, const void* x, const void* h_prev, const void* dz, float* dW, float* db,
                          cudaStream_t st, int kind) {
  int rc;
  TcGeom g;
  pick_spatial_tile(d->H, d->W, &g, 6);   // 64-pixel blocks
  plc::WgradTcParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W;
  p.ksize = d->k; p.pad = d->k / 2;
  p.tw = g.tw; p.th = g.th; p.tiles_x = g.tiles_x; p.tiles_y = g.tiles_y;
  p.PB = d->B * g.tiles_x * g.tiles_y;
  p.chunks0 = cdiv(d->Cin, 64); p.chunks1 = cdiv(d->Ch, 64);
  p.CB = d->kt * d->k * d->k * (p.chunks0 + p.chunks1);
  const bool nd5 = d->kt > 1 || d->stride_t > 1;
  const bool general = nd5 || d->stride > 1;                 // strided / 3-D: single-CTA kernel only
  p.stride = d->stride; p.nd5 = nd5; p.kt = d->kt; p.pad_t = d->kt / 2; p.stride_t = d->stride_t; p.T_out = d->T_out;
  // CTA pairs (cta_group::2, 256-row output tiles) once dZ has >= 256 channels and there is enough work
  const bool pair = !general && d->N >= 256 && pick_cta_group(p.PB / 2) == 2;
  p.num_groups = cdiv(p.CB, plc::kWgMaxGB);
  p.GB = cdiv(p.CB, p.num_groups);
  if (pair) p.GB = 2 * cdiv(p.GB, 2);                      // even: the pair splits the columns in halves
  {
    static const int gb_env = [] { const char* e = getenv("PLC_WGRAD_GB"); return e ? atoi(e) : 0; }();   // experiments
    if (gb_env >= 2 && gb_env <= plc::kWgMaxGB && (!pair || gb_env % 2 == 0)) p.GB = gb_env;
  }
  p.num_groups = cdiv(p.CB, p.GB);
  p.n_tiles = cdiv(d->N, pair ? 256 : 128);
  const int tiles = p.n_tiles * p.num_groups;
  int S = (sm_count() / (pair ? 2 : 1)) / tiles;
  if (S < 1) S = 1;
  if (S > p.PB) S = p.PB;
  p.S = S;
  p.C0 = d->Cin; p.C1 = d->Ch; p.N4 = d->N; p.Ctot = d->Cin + d->Ch;
  p.dW = dW;
  p.db = db;
  p.prof = g_prof_buf;
  CUtensorMap tz, t0, t1;
  if ((rc = make_tmap_act(&tz, dz, d->B, d->H, d->W, d->N, g.tw, g.th))) return rc;
  if (d->Ch > 0) {
    if ((rc = make_tmap_act(&t1, h_prev, d->B, d->H, d->W, d->Ch, g.tw, g.th))) return rc;
  }
  if (d->Cin > 0 && general) {
    if ((rc = make_tmap_act(&t0, x, d->Bs, d->Hs, d->Ws, d->Cin, g.tw, g.th, 2, 64, CU_TENSOR_MAP_SWIZZLE_128B, d->stride,
                            nd5 ? d->Ts : 0)))
      return rc;
  } else if (d->Cin > 0) {
    if ((rc = make_tmap_act(&t0, x, d->B, d->H, d->W, d->Cin, g.tw, g.th))) return rc;
  } else {
    t0 = t1;
  }
  if (d->Ch == 0) t1 = t0;
  LaunchTimer timer(kind, st, conv_flops(d->B, d->H, d->W, d->Cin + d->Ch, d->N, d->k) * d->kt);
  if (pair) {
    PLC_CUDA(set_smem_once<TagW2>(plc::wgrad_tc_kernel2, plc::kW2SmemBytes));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * tiles * S);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = plc::kW2SmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    PLC_CUDA(cudaLaunchKernelEx(&cfg, plc::wgrad_tc_kernel2, p, tz, t0, t1));
  } else {
    PLC_CUDA(set_smem_once<TagW1>(plc::wgrad_tc_kernel, plc::kWgSmemBytes));
    plc::wgrad_tc_kernel<<<tiles * S, 256, plc::kWgSmemBytes, st>>>(p, tz, t0, t1);
    PLC_CUDA(cudaGetLastError());
  }
  return PLC_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int plc_abi_version(void) { return PLC_ABI_VERSION; }

const char* plc_last_error(void) { return g_err.c_str(); }

int plc_device_supported(int dev) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

size_t plc_packed_weight_bytes(const PlcCellDesc* d, int pack_kind) {
  if (check_desc(d) != PLC_OK) return 0;
  const size_t kk = static_cast<size_t>(d->k) * d->k;
  if (d->mode == PLC_MODE_FP32) return kk * (d->Cin + d->Ch) * 4 * d->Ch * sizeof(float);
  if (pack_kind == PLC_PACK_FWD) {
    const size_t ktot = static_cast<size_t>(kgeom(d->Cin, d->Ch, d->k).num_kb) * 64;
    return static_cast<size_t>(4) * d->Ch * ktot * 2;
  }
  if (pack_kind == PLC_PACK_DGRAD) {
    const size_t ktot = static_cast<size_t>(kgeom(4 * d->Ch, 0, d->k).num_kb) * 64;
    return static_cast<size_t>(d->Cin + d->Ch) * ktot * 2;
  }
  fail(PLC_ERR_BAD_DESC, "bad pack kind %d", pack_kind);
  return 0;
}

int plc_pack_weight(const PlcCellDesc* d, int pack_kind, const float* w_oihw, void* w_packed, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!w_oihw || !w_packed) return fail(PLC_ERR_NULL_ARG, "plc_pack_weight: null pointer");
  if (pack_kind != PLC_PACK_FWD && pack_kind != PLC_PACK_DGRAD) return fail(PLC_ERR_BAD_DESC, "bad pack kind");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LaunchTimer timer(PLC_K_PACK, st);
  const int threads = 256, blocks = 148 * 4;
  if (d->mode == PLC_MODE_FP32) {
    if (pack_kind == PLC_PACK_FWD)
      plc::pack_w_f32_fwd_kernel<<<blocks, threads, 0, st>>>(w_oihw, static_cast<float*>(w_packed), d->Cin, d->Ch, d->k);
    else
      plc::pack_w_f32_dgrad_kernel<<<blocks, threads, 0, st>>>(w_oihw, static_cast<float*>(w_packed), d->Cin, d->Ch,
                                                               d->k);
  } else {
    if (!aligned16(w_packed)) return fail(PLC_ERR_ALIGNMENT, "w_packed must be 16-byte aligned");
    if (pack_kind == PLC_PACK_FWD)
      pack_w_tc_fwd_kernel<<<blocks, threads, 0, st>>>(w_oihw, static_cast<__nv_bfloat16*>(w_packed), d->Cin, d->Ch,
                                                       d->k, pick_ch_tile(d->Ch), kgeom(d->Cin, d->Ch, d->k));
    else
      pack_w_conv_dgrad_kernel<<<blocks, threads, 0, st>>>(w_oihw, static_cast<__nv_bfloat16*>(w_packed),
                                                           d->Cin + d->Ch, 4 * d->Ch, d->k * d->k,
                                                           kgeom(4 * d->Ch, 0, d->k));
  }
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

int plc_cell_fwd_zero_state_ok(const PlcCellDesc* d) {
  if (check_desc(d) != PLC_OK) return 0;
  if (d->mode != PLC_MODE_BF16_TC || d->Cin <= 0) return 0;
  const KGeom kg = kgeom(d->Cin, d->Ch, d->k);
  return (d->k * d->k * kg.chunks0) % (64 / kg.kc) == 0;   // the x taps fill whole K stages
}

// saved-gates mode: bf16 tensor-core mode with full 64-channel slices (N_TILE = 256, the TMA epilogues)
static bool saved_gates_supported(const PlcCellDesc* d) {
  return d->mode == PLC_MODE_BF16_TC && pick_ch_tile(d->Ch) == 64 && plc::tma_store_epilogue<256, plc::EPI_LSTM_FWD>() &&
         plc::tma_store_epilogue<256, plc::EPI_LSTM_BWD_GATES>();
}
size_t plc_saved_gates_bytes(const PlcCellDesc* d) {
  if (check_desc(d) != PLC_OK || !saved_gates_supported(d)) return 0;
  TcGeom g;
  plc::ConvTcParams p;
  lstm_tc_setup(d, &g, &p, plc::EPI_LSTM_FWD);       // the forward kernel's tile geometry defines the layout
  return static_cast<size_t>(p.num_m_tiles) * p.num_n_tiles * (4 * 8 * 128 * 16);
}

static int cell_fwd_impl(const PlcCellDesc* d, const void* x, const void* h_prev, const void* c_prev,
                         const void* w_packed_fwd, const float* bias, void* h_out, void* c_out, void* gates_out,
                         void* gates_saved, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (gates_saved && !saved_gates_supported(d))
    return fail(PLC_ERR_UNSUPPORTED, "plc_cell_fwd_save: saved-gates mode needs bf16 mode and Ch %% 64 == 0 "
                                     "(plc_saved_gates_bytes returns 0 for this descriptor)");
  // zero-initial-state form (generator.py:156-160 starts every sequence from h = c = 0): h_prev == c_prev == NULL
  const bool zero_state = !h_prev && !c_prev;
  if (zero_state && !plc_cell_fwd_zero_state_ok(d))
    return fail(PLC_ERR_UNSUPPORTED, "plc_cell_fwd: the zero-state form (h_prev = c_prev = NULL) is not available for "
                                     "this shape/mode (see plc_cell_fwd_zero_state_ok); pass zero tensors");
  if ((d->Cin > 0 && !x) || (!zero_state && (!h_prev || !c_prev)) || !w_packed_fwd || !h_out || !c_out)
    return fail(PLC_ERR_NULL_ARG, "plc_cell_fwd: null pointer");
  if (d->has_bias && !bias) return fail(PLC_ERR_NULL_ARG, "plc_cell_fwd: has_bias set but bias is null");
  if (h_out == h_prev) return fail(PLC_ERR_BAD_DESC, "h_out must not alias h_prev (halo reads)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  if (d->mode == PLC_MODE_FP32) {
    plc::ConvSimtParams p;
    memset(&p, 0, sizeof(p));
    p.B = d->B; p.H = d->H; p.W = d->W; p.M = d->B * d->H * d->W;
    p.ksize = d->k; p.pad = d->k / 2;
    p.C0 = d->Cin; p.C1 = d->Ch;
    p.K = d->k * d->k * (d->Cin + d->Ch);
    p.N = 4 * d->Ch; p.Ch = d->Ch; p.Cin = d->Cin;
    p.src0 = static_cast<const float*>(x); p.src1 = static_cast<const float*>(h_prev);
    p.w = static_cast<const float*>(w_packed_fwd);
    p.bias = d->has_bias ? bias : nullptr;
    p.c_prev = static_cast<const float*>(c_prev);
    p.c_out = static_cast<float*>(c_out); p.h_out = static_cast<float*>(h_out);
    p.gates_out = static_cast<float*>(gates_out);
    dim3 grid(cdiv(p.M, plc::SBM), cdiv(p.N, plc::SBN));
    plc::conv_simt_kernel<plc::SEPI_LSTM_FWD><<<grid, 256, 0, st>>>(p);
    PLC_CUDA(cudaGetLastError());
    return PLC_OK;
  }

  if (!aligned16(x) || !aligned16(h_prev) || !aligned16(c_prev) || !aligned16(w_packed_fwd) || !aligned16(h_out) ||
      !aligned16(c_out) || !aligned16(gates_out))
    return fail(PLC_ERR_ALIGNMENT, "plc_cell_fwd: all device pointers must be 16-byte aligned");
  TcGeom g;
  plc::ConvTcParams p;
  lstm_tc_setup(d, &g, &p, plc::EPI_LSTM_FWD);
  p.bias = d->has_bias ? bias : nullptr;
  p.c_prev = static_cast<const float*>(c_prev);
  p.c_out = static_cast<float*>(c_out);
  p.h_out = static_cast<__nv_bfloat16*>(h_out);
  p.gates_out = static_cast<__nv_bfloat16*>(gates_out);
  p.gates_saved = static_cast<uint4*>(gates_saved);
  const long full_kb = p.num_kb;               // K extent of the packed weight image (both sources)
  if (zero_state) {
    // h_prev == 0: its taps contribute nothing.  The packed K order is source-major, so the x part is the leading
    // num_boxes0 / G stages (plc_cell_fwd_zero_state_ok guarantees it ends on a stage boundary): skip the rest.
    p.chunks1 = 0;
    p.num_boxes = d->k * d->k * p.chunks0;
    p.num_kb = p.num_boxes / (64 / p.kc);
  }
  CUtensorMap ta0, ta1, tb;
  if (d->Cin > 0 && (rc = make_tmap_src(&ta0, x, d->B, d->H, d->W, d->Cin, p))) return rc;
  if (zero_state) {
    ta1 = ta0;
  } else {
    if ((rc = make_tmap_src(&ta1, h_prev, d->B, d->H, d->W, d->Ch, p))) return rc;
    if (d->Cin == 0) ta0 = ta1;
  }
  const int cta = pick_cta_group(p.num_m_tiles);
  if ((rc = make_tmap_mat(&tb, w_packed_fwd, 4L * d->Ch, full_kb * 64, 64, g.n_tile / cta))) return rc;
  CUtensorMap to0 = ta1, to1 = ta1;
  if (plc::tma_store_epilogue<256, plc::EPI_LSTM_FWD>() && g.n_tile == 256) {
    if ((rc = make_tmap_act(&to0, c_out, d->B, d->H, d->W, d->Ch, g.tw, g.th, 4, 32))) return rc;
    if ((rc = make_tmap_act(&to1, h_out, d->B, d->H, d->W, d->Ch, g.tw, g.th, 2, 64))) return rc;
  }
  return launch_conv_tc<plc::EPI_LSTM_FWD>(g.n_tile, cta, p, ta0, ta1, tb, to0, to1, st,
                                           zero_state ? PLC_K_CELL_FWD_ZERO : PLC_K_CELL_FWD,
                                           conv_flops(d->B, d->H, d->W, (zero_state ? 0 : d->Ch) + d->Cin, 4 * d->Ch, d->k));
}

int plc_cell_fwd(const PlcCellDesc* d, const void* x, const void* h_prev, const void* c_prev,
                 const void* w_packed_fwd, const float* bias, void* h_out, void* c_out, void* gates_out,
                 void* stream) {
  return cell_fwd_impl(d, x, h_prev, c_prev, w_packed_fwd, bias, h_out, c_out, gates_out, nullptr, stream);
}
int plc_cell_fwd_save(const PlcCellDesc* d, const void* x, const void* h_prev, const void* c_prev,
                      const void* w_packed_fwd, const float* bias, void* h_out, void* c_out, void* gates_saved,
                      void* stream) {
  if (!gates_saved) return fail(PLC_ERR_NULL_ARG, "plc_cell_fwd_save: gates_saved is null");
  if (!aligned16(gates_saved)) return fail(PLC_ERR_ALIGNMENT, "plc_cell_fwd_save: gates_saved must be 16-byte aligned");
  return cell_fwd_impl(d, x, h_prev, c_prev, w_packed_fwd, bias, h_out, c_out, nullptr, gates_saved, stream);
}

size_t plc_bwd_workspace_bytes(const PlcCellDesc* d) {
  if (check_desc(d) != PLC_OK) return 0;
  const size_t m = static_cast<size_t>(d->B) * d->H * d->W;
  return m * 4 * d->Ch * (d->mode == PLC_MODE_FP32 ? 4 : 2);
}

// fp32 validation mode: weight + bias gradient of one call (dW_acc is the OIHW layout itself)
static int wgrad_fp32(const PlcCellDesc* d, const void* x, const void* h_prev, const void* dz, float* dW_acc, float* db_acc,
                      cudaStream_t st) {
  const int M = d->B * d->H * d->W, kk = d->k * d->k;
  plc::WgradParams w;
  memset(&w, 0, sizeof(w));
  w.B = d->B; w.H = d->H; w.W = d->W; w.M = M;
  w.ksize = d->k; w.pad = d->k / 2;
  w.C0 = d->Cin; w.C1 = d->Ch; w.K = kk * (d->Cin + d->Ch); w.N = 4 * d->Ch;
  w.src0 = x; w.src1 = h_prev; w.dz = dz; w.dW = dW_acc; w.db = d->has_bias ? db_acc : nullptr;
  const int tiles = cdiv(w.N, plc::SBN) * cdiv(w.K, plc::SBM);
  int splits = cdiv(sm_count() * 4, tiles);
  if (splits < 1) splits = 1;
  int ppb = cdiv(cdiv(M, splits), plc::SBK) * plc::SBK;
  if (ppb < plc::SBK) ppb = plc::SBK;
  w.pix_per_block = ppb;
  dim3 g3(cdiv(w.N, plc::SBN), cdiv(w.K, plc::SBM), cdiv(M, ppb));
  plc::wgrad_simt_kernel<float><<<g3, 256, 0, st>>>(w);
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

static int cell_bwd_impl(const PlcCellDesc* d, const void* x, const void* h_prev, const void* c_prev,
                         const void* w_packed_fwd, const void* w_packed_dgrad, const float* bias, const void* gates_saved,
                         const void* dh, const void* dh2, const float* dc_next, void* dx, void* dh_prev, float* dc_prev,
                         float* dW_acc, float* db_acc, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if ((d->Cin > 0 && !x) || !h_prev || !c_prev || (!w_packed_fwd && !gates_saved) || !dh || !dc_prev || !workspace)
    return fail(PLC_ERR_NULL_ARG, "plc_cell_bwd: null pointer");
  if ((dx || dh_prev) && !w_packed_dgrad) return fail(PLC_ERR_NULL_ARG, "plc_cell_bwd: dgrad image missing");
  if (d->has_bias && !bias && !gates_saved) return fail(PLC_ERR_NULL_ARG, "plc_cell_bwd: has_bias set but bias is null");
  if (gates_saved && !saved_gates_supported(d))
    return fail(PLC_ERR_UNSUPPORTED, "plc_cell_bwd_saved: saved-gates mode needs bf16 mode and Ch %% 64 == 0");
  if (workspace_bytes < plc_bwd_workspace_bytes(d))
    return fail(PLC_ERR_WORKSPACE, "workspace %zu < required %zu bytes", workspace_bytes, plc_bwd_workspace_bytes(d));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = d->B * d->H * d->W;
  const int kk = d->k * d->k;

  if (d->mode == PLC_MODE_FP32) {
    // 1) recompute gates, dZ + dc_prev
    plc::ConvSimtParams p;
    memset(&p, 0, sizeof(p));
    p.B = d->B; p.H = d->H; p.W = d->W; p.M = M;
    p.ksize = d->k; p.pad = d->k / 2;
    p.C0 = d->Cin; p.C1 = d->Ch; p.K = kk * (d->Cin + d->Ch);
    p.N = 4 * d->Ch; p.Ch = d->Ch; p.Cin = d->Cin;
    p.src0 = static_cast<const float*>(x); p.src1 = static_cast<const float*>(h_prev);
    p.w = static_cast<const float*>(w_packed_fwd);
    p.bias = d->has_bias ? bias : nullptr;
    p.c_prev = static_cast<const float*>(c_prev);
    p.dh = static_cast<const float*>(dh); p.dh2 = static_cast<const float*>(dh2);
    p.dc_next = dc_next; p.dc_prev = dc_prev;
    p.dz = static_cast<float*>(workspace);
    dim3 grid(cdiv(M, plc::SBM), cdiv(p.N, plc::SBN));
    plc::conv_simt_kernel<plc::SEPI_LSTM_BWD_GATES><<<grid, 256, 0, st>>>(p);
    PLC_CUDA(cudaGetLastError());
    // 2) dgrad
    if (dx || dh_prev) {
      plc::ConvSimtParams q;
      memset(&q, 0, sizeof(q));
      q.B = d->B; q.H = d->H; q.W = d->W; q.M = M;
      q.ksize = d->k; q.pad = d->k / 2;
      q.C0 = 4 * d->Ch; q.C1 = 0; q.K = kk * 4 * d->Ch;
      q.N = d->Cin + d->Ch; q.Ch = d->Ch; q.Cin = d->Cin;
      q.src0 = static_cast<const float*>(workspace);
      q.w = static_cast<const float*>(w_packed_dgrad);
      q.out0 = static_cast<float*>(dx); q.out1 = static_cast<float*>(dh_prev);
      dim3 g2(cdiv(M, plc::SBM), cdiv(q.N, plc::SBN));
      plc::conv_simt_kernel<plc::SEPI_PLAIN><<<g2, 256, 0, st>>>(q);
      PLC_CUDA(cudaGetLastError());
    }
    // 3) wgrad (+ bias grad)
    if (dW_acc) return wgrad_fp32(d, x, h_prev, workspace, dW_acc, db_acc, st);
    return PLC_OK;
  }

  // ---------------- bf16 tensor-core mode
  if (!aligned16(x) || !aligned16(h_prev) || !aligned16(c_prev) || !aligned16(w_packed_fwd) ||
      !aligned16(w_packed_dgrad) || !aligned16(dh) || !aligned16(dh2) || !aligned16(dc_next) || !aligned16(dx) ||
      !aligned16(dh_prev) || !aligned16(dc_prev) || !aligned16(workspace))
    return fail(PLC_ERR_ALIGNMENT, "plc_cell_bwd: all device pointers must be 16-byte aligned");
  TcGeom g;
  plc::ConvTcParams p;
  // saved-gates form: the tile geometry must be the FORWARD kernel's (it defines the saved layout); no mainloop runs
  lstm_tc_setup(d, &g, &p, gates_saved ? plc::EPI_LSTM_FWD : plc::EPI_LSTM_BWD_GATES);
  if (gates_saved) {
    if (!aligned16(gates_saved)) return fail(PLC_ERR_ALIGNMENT, "plc_cell_bwd_saved: gates_saved must be 16-byte aligned");
    p.patch = 0;
    p.gates_saved = static_cast<uint4*>(const_cast<void*>(gates_saved));
  }
  // 1) gate recompute (same mainloop as the forward) + dZ / dc_prev epilogue
  {
    static const int skip = [] { const char* e = getenv("PLC_EXP_SKIP_GATE_MMA"); return e ? atoi(e) : 0; }();
    p.exp_skip_mma = skip;
  }
  p.bias = d->has_bias ? bias : nullptr;
  p.c_prev = static_cast<const float*>(c_prev);
  p.dh = static_cast<const __nv_bfloat16*>(dh);
  p.dh2 = static_cast<const __nv_bfloat16*>(dh2);
  p.dc_next = dc_next;
  p.dc_prev = dc_prev;
  p.dz = static_cast<__nv_bfloat16*>(workspace);
  CUtensorMap ta0, ta1, tb;
  if ((rc = make_tmap_src(&ta1, h_prev, d->B, d->H, d->W, d->Ch, p))) return rc;
  if (d->Cin > 0) {
    if ((rc = make_tmap_src(&ta0, x, d->B, d->H, d->W, d->Cin, p))) return rc;
  } else {
    ta0 = ta1;
  }
  const int cta = pick_cta_group(p.num_m_tiles);
  if (gates_saved) tb = ta1;      // never dereferenced (no mainloop); must still be a valid descriptor for the prefetch
  else if ((rc = make_tmap_mat(&tb, w_packed_fwd, 4L * d->Ch, (long)p.num_kb * 64, 64, g.n_tile / cta))) return rc;
  CUtensorMap tzo = ta1, tdc = ta1;
  plc::GateMaps gm{};
  if (plc::tma_store_epilogue<256, plc::EPI_LSTM_BWD_GATES>() && g.n_tile == 256) {
    // 16-channel round boxes (ConvTcCfg::kGateRoundCh): bf16 = 32-byte rows -> SWIZZLE_32B (dZ gate slices, dh, dh2);
    // fp32 = 64-byte rows -> SWIZZLE_64B (dc_prev, c_prev, dc_next)
    const int rc_ch = 16;
    const CUtensorMapSwizzle s32 = CU_TENSOR_MAP_SWIZZLE_32B, s64 = CU_TENSOR_MAP_SWIZZLE_64B;
    if ((rc = make_tmap_act(&tzo, workspace, d->B, d->H, d->W, 4 * d->Ch, g.tw, g.th, 2, rc_ch, s32))) return rc;
    if ((rc = make_tmap_act(&tdc, dc_prev, d->B, d->H, d->W, d->Ch, g.tw, g.th, 4, rc_ch, s64))) return rc;
    if ((rc = make_tmap_act(&gm.c_prev, c_prev, d->B, d->H, d->W, d->Ch, g.tw, g.th, 4, rc_ch, s64))) return rc;
    gm.dc_next = gm.c_prev;
    if (dc_next && (rc = make_tmap_act(&gm.dc_next, dc_next, d->B, d->H, d->W, d->Ch, g.tw, g.th, 4, rc_ch, s64))) return rc;
    if ((rc = make_tmap_act(&gm.dh, dh, d->B, d->H, d->W, d->Ch, g.tw, g.th, 2, rc_ch, s32))) return rc;
    gm.dh2 = gm.dh;
    if (dh2 && (rc = make_tmap_act(&gm.dh2, dh2, d->B, d->H, d->W, d->Ch, g.tw, g.th, 2, rc_ch, s32))) return rc;
  }
  if ((rc = launch_conv_tc<plc::EPI_LSTM_BWD_GATES>(g.n_tile, cta, p, ta0, ta1, tb, tzo, tdc, st, PLC_K_BWD_GATES,
                                                    gates_saved ? 0.0 : conv_flops(d->B, d->H, d->W, d->Cin + d->Ch,
                                                                                   4 * d->Ch, d->k), &gm)))
    return rc;

  // 2) dgrad: conv of dZ (4Ch channels) with the flipped/transposed image -> dx, dh_prev
  if (dx || dh_prev) {
    plc::ConvTcParams q;
    TcGeom gq;
    pick_spatial_tile(d->H, d->W, &gq);
    fill_geom(d, gq, &q);
    const int n_total = d->Cin + d->Ch;
    const int nt = pick_plain_n_tile(n_total);
    q.num_n_tiles = cdiv(n_total, nt);
    q.num_tiles = q.num_m_tiles * q.num_n_tiles;
    set_kgeom(&q, kgeom(4 * d->Ch, 0, d->k));
    maybe_patch(&gq, &q, plc::EPI_PLAIN, nt);
    q.n_total = n_total;
    q.out0 = static_cast<__nv_bfloat16*>(dx);
    q.out1 = static_cast<__nv_bfloat16*>(dh_prev);
    const int ctaq = pick_cta_group(q.num_m_tiles);
    CUtensorMap tz, tbd;
    if ((rc = make_tmap_src(&tz, workspace, d->B, d->H, d->W, 4 * d->Ch, q))) return rc;
    if ((rc = make_tmap_mat(&tbd, w_packed_dgrad, n_total, (long)q.num_kb * 64, 64, nt / ctaq))) return rc;
    CUtensorMap tq0 = tz, tq1 = tz;
    if ((rc = setup_plain_stores(&q, d->B, d->H, d->W, &tq0, &tq1))) return rc;
    if ((rc = launch_conv_tc<plc::EPI_PLAIN>(nt, ctaq, q, tz, tz, tbd, tq0, tq1, st, PLC_K_BWD_DGRAD,
                                             conv_flops(d->B, d->H, d->W, 4 * d->Ch, (dx ? d->Cin : 0) + (dh_prev ? d->Ch : 0),
                                                        d->k))))
      return rc;
  }

  // 3) wgrad + bias grad
  if (dW_acc) {
    if ((rc = launch_wgrad_tc(d, x, h_prev, workspace, dW_acc, d->has_bias ? db_acc : nullptr, st))) return rc;
  }
  return PLC_OK;
}

int plc_cell_bwd(const PlcCellDesc* d, const void* x, const void* h_prev, const void* c_prev,
                 const void* w_packed_fwd, const void* w_packed_dgrad, const float* bias, const void* dh,
                 const void* dh2, const float* dc_next, void* dx, void* dh_prev, float* dc_prev, float* dW_acc,
                 float* db_acc, void* workspace, size_t workspace_bytes, void* stream) {
  if (!w_packed_fwd) return fail(PLC_ERR_NULL_ARG, "plc_cell_bwd: null pointer");
  return cell_bwd_impl(d, x, h_prev, c_prev, w_packed_fwd, w_packed_dgrad, bias, nullptr, dh, dh2, dc_next, dx, dh_prev,
                       dc_prev, dW_acc, db_acc, workspace, workspace_bytes, stream);
}
int plc_cell_bwd_saved(const PlcCellDesc* d, const void* x, const void* h_prev, const void* c_prev,
                       const void* gates_saved, const void* w_packed_dgrad, const void* dh, const void* dh2,
                       const float* dc_next, void* dx, void* dh_prev, float* dc_prev, float* dW_acc, float* db_acc,
                       void* workspace, size_t workspace_bytes, void* stream) {
  if (!gates_saved) return fail(PLC_ERR_NULL_ARG, "plc_cell_bwd_saved: gates_saved is null");
  return cell_bwd_impl(d, x, h_prev, c_prev, nullptr, w_packed_dgrad, nullptr, gates_saved, dh, dh2, dc_next, dx, dh_prev,
                       dc_prev, dW_acc, db_acc, workspace, workspace_bytes, stream);
}

int plc_cell_wgrad(const PlcCellDesc* d, const void* x, const void* h_prev, const void* dz, float* dW_acc, float* db_acc,
                   void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if ((d->Cin > 0 && !x) || !h_prev || !dz || !dW_acc) return fail(PLC_ERR_NULL_ARG, "plc_cell_wgrad: null pointer");
  if (d->has_bias && !db_acc) return fail(PLC_ERR_NULL_ARG, "plc_cell_wgrad: has_bias set but db_acc is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->mode == PLC_MODE_FP32) return wgrad_fp32(d, x, h_prev, dz, dW_acc, db_acc, st);
  if (!aligned16(x) || !aligned16(h_prev) || !aligned16(dz))
    return fail(PLC_ERR_ALIGNMENT, "plc_cell_wgrad: all device pointers must be 16-byte aligned");
  return launch_wgrad_tc(d, x, h_prev, dz, dW_acc, d->has_bias ? db_acc : nullptr, st);
}

// ---------------------------------------------------------------------------------- weight-gradient accumulator
static size_t wgrad_acc_elems(int mode, int Cin, int Ch, int N, int taps) {
  const size_t kk = static_cast<size_t>(taps);
  if (mode == PLC_MODE_FP32) return static_cast<size_t>(N) * (Cin + Ch) * kk;              // OIHW itself
  return static_cast<size_t>(N) * kk * (cdiv(Cin, 64) + cdiv(Ch, 64)) * 64;                 // packed [N][CB*64]
}
static int wgrad_unpack(int mode, int Cin, int Ch, int N, int taps, const float* acc, float* dW, cudaStream_t st) {
  if (!acc || !dW) return fail(PLC_ERR_NULL_ARG, "wgrad unpack: null pointer");
  LaunchTimer timer(PLC_K_ELEMENTWISE, st);
  if (mode == PLC_MODE_FP32)
    plc::add_inplace_kernel<<<148 * 4, 256, 0, st>>>(acc, dW, wgrad_acc_elems(mode, Cin, Ch, N, taps));
  else
    plc::wgrad_unpack_kernel<<<148 * 4, 256, 0, st>>>(acc, dW, N, Cin, Ch, taps);
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

size_t plc_wgrad_acc_bytes(const PlcCellDesc* d) {
  if (check_desc(d) != PLC_OK) return 0;
  return wgrad_acc_elems(d->mode, d->Cin, d->Ch, 4 * d->Ch, d->k * d->k) * sizeof(float);
}
int plc_wgrad_unpack(const PlcCellDesc* d, const float* acc, float* dW_oihw, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  return wgrad_unpack(d->mode, d->Cin, d->Ch, 4 * d->Ch, d->k * d->k, acc, dW_oihw, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------- generic "same" conv (bf16)
static int check_conv(const PlcConvDesc* d) {
  if (!d) return fail(PLC_ERR_BAD_DESC, "null conv descriptor");
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0)
    return fail(PLC_ERR_BAD_DESC, "bad conv sizes B=%d H=%d W=%d Cin=%d Cout=%d", d->B, d->H, d->W, d->Cin, d->Cout);
  if (d->k <= 0 || (d->k % 2) == 0 || d->k > 7) return fail(PLC_ERR_BAD_DESC, "conv kernel size %d must be odd, <= 7", d->k);
  if (d->Cin % 8 || d->Cout % 8)
    return fail(PLC_ERR_ALIGNMENT, "conv needs Cin %% 8 == 0 and Cout %% 8 == 0 (got %d, %d): pad channels", d->Cin, d->Cout);
  if (d->pixel_shuffle && d->Cout % 32) return fail(PLC_ERR_ALIGNMENT, "PixelShuffle(2) store needs Cout %% 32 == 0");
  if ((long long)d->B * d->H * d->W >= (1ll << 31)) return fail(PLC_ERR_UNSUPPORTED, "B*H*W must be < 2^31");
  return PLC_OK;
}

size_t plc_conv_packed_weight_bytes(const PlcConvDesc* d, int pack_kind) {
  if (check_conv(d) != PLC_OK) return 0;
  const size_t kk = static_cast<size_t>(d->k) * d->k;
  (void)kk;
  if (pack_kind == PLC_PACK_FWD) return static_cast<size_t>(d->Cout) * kgeom(d->Cin, 0, d->k).num_kb * 64 * 2;
  if (pack_kind == PLC_PACK_DGRAD) return static_cast<size_t>(d->Cin) * kgeom(d->Cout, 0, d->k).num_kb * 64 * 2;
  fail(PLC_ERR_BAD_DESC, "bad pack kind %d", pack_kind);
  return 0;
}

int plc_conv_pack_weight(const PlcConvDesc* d, int pack_kind, const float* w_oihw, const float* bias, void* w_packed,
                         float* bias_packed, void* stream) {
  int rc = check_conv(d);
  if (rc) return rc;
  if (!w_oihw || !w_packed) return fail(PLC_ERR_NULL_ARG, "plc_conv_pack_weight: null pointer");
  if (!aligned16(w_packed) || !aligned16(bias_packed)) return fail(PLC_ERR_ALIGNMENT, "packed buffers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LaunchTimer timer(PLC_K_PACK, st);
  if (pack_kind == PLC_PACK_FWD)
    pack_w_conv_fwd_kernel<<<148 * 4, 256, 0, st>>>(w_oihw, bias, static_cast<__nv_bfloat16*>(w_packed), bias_packed,
                                                    d->Cin, d->Cout, d->k * d->k, kgeom(d->Cin, 0, d->k),
                                                    d->pixel_shuffle);
  else if (pack_kind == PLC_PACK_DGRAD)
    pack_w_conv_dgrad_kernel<<<148 * 4, 256, 0, st>>>(w_oihw, static_cast<__nv_bfloat16*>(w_packed), d->Cin, d->Cout,
                                                      d->k * d->k, kgeom(d->Cout, 0, d->k));
  else
    return fail(PLC_ERR_BAD_DESC, "bad pack kind %d", pack_kind);
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

static int conv_fwd_impl(const PlcConvDesc* d, const void* x, const void* w_packed_fwd, const float* bias_packed, void* out,
                         bool out_f32, void* stream) {
  int rc = check_conv(d);
  if (rc) return rc;
  if (!x || !w_packed_fwd || !out) return fail(PLC_ERR_NULL_ARG, "plc_conv_fwd: null pointer");
  if (out_f32 && d->pixel_shuffle) return fail(PLC_ERR_UNSUPPORTED, "plc_conv_fwd_f32: no PixelShuffle store in fp32");
  if (!aligned16(x) || !aligned16(w_packed_fwd) || !aligned16(out) || !aligned16(bias_packed))
    return fail(PLC_ERR_ALIGNMENT, "plc_conv_fwd: all device pointers must be 16-byte aligned");
  PlcCellDesc cd{d->B, d->H, d->W, d->Cin, d->Cout, d->k, PLC_MODE_BF16_TC, 0};
  TcGeom g;
  pick_spatial_tile(d->H, d->W, &g);
  plc::ConvTcParams q;
  fill_geom(&cd, g, &q);
  const int nt = pick_plain_n_tile(d->Cout);
  q.num_n_tiles = cdiv(d->Cout, nt);
  q.num_tiles = q.num_m_tiles * q.num_n_tiles;
  set_kgeom(&q, kgeom(d->Cin, 0, d->k));
  maybe_patch(&g, &q, plc::EPI_PLAIN, nt);
  q.n_total = d->Cout;
  q.Cin = d->Cout;                     // no column split: everything goes to out0
  q.out0 = static_cast<__nv_bfloat16*>(out);
  q.plain_bias = d->has_bias ? bias_packed : nullptr;
  q.plain_relu = d->relu;
  q.plain_shuffle = d->pixel_shuffle;
  q.plain_f32 = out_f32 ? 1 : 0;
  const int cta = pick_cta_group(q.num_m_tiles);
  CUtensorMap ta, tb;
  if ((rc = make_tmap_src(&ta, x, d->B, d->H, d->W, d->Cin, q))) return rc;
  if ((rc = make_tmap_mat(&tb, w_packed_fwd, d->Cout, (long)q.num_kb * 64, 64, nt / cta))) return rc;
  CUtensorMap to0 = ta, to1 = ta;
  if (!out_f32 && (rc = setup_plain_stores(&q, d->B, d->H, d->W, &to0, &to1))) return rc;   // fp32: per-thread stores
  return launch_conv_tc<plc::EPI_PLAIN>(nt, cta, q, ta, ta, tb, to0, to1, static_cast<cudaStream_t>(stream),
                                        PLC_K_CONV_FWD, conv_flops(d->B, d->H, d->W, d->Cin, d->Cout, d->k));
}
int plc_conv_fwd(const PlcConvDesc* d, const void* x, const void* w_packed_fwd, const float* bias_packed, void* out,
                 void* stream) {
  return conv_fwd_impl(d, x, w_packed_fwd, bias_packed, out, false, stream);
}
int plc_conv_fwd_f32(const PlcConvDesc* d, const void* x, const void* w_packed_fwd, const float* bias_packed, float* out,
                     void* stream) {
  return conv_fwd_impl(d, x, w_packed_fwd, bias_packed, out, true, stream);
}

size_t plc_conv_wgrad_acc_bytes(const PlcConvDesc* d) {
  if (check_conv(d) != PLC_OK) return 0;
  return wgrad_acc_elems(PLC_MODE_BF16_TC, d->Cin, 0, d->Cout, d->k * d->k) * sizeof(float);
}
int plc_conv_wgrad_unpack(const PlcConvDesc* d, const float* acc, float* dW_oihw, void* stream) {
  int rc = check_conv(d);
  if (rc) return rc;
  return wgrad_unpack(PLC_MODE_BF16_TC, d->Cin, 0, d->Cout, d->k * d->k, acc, dW_oihw,
                      static_cast<cudaStream_t>(stream));
}

int plc_conv_grad_mask(const PlcConvDesc* d, const void* y, const void* dy, void* dz, void* stream) {
  int rc = check_conv(d);
  if (rc) return rc;
  if (!y || !dy || !dz) return fail(PLC_ERR_NULL_ARG, "plc_conv_grad_mask: null pointer");
  const size_t npix = static_cast<size_t>(d->B) * d->H * d->W;
  LaunchTimer timer(PLC_K_ELEMENTWISE, static_cast<cudaStream_t>(stream));
  if (!d->pixel_shuffle && d->relu && d->Cout % 8 == 0 && aligned16(y) && aligned16(dy) && aligned16(dz))
    conv_grad_mask_vec_kernel<<<sm_count() * 16, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(y), static_cast<const uint4*>(dy), static_cast<uint4*>(dz), npix * (d->Cout / 8));
  else
    conv_grad_mask_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dz), npix,
        d->H, d->W, d->Cout, d->relu, d->pixel_shuffle);
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

int plc_conv_im2col_narrow(const PlcConvDesc* d, int cin_true, const void* x, void* col, void* stream) {
  int rc = check_conv(d);
  if (rc) return rc;
  if (!x || !col) return fail(PLC_ERR_NULL_ARG, "plc_conv_im2col_narrow: null pointer");
  if (d->Cin != 8 || cin_true < 1 || cin_true > 8 || d->k * d->k * cin_true > 32)
    return fail(PLC_ERR_UNSUPPORTED, "plc_conv_im2col_narrow: needs an 8-channel (padded) input and k*k*cin_true <= 32 "
                                     "(got Cin %d, k %d, cin_true %d)", d->Cin, d->k, cin_true);
  if (!aligned16(x) || !aligned16(col))
    return fail(PLC_ERR_ALIGNMENT, "plc_conv_im2col_narrow: all device pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LaunchTimer timer(PLC_K_ELEMENTWISE, st);
  im2col_narrow_kernel<<<sm_count() * 16, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<uint4*>(col),
                                                        d->B, d->H, d->W, cin_true, d->k);
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

int plc_conv_bwd(const PlcConvDesc* d, const void* x, const void* dz, const void* w_packed_dgrad, void* dx,
                 float* dW_acc, float* db_acc, void* stream) {
  int rc = check_conv(d);
  if (rc) return rc;
  if (!x || !dz) return fail(PLC_ERR_NULL_ARG, "plc_conv_bwd: null pointer");
  if (dx && !w_packed_dgrad) return fail(PLC_ERR_NULL_ARG, "plc_conv_bwd: dgrad image missing");
  if (!aligned16(x) || !aligned16(dz) || !aligned16(dx) || !aligned16(w_packed_dgrad))
    return fail(PLC_ERR_ALIGNMENT, "plc_conv_bwd: all device pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dx) {   // dgrad: conv of dZ with the flipped / transposed image
    PlcCellDesc cd{d->B, d->H, d->W, d->Cout, d->Cin, d->k, PLC_MODE_BF16_TC, 0};
    TcGeom g;
    pick_spatial_tile(d->H, d->W, &g);
    plc::ConvTcParams q;
    fill_geom(&cd, g, &q);
    const int nt = pick_plain_n_tile(d->Cin);
    q.num_n_tiles = cdiv(d->Cin, nt);
    q.num_tiles = q.num_m_tiles * q.num_n_tiles;
    set_kgeom(&q, kgeom(d->Cout, 0, d->k));
    maybe_patch(&g, &q, plc::EPI_PLAIN, nt);
    q.n_total = d->Cin;
    q.Cin = d->Cin;
    q.out0 = static_cast<__nv_bfloat16*>(dx);
    const int cta = pick_cta_group(q.num_m_tiles);
    CUtensorMap tz, tb;
    if ((rc = make_tmap_src(&tz, dz, d->B, d->H, d->W, d->Cout, q))) return rc;
    if ((rc = make_tmap_mat(&tb, w_packed_dgrad, d->Cin, (long)q.num_kb * 64, 64, nt / cta))) return rc;
    CUtensorMap to0 = tz, to1 = tz;
    if ((rc = setup_plain_stores(&q, d->B, d->H, d->W, &to0, &to1))) return rc;
    if ((rc = launch_conv_tc<plc::EPI_PLAIN>(nt, cta, q, tz, tz, tb, to0, to1, st, PLC_K_CONV_DGRAD,
                                             conv_flops(d->B, d->H, d->W, d->Cout, d->Cin, d->k))))
      return rc;
  }
  if (dW_acc) {
    WgradShape w;
    w.B = d->B; w.H = d->H; w.W = d->W; w.k = d->k; w.Cin = d->Cin; w.Ch = 0; w.N = d->Cout;
    if ((rc = launch_wgrad_tc_shape(&w, x, nullptr, dz, dW_acc, d->has_bias ? db_acc : nullptr, st, PLC_K_CONV_WGRAD)))
      return rc;
  }
  return PLC_OK;
}


// ---------------------------------------------------------------------------------- strided 2-D / 3-D conv (bf16)
// The discriminator's convolutions (north_star: "strided 2D/3D convolutions reuse the same implicit-GEMM core"): the
// default pipeline of conv_igemm_tc_kernel with strided (elementStrides) and, for a time kernel, 5-D tensor maps.
struct NdGeom { int To, Ho, Wo, taps, nd5; };

static int check_convnd(const PlcConvNdDesc* d, NdGeom* g) {
  if (!d) return fail(PLC_ERR_BAD_DESC, "null conv descriptor");
  if (d->B <= 0 || d->T <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0)
    return fail(PLC_ERR_BAD_DESC, "bad conv sizes B=%d T=%d H=%d W=%d Cin=%d Cout=%d", d->B, d->T, d->H, d->W, d->Cin,
                d->Cout);
  if (d->k <= 0 || (d->k % 2) == 0 || d->k > 7 || d->kt <= 0 || (d->kt % 2) == 0 || d->kt > 7)
    return fail(PLC_ERR_BAD_DESC, "conv kernel sizes (kt=%d, k=%d) must be odd and <= 7 (zero padding k/2)", d->kt, d->k);
  if ((d->stride != 1 && d->stride != 2) || (d->stride_t != 1 && d->stride_t != 2))
    return fail(PLC_ERR_UNSUPPORTED, "conv strides (%d, %d) must be 1 or 2", d->stride_t, d->stride);
  if (d->Cin % 8 || d->Cout % 8)
    return fail(PLC_ERR_ALIGNMENT, "conv needs Cin %% 8 == 0 and Cout %% 8 == 0 (got %d, %d): pad channels", d->Cin, d->Cout);
  if (d->act < 0 || d->act > 2) return fail(PLC_ERR_BAD_DESC, "act must be 0 (none), 1 (ReLU) or 2 (LeakyReLU)");
  g->To = (d->T - 1) / d->stride_t + 1;
  g->Ho = (d->H - 1) / d->stride + 1;
  g->Wo = (d->W - 1) / d->stride + 1;
  g->taps = d->kt * d->k * d->k;
  g->nd5 = (d->kt > 1 || d->stride_t > 1) ? 1 : 0;
  if ((long long)d->B * d->T * d->H * d->W >= (1ll << 31)) return fail(PLC_ERR_UNSUPPORTED, "B*T*H*W must be < 2^31");
  return PLC_OK;
}

// one EPI_PLAIN launch: input grid [B, T, H, W, cin] -> output grid [B*To, Ho, Wo, cout]
static int convnd_run(const void* x, int B, int T, int H, int W, int cin, int To, int Ho, int Wo, int cout, int kt, int k,
                      int st_t, int st_s, const void* w_packed, const float* bias_packed, int relu, float slope, void* out,
                      int kind, cudaStream_t st) {
  int rc;
  const int nd5 = (kt > 1 || st_t > 1) ? 1 : 0;
  const int imgs = nd5 ? B * To : B * T;
  PlcCellDesc cd{imgs, Ho, Wo, cin, cout, k, PLC_MODE_BF16_TC, 0};
  TcGeom g;
  pick_spatial_tile(Ho, Wo, &g);
  plc::ConvTcParams q;
  fill_geom(&cd, g, &q);
  const int nt = pick_plain_n_tile(cout);
  q.num_n_tiles = cdiv(cout, nt);
  q.num_tiles = q.num_m_tiles * q.num_n_tiles;
  set_kgeom(&q, kgeom_taps(cin, 0, kt * k * k));
  q.stride = st_s; q.nd5 = nd5; q.kt = kt; q.pad_t = kt / 2; q.stride_t = st_t; q.T_out = To;
  q.tap_t0 = -(kt / 2);
  if (!nd5 && st_s == 1) maybe_patch(&g, &q, plc::EPI_PLAIN, nt);
  q.n_total = cout;
  q.Cin = cout;                        // no column split: everything goes to out0
  q.out0 = static_cast<__nv_bfloat16*>(out);
  q.plain_bias = bias_packed;
  q.plain_relu = relu;
  q.plain_slope = slope;
  const int cta = pick_cta_group(q.num_m_tiles);
  CUtensorMap ta, tb;
  if (q.patch) {
    if ((rc = make_tmap_src(&ta, x, imgs, H, W, cin, q))) return rc;
  } else if ((rc = make_tmap_act(&ta, x, nd5 ? B : imgs, H, W, cin, q.tw, q.th, 2, q.kc, swizzle_for_kc(q.kc), st_s,
                                 nd5 ? T : 0))) {
    return rc;
  }
  if ((rc = make_tmap_mat(&tb, w_packed, cout, (long)q.num_kb * 64, 64, nt / cta))) return rc;
  CUtensorMap to0 = ta, to1 = ta;
  if ((rc = setup_plain_stores(&q, imgs, Ho, Wo, &to0, &to1))) return rc;
  return launch_conv_tc<plc::EPI_PLAIN>(nt, cta, q, ta, ta, tb, to0, to1, st, kind,
                                        conv_flops(imgs, Ho, Wo, cin, cout, k) * kt);
}

int plc_convnd_out_shape(const PlcConvNdDesc* d, int* T_out, int* H_out, int* W_out) {
  NdGeom g;
  int rc = check_convnd(d, &g);
  if (rc) return rc;
  if (T_out) *T_out = g.To;
  if (H_out) *H_out = g.Ho;
  if (W_out) *W_out = g.Wo;
  return PLC_OK;
}

// Transposed conv of a STRIDED layer, decomposed by output phase.  Along one dimension (kernel k, stride s, padding p) the
// dX positions of parity class ph = i mod s receive exactly the taps k0, k0+s, ... with k0 = (ph + p) mod s, read from dZ
// at offsets o0, o0-1, ... (o0 = (ph + p - k0) / s) relative to i / s:   dX[s*i' + ph] = sum_j dZ[i' + o0 - j] * W[k0 + j*s].
// Each phase is therefore a STRIDE-1 conv of dZ with its own small tap set: no zero-inserted copy of dZ, and exactly the
// layer's MMAs in total (a zero-insertion formulation executes s^d times as many and streams an s^d times larger tensor).
struct PhaseDim { int k0, n, o0; };
static PhaseDim phase_dim(int k, int s, int pad, int ph) {
  PhaseDim r;
  r.k0 = (ph + pad) % s;
  r.n = r.k0 < k ? (k - r.k0 + s - 1) / s : 0;
  r.o0 = (ph + pad - r.k0) / s;
  return r;
}
// phase image Wd_ph[c][(tap j = (jz, jy, jx), dZ-channel chunk)*64 + jj] = w[n][c][k0z + jz*st][k0y + jy*s][k0x + jx*s]
__global__ void pack_w_phase_dgrad_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cin, int cout,
                                          int kt, int k, int st, int s, PhaseDim pz, PhaseDim py, PhaseDim px, KGeom kg) {
  const int taps = pz.n * py.n * px.n, ktot = kg.num_kb * 64;
  const size_t total = static_cast<size_t>(cin) * ktot;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int kp = idx % ktot, c = idx / ktot;
    int tap;
    const int n = decode_packed_k(kp, kg.kc, taps, kg.chunks0, 0, cout, 0, tap);
    float v = 0.f;
    if (n >= 0) {
      const int jx = tap % px.n, jy = (tap / px.n) % py.n, jz = tap / (px.n * py.n);
      const int kz = pz.k0 + jz * st, ky = py.k0 + jy * s, kx = px.k0 + jx * s;
      v = w[(((static_cast<size_t>(n) * cin + c) * kt + kz) * k + ky) * k + kx];
    }
    out[idx] = __float2bfloat16(v);
  }
}
static bool convnd_strided(const PlcConvNdDesc* d) { return d->stride > 1 || d->stride_t > 1; }
static size_t phase_image_bytes(const PlcConvNdDesc* d, int pt, int py, int px, int* taps_out) {
  const int taps = phase_dim(d->kt, d->stride_t, d->kt / 2, pt).n * phase_dim(d->k, d->stride, d->k / 2, py).n *
                   phase_dim(d->k, d->stride, d->k / 2, px).n;
  if (taps_out) *taps_out = taps;
  return taps ? static_cast<size_t>(d->Cin) * kgeom_taps(d->Cout, 0, taps).num_kb * 128 : 0;
}

size_t plc_convnd_packed_weight_bytes(const PlcConvNdDesc* d, int pack_kind) {
  NdGeom g;
  if (check_convnd(d, &g) != PLC_OK) return 0;
  if (pack_kind == PLC_PACK_FWD) return static_cast<size_t>(d->Cout) * kgeom_taps(d->Cin, 0, g.taps).num_kb * 64 * 2;
  if (pack_kind == PLC_PACK_DGRAD) {
    if (!convnd_strided(d)) return static_cast<size_t>(d->Cin) * kgeom_taps(d->Cout, 0, g.taps).num_kb * 64 * 2;
    size_t tot = 0;   // one image per output phase, back to back
    for (int pt = 0; pt < d->stride_t; ++pt)
      for (int py = 0; py < d->stride; ++py)
        for (int px = 0; px < d->stride; ++px) tot += phase_image_bytes(d, pt, py, px, nullptr);
    return tot;
  }
  fail(PLC_ERR_BAD_DESC, "bad pack kind %d", pack_kind);
  return 0;
}

int plc_convnd_pack_weight(const PlcConvNdDesc* d, int pack_kind, const float* w, const float* bias, void* w_packed,
                           float* bias_packed, void* stream) {
  NdGeom g;
  int rc = check_convnd(d, &g);
  if (rc) return rc;
  if (!w || !w_packed) return fail(PLC_ERR_NULL_ARG, "plc_convnd_pack_weight: null pointer");
  if (!aligned16(w_packed) || !aligned16(bias_packed)) return fail(PLC_ERR_ALIGNMENT, "packed buffers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LaunchTimer timer(PLC_K_PACK, st);
  if (pack_kind == PLC_PACK_FWD) {
    pack_w_conv_fwd_kernel<<<148 * 4, 256, 0, st>>>(w, bias, static_cast<__nv_bfloat16*>(w_packed), bias_packed, d->Cin,
                                                    d->Cout, g.taps, kgeom_taps(d->Cin, 0, g.taps), 0);
  } else if (pack_kind == PLC_PACK_DGRAD && !convnd_strided(d)) {
    pack_w_conv_dgrad_kernel<<<148 * 4, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(w_packed), d->Cin, d->Cout, g.taps,
                                                      kgeom_taps(d->Cout, 0, g.taps));
  } else if (pack_kind == PLC_PACK_DGRAD) {
    size_t off = 0;
    for (int pt = 0; pt < d->stride_t; ++pt)
      for (int py = 0; py < d->stride; ++py)
        for (int px = 0; px < d->stride; ++px) {
          int taps;
          const size_t bytes = phase_image_bytes(d, pt, py, px, &taps);
          if (!taps) continue;
          pack_w_phase_dgrad_kernel<<<148 * 2, 256, 0, st>>>(
              w, reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(w_packed) + off), d->Cin, d->Cout, d->kt, d->k,
              d->stride_t, d->stride, phase_dim(d->kt, d->stride_t, d->kt / 2, pt), phase_dim(d->k, d->stride, d->k / 2, py),
              phase_dim(d->k, d->stride, d->k / 2, px), kgeom_taps(d->Cout, 0, taps));
          off += bytes;
        }
  } else {
    return fail(PLC_ERR_BAD_DESC, "bad pack kind %d", pack_kind);
  }
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

int plc_convnd_fwd(const PlcConvNdDesc* d, const void* x, const void* w_packed_fwd, const float* bias_packed, void* out,
                   void* stream) {
  NdGeom g;
  int rc = check_convnd(d, &g);
  if (rc) return rc;
  if (!x || !w_packed_fwd || !out) return fail(PLC_ERR_NULL_ARG, "plc_convnd_fwd: null pointer");
  if (!aligned16(x) || !aligned16(w_packed_fwd) || !aligned16(out) || !aligned16(bias_packed))
    return fail(PLC_ERR_ALIGNMENT, "plc_convnd_fwd: all device pointers must be 16-byte aligned");
  return convnd_run(x, d->B, d->T, d->H, d->W, d->Cin, g.To, g.Ho, g.Wo, d->Cout, d->kt, d->k, d->stride_t, d->stride,
                    w_packed_fwd, d->has_bias ? bias_packed : nullptr, d->act != 0, d->act == 2 ? d->slope : 0.f, out,
                    PLC_K_CONV_FWD, static_cast<cudaStream_t>(stream));
}

// dz = dy * act'(y) on [B,To,Ho,Wo,C], 8 channels per thread
__global__ void convnd_grad_mask_kernel(const uint4* __restrict__ y, const uint4* __restrict__ dy, uint4* __restrict__ dz,
                                        size_t total, float slope) {
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 g = dy[idx], yv = y[idx];
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w};
    uint32_t ow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 g2 = *reinterpret_cast<const __nv_bfloat162*>(&gw[i]);
      const __nv_bfloat162 y2 = *reinterpret_cast<const __nv_bfloat162*>(&yw[i]);
      ow[i] = plc::pack_bf16x2(__low2float(g2) * (__low2float(y2) > 0.f ? 1.f : slope),
                               __high2float(g2) * (__high2float(y2) > 0.f ? 1.f : slope));
    }
    dz[idx] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

int plc_convnd_grad_mask(const PlcConvNdDesc* d, const void* y, const void* dy, void* dz, void* stream) {
  NdGeom g;
  int rc = check_convnd(d, &g);
  if (rc) return rc;
  if (!dy || !dz || !y) return fail(PLC_ERR_NULL_ARG, "plc_convnd_grad_mask: null pointer");
  if (d->act == 0) return fail(PLC_ERR_BAD_DESC, "plc_convnd_grad_mask: the layer has no activation (dZ == dY)");
  if (!aligned16(y) || !aligned16(dy) || !aligned16(dz))
    return fail(PLC_ERR_ALIGNMENT, "plc_convnd_grad_mask: all device pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LaunchTimer timer(PLC_K_ELEMENTWISE, st);
  const size_t total = static_cast<size_t>(d->B) * g.To * g.Ho * g.Wo * (d->Cout / 8);
  convnd_grad_mask_kernel<<<sm_count() * 8, 256, 0, st>>>(static_cast<const uint4*>(y), static_cast<const uint4*>(dy),
                                                          static_cast<uint4*>(dz), total, d->act == 2 ? d->slope : 0.f);
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

size_t plc_convnd_wgrad_acc_bytes(const PlcConvNdDesc* d) {
  NdGeom g;
  if (check_convnd(d, &g) != PLC_OK) return 0;
  return wgrad_acc_elems(PLC_MODE_BF16_TC, d->Cin, 0, d->Cout, g.taps) * sizeof(float);
}
int plc_convnd_wgrad_unpack(const PlcConvNdDesc* d, const float* acc, float* dW, void* stream) {
  NdGeom g;
  int rc = check_convnd(d, &g);
  if (rc) return rc;
  return wgrad_unpack(PLC_MODE_BF16_TC, d->Cin, 0, d->Cout, g.taps, acc, dW, static_cast<cudaStream_t>(stream));
}

// dX of a strided layer: one stride-1 launch per output phase, writing that phase's sub-lattice of dX
static int convnd_dgrad_phases(const PlcConvNdDesc* d, const NdGeom& g, const void* dz, const void* w_packed, void* dx,
                               cudaStream_t st) {
  int rc;
  size_t off = 0;
  for (int pt = 0; pt < d->stride_t; ++pt)
    for (int py = 0; py < d->stride; ++py)
      for (int px = 0; px < d->stride; ++px) {
        const PhaseDim dzp = phase_dim(d->kt, d->stride_t, d->kt / 2, pt);
        const PhaseDim dyp = phase_dim(d->k, d->stride, d->k / 2, py);
        const PhaseDim dxp = phase_dim(d->k, d->stride, d->k / 2, px);
        int taps;
        const size_t bytes = phase_image_bytes(d, pt, py, px, &taps);
        if (!taps)
          return fail(PLC_ERR_UNSUPPORTED, "transposed conv: kernel (%d,%d) smaller than stride (%d,%d) leaves dX phases "
                                           "without a tap", d->kt, d->k, d->stride_t, d->stride);
        const int Tp = (d->T - pt + d->stride_t - 1) / d->stride_t;
        const int Hp = (d->H - py + d->stride - 1) / d->stride, Wp = (d->W - px + d->stride - 1) / d->stride;
        if (Tp <= 0 || Hp <= 0 || Wp <= 0) { off += bytes; continue; }
        const int imgs = g.nd5 ? d->B * Tp : d->B * d->T;
        PlcCellDesc cd{imgs, Hp, Wp, d->Cout, d->Cin, d->k, PLC_MODE_BF16_TC, 0};
        TcGeom tg;
        pick_spatial_tile(Hp, Wp, &tg);
        plc::ConvTcParams q;
        fill_geom(&cd, tg, &q);
        const int nt = pick_plain_n_tile(d->Cin);
        q.num_n_tiles = cdiv(d->Cin, nt);
        q.num_tiles = q.num_m_tiles * q.num_n_tiles;
        set_kgeom(&q, kgeom_taps(d->Cout, 0, taps));
        q.nd5 = g.nd5; q.kt = dzp.n; q.T_out = Tp;
        q.ky_n = dyp.n; q.kx_n = dxp.n;
        q.tap_x0 = dxp.o0; q.tap_y0 = dyp.o0; q.tap_t0 = dzp.o0; q.tap_dir = -1;
        q.n_total = d->Cin; q.Cin = d->Cin;
        q.out0 = static_cast<__nv_bfloat16*>(dx);
        q.o_map = 1; q.o_s = d->stride; q.o_py = py; q.o_px = px; q.o_H = d->H; q.o_W = d->W;
        q.o_st = d->stride_t; q.o_pt = pt; q.o_T = d->T;
        const int cta = pick_cta_group(q.num_m_tiles);
        CUtensorMap ta, tb;
        if ((rc = make_tmap_act(&ta, dz, g.nd5 ? d->B : d->B * d->T, g.Ho, g.Wo, d->Cout, q.tw, q.th, 2, q.kc,
                                swizzle_for_kc(q.kc), 1, g.nd5 ? g.To : 0)))
          return rc;
        if ((rc = make_tmap_mat(&tb, static_cast<const char*>(w_packed) + off, d->Cin, (long)q.num_kb * 64, 64, nt / cta)))
          return rc;
        if ((rc = launch_conv_tc<plc::EPI_PLAIN>(nt, cta, q, ta, ta, tb, ta, ta, st, PLC_K_CONV_DGRAD,
                                                 conv_flops(imgs, Hp, Wp, d->Cout, d->Cin, 1) * taps)))
          return rc;
        off += bytes;
      }
  return PLC_OK;
}

int plc_convnd_bwd(const PlcConvNdDesc* d, const void* x, const void* dz, const void* w_packed_dgrad, void* dx,
                   float* dW_acc, float* db_acc, void* stream) {
  NdGeom g;
  int rc = check_convnd(d, &g);
  if (rc) return rc;
  if (!x || !dz) return fail(PLC_ERR_NULL_ARG, "plc_convnd_bwd: null pointer");
  if (dx && !w_packed_dgrad) return fail(PLC_ERR_NULL_ARG, "plc_convnd_bwd: dx needs the dgrad image");
  if (!aligned16(x) || !aligned16(dz) || !aligned16(dx) || !aligned16(w_packed_dgrad))
    return fail(PLC_ERR_ALIGNMENT, "plc_convnd_bwd: all device pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dx && convnd_strided(d)) {
    if ((rc = convnd_dgrad_phases(d, g, dz, w_packed_dgrad, dx, st))) return rc;
  } else if (dx) {
    // stride 1: dX = "same" conv of dZ with the flipped / transposed image
    if ((rc = convnd_run(dz, d->B, d->T, d->H, d->W, d->Cout, d->T, d->H, d->W, d->Cin, d->kt, d->k, 1, 1,
                         w_packed_dgrad, nullptr, 0, 0.f, dx, PLC_K_CONV_DGRAD, st)))
      return rc;
  }
  if (dW_acc) {
    WgradShape w;
    w.B = d->B * g.To; w.H = g.Ho; w.W = g.Wo; w.k = d->k; w.Cin = d->Cin; w.Ch = 0; w.N = d->Cout;
    w.stride = d->stride; w.kt = d->kt; w.stride_t = d->stride_t; w.T_out = g.To;
    w.Bs = g.nd5 ? d->B : d->B * d->T; w.Ts = d->T; w.Hs = d->H; w.Ws = d->W;
    if ((rc = launch_wgrad_tc_shape(&w, x, nullptr, dz, dW_acc, d->has_bias ? db_acc : nullptr, st, PLC_K_CONV_WGRAD)))
      return rc;
  }
  return PLC_OK;
}

// ---------------------------------------------------------------------------------- frame-level first layer
static int check_frameconv(const PlcFrameConvDesc* d, const char* fn) {
  if (!d) return fail(PLC_ERR_NULL_ARG, "%s: null descriptor", fn);
  if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->Cf < 1 || d->Cf > 4)
    return fail(PLC_ERR_BAD_DESC, "%s: need N, H, W > 0 and 1 <= Cf <= 4 (got N=%d H=%d W=%d Cf=%d)", fn, d->N, d->H, d->W,
                d->Cf);
  if (static_cast<long long>(d->N) * d->H >= (1ll << 31)) return fail(PLC_ERR_UNSUPPORTED, "%s: N*H must be < 2^31", fn);
  if (d->stride != 1 && d->stride != 2) return fail(PLC_ERR_UNSUPPORTED, "%s: stride must be 1 or 2 (got %d)", fn, d->stride);
  if (d->act < 0 || d->act > 2) return fail(PLC_ERR_BAD_DESC, "%s: act must be 0, 1 or 2", fn);
  const int G = d->Cout / 8;
  if (d->Cout % 8 || G < 1 || G > 32 || (G & (G - 1)))
    return fail(PLC_ERR_UNSUPPORTED, "%s: Cout must be 8, 16, 32, 64, 128 or 256 (got %d)", fn, d->Cout);
  return PLC_OK;
}

static void fill_frameconv(const PlcFrameConvDesc* d, plc::FrameConvParams* p) {
  memset(p, 0, sizeof(*p));
  p->N = d->N; p->Cf = d->Cf; p->H = d->H; p->W = d->W; p->Cout = d->Cout; p->stride = d->stride;
  p->Ho = (d->H - 1) / d->stride + 1; p->Wo = (d->W - 1) / d->stride + 1;
  p->act = d->act; p->slope = d->slope;
}

int plc_frameconv_out_shape(const PlcFrameConvDesc* d, int* H_out, int* W_out) {
  if (int rc = check_frameconv(d, "plc_frameconv_out_shape")) return rc;
  if (H_out) *H_out = (d->H - 1) / d->stride + 1;
  if (W_out) *W_out = (d->W - 1) / d->stride + 1;
  return PLC_OK;
}

int plc_frameconv_fwd(const PlcFrameConvDesc* d, const float* frames, const float* w_oihw, const float* bias, void* out,
                      void* stream) {
  if (int rc = check_frameconv(d, "plc_frameconv_fwd")) return rc;
  if (!frames || !w_oihw || !out || (d->has_bias && !bias)) return fail(PLC_ERR_NULL_ARG, "plc_frameconv_fwd: null pointer");
  if (!aligned16(out)) return fail(PLC_ERR_ALIGNMENT, "plc_frameconv_fwd: out must be 16-byte aligned");
  plc::FrameConvParams p;
  fill_frameconv(d, &p);
  p.frames = frames; p.w = w_oihw; p.bias = d->has_bias ? bias : nullptr;
  p.out = static_cast<__nv_bfloat16*>(out);
  const size_t rows = static_cast<size_t>(p.N) * p.Ho, cap = static_cast<size_t>(sm_count()) * 8;
  const size_t blocks = (rows + 7) / 8;           // 8 warps per block, one output row per warp at a time
  const size_t smem = (static_cast<size_t>(d->Cf) * 9 * d->Cout + d->Cout) * sizeof(float);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LaunchTimer timer(PLC_K_CONV_FWD, st, 2.0 * p.N * p.Ho * p.Wo * 9.0 * d->Cf * d->Cout);
  plc::frameconv_fwd_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), 256, smem, st>>>(p);
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

int plc_frameconv_bwd(const PlcFrameConvDesc* d, const float* frames, const float* w_oihw, const void* y, const void* dy,
                      float* dframes, float* dW, float* db, void* stream) {
  if (int rc = check_frameconv(d, "plc_frameconv_bwd")) return rc;
  if (!frames || !w_oihw || !dy || (d->act != 0 && !y)) return fail(PLC_ERR_NULL_ARG, "plc_frameconv_bwd: null pointer");
  if (!aligned16(y) || !aligned16(dy)) return fail(PLC_ERR_ALIGNMENT, "plc_frameconv_bwd: y / dy must be 16-byte aligned");
  plc::FrameConvParams p;
  fill_frameconv(d, &p);
  p.frames = frames; p.w = w_oihw;
  p.y = static_cast<const __nv_bfloat16*>(y); p.dy = static_cast<const __nv_bfloat16*>(dy);
  p.dframes = dframes; p.dW = dW; p.db = d->has_bias ? db : nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double flops = 2.0 * p.N * p.Ho * p.Wo * 9.0 * d->Cf * d->Cout;
  if (dW) {
    const size_t rows = static_cast<size_t>(p.N) * p.Ho, cap = static_cast<size_t>(sm_count()) * 4;
    const size_t blocks = (rows + 7) / 8;
    LaunchTimer timer(PLC_K_CONV_WGRAD, st, flops);
    plc::frameconv_wgrad_kernel<<<dim3(static_cast<unsigned>(blocks < cap ? blocks : cap), d->Cf), 256, 0, st>>>(p);
    PLC_CUDA(cudaGetLastError());
  }
  if (dframes) {
    const int R = d->stride == 2 ? 2 : 3;
    const size_t smem = (static_cast<size_t>(9) * d->Cf * d->Cout + static_cast<size_t>(R) * p.Wo * 9 * d->Cf) * sizeof(float);
    if (smem > 200 * 1024) return fail(PLC_ERR_UNSUPPORTED, "plc_frameconv_bwd: output row too wide for the dgrad ring (W=%d)", d->W);
    const int cap = sm_count() * 8;
    const unsigned grid = static_cast<unsigned>(p.N < cap ? p.N : cap);
    LaunchTimer timer(PLC_K_CONV_DGRAD, st, flops);
#define PLC_FC_DGRAD(SS, CC)                                                                                        \
  do {                                                                                                              \
    if (smem > 48 * 1024)                                                                                           \
      PLC_CUDA(cudaFuncSetAttribute(plc::frameconv_dgrad_kernel<SS, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    static_cast<int>(smem)));                                                       \
    plc::frameconv_dgrad_kernel<SS, CC><<<grid, 256, smem, st>>>(p);                                                \
  } while (0)
    if (d->stride == 2) {
      switch (d->Cf) { case 1: PLC_FC_DGRAD(2, 1); break; case 2: PLC_FC_DGRAD(2, 2); break;
                       case 3: PLC_FC_DGRAD(2, 3); break; default: PLC_FC_DGRAD(2, 4); }
    } else {
      switch (d->Cf) { case 1: PLC_FC_DGRAD(1, 1); break; case 2: PLC_FC_DGRAD(1, 2); break;
                       case 3: PLC_FC_DGRAD(1, 3); break; default: PLC_FC_DGRAD(1, 4); }
    }
#undef PLC_FC_DGRAD
    PLC_CUDA(cudaGetLastError());
  }
  return PLC_OK;
}

int plc_debug_set_patch(int mode) {
  if (mode < -1 || mode > 1) return fail(PLC_ERR_BAD_DESC, "plc_debug_set_patch: mode must be -1 (auto), 0 or 1");
  g_patch_override.store(mode);
  return PLC_OK;
}

int plc_debug_set_cta_group(int cta_group) {
  if (cta_group < 0 || cta_group > 2) return fail(PLC_ERR_BAD_DESC, "cta_group must be 0 (auto), 1 or 2");
  g_cta_override.store(cta_group);
  return PLC_OK;
}

int plc_timing_enable(int on) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (auto& t : g_timed) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
  g_timed.clear();
  g_timing_on.store(on ? 1 : 0);
  return PLC_OK;
}

int plc_timing_collect(int* kinds, float* ms, double* flops, int capacity) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  int n = 0;
  for (auto& t : g_timed) {
    float v = 0.f;
    if (cudaEventSynchronize(t.e1) != cudaSuccess || cudaEventElapsedTime(&v, t.e0, t.e1) != cudaSuccess) {
      cudaGetLastError();
      v = -1.f;
    }
    if (n < capacity && kinds && ms) {
      kinds[n] = t.kind; ms[n] = v;
      if (flops) flops[n] = t.flops;
    }
    ++n;
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  g_timed.clear();
  return n;
}

int plc_debug_set_prof(void* device_buf_u64) {
  g_prof_buf = static_cast<unsigned long long*>(device_buf_u64);
  return PLC_OK;
}

int plc_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int B, int C_src, int C_dst, int H, int W,
                              void* stream) {
  if (!src || !dst) return fail(PLC_ERR_NULL_ARG, "null pointer");
  if (B <= 0 || C_src <= 0 || C_dst < C_src || H <= 0 || W <= 0) return fail(PLC_ERR_BAD_DESC, "bad sizes");
  dim3 grid(cdiv(H * W, 32), cdiv(C_dst, 32), B), block(32, 8);
  LaunchTimer timer(PLC_K_ELEMENTWISE, static_cast<cudaStream_t>(stream));
  nchw_f32_to_nhwc_bf16_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), C_src, C_dst, H * W);
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

int plc_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int B, int C, int H, int W, void* stream) {
  if (!src || !dst) return fail(PLC_ERR_NULL_ARG, "null pointer");
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(PLC_ERR_BAD_DESC, "bad sizes");
  dim3 grid(cdiv(H * W, 32), cdiv(C, 32), B), block(32, 8);
  LaunchTimer timer(PLC_K_ELEMENTWISE, static_cast<cudaStream_t>(stream));
  nhwc_bf16_to_nchw_f32_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), dst, C, H * W);
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

int plc_frontend_fwd(const float* frames, int N, int Cf, int H, int W, const float* w_oihw, const float* bias, int C,
                     int C_stride, int mode, void* out, void* stream) {
  if (!frames || !w_oihw || !out) return fail(PLC_ERR_NULL_ARG, "plc_frontend_fwd: null pointer");
  if (N <= 0 || Cf <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 || C > 256 || C_stride < C || C_stride % 8)
    return fail(PLC_ERR_BAD_DESC, "plc_frontend_fwd: need C %% 8 == 0, C <= 256, C_stride >= C (got C=%d stride=%d)", C,
                C_stride);
  if (!aligned16(out)) return fail(PLC_ERR_ALIGNMENT, "plc_frontend_fwd: out must be 16-byte aligned");
  const int G = C / 8, P = G >= 256 ? 1 : 256 / G;
  if (G > 256) return fail(PLC_ERR_UNSUPPORTED, "plc_frontend_fwd: C must be <= 2048");
  const size_t quads = static_cast<size_t>(N) * H * ((W + plc::kFrontPx - 1) / plc::kFrontPx);
  const size_t groups = (quads + P - 1) / P;
  const size_t max_blocks = static_cast<size_t>(sm_count()) * 8;
  dim3 grid(static_cast<unsigned>(groups < max_blocks ? groups : max_blocks)), block(G, P);
  const size_t smem = (static_cast<size_t>(Cf + 2) * 9 * C + C) * sizeof(float);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LaunchTimer timer(PLC_K_FRONTEND, st);
  if (mode == PLC_MODE_BF16_TC) {
    PLC_CUDA(cudaFuncSetAttribute(plc::frontend_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    plc::frontend_kernel<__nv_bfloat16><<<grid, block, smem, st>>>(frames, w_oihw, bias, static_cast<__nv_bfloat16*>(out),
                                                                  N, Cf, H, W, C, C_stride);
  } else {
    PLC_CUDA(cudaFuncSetAttribute(plc::frontend_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    plc::frontend_kernel<float><<<grid, block, smem, st>>>(frames, w_oihw, bias, static_cast<float*>(out), N, Cf, H, W, C,
                                                          C_stride);
  }
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

int plc_head_fwd(const void* h, long npix, int C, const float* w, const float* bias, int mode, float* out,
                 void* stream) {
  if (!h || !w || !out) return fail(PLC_ERR_NULL_ARG, "plc_head_fwd: null pointer");
  if (npix <= 0 || C <= 0 || C % 8) return fail(PLC_ERR_BAD_DESC, "plc_head_fwd: need C %% 8 == 0");
  if (!aligned16(h)) return fail(PLC_ERR_ALIGNMENT, "plc_head_fwd: h must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LaunchTimer timer(PLC_K_HEAD, st);
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((npix + threads - 1) / threads);
  if (mode == PLC_MODE_BF16_TC) {
    const __nv_bfloat16* hb = static_cast<const __nv_bfloat16*>(h);
    const int G = C / 8;
    const unsigned cb = static_cast<unsigned>(((size_t)npix * G + threads - 1) / threads);
    switch (G) {
      case 1: plc::head_kernel_bf16_coalesced<1><<<cb, threads, 0, st>>>(hb, w, bias, out, (size_t)npix); break;
      case 2: plc::head_kernel_bf16_coalesced<2><<<cb, threads, 0, st>>>(hb, w, bias, out, (size_t)npix); break;
      case 4: plc::head_kernel_bf16_coalesced<4><<<cb, threads, 0, st>>>(hb, w, bias, out, (size_t)npix); break;
      case 8: plc::head_kernel_bf16_coalesced<8><<<cb, threads, 0, st>>>(hb, w, bias, out, (size_t)npix); break;
      case 16: plc::head_kernel_bf16_coalesced<16><<<cb, threads, 0, st>>>(hb, w, bias, out, (size_t)npix); break;
      case 32: plc::head_kernel_bf16_coalesced<32><<<cb, threads, 0, st>>>(hb, w, bias, out, (size_t)npix); break;
      default: plc::head_kernel<__nv_bfloat16><<<blocks, threads, 0, st>>>(hb, w, bias, out, (size_t)npix, C);
    }
  } else
    plc::head_kernel<float><<<blocks, threads, 0, st>>>(static_cast<const float*>(h), w, bias, out, (size_t)npix, C);
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

int plc_frames_to_nhwc(const float* frames, int B, int T, int Cf, int H, int W, int Cp, void* out, void* stream) {
  if (!frames || !out) return fail(PLC_ERR_NULL_ARG, "plc_frames_to_nhwc: null pointer");
  if (B <= 0 || T <= 0 || Cf <= 0 || H <= 0 || W <= 0 || Cp < Cf + 2 || Cp % 8)
    return fail(PLC_ERR_BAD_DESC, "plc_frames_to_nhwc: need Cp %% 8 == 0 and Cp >= Cf + 2 (got Cf=%d Cp=%d)", Cf, Cp);
  if (!aligned16(out)) return fail(PLC_ERR_ALIGNMENT, "plc_frames_to_nhwc: out must be 16-byte aligned");
  LaunchTimer timer(PLC_K_ELEMENTWISE, static_cast<cudaStream_t>(stream));
  plc::frames_to_nhwc_kernel<<<sm_count() * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      frames, static_cast<__nv_bfloat16*>(out), B, T, Cf, H, W, Cp);
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

int plc_head_bwd(const void* h, long npix, int C, const float* w, const float* dy, void* dh, float* dw_acc,
                 float* db_acc, void* stream) {
  if (!h || !w || !dy || !dh || !dw_acc) return fail(PLC_ERR_NULL_ARG, "plc_head_bwd: null pointer");
  const int G = C / 8;
  if (npix <= 0 || C % 8 || (G != 1 && G != 2 && G != 4 && G != 8 && G != 16 && G != 32))
    return fail(PLC_ERR_UNSUPPORTED, "plc_head_bwd: C must be 8, 16, 32, 64, 128 or 256 (got %d)", C);
  if (!aligned16(h) || !aligned16(dh)) return fail(PLC_ERR_ALIGNMENT, "plc_head_bwd: h/dh must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* hb = static_cast<const __nv_bfloat16*>(h);
  __nv_bfloat16* dhb = static_cast<__nv_bfloat16*>(dh);
  LaunchTimer timer(PLC_K_HEAD, st);
  const int blocks = sm_count() * 8;
#define PLC_HB(GG) case GG: plc::head_bwd_kernel_bf16<GG><<<blocks, 256, 0, st>>>(hb, w, dy, dhb, dw_acc, db_acc, (size_t)npix); break;
  switch (G) { PLC_HB(1) PLC_HB(2) PLC_HB(4) PLC_HB(8) PLC_HB(16) PLC_HB(32) }
#undef PLC_HB
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

static int loss_check(const PlcLossDesc* d, const char* fn) {
  if (!d) return fail(PLC_ERR_NULL_ARG, "%s: null descriptor", fn);
  if (d->B <= 0 || d->T <= 0 || d->H <= 0 || d->W <= 0 || d->scale <= 0 || d->n_stations < 0)
    return fail(PLC_ERR_BAD_DESC, "%s: B,T,H,W,scale must be positive (got B=%d T=%d H=%d W=%d scale=%d N=%d)", fn, d->B,
                d->T, d->H, d->W, d->scale, d->n_stations);
  if (d->weight_mode < 0 || d->weight_mode > 3)
    return fail(PLC_ERR_BAD_DESC, "%s: weight_mode must be 0 (off), 1 (log), 2 (sqrt) or 3 (stratified)", fn);
  if (static_cast<long long>(d->H) * d->scale > 0x7fffffffLL || static_cast<long long>(d->W) * d->scale > 0x7fffffffLL)
    return fail(PLC_ERR_BAD_DESC, "%s: high-resolution grid too large", fn);
  return PLC_OK;
}

size_t plc_loss_workspace_bytes(const PlcLossDesc* d) {
  if (loss_check(d, "plc_loss_workspace_bytes") != PLC_OK) return 0;
  return 64;   /* the reduction cells; the conservation residual signs stay in shared memory */
}

int plc_combined_loss(const PlcLossDesc* d, const float* pred, const float* lr_input, const long long* s_coords,
                      const float* s_values, void* workspace, float* terms_out, float* dpred, const float* grad_scale,
                      void* stream) {
  if (int rc = loss_check(d, "plc_combined_loss")) return rc;
  if (!pred || !lr_input || !workspace || !terms_out) return fail(PLC_ERR_NULL_ARG, "plc_combined_loss: null pointer");
  if (d->n_stations > 0 && (!s_coords || !s_values))
    return fail(PLC_ERR_NULL_ARG, "plc_combined_loss: n_stations > 0 needs s_coords and s_values");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LaunchTimer timer(PLC_K_LOSS, st);
  plc::LossParams p{};
  p.B = d->B; p.T = d->T; p.H = d->H; p.W = d->W; p.s = d->scale;
  p.Hs = d->H * d->scale; p.Ws = d->W * d->scale;
  p.n_st = d->n_stations; p.svals_has_batch = d->svals_has_batch; p.coord_scale = d->coord_scale;
  p.weight_mode = d->weight_mode;
  p.l_point = d->lambda_point; p.l_cons = d->lambda_conserve; p.l_smooth = d->lambda_smooth; p.l_temp = d->lambda_temporal;
  p.pred = pred; p.lr = lr_input; p.coords = s_coords; p.svals = s_values;
  p.sums = static_cast<float*>(workspace);
  p.dpred = dpred;
  p.grad_scale = grad_scale;
  PLC_CUDA(cudaMemsetAsync(workspace, 0, 64, st));
  const size_t cap = static_cast<size_t>(sm_count()) * 8;
  auto blocks = [&](size_t n) { return static_cast<unsigned>(std::min(cap, (n + 255) / 256)); };
  // block = 256 threads as (bx along x) x (by strips of ~8 rows); band = m LR rows ~ 8 * by HR rows; the sign table
  // (m * W floats) lives in shared memory
  const bool vec4 = p.Ws % 4 == 0 && aligned16(pred) && (!dpred || aligned16(dpred));
  int bx = 4;
  while (bx < 256 && bx * (vec4 ? 4 : 1) < p.Ws) bx *= 2;
  const int by = 256 / bx;
  const int m = std::min(p.H, std::max(1, (8 * by + p.s / 2) / p.s));
  const int rpt = (m * p.s + by - 1) / by;
  const int bands = (p.H + m - 1) / m;
  const size_t smem = sizeof(float) * m * p.W;
  const long long nblk = static_cast<long long>(p.B) * p.T * bands;
  if (smem > 160 * 1024 || nblk > 0x7fffffffLL)
    return fail(PLC_ERR_UNSUPPORTED, "plc_combined_loss: grid too large (W=%d, %lld blocks)", p.W, nblk);
  auto kern = vec4 ? plc::loss_grid_kernel<4> : plc::loss_grid_kernel<1>;
  if (smem > 48 * 1024)
    PLC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<static_cast<unsigned>(std::min<long long>(nblk, cap)), dim3(bx, by), smem, st>>>(p, m, bands,
                                                                                         static_cast<int>(nblk), rpt);
  const size_t n_obs = static_cast<size_t>(p.B) * p.T * p.n_st;
  if (n_obs) {
    plc::loss_point_kernel<<<blocks(n_obs), 256, 0, st>>>(p);
    if (dpred) plc::loss_point_grad_kernel<<<blocks(n_obs), 256, 0, st>>>(p);
  }
  plc::loss_finalize_kernel<<<1, 1, 0, st>>>(p, terms_out);
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

int plc_frontend_tc_supported(int Cf, int C, int C_stride) {
  return (Cf >= 1 && Cf <= 3 && C == plc::kFeN && C_stride == plc::kFeN) ? 1 : 0;
}

int plc_frontend_tc_fwd(const float* frames, int B, int T, int Cf, int H, int W, const float* w_oihw, const float* bias,
                        int C, void* out, void* stream) {
  if (!frames || !w_oihw || !out) return fail(PLC_ERR_NULL_ARG, "plc_frontend_tc_fwd: null pointer");
  if (B <= 0 || T <= 0 || H <= 0 || W <= 0) return fail(PLC_ERR_BAD_DESC, "plc_frontend_tc_fwd: bad sizes");
  if (!plc_frontend_tc_supported(Cf, C, C))
    return fail(PLC_ERR_UNSUPPORTED, "plc_frontend_tc_fwd: needs 1..3 frame channels and C == 64 (got Cf=%d C=%d); use "
                                     "plc_frames_to_nhwc + plc_conv_fwd", Cf, C);
  const long long total = static_cast<long long>(B) * T * H * W;
  if (total >= (1ll << 31) - 128) return fail(PLC_ERR_UNSUPPORTED, "plc_frontend_tc_fwd: B*T*H*W must be < 2^31");
  if (!aligned16(out)) return fail(PLC_ERR_ALIGNMENT, "plc_frontend_tc_fwd: out must be 16-byte aligned");
  int rc;
  CUtensorMap tm;
  if ((rc = make_tmap_mat(&tm, out, total, plc::kFeN, plc::kFeN, 128))) return rc;
  plc::FrontendTcParams p;
  p.B = B; p.T = T; p.H = H; p.W = W;
  p.num_tiles = static_cast<int>((total + 127) / 128);
  p.frames = frames; p.w = w_oihw; p.bias = bias;
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LaunchTimer timer(PLC_K_FRONTEND, st);
#define PLC_FE(CF)                                                                                         \
  case CF:                                                                                                 \
    PLC_CUDA(set_smem_once<IntTag<CF>>(plc::frontend_tc_kernel<CF>, plc::kFeSmemBytes));                   \
    plc::frontend_tc_kernel<CF><<<grid, plc::kFeThreads, plc::kFeSmemBytes, st>>>(p, tm);                  \
    break;
  switch (Cf) { PLC_FE(1) PLC_FE(2) PLC_FE(3) }
#undef PLC_FE
  PLC_CUDA(cudaGetLastError());
  return PLC_OK;
}

}  // extern "C"
