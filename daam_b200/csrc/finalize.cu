// Finalize kernels: per-key bicubic upsample -> clamp -> mean over keys (-> normalise), word-map row mean,
// and the image-size expansion of a word map.
//
// Replaces DiffusionHeatMapHooker.compute_global_heat_map (daam/trace.py:109-130), GlobalHeatMap.
// compute_word_heat_map (daam/heatmap.py:121-123) and WordHeatMap.expand_as (daam/heatmap.py:77-93).
// The interpolation is torch's `upsample_bicubic2d` with align_corners=False: source index
// (dst + 0.5) * in/out - 0.5 (not clamped), Keys' cubic convolution with A = -0.75 on the 4 taps floor-1..floor+2,
// taps clamped to the border; rows are combined horizontally first, then vertically, all in fp32.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace daam {
namespace {

constexpr int kMaxGroups = 160;   // key groups (layer x prompt slices) per finalize launch
constexpr int kMaxRows = 128;     // selected rows of a word map

struct FinalizeParams {
  int n_groups, x, n_rows, n_keys;
  daam_key_group g[kMaxGroups];
};

struct Taps {
  int idx[4];
  float w[4];
};

__device__ __forceinline__ float cubic_near(float t, float a) { return ((a + 2.f) * t - (a + 3.f)) * t * t + 1.f; }
__device__ __forceinline__ float cubic_far(float t, float a) { return ((a * t - 5.f * a) * t + 8.f * a) * t - 4.f * a; }

__device__ __forceinline__ Taps make_taps(int dst, int n_in, int n_out) {
  const float a = -0.75f;
  const float scale = (float)n_in / (float)n_out;
  const float src = scale * ((float)dst + 0.5f) - 0.5f;
  const float fl = floorf(src);
  const float t = src - fl;
  const int base = (int)fl;
  Taps r;
  r.w[0] = cubic_far(t + 1.f, a);
  r.w[1] = cubic_near(t, a);
  r.w[2] = cubic_near(1.f - t, a);
  r.w[3] = cubic_far(2.f - t, a);
#pragma unroll
  for (int i = 0; i < 4; ++i) r.idx[i] = min(max(base - 1 + i, 0), n_in - 1);
  return r;
}

__device__ __forceinline__ float bicubic_at(const float* __restrict__ src, int w, const Taps& ty, const Taps& tx) {
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float* row = src + ty.idx[i] * w;
    float r = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) r += tx.w[j] * __ldg(row + tx.idx[j]);
    v += ty.w[i] * r;
  }
  return v;
}

// grid: (ceil(x*x / 256), n_rows). One thread = one output element (row t, pixel o); it walks every selected key.
__global__ void __launch_bounds__(256) finalize_kernel(const __grid_constant__ FinalizeParams P, float* __restrict__ out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  const int x = P.x;
  if (o >= x * x) return;
  const int oy = o / x, ox = o - oy * x;
  float sum = 0.f;
  int ch = -1, cw = -1;
  Taps ty, tx;
  for (int g = 0; g < P.n_groups; ++g) {
    const daam_key_group& G = P.g[g];
    const int hw = G.h * G.w;
    const int h0 = G.head_sel < 0 ? 0 : G.head_sel;
    const int h1 = G.head_sel < 0 ? G.heads : G.head_sel + 1;
    const bool same = (G.h == x && G.w == x);
    if (!same && (G.h != ch || G.w != cw)) {
      ty = make_taps(oy, G.h, x);
      tx = make_taps(ox, G.w, x);
      ch = G.h; cw = G.w;
    }
    const float* base = G.acc + (long long)t * hw;
    const long long head_stride = (long long)G.tokens * hw;
    if (same) {   // scale 1: the cubic weights are exactly (0, 1, 0, 0)
#pragma unroll 4
      for (int head = h0; head < h1; ++head) sum += fmaxf(__ldg(base + head * head_stride + o), 0.f);
    } else {
#pragma unroll 2
      for (int head = h0; head < h1; ++head) sum += fmaxf(bicubic_at(base + head * head_stride, G.w, ty, tx), 0.f);
    }
  }
  out[(long long)t * x * x + o] = sum / (float)P.n_keys;
}


// ---- fast path: integer upsampling factors 1 / 2 / 4 -------------------------------------------------------------
// One CTA owns one output band (16 rows x x columns) of one token row and walks the key classes (distinct source
// resolutions) one after the other. Within a class a thread owns ONE source pixel and produces its F x F output block:
// with an integer factor the cubic weights depend only on the output phase (F distinct weight sets per axis), and the
// F x F outputs of a source pixel read the same 5 x 5 source window -- 25 loads and 9-14 FMAs per output instead of 16
// loads and 20 FMAs per output for the generic gather. When a class has fewer source pixels per band than threads, the
// spare thread groups take every kg-th key and the groups are merged through the shared-memory band in a fixed order
// (deterministic sums). Arithmetic per key is bit-identical to bicubic_at (same taps, same weights, same order).
constexpr int kBandRows = 8;
constexpr int kMaxClassKeys = 2048;        // key pointers of one class staged in shared memory

template <int F>
struct PhaseWeights {
  float w[F][4];
  __device__ __forceinline__ void init() {
    const float a = -0.75f;
#pragma unroll
    for (int p = 0; p < F; ++p) {
      const float src = (1.0f / (float)F) * ((float)p + 0.5f) - 0.5f;   // source coordinate relative to the block's pixel
      const float t = src - floorf(src);
      w[p][0] = cubic_far(t + 1.f, a);
      w[p][1] = cubic_near(t, a);
      w[p][2] = cubic_near(1.f - t, a);
      w[p][3] = cubic_far(2.f - t, a);
    }
  }
};

// Fills keys[] (shared) with the source pointer (token row t) of every selected key of the class; returns the count.
__device__ __forceinline__ int stage_class_keys(const FinalizeParams& P, int cls_side, int t, const float** keys) {
  __shared__ int n_keys_s;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int g = 0; g < P.n_groups; ++g) {
      const daam_key_group& G = P.g[g];
      if (G.h != cls_side || G.w != cls_side) continue;
      const int h0 = G.head_sel < 0 ? 0 : G.head_sel;
      const int h1 = G.head_sel < 0 ? G.heads : G.head_sel + 1;
      const long long hw = (long long)G.h * G.w;
      for (int head = h0; head < h1; ++head) keys[n++] = G.acc + ((long long)head * G.tokens + t) * hw;
    }
    n_keys_s = n;
  }
  __syncthreads();
  return n_keys_s;
}

template <int F>
__device__ __forceinline__ void load_window(const float* src, const int (&iy)[5], const int (&ix)[5], float (&v)[5][5]) {
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) v[i][j] = __ldg(src + iy[i] + ix[j]);
}

template <int F>
__device__ __forceinline__ void add_key(const PhaseWeights<F>& pw, const float (&v)[5][5], float (&acc)[F][F]) {
  float r[5][F];                                       // horizontal pass, per source row and output phase
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int px = 0; px < F; ++px) {
      const int off = px < F / 2 ? 0 : 1;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) q += pw.w[px][j] * v[i][off + j];
      r[i][px] = q;
    }
#pragma unroll
  for (int py = 0; py < F; ++py) {
    const int off = py < F / 2 ? 0 : 1;
#pragma unroll
    for (int px = 0; px < F; ++px) {
      float o = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) o += pw.w[py][i] * r[off + i][px];
      acc[py][px] += fmaxf(o, 0.f);
    }
  }
}

template <int F>
__device__ __forceinline__ void class_pass(const FinalizeParams& P, int cls_h, int t, int band, float* tile,
                                           const float** keys) {
  constexpr int R = kBandRows / F;                     // source rows under this band
  const int x = P.x, w = cls_h, h = cls_h;
  const int n_src = R * w;
  const int kg = n_src >= 256 ? 1 : 256 / n_src;       // thread groups that split the keys
  const int passes = (n_src + 255) / 256;
  const int nk = stage_class_keys(P, cls_h, t, keys);
  PhaseWeights<F> pw;
  pw.init();
  for (int pass = 0; pass < passes; ++pass) {
    const int group = n_src >= 256 ? 0 : (int)threadIdx.x / n_src;
    const int s = n_src >= 256 ? pass * 256 + (int)threadIdx.x : (int)threadIdx.x % n_src;
    const bool live = s < n_src && group < kg;
    const int sy = band * R + s / w, sx = s % w;
    int iy[5], ix[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      iy[i] = min(max(sy - 2 + i, 0), h - 1) * w;
      ix[i] = min(max(sx - 2 + i, 0), w - 1);
    }
    float acc[F][F];
#pragma unroll
    for (int py = 0; py < F; ++py)
#pragma unroll
      for (int px = 0; px < F; ++px) acc[py][px] = 0.f;
    if (live) {
      // two windows in flight: the loads of the next key are issued before the arithmetic of the current one
      float va[5][5], vb[5][5];
      int k = group;
      if (k < nk) load_window<F>(keys[k], iy, ix, va);
      for (; k < nk; k += 2 * kg) {
        const bool has_b = k + kg < nk;
        if (has_b) load_window<F>(keys[k + kg], iy, ix, vb);
        add_key<F>(pw, va, acc);
        if (k + 2 * kg < nk) load_window<F>(keys[k + 2 * kg], iy, ix, va);
        if (has_b) add_key<F>(pw, vb, acc);
      }
    }
    for (int g = 0; g < kg; ++g) {                      // merge the key groups in a fixed order
      if (live && group == g) {
        const int oy0 = (s / w) * F, ox0 = sx * F;
#pragma unroll
        for (int py = 0; py < F; ++py)
#pragma unroll
          for (int px = 0; px < F; ++px) tile[(oy0 + py) * x + ox0 + px] += acc[py][px];
      }
      __syncthreads();
    }
  }
}

// factor 1: bicubic at scale 1 is the identity, the class contributes clamp(src) -- coalesced float4 reads; when the
// band has fewer float4s than threads, the spare thread groups take every kg-th key (merged in a fixed order)
__device__ __forceinline__ void class_pass_identity(const FinalizeParams& P, int t, int band, float* tile,
                                                    const float** keys) {
  const int x = P.x;
  const int n4 = kBandRows * x / 4;
  const int nk = stage_class_keys(P, x, t, keys);
  const int kg = n4 >= 256 ? 1 : 256 / n4;
  const int passes = (n4 + 255) / 256;
  for (int pass = 0; pass < passes; ++pass) {
    const int group = n4 >= 256 ? 0 : (int)threadIdx.x / n4;
    const int i = n4 >= 256 ? pass * 256 + (int)threadIdx.x : (int)threadIdx.x % n4;
    const bool live = i < n4 && group < kg;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      const long long off = (long long)band * kBandRows * x + 4 * i;
#pragma unroll 4
      for (int k = group; k < nk; k += kg) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(keys[k] + off));
        acc.x += fmaxf(v.x, 0.f); acc.y += fmaxf(v.y, 0.f); acc.z += fmaxf(v.z, 0.f); acc.w += fmaxf(v.w, 0.f);
      }
    }
    for (int g = 0; g < kg; ++g) {
      if (live && group == g) {
        float4* dst = reinterpret_cast<float4*>(tile + 4 * i);
        float4 cur = *dst;
        cur.x += acc.x; cur.y += acc.y; cur.z += acc.z; cur.w += acc.w;
        *dst = cur;
      }
      __syncthreads();
    }
  }
}

struct ClassList {
  int n;
  int side[8];      // distinct source sides, all dividing x with factor 1, 2 or 4
};

// grid: (x / kBandRows bands, n_rows); dynamic smem: kBandRows * x floats
__global__ void __launch_bounds__(256) finalize_fast_kernel(const __grid_constant__ FinalizeParams P,
                                                            const __grid_constant__ ClassList C,
                                                            float* __restrict__ out) {
  extern __shared__ __align__(16) float tile[];
  __shared__ const float* keys[kMaxClassKeys];
  const int band = blockIdx.x, t = blockIdx.y, x = P.x;
  for (int i = threadIdx.x; i < kBandRows * x; i += blockDim.x) tile[i] = 0.f;
  __syncthreads();
  for (int c = 0; c < C.n; ++c) {
    const int side = C.side[c], f = x / side;
    if (f == 1) class_pass_identity(P, t, band, tile, keys);
    else if (f == 2) class_pass<2>(P, side, t, band, tile, keys);
    else class_pass<4>(P, side, t, band, tile, keys);
  }
  float* dst = out + (long long)t * x * x + (long long)band * kBandRows * x;
  for (int i = threadIdx.x; i < kBandRows * x; i += blockDim.x) dst[i] = tile[i] / (float)P.n_keys;
}

// One output map per selected key (no mean): out[key][row][x][x] = clamp(bicubic(key[row])). grid: (ceil(x*x/256), n_rows,
// n_keys); key k belongs to group g with first_key[g] <= k < first_key[g + 1].
struct PerKeyParams {
  int n_groups, x, n_rows, n_keys;
  int first_key[kMaxGroups + 1];
  daam_key_group g[kMaxGroups];
};

__global__ void __launch_bounds__(256) finalize_per_key_kernel(const __grid_constant__ PerKeyParams P, float* __restrict__ out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y, key = blockIdx.z;
  const int x = P.x;
  if (o >= x * x) return;
  int g = 0;
  while (g + 1 < P.n_groups && key >= P.first_key[g + 1]) ++g;
  const daam_key_group& G = P.g[g];
  const int head = (G.head_sel < 0 ? 0 : G.head_sel) + (key - P.first_key[g]);
  const int hw = G.h * G.w;
  const float* src = G.acc + ((long long)head * G.tokens + t) * hw;
  float v;
  if (G.h == x && G.w == x) {
    v = __ldg(src + o);
  } else {
    const int oy = o / x, ox = o - oy * x;
    v = bicubic_at(src, G.w, make_taps(oy, G.h, x), make_taps(ox, G.w, x));
  }
  out[((long long)key * P.n_rows + t) * x * x + o] = fmaxf(v, 0.f);
}

// maps / (maps[1:-1].sum(0) + 1e-6), in place (daam/trace.py:129-130)
__global__ void normalize_kernel(float* __restrict__ maps, int n_rows, int xx) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= xx) return;
  maps += (long long)blockIdx.y * n_rows * xx;      // blockIdx.y: independent map stacks (per-key finalize)
  float s = 0.f;
  for (int t = 1; t < n_rows - 1; ++t) s += maps[(long long)t * xx + o];
  s += 1e-6f;
  for (int t = 0; t < n_rows; ++t) maps[(long long)t * xx + o] = maps[(long long)t * xx + o] / s;
}

struct RowSel {
  int n;
  int rows[kMaxRows];
};

__global__ void word_map_kernel(const float* __restrict__ maps, const __grid_constant__ RowSel sel, int xx,
                                float* __restrict__ out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= xx) return;
  float s = 0.f;
  for (int i = 0; i < sel.n; ++i) s += __ldg(maps + (long long)sel.rows[i] * xx + o);
  out[o] = s / (float)sel.n;
}

// ---- expand_as ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ordered(float f) {   // monotone float -> uint map for atomicMin/Max
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unordered(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void expand_init_kernel(unsigned* mm) { mm[0] = 0xffffffffu; mm[1] = 0u; }

__global__ void __launch_bounds__(256) expand_upsample_kernel(const float* __restrict__ src, int x, int oh, int ow,
                                                              float* __restrict__ out, unsigned* mm) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  float v = 0.f;
  const bool live = o < oh * ow;
  if (live) {
    const int oy = o / ow, ox = o - oy * ow;
    const Taps ty = make_taps(oy, x, oh), tx = make_taps(ox, x, ow);
    v = bicubic_at(src, x, ty, tx);
    out[o] = v;
  }
  float lo = live ? v : INFINITY, hi = live ? v : -INFINITY;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm, ordered(lo));
    atomicMax(mm + 1, ordered(hi));
  }
}

__global__ void expand_normalize_kernel(float* __restrict__ out, int n, const unsigned* __restrict__ mm, int absolute,
                                        int use_threshold, float threshold) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  float v = out[o];
  if (!absolute) {
    const float lo = unordered(mm[0]), hi = unordered(mm[1]);
    v = (v - lo) / (hi - lo + 1e-8f);
  }
  if (use_threshold) v = v > threshold ? 1.f : 0.f;
  out[o] = v;
}

}  // namespace
}  // namespace daam

using namespace daam;

// DAAM_FINALIZE_GENERIC=1 forces the generic gather kernel (tests compare the two paths)
static bool force_generic_finalize() {
  const char* e = getenv("DAAM_FINALIZE_GENERIC");
  return e && e[0] == '1';
}

extern "C" int daam_finalize(const daam_key_group* groups, int32_t n_groups, int32_t x, int32_t n_rows,
                             int32_t normalize, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!groups || !out || x <= 0 || n_rows <= 0) { set_error("daam_finalize: null pointer or non-positive size"); return DAAM_E_INVALID; }
  if (n_groups <= 0) { set_error("daam_finalize: no key selected"); return DAAM_E_INVALID; }
  if (n_groups > kMaxGroups) { set_error("daam_finalize: %d key groups > %d", n_groups, kMaxGroups); return DAAM_E_UNSUPPORTED; }
  DeviceInfo dev;
  if (int rc = get_device_info(&dev)) return rc;
  FinalizeParams p;
  p.n_groups = n_groups; p.x = x; p.n_rows = n_rows; p.n_keys = 0;
  for (int i = 0; i < n_groups; ++i) {
    const daam_key_group& g = groups[i];
    if (!g.acc || g.heads <= 0 || g.h <= 0 || g.w <= 0 || g.tokens < n_rows || g.head_sel >= g.heads) {
      set_error("daam_finalize: bad key group %d (heads %d, h %d, w %d, tokens %d, head_sel %d, n_rows %d)", i, g.heads,
                g.h, g.w, g.tokens, g.head_sel, n_rows);
      return DAAM_E_INVALID;
    }
    p.g[i] = g;
    p.n_keys += g.head_sel < 0 ? g.heads : 1;
  }
  const int xx = x * x;
  // fast path: every key is square with an integer factor 1 / 2 / 4 (all SD / SDXL layers that are ever traced)
  ClassList cls;
  cls.n = 0;
  bool fast = x % kBandRows == 0 && x % 4 == 0 && x <= 256 && p.n_keys <= kMaxClassKeys && !force_generic_finalize();
  for (int i = 0; i < n_groups && fast; ++i) {
    const daam_key_group& g = groups[i];
    if (g.h != g.w || x % g.h != 0 || (x / g.h != 1 && x / g.h != 2 && x / g.h != 4) ||
        (x / g.h == 1 && reinterpret_cast<uintptr_t>(g.acc) % 16 != 0)) { fast = false; break; }
    bool seen = false;
    for (int c = 0; c < cls.n; ++c) seen = seen || cls.side[c] == g.h;
    if (!seen) {
      if (cls.n == 8) { fast = false; break; }
      cls.side[cls.n++] = g.h;
    }
  }
  if (fast) {
    finalize_fast_kernel<<<dim3(x / kBandRows, n_rows), 256, kBandRows * x * sizeof(float), stream>>>(p, cls, out);
  } else {
    dim3 grid((xx + 255) / 256, n_rows);
    finalize_kernel<<<grid, 256, 0, stream>>>(p, out);
  }
  DAAM_CUDA_TRY(cudaGetLastError());
  count_launch();
  if (normalize) {
    normalize_kernel<<<(xx + 255) / 256, 256, 0, stream>>>(out, n_rows, xx);
    DAAM_CUDA_TRY(cudaGetLastError());
    count_launch();
  }
  return DAAM_OK;
}

extern "C" int daam_finalize_per_key(const daam_key_group* groups, int32_t n_groups, int32_t x, int32_t n_rows,
                                     int32_t normalize, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!groups || !out || x <= 0 || n_rows <= 0) { set_error("daam_finalize_per_key: null pointer or non-positive size"); return DAAM_E_INVALID; }
  if (n_groups <= 0) { set_error("daam_finalize_per_key: no key selected"); return DAAM_E_INVALID; }
  if (n_groups > kMaxGroups) { set_error("daam_finalize_per_key: %d key groups > %d", n_groups, kMaxGroups); return DAAM_E_UNSUPPORTED; }
  DeviceInfo dev;
  if (int rc = get_device_info(&dev)) return rc;
  static thread_local PerKeyParams p;
  p.n_groups = n_groups; p.x = x; p.n_rows = n_rows; p.n_keys = 0;
  for (int i = 0; i < n_groups; ++i) {
    const daam_key_group& g = groups[i];
    if (!g.acc || g.heads <= 0 || g.h <= 0 || g.w <= 0 || g.tokens < n_rows || g.head_sel >= g.heads) {
      set_error("daam_finalize_per_key: bad key group %d", i);
      return DAAM_E_INVALID;
    }
    p.g[i] = g;
    p.first_key[i] = p.n_keys;
    p.n_keys += g.head_sel < 0 ? g.heads : 1;
  }
  p.first_key[n_groups] = p.n_keys;
  if (p.n_keys > 65535) { set_error("daam_finalize_per_key: %d keys > 65535", p.n_keys); return DAAM_E_UNSUPPORTED; }
  const int xx = x * x;
  dim3 grid((xx + 255) / 256, n_rows, p.n_keys);
  finalize_per_key_kernel<<<grid, 256, 0, stream>>>(p, out);
  DAAM_CUDA_TRY(cudaGetLastError());
  count_launch();
  if (normalize) {   // every key's map is an independent [n_rows, x, x] block
    normalize_kernel<<<dim3((xx + 255) / 256, p.n_keys), 256, 0, stream>>>(out, n_rows, xx);
    DAAM_CUDA_TRY(cudaGetLastError());
    count_launch();
  }
  return DAAM_OK;
}

extern "C" int daam_word_heat_map(const float* global_maps, int32_t n_rows, int32_t x, const int32_t* rows,
                                  int32_t n_sel, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!global_maps || !rows || !out || x <= 0 || n_sel <= 0) { set_error("daam_word_heat_map: null pointer or empty selection"); return DAAM_E_INVALID; }
  if (n_sel > kMaxRows) { set_error("daam_word_heat_map: %d rows > %d", n_sel, kMaxRows); return DAAM_E_UNSUPPORTED; }
  DeviceInfo dev;
  if (int rc = get_device_info(&dev)) return rc;
  RowSel sel;
  sel.n = n_sel;
  for (int i = 0; i < n_sel; ++i) {
    int r = rows[i];
    if (r < 0) r += n_rows;   // torch-style negative index
    if (r < 0 || r >= n_rows) { set_error("daam_word_heat_map: row %d out of range [0, %d)", rows[i], n_rows); return DAAM_E_INVALID; }
    sel.rows[i] = r;
  }
  const int xx = x * x;
  word_map_kernel<<<(xx + 255) / 256, 256, 0, stream>>>(global_maps, sel, xx, out);
  DAAM_CUDA_TRY(cudaGetLastError());
  count_launch();
  return DAAM_OK;
}

extern "C" int daam_expand_as(const float* word_map, int32_t x, int32_t out_h, int32_t out_w, int32_t absolute,
                              int32_t use_threshold, float threshold, float* out, float* scratch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!word_map || !out || !scratch || x <= 0 || out_h <= 0 || out_w <= 0) { set_error("daam_expand_as: null pointer or non-positive size"); return DAAM_E_INVALID; }
  DeviceInfo dev;
  if (int rc = get_device_info(&dev)) return rc;
  unsigned* mm = reinterpret_cast<unsigned*>(scratch);
  const int n = out_h * out_w;
  expand_init_kernel<<<1, 1, 0, stream>>>(mm);
  expand_upsample_kernel<<<(n + 255) / 256, 256, 0, stream>>>(word_map, x, out_h, out_w, out, mm);
  expand_normalize_kernel<<<(n + 255) / 256, 256, 0, stream>>>(out, n, mm, absolute, use_threshold, threshold);
  DAAM_CUDA_TRY(cudaGetLastError());
  count_launch(3);
  return DAAM_OK;
}
