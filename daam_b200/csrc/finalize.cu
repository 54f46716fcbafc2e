// Finalize kernels: per-key bicubic upsample -> clamp -> mean over keys (-> normalise), word-map row mean,
// and the image-size expansion of a word map.
//
// Replaces DiffusionHeatMapHooker.compute_global_heat_map (daam/trace.py:109-130), GlobalHeatMap.
// compute_word_heat_map (daam/heatmap.py:121-123) and WordHeatMap.expand_as (daam/heatmap.py:77-93).
// The interpolation is torch's `upsample_bicubic2d` with align_corners=False: source index
// (dst + 0.5) * in/out - 0.5 (not clamped), Keys' cubic convolution with A = -0.75 on the 4 taps floor-1..floor+2,
// taps clamped to the border; rows are combined horizontally first, then vertically, all in fp32.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>

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
// One CTA owns one output band (BR = 8 or 4 rows x x columns) of one token row and walks the key classes (distinct
// source resolutions) one after the other. With an integer factor F the cubic weights depend only on the output phase
// (F distinct weight sets per axis), and the F x F outputs under one source pixel read the same 5 x 5 source window: a
// thread owns ONE source pixel and emits its F x F outputs from one window (25 values, 9-14 FMAs per output).
// The windows come from shared memory: the band's source rows (+2 halo rows each side, clamped at the borders) of a
// CHUNK of keys are streamed in with 16-byte cp.async (every thread has several independent copies in flight, two
// chunks double-buffered), so the key loop is no longer a chain of dependent global loads -- the round-1 kernel's bound
// (34.8 us for the 175-key SD-2.1 case, 222 registers, 0.86 waves) -- and the kernel fits two CTAs per SM.
// When a class has fewer source pixels per band than threads, the spare thread groups take every kg-th key and the
// groups are merged through the shared-memory band in a fixed order (deterministic sums). Arithmetic per key is
// bit-identical to bicubic_at (same taps, same weights, same order).
constexpr int kMaxClassKeys = 2048;        // key pointers of one class staged in shared memory
constexpr int kStageFloats = 4096;         // one chunk buffer (16 KB); two of them

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

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Streams chunk `chunk` of a factor-2 / factor-4 class into `buf` (one commit group): the band's source rows (+2 halo rows
// each side, border rows replicated like the clamped taps) of up to kc keys. Every thread copies one fixed 16-byte unit
// of every keys_par-th key. Shared by the class passes and by the cross-class prefetch (next class's first chunk).
struct ChunkGeom {
  int region, kc, n_src, kg, VR;
  __device__ __forceinline__ ChunkGeom(int side, int f, int br) {
    const int R = br / f;                              // source rows under the band
    VR = R + 4;                                        // + 2 halo rows above and below
    region = VR * side;                                // floats of one key's staged rows
    n_src = R * side;                                  // source pixels under the band (<= 256, checked by the host)
    kg = 256 / n_src;                                  // thread groups that split the keys
    kc = kStageFloats / region;                        // keys per chunk, a multiple of kg so every group keeps its stride
    kc -= kc % kg;
  }
};
__device__ __forceinline__ void issue_chunk(int side, int f, int band, int br, const float* const* keys, int nk, int chunk,
                                            float* buf) {
  const ChunkGeom G(side, f, br);
  const int row_units = side / 4, key_units = G.VR * row_units;     // 16-byte units
  const int keys_par = 256 / key_units, copy_k = (int)threadIdx.x / key_units;
  const int k0 = chunk * G.kc, kn = min(G.kc, nk - k0);
  if (copy_k < keys_par) {
    const int pos = (int)threadIdx.x - copy_k * key_units;
    const int vr = pos / row_units, c4 = pos - vr * row_units;
    const int src = min(max(band * (br / f) - 2 + vr, 0), side - 1) * side + 4 * c4;
    float* dst = buf + vr * side + 4 * c4;
    for (int k = copy_k; k < kn; k += keys_par) cp_async16(dst + k * G.region, keys[k0 + k] + src);
  }
  cp_async_commit();
}

// What to prefetch while a class is being merged: the first chunk of the next factor-2 / factor-4 class (into chunk
// buffer 1; the merge parks the key groups in buffer 0).
struct NextClass {
  int side, f, nk;                                     // f == 0: nothing to prefetch
  const float* const* keys;
};

// `first_buf`: the chunk buffer holding this class's chunk 0 (1 when the previous class prefetched it, else 0 and the
// chunk is issued here).
template <int F>
__device__ __forceinline__ void class_pass(const FinalizeParams& P, int side, int nk, int band, int br, float* tile,
                                           const float* const* keys, float* stage, bool prefetched, const NextClass& next) {
  const int x = P.x;
  const ChunkGeom G(side, F, br);
  const int region = G.region, n_src = G.n_src, kg = G.kg, kc = G.kc;
  const int n_chunks = (nk + kc - 1) / kc;
  const int first_buf = prefetched ? 1 : 0;
  auto issue = [&](int c) { issue_chunk(side, F, band, br, keys, nk, c, stage + ((c + first_buf) & 1) * kStageFloats); };

  const int group = (int)threadIdx.x / n_src;
  const int s = (int)threadIdx.x - group * n_src;
  const bool live = group < kg;
  const int ly = s / side, sx = s - ly * side;
  int ix[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) ix[j] = min(max(sx - 2 + j, 0), side - 1);
  PhaseWeights<F> pw;
  pw.init();
  float acc[F][F];
#pragma unroll
  for (int py = 0; py < F; ++py)
#pragma unroll
    for (int px = 0; px < F; ++px) acc[py][px] = 0.f;

  if (!prefetched) issue(0);
  for (int c = 0; c < n_chunks; ++c) {
    if (c + 1 < n_chunks) { issue(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();                                   // chunk c has landed for every thread
    if (live) {
      const float* buf = stage + ((c + first_buf) & 1) * kStageFloats + ly * side;
      const int kn = min(kc, nk - c * kc);
      for (int k = group; k < kn; k += kg) {
        const float* src = buf + k * region;
        float v[5][5];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
          for (int j = 0; j < 5; ++j) v[i][j] = src[i * side + ix[j]];
        add_key<F>(pw, v, acc);
      }
    }
    __syncthreads();                                   // buffer (c & 1) may be overwritten by chunk c + 2
  }
  // both chunk buffers are free now: the next class's first chunk streams into buffer 1 while this class is merged
  if (next.f) issue_chunk(next.side, next.f, band, br, next.keys, next.nk, 0, stage + kStageFloats);
  // merge the key groups in a fixed order (deterministic sums): every group parks its band in the (now free) first
  // chunk buffer, then each band element is summed over the groups by one thread
  const int band_elems = br * x;
  if (live) {
    float* mine = stage + group * band_elems + (ly * F) * x + sx * F;
#pragma unroll
    for (int py = 0; py < F; ++py)
#pragma unroll
      for (int px = 0; px < F; ++px) mine[py * x + px] = acc[py][px];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < band_elems; i += blockDim.x) {
    float sum = tile[i];
    for (int g = 0; g < kg; ++g) sum += stage[g * band_elems + i];
    tile[i] = sum;
  }
  __syncthreads();
}

// factor 1: bicubic at scale 1 is the identity, the class contributes clamp(src) -- coalesced float4 reads; when the
// band has fewer float4s than threads, the spare thread groups take every kg-th key (merged in a fixed order)
__device__ __forceinline__ void class_pass_identity(const FinalizeParams& P, int nk, int band, int br, float* tile,
                                                    const float* const* keys, float* stage, const NextClass& next) {
  const int x = P.x;
  // this pass reads its keys straight from global memory: the next class's first chunk streams in underneath it
  if (next.f) issue_chunk(next.side, next.f, band, br, next.keys, next.nk, 0, stage + kStageFloats);
  const int n4 = br * x / 4;
  const int kg = n4 >= 256 ? 1 : 256 / n4;
  const int passes = (n4 + 255) / 256;
  for (int pass = 0; pass < passes; ++pass) {
    const int group = n4 >= 256 ? 0 : (int)threadIdx.x / n4;
    const int i = n4 >= 256 ? pass * 256 + (int)threadIdx.x : (int)threadIdx.x % n4;
    const bool live = i < n4 && group < kg;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      const long long off = (long long)band * br * x + 4 * i;
#pragma unroll 8
      for (int k = group; k < nk; k += kg) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(keys[k] + off));
        acc.x += fmaxf(v.x, 0.f); acc.y += fmaxf(v.y, 0.f); acc.z += fmaxf(v.z, 0.f); acc.w += fmaxf(v.w, 0.f);
      }
    }
    if (kg == 1) {                                       // every float4 of the band has one owner
      if (live) {
        float4* dst = reinterpret_cast<float4*>(tile + 4 * i);
        float4 cur = *dst;
        cur.x += acc.x; cur.y += acc.y; cur.z += acc.z; cur.w += acc.w;
        *dst = cur;
      }
      __syncthreads();
    } else {                                             // park the groups' bands, then sum them in a fixed order
      if (live) *reinterpret_cast<float4*>(stage + group * (br * x) + 4 * i) = acc;
      __syncthreads();
      for (int e = threadIdx.x; e < br * x; e += blockDim.x) {
        float sum = tile[e];
        for (int g = 0; g < kg; ++g) sum += stage[g * (br * x) + e];
        tile[e] = sum;
      }
      __syncthreads();
    }
  }
}

struct ClassList {
  int n;
  int band_rows;            // 8, or 4 when that is what fills the machine / keeps a band's source pixels within one CTA
  int side[8];              // distinct source sides, all dividing x with factor 1, 2 or 4
  int key_begin[9];         // keys[] is ordered class by class: class c owns [key_begin[c], key_begin[c + 1])
  int key_slot[kMaxGroups]; // where group g's first selected key goes in keys[]
};

// grid: (x / band_rows bands, n_rows); dynamic smem: two chunk buffers + the band tile (band_rows * x floats)
__global__ void __launch_bounds__(256, 2) finalize_fast_kernel(const __grid_constant__ FinalizeParams P,
                                                               const __grid_constant__ ClassList C,
                                                               float* __restrict__ out) {
  extern __shared__ __align__(16) float dyn[];
  __shared__ const float* keys[kMaxClassKeys];
  float* stage = dyn;                                  // 2 x kStageFloats
  float* tile = dyn + 2 * kStageFloats;
  const int band = blockIdx.x, t = blockIdx.y, x = P.x, br = C.band_rows;
  // key pointers (token row t) of every selected key, class by class; one thread per key group
  for (int g = threadIdx.x; g < P.n_groups; g += blockDim.x) {
    const daam_key_group& G = P.g[g];
    const int h0 = G.head_sel < 0 ? 0 : G.head_sel;
    const int h1 = G.head_sel < 0 ? G.heads : G.head_sel + 1;
    const long long hw = (long long)G.h * G.w;
    for (int head = h0; head < h1; ++head) keys[C.key_slot[g] + head - h0] = G.acc + ((long long)head * G.tokens + t) * hw;
  }
  for (int i = threadIdx.x; i < br * x; i += blockDim.x) tile[i] = 0.f;
  __syncthreads();
  bool prefetched = false;                             // chunk 0 of class c is already streaming into buffer 1
  for (int c = 0; c < C.n; ++c) {
    const int side = C.side[c], f = x / side, nk = C.key_begin[c + 1] - C.key_begin[c];
    const float* const* ck = keys + C.key_begin[c];
    NextClass next = {0, 0, 0, nullptr};
    if (c + 1 < C.n && C.side[c + 1] != x)
      next = {C.side[c + 1], x / C.side[c + 1], C.key_begin[c + 2] - C.key_begin[c + 1], keys + C.key_begin[c + 1]};
    if (f == 1) class_pass_identity(P, nk, band, br, tile, ck, stage, next);
    else if (f == 2) class_pass<2>(P, side, nk, band, br, tile, ck, stage, prefetched, next);
    else class_pass<4>(P, side, nk, band, br, tile, ck, stage, prefetched, next);
    prefetched = next.f != 0;
  }
  __syncthreads();
  float* dst = out + (long long)t * x * x + (long long)band * br * x;
  for (int i = threadIdx.x; i < br * x; i += blockDim.x) dst[i] = tile[i] / (float)P.n_keys;
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

// ---- fused word list -> image-size masks -------------------------------------------------------------------------
// One cooperative launch for a LIST of words: gather-mean of the word's rows of the global map (heatmap.py:121-123) ->
// bicubic to (out_h, out_w) -> min / max over the image -> normalise / threshold (heatmap.py:77-93). CTA = (word, chunk
// of output pixels). The word map lives in shared memory; the min/max pass and the write pass both interpolate from it
// (16 shared loads + 20 FMAs per pixel), so nothing but the final image is written and nothing is read back: per-chunk
// partial min/max go through `scratch`, one grid-wide barrier separates the passes. With `absolute` there is no
// min/max pass and no barrier. Deterministic (no atomics).
constexpr int kMaxWords = 96;
constexpr int kMaxWordRows = 320;       // selected rows over all words of a launch
constexpr int kMaxChunks = 32;          // CTAs per word; scratch holds 2 floats per (word, chunk)

struct ExpandWordsParams {
  const float* maps;                    // [n_map_rows][x][x]
  float* word_maps;                     // optional [n_words][x][x]
  float* out;                           // [n_words][oh][ow]
  float* scratch;                       // [n_words][chunks][2]
  int x, oh, ow, n_words, chunks, absolute, use_threshold;
  float threshold;
  int row_begin[kMaxWords + 1];
  int rows[kMaxWordRows];
};

__device__ __forceinline__ float bicubic_shared(const float* sm, int w, const Taps& ty, const Taps& tx) {
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float* row = sm + ty.idx[i] * w;
    float r = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) r += tx.w[j] * row[tx.idx[j]];
    v += ty.w[i] * r;
  }
  return v;
}

__global__ void __launch_bounds__(256) expand_words_kernel(const __grid_constant__ ExpandWordsParams P) {
  extern __shared__ __align__(16) float wm[];          // the word map [x][x]
  __shared__ float red_lo[8], red_hi[8];
  const int word = blockIdx.x / P.chunks, chunk = blockIdx.x - word * P.chunks;
  const int x = P.x, xx = x * x, n = P.oh * P.ow;
  const int r0 = P.row_begin[word], r1 = P.row_begin[word + 1];
  for (int i = threadIdx.x; i < xx; i += blockDim.x) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += __ldg(P.maps + (long long)P.rows[r] * xx + i);
    s = s / (float)(r1 - r0);
    wm[i] = s;
    if (chunk == 0 && P.word_maps) P.word_maps[(long long)word * xx + i] = s;
  }
  __syncthreads();
  const int per = (n + P.chunks - 1) / P.chunks;
  const int begin = chunk * per, end = min(n, begin + per);
  float lo = 0.f, hi = 0.f;
  if (!P.absolute) {
    lo = INFINITY; hi = -INFINITY;
    for (int o = begin + threadIdx.x; o < end; o += blockDim.x) {
      const int oy = o / P.ow, ox = o - oy * P.ow;
      const float v = bicubic_shared(wm, x, make_taps(oy, x, P.oh), make_taps(ox, x, P.ow));
      lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, s));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if ((threadIdx.x & 31) == 0) { red_lo[threadIdx.x >> 5] = lo; red_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < (int)blockDim.x / 32; ++i) { lo = fminf(lo, red_lo[i]); hi = fmaxf(hi, red_hi[i]); }
      float* slot = P.scratch + 2 * ((long long)word * P.chunks + chunk);
      slot[0] = lo; slot[1] = hi;
      __threadfence();
    }
    cooperative_groups::this_grid().sync();
    if (threadIdx.x == 0) {
      lo = INFINITY; hi = -INFINITY;
      const volatile float* slots = P.scratch + 2 * (long long)word * P.chunks;
      for (int c = 0; c < P.chunks; ++c) { lo = fminf(lo, slots[2 * c]); hi = fmaxf(hi, slots[2 * c + 1]); }
      red_lo[0] = lo; red_hi[0] = hi;
    }
    __syncthreads();
    lo = red_lo[0]; hi = red_hi[0];
  }
  float* dst = P.out + (long long)word * n;
  for (int o = begin + threadIdx.x; o < end; o += blockDim.x) {
    const int oy = o / P.ow, ox = o - oy * P.ow;
    float v = bicubic_shared(wm, x, make_taps(oy, x, P.oh), make_taps(ox, x, P.ow));
    if (!P.absolute) v = (v - lo) / (hi - lo + 1e-8f);
    if (P.use_threshold) v = v > P.threshold ? 1.f : 0.f;
    dst[o] = v;
  }
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
  // fast path: every key is square with an integer factor 1 / 2 / 4 (all SD / SDXL layers that are ever traced) and
  // 16-byte-aligned rows (cp.async / float4)
  ClassList cls;
  cls.n = 0;
  bool fast = x % 16 == 0 && x <= 256 && p.n_keys <= kMaxClassKeys && !force_generic_finalize();
  for (int i = 0; i < n_groups && fast; ++i) {
    const daam_key_group& g = groups[i];
    if (g.h != g.w || x % g.h != 0 || (x / g.h != 1 && x / g.h != 2 && x / g.h != 4) ||
        reinterpret_cast<uintptr_t>(g.acc) % 16 != 0) { fast = false; break; }
    bool seen = false;
    for (int c = 0; c < cls.n; ++c) seen = seen || cls.side[c] == g.h;
    if (!seen) {
      if (cls.n == 8) { fast = false; break; }
      cls.side[cls.n++] = g.h;
    }
  }
  if (fast) {
    // 8-row bands unless that leaves the machine under-filled (< 2 CTAs per SM) or a band's source pixels of the
    // factor-2 class would exceed one CTA's 256 threads (x > 128); forcing either height measured the same within 1 %
    // for the 175-key SD-2.1 case
    cls.band_rows = ((x / 8) * n_rows >= 2 * dev.sm_count && x <= 128) ? 8 : 4;
    int next = 0;
    for (int c = 0; c < cls.n; ++c) {                   // keys[] of the kernel: class by class, groups in call order
      cls.key_begin[c] = next;
      for (int i = 0; i < n_groups; ++i)
        if (groups[i].h == cls.side[c]) {
          cls.key_slot[i] = next;
          next += groups[i].head_sel < 0 ? groups[i].heads : 1;
        }
    }
    cls.key_begin[cls.n] = next;
    const size_t smem = (2 * kStageFloats + (size_t)cls.band_rows * x) * sizeof(float);
    static std::once_flag attr_once[64];
    cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once[dev.device & 63], [&] {
      attr_err = cudaFuncSetAttribute(finalize_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((2 * kStageFloats + 8 * 256) * sizeof(float)));
    });
    DAAM_CUDA_TRY(attr_err);
    finalize_fast_kernel<<<dim3(x / cls.band_rows, n_rows), 256, smem, stream>>>(p, cls, out);
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

static int launch_expand_words(ExpandWordsParams& p, const DeviceInfo& dev, cudaStream_t stream) {
  const size_t smem = (size_t)p.x * p.x * sizeof(float);
  static std::mutex mu;
  static size_t configured_dev[64] = {};
  static int blocks_per_sm[64] = {};
  int per_sm;
  {
    std::lock_guard<std::mutex> lock(mu);
    size_t& configured = configured_dev[dev.device & 63];
    if (smem > configured) {
      if (smem > 48 * 1024)
        DAAM_CUDA_TRY(cudaFuncSetAttribute(expand_words_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
      blocks_per_sm[dev.device & 63] = 0;
    }
    if (blocks_per_sm[dev.device & 63] == 0) {
      int occ = 0;
      DAAM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, expand_words_kernel, 256, configured));
      blocks_per_sm[dev.device & 63] = occ < 1 ? 1 : occ;
    }
    per_sm = blocks_per_sm[dev.device & 63];
  }
  const int capacity = per_sm * dev.sm_count;          // a cooperative grid must be co-resident
  const int n = p.oh * p.ow;
  int done = 0;
  const int total = p.n_words;
  ExpandWordsParams q = p;
  while (done < total) {                                // more words than the device holds at once: several launches
    const int batch = total - done < capacity ? total - done : capacity;
    int chunks = capacity / batch;
    if (chunks > kMaxChunks) chunks = kMaxChunks;
    if (chunks > (n + 255) / 256) chunks = (n + 255) / 256;
    if (chunks < 1) chunks = 1;
    q.n_words = batch;
    q.chunks = chunks;
    q.out = p.out + (long long)done * n;
    q.word_maps = p.word_maps ? p.word_maps + (long long)done * p.x * p.x : nullptr;
    q.scratch = p.scratch + 2LL * kMaxChunks * done;
    for (int i = 0; i <= batch; ++i) q.row_begin[i] = p.row_begin[done + i];
    void* args[] = {&q};
    DAAM_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)expand_words_kernel, dim3(batch * chunks), dim3(256), args, smem,
                                              stream));
    count_launch();
    done += batch;
  }
  return DAAM_OK;
}

extern "C" int daam_expand_words(const float* global_maps, int32_t n_rows, int32_t x, const int32_t* rows,
                                 const int32_t* row_begin, int32_t n_words, int32_t out_h, int32_t out_w,
                                 int32_t absolute, int32_t use_threshold, float threshold, float* word_maps, float* out,
                                 float* scratch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!global_maps || !rows || !row_begin || !out || !scratch || x <= 0 || out_h <= 0 || out_w <= 0 || n_rows <= 0) { set_error("daam_expand_words: null pointer or non-positive size"); return DAAM_E_INVALID; }
  if (n_words <= 0) { set_error("daam_expand_words: empty word list"); return DAAM_E_INVALID; }
  if (n_words > kMaxWords) { set_error("daam_expand_words: %d words > %d", n_words, kMaxWords); return DAAM_E_UNSUPPORTED; }
  if (row_begin[0] != 0 || row_begin[n_words] > kMaxWordRows) { set_error("daam_expand_words: row_begin must start at 0 and select at most %d rows", kMaxWordRows); return DAAM_E_UNSUPPORTED; }
  if ((size_t)x * x * sizeof(float) > 200 * 1024) { set_error("daam_expand_words: x = %d does not fit shared memory", x); return DAAM_E_UNSUPPORTED; }
  DeviceInfo dev;
  if (int rc = get_device_info(&dev)) return rc;
  static thread_local ExpandWordsParams p;
  p.maps = global_maps; p.word_maps = word_maps; p.out = out; p.scratch = scratch;
  p.x = x; p.oh = out_h; p.ow = out_w; p.n_words = n_words; p.chunks = 1;
  p.absolute = absolute ? 1 : 0; p.use_threshold = use_threshold ? 1 : 0; p.threshold = threshold;
  for (int w = 0; w < n_words; ++w) {
    if (row_begin[w + 1] <= row_begin[w]) { set_error("daam_expand_words: word %d selects no row", w); return DAAM_E_INVALID; }
    p.row_begin[w] = row_begin[w];
  }
  p.row_begin[n_words] = row_begin[n_words];
  for (int i = 0; i < row_begin[n_words]; ++i) {
    int r = rows[i];
    if (r < 0) r += n_rows;   // torch-style negative index
    if (r < 0 || r >= n_rows) { set_error("daam_expand_words: row %d out of range [0, %d)", rows[i], n_rows); return DAAM_E_INVALID; }
    p.rows[i] = r;
  }
  return launch_expand_words(p, dev, stream);
}

extern "C" int daam_expand_as(const float* word_map, int32_t x, int32_t out_h, int32_t out_w, int32_t absolute,
                              int32_t use_threshold, float threshold, float* out, float* scratch, void* stream_) {
  // one word whose "rows" are the word map itself
  const int32_t rows[1] = {0}, row_begin[2] = {0, 1};
  if (!word_map) { set_error("daam_expand_as: null pointer or non-positive size"); return DAAM_E_INVALID; }
  return daam_expand_words(word_map, 1, x, rows, row_begin, 1, out_h, out_w, absolute, use_threshold, threshold, nullptr,
                           out, scratch, stream_);
}
