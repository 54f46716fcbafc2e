// Fused softmax(QK^T) -> unravel -> accumulate, SIMT fp32 variant ("warp dot" path).
//
// Serves fp32 projections (BASELINE config 1: the reference's own fp32 numerics, which tensor cores cannot give)
// and every head_dim the tcgen05 variant does not take (SD-1.x: 40/80/160). One thread owns one pixel: its 77
// logits live in registers, K^T sits in shared memory and is read as warp-wide broadcasts, so softmax needs no
// shuffles and the accumulator update `acc[t][pixel] += p[t]` is one fully coalesced 128-byte access per warp and
// token. Replaces daam/trace.py:276 (get_attention_scores), :219-244 (_unravel_attn) and :293-294 (update loop).
#include "common.cuh"

namespace daam {
namespace {

template <typename T> struct Vec;  // 16-byte global loads converted to fp32
template <> struct Vec<float> {
  static constexpr int kElems = 4;
  static __device__ __forceinline__ void load(const float* p, float* out) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  }
  static __device__ __forceinline__ float one(const float* p) { return __ldg(p); }
};
template <> struct Vec<__half> {
  static constexpr int kElems = 8;
  static __device__ __forceinline__ void load(const __half* p, float* out) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ float one(const __half* p) { return __half2float(__ldg(p)); }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int kElems = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* out) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ float one(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }
};

// Stage K^T (ks[dim][token], tokens padded to 80 with zeros) and the Q tile (qs[pixel][dim], row stride d+1: odd,
// hence bank-conflict free for the per-thread row walk) into shared memory with coalesced 16-byte global loads.
template <typename T>
__device__ __forceinline__ void stage_tile(const LayerParams& L, int prompt, int head, int pixel0, float* ks,
                                           float* qs, bool load_k) {
  const int d = L.head_dim;
  const T* kbase = static_cast<const T*>(L.k) + prompt * L.ks_prompt + head * L.ks_head;
  const T* qbase = static_cast<const T*>(L.q) + prompt * L.qs_prompt + head * L.qs_head;
  constexpr int V = Vec<T>::kElems;
  const int qstride = d + 1;
  if (L.vec_ok) {
    const int vec_per_row = d / V;
    for (int c = threadIdx.x; load_k && c < kTokensPad * vec_per_row; c += blockDim.x) {
      const int t = c / vec_per_row, v = c - t * vec_per_row;
      float f[V];
      if (t < kTokens) {
        Vec<T>::load(kbase + t * L.ks_token + v * V, f);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) f[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < V; ++i) ks[(v * V + i) * kTokensPad + t] = f[i];
    }
    for (int c = threadIdx.x; c < kTilePixels * vec_per_row; c += blockDim.x) {
      const int r = c / vec_per_row, v = c - r * vec_per_row;
      float f[V];
      if (pixel0 + r < L.hw) {
        Vec<T>::load(qbase + (long long)(pixel0 + r) * L.qs_pixel + v * V, f);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) f[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < V; ++i) qs[r * qstride + v * V + i] = f[i];
    }
  } else {  // unaligned views: scalar loads
    for (int c = threadIdx.x; load_k && c < kTokensPad * d; c += blockDim.x) {
      const int t = c / d, e = c - t * d;
      ks[e * kTokensPad + t] = t < kTokens ? Vec<T>::one(kbase + t * L.ks_token + e) : 0.f;
    }
    for (int c = threadIdx.x; c < kTilePixels * d; c += blockDim.x) {
      const int r = c / d, e = c - r * d;
      qs[r * qstride + e] = pixel0 + r < L.hw ? Vec<T>::one(qbase + (long long)(pixel0 + r) * L.qs_pixel + e) : 0.f;
    }
  }
}

__global__ void __launch_bounds__(kTilePixels, 3) accumulate_simt_kernel(const __grid_constant__ LaunchParams P) {
  extern __shared__ __align__(16) float smem[];
  // contiguous chunk of tiles per CTA: consecutive tiles share (layer, prompt, head), so K^T is staged once per run
  const int per = P.total_tiles / gridDim.x, rem = P.total_tiles % gridDim.x;
  const int first = blockIdx.x * per + min((int)blockIdx.x, rem);
  const int count = per + ((int)blockIdx.x < rem ? 1 : 0);

  int li = 0, last_run = -1;
  for (int tile = first; tile < first + count; ++tile) {
    while (li + 1 < P.n_layers && tile >= P.layer[li + 1].tile_begin) ++li;
    const LayerParams& L = P.layer[li];
    const int local = tile - L.tile_begin;
    const int ptile = local % L.tiles_per_head;
    const int ph = local / L.tiles_per_head;
    const int head = ph % L.heads, prompt = ph / L.heads;
    const int pixel0 = ptile * kTilePixels;
    const int d = L.head_dim;
    float* ks = smem;                         // [d][80]
    float* qs = smem + d * kTokensPad;        // [128][d + 1]

    const int run = L.tile_begin + ph;        // unique per (layer, prompt, head) within the launch
    const bool load_k = run != last_run;
    last_run = run;

    __syncthreads();                          // previous tile's readers are done
    if (L.dtype == DAAM_F32) stage_tile<float>(L, prompt, head, pixel0, ks, qs, load_k);
    else if (L.dtype == DAAM_F16) stage_tile<__half>(L, prompt, head, pixel0, ks, qs, load_k);
    else stage_tile<__nv_bfloat16>(L, prompt, head, pixel0, ks, qs, load_k);
    __syncthreads();

    float s[kTokensPad];
#pragma unroll
    for (int t = 0; t < kTokensPad; ++t) s[t] = 0.f;
    const float* qrow = qs + threadIdx.x * (d + 1);
#pragma unroll 2
    for (int e = 0; e < d; ++e) {
      const float qv = qrow[e];
      const float4* kr = reinterpret_cast<const float4*>(ks + e * kTokensPad);
#pragma unroll
      for (int j = 0; j < kTokensPad / 4; ++j) {
        const float4 kv = kr[j];
        s[4 * j + 0] = fmaf(qv, kv.x, s[4 * j + 0]);
        s[4 * j + 1] = fmaf(qv, kv.y, s[4 * j + 1]);
        s[4 * j + 2] = fmaf(qv, kv.z, s[4 * j + 2]);
        s[4 * j + 3] = fmaf(qv, kv.w, s[4 * j + 3]);
      }
    }
    // softmax over the 77 real tokens (columns 77..79 are padding and never read)
    float m = s[0];
#pragma unroll
    for (int t = 1; t < kTokens; ++t) m = fmaxf(m, s[t]);
    const float c = L.scale_log2e, mc = m * c;
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < kTokens; ++t) { s[t] = fast_exp2(fmaf(s[t], c, -mc)); sum += s[t]; }
    const float inv = 1.0f / sum;

    const int pixel = pixel0 + threadIdx.x;
    if (pixel < L.hw) {
      float* a = L.acc + ((long long)(prompt * L.heads + head) * kTokens) * L.hw + pixel;
      const long long hw = L.hw;
      if (P.rmw_mode == 1) {
#pragma unroll
        for (int t = 0; t < kTokens; ++t) atomicAdd(a + t * hw, s[t] * inv);   // result unused -> RED
      } else {
        constexpr int kChunk = 11;             // 77 = 7 x 11 loads in flight per thread
#pragma unroll
        for (int t0 = 0; t0 < kTokens; t0 += kChunk) {
          float old[kChunk];
#pragma unroll
          for (int i = 0; i < kChunk; ++i) old[i] = a[(t0 + i) * hw];
#pragma unroll
          for (int i = 0; i < kChunk; ++i) a[(t0 + i) * hw] = fmaf(s[t0 + i], inv, old[i]);
        }
      }
    }
  }
}

}  // namespace

int launch_accumulate_simt(const LaunchParams& p, const DeviceInfo& dev, cudaStream_t stream) {
  int dmax = 0;
  for (int i = 0; i < p.n_layers; ++i) dmax = p.layer[i].head_dim > dmax ? p.layer[i].head_dim : dmax;
  const size_t smem = sizeof(float) * ((size_t)dmax * kTokensPad + (size_t)kTilePixels * (dmax + 1));
  static size_t configured = 0;
  if (smem > configured) {
    DAAM_CUDA_TRY(cudaFuncSetAttribute(accumulate_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    configured = smem;
  }
  int occ = 0;
  DAAM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, accumulate_simt_kernel, kTilePixels, smem));
  if (occ < 1) occ = 1;
  int grid = dev.sm_count * occ;
  if (grid > p.total_tiles) grid = p.total_tiles;
  accumulate_simt_kernel<<<grid, kTilePixels, smem, stream>>>(p);
  DAAM_CUDA_TRY(cudaGetLastError());
  count_launch();
  return DAAM_OK;
}

}  // namespace daam
