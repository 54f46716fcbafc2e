// Fused softmax(QK^T) -> unravel -> accumulate, tcgen05 / TMA / TMEM variant (head_dim up to 192 in 64-wide K chunks).
//
// One tile = 128 pixels x 77 tokens of one (layer, prompt, head). Per tile:
//   TMA        Q tile [128 x 64] and K [77(+3 zero rows) x 64] -> shared memory, 128B-swizzled K-major (the UMMA
//              canonical layout), straight from the strided `to_q`/`to_k` outputs via 4-D tensor maps
//              {dim, head, row, prompt}; partial tiles and the 3 padding token rows are zero-filled by the TMA unit.
//   tcgen05    S = Q K^T as 4 x tcgen05.mma (M128 N80 K16, kind::f16, fp32 accumulate) into a TMEM accumulator
//              (2 accumulators, so the MMA of tile i+1 overlaps the epilogue of tile i).
//   epilogue   4 warps: tcgen05.ld gives every thread the 77 logits of ITS pixel (TMEM lane == pixel), so the
//              softmax is thread-local (no shuffles); the probabilities are then added into the fp32 accumulator
//              acc[head][token][pixel] either
//                red mode : staged token-major in shared memory and sent as ONE bulk-tensor reduce-add
//                           (cp.reduce.async.bulk.tensor .add.f32): the read-modify-write happens in L2, the SM never
//                           loads the accumulator;
//                ldst mode: coalesced 128-byte load / add / store per warp and token, straight from registers.
// Warp roles: 0-3 epilogue, 4 TMA producer (one elected thread), 5 TMEM allocator + MMA issuer (one thread).
// Persistent: every CTA walks a contiguous chunk of the launch's tiles; up to 2 CTAs per SM (256 TMEM columns each).
//
// fp32 projections (the reference's default dtype for SD-1.x/2.x, daam/run/generate.py:205) take the same kernel in
// "split" form: tensor cores have no fp32 operand type and kind::tf32 would drop 13 mantissa bits, so four converter
// warps load the fp32 Q/K tiles with coalesced 16-byte loads, split every value into three bf16 terms
// (x = x1 + x2 + x3 carries all 24 significand bits), store them as three swizzled operand tiles each, and the MMA
// warp accumulates the six products of order <= 2^-16 (q1k1 + q1k2 + q2k1 + q1k3 + q2k2 + q3k1) in fp32 in TMEM:
// 24 MMAs per tile instead of 4, still far from the tensor pipe's limit, and the path stays HBM-bound.
//
// head_dim other than 64 (SD-1.x: 40 / 80 / 160): the contraction runs in 64-wide K chunks, one chunk per smem stage,
// accumulated into the same TMEM accumulator; the last chunk is zero-filled beyond head_dim (by the TMA unit, or by the
// converter warps) and issues only the MMAs that cover live columns.
//
// Replaces daam/trace.py:276 (get_attention_scores), :219-244 (_unravel_attn) and :293-294 (update loop).
#include <cuda.h>

#include <mutex>
#include <unordered_map>
#include <string>

#include "common.cuh"

namespace daam {
namespace {

constexpr int kStages = 2;
constexpr int kQBytes = kTilePixels * 128;            // 128 rows x 64 x 2 B
constexpr int kKBytes = kTokensPad * 128;             // 80 rows x 64 x 2 B
constexpr int kStageBytes = kQBytes + kKBytes;        // 26624 = 26 x 1024 (keeps every tile 1024-byte aligned)
constexpr int kPBytes = kTokens * kTilePixels * 4;    // staged probabilities [77][128] fp32
constexpr int kTmemCols = 256;
constexpr int kAccCols = 128;                         // column distance between the two accumulators
constexpr int kThreads = 192;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kPBytes + 128;
// split (fp32) form: a stage holds Q1 Q2 Q3 K1 K2 K3; warps 6-9 convert; one CTA per SM
constexpr int kSplitStageBytes = 3 * kStageBytes;
constexpr int kSplitThreads = 448;                 // + two converter groups of 4 warps, one per stage
constexpr int kSplitSmemBytes = 1024 + kStages * kSplitStageBytes + kPBytes + 128;

struct MmaParams {
  LaunchParams base;
  CUtensorMap qmap[kMaxLayersPerLaunch];
  CUtensorMap kmap[kMaxLayersPerLaunch];
  CUtensorMap amap[kMaxLayersPerLaunch];
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) __trap();    // ~10 s: far beyond any legitimate wait, preemption included
  }
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major operand tile, 128B swizzle, rows of 128 bytes, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) /* LBO (unused with swizzle) */ |
         ((uint64_t)(1024 >> 4) << 32) /* SBO */ | (1ull << 46) /* descriptor version (sm_100) */ |
         (2ull << 61) /* SWIZZLE_128B */;
}
// Instruction descriptor, kind::f16: fp32 accumulate, A/B both K-major, M = 128, N = 80.
__device__ __forceinline__ uint32_t umma_idesc(bool bf16) {
  const uint32_t fmt = bf16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kTokensPad >> 3) << 17) |
         ((uint32_t)(kTilePixels >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct Tile {
  int li, prompt, head, pixel0;
};
__device__ __forceinline__ Tile decode_tile(const LaunchParams& P, int tile, int& li) {
  while (li + 1 < P.n_layers && tile >= P.layer[li + 1].tile_begin) ++li;
  const LayerParams& L = P.layer[li];
  const int local = tile - L.tile_begin;
  const int ptile = local % L.tiles_per_head;
  const int ph = local / L.tiles_per_head;
  Tile t;
  t.li = li;
  t.head = ph % L.heads;
  t.prompt = ph / L.heads;
  t.pixel0 = ptile * kTilePixels;
  return t;
}

// fp32 -> three bf16 terms, 8 values -> one 16-byte chunk per term
__device__ __forceinline__ void split8(const float (&x)[8], uint4& p1, uint4& p2, uint4& p3) {
  __nv_bfloat162 a[4], b[4], c[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float lo = x[2 * i], hi = x[2 * i + 1];
    const __nv_bfloat16 a0 = __float2bfloat16_rn(lo), a1 = __float2bfloat16_rn(hi);
    lo -= __bfloat162float(a0); hi -= __bfloat162float(a1);
    const __nv_bfloat16 b0 = __float2bfloat16_rn(lo), b1 = __float2bfloat16_rn(hi);
    lo -= __bfloat162float(b0); hi -= __bfloat162float(b1);
    a[i] = __halves2bfloat162(a0, a1);
    b[i] = __halves2bfloat162(b0, b1);
    c[i] = __halves2bfloat162(__float2bfloat16_rn(lo), __float2bfloat16_rn(hi));
  }
  p1 = *reinterpret_cast<uint4*>(a);
  p2 = *reinterpret_cast<uint4*>(b);
  p3 = *reinterpret_cast<uint4*>(c);
}

// Converter warps (split form): rows [row0, row0 + n_rows_live) of a fp32 [rows x 64] operand -> three 128B-swizzled
// K-major bf16 tiles at dst, dst + part_bytes, dst + 2 * part_bytes. 16-byte chunks of 8 floats, chunk = row * 8 +
// group; rows >= n_rows_live and columns >= n_cols_live (a multiple of 8) are written as zeros. ctid: 0..127.
template <int kChunksPerThread>
__device__ __forceinline__ void convert_operand(const float* src_base, long long row_stride, int n_rows_live,
                                                int n_cols_live, uint8_t* dst, int part_bytes, int ctid) {
  float4 lo[kChunksPerThread], hi[kChunksPerThread];
#pragma unroll
  for (int c = 0; c < kChunksPerThread; ++c) {            // all loads first: 2 x kChunksPerThread in flight per thread
    const int chunk = ctid + 128 * c, r = chunk >> 3, g = chunk & 7;
    if (r < n_rows_live && g * 8 < n_cols_live) {
      const float4* srcp = reinterpret_cast<const float4*>(src_base + (long long)r * row_stride + g * 8);
      lo[c] = __ldg(srcp);
      hi[c] = __ldg(srcp + 1);
    } else {
      lo[c] = hi[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int c = 0; c < kChunksPerThread; ++c) {
    const int chunk = ctid + 128 * c, r = chunk >> 3, g = chunk & 7;
    const float x[8] = {lo[c].x, lo[c].y, lo[c].z, lo[c].w, hi[c].x, hi[c].y, hi[c].z, hi[c].w};
    uint4 p1, p2, p3;
    split8(x, p1, p2, p3);
    const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((g ^ (r & 7)) << 4);     // SWIZZLE_128B, K-major
    *reinterpret_cast<uint4*>(dst + off) = p1;
    *reinterpret_cast<uint4*>(dst + part_bytes + off) = p2;
    *reinterpret_cast<uint4*>(dst + 2 * part_bytes + off) = p3;
  }
}

// kChunked: some layer of the launch has head_dim > 64 (several K chunks per tile); the common single-chunk case keeps
// its simpler loops (one load iteration per tile).
template <bool kSplit, bool kChunked>
__global__ void __launch_bounds__(kSplit ? kSplitThreads : kThreads, kSplit ? 1 : 2)
accumulate_mma_kernel(const __grid_constant__ MmaParams MP) {
  constexpr int kStageBytesT = kSplit ? kSplitStageBytes : kStageBytes;
  const LaunchParams& P = MP.base;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                       // 1024-byte alignment for the swizzled tiles
  uint8_t* gen = smem_raw + (base - raw);
  float* sP = reinterpret_cast<float*>(gen + kStages * kStageBytesT);
  const uint32_t sP_u32 = base + kStages * kStageBytesT;
  const uint32_t bars = sP_u32 + kPBytes;                             // 8 mbarriers + the TMEM base address
  const uint32_t full0 = bars, empty0 = bars + 16, tfull0 = bars + 32, tempty0 = bars + 48;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + kStages * kStageBytesT + kPBytes + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = P.total_tiles / gridDim.x, rem = P.total_tiles % gridDim.x;
  const int first = blockIdx.x * per + min((int)blockIdx.x, rem);
  const int count = per + ((int)blockIdx.x < rem ? 1 : 0);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, kSplit ? 128 : 1);   // split form: one arrival per converter thread
      mbar_init(empty0 + 8 * s, 1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4);       // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // descriptor fetches of the first tile overlap the barrier / TMEM set-up (and, under PDL, the previous kernel's tail)
  // (not in the split form: it sits at the 128-register cap and any extra live state costs it ~15 %)
  if constexpr (!kSplit) {
    if (count > 0 && lane == 0 && (warp == 0 || warp == 4)) {
      int li0 = 0;
      const Tile t0 = decode_tile(P, first, li0);
      if (warp == 0) {
        prefetch_tensormap(&MP.amap[t0.li]);
      } else {
        prefetch_tensormap(&MP.qmap[t0.li]);
        prefetch_tensormap(&MP.kmap[t0.li]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation) may overlap the tail of the
  // previous kernel on the stream; nothing below (TMA loads, reduce-adds) may start before that kernel has completed
  // and flushed. Our own dependents may be scheduled as soon as every CTA of this grid is past this point.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (kSplit && warp >= 6) {
    // ===== converter warps (fp32 projections): global fp32 -> three bf16 operand tiles per operand =====
    // group g (warps 6-9 / 10-13) fills stage g with the load iterations j = g, g + 2, ... (one iteration = one 64-wide
    // K chunk of one tile): two iterations' loads and conversions overlap
    const int group = (warp - 6) >> 2;
    const int ctid = (threadIdx.x - 192) & 127;
    auto convert_chunk = [&](const Tile& t, const LayerParams& L, int c, uint32_t ph) {
      mbar_wait(empty0 + 8 * group, ph ^ 1u);
      uint8_t* stage = gen + group * kStageBytesT;
      const int cols = min(64, L.head_dim - 64 * c);
      const float* qsrc = static_cast<const float*>(L.q) + t.prompt * L.qs_prompt + t.head * L.qs_head +
                          (long long)t.pixel0 * L.qs_pixel + 64 * c;
      const float* ksrc = static_cast<const float*>(L.k) + t.prompt * L.ks_prompt + t.head * L.ks_head + 64 * c;
      convert_operand<8>(qsrc, L.qs_pixel, min(kTilePixels, L.hw - t.pixel0), cols, stage, kQBytes, ctid);
      convert_operand<5>(ksrc, L.ks_token, kTokens, cols, stage + 3 * kQBytes, kKBytes, ctid);
      fence_proxy_async();                             // generic-proxy stores -> visible to the tensor core's reads
      mbar_arrive(full0 + 8 * group);
    };
    int li = 0;
    if constexpr (!kChunked) {
      for (int i = group; i < count; i += kStages) {   // load iteration == tile
        const Tile t = decode_tile(P, first + i, li);
        convert_chunk(t, P.layer[t.li], 0, (uint32_t)(i >> 1) & 1u);
      }
    } else {
      int j = 0;
      for (int i = 0; i < count; ++i) {
        const Tile t = decode_tile(P, first + i, li);
        const LayerParams& L = P.layer[t.li];
        const int n_chunks = (L.head_dim + 63) >> 6;
        for (int c = 0; c < n_chunks; ++c, ++j)
          if ((j & 1) == group) convert_chunk(t, L, c, (uint32_t)(j >> 1) & 1u);
      }
    }
  } else if (warp == 4) {
    // ===== TMA producer (16-bit projections) =====
    if (!kSplit && lane == 0) {
      int li = 0, j = 0;
      for (int i = 0; i < count; ++i) {
        const Tile t = decode_tile(P, first + i, li);
        const int n_chunks = kChunked ? (P.layer[t.li].head_dim + 63) >> 6 : 1;
        for (int c = 0; c < n_chunks; ++c, ++j) {      // one load iteration = one 64-wide K chunk of one tile
          const int s = j % kStages;
          const uint32_t ph = (uint32_t)(j / kStages) & 1u;
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          mbar_expect_tx(full0 + 8 * s, kStageBytes);
          const uint32_t q_dst = base + s * kStageBytesT, k_dst = q_dst + kQBytes;
          tma_load_4d(&MP.qmap[t.li], full0 + 8 * s, q_dst, 64 * c, t.head, t.pixel0, t.prompt);
          tma_load_4d(&MP.kmap[t.li], full0 + 8 * s, k_dst, 64 * c, t.head, 0, t.prompt);
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int li = 0, j = 0;
      for (int i = 0; i < count; ++i) {
        const Tile t = decode_tile(P, first + i, li);
        const LayerParams& L = P.layer[t.li];
        const int a = i & 1;
        const uint32_t aph = (uint32_t)(i >> 1) & 1u;
        const int n_chunks = kChunked ? (L.head_dim + 63) >> 6 : 1;
        const uint32_t d_tmem = tmem_base + a * kAccCols;
        mbar_wait(tempty0 + 8 * a, aph ^ 1u);          // epilogue has drained this accumulator
        for (int c = 0; c < n_chunks; ++c, ++j) {
          const int s = j % kStages;
          const uint32_t ph = (uint32_t)(j / kStages) & 1u;
          mbar_wait(full0 + 8 * s, ph);                // the chunk's operand tiles have landed
          tc_fence_after();
          const uint32_t q_src = base + s * kStageBytesT;
          const int k_steps = (min(64, L.head_dim - 64 * c) + 15) >> 4;   // UMMA_K 16 = 32 bytes along the swizzled row
          if constexpr (kSplit) {
            // q.k = sum of the six split products up to order 2^-16, smallest first; every operand is bf16
            const uint32_t k_src = q_src + 3 * kQBytes;
            const uint32_t idesc = umma_idesc(true);
            constexpr int qa[6] = {2, 0, 1, 1, 0, 0}, kb[6] = {0, 2, 1, 0, 1, 0};
#pragma unroll
            for (int p = 0; p < 6; ++p)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < k_steps)
                  umma_f16(d_tmem, umma_desc_sw128(q_src + qa[p] * kQBytes + 32 * k),
                           umma_desc_sw128(k_src + kb[p] * kKBytes + 32 * k), idesc, (c | p | k) != 0);
          } else {
            const uint32_t k_src = q_src + kQBytes;
            const uint32_t idesc = umma_idesc(L.dtype == DAAM_BF16);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < k_steps)
                umma_f16(d_tmem, umma_desc_sw128(q_src + 32 * k), umma_desc_sw128(k_src + 32 * k), idesc, (c | k) != 0);
          }
          umma_commit(empty0 + 8 * s);                 // frees the smem stage once the MMAs have read it
        }
        umma_commit(tfull0 + 8 * a);                   // accumulator ready for the epilogue
      }
    }
  } else if (warp < 4) {
    // ===== epilogue warps: softmax + accumulate =====
    int li = 0;
    const int tid = threadIdx.x;                       // 0..127 == pixel within the tile == TMEM lane
    bool issued = false;
    for (int i = 0; i < count; ++i) {
      const Tile t = decode_tile(P, first + i, li);
      const LayerParams& L = P.layer[t.li];
      const int a = i & 1;
      const uint32_t aph = (uint32_t)(i >> 1) & 1u;
      mbar_wait(tfull0 + 8 * a, aph);
      tc_fence_after();
      float v[kTokensPad];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + a * kAccCols;
#pragma unroll
      for (int c = 0; c < kTokensPad / 16; ++c) tmem_ld16(taddr + c * 16, v + c * 16);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * a);

      float m = v[0];
#pragma unroll
      for (int j = 1; j < kTokens; ++j) m = fmaxf(m, v[j]);
      const float c = L.scale_log2e, mc = m * c;
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < kTokens; ++j) {
        v[j] = fast_exp2(fmaf(v[j], c, -mc));
        sum += v[j];
      }
      const float inv = 1.0f / sum;

      if (P.rmw_mode == 1) {
        if (tid == 0 && issued) bulk_wait_read0();     // the previous reduce has finished reading sP
        epi_barrier();
#pragma unroll
        for (int j = 0; j < kTokens; ++j) sP[j * kTilePixels + tid] = v[j] * inv;
        fence_proxy_async();                           // generic-proxy writes -> visible to the bulk-async proxy
        epi_barrier();
        if (tid == 0) {
          tma_reduce_add_2d(&MP.amap[t.li], sP_u32, t.pixel0, (t.prompt * L.heads + t.head) * kTokens);
          bulk_commit();
        }
        issued = true;
      } else {
        const int pixel = t.pixel0 + tid;
        if (pixel < L.hw) {
          const long long hw = L.hw;
          float* acc = L.acc + ((long long)(t.prompt * L.heads + t.head) * kTokens) * hw + pixel;
          constexpr int kChunk = 11;
#pragma unroll
          for (int j0 = 0; j0 < kTokens; j0 += kChunk) {
            float old[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) old[j] = acc[(j0 + j) * hw];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) acc[(j0 + j) * hw] = fmaf(v[j0 + j], inv, old[j]);
          }
        }
      }
    }
    if (tid == 0 && issued) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---- host: tensor maps ----------------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  long long s1, s2, s3;
  int d1, d2, d3, kind;     // kind: 0 q/k 16-bit (dtype in bit 4), 1 accumulator
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && s1 == o.s1 && s2 == o.s2 && s3 == o.s3 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 &&
           kind == o.kind;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](long long v) { h ^= (size_t)v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.s1); mix(k.s2); mix(k.s3); mix(k.d1); mix(k.d2); mix(k.d3); mix(k.kind);
    return h;
  }
};

// Tensor maps are pure functions of (pointer, shape, strides): cache them, the allocator hands the same Q/K
// addresses back every denoising step.
std::unordered_map<MapKey, CUtensorMap, MapKeyHash>& map_cache() {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> c;
  return c;
}
std::mutex g_map_mu;

// {head_dim, heads, rows, prompts} view of a projection; box = [box_rows x 64 dims] of one head, 128B-swizzled (columns
// beyond head_dim in the last K chunk are zero-filled).
int make_qk_map(const void* ptr, int dtype, int head_dim, int heads, int rows, int prompts, long long s_head,
                long long s_row, long long s_prompt, int box_rows, CUtensorMap* out) {
  MapKey key{ptr, s_head, s_row, s_prompt, heads, rows, prompts * 1024 + box_rows, (dtype << 4) | (head_dim << 8)};
  {
    std::lock_guard<std::mutex> lock(g_map_mu);
    auto it = map_cache().find(key);
    if (it != map_cache().end()) { *out = it->second; return DAAM_OK; }
  }
  EncodeFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return DAAM_E_CUDA; }
  const cuuint64_t dims[4] = {(cuuint64_t)head_dim, (cuuint64_t)heads, (cuuint64_t)rows, (cuuint64_t)prompts};
  auto bytes = [](long long s) { return (cuuint64_t)(s > 0 ? s : 8) * 2; };
  const cuuint64_t strides[3] = {bytes(s_head), bytes(s_row), bytes(s_prompt)};
  const cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, dtype == DAAM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(q/k) failed with CUresult %d", (int)r); return DAAM_E_CUDA; }
  std::lock_guard<std::mutex> lock(g_map_mu);
  if (map_cache().size() > 8192) map_cache().clear();
  map_cache()[key] = *out;
  return DAAM_OK;
}

// accumulator as a 2-D fp32 tensor {hw, prompts*heads*77}; box = [77 tokens x 128 pixels], no swizzle.
int make_acc_map(float* acc, int hw, int rows, CUtensorMap* out) {
  MapKey key{acc, 0, 0, 0, hw, rows, 0, 1};
  {
    std::lock_guard<std::mutex> lock(g_map_mu);
    auto it = map_cache().find(key);
    if (it != map_cache().end()) { *out = it->second; return DAAM_OK; }
  }
  EncodeFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return DAAM_E_CUDA; }
  const cuuint64_t dims[2] = {(cuuint64_t)hw, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)hw * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kTilePixels, (cuuint32_t)kTokens};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, acc, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(acc) failed with CUresult %d", (int)r); return DAAM_E_CUDA; }
  std::lock_guard<std::mutex> lock(g_map_mu);
  if (map_cache().size() > 8192) map_cache().clear();
  map_cache()[key] = *out;
  return DAAM_OK;
}

}  // namespace

bool mma_supported(const LayerParams& L) {
  return L.head_dim % 8 == 0 && L.head_dim <= 192 && L.vec_ok && L.qs_head > 0 && L.qs_pixel > 0 && L.ks_head > 0 &&
         L.ks_token > 0;
}

int launch_accumulate_mma(const LaunchParams& p, const DeviceInfo& dev, cudaStream_t stream) {
  if (dev.cc_major != 10) { set_error("the tcgen05 kernel needs an sm_100 device (found sm_%d%d)", dev.cc_major, dev.cc_minor); return DAAM_E_UNSUPPORTED; }
  static thread_local MmaParams mp;
  mp.base = p;
  const bool split = p.n_layers > 0 && p.layer[0].dtype == DAAM_F32;     // a pack holds one operand class (api.cu)
  for (int i = 0; i < p.n_layers; ++i) {
    const LayerParams& L = p.layer[i];
    if ((L.dtype == DAAM_F32) != split) { set_error("mixed fp32 / 16-bit layers in one tcgen05 pack"); return DAAM_E_INVALID; }
    if (!split) {
      if (int rc = make_qk_map(L.q, L.dtype, L.head_dim, L.heads, L.hw, L.n_prompts, L.qs_head, L.qs_pixel, L.qs_prompt, kTilePixels, &mp.qmap[i])) return rc;
      if (int rc = make_qk_map(L.k, L.dtype, L.head_dim, L.heads, kTokens, L.n_prompts, L.ks_head, L.ks_token, L.ks_prompt, kTokensPad, &mp.kmap[i])) return rc;
    }
    if (int rc = make_acc_map(L.acc, L.hw, L.n_prompts * L.heads * kTokens, &mp.amap[i])) return rc;
  }
  static bool configured_dev[64] = {};                // the attribute is per device
  bool& configured = configured_dev[dev.device & 63];
  if (!configured) {
    DAAM_CUDA_TRY(cudaFuncSetAttribute(accumulate_mma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    DAAM_CUDA_TRY(cudaFuncSetAttribute(accumulate_mma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    DAAM_CUDA_TRY(cudaFuncSetAttribute(accumulate_mma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSplitSmemBytes));
    DAAM_CUDA_TRY(cudaFuncSetAttribute(accumulate_mma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSplitSmemBytes));
    configured = true;
  }
  int grid = dev.sm_count * (split ? 1 : 2);
  if (grid > p.total_tiles) grid = p.total_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(split ? kSplitThreads : kThreads);
  cfg.dynamicSmemBytes = split ? kSplitSmemBytes : kSmemBytes;
  cfg.stream = stream;
  // Inside a stream capture the launch becomes a plain kernel node (programmatic edges are left to the graph owner).
  cudaStreamCaptureStatus capture = cudaStreamCaptureStatusNone;
  DAAM_CUDA_TRY(cudaStreamIsCapturing(stream, &capture));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (p.pdl && capture == cudaStreamCaptureStatusNone) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  bool chunked = false;
  for (int i = 0; i < p.n_layers; ++i) chunked = chunked || p.layer[i].head_dim > 64;
  if (split && chunked) DAAM_CUDA_TRY(cudaLaunchKernelEx(&cfg, accumulate_mma_kernel<true, true>, mp));
  else if (split) DAAM_CUDA_TRY(cudaLaunchKernelEx(&cfg, accumulate_mma_kernel<true, false>, mp));
  else if (chunked) DAAM_CUDA_TRY(cudaLaunchKernelEx(&cfg, accumulate_mma_kernel<false, true>, mp));
  else DAAM_CUDA_TRY(cudaLaunchKernelEx(&cfg, accumulate_mma_kernel<false, false>, mp));
  DAAM_CUDA_TRY(cudaGetLastError());
  count_launch();
  return DAAM_OK;
}

}  // namespace daam
