// C ABI entry points of libdaam_b200.so (include/daam_b200.h): argument validation, packing of layer calls into
// persistent launches, error strings. The kernels live in accumulate_simt.cu, accumulate_mma.cu and finalize.cu.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace daam {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return DAAM_E_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int get_device_info(DeviceInfo* out) {
  static std::mutex mu;
  static DeviceInfo cache[64];
  int dev = 0;
  DAAM_CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return DAAM_E_CUDA; }
  if (cache[dev].device != dev) {
    cudaDeviceProp prop;
    DAAM_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    cache[dev].device = dev;
    cache[dev].sm_count = prop.multiProcessorCount;
    cache[dev].cc_major = prop.major;
    cache[dev].cc_minor = prop.minor;
  }
  *out = cache[dev];
  return DAAM_OK;
}

static size_t dtype_size(int dtype) { return dtype == DAAM_F32 ? 4 : 2; }

// Validates one layer call and fills the device-side descriptor (tile_begin is set by the packer).
int make_layer_params(const daam_layer& in, int index, LayerParams* out, bool need_acc);
int make_layer_params(const daam_layer& in, int index, LayerParams* out, bool need_acc) {
  if (!in.q || !in.k || (need_acc && !in.acc)) { set_error("daam_accumulate: layer %d has a null pointer", index); return DAAM_E_INVALID; }
  if (in.dtype != DAAM_F32 && in.dtype != DAAM_F16 && in.dtype != DAAM_BF16) { set_error("daam_accumulate: layer %d: unknown dtype %d", index, in.dtype); return DAAM_E_INVALID; }
  if (in.tokens != kTokens) { set_error("daam_accumulate: layer %d: tokens = %d, only %d is traced (daam/trace.py:289)", index, in.tokens, kTokens); return DAAM_E_UNSUPPORTED; }
  if (in.head_dim <= 0 || in.head_dim % 8 != 0 || in.head_dim > DAAM_MAX_HEAD_DIM) { set_error("daam_accumulate: layer %d: head_dim = %d must be a multiple of 8 in (0, %d]", index, in.head_dim, DAAM_MAX_HEAD_DIM); return DAAM_E_UNSUPPORTED; }
  if (in.n_prompts <= 0 || in.heads <= 0 || in.hw <= 0) { set_error("daam_accumulate: layer %d: non-positive n_prompts/heads/hw", index); return DAAM_E_INVALID; }
  if (in.hw % 4 != 0) { set_error("daam_accumulate: layer %d: hw = %d must be a multiple of 4", index, in.hw); return DAAM_E_UNSUPPORTED; }
  if (need_acc && reinterpret_cast<uintptr_t>(in.acc) % 16 != 0) { set_error("daam_accumulate: layer %d: acc is not 16-byte aligned", index); return DAAM_E_INVALID; }
  if (!(in.scale > 0.f)) { set_error("daam_accumulate: layer %d: scale must be positive", index); return DAAM_E_INVALID; }
  LayerParams& L = *out;
  L.q = in.q; L.k = in.k; L.acc = in.acc;
  L.qs_prompt = in.q_stride_prompt; L.qs_pixel = in.q_stride_pixel; L.qs_head = in.q_stride_head;
  L.ks_prompt = in.k_stride_prompt; L.ks_token = in.k_stride_token; L.ks_head = in.k_stride_head;
  L.n_prompts = in.n_prompts; L.heads = in.heads; L.hw = in.hw; L.head_dim = in.head_dim;
  L.dtype = in.dtype;
  L.scale_log2e = in.scale * 1.4426950408889634f;
  L.tiles_per_head = (in.hw + kTilePixels - 1) / kTilePixels;
  L.tile_begin = 0;
  const size_t es = dtype_size(in.dtype);
  auto aligned = [&](long long stride) { return (stride * (long long)es) % 16 == 0; };
  L.vec_ok = reinterpret_cast<uintptr_t>(in.q) % 16 == 0 && reinterpret_cast<uintptr_t>(in.k) % 16 == 0 &&
             aligned(in.q_stride_prompt) && aligned(in.q_stride_pixel) && aligned(in.q_stride_head) &&
             aligned(in.k_stride_prompt) && aligned(in.k_stride_token) && aligned(in.k_stride_head);
  L.pad_ = 0;
  return DAAM_OK;
}

}  // namespace daam

using namespace daam;

extern "C" int daam_accumulate(const daam_layer* layers, int32_t n_layers, uint32_t flags, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_layers < 0 || (n_layers > 0 && !layers)) { set_error("daam_accumulate: bad layer array"); return DAAM_E_INVALID; }
  if (n_layers == 0) return DAAM_OK;
  DeviceInfo dev;
  if (int rc = get_device_info(&dev)) return rc;
  const uint32_t path = flags & 3u, rmw = flags & DAAM_ACC_RMW_MASK;

  // Three packs: 16-bit layers for the tcgen05 kernel (TMA form), fp32 layers for its split form, and the rest for
  // the SIMT kernel. Each is flushed when its parameter block is full.
  static thread_local LaunchParams mma, mma32, simt;
  mma.n_layers = mma32.n_layers = simt.n_layers = 0;
  mma.total_tiles = mma32.total_tiles = simt.total_tiles = 0;
  mma.rmw_mode = mma32.rmw_mode = simt.rmw_mode = (rmw == DAAM_ACC_RMW_LDST) ? 0 : 1;   // default: reduce-add
  mma.pdl = mma32.pdl = simt.pdl = (flags & DAAM_ACC_NO_PDL) ? 0 : 1;
  auto flush = [&](LaunchParams& p, bool is_mma) -> int {
    if (p.n_layers == 0) return DAAM_OK;
    int rc = is_mma ? launch_accumulate_mma(p, dev, stream) : launch_accumulate_simt(p, dev, stream);
    p.n_layers = 0;
    p.total_tiles = 0;
    return rc;
  };
  for (int i = 0; i < n_layers; ++i) {
    LayerParams L;
    if (int rc = make_layer_params(layers[i], i, &L, /*need_acc=*/true)) return rc;
    bool use_mma = path != DAAM_ACC_FORCE_SIMT && mma_supported(L);
    if (path == DAAM_ACC_FORCE_MMA && !use_mma) {
      set_error("daam_accumulate: layer %d cannot take the tcgen05 path (dtype %d, head_dim %d, alignment %d)", i,
                L.dtype, L.head_dim, L.vec_ok);
      return DAAM_E_UNSUPPORTED;
    }
    LaunchParams& p = use_mma ? (L.dtype == DAAM_F32 ? mma32 : mma) : simt;
    L.tile_begin = p.total_tiles;
    p.layer[p.n_layers++] = L;
    p.total_tiles += L.tiles_per_head * L.heads * L.n_prompts;
    if (p.n_layers == kMaxLayersPerLaunch)
      if (int rc = flush(p, use_mma)) return rc;
  }
  if (int rc = flush(mma, true)) return rc;
  if (int rc = flush(mma32, true)) return rc;
  if (int rc = flush(simt, false)) return rc;
  return DAAM_OK;
}

extern "C" int daam_abi_version(void) { return DAAM_ABI_VERSION; }
extern "C" const char* daam_last_error(void) { return g_error; }
extern "C" int64_t daam_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int daam_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  DeviceInfo dev;
  if (int rc = get_device_info(&dev)) return rc;
  if (sm_count) *sm_count = dev.sm_count;
  if (cc_major) *cc_major = dev.cc_major;
  if (cc_minor) *cc_minor = dev.cc_minor;
  return DAAM_OK;
}
