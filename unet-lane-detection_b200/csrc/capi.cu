// C ABI of libunet_b200.so (see include/unet_b200.h). Host-side plan: layer table, workspace layout,
// TMA tensor maps, launches. No torch types, no exceptions across the boundary, no CPU fallback.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "../../include/unet_b200.h"
#include "aux_kernels.cuh"
#include "conv_halo.cuh"
#include "conv_umma.cuh"
#include "stem_umma.cuh"
#include "stem_halo.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define UB_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail(UB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                        \
  } while (0)

// ---- driver entry point for tensor-map encoding (no link-time dependency on libcuda) ------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// ---- kernel-selection switches (unet_b200_set_option) ------------------------------------------------------------------
// g_opts holds the PROCESS DEFAULTS. A plan / trainer copies them when it is created and runs with its own copy from then on
// (its entry points install the copy as this thread's current options, OptScope), so changing a default never changes the
// behaviour of an existing plan, and two plans with different options can run side by side. The single-layer entry points
// have no plan and read the defaults at call time.
struct Opts {
  int halo = 1;          // 3x3 convs with Cout 64/128 on the halo-patch kernel
  int fuse_head = 1;     // 1x1 head + sigmoid + mask inside the last conv's epilogue
  int stem_umma = 1;     // Cout == 64 stem on tensor cores (stem_umma.cuh) instead of the FP32-pipe kernel
  int halo2 = 1;         // halo layers on the CTA-pair kernel when at least two tiles exist
  int umma2 = 1;         // per-tap layers on the CTA-pair kernel when at least two pixel tiles exist
  int pdl = -1;          // programmatic dependent launch: -1 = per plan (on for small plans, see plan_create), 0 / 1 = off / on
  int wgrad_rows64 = 1;  // 64-pixel reduction tiles in the BLOCK_N = 256 weight-gradient kernel
  int wgrad2 = 1;        // CTA-pair weight-gradient kernel for BLOCK_N >= 128
  int wgrad_stream = 1;  // weight-gradient GEMMs on a side stream
  int stem_wide = 0;     // tensor-core stem on 4 x 32 tiles (4 KB contiguous output rows per store) instead of 16 x 8
  int bwd_fuse = 2;      // training: BatchNorm-backward reduction fused into the elementwise pass that produces the gradient:
                         // 2 = the head backward only (16.9 vs 17.0-17.2 ms/step), 1 = also the four max-pool backward passes
                         // (slower: 140 registers), 0 = off
  int dgrad_fuse = 1;    // training: BatchNorm-backward reduction inside the tcgen05 dgrad epilogues that write the gradient
  int wgrad_halo = 1;    // training: Cout == 64 weight gradients on the halo-patch kernel (all nine taps per CTA)
  int pack_split = 0;    // training: bulk of the operand pack on the side stream (measured: the block scheduler runs it first anyway)
  int host_pieces = 8;   // host-buffer entry point: pieces per pass whose copies are pipelined with the first / last layers (0: off)
  int stem_fuse = 0;     // inference: the stem runs inside enc0.conv1's patch producer (stem_halo2_kernel), its output never stored
                         // (bit-identical; measured the same speed as the two kernels - shared-memory bound - so off)
  int small_n = 1;       // small plans (per-frame executor, batches up to ~32): narrower column blocks (BLOCK_N 128 / 64) where the
                         // wave model says they are faster; 2 = only while a layer cannot give every SM a tile; 0 = always 256
  int host_geometric = 1; // host-buffer entry point: pieces of 16, 32, 64, ... frames (largest first at the output end) instead of equal ones
  int host_hybrid = 1;   // host-buffer entry point, source frames much larger than the network input (copy-bound): short first pass
                         // AND pieces inside every pass (0: pass-granular pipeline without pieces, the earlier form)
  int pre_rows = 0;      // resizing preprocess, bulk form: output rows per tile (0: 8, halved until the stages fit)
  int pre_stages = 0;    // resizing preprocess, bulk form: shared-memory stages (0: 2)
  int pre_bulk = 1;      // resizing preprocess: source rows staged by the copy engine (cp.async.bulk, two stages) instead of by the threads
};
Opts g_opts;
thread_local const Opts* tl_opts = &g_opts;
struct OptScope {
  const Opts* prev;
  explicit OptScope(const Opts* o) : prev(tl_opts) { tl_opts = o; }
  ~OptScope() { tl_opts = prev; }
};

// ---- per-device state -----------------------------------------------------------------------------------------------
// Everything the library remembers about a device lives here, indexed by the CUDA device ordinal that is current when an
// entry point runs: SM count, which kernels already had their dynamic-shared-memory limit raised (cudaFuncSetAttribute is
// per device), and a small all-zero bias vector for the single-op entry points. One process may drive several devices.
constexpr int UB_MAX_DEVICES = 64;
enum AttrSlot : int {
  AT_UMMA = 0,        // + {0,1,2} for BLOCK_N 64/128/256
  AT_UMMA2 = 3,
  AT_HALO = 6,        // + {0,1} for BLOCK_N 64/128
  AT_HALO2 = 8,
  AT_STEM = 10,
  AT_PRE = 11,
  AT_WGRAD = 12,      // + {0,1,2}
  AT_WGRAD2 = 15,     // + {1,2}
  AT_STEM_WGRAD_TC = 18,
  AT_STEM_WGRAD = 19,
  AT_PRE_ROWS = 20,
  AT_BNFUSE = 21,
  AT_STEM_WIDE = 22,
  AT_WGRAD_HALO = 23,
  AT_STEM_HALO = 24,
  AT_PRE_BULK = 25,   // + {0..7}: {2x2 decimation, swap, generic outputs} instantiations
};
struct DevState {
  std::atomic<int> num_sms{0};
  std::atomic<unsigned long long> attrs{0};
  std::atomic<float*> zero_bias{nullptr};
};
DevState g_dev[UB_MAX_DEVICES];

DevState* cur_dev() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= UB_MAX_DEVICES) dev = 0;
  return &g_dev[dev];
}
// SM count of the current device (device_check() has run for it: every entry point that launches calls it first)
int cur_sms() {
  const int n = cur_dev()->num_sms.load(std::memory_order_relaxed);
  return n > 0 ? n : 148;
}
// Raise a kernel's dynamic shared-memory limit once per device.
template <typename K>
cudaError_t ensure_smem(K kernel, int slot, int bytes) {
  DevState* d = cur_dev();
  const unsigned long long bit = 1ull << slot;
  if (d->attrs.load(std::memory_order_acquire) & bit) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) d->attrs.fetch_or(bit, std::memory_order_release);
  return e;
}

// Every kernel launch of the library. With the "pdl" option a kernel may be scheduled before its
// stream predecessor has drained (programmatic dependent launch; every kernel executes griddepcontrol.wait before its first
// global access - pdl_enter() / pdl_wait(), ptx.cuh - so the ordering of the data is unchanged). For LARGE batches it is
// slower (tools/ab_option.py pdl, same box, alternating: inference 17.0 k -> 16.4 k frames/s, training 18.9 -> 19.3 ms/step -
// the persistent one-CTA-per-SM kernels leave no room for an early dependent, and its parked CTAs only get in the way of the
// tail); for small plans it hides the kernels' prologues: the default (-1) lets plan_create decide.
template <typename... KArgs, typename... Args>
void ub_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = tl_opts->pdl > 0 ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);   // errors surface through cudaGetLastError() at the call site
}

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

// bf16 NHWC activation [Bc][H][W][C] -> 4-D map (C, W, H, B), box (64, TW, TH, TB), 128B swizzle, zero OOB fill.
int make_act_map(CUtensorMap* m, const void* base, int Bc, int H, int W, int C, int TW, int TH, int TB) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail(UB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bc};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  // a box may not exceed the tensor extent along the batch axis; the kernel is told the smaller byte count
  cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)(TB < Bc ? TB : Bc)};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(UB_ERR_CUDA, "cuTensorMapEncodeTiled(act B=%d H=%d W=%d C=%d box %dx%dx%d) -> %d", Bc, H, W, C, TW,
                TH, TB, (int)r);
  }
  return UB_OK;
}

// bf16 weights [N][K] K-major -> 2-D map (K, N), box (64, box_rows).
int make_w_map_box(CUtensorMap* m, const void* base, int N, int K, int box_rows);
// Weight map of the conv kernels: the box is HALF a BLOCK_N tile, so that the same map serves the 1-CTA kernels (two loads
// per tile) and the CTA-pair kernels (each CTA of the pair loads its half), see conv_umma.cuh / conv_halo.cuh.
int make_w_map(CUtensorMap* m, const void* base, int N, int K, int block_n) { return make_w_map_box(m, base, N, K, block_n / 2); }
int make_w_map_box(CUtensorMap* m, const void* base, int N, int K, int block_n) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail(UB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)block_n};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(UB_ERR_CUDA, "cuTensorMapEncodeTiled(w N=%d K=%d) -> %d", N, K, (int)r);
  return UB_OK;
}

// Halo'd patch map for conv_halo.cuh: same tensor, box (64 ch, 10 px, 18 rows, 1 image).
int make_halo_map(CUtensorMap* m, const void* base, int Bc, int H, int W, int C) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail(UB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bc};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, 10, 18, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(UB_ERR_CUDA, "cuTensorMapEncodeTiled(halo B=%d H=%d W=%d C=%d) -> %d", Bc, H, W, C, (int)r);
  return UB_OK;
}

// NHWC bf16 [Bc][H][W][C] viewed as (C, W, H, B) with a (64, bw, bh, 1) box: TMA-store target of the staged epilogues.
int make_box_map(CUtensorMap* m, const void* base, int Bc, int H, int W, int C, int bw, int bh) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail(UB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bc};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(UB_ERR_CUDA, "cuTensorMapEncodeTiled(box B=%d H=%d W=%d C=%d %dx%d) -> %d", Bc, H, W, C, bw, bh, (int)r);
  return UB_OK;
}

// Tile-box (64, TW, TH, TB) store view of an NHWC bf16 tensor with arbitrary pixel strides (in elements):
// used for conv outputs, pooled outputs and the four (dy,dx) quads of a ConvT output.
int make_tile_store_map(CUtensorMap* m, const void* base, int Bc, int H, int W, int C, size_t sw, size_t sh, size_t sb,
                        int TW, int TH, int TB) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail(UB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bc};
  cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)(TB < Bc ? TB : Bc)};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(UB_ERR_CUDA, "cuTensorMapEncodeTiled(store B=%d H=%d W=%d C=%d) -> %d", Bc, H, W, C, (int)r);
  return UB_OK;
}

struct Layer;
int make_umma_store_maps(Layer& l, void* out, void* pool, int Bc, int cm = 1);

// Shared-memory carve-up of conv_halo_kernel: prefer resident weights, then the deepest patch ring that fits.
bool halo_smem_plan(int block_n, int kc, int head, ub::HaloArgs* a) {
  const int ystg = a->bn.y != nullptr;   // fused BatchNorm-backward sums: a second set of warp-private tiles
  const int need = 3 * kc;  // weight stages (3 taps each)
  if (need <= ub::HaloCfg::MAX_B) {
    for (int as = 4; as >= 2; --as) {
      if (ub::halo_smem_bytes(block_n, as, need, head, 0, ystg) <= ub::HaloCfg::SMEM_LIMIT) {
        a->resident = 1; a->a_stages = as; a->b_stages = need;
        return true;
      }
    }
  }
  // streamed weights: a full kernel (3 stages) in flight first, then as many patch stages as fit
  for (int b = 3; b >= 2; --b) {
    for (int as = 3; as >= 2; --as) {
      if (ub::halo_smem_bytes(block_n, as, b, head, 0, ystg) <= ub::HaloCfg::SMEM_LIMIT) {
        a->resident = 0; a->a_stages = as; a->b_stages = b;
        return true;
      }
    }
  }
  return false;
}

bool halo_eligible(int H, int W, int C0, int C1, int Cout) {
  (void)H;
  return tl_opts->halo && (W % 8 == 0) && (Cout == 64 || Cout == 128) && (C0 % 64 == 0) && (C1 % 64 == 0);
}

int pow2_divisor(int v, int cap) {
  int p = 1;
  while (p * 2 <= cap && v % (p * 2) == 0) p *= 2;
  return p;
}

// 128-pixel box: TW | W (<= 16 so the fused pool's partners stay inside a warp), TH | H, rest from batch.
// Small batches (the per-frame path of the ROS node): when the batch cannot fill the box's image dimension, the box is made
// as large as the image allows instead - TW / TH = the largest powers of two not exceeding W / H (and 16, 128/TW), whether or
// not they divide the image: TMA zero-fills the overhang on loads (which is also the conv's padding) and clips it on stores.
// A 14x14 level at batch 1 is then 4 tiles per column block instead of 49 tiles with 4 live rows each.
void pick_tile(int H, int W, int* TW, int* TH, int* TB, int batch = 1 << 30) {
  *TW = pow2_divisor(W, 16);
  *TH = pow2_divisor(H, 128 / *TW);
  *TB = 128 / (*TW * *TH);
  if (*TB > batch && *TB > 1) {
    int tw = 1, th = 1;
    while (tw * 2 <= W && tw * 2 <= 16) tw *= 2;
    while (th * 2 <= H && th * 2 <= 128 / tw) th *= 2;
    if (tw * th > *TW * *TH) {   // only when it actually puts more pixels of one image into the box
      *TW = tw;
      *TH = th;
      *TB = 128 / (tw * th);
    }
  }
}

int pick_block_n(int N) { return (N % 256 == 0) ? 256 : (N % 128 == 0) ? 128 : 64; }

int device_check() {
  int dev = 0;
  UB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= UB_MAX_DEVICES) return fail(UB_ERR_DEVICE, "device ordinal %d outside [0,%d)", dev, UB_MAX_DEVICES);
  DevState* d = &g_dev[dev];
  if (d->num_sms.load(std::memory_order_acquire) > 0) return UB_OK;  // cudaGetDeviceProperties is slow: once per device
  cudaDeviceProp prop;
  UB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    return fail(UB_ERR_DEVICE, "device %d '%s' is sm_%d%d; libunet_b200 runs on sm_100 only (no fallback)", dev, prop.name,
                prop.major, prop.minor);
  }
  d->num_sms.store(prop.multiProcessorCount, std::memory_order_release);
  return UB_OK;
}


template <int BN>
int launch_conv_t(const CUtensorMap* ma, const CUtensorMap& w, const CUtensorMap* mo, const ub::ConvArgs& args, int slot,
                  cudaStream_t st, const CUtensorMap* my) {
  const int m_tiles = args.tiles_w * args.tiles_h * args.tiles_b;
  const int sms = cur_sms();
  const bool ystg = args.bn.y != nullptr;     // fused BatchNorm-backward sums need the y map and a second set of staging tiles
  if (ystg && (my == nullptr || args.epi != ub::EPI_STORE || args.split)) return fail(UB_ERR_ARG, "fused BN-backward sums: bad layer");
  const CUtensorMap& tmy = my != nullptr ? *my : mo[0];
  if (tl_opts->umma2 && m_tiles >= 2) {
    using Cfg2 = ub::ConvCfg<BN, true>;
    UB_CUDA(ensure_smem(ub::conv_umma2_kernel<BN>, AT_UMMA2 + slot, Cfg2::SMEM_LIMIT));
    ub::ConvArgs args2 = args;
    args2.stages = Cfg2::plan_stages(ystg);
    const int smem = Cfg2::smem_bytes(args2.stages, ystg);
    const int pair_tiles = ((m_tiles + 1) / 2) * args.n_tiles;
    const int max_pairs = sms / 2;
    const int grid = 2 * (pair_tiles < max_pairs ? pair_tiles : max_pairs);   // cluster size 2 (__cluster_dims__)
    ub_launch(ub::conv_umma2_kernel<BN>, grid, ub::CONV_THREADS, smem, st, ma[0], ma[1], ma[2], ma[3], w, mo[0], mo[1], mo[2], mo[3], tmy, args2);
    UB_CUDA(cudaGetLastError());
    return UB_OK;
  }
  using Cfg = ub::ConvCfg<BN>;
  UB_CUDA(ensure_smem(ub::conv_umma_kernel<BN>, AT_UMMA + slot, Cfg::SMEM_LIMIT));
  ub::ConvArgs args2 = args;
  args2.stages = Cfg::plan_stages(ystg);
  const int smem = Cfg::smem_bytes(args2.stages, ystg);
  const int total = m_tiles * args.n_tiles;
  const int grid = total < sms ? total : sms;
  ub_launch(ub::conv_umma_kernel<BN>, grid, ub::CONV_THREADS, smem, st, ma[0], ma[1], ma[2], ma[3], w, mo[0], mo[1], mo[2], mo[3], tmy, args2);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

// Shared-memory carve-up of the CTA-pair kernel (half-size weight tiles): resident weights first.
bool halo2_smem_plan(int block_n, int kc, int head, ub::HaloArgs* a) {
  const int ystg = a->bn.y != nullptr;
  const int need = 3 * kc;
  if (need <= ub::HaloCfg::MAX_B) {
    for (int as = 4; as >= 2; --as) {
      if (ub::halo_smem_bytes(block_n, as, need, head, 1, ystg) <= ub::HaloCfg::SMEM_LIMIT) {
        a->resident = 1; a->a_stages = as; a->b_stages = need;
        return true;
      }
    }
  }
  for (int b = 6; b >= 2; --b) {
    for (int as = 4; as >= 3; --as) {
      if (ub::halo_smem_bytes(block_n, as, b, head, 1, ystg) <= ub::HaloCfg::SMEM_LIMIT) {
        a->resident = 0; a->a_stages = as; a->b_stages = b;
        return true;
      }
    }
  }
  return false;
}

template <int BN>
int launch_halo_t(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& mo,
                  ub::HaloArgs args, int slot, cudaStream_t st, const CUtensorMap* my) {
  const int head = args.epi == ub::HEPI_HEAD;
  const int ystg = args.bn.y != nullptr;
  if (ystg && (my == nullptr || head)) return fail(UB_ERR_ARG, "fused BN-backward sums: bad layer");
  const CUtensorMap& tmy = my != nullptr ? *my : mo;
  if (head && BN != 64) return fail(UB_ERR_ARG, "the fused head epilogue needs Cout == 64");
  const int total = args.tiles_w * args.tiles_h * args.B;
  const int sms = cur_sms();
  if (tl_opts->halo2 && total >= 2 && halo2_smem_plan(BN, args.kc0 + args.kc1, head, &args)) {
    UB_CUDA(ensure_smem(ub::conv_halo2_kernel<BN>, AT_HALO2 + slot, ub::HaloCfg::SMEM_LIMIT));
    const int smem = ub::halo_smem_bytes(BN, args.a_stages, args.b_stages, head, 1, ystg);
    const int pairs = (total + 1) / 2;
    const int max_pairs = sms / 2;
    const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);   // cluster size 2 (__cluster_dims__)
    ub_launch(ub::conv_halo2_kernel<BN>, grid, ub::HALO_THREADS, smem, st, a0, a1, w, mo, tmy, args);
    UB_CUDA(cudaGetLastError());
    return UB_OK;
  }
  UB_CUDA(ensure_smem(ub::conv_halo_kernel<BN>, AT_HALO + slot, ub::HaloCfg::SMEM_LIMIT));
  if (!halo_smem_plan(BN, args.kc0 + args.kc1, head, &args)) {
    return fail(UB_ERR_ARG, "no shared-memory plan for halo conv (N=%d, KC=%d)", BN, args.kc0 + args.kc1);
  }
  const int smem = ub::halo_smem_bytes(BN, args.a_stages, args.b_stages, head, 0, ystg);
  const int grid = total < sms ? total : sms;
  ub_launch(ub::conv_halo_kernel<BN>, grid, ub::HALO_THREADS, smem, st, a0, a1, w, mo, tmy, args);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int launch_halo(int block_n, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& mo,
                const ub::HaloArgs& args, cudaStream_t st, const CUtensorMap* my = nullptr) {
  int rc = device_check();
  if (rc != UB_OK) return rc;
  if (block_n == 64) return launch_halo_t<64>(a0, a1, w, mo, args, 0, st, my);
  if (block_n == 128) return launch_halo_t<128>(a0, a1, w, mo, args, 1, st, my);
  return fail(UB_ERR_ARG, "halo kernel supports Cout 64/128, got %d", block_n);
}

// Tile width of the tensor-core stem (stem_umma.cuh): 8 (16 x 8 tiles) or 32 (4 x 32 tiles, option "stem_wide").
int stem_tile_w() { return tl_opts->stem_wide ? 32 : 8; }
// TMA-store target of the stem: one TMEM lane quarter = (tw pixels) x (32 / tw rows)
int make_stem_out_map(CUtensorMap* m, void* out, int Bc, int H, int W, int tw) { return make_box_map(m, out, Bc, H, W, 64, tw, 32 / tw); }

template <int TW>
int launch_stem_umma_t(const CUtensorMap& mw, const CUtensorMap& mo, ub::StemArgs a, int slot, cudaStream_t st) {
  UB_CUDA(ensure_smem(ub::stem_umma_kernel<TW>, slot, ub::StemCfg::SMEM_BYTES));
  a.tiles_w = (a.W + TW - 1) / TW;
  a.tiles_h = (a.H + 128 / TW - 1) / (128 / TW);
  const int total = a.tiles_w * a.tiles_h * a.B;
  const int sms = cur_sms();
  const int grid = total < sms ? total : sms;
  ub_launch(ub::stem_umma_kernel<TW>, grid, ub::StemCfg::THREADS, ub::StemCfg::SMEM_BYTES, st, mw, mo, a);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

// tw must be the tile width the output map `mo` was built for (make_stem_out_map)
// b0: first image (x and the output map address the whole tensors; images [b0, b0 + B) are processed)
int launch_stem_umma(const CUtensorMap& mw, const CUtensorMap& mo, const void* x, const float* bias, int B, int H, int W,
                     int relu, cudaStream_t st, double* stat_sum = nullptr, double* stat_sumsq = nullptr, int tw = 8, int b0 = 0) {
  int rc = device_check();
  if (rc != UB_OK) return rc;
  ub::StemArgs a;
  a.B = B;
  a.b0 = b0;
  a.H = H;
  a.W = W;
  a.tiles_w = a.tiles_h = 0;
  a.relu = relu;
  a.x = reinterpret_cast<const uint2*>(x);
  a.bias = bias;
  a.stat_sum = stat_sum;
  a.stat_sumsq = stat_sumsq;
  if (tw == 32) return launch_stem_umma_t<32>(mw, mo, a, AT_STEM_WIDE, st);
  return launch_stem_umma_t<8>(mw, mo, a, AT_STEM, st);
}

int launch_conv(int block_n, const CUtensorMap* ma, const CUtensorMap& w, const CUtensorMap* mo, const ub::ConvArgs& args,
                cudaStream_t st, const CUtensorMap* my = nullptr) {
  int rc = device_check();
  if (rc != UB_OK) return rc;
  switch (block_n) {
    case 64: return launch_conv_t<64>(ma, w, mo, args, 0, st, my);
    case 128: return launch_conv_t<128>(ma, w, mo, args, 1, st, my);
    case 256: return launch_conv_t<256>(ma, w, mo, args, 2, st, my);
  }
  return fail(UB_ERR_ARG, "unsupported BLOCK_N %d", block_n);
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- plan -------------------------------------------------------------------------------------------
enum LayerKind { L_STEM, L_CONV, L_CONVT };

struct Layer {
  LayerKind kind;
  int H, W;          // GEMM-row grid (input spatial size)
  int C0, C1, Cout;  // conv: Cin split / Cout.  convT: C0 = Cin, Cout = f   (physical: multiples of 64 except the stem input)
  int lC0, lC1, lCout;  // logical channel counts of the reference tensors (<= physical; extra channels are exact zeros)
  int relu;
  int in0, in1, out, pool;  // activation buffer ids (-1 = none; in0 == -2: network input)
  size_t w_off, b_off;      // offsets into the weight buffer
  int block_n, TW, TH, TB;
  bool set;
  bool halo;       // runs on conv_halo_kernel
  bool fuse_head;  // last conv: 1x1 head + sigmoid + mask evaluated in its epilogue
  bool stem_tc;    // stem on tensor cores (Cout == 64)
  bool fuse_stem;  // enc0.conv1 only: the stem is computed inside this layer's patch producer (stem_halo.cuh), no stem launch
  CUtensorMap mWs; //   its weight map (the stem's packed [64][64] weights, box of 32 rows)
  int stem_tw;     // its tile width (8 or 32), fixed when the layer is created
  CUtensorMap mA0, mA1, mW;
  CUtensorMap mOut, mPool;  // TMA-store targets (halo layers)
  CUtensorMap mO[4];        // TMA-store targets of conv_umma_kernel: {out, pool, -, -} or the four ConvT quads
  CUtensorMap mY;           // training backward (dgrad layers): load view of the y tensor that has the output's geometry, same box
  int y_bytes;              //   as the output store view; bytes one such box delivers (0: no view built)
};

struct Buf {
  int H, W, C;
  size_t off, bytes;
  int first, last;   // layer index that writes it / last layer index that reads it (-1: never read; n_layers: the head kernel)
};

// Store views for a conv_umma_kernel layer. Conv: out (+ pooled out). ConvT: quad (dy,dx) of the [B,2H,2W,f] output is
// the tensor (f, W, H, B) at base + (dy*2W + dx)*f with pixel strides (2f, 4W*f, 4HW*f).
int make_umma_store_maps(Layer& l, void* out, void* pool, int Bc, int cm) {
  // cm = 2: split-precision tensors, 2*Cout physical channels [hi | lo] per pixel (the kernel stores lo at channel Cout + co)
  int rc;
  (void)pool;
  // one TMEM lane quarter (32 rows) of the tile box, see conv_args(): sub-box (TW, sh, sb)
  const int sb = l.TB >= 4 ? l.TB / 4 : 1;
  const int sh = l.TB >= 4 ? l.TH : (l.TB == 2 ? l.TH / 2 : l.TH / 4);
  if (l.kind == L_CONV) {
    const size_t C = (size_t)cm * l.Cout;
    rc = make_tile_store_map(&l.mO[0], out, Bc, l.H, l.W, (int)C, C, (size_t)l.W * C, (size_t)l.H * l.W * C, l.TW, sh, sb);
    if (rc != UB_OK) return rc;
    l.mO[1] = l.mO[0];
    l.mO[2] = l.mO[0];
    l.mO[3] = l.mO[0];
  } else {
    const size_t f = (size_t)cm * l.Cout;
    for (int qd = 0; qd < 4; ++qd) {
      const int dy = qd >> 1, dx = qd & 1;
      uint8_t* base = static_cast<uint8_t*>(out) + ((size_t)dy * 2 * l.W + dx) * f * 2;
      rc = make_tile_store_map(&l.mO[qd], base, Bc, l.H, l.W, (int)f, 2 * f, (size_t)4 * l.W * f, (size_t)4 * l.H * l.W * f, l.TW, sh, sb);
      if (rc != UB_OK) return rc;
    }
  }
  return UB_OK;
}

}  // namespace

struct unet_b200_plan {
  int Bc, H, W, in_ch, out_ch, levels;
  int split;                   // 1: fp32-class plan (every tensor [hi | lo] bf16, three K passes per product, fp32 network input)
  int feat[UB_MAX_LEVELS];
  Opts opt;                    // the switches this plan was created with (unet_b200_set_option changes later plans only)
  size_t ws_unshared_bytes;    // what the workspace would be with one private buffer per layer output (reporting)
  std::vector<Layer> layers;
  std::vector<int> conv_ids;   // layer index of 3x3 conv #i (stem included as conv 0)
  std::vector<int> convt_ids;  // layer index of ConvT #i
  std::vector<Buf> bufs;
  int final_buf;
  size_t ws_bytes, wt_bytes;
  size_t head_w_off, head_b_off;   // fp32 [out_ch][f0 physical], fp32 [out_ch]
  float head_bias;                 // host copy of bias[0] for the fused head epilogue (out_ch == 1)
  bool head_set;
  uint8_t* ws;
  uint8_t* wt;
  // copy/compute overlap of the host-buffer entry point (created on first use, destroyed with the plan)
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_in_free[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr},
              ev_out_free[2] = {nullptr, nullptr}, ev_start = nullptr;
  static constexpr int MAX_PIECES = 16;     // pieces of one pass (PieceHooks)
  cudaEvent_t ev_piece_in[MAX_PIECES] = {}, ev_piece_cmp[MAX_PIECES] = {};
};

namespace {

int add_buf(unet_b200_plan* p, int H, int W, int C) {
  Buf b;
  b.H = H;
  b.W = W;
  b.C = C;
  b.off = 0;
  b.bytes = align_up((size_t)p->Bc * H * W * C * 2 * (p->split ? 2 : 1), 1024);   // split tensors carry [hi C | lo C]
  b.first = -1;
  b.last = -1;
  p->bufs.push_back(b);
  return (int)p->bufs.size() - 1;
}

void add_conv(unet_b200_plan* p, LayerKind kind, int H, int W, int C0, int C1, int Cout, int in0, int in1, int out,
              int pool, int lC0 = -1, int lC1 = -1, int lCout = -1) {
  Layer l;
  memset(&l, 0, sizeof(l));
  l.kind = kind;
  l.H = H;
  l.W = W;
  l.C0 = C0;
  l.C1 = C1;
  l.Cout = Cout;
  l.lC0 = lC0 < 0 ? C0 : lC0;
  l.lC1 = lC1 < 0 ? C1 : lC1;
  l.lCout = lCout < 0 ? Cout : lCout;
  l.relu = (kind != L_CONVT);
  l.in0 = in0;
  l.in1 = in1;
  l.out = out;
  l.pool = pool;
  l.set = false;
  l.w_off = p->wt_bytes;
  const size_t kmul = p->split ? 3 : 1;   // split weights: [w_hi | w_hi | w_lo] per K source
  if (kind == L_STEM) {
    p->wt_bytes += align_up((size_t)36 * Cout * 4, 256);
  } else if (kind == L_CONV) {
    p->wt_bytes += align_up((size_t)Cout * 9 * (C0 + C1) * 2 * kmul, 256);
  } else {
    p->wt_bytes += align_up((size_t)4 * Cout * C0 * 2 * kmul, 256);
  }
  l.b_off = p->wt_bytes;
  p->wt_bytes += align_up((size_t)Cout * 4, 256);
  if (kind == L_STEM) {
    l.stem_tc = !p->split && tl_opts->stem_umma && Cout == 64;   // split plan: FP32-pipe stem on the fp32 image
    l.stem_tw = stem_tile_w();
  }
  if (kind != L_STEM) {
    pick_tile(H, W, &l.TW, &l.TH, &l.TB, p->Bc);
    l.block_n = pick_block_n(kind == L_CONV ? Cout : 4 * Cout);
    l.halo = !p->split && (kind == L_CONV) && halo_eligible(H, W, C0, C1, Cout);   // split plan: per-tap kernel only (4 K sources)
    if (l.halo) l.block_n = Cout;
    // Small plans (the per-frame executor path): a 14 x 14 level at batch 1 is 4 pixel tiles x 4 column blocks of 256 = 16
    // CTAs on 148 SMs. Narrower column blocks multiply the CTAs (the pixel tile is re-read from L2, which costs little at
    // this size). small_n = 1 (default): the block width with the lowest modelled time; 2: the first rule - halve BLOCK_N
    // while the layer cannot give every SM a tile. Plans of 64 frames and more keep 256 under both.
    if (!l.halo && tl_opts->small_n) {
      const int m_tiles = ((W + l.TW - 1) / l.TW) * ((H + l.TH - 1) / l.TH) * ((p->Bc + l.TB - 1) / l.TB);
      const int N = kind == L_CONV ? Cout : 4 * Cout;
      if (tl_opts->small_n == 1) {
        // wave model: rounds of the persistent grid x cost of one tile. Measured (tools/ab_small_n.py): a 128-wide tile
        // costs 0.56 of a 256-wide one, not 0.5 (the pixel tile is fetched twice as often per FLOP) - with 0.5 the model moved
        // a 16-frame 480 x 640 plan to 128-wide tiles and lost 7 %; a 64-wide tile 3/8 (its MMAs are bound by the A operand's
        // shared-memory reads)
        const bool pair = tl_opts->umma2 && m_tiles >= 2;
        const int slots = pair ? cur_sms() / 2 : cur_sms();
        long best_cost = -1;
        int best_bn = l.block_n;
        for (int bn = l.block_n; bn >= 64; bn /= 2) {
          if (N % bn != 0) continue;
          const long tiles = (long)(pair ? (m_tiles + 1) / 2 : m_tiles) * (N / bn);
          const long cost = ((tiles + slots - 1) / slots) * (bn == 256 ? 16 : bn == 128 ? 9 : 6);
          if (best_cost < 0 || cost * 103 < best_cost * 100) {    // narrower only for a gain of more than 3 %
            best_cost = cost;
            best_bn = bn;
          }
        }
        l.block_n = best_bn;
      } else {
        while (l.block_n > 64 && m_tiles * (N / l.block_n) < cur_sms()) l.block_n /= 2;
      }
    }
  }
  p->layers.push_back(l);
  if (kind == L_CONVT) {
    p->convt_ids.push_back((int)p->layers.size() - 1);
  } else {
    p->conv_ids.push_back((int)p->layers.size() - 1);
  }
}

ub::HaloArgs halo_args(const Layer& l, int batch, const float* bias, void* out, void* pool) {
  ub::HaloArgs a;
  memset(&a, 0, sizeof(a));
  a.B = batch;
  a.H = l.H;
  a.W = l.W;
  a.tiles_w = (l.W + 7) / 8;
  a.tiles_h = (l.H + 15) / 16;
  a.kc0 = l.C0 / 64;
  a.kc1 = l.C1 / 64;
  a.epi = ub::HEPI_STORE;
  a.relu = l.relu;
  a.pool_out = reinterpret_cast<__nv_bfloat16*>(pool);
  a.Cout = l.Cout;
  a.bias = bias;
  (void)out;
  return a;
}

ub::ConvArgs conv_args(const Layer& l, int batch, int batch_cap, const float* bias, void* out, void* pool) {
  ub::ConvArgs a;
  a.a_bytes = 128 * l.TW * l.TH * (l.TB < batch_cap ? l.TB : batch_cap);
  a.B = batch;
  a.H = l.H;
  a.W = l.W;
  a.TW = l.TW;
  a.TH = l.TH;
  a.TB = l.TB;
  a.tiles_w = (l.W + l.TW - 1) / l.TW;
  a.tiles_h = (l.H + l.TH - 1) / l.TH;
  a.tiles_b = (batch + l.TB - 1) / l.TB;
  const int N = (l.kind == L_CONV) ? l.Cout : 4 * l.Cout;
  a.n_tiles = N / l.block_n;
  a.taps = (l.kind == L_CONV) ? 9 : 1;
  a.kc0 = l.C0 / 64;
  a.kc1 = l.C1 / 64;
  a.kc2 = 0;
  a.kc3 = 0;
  a.epi = (l.kind == L_CONV) ? ub::EPI_STORE : ub::EPI_CONVT;
  a.relu = l.relu;
  a.Cout = l.Cout;
  a.bias = bias;
  a.pool_out = reinterpret_cast<__nv_bfloat16*>(pool);
  a.stat_sum = nullptr;
  a.stat_sumsq = nullptr;
  a.split = 0;
  memset(&a.bn, 0, sizeof(a.bn));
  if (l.TB >= 4) {
    a.sub_b = l.TB / 4;
    a.sub_h = l.TH;
  } else if (l.TB == 2) {
    a.sub_b = 1;
    a.sub_h = l.TH / 2;
  } else {
    a.sub_b = 1;
    a.sub_h = l.TH / 4;
  }
  (void)out;
  return a;
}

// Build the tensor maps of one stand-alone 3x3 conv (or ConvT when kind == L_CONVT) on explicit device pointers.
// x1 may be null (C1 == 0). Used by the single-layer entry points and by the trainer (train_capi.cuh).
int conv_layer_setup(Layer& l, LayerKind kind, const void* x0, int C0, const void* x1, int C1, const void* wp, void* y,
                     void* pool, int B, int H, int W, int Cout, int relu, bool allow_halo) {
  memset(&l, 0, sizeof(l));
  l.kind = kind;
  l.H = H;
  l.W = W;
  l.C0 = C0;
  l.C1 = C1;
  l.Cout = Cout;
  l.relu = relu;
  l.set = true;
  pick_tile(H, W, &l.TW, &l.TH, &l.TB, B);
  int rc;
  if (kind == L_CONVT) {
    l.block_n = pick_block_n(4 * Cout);
    rc = make_act_map(&l.mA0, x0, B, H, W, C0, l.TW, l.TH, l.TB);
    if (rc != UB_OK) return rc;
    l.mA1 = l.mA0;
    rc = make_w_map(&l.mW, wp, 4 * Cout, C0, l.block_n);
    if (rc != UB_OK) return rc;
    return make_umma_store_maps(l, y, nullptr, B);
  }
  l.block_n = pick_block_n(Cout);
  if (allow_halo && halo_eligible(H, W, C0, C1, Cout)) {
    l.halo = true;
    l.block_n = Cout;
    rc = make_halo_map(&l.mA0, x0, B, H, W, C0);
    if (rc != UB_OK) return rc;
    if (C1 > 0) {
      rc = make_halo_map(&l.mA1, x1, B, H, W, C1);
      if (rc != UB_OK) return rc;
    } else {
      l.mA1 = l.mA0;
    }
    rc = make_w_map(&l.mW, wp, Cout, 9 * (C0 + C1), l.block_n);
    if (rc != UB_OK) return rc;
    return make_box_map(&l.mOut, y, B, H, W, Cout, 8, 4);
  }
  if (pool != nullptr && (l.TW < 2 || l.TH < 2)) return fail(UB_ERR_ARG, "tile %dx%d cannot fuse the pool", l.TW, l.TH);
  rc = make_act_map(&l.mA0, x0, B, H, W, C0, l.TW, l.TH, l.TB);
  if (rc != UB_OK) return rc;
  if (C1 > 0) {
    rc = make_act_map(&l.mA1, x1, B, H, W, C1, l.TW, l.TH, l.TB);
    if (rc != UB_OK) return rc;
  } else {
    l.mA1 = l.mA0;
  }
  rc = make_w_map(&l.mW, wp, Cout, 9 * (C0 + C1), l.block_n);
  if (rc != UB_OK) return rc;
  return make_umma_store_maps(l, y, pool, B);
}

// Launch a layer prepared by conv_layer_setup (maps were built for batch capacity Bc).
// bn != null (training backward): the epilogue also takes the BatchNorm-backward sums of the layer whose gradient it writes;
// needs the y view of make_epi_y_map.
int conv_layer_launch(const Layer& l, int batch, int Bc, const float* bias, void* y, void* pool, cudaStream_t st,
                      double* stat_sum = nullptr, double* stat_sumsq = nullptr, const ub::EpiBnBwd* bn = nullptr) {
  if (bn != nullptr && l.y_bytes == 0) return fail(UB_ERR_STATE, "fused BN-backward sums without a y view");
  if (l.halo) {
    ub::HaloArgs ha = halo_args(l, batch, bias, y, pool);
    ha.stat_sum = stat_sum;
    ha.stat_sumsq = stat_sumsq;
    if (bn != nullptr) {
      ha.bn = *bn;
      ha.bn.bytes = l.y_bytes;
    }
    return launch_halo(l.block_n, l.mA0, l.mA1, l.mW, l.mOut, ha, st, bn != nullptr ? &l.mY : nullptr);
  }
  ub::ConvArgs a = conv_args(l, batch, Bc, bias, y, pool);
  a.stat_sum = stat_sum;
  a.stat_sumsq = stat_sumsq;
  if (bn != nullptr) {
    a.bn = *bn;
    a.bn.bytes = l.y_bytes;
  }
  const CUtensorMap ma[4] = {l.mA0, l.mA1, l.mA0, l.mA0};
  return launch_conv(l.block_n, ma, l.mW, l.mO, a, st, bn != nullptr ? &l.mY : nullptr);
}

// Load view of `ysrc` ([Bc][H][W][Cout] bf16, the geometry of the layer's output) with the box of the output store view, for
// the fused BatchNorm-backward sums of a dgrad layer (L_CONV, prepared by conv_layer_setup or setup_up_backward).
int make_epi_y_map(Layer& l, const void* ysrc, int Bc) {
  const size_t C = (size_t)l.Cout;
  if (l.halo) {
    l.y_bytes = 4096;
    return make_box_map(&l.mY, ysrc, Bc, l.H, l.W, l.Cout, 8, 4);
  }
  const int sb = l.TB >= 4 ? l.TB / 4 : 1;
  const int sh = l.TB >= 4 ? l.TH : (l.TB == 2 ? l.TH / 2 : l.TH / 4);
  l.y_bytes = 128 * l.TW * sh * (sb < Bc ? sb : Bc);
  return make_tile_store_map(&l.mY, ysrc, Bc, l.H, l.W, l.Cout, C, (size_t)l.W * C, (size_t)l.H * l.W * C, l.TW, sh, sb);
}

// Workspace layout by liveness: a layer output gets its address when the layer runs and gives it back after its last
// reader, so the workspace holds the live set (the skips + the tensors of the layer in flight: <= 19.3 MB per frame for the
// default network instead of 64.1 MB with one private buffer per output). Kernels run in stream order, so reuse needs no
// extra synchronisation; a layer's outputs are placed BEFORE its inputs are released (a conv reads neighbouring pixels of
// its input while it writes). Best fit over a coalescing free list, growing the top when nothing fits.
void plan_assign_workspace(unet_b200_plan* p) {
  const int n = (int)p->layers.size();
  for (Buf& b : p->bufs) b.first = b.last = -1;
  for (int li = 0; li < n; ++li) {
    const Layer& l = p->layers[li];
    if (l.out >= 0 && !l.fuse_head) p->bufs[l.out].first = li;
    if (l.pool >= 0) p->bufs[l.pool].first = li;
    if (l.in0 >= 0) p->bufs[l.in0].last = li;
    if (l.in1 >= 0) p->bufs[l.in1].last = li;
  }
  if (!p->layers.back().fuse_head) p->bufs[p->final_buf].last = n;   // read by the head kernel
  struct Blk { size_t off, size; };
  std::vector<Blk> fl;
  size_t top = 0;
  p->ws_unshared_bytes = 0;
  auto take = [&](size_t bytes) -> size_t {
    int best = -1;
    for (int i = 0; i < (int)fl.size(); ++i) {
      if (fl[i].size >= bytes && (best < 0 || fl[i].size < fl[best].size)) best = i;
    }
    if (best >= 0) {
      const size_t off = fl[best].off;
      fl[best].off += bytes;
      fl[best].size -= bytes;
      if (fl[best].size == 0) fl.erase(fl.begin() + best);
      return off;
    }
    // grow: a free block that touches the top is extended instead of left behind
    for (int i = 0; i < (int)fl.size(); ++i) {
      if (fl[i].off + fl[i].size == top) {
        const size_t off = fl[i].off;
        top = off + bytes;
        fl.erase(fl.begin() + i);
        return off;
      }
    }
    const size_t off = top;
    top += bytes;
    return off;
  };
  auto give = [&](size_t off, size_t bytes) {
    fl.push_back({off, bytes});
    std::sort(fl.begin(), fl.end(), [](const Blk& a, const Blk& b) { return a.off < b.off; });
    for (int i = 0; i + 1 < (int)fl.size();) {
      if (fl[i].off + fl[i].size == fl[i + 1].off) {
        fl[i].size += fl[i + 1].size;
        fl.erase(fl.begin() + i + 1);
      } else {
        ++i;
      }
    }
  };
  for (int li = 0; li < n; ++li) {
    const Layer& l = p->layers[li];
    const int outs[2] = {l.fuse_head ? -1 : l.out, l.pool};
    for (int o : outs) {
      if (o < 0) continue;
      Buf& b = p->bufs[o];
      b.off = take(b.bytes);
      p->ws_unshared_bytes += b.bytes;
    }
    for (int o : outs) {   // written but never read: free again once the layer's other output has its own space
      if (o >= 0 && p->bufs[o].last < 0) give(p->bufs[o].off, p->bufs[o].bytes);
    }
    const int ins[2] = {l.in0, l.in1};
    for (int i : ins) {
      if (i >= 0 && p->bufs[i].last == li) give(p->bufs[i].off, p->bufs[i].bytes);
    }
  }
  p->ws_bytes = top > 0 ? top : 1024;
}

int grid_for(size_t work_items, int threads) {
  size_t g = (work_items + threads - 1) / threads;
  const size_t cap = (size_t)cur_sms() * 16;
  if (g > cap) g = cap;
  if (g == 0) g = 1;
  return (int)g;
}

}  // namespace

extern "C" {

const char* unet_b200_last_error(void) { return g_err; }
int unet_b200_version(void) { return 100; }
int unet_b200_device_ok(void) { return device_check(); }

int unet_b200_plan_create(unet_b200_plan** out, int max_batch, int H, int W, int in_channels, int out_channels,
                          const int* features, int levels) {
  return unet_b200_plan_create_ex(out, max_batch, H, W, in_channels, out_channels, features, levels, UB_PRECISION_BF16);
}

int unet_b200_plan_precision(const unet_b200_plan* p) { return p ? (p->split ? UB_PRECISION_FP32 : UB_PRECISION_BF16) : -1; }

int unet_b200_plan_create_ex(unet_b200_plan** out, int max_batch, int H, int W, int in_channels, int out_channels,
                             const int* features, int levels, int precision) {
  if (out == nullptr || features == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (precision != UB_PRECISION_BF16 && precision != UB_PRECISION_FP32) return fail(UB_ERR_ARG, "unknown precision %d", precision);
  if (levels < 1 || levels > UB_MAX_LEVELS) return fail(UB_ERR_ARG, "levels must be in [1,%d]", UB_MAX_LEVELS);
  if (max_batch < 1) return fail(UB_ERR_ARG, "max_batch must be >= 1");
  if (in_channels < 1 || in_channels > 4) return fail(UB_ERR_ARG, "in_channels must be in [1,4] (got %d)", in_channels);
  if (out_channels < 1 || out_channels > 64) return fail(UB_ERR_ARG, "out_channels must be in [1,64] (got %d)", out_channels);
  if (H % (1 << levels) != 0 || W % (1 << levels) != 0) {
    return fail(UB_ERR_ARG, "H=%d and W=%d must be divisible by 2^levels=%d", H, W, 1 << levels);
  }
  for (int i = 0; i < levels; ++i) {
    if (features[i] <= 0 || features[i] > 4096) return fail(UB_ERR_ARG, "features[%d]=%d must be in [1,4096]", i, features[i]);
    if (precision == UB_PRECISION_FP32 && features[i] % 64 != 0) {
      return fail(UB_ERR_ARG, "the fp32-class plan needs features that are multiples of 64 (features[%d]=%d)", i, features[i]);
    }
  }
  if (precision == UB_PRECISION_FP32 && out_channels != 1) return fail(UB_ERR_ARG, "the fp32-class plan needs out_channels == 1");
  if (features[0] > 256) return fail(UB_ERR_ARG, "features[0]=%d: the first block's width must be <= 256 (stem kernel)", features[0]);
  // physical channel counts: 64-aligned (one 128-byte swizzled row per pixel and channel block); a logical width such as
  // the deployed topology's 32 is stored zero-extended, see pack_conv3x3_pad_kernel
  int fp[UB_MAX_LEVELS];
  for (int i = 0; i < levels; ++i) fp[i] = (features[i] + 63) / 64 * 64;
  unet_b200_plan* p = new (std::nothrow) unet_b200_plan();
  if (p == nullptr) return fail(UB_ERR_ARG, "out of host memory");
  p->Bc = max_batch;
  p->H = H;
  p->W = W;
  p->in_ch = in_channels;
  p->out_ch = out_channels;
  p->levels = levels;
  p->split = precision == UB_PRECISION_FP32 ? 1 : 0;
  p->opt = g_opts;                // this plan's switches from here on
  // Programmatic dependent launch: the conv / stem kernels run their prologue (barrier init, TMEM allocation, descriptor
  // prefetch) before griddepcontrol.wait, so with the launch attribute set it overlaps the predecessor's tail. That pays
  // when kernels are short - one 224 x 224 frame: 0.319 -> 0.262 ms per pass, 8 frames 0.62 -> 0.59 - is neutral at 16
  // frames and costs 2-3 % from 32 frames on (an early dependent's parked CTAs get in the way of a persistent grid's tail;
  // tools/ab_plan_option.py pdl 0 1, profiles/r2_ab_pdl.log): on by default for plans of up to 600 k pixels.
  if (p->opt.pdl < 0) p->opt.pdl = (size_t)max_batch * H * W <= 600000 ? 1 : 0;
  OptScope opt_scope(&p->opt);
  p->ws_bytes = 0;
  p->wt_bytes = 0;
  p->ws = nullptr;
  p->wt = nullptr;
  p->head_set = false;
  p->head_bias = 0.f;
  for (int i = 0; i < levels; ++i) p->feat[i] = features[i];

  // encoder (README.md:1432-1434, 1464-1467)
  int cur = -2;  // network input
  int cin = in_channels, lcin = in_channels;
  std::vector<int> skips;
  for (int i = 0; i < levels; ++i) {
    const int h = H >> i, w = W >> i, f = fp[i], lf = features[i];
    const int ea = add_buf(p, h, w, f);
    const int sk = add_buf(p, h, w, f);
    const int pl = add_buf(p, h / 2, w / 2, f);
    if (i == 0) {
      add_conv(p, L_STEM, h, w, cin, 0, f, cur, -1, ea, -1, lcin, 0, lf);
    } else {
      add_conv(p, L_CONV, h, w, cin, 0, f, cur, -1, ea, -1, lcin, 0, lf);
    }
    add_conv(p, L_CONV, h, w, f, 0, f, ea, -1, sk, pl, lf, 0, lf);
    skips.push_back(sk);
    cur = pl;
    cin = f;
    lcin = lf;
  }
  // bottleneck (README.md:1437, 1470)
  {
    const int h = H >> levels, w = W >> levels, lf = features[levels - 1] * 2, f = (lf + 63) / 64 * 64;
    const int ba = add_buf(p, h, w, f);
    const int bb = add_buf(p, h, w, f);
    add_conv(p, L_CONV, h, w, cin, 0, f, cur, -1, ba, -1, lcin, 0, lf);
    add_conv(p, L_CONV, h, w, f, 0, f, ba, -1, bb, -1, lf, 0, lf);
    cur = bb;
    cin = f;
    lcin = lf;
  }
  // decoder (README.md:1440-1444, 1473-1479): ConvT, then double conv over cat([skip, up])
  for (int j = 0; j < levels; ++j) {
    const int i = levels - 1 - j;
    const int h = H >> i, w = W >> i, f = fp[i], lf = features[i];
    const int up = add_buf(p, h, w, f);
    const int da = add_buf(p, h, w, f);
    const int db = add_buf(p, h, w, f);
    add_conv(p, L_CONVT, h / 2, w / 2, cin, 0, f, cur, -1, up, -1, lcin, 0, lf);
    add_conv(p, L_CONV, h, w, f, f, f, skips[i], up, da, -1, lf, lf, lf);
    add_conv(p, L_CONV, h, w, f, 0, f, da, -1, db, -1, lf, 0, lf);
    cur = db;
    cin = f;
    lcin = lf;
  }
  p->final_buf = cur;
  {
    Layer& last = p->layers.back();
    last.fuse_head = tl_opts->fuse_head && out_channels == 1 && last.kind == L_CONV && last.halo && last.Cout == 64;
  }
  if (p->layers.size() >= 2) {
    // the stem inside enc0.conv1 (stem_halo.cuh): tensor-core stem, second conv 64 -> 64 on the CTA-pair halo kernel
    const Layer& l0 = p->layers[0];
    Layer& l1 = p->layers[1];
    l1.fuse_stem = tl_opts->stem_fuse && tl_opts->halo2 && !p->split && l0.kind == L_STEM && l0.stem_tc && l1.kind == L_CONV &&
                   l1.halo && l1.C0 == 64 && l1.C1 == 0 && l1.Cout == 64;
  }
  p->head_w_off = p->wt_bytes;
  p->wt_bytes += align_up((size_t)out_channels * fp[0] * 4 * (p->split ? 2 : 1), 256);   // split: the weight vector twice (hi + lo)
  p->head_b_off = p->wt_bytes;
  p->wt_bytes += align_up((size_t)out_channels * 4, 256);
  plan_assign_workspace(p);
  *out = p;
  return UB_OK;
}

void unet_b200_plan_destroy(unet_b200_plan* p) {
  if (p == nullptr) return;
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  for (int i = 0; i < 2; ++i) {
    if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
    if (p->ev_in_free[i]) cudaEventDestroy(p->ev_in_free[i]);
    if (p->ev_cmp[i]) cudaEventDestroy(p->ev_cmp[i]);
    if (p->ev_out_free[i]) cudaEventDestroy(p->ev_out_free[i]);
  }
  if (p->ev_start) cudaEventDestroy(p->ev_start);
  for (int i = 0; i < unet_b200_plan::MAX_PIECES; ++i) {
    if (p->ev_piece_in[i]) cudaEventDestroy(p->ev_piece_in[i]);
    if (p->ev_piece_cmp[i]) cudaEventDestroy(p->ev_piece_cmp[i]);
  }
  delete p;
}
size_t unet_b200_plan_workspace_bytes(const unet_b200_plan* p) { return p ? p->ws_bytes : 0; }
size_t unet_b200_plan_workspace_unshared_bytes(const unet_b200_plan* p) { return p ? p->ws_unshared_bytes : 0; }
size_t unet_b200_plan_weight_bytes(const unet_b200_plan* p) { return p ? p->wt_bytes : 0; }
int unet_b200_plan_num_convs(const unet_b200_plan* p) { return p ? (int)p->conv_ids.size() : 0; }

int unet_b200_plan_bind(unet_b200_plan* p, void* workspace_dev, void* weights_dev) {
  if (p == nullptr || workspace_dev == nullptr || weights_dev == nullptr) return fail(UB_ERR_ARG, "null argument");
  if ((reinterpret_cast<uintptr_t>(workspace_dev) & 255) || (reinterpret_cast<uintptr_t>(weights_dev) & 255)) {
    return fail(UB_ERR_ARG, "workspace and weight buffers must be 256-byte aligned");
  }
  int rc = device_check();
  if (rc != UB_OK) return rc;
  OptScope opt_scope(&p->opt);
  p->ws = static_cast<uint8_t*>(workspace_dev);
  p->wt = static_cast<uint8_t*>(weights_dev);
  for (Layer& l : p->layers) {
    if (l.kind == L_STEM) {
      if (l.stem_tc) {
        rc = make_w_map_box(&l.mW, p->wt + l.w_off, 64, 64, 64);
        if (rc != UB_OK) return rc;
        rc = make_stem_out_map(&l.mOut, p->ws + p->bufs[l.out].off, p->Bc, l.H, l.W, l.stem_tw);
        if (rc != UB_OK) return rc;
        Layer& l1 = p->layers[1];
        if (l1.fuse_stem) {
          rc = make_w_map_box(&l1.mWs, p->wt + l.w_off, 64, 64, 32);
          if (rc != UB_OK) return rc;
        }
      }
      continue;
    }
    const Buf& b0 = p->bufs[l.in0];
    if (p->split) {
      // fp32-class plan: tensors are [hi C | lo C]; one map per input over its 2C channels serves both K sources (hi|lo, hi)
      rc = make_act_map(&l.mA0, p->ws + b0.off, p->Bc, l.H, l.W, 2 * l.C0, l.TW, l.TH, l.TB);
      if (rc != UB_OK) return rc;
      l.mA1 = l.mA0;
      if (l.C1 > 0) {
        rc = make_act_map(&l.mA1, p->ws + p->bufs[l.in1].off, p->Bc, l.H, l.W, 2 * l.C1, l.TW, l.TH, l.TB);
        if (rc != UB_OK) return rc;
      }
      rc = make_umma_store_maps(l, p->ws + p->bufs[l.out].off, nullptr, p->Bc, 2);
      if (rc != UB_OK) return rc;
      rc = make_w_map(&l.mW, p->wt + l.w_off, l.kind == L_CONV ? l.Cout : 4 * l.Cout, (l.kind == L_CONV ? 9 : 1) * 3 * (l.C0 + l.C1),
                      l.block_n);
      if (rc != UB_OK) return rc;
      continue;
    }
    rc = l.halo ? make_halo_map(&l.mA0, p->ws + b0.off, p->Bc, l.H, l.W, l.C0)
                : make_act_map(&l.mA0, p->ws + b0.off, p->Bc, l.H, l.W, l.C0, l.TW, l.TH, l.TB);
    if (rc != UB_OK) return rc;
    if (l.C1 > 0) {
      const Buf& b1 = p->bufs[l.in1];
      rc = l.halo ? make_halo_map(&l.mA1, p->ws + b1.off, p->Bc, l.H, l.W, l.C1)
                  : make_act_map(&l.mA1, p->ws + b1.off, p->Bc, l.H, l.W, l.C1, l.TW, l.TH, l.TB);
    } else {
      l.mA1 = l.mA0;
    }
    if (rc != UB_OK) return rc;
    if (!l.halo) {
      rc = make_umma_store_maps(l, p->ws + p->bufs[l.out].off, l.pool >= 0 ? p->ws + p->bufs[l.pool].off : nullptr, p->Bc);
      if (rc != UB_OK) return rc;
    }
    if (l.halo) {
      const Buf& bo = p->bufs[l.out];
      rc = make_box_map(&l.mOut, p->ws + bo.off, p->Bc, l.H, l.W, l.Cout, 8, 4);
      if (rc != UB_OK) return rc;
    }
    if (l.kind == L_CONV) {
      rc = make_w_map(&l.mW, p->wt + l.w_off, l.Cout, 9 * (l.C0 + l.C1), l.block_n);
    } else {
      rc = make_w_map(&l.mW, p->wt + l.w_off, 4 * l.Cout, l.C0, l.block_n);
    }
    if (rc != UB_OK) return rc;
  }
  return UB_OK;
}

int unet_b200_plan_set_conv(unet_b200_plan* p, int idx, const float* w, const float* gamma, const float* beta,
                            const float* mean, const float* var, float eps, void* stream) {
  if (p == nullptr || w == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (p->wt == nullptr) return fail(UB_ERR_STATE, "plan_bind must be called before set_conv");
  if (idx < 0 || idx >= (int)p->conv_ids.size()) return fail(UB_ERR_ARG, "conv index %d out of range", idx);
  Layer& l = p->layers[p->conv_ids[idx]];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* bias = reinterpret_cast<float*>(p->wt + l.b_off);
  if (p->split) {
    if (l.kind == L_STEM) {   // fp32 weights, no bf16 rounding
      ub_launch(ub::pack_stem_kernel, grid_for(36 * l.Cout, 256), 256, 0, st, w, gamma, beta, mean, var, eps, l.Cout, l.C0,
                reinterpret_cast<float*>(p->wt + l.w_off), bias, 1, l.lCout);
    } else {
      ub_launch(ub::pack_conv3x3_split_kernel, grid_for((size_t)l.Cout * 27 * (l.C0 + l.C1), 256), 256, 0, st, w, gamma, beta, mean,
                var, eps, l.Cout, l.C0, l.C1, reinterpret_cast<__nv_bfloat16*>(p->wt + l.w_off), bias);
    }
    UB_CUDA(cudaGetLastError());
    l.set = true;
    return UB_OK;
  }
  if (l.kind == L_STEM && l.stem_tc) {
    // logical Cout rows are written; rows / bias entries of padded channels keep the zeros the buffer was created with
    UB_CUDA(cudaMemsetAsync(p->wt + l.w_off, 0, (size_t)64 * 64 * 2, st));
    UB_CUDA(cudaMemsetAsync(bias, 0, (size_t)l.Cout * 4, st));
    ub_launch(ub::pack_stem_umma_kernel, grid_for(64 * l.lCout, 256), 256, 0, st, 
        w, gamma, beta, mean, var, eps, l.lCout, l.C0, reinterpret_cast<__nv_bfloat16*>(p->wt + l.w_off), bias);
  } else if (l.kind == L_STEM) {
    ub_launch(ub::pack_stem_kernel, grid_for(36 * l.Cout, 256), 256, 0, st, w, gamma, beta, mean, var, eps, l.Cout, l.C0,
                                                                     reinterpret_cast<float*>(p->wt + l.w_off), bias, 0, l.lCout);
  } else {
    const int cin = l.C0 + l.C1;
    ub_launch(ub::pack_conv3x3_pad_kernel, grid_for((size_t)l.Cout * 9 * cin, 256), 256, 0, st, 
        w, gamma, beta, mean, var, eps, l.lCout, l.lC0, l.lC1, l.Cout, l.C0, l.C1,
        reinterpret_cast<__nv_bfloat16*>(p->wt + l.w_off), bias);
  }
  UB_CUDA(cudaGetLastError());
  l.set = true;
  return UB_OK;
}

int unet_b200_plan_set_convT(unet_b200_plan* p, int idx, const float* w, const float* bias, void* stream) {
  if (p == nullptr || w == nullptr || bias == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (p->wt == nullptr) return fail(UB_ERR_STATE, "plan_bind must be called before set_convT");
  if (idx < 0 || idx >= (int)p->convt_ids.size()) return fail(UB_ERR_ARG, "convT index %d out of range", idx);
  Layer& l = p->layers[p->convt_ids[idx]];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->split) {
    ub_launch(ub::pack_convT_split_kernel, grid_for((size_t)12 * l.Cout * l.C0, 256), 256, 0, st, w, l.C0, l.Cout,
              reinterpret_cast<__nv_bfloat16*>(p->wt + l.w_off));
    UB_CUDA(cudaGetLastError());
    UB_CUDA(cudaMemcpyAsync(p->wt + l.b_off, bias, (size_t)l.Cout * 4, cudaMemcpyDeviceToDevice, st));
    l.set = true;
    return UB_OK;
  }
  ub_launch(ub::pack_convT_pad_kernel, grid_for((size_t)4 * l.Cout * l.C0, 256), 256, 0, st, 
      w, bias, l.lC0, l.lCout, l.C0, l.Cout, reinterpret_cast<__nv_bfloat16*>(p->wt + l.w_off),
      reinterpret_cast<float*>(p->wt + l.b_off));
  UB_CUDA(cudaGetLastError());
  l.set = true;
  return UB_OK;
}

int unet_b200_plan_set_head(unet_b200_plan* p, const float* w, const float* bias, void* stream) {
  if (p == nullptr || w == nullptr || bias == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (p->wt == nullptr) return fail(UB_ERR_STATE, "plan_bind must be called before set_head");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t f0 = (size_t)p->feat[0], f0p = (f0 + 63) / 64 * 64;
  // rows of the logical width copied into zero-extended rows of the physical width
  UB_CUDA(cudaMemsetAsync(p->wt + p->head_w_off, 0, (size_t)p->out_ch * f0p * 4, st));
  UB_CUDA(cudaMemcpy2DAsync(p->wt + p->head_w_off, f0p * 4, w, f0 * 4, f0 * 4, (size_t)p->out_ch, cudaMemcpyDeviceToDevice, st));
  if (p->split) {   // the head runs over [hi | lo]: the same weights for both halves
    UB_CUDA(cudaMemcpyAsync(p->wt + p->head_w_off + f0p * 4, w, f0 * 4, cudaMemcpyDeviceToDevice, st));
  }
  UB_CUDA(cudaMemcpyAsync(p->wt + p->head_b_off, bias, (size_t)p->out_ch * 4, cudaMemcpyDeviceToDevice, st));
  UB_CUDA(cudaMemcpyAsync(&p->head_bias, bias, 4, cudaMemcpyDeviceToHost, st));
  UB_CUDA(cudaStreamSynchronize(st));
  p->head_set = true;
  return UB_OK;
}

int unet_b200_forward_launches(const unet_b200_plan* p) {
  if (p == nullptr) return 0;
  int n = (int)p->layers.size() + (p->layers.back().fuse_head ? 0 : 1);
  if (p->layers.size() > 1 && p->layers[1].fuse_stem) --n;   // the stem runs inside enc0.conv1
  if (p->split) {
    for (const Layer& l : p->layers) n += l.pool >= 0 ? 1 : 0;   // the 2x2 pools are their own kernels there
  }
  return n;
}

int unet_b200_set_option(const char* name, int value) {
  if (name == nullptr) return fail(UB_ERR_ARG, "null option name");
  struct { const char* n; int* v; } tab[] = {
      {"halo", &g_opts.halo}, {"pdl", &g_opts.pdl}, {"halo2", &g_opts.halo2}, {"umma2", &g_opts.umma2},
      {"fuse_head", &g_opts.fuse_head}, {"stem_umma", &g_opts.stem_umma}, {"wgrad_rows64", &g_opts.wgrad_rows64},
      {"wgrad2", &g_opts.wgrad2}, {"wgrad_stream", &g_opts.wgrad_stream}, {"bwd_fuse", &g_opts.bwd_fuse},
      {"stem_wide", &g_opts.stem_wide}, {"dgrad_fuse", &g_opts.dgrad_fuse},
      {"wgrad_halo", &g_opts.wgrad_halo}, {"pack_split", &g_opts.pack_split},
      {"host_pieces", &g_opts.host_pieces}, {"stem_fuse", &g_opts.stem_fuse}, {"pre_bulk", &g_opts.pre_bulk}, {"pre_rows", &g_opts.pre_rows}, {"pre_stages", &g_opts.pre_stages},
      {"host_hybrid", &g_opts.host_hybrid}, {"host_geometric", &g_opts.host_geometric},
      {"small_n", &g_opts.small_n}};
  for (auto& e : tab) {
    if (strcmp(name, e.n) == 0) {
      *e.v = value;
      return UB_OK;
    }
  }
  return fail(UB_ERR_ARG, "unknown option '%s'", name);
}

// Stem + enc0.conv1 as one kernel (stem_halo.cuh) on images [b0, b0 + batch); needs at least two tiles (a CTA pair).
static bool stem_fused(const unet_b200_plan* p, int batch) {
  if (p->layers.size() < 2 || !p->layers[1].fuse_stem) return false;
  const Layer& l1 = p->layers[1];
  return ((l1.W + 7) / 8) * ((l1.H + 15) / 16) * batch >= 2;
}
static int launch_stem_halo(unet_b200_plan* p, const void* x, int batch, int b0, cudaStream_t st) {
  const Layer& l0 = p->layers[0];
  const Layer& l1 = p->layers[1];
  UB_CUDA(ensure_smem(ub::stem_halo2_kernel, AT_STEM_HALO, ub::StemHaloCfg::SMEM_BYTES));
  ub::StemHaloArgs a;
  a.B = batch;
  a.H = l1.H;
  a.W = l1.W;
  a.b0 = b0;
  a.tiles_w = (l1.W + 7) / 8;
  a.tiles_h = (l1.H + 15) / 16;
  a.x = reinterpret_cast<const uint2*>(x);
  a.stem_bias = reinterpret_cast<const float*>(p->wt + l0.b_off);
  a.bias = reinterpret_cast<const float*>(p->wt + l1.b_off);
  a.relu = l1.relu;
  a.pool_out = l1.pool >= 0 ? reinterpret_cast<__nv_bfloat16*>(p->ws + p->bufs[l1.pool].off) : nullptr;
  const int total = a.tiles_w * a.tiles_h * batch;
  const int pairs = (total + 1) / 2;
  const int max_pairs = cur_sms() / 2;
  const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);   // cluster size 2 (__cluster_dims__)
  ub_launch(ub::stem_halo2_kernel, grid, ub::StemHaloCfg::THREADS, ub::StemHaloCfg::SMEM_BYTES, st, l1.mW, l1.mWs, l1.mOut, a);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

// Pieces (host-buffer entry point): the two layers of the full-resolution level at the START of the network (stem, enc0.conv1)
// and the LAST layer (dec3.conv1 with the fused head) run once per piece of `piece` images instead of once per batch, with a
// callback before the first / after the last: the caller hangs the piece's H2D copy + preprocess in front and its D2H copy
// behind, so that only the first piece's input copy and the last piece's output copy are not hidden by kernels. Those three
// layers have >= 392 tiles per image, so a piece of 32 images still fills the GPU; every other layer runs on the whole batch.
struct PieceHooks {
  int np;                                               // pieces of this pass (<= unet_b200_plan::MAX_PIECES)
  int front_b0[unet_b200_plan::MAX_PIECES], front_n[unet_b200_plan::MAX_PIECES];   // pieces of the first two layers, in run order
  int tail_b0[unet_b200_plan::MAX_PIECES], tail_n[unet_b200_plan::MAX_PIECES];     // pieces of the last layer, in run order
  void* ctx;
  int (*before)(void* ctx, int i, int b0, int n);       // before the first layer works on front piece i = images [b0, b0 + n)
  int (*after)(void* ctx, int i, int b0, int n);        // after the last layer has been enqueued for tail piece i
};
static bool plan_supports_pieces(const unet_b200_plan* p) {
  if (p->split || p->layers.size() < 4) return false;
  const Layer& l0 = p->layers[0];
  const Layer& l1 = p->layers[1];
  const Layer& ll = p->layers.back();
  return l0.kind == L_STEM && l0.stem_tc && l1.kind == L_CONV && l1.halo && ll.kind == L_CONV && ll.halo && ll.fuse_head;
}

static int forward_impl(unet_b200_plan* p, const void* x, int batch, float* logits, float* probs, uint8_t* mask,
                        float threshold, cudaStream_t st, std::vector<cudaEvent_t>* ev, const PieceHooks* ph = nullptr) {
  if (p == nullptr || x == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (p->ws == nullptr) return fail(UB_ERR_STATE, "plan is not bound");
  if (batch < 1 || batch > p->Bc) return fail(UB_ERR_ARG, "batch %d outside [1,%d]", batch, p->Bc);
  if (!p->head_set) return fail(UB_ERR_STATE, "head weights not set");
  for (const Layer& l : p->layers) {
    if (!l.set) return fail(UB_ERR_STATE, "layer weights not set");
  }
  OptScope opt_scope(&p->opt);
  size_t ei = 0;
  if (ev) UB_CUDA(cudaEventRecord((*ev)[ei++], st));
  if (ph != nullptr && (ph->np < 1 || !plan_supports_pieces(p))) return fail(UB_ERR_STATE, "plan cannot run in pieces");
  const int n_layers = (int)p->layers.size();
  for (int li = 0; li < n_layers; ++li) {
    Layer& l = p->layers[li];
    const float* bias = reinterpret_cast<const float*>(p->wt + l.b_off);
    void* out = p->ws + p->bufs[l.out].off;
    void* pool = l.pool >= 0 ? p->ws + p->bufs[l.pool].off : nullptr;
    if (ph != nullptr && (li == 0 || li == n_layers - 1)) {
      // piece-wise: (stem + enc0.conv1) per piece up front, the fused-head conv per piece at the end
      for (int i = 0; i < ph->np; ++i) {
        const int b0 = li == 0 ? ph->front_b0[i] : ph->tail_b0[i];
        const int n = li == 0 ? ph->front_n[i] : ph->tail_n[i];
        int rc;
        if (li == 0) {
          rc = ph->before(ph->ctx, i, b0, n);
          if (rc != UB_OK) return rc;
          if (stem_fused(p, n)) {
            rc = launch_stem_halo(p, x, n, b0, st);
            if (rc != UB_OK) return rc;
          } else {
            rc = launch_stem_umma(l.mW, l.mOut, x, bias, n, l.H, l.W, l.relu, st, nullptr, nullptr, l.stem_tw, b0);
            if (rc != UB_OK) return rc;
            Layer& l1 = p->layers[1];
            ub::HaloArgs a = halo_args(l1, n, reinterpret_cast<const float*>(p->wt + l1.b_off), p->ws + p->bufs[l1.out].off,
                                       l1.pool >= 0 ? p->ws + p->bufs[l1.pool].off : nullptr);
            a.b0 = b0;
            rc = launch_halo(l1.block_n, l1.mA0, l1.mA1, l1.mW, l1.mOut, a, st);
            if (rc != UB_OK) return rc;
          }
        } else {
          ub::HaloArgs a = halo_args(l, n, bias, out, pool);
          a.b0 = b0;
          a.epi = ub::HEPI_HEAD;
          a.head_w = reinterpret_cast<const float*>(p->wt + p->head_w_off);
          a.head_b = p->head_bias;
          a.thr = threshold;
          a.logits = logits;
          a.probs = probs;
          a.mask = mask;
          rc = launch_halo(l.block_n, l.mA0, l.mA1, l.mW, l.mOut, a, st);
          if (rc != UB_OK) return rc;
          rc = ph->after(ph->ctx, i, b0, n);
          if (rc != UB_OK) return rc;
        }
      }
      if (li == 0) ++li;   // enc0.conv1 ran with the stem
      continue;
    }
    if (p->split) {
      // fp32-class plan: stem on the FP32 pipes from the fp32 NHWC4 image, every other layer = the per-tap
      // tcgen05 kernel over K sources (x[hi|lo] : 2C), (x[hi] : C) against weights [w_hi | w_hi | w_lo]; pools on hi + lo
      if (l.kind == L_STEM) {
        const int tiles = ((l.W + 15) / 16) * ((l.H + 15) / 16) * batch;
        const size_t smem = (size_t)(36 * l.Cout + l.Cout) * 4 + 18 * 18 * 16;
        ub_launch(ub::stem_conv_kernel<2>, tiles, 256, smem, st, x, reinterpret_cast<const float*>(p->wt + l.w_off), bias, batch,
                  l.H, l.W, l.C0, l.Cout, l.relu, reinterpret_cast<__nv_bfloat16*>(out));
        UB_CUDA(cudaGetLastError());
      } else {
        ub::ConvArgs a = conv_args(l, batch, p->Bc, bias, out, nullptr);
        a.kc0 = 2 * l.C0 / 64;
        a.kc1 = l.C0 / 64;
        a.kc2 = 2 * l.C1 / 64;
        a.kc3 = l.C1 / 64;
        a.split = 1;
        a.pool_out = nullptr;
        const CUtensorMap ma[4] = {l.mA0, l.mA0, l.mA1, l.mA1};
        int rc = launch_conv(l.block_n, ma, l.mW, l.mO, a, st);
        if (rc != UB_OK) return rc;
      }
      if (l.pool >= 0) {
        const size_t n = (size_t)batch * (l.H / 2) * (l.W / 2) * (l.Cout / 8);
        ub_launch(ub::maxpool2x2_split_kernel, grid_for(n, 256), 256, 0, st, reinterpret_cast<const uint4*>(out), batch, l.H, l.W,
                  l.Cout / 8, reinterpret_cast<uint4*>(pool));
        UB_CUDA(cudaGetLastError());
      }
    } else if (l.kind == L_STEM && l.stem_tc && stem_fused(p, batch)) {
      // (computed inside the next layer)
    } else if (li == 1 && l.fuse_stem && stem_fused(p, batch)) {
      int rc = launch_stem_halo(p, x, batch, 0, st);
      if (rc != UB_OK) return rc;
    } else if (l.kind == L_STEM && l.stem_tc) {
      int rc = launch_stem_umma(l.mW, l.mOut, x, bias, batch, l.H, l.W, l.relu, st, nullptr, nullptr, l.stem_tw);
      if (rc != UB_OK) return rc;
    } else if (l.kind == L_STEM) {
      const int tiles = ((l.W + 15) / 16) * ((l.H + 15) / 16) * batch;
      const size_t smem = (size_t)(36 * l.Cout + l.Cout) * 4 + 18 * 18 * 16;
      ub_launch(ub::stem_conv_kernel<0>, tiles, 256, smem, st, reinterpret_cast<const uint2*>(x),
                                                      reinterpret_cast<const float*>(p->wt + l.w_off), bias, batch, l.H,
                                                      l.W, l.C0, l.Cout, l.relu, reinterpret_cast<__nv_bfloat16*>(out));
      UB_CUDA(cudaGetLastError());
    } else if (l.halo) {
      ub::HaloArgs a = halo_args(l, batch, bias, out, pool);
      if (l.fuse_head) {
        a.epi = ub::HEPI_HEAD;
        a.head_w = reinterpret_cast<const float*>(p->wt + p->head_w_off);
        a.head_b = p->head_bias;
        a.thr = threshold;
        a.logits = logits;
        a.probs = probs;
        a.mask = mask;
      }
      int rc = launch_halo(l.block_n, l.mA0, l.mA1, l.mW, l.mOut, a, st);
      if (rc != UB_OK) return rc;
    } else {
      ub::ConvArgs a = conv_args(l, batch, p->Bc, bias, out, pool);
      const CUtensorMap ma[4] = {l.mA0, l.mA1, l.mA0, l.mA0};
      int rc = launch_conv(l.block_n, ma, l.mW, l.mO, a, st);
      if (rc != UB_OK) return rc;
    }
    if (ev) UB_CUDA(cudaEventRecord((*ev)[ei++], st));
  }
  if (!p->layers.back().fuse_head) {
    const Buf& fb = p->bufs[p->final_buf];
    const size_t npix = (size_t)batch * fb.H * fb.W;
    if (p->out_ch == 1) {
      ub_launch(ub::head_kernel, grid_for(npix * 8, 256), 256, 0, st,
          reinterpret_cast<const __nv_bfloat16*>(p->ws + fb.off), reinterpret_cast<const float*>(p->wt + p->head_w_off),
          p->head_bias, npix, fb.C * (p->split ? 2 : 1), logits, probs, mask, threshold);
    } else {
      // out_channels > 1 (README.md:1447 builds any): outputs are NCHW [batch][out_ch][H][W]
      const size_t smem = ((size_t)p->out_ch * fb.C + p->out_ch) * 4;
      ub_launch(ub::head_multi_kernel, grid_for(npix, 256), 256, smem, st,
          reinterpret_cast<const __nv_bfloat16*>(p->ws + fb.off), reinterpret_cast<const float*>(p->wt + p->head_w_off),
          reinterpret_cast<const float*>(p->wt + p->head_b_off), batch, (size_t)fb.H * fb.W, fb.C, p->out_ch, logits, probs, mask,
          threshold);
    }
    UB_CUDA(cudaGetLastError());
  }
  if (ev) UB_CUDA(cudaEventRecord((*ev)[ei++], st));
  return UB_OK;
}

int unet_b200_forward(unet_b200_plan* p, const void* x, int batch, float* logits, float* probs, uint8_t* mask,
                      float threshold, void* stream) {
  return forward_impl(p, x, batch, logits, probs, mask, threshold, static_cast<cudaStream_t>(stream), nullptr);
}

int unet_b200_forward_profile(unet_b200_plan* p, const void* x, int batch, float* logits, float* probs, uint8_t* mask,
                              float threshold, void* stream, float* ms_out, int n_out) {
  if (p == nullptr || ms_out == nullptr) return fail(UB_ERR_ARG, "null argument");
  const int n = (int)p->layers.size() + 1;
  if (n_out < n) return fail(UB_ERR_ARG, "ms_out holds %d entries, need %d", n_out, n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) UB_CUDA(cudaEventCreate(&e));
  int rc = forward_impl(p, x, batch, logits, probs, mask, threshold, st, &ev);
  if (rc == UB_OK) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(UB_ERR_CUDA, "profile sync failed: %s", cudaGetErrorString(e));
  }
  if (rc == UB_OK) {
    for (int i = 0; i < n; ++i) cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]);
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

int unet_b200_plan_num_layers(const unet_b200_plan* p) { return p ? (int)p->layers.size() + 1 : 0; }

int unet_b200_plan_layer_info(const unet_b200_plan* p, int idx, int* info8) {
  if (p == nullptr || info8 == nullptr) return fail(UB_ERR_ARG, "null argument");
  const int n = (int)p->layers.size();
  if (idx < 0 || idx > n) return fail(UB_ERR_ARG, "layer index %d out of range", idx);
  if (idx == n) {  // head
    const Buf& fb = p->bufs[p->final_buf];
    const int v[8] = {3, fb.H, fb.W, fb.C, 1, 1, 0, 0};
    memcpy(info8, v, sizeof(v));
    return UB_OK;
  }
  const Layer& l = p->layers[idx];
  const int v[8] = {(int)l.kind, l.H, l.W, l.C0 + l.C1, l.Cout, l.kind == L_CONVT ? 1 : 9, l.block_n,
                    (l.pool >= 0 ? 1 : 0) | (l.halo ? 2 : 0) | (l.fuse_head ? 4 : 0)};
  memcpy(info8, v, sizeof(v));
  return UB_OK;
}

int unet_b200_nchw_to_nhwc4(const float* x, int batch, int C, int H, int W, void* y, void* stream) {
  if (x == nullptr || y == nullptr || C < 1 || C > 4) return fail(UB_ERR_ARG, "bad argument");
  const size_t n = (size_t)batch * H * W;
  ub_launch(ub::nchw_to_nhwc4_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      x, batch, C, H, W, reinterpret_cast<uint2*>(y));
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_nchw_to_nhwc4_f32(const float* x, int batch, int C, int H, int W, float* y, void* stream) {
  if (x == nullptr || y == nullptr || C < 1 || C > 4) return fail(UB_ERR_ARG, "bad argument");
  const size_t n = (size_t)batch * H * W;
  ub_launch(ub::nchw_to_nhwc4_f32_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), x, batch, C, H, W,
            reinterpret_cast<float4*>(y));
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

static int preprocess_impl(const uint8_t* src, int batch, int Hs, int Ws, size_t pitch, size_t frame_stride, int H, int W,
                           int swap_rb, const float* mean3, const float* std3, void* y, bool y_f32, uint8_t* resized, void* stream);

int unet_b200_preprocess_u8(const uint8_t* src, int batch, int Hs, int Ws, size_t pitch, size_t frame_stride, int H,
                            int W, int swap_rb, const float* mean3, const float* std3, void* y, uint8_t* resized,
                            void* stream) {
  return preprocess_impl(src, batch, Hs, Ws, pitch, frame_stride, H, W, swap_rb, mean3, std3, y, false, resized, stream);
}

int unet_b200_preprocess_u8_f32(const uint8_t* src, int batch, int Hs, int Ws, size_t pitch, size_t frame_stride, int H,
                                int W, int swap_rb, const float* mean3, const float* std3, float* y, uint8_t* resized,
                                void* stream) {
  return preprocess_impl(src, batch, Hs, Ws, pitch, frame_stride, H, W, swap_rb, mean3, std3, y, true, resized, stream);
}

static int preprocess_impl(const uint8_t* src, int batch, int Hs, int Ws, size_t pitch, size_t frame_stride, int H, int W,
                           int swap_rb, const float* mean3, const float* std3, void* y, bool y_f32, uint8_t* resized, void* stream) {
  if (src == nullptr || y == nullptr || mean3 == nullptr || std3 == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (batch < 1 || Hs < 1 || Ws < 1 || H < 1 || W < 1) return fail(UB_ERR_ARG, "bad size");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  // output rows per tile: as many as keep the staged source rows within 64 KB (several CTAs per SM), at least one
  int rows = ub::PRE_ROWS;
  while (rows > 1 && ub::pre_smem_bytes(Ws, W, H, rows) > 64 * 1024) rows >>= 1;
  const size_t smem = ub::pre_smem_bytes(Ws, W, H, rows);
  if (smem > 200 * 1024) return fail(UB_ERR_ARG, "frames too large for shared-memory staging (%d pixels wide, %d x %d output)", Ws, H, W);
  ub::PreArgs a;
  a.src = src;
  a.pitch = pitch;
  a.frame_stride = frame_stride;
  a.B = batch;
  a.Hs = Hs;
  a.Ws = Ws;
  a.H = H;
  a.W = W;
  a.swap_rb = swap_rb;
  for (int c = 0; c < 3; ++c) {
    a.mean[c] = mean3[c];
    a.inv_std[c] = 1.f / std3[c];
  }
  a.dst = y_f32 ? nullptr : reinterpret_cast<uint2*>(y);
  a.dst_f32 = y_f32 ? reinterpret_cast<float4*>(y) : nullptr;
  a.dst_u8 = resized;
  // same-size frames: cv2.resize is a copy - swap + normalise straight from global memory, four pixels per thread
  if (Hs == H && Ws == W && (W & 3) == 0 && (pitch & 3) == 0 && (frame_stride & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(src) & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    ub_launch(ub::preprocess_copy_u8_kernel, grid_for((size_t)batch * H * (W / 4), 256), 256, 0, static_cast<cudaStream_t>(stream), a);
    UB_CUDA(cudaGetLastError());
    return UB_OK;
  }
  // the copy engine stages the rows (two stages in flight), see preprocess_bulk_u8_kernel
  if (tl_opts->pre_bulk) {
    // tile rows / stages: the staged rows of all stages within ~72 KB (three CTAs per SM). Options pre_rows / pre_stages
    // override the choice for A/B runs.
    int st_b = tl_opts->pre_stages >= 2 && tl_opts->pre_stages <= ub::PREB_MAX_STAGES ? tl_opts->pre_stages : 2;
    int rb = tl_opts->pre_rows >= 1 && tl_opts->pre_rows <= ub::PRE_ROWS ? tl_opts->pre_rows : ub::PRE_ROWS;
    const size_t cap_b = (tl_opts->pre_rows >= 1 ? 200 : 72) * 1024;     // explicit rows: only the hardware limit
    while (rb > 1 && ub::preb_smem_bytes(Ws, W, H, rb, st_b) > cap_b) rb >>= 1;
    const size_t smem_b = ub::preb_smem_bytes(Ws, W, H, rb, st_b);
    if (smem_b <= 200 * 1024) {
      int threads = ((W + 31) / 32) * 32;       // a thread owns output columns: no idle warps for W < 256
      if (threads > ub::PREB_THREADS) threads = ub::PREB_THREADS;
      const int tiles_b = batch * ((H + rb - 1) / rb);
      const bool area2 = Hs == 2 * H && Ws == 2 * W, generic = resized != nullptr || y_f32;
      // instantiation: optional outputs -> the generic form; else {2x2 decimation} x {R<->B swap}
      auto go = [&](auto kernel, int slot) -> int {
        if (smem_b > 48 * 1024) UB_CUDA(ensure_smem(kernel, slot, 200 * 1024));
        int per_sm_b = 0;
        UB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, kernel, threads, smem_b));
        per_sm_b = per_sm_b < 1 ? 1 : per_sm_b;
        const int resident_b = cur_sms() * per_sm_b;
        ub_launch(kernel, tiles_b < resident_b ? tiles_b : resident_b, threads, smem_b, static_cast<cudaStream_t>(stream), a, rb, st_b);
        UB_CUDA(cudaGetLastError());
        return UB_OK;
      };
      if (generic) {
        if (area2) return swap_rb ? go(ub::preprocess_bulk_u8_kernel<true, true, true>, AT_PRE_BULK + 4)
                                  : go(ub::preprocess_bulk_u8_kernel<true, false, true>, AT_PRE_BULK + 5);
        return swap_rb ? go(ub::preprocess_bulk_u8_kernel<false, true, true>, AT_PRE_BULK + 6)
                       : go(ub::preprocess_bulk_u8_kernel<false, false, true>, AT_PRE_BULK + 7);
      }
      if (area2) return swap_rb ? go(ub::preprocess_bulk_u8_kernel<true, true, false>, AT_PRE_BULK + 0)
                                : go(ub::preprocess_bulk_u8_kernel<true, false, false>, AT_PRE_BULK + 1);
      return swap_rb ? go(ub::preprocess_bulk_u8_kernel<false, true, false>, AT_PRE_BULK + 2)
                     : go(ub::preprocess_bulk_u8_kernel<false, false, false>, AT_PRE_BULK + 3);
    }
  }
  if (smem > 48 * 1024) UB_CUDA(ensure_smem(ub::preprocess_u8_kernel, AT_PRE, 200 * 1024));
  // persistent CTAs: as many as are resident at once (shared memory bound), each walks tiles of `rows` output rows
  const int tiles = batch * ((H + rows - 1) / rows);
  int per_sm = 0;   // what the hardware really keeps resident (registers AND shared memory): a second wave of persistent CTAs
                    // would start when the first has finished all its tiles
  UB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ub::preprocess_u8_kernel, ub::PRE_THREADS, smem));
  per_sm = per_sm < 1 ? 1 : per_sm;
  const int resident = cur_sms() * per_sm;
  ub_launch(ub::preprocess_u8_kernel, tiles < resident ? tiles : resident, ub::PRE_THREADS, smem, static_cast<cudaStream_t>(stream), a, rows);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_preprocess_warp_u8(const uint8_t* src, int batch, int Hs, int Ws, size_t pitch, size_t frame_stride,
                                 const double* m_inv9, int Hw, int Ww, int H, int W, int swap_rb, const float* mean3,
                                 const float* std3, void* y, uint8_t* resized, uint8_t* warped, void* stream) {
  if (src == nullptr || m_inv9 == nullptr || mean3 == nullptr || std3 == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (y == nullptr && warped == nullptr) return fail(UB_ERR_ARG, "nothing to compute: y and warped are both null");
  if (batch < 1 || Hs < 1 || Ws < 1 || Hw < 1 || Ww < 1 || H < 1 || W < 1) return fail(UB_ERR_ARG, "bad size");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  ub::WarpPreArgs a;
  a.src = src;
  a.pitch = pitch;
  a.frame_stride = frame_stride;
  a.B = batch;
  a.Hs = Hs;
  a.Ws = Ws;
  a.Hw = Hw;
  a.Ww = Ww;
  a.H = H;
  a.W = W;
  a.swap_rb = swap_rb;
  for (int k = 0; k < 9; ++k) a.m[k] = m_inv9[k];
  for (int c = 0; c < 3; ++c) {
    a.mean[c] = mean3[c];
    a.inv_std[c] = 1.f / std3[c];
  }
  a.dst = reinterpret_cast<uint2*>(y);
  a.dst_u8 = resized;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (y != nullptr) {
    ub_launch(ub::warp_preprocess_u8_kernel, grid_for((size_t)batch * H * W, 256), 256, 0, st, a);
    UB_CUDA(cudaGetLastError());
  }
  if (warped != nullptr) {
    ub_launch(ub::warp_perspective_u8_kernel, grid_for((size_t)batch * Hw * Ww, 256), 256, 0, st, a, warped);
    UB_CUDA(cudaGetLastError());
  }
  return UB_OK;
}

int unet_b200_resize_gray_u8(const uint8_t* src, int batch, int Hs, int Ws, uint8_t* dst, int Hd, int Wd, void* stream) {
  if (src == nullptr || dst == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (batch < 1 || Hs < 1 || Ws < 1 || Hd < 1 || Wd < 1) return fail(UB_ERR_ARG, "bad size");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  const size_t work = (size_t)batch * Hd * ((Wd + 3) / 4);
  ub_launch(ub::resize_gray_u8_kernel, grid_for(work, 256), 256, 0, static_cast<cudaStream_t>(stream), src, batch, Hs, Ws, dst, Hd, Wd);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

size_t unet_b200_infer_staging_bytes(const unet_b200_plan* p, int Hs, int Ws) {
  if (p == nullptr) return 0;
  const size_t npix = (size_t)p->Bc * p->H * p->W, nout = npix * p->out_ch;
  size_t n = align_up((size_t)p->Bc * Hs * Ws * 3, 256);  // frames
  n += align_up(npix * (p->split ? 16 : 8), 256);         // network input: NHWC4 bf16 (fp32-class plan: NHWC4 fp32)
  n += 2 * align_up(nout * 4, 256);                       // logits, probs
  n += align_up(nout, 256);                               // mask
  return n;
}

int unet_b200_infer_u8_host(unet_b200_plan* p, void* staging, const uint8_t* frames, int batch, int Hs, int Ws,
                            int swap_rb, const float* mean3, const float* std3, float threshold, float* logits_h,
                            float* probs_h, uint8_t* mask_h, void* stream) {
  if (p == nullptr || staging == nullptr || frames == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (batch < 1 || batch > p->Bc) return fail(UB_ERR_ARG, "batch %d outside [1,%d]", batch, p->Bc);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t npix_c = (size_t)p->Bc * p->H * p->W, nout_c = npix_c * p->out_ch;
  const size_t npix = (size_t)batch * p->H * p->W * p->out_ch;   // output elements of this call
  uint8_t* s = static_cast<uint8_t*>(staging);
  uint8_t* d_frames = s;
  s += align_up((size_t)p->Bc * Hs * Ws * 3, 256);
  void* d_x = s;
  s += align_up(npix_c * (p->split ? 16 : 8), 256);
  float* d_logits = reinterpret_cast<float*>(s);
  s += align_up(nout_c * 4, 256);
  float* d_probs = reinterpret_cast<float*>(s);
  s += align_up(nout_c * 4, 256);
  uint8_t* d_mask = s;
  const size_t frame_bytes = (size_t)Hs * Ws * 3;
  UB_CUDA(cudaMemcpyAsync(d_frames, frames, frame_bytes * batch, cudaMemcpyHostToDevice, st));
  int rc = preprocess_impl(d_frames, batch, Hs, Ws, (size_t)Ws * 3, frame_bytes, p->H, p->W, swap_rb, mean3, std3, d_x,
                           p->split != 0, nullptr, st);
  if (rc != UB_OK) return rc;
  rc = unet_b200_forward(p, d_x, batch, logits_h ? d_logits : nullptr, probs_h ? d_probs : nullptr,
                         mask_h ? d_mask : nullptr, threshold, st);
  if (rc != UB_OK) return rc;
  if (logits_h) UB_CUDA(cudaMemcpyAsync(logits_h, d_logits, npix * 4, cudaMemcpyDeviceToHost, st));
  if (probs_h) UB_CUDA(cudaMemcpyAsync(probs_h, d_probs, npix * 4, cudaMemcpyDeviceToHost, st));
  if (mask_h) UB_CUDA(cudaMemcpyAsync(mask_h, d_mask, npix, cudaMemcpyDeviceToHost, st));
  UB_CUDA(cudaStreamSynchronize(st));
  return UB_OK;
}

// ---- host-buffer entry point with copy/compute overlap over chunks ----------------------------------------------------
size_t unet_b200_infer_stream_staging_bytes(const unet_b200_plan* p, int Hs, int Ws) {
  if (p == nullptr) return 0;
  const size_t npix = (size_t)p->Bc * p->H * p->W, nout = npix * p->out_ch;
  size_t n = 2 * align_up((size_t)p->Bc * Hs * Ws * 3, 256);  // two frame slots
  n += align_up(npix * (p->split ? 16 : 8), 256);             // NHWC4 input (consumed by the stem before the next chunk's preprocess)
  n += 2 * (2 * align_up(nout * 4, 256) + align_up(nout, 256));  // two output slots: logits, probs, mask
  return n;
}

// Pieces of a pass of n frames: sizes[] in FRONT order, returns their number (0: the plan / the options do not run in pieces).
// host_geometric (default): 16, 32, 64, ... - a piece's input copy takes about half as long as the two front layers on the
// same frames, so the copy of a piece twice the size still hides behind the piece before it: the exposed first copy is 16
// frames, and a pass is four or five pieces instead of eight (every per-piece launch has a tail). The last layer runs the
// same sizes in reverse order (largest first), so the exposed last output copy is 16 frames as well.
// Otherwise: `host_pieces` equal pieces (multiples of 8 frames, at least 16).
static int host_piece_table(const unet_b200_plan* p, int n, int* sizes) {
  if (p->opt.host_pieces <= 1 || !plan_supports_pieces(p) || n < 1) return 0;
  const int max_np = p->opt.host_pieces < unet_b200_plan::MAX_PIECES ? p->opt.host_pieces : unet_b200_plan::MAX_PIECES;
  int np = 0;
  if (p->opt.host_geometric) {
    int left = n, sz = 16;
    while (left > 0) {
      int take = sz < left ? sz : left;
      if (np == max_np - 1 || left - take < take) take = left;  // the last piece takes a remainder smaller than itself
      sizes[np++] = take;
      left -= take;
      sz *= 2;
    }
    return np;
  }
  int piece = ((n + max_np - 1) / max_np + 7) & ~7;
  if (piece < 16) piece = 16;
  for (int b0 = 0; b0 < n; b0 += piece) sizes[np++] = n - b0 < piece ? n - b0 : piece;
  return np;
}
static void host_fill_hooks(PieceHooks* ph, const int* sizes, int np) {
  ph->np = np;
  for (int i = 0, b0 = 0; i < np; ++i) {
    ph->front_b0[i] = b0;
    ph->front_n[i] = sizes[i];
    b0 += sizes[i];
  }
  for (int i = 0, b0 = 0; i < np; ++i) {      // largest first
    ph->tail_b0[i] = b0;
    ph->tail_n[i] = sizes[np - 1 - i];
    b0 += sizes[np - 1 - i];
  }
}

int unet_b200_plan_host_pieces(const unet_b200_plan* p) {
  if (p == nullptr) return 0;
  int sizes[unet_b200_plan::MAX_PIECES];
  return host_piece_table(p, p->Bc, sizes);
}

// Pass schedule of unet_b200_infer_u8_host_stream: frames of pass `it` when `left` of `total` frames remain. With pieces a
// pass takes the plan's capacity; without, the FIRST pass is a quarter of the capacity, so that the bulk of the copies runs
// under the first pass's kernels. Pieces are used unless the source frames are much larger than the network input
// (copy-bound: the H2D copy of a pass takes longer than the two layers its pieces could hide it behind - 480x640 camera
// frames measured 15.9 k frames/s piece-wise, 16.0 k pass-granular; 224x224 frames 16.7 k against 16.2 k).
static bool host_copy_bound(const unet_b200_plan* p, size_t frame_bytes) { return 2 * frame_bytes > 3 * (size_t)p->H * p->W * 3; }
// pieces for frames of this size? Copy-bound sources take them only in the hybrid schedule (option host_hybrid): a SHORT first
// pass whose kernels cover the copies of the rest, and inside every pass the pieces, so that only one piece's copy is exposed.
static bool host_use_pieces(const unet_b200_plan* p, size_t frame_bytes) {
  int sizes[unet_b200_plan::MAX_PIECES];
  if (host_piece_table(p, p->Bc, sizes) <= 0) return false;
  return p->opt.host_hybrid != 0 || !host_copy_bound(p, frame_bytes);
}
static int host_pass_size(const unet_b200_plan* p, int it, int left, int total, bool pieces, size_t frame_bytes) {
  int n = p->Bc;
  if (it == 0 && (!pieces || host_copy_bound(p, frame_bytes))) {
    int first = (p->Bc / 4) & ~7;
    if (first < 8) first = 8;
    if (first * 2 >= total) first = total < p->Bc ? total : p->Bc;
    n = first;
  }
  return n < left ? n : left;
}

// kernels unet_b200_infer_u8_host_stream launches for `total` frames of Hs x Ws pixels (preprocess included)
int unet_b200_infer_stream_launches(const unet_b200_plan* p, int total, int Hs, int Ws) {
  if (p == nullptr || total < 1) return 0;
  const int per_pass = unet_b200_forward_launches(p) + 1;
  int launches = 0;
  const bool pieces = host_use_pieces(p, (size_t)Hs * Ws * 3);
  int it = 0;
  for (int b0 = 0, n = 0; b0 < total; b0 += n, ++it) {
    n = host_pass_size(p, it, total - b0, total, pieces, (size_t)Hs * Ws * 3);
    int sizes[unet_b200_plan::MAX_PIECES];
    const int np = pieces ? host_piece_table(p, n, sizes) : 1;
    launches += per_pass + 4 * (np - 1);   // preprocess, stem, enc0.conv1 and the fused-head conv once per piece
  }
  return launches;
}

int unet_b200_infer_u8_host_stream(unet_b200_plan* p, void* staging, const uint8_t* frames, int total, int Hs, int Ws,
                                   int swap_rb, const float* mean3, const float* std3, float threshold, float* logits_h,
                                   float* probs_h, uint8_t* mask_h, void* stream) {
  if (p == nullptr || staging == nullptr || frames == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (total < 1) return fail(UB_ERR_ARG, "total must be >= 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->s_in == nullptr) {
    UB_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    UB_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      UB_CUDA(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
      UB_CUDA(cudaEventCreateWithFlags(&p->ev_in_free[i], cudaEventDisableTiming));
      UB_CUDA(cudaEventCreateWithFlags(&p->ev_cmp[i], cudaEventDisableTiming));
      UB_CUDA(cudaEventCreateWithFlags(&p->ev_out_free[i], cudaEventDisableTiming));
    }
    UB_CUDA(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
  }
  const size_t npix_c = (size_t)p->Bc * p->H * p->W;
  const size_t frame_bytes = (size_t)Hs * Ws * 3;
  uint8_t* s = static_cast<uint8_t*>(staging);
  uint8_t* d_frames[2];
  for (int i = 0; i < 2; ++i) {
    d_frames[i] = s;
    s += align_up((size_t)p->Bc * frame_bytes, 256);
  }
  void* d_x = s;
  s += align_up(npix_c * (p->split ? 16 : 8), 256);
  float *d_logits[2], *d_probs[2];
  uint8_t* d_mask[2];
  const size_t nout_c = npix_c * p->out_ch;
  for (int i = 0; i < 2; ++i) {
    d_logits[i] = reinterpret_cast<float*>(s);
    s += align_up(nout_c * 4, 256);
    d_probs[i] = reinterpret_cast<float*>(s);
    s += align_up(nout_c * 4, 256);
    d_mask[i] = s;
    s += align_up(nout_c, 256);
  }
  // the side streams start after whatever the caller already queued on `stream`
  UB_CUDA(cudaEventRecord(p->ev_start, st));
  UB_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_start, 0));
  UB_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_start, 0));
  const size_t hw = (size_t)p->H * p->W * p->out_ch;   // output elements per frame
  OptScope opt_scope(&p->opt);
  // ---- piece-wise passes (default): every pass takes up to Bc frames; inside a pass the input copy, the preprocess and the
  // two full-resolution layers at the start run per PIECE, and so do the fused-head conv and the output copies at the end
  // (forward_impl, PieceHooks). Not hidden: the first piece's H2D copy and the last piece's D2H copy - 1/8 of what a
  // pass-granular pipeline leaves exposed - and no pass has to be cut short to get the pipeline going.
  if (host_use_pieces(p, frame_bytes)) {
    struct Ctx {
      unet_b200_plan* p;
      cudaStream_t st;
      const uint8_t* frames_h;      // this pass's frames on the host
      uint8_t* d_frames;            // this pass's frame slot
      void* d_x;
      float *d_logits, *d_probs;
      uint8_t* d_mask;
      float *logits_h, *probs_h;    // this pass's rows of the caller's outputs (or null)
      uint8_t* mask_h;
      size_t frame_bytes, hw;
      int Hs, Ws, swap_rb, n, slot;
      const float *mean3, *std3;
    } c;
    c.p = p;
    c.st = st;
    c.d_x = d_x;
    c.frame_bytes = frame_bytes;
    c.hw = hw;
    c.Hs = Hs;
    c.Ws = Ws;
    c.swap_rb = swap_rb;
    c.mean3 = mean3;
    c.std3 = std3;
    PieceHooks ph;
    ph.ctx = &c;
    ph.before = [](void* v, int i, int b0, int n) -> int {
      Ctx* c = static_cast<Ctx*>(v);
      unet_b200_plan* p = c->p;
      UB_CUDA(cudaStreamWaitEvent(c->st, p->ev_piece_in[i], 0));
      int rc = preprocess_impl(c->d_frames + (size_t)b0 * c->frame_bytes, n, c->Hs, c->Ws, (size_t)c->Ws * 3, c->frame_bytes, p->H,
                               p->W, c->swap_rb, c->mean3, c->std3, static_cast<uint8_t*>(c->d_x) + (size_t)b0 * p->H * p->W * 8,
                               false, nullptr, c->st);
      if (rc != UB_OK) return rc;
      if (b0 + n >= c->n) UB_CUDA(cudaEventRecord(p->ev_in_free[c->slot], c->st));   // the frame slot has been read
      return UB_OK;
    };
    ph.after = [](void* v, int i, int b0, int n) -> int {
      Ctx* c = static_cast<Ctx*>(v);
      unet_b200_plan* p = c->p;
      const size_t o = (size_t)b0 * c->hw, cnt = (size_t)n * c->hw;
      UB_CUDA(cudaEventRecord(p->ev_piece_cmp[i], c->st));
      UB_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_piece_cmp[i], 0));
      if (c->logits_h) UB_CUDA(cudaMemcpyAsync(c->logits_h + o, c->d_logits + o, cnt * 4, cudaMemcpyDeviceToHost, p->s_out));
      if (c->probs_h) UB_CUDA(cudaMemcpyAsync(c->probs_h + o, c->d_probs + o, cnt * 4, cudaMemcpyDeviceToHost, p->s_out));
      if (c->mask_h) UB_CUDA(cudaMemcpyAsync(c->mask_h + o, c->d_mask + o, cnt, cudaMemcpyDeviceToHost, p->s_out));
      if (b0 + n >= c->n) UB_CUDA(cudaEventRecord(p->ev_out_free[c->slot], p->s_out));
      return UB_OK;
    };
    int it = 0;
    for (int b0 = 0, n = 0; b0 < total; b0 += n, ++it) {
      n = host_pass_size(p, it, total - b0, total, true, frame_bytes);
      const int slot = it & 1;
      int sizes[unet_b200_plan::MAX_PIECES];
      const int np = host_piece_table(p, n, sizes);
      host_fill_hooks(&ph, sizes, np);
      for (int i = 0; i < np; ++i) {
        if (p->ev_piece_in[i] == nullptr) {
          UB_CUDA(cudaEventCreateWithFlags(&p->ev_piece_in[i], cudaEventDisableTiming));
          UB_CUDA(cudaEventCreateWithFlags(&p->ev_piece_cmp[i], cudaEventDisableTiming));
        }
      }
      // H2D of every piece of this pass (the slot is free once pass it-2's last preprocess has read it)
      if (it >= 2) UB_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_in_free[slot], 0));
      for (int i = 0; i < np; ++i) {
        const int pb = ph.front_b0[i], pn = ph.front_n[i];
        UB_CUDA(cudaMemcpyAsync(d_frames[slot] + (size_t)pb * frame_bytes, frames + (size_t)(b0 + pb) * frame_bytes,
                                frame_bytes * pn, cudaMemcpyHostToDevice, p->s_in));
        UB_CUDA(cudaEventRecord(p->ev_piece_in[i], p->s_in));
      }
      if (it >= 2) UB_CUDA(cudaStreamWaitEvent(st, p->ev_out_free[slot], 0));  // pass it-2's results have left the output slot
      c.frames_h = frames + (size_t)b0 * frame_bytes;
      c.d_frames = d_frames[slot];
      c.d_logits = d_logits[slot];
      c.d_probs = d_probs[slot];
      c.d_mask = d_mask[slot];
      c.logits_h = logits_h ? logits_h + (size_t)b0 * hw : nullptr;
      c.probs_h = probs_h ? probs_h + (size_t)b0 * hw : nullptr;
      c.mask_h = mask_h ? mask_h + (size_t)b0 * hw : nullptr;
      c.n = n;
      c.slot = slot;
      int rc = forward_impl(p, d_x, n, logits_h ? d_logits[slot] : nullptr, probs_h ? d_probs[slot] : nullptr,
                            mask_h ? d_mask[slot] : nullptr, threshold, st, nullptr, &ph);
      if (rc != UB_OK) return rc;
    }
    UB_CUDA(cudaStreamSynchronize(p->s_out));
    UB_CUDA(cudaStreamSynchronize(st));
    return UB_OK;
  }
  // Pass-granular fallback: the H2D copy of the FIRST pass is the one copy no kernel can hide, so the first pass is a quarter
  // of the plan's capacity (host_pass_size); every later pass is plan-sized and its copy runs under the previous pass's kernels.
  int it = 0;
  for (int b0 = 0, n = 0; b0 < total; b0 += n, ++it) {
    n = host_pass_size(p, it, total - b0, total, false, frame_bytes);
    const int slot = it & 1;
    const size_t npix = (size_t)n * hw;
    // H2D of chunk `it` (runs while chunk it-1 computes); the slot is free once chunk it-2's preprocess has read it
    if (it >= 2) UB_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_in_free[slot], 0));
    UB_CUDA(cudaMemcpyAsync(d_frames[slot], frames + (size_t)b0 * frame_bytes, frame_bytes * n, cudaMemcpyHostToDevice, p->s_in));
    UB_CUDA(cudaEventRecord(p->ev_in[slot], p->s_in));
    // compute
    UB_CUDA(cudaStreamWaitEvent(st, p->ev_in[slot], 0));
    int rc = preprocess_impl(d_frames[slot], n, Hs, Ws, (size_t)Ws * 3, frame_bytes, p->H, p->W, swap_rb, mean3, std3, d_x,
                             p->split != 0, nullptr, st);
    if (rc != UB_OK) return rc;
    UB_CUDA(cudaEventRecord(p->ev_in_free[slot], st));
    if (it >= 2) UB_CUDA(cudaStreamWaitEvent(st, p->ev_out_free[slot], 0));  // chunk it-2's results have left the slot
    rc = unet_b200_forward(p, d_x, n, logits_h ? d_logits[slot] : nullptr, probs_h ? d_probs[slot] : nullptr,
                           mask_h ? d_mask[slot] : nullptr, threshold, st);
    if (rc != UB_OK) return rc;
    UB_CUDA(cudaEventRecord(p->ev_cmp[slot], st));
    // D2H of chunk `it` (runs while chunk it+1 computes)
    UB_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_cmp[slot], 0));
    if (logits_h) UB_CUDA(cudaMemcpyAsync(logits_h + (size_t)b0 * hw, d_logits[slot], npix * 4, cudaMemcpyDeviceToHost, p->s_out));
    if (probs_h) UB_CUDA(cudaMemcpyAsync(probs_h + (size_t)b0 * hw, d_probs[slot], npix * 4, cudaMemcpyDeviceToHost, p->s_out));
    if (mask_h) UB_CUDA(cudaMemcpyAsync(mask_h + (size_t)b0 * hw, d_mask[slot], npix, cudaMemcpyDeviceToHost, p->s_out));
    UB_CUDA(cudaEventRecord(p->ev_out_free[slot], p->s_out));
  }
  UB_CUDA(cudaStreamSynchronize(p->s_out));
  UB_CUDA(cudaStreamSynchronize(st));
  return UB_OK;
}

// ---- single layers ----------------------------------------------------------------------------------
int unet_b200_conv3x3(const void* x0, int C0, const void* x1, int C1, const void* wp, const float* bias, int B, int H,
                      int W, int Cout, int relu, void* y, void* pool, void* stream) {
  if (x0 == nullptr || wp == nullptr || bias == nullptr || y == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (C0 % 64 != 0 || C1 % 64 != 0 || Cout % 64 != 0 || C0 <= 0 || C1 < 0 || Cout <= 0) {
    return fail(UB_ERR_ARG, "C0=%d C1=%d Cout=%d must be multiples of 64", C0, C1, Cout);
  }
  if (C1 > 0 && x1 == nullptr) return fail(UB_ERR_ARG, "x1 is null but C1 > 0");
  if (pool != nullptr && ((H | W) & 1)) return fail(UB_ERR_ARG, "fused pool needs even H and W");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  Layer l;
  rc = conv_layer_setup(l, L_CONV, x0, C0, x1, C1, wp, y, pool, B, H, W, Cout, relu, true);
  if (rc != UB_OK) return rc;
  return conv_layer_launch(l, B, B, bias, y, pool, static_cast<cudaStream_t>(stream));
}

int unet_b200_convT2x2(const void* x, int Cin, const void* wp, const float* bias, int B, int H, int W, int f, void* y,
                       void* stream) {
  if (x == nullptr || wp == nullptr || bias == nullptr || y == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (Cin % 64 != 0 || f % 64 != 0 || Cin <= 0 || f <= 0) return fail(UB_ERR_ARG, "Cin=%d f=%d must be multiples of 64", Cin, f);
  int rc = device_check();
  if (rc != UB_OK) return rc;
  Layer l;
  rc = conv_layer_setup(l, L_CONVT, x, Cin, nullptr, 0, wp, y, nullptr, B, H, W, f, 0, false);
  if (rc != UB_OK) return rc;
  return conv_layer_launch(l, B, B, bias, y, nullptr, static_cast<cudaStream_t>(stream));
}

int unet_b200_pack_conv3x3(const float* w, const float* gamma, const float* beta, const float* mean, const float* var,
                           float eps, int Cout, int Cin, void* wp, float* bias, void* stream) {
  if (w == nullptr || wp == nullptr || bias == nullptr) return fail(UB_ERR_ARG, "null argument");
  ub_launch(ub::pack_conv3x3_kernel, grid_for((size_t)Cout * 9 * Cin, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      w, gamma, beta, mean, var, eps, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(wp), bias);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_pack_convT2x2(const float* w, int Cin, int f, void* wp, void* stream) {
  if (w == nullptr || wp == nullptr) return fail(UB_ERR_ARG, "null argument");
  ub_launch(ub::pack_convT_kernel, grid_for((size_t)4 * f * Cin, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      w, Cin, f, reinterpret_cast<__nv_bfloat16*>(wp));
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_pack_stem(const float* w, const float* gamma, const float* beta, const float* mean, const float* var,
                        float eps, int Cout, int Cin, float* ws, float* bias, void* stream) {
  if (w == nullptr || ws == nullptr || bias == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (Cin < 1 || Cin > 4) return fail(UB_ERR_ARG, "stem Cin must be in [1,4]");
  ub_launch(ub::pack_stem_kernel, grid_for((size_t)36 * Cout, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      w, gamma, beta, mean, var, eps, Cout, Cin, ws, bias, 0, Cout);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_stem_conv(const void* x, const float* ws, const float* bias, int B, int H, int W, int Cin, int Cout,
                        int relu, void* y, void* stream) {
  if (x == nullptr || ws == nullptr || bias == nullptr || y == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (Cout % 32 != 0 || Cout <= 0 || Cout > 256) return fail(UB_ERR_ARG, "stem Cout=%d must be a multiple of 32, <= 256", Cout);
  const int tiles = ((W + 15) / 16) * ((H + 15) / 16) * B;
  const size_t smem = (size_t)(36 * Cout + Cout) * 4 + 18 * 18 * 16;
  ub_launch(ub::stem_conv_kernel<0>, tiles, 256, smem, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const uint2*>(x), ws, bias, B, H, W, Cin, Cout, relu, reinterpret_cast<__nv_bfloat16*>(y));
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_pack_stem_tc(const float* w, const float* gamma, const float* beta, const float* mean, const float* var,
                           float eps, int Cout, int Cin, void* wp, float* bias, void* stream) {
  if (w == nullptr || wp == nullptr || bias == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (Cin < 1 || Cin > 4 || Cout != 64) return fail(UB_ERR_ARG, "tensor-core stem needs Cin in [1,4] and Cout == 64");
  ub_launch(ub::pack_stem_umma_kernel, grid_for((size_t)64 * Cout, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      w, gamma, beta, mean, var, eps, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(wp), bias);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_stem_conv_tc(const void* x, const void* wp, const float* bias, int B, int H, int W, int relu, void* y,
                           void* stream) {
  if (x == nullptr || wp == nullptr || bias == nullptr || y == nullptr) return fail(UB_ERR_ARG, "null argument");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  CUtensorMap mw, mo;
  rc = make_w_map_box(&mw, wp, 64, 64, 64);
  if (rc != UB_OK) return rc;
  const int tw = stem_tile_w();
  rc = make_stem_out_map(&mo, y, B, H, W, tw);
  if (rc != UB_OK) return rc;
  return launch_stem_umma(mw, mo, x, bias, B, H, W, relu, static_cast<cudaStream_t>(stream), nullptr, nullptr, tw);
}

int unet_b200_head(const void* x, const float* w, float bias, size_t npix, int C, float* logits, float* probs,
                   uint8_t* mask, float threshold, void* stream) {
  if (x == nullptr || w == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (C % 8 != 0) return fail(UB_ERR_ARG, "head C=%d must be a multiple of 8", C);
  ub_launch(ub::head_kernel, grid_for(npix * 8, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(x), w, bias, npix, C, logits, probs, mask, threshold);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_maxpool2x2(const void* x, int B, int H, int W, int C, void* y, void* stream) {
  if (x == nullptr || y == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (C % 8 != 0 || ((H | W) & 1)) return fail(UB_ERR_ARG, "maxpool needs C %% 8 == 0 and even H, W");
  const size_t n = (size_t)B * (H / 2) * (W / 2) * (C / 8);
  ub_launch(ub::maxpool2x2_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const uint4*>(x), B, H, W, C / 8, reinterpret_cast<uint4*>(y));
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

// ---- split-precision (fp32-class) single layers: tensors are [B,H,W,2C] = [hi C | lo C] bf16 (aux_kernels.cuh) ----------
namespace {
int split_layer_launch(LayerKind kind, const void* x0, int C0, const void* x1, int C1, const void* wp, const float* bias, int B,
                       int H, int W, int Cout, int relu, void* y, cudaStream_t st) {
  Layer l;
  memset(&l, 0, sizeof(l));
  l.kind = kind;
  l.H = H;
  l.W = W;
  l.C0 = C0;
  l.C1 = C1;
  l.Cout = Cout;
  l.relu = relu;
  pick_tile(H, W, &l.TW, &l.TH, &l.TB, B);
  const int N = (kind == L_CONV) ? Cout : 4 * Cout;
  l.block_n = pick_block_n(N);
  CUtensorMap m0, m1;
  int rc = make_act_map(&m0, x0, B, H, W, 2 * C0, l.TW, l.TH, l.TB);
  if (rc != UB_OK) return rc;
  m1 = m0;
  if (C1 > 0) {
    rc = make_act_map(&m1, x1, B, H, W, 2 * C1, l.TW, l.TH, l.TB);
    if (rc != UB_OK) return rc;
  }
  rc = make_w_map(&l.mW, wp, N, (kind == L_CONV ? 9 : 1) * 3 * (C0 + C1), l.block_n);
  if (rc != UB_OK) return rc;
  rc = make_umma_store_maps(l, y, nullptr, B, 2);
  if (rc != UB_OK) return rc;
  ub::ConvArgs a = conv_args(l, B, B, bias, y, nullptr);
  // K sources: x0[hi|lo], x0[hi], x1[hi|lo], x1[hi]  against  [w_hi | w_hi | w_lo] per source
  a.kc0 = 2 * C0 / 64;
  a.kc1 = C0 / 64;
  a.kc2 = 2 * C1 / 64;
  a.kc3 = C1 / 64;
  a.split = 1;
  const CUtensorMap ma[4] = {m0, m0, m1, m1};
  return launch_conv(l.block_n, ma, l.mW, l.mO, a, st);
}
}  // namespace

int unet_b200_conv3x3_split(const void* x0, int C0, const void* x1, int C1, const void* wp, const float* bias, int B, int H,
                            int W, int Cout, int relu, void* y, void* stream) {
  if (x0 == nullptr || wp == nullptr || bias == nullptr || y == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (C0 % 64 != 0 || C1 % 64 != 0 || Cout % 64 != 0 || C0 <= 0 || C1 < 0 || Cout <= 0) {
    return fail(UB_ERR_ARG, "C0=%d C1=%d Cout=%d must be multiples of 64", C0, C1, Cout);
  }
  if (C1 > 0 && x1 == nullptr) return fail(UB_ERR_ARG, "x1 is null but C1 > 0");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  return split_layer_launch(L_CONV, x0, C0, x1, C1, wp, bias, B, H, W, Cout, relu, y, static_cast<cudaStream_t>(stream));
}

int unet_b200_convT2x2_split(const void* x, int Cin, const void* wp, const float* bias, int B, int H, int W, int f, void* y,
                             void* stream) {
  if (x == nullptr || wp == nullptr || bias == nullptr || y == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (Cin % 64 != 0 || f % 64 != 0 || Cin <= 0 || f <= 0) return fail(UB_ERR_ARG, "Cin=%d f=%d must be multiples of 64", Cin, f);
  int rc = device_check();
  if (rc != UB_OK) return rc;
  return split_layer_launch(L_CONVT, x, Cin, nullptr, 0, wp, bias, B, H, W, f, 0, y, static_cast<cudaStream_t>(stream));
}

int unet_b200_pack_conv3x3_split(const float* w, const float* gamma, const float* beta, const float* mean, const float* var,
                                 float eps, int Cout, int C0, int C1, void* wp, float* bias, void* stream) {
  if (w == nullptr || wp == nullptr || bias == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (C0 <= 0 || C1 < 0) return fail(UB_ERR_ARG, "bad channel split");
  ub_launch(ub::pack_conv3x3_split_kernel, grid_for((size_t)Cout * 27 * (C0 + C1), 256), 256, 0, static_cast<cudaStream_t>(stream), 
      w, gamma, beta, mean, var, eps, Cout, C0, C1, reinterpret_cast<__nv_bfloat16*>(wp), bias);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_pack_convT2x2_split(const float* w, int Cin, int f, void* wp, void* stream) {
  if (w == nullptr || wp == nullptr) return fail(UB_ERR_ARG, "null argument");
  ub_launch(ub::pack_convT_split_kernel, grid_for((size_t)12 * f * Cin, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      w, Cin, f, reinterpret_cast<__nv_bfloat16*>(wp));
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_pack_stem_fp32(const float* w, const float* gamma, const float* beta, const float* mean, const float* var,
                             float eps, int Cout, int Cin, float* ws, float* bias, void* stream) {
  if (w == nullptr || ws == nullptr || bias == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (Cin < 1 || Cin > 4) return fail(UB_ERR_ARG, "stem Cin must be in [1,4]");
  ub_launch(ub::pack_stem_kernel, grid_for((size_t)36 * Cout, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      w, gamma, beta, mean, var, eps, Cout, Cin, ws, bias, 1, Cout);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_stem_conv_split(const float* x_nchw, const float* ws, const float* bias, int B, int H, int W, int Cin, int Cout,
                              int relu, void* y, void* stream) {
  if (x_nchw == nullptr || ws == nullptr || bias == nullptr || y == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (Cin < 1 || Cin > 4) return fail(UB_ERR_ARG, "stem Cin must be in [1,4]");
  if (Cout % 32 != 0 || Cout <= 0 || Cout > 256) return fail(UB_ERR_ARG, "stem Cout=%d must be a multiple of 32, <= 256", Cout);
  const int tiles = ((W + 15) / 16) * ((H + 15) / 16) * B;
  const size_t smem = (size_t)(36 * Cout + Cout) * 4 + 18 * 18 * 16;
  ub_launch(ub::stem_conv_kernel<1>, tiles, 256, smem, static_cast<cudaStream_t>(stream), 
      x_nchw, ws, bias, B, H, W, Cin, Cout, relu, reinterpret_cast<__nv_bfloat16*>(y));
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_maxpool2x2_split(const void* x, int B, int H, int W, int C, void* y, void* stream) {
  if (x == nullptr || y == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (C % 8 != 0 || ((H | W) & 1)) return fail(UB_ERR_ARG, "maxpool needs C %% 8 == 0 and even H, W");
  const size_t n = (size_t)B * (H / 2) * (W / 2) * (C / 8);
  ub_launch(ub::maxpool2x2_split_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const uint4*>(x), B, H, W, C / 8, reinterpret_cast<uint4*>(y));
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

}  // extern "C"

#include "train_capi.cuh"
