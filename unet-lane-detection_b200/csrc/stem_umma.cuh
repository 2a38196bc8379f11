// Stem convolution on tensor cores: Conv2d(Cin<=4 -> 64, 3x3, pad 1) + folded BN + ReLU
// (README.md:1452 with in_channels = 3), NHWC4 bf16 input -> NHWC bf16 output.
//
// K = 9 taps x 4 (padded) channels = 36, far below a 64-deep TMA K block, so the im2col tile is built
// by producer warps instead of TMA: a producer group stages the tile's halo'd [18][10] input patch in shared
// memory with coalesced loads, then each thread assembles one 128-byte row (pixel) of the A operand from it -
// taps 2j,2j+1 form 16-byte chunk j, chunks 0..5 cover K = 48 (columns 36..47 are zero) - and writes
// it in the 128B-swizzled layout the UMMA descriptor expects. Three K=16 MMAs per 128-pixel tile.
// The layer is bound by writing its 128 B/pixel output, so everything else is sized to stay out of
// the way: 2 x 4 producer warps (alternating tiles), 4 operand stages, 4 TMEM accumulator stages,
// 8 epilogue warps with double-buffered staging + TMA store.
#pragma once
#include "epilogue.cuh"
#include "ptx.cuh"

namespace ub {

struct StemArgs {
  int B, H, W;            // images [b0, b0 + B) of the tensors behind x / the output map
  int b0;
  int tiles_w, tiles_h;   // TW-pixel x (128/TW)-row tiles per image
  int relu;
  const uint2* x;         // [B,H,W] x 4 bf16
  const float* bias;      // [64]
  double* stat_sum;       // optional (training): per-channel sum / sum of squares of the bf16 output
  double* stat_sumsq;
};

struct StemCfg {
  static constexpr int N = 64;
  static constexpr int A_STAGES = 4, ACC_STAGES = 4;
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = N * 128;
  static constexpr int THREADS = 17 * 32;  // warp 0 MMA, warps 1..8 epilogue, warps 9..16 producers
  static constexpr int PATCH_BYTES = 2048;     // halo'd input patch of one tile x 8 B ([18][10] = 1440 B or [6][34] = 1632 B), one per producer group
  static constexpr int SMEM_BYTES = A_STAGES * A_BYTES + B_BYTES + 8 * 4096 + 2 * PATCH_BYTES + 512 + 1024;
};

// Weights for the stem GEMM: wp[co][k] bf16, k = tap*4 + ci (zero for ci >= Cin and k >= 36), bias fp32.
__global__ void pack_stem_umma_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ mean,
                                      const float* __restrict__ var, float eps, int Cout, int Cin,
                                      __nv_bfloat16* __restrict__ wp, float* __restrict__ bias) {
  pdl_enter();
  const int total = Cout * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % 64, co = i / 64;
    const int tap = k / 4, ci = k % 4;
    float v = 0.f;
    if (tap < 9 && ci < Cin) {
      float s = 1.f;
      if (gamma != nullptr) s = gamma[co] / sqrtf(var[co] + eps);
      v = w[(static_cast<size_t>(co) * Cin + ci) * 9 + tap] * s;
    }
    wp[i] = __float2bfloat16_rn(v);
  }
  for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < Cout; co += gridDim.x * blockDim.x) {
    float bv = 0.f;
    if (gamma != nullptr) bv = beta[co] - mean[co] * (gamma[co] / sqrtf(var[co] + eps));
    bias[co] = bv;
  }
}

// TW = tile width in pixels (tile = 128/TW rows x TW pixels):
//   8  -> 16 x 8 tiles, the smallest halo ([18][10] patch); a TMEM lane quarter stores a 4-row x 8-pixel box = four 1 KB pieces;
//   32 -> 4 x 32 tiles ([6][34] patch, 13 % more input reads); a lane quarter is ONE image-row segment of 32 pixels = 4 KB
//         contiguous in HBM. The layer is bound by its output writes, so the store pattern is what matters.
template <int TW>
__global__ void __launch_bounds__(StemCfg::THREADS, 1)
stem_umma_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut, const StemArgs a) {
  pdl_launch();
  using Cfg = StemCfg;
  constexpr int TH = 128 / TW;             // tile rows
  constexpr int PW = TW + 2;               // patch pitch in pixels
  constexpr int PATCH = (TH + 2) * PW;     // patch entries (<= 256: at most two per producer thread)
  constexpr int QROWS = 32 / TW;           // tile rows covered by one TMEM lane quarter (TW <= 32)
  static_assert(TW == 8 || TW == 16 || TW == 32, "tile width");
  static_assert(PATCH <= 256 && PATCH * 8 <= Cfg::PATCH_BYTES, "patch size");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smA + Cfg::A_STAGES * Cfg::A_BYTES;
  uint8_t* smS = smB + Cfg::B_BYTES;  // [8 warps][4 KB] private output staging
  uint8_t* smP = smS + 8 * 4096;      // [2 producer groups][18][10] uint2: the tile's halo'd input patch
  uint64_t* bars = reinterpret_cast<uint64_t*>(smP + 2 * Cfg::PATCH_BYTES);
  uint64_t* a_full = bars;                       // [4] producers (128 arrivals) -> MMA
  uint64_t* a_empty = a_full + Cfg::A_STAGES;    // [4] MMA -> producers
  uint64_t* tfull = a_empty + Cfg::A_STAGES;     // [4] MMA -> epilogue
  uint64_t* tempty = tfull + Cfg::ACC_STAGES;    // [4] epilogue (128 arrivals: one warp group per tile) -> MMA
  uint64_t* w_full = tempty + Cfg::ACC_STAGES;   // [1] weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < Cfg::A_STAGES; ++s) {
      mbar_init(&a_full[s], 128);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < Cfg::ACC_STAGES; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, Cfg::ACC_STAGES * Cfg::N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_img = a.tiles_w * a.tiles_h;
  const int total_tiles = tiles_per_img * a.B;

  if (warp == 0) {
    // ------------------------------------------------------------ MMA issuer (+ one-off weight load)
    if (elect_one()) {
      mbar_expect_tx(w_full, Cfg::B_BYTES);
      tma_load_2d(smB, &tmW, w_full, 0, 0);
      mbar_wait(w_full, 0);
      constexpr uint32_t idesc = make_idesc_bf16_f32(128, Cfg::N);
      const uint64_t d_hi = make_sw128_kmajor_desc(0, 1024, 0);
      const uint64_t db = d_hi + (smem_u32(smB) >> 4);
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int s = it & (Cfg::A_STAGES - 1);
        const int acc = it & (Cfg::ACC_STAGES - 1);
        mbar_wait(&tempty[acc], ((it / Cfg::ACC_STAGES) & 1) ^ 1);
        mbar_wait(&a_full[s], (it / Cfg::A_STAGES) & 1);
        tc_fence_after();
        const uint64_t da = d_hi + (smem_u32(smA + s * Cfg::A_BYTES) >> 4);
        const uint32_t d_tmem = tmem_base + acc * Cfg::N;
#pragma unroll
        for (int k = 0; k < 3; ++k) umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
        umma_commit(&a_empty[s]);
        umma_commit(&tfull[acc]);
      }
    }
    __syncwarp();
  } else if (warp >= 9) {
    // ------------------------------------------------------------ im2col producers: group pg builds tiles it = pg, pg+2, ...
    // The 128 threads of a group first stage the tile's [18][10] input patch in shared memory (180 pixels x 8 B: one or two
    // coalesced loads per thread, ONE bounds test each) and then pick their 3x3 neighbourhood from it with nine LDS.64 -
    // the first version did nine bounds-checked global loads per thread and the kernel was instruction-issue bound
    // (profiles/r1_ncu_full_stem_umma_v2.txt: 128 M warp instructions, IPC 1.9). The global loads of the group's NEXT tile
    // are issued before the current tile is assembled, so their latency is off the critical path.
    const int pg = (warp - 9) >> 2;
    const int pt = ((warp - 9) & 3) * 32 + lane;  // thread of the group == row of the A tile == pixel of the tile
    const int tw = pt % TW, th = pt / TW;
    uint2* patch = reinterpret_cast<uint2*>(smP + pg * Cfg::PATCH_BYTES);
    const uint32_t bar_id = 1 + pg;
    // patch entries of this thread: e0 = pt, e1 = pt + 128 (only the first PATCH - 128 threads have one)
    const int r0 = pt / PW, c0 = pt % PW;
    const int r1 = (pt + 128) / PW, c1 = (pt + 128) % PW;
    const bool has1 = pt + 128 < PATCH;
    auto fetch = [&](int t, uint2& v0, uint2& v1) {
      const int br = t / tiles_per_img;
      const int ti = t - br * tiles_per_img;
      const int b = a.b0 + br;
      const int w0 = (ti % a.tiles_w) * TW - 1;
      const int h0 = (ti / a.tiles_w) * TH - 1;
      const uint2* img = a.x + static_cast<size_t>(b) * a.H * a.W;
      v0 = make_uint2(0u, 0u);
      v1 = make_uint2(0u, 0u);
      int hh = h0 + r0, ww = w0 + c0;
      if (static_cast<unsigned>(hh) < static_cast<unsigned>(a.H) && static_cast<unsigned>(ww) < static_cast<unsigned>(a.W)) {
        v0 = __ldg(img + hh * a.W + ww);
      }
      hh = h0 + r1;
      ww = w0 + c1;
      if (has1 && static_cast<unsigned>(hh) < static_cast<unsigned>(a.H) && static_cast<unsigned>(ww) < static_cast<unsigned>(a.W)) {
        v1 = __ldg(img + hh * a.W + ww);
      }
    };
    const int t_step = 2 * static_cast<int>(gridDim.x);
    int t = static_cast<int>(blockIdx.x) + pg * static_cast<int>(gridDim.x);
    uint2 v0 = make_uint2(0u, 0u), v1 = make_uint2(0u, 0u);
    if (t < total_tiles) fetch(t, v0, v1);
    for (int it = pg; t < total_tiles; t += t_step, it += 2) {
      const int s = it & (Cfg::A_STAGES - 1);
      patch[pt] = v0;
      if (has1) patch[pt + 128] = v1;
      named_bar_sync(bar_id, 128);                      // patch complete
      if (t + t_step < total_tiles) fetch(t + t_step, v0, v1);   // next tile's loads in flight from here on
      uint2 tap[9];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) tap[r * 3 + c] = patch[(th + r) * PW + tw + c];
      }
      named_bar_sync(bar_id, 128);                      // everyone has read the patch: it may be overwritten
      mbar_wait_parked(&a_empty[s], ((it / Cfg::A_STAGES) & 1) ^ 1, 2000);
      const uint32_t row = smem_u32(smA + s * Cfg::A_BYTES + pt * 128);
      const int sw = pt & 7;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        st_shared_v4(row + ((j ^ sw) << 4), tap[2 * j].x, tap[2 * j].y, tap[2 * j + 1].x, tap[2 * j + 1].y);
      }
      st_shared_v4(row + ((4 ^ sw) << 4), tap[8].x, tap[8].y, 0u, 0u);
      st_shared_v4(row + ((5 ^ sw) << 4), 0u, 0u, 0u, 0u);
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      mbar_arrive(&a_full[s]);
    }
  } else {
    // ------------------------------------------------------------ epilogue: warps 1..8, warp-private units (epilogue.cuh);
    // the two warps of a TMEM lane quarter alternate tiles (group cg takes tiles with it % 2 == cg)
    const int q = warp & 3;
    const int cg = (warp - 1) >> 2;
    uint8_t* stg = smS + (warp - 1) * 4096;
    float st0[4] = {0.f, 0.f, 0.f, 0.f};  // fused batch statistics (epilogue.cuh)
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      if ((it & 1) != cg) continue;
      const int acc = it & (Cfg::ACC_STAGES - 1);
      const int br = t / tiles_per_img;
      const int ti = t - br * tiles_per_img;
      const int b = a.b0 + br;
      const int w0 = (ti % a.tiles_w) * TW;
      const int h0 = (ti / a.tiles_w) * TH;
      mbar_wait(&tfull[acc], (it / Cfg::ACC_STAGES) & 1);
      tc_fence_after();
      uint32_t p[32];
      epi_load_unit(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::N, a.bias, a.relu, p);
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (lane == 0) bulk_wait_group_read<0>();
      __syncwarp();
      epi_stage_row(stg, lane, p);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&tmOut, stg, 0, w0, h0 + QROWS * q, b);   // box (64 ch, TW pixels, QROWS rows, 1 image)
        bulk_commit_group();
      }
      if (a.stat_sum != nullptr) {
        const bool valid = (w0 + (lane % TW) < a.W) && (h0 + QROWS * q + (lane / TW) < a.H);
        epi_stats_accumulate(stg, lane, __ballot_sync(0xffffffffu, valid), st0);
      }
    }
    if (a.stat_sum != nullptr) epi_stats_flush(a.stat_sum, a.stat_sumsq, 0, lane, st0);
    if (lane == 0) bulk_wait_group_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::ACC_STAGES * Cfg::N);
  }
}

}  // namespace ub
