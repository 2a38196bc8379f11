// Weight gradient of the 3x3 convs with Cout == 64 (the 224^2 level: enc0.conv1, dec3.conv0, dec3.conv1 - README.md:1451-1458
// under loss.backward(), README.md:2078), all nine taps from ONE halo'd activation patch.
//
//   dW[co][tap][ci] += sum over pixels p of  dy[p][co] * x[p + shift(tap)][ci]
//
// wgrad_umma.cuh gives every (tap pair, K slice) its own CTA, so the pixel boxes of x and dy are pulled through L2 five
// (dy) and ten (x) times: 48 KB of operands per eight N = 64 MMAs, 125 B per cycle and SM - the three Cout = 64 layers sat at
// 45 % of the bf16 peak, L2-bound (profiles/r1_ncu_full_train_wgrad64.txt), and because the weight-gradient stream shares
// the SMs with the backward's dgrad GEMMs they held up the critical path at exactly the two largest levels.
// Here a CTA owns ALL nine taps of one 64-channel block of x over its slice of the 16 x 8 pixel tiles:
//   * the x tile arrives once as the halo'd patch [18][10][64 ch] of conv_halo.cuh (23 KB); tap (r,s) is the MN-major UMMA
//     descriptor  start = patch + ((2j + r)*10 + s)*128 B, SBO = 10*128 B (the next pixel row), exactly the shifted-descriptor
//     trick of the forward kernel (the 128B swizzle is a function of the shared-memory address, tools/halo_probe.cu) applied
//     to the MN-major operand form of wgrad_umma.cuh (channels contiguous, pixels = K);
//   * M = 128 rows are two taps x 64 input channels: the second 64-row block is the first one displaced by
//     LBO = (tap distance in patch rows)*128 B (128 B for horizontal neighbours, 1024 B from tap 2 to tap 3), so five MMAs
//     per 16-pixel K step cover the nine taps (the tenth half-tile repeats tap 8 and is dropped);
//   * dy arrives once as the [16 x 8 pixels][64 ch] box (16 KB), N = 64;
//   * five fp32 accumulators of 64 columns live in TMEM (320 of 512 columns) for the whole kernel; at the end four warps add
//     them into the PyTorch-layout gradient with fp32 atomics, as wgrad_umma.cuh does.
// 39 KB of operands per forty MMAs: 20 B per cycle and SM, so the kernel is bound by the N = 64 MMA rate (48 cycles against
// a math floor of 32: the A operand's shared-memory reads) instead of L2.
#pragma once
#include "ptx.cuh"
#include "wgrad_umma.cuh"

namespace ub {

struct WgradHaloArgs {
  int B, H, W;
  int tiles_w, tiles_h;     // 8-pixel / 16-row tiles per image
  int ncb;                  // 64-channel blocks of x = cat(x0, x1); CTA c works on block c % ncb
  int cb_split;             // blocks [0, cb_split) come from x0, the rest from x1
  int kslices;              // CTAs per channel block: CTA c takes tiles c / ncb, + kslices, ...
  int stages;
  int lc0, lc1, lcout;      // logical channels of x0 / x1 / dy (physical channels beyond them are exact zeros, no gradient slot)
  long long s_co, s_ci, s_tap;   // dW index = co*s_co + ci*s_ci + tap*s_tap
  GradRoute route;
  long long off;
};

struct WgradHaloCfg {
  static constexpr int A_BYTES = 18 * 10 * 128;      // 23040: the halo'd x patch
  static constexpr int A_PITCH = 23552;              // next multiple of 1024
  static constexpr int D_BYTES = 128 * 128;          // the dy box
  static constexpr int STAGE_BYTES = A_PITCH + D_BYTES;
  static constexpr int STAGES = 5;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
  static constexpr int TMEM_COLS = 512;              // five 64-column accumulators
  static constexpr int THREADS = 192;                // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..5 epilogue
};

__global__ void __launch_bounds__(WgradHaloCfg::THREADS, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX0 /* halo box (64,10,18,1) */,
                  const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmD /* box (64,8,16,1) */,
                  const WgradHaloArgs a) {
  pdl_enter();
  using Cfg = WgradHaloCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* done = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int cb = blockIdx.x % a.ncb;
  const int ks = blockIdx.x / a.ncb;
  const int tiles_per_img = a.tiles_w * a.tiles_h;
  const int total_tiles = tiles_per_img * a.B;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const CUtensorMap* mx = cb < a.cb_split ? &tmX0 : &tmX1;
      const int c0 = (cb < a.cb_split ? cb : cb - a.cb_split) * 64;
      for (int t = ks; t < total_tiles; t += a.kslices) {
        const int b = t / tiles_per_img;
        const int ti = t - b * tiles_per_img;
        const int w0 = (ti % a.tiles_w) * 8;
        const int h0 = (ti / a.tiles_w) * 16;
        mbar_wait_parked(&empty[stage], phase ^ 1);
        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::D_BYTES);
        tma_load_4d(sA, mx, &full[stage], c0, w0 - 1, h0 - 1, b);          // zero fill outside the image = the conv's padding
        tma_load_4d(sA + Cfg::A_PITCH, &tmD, &full[stage], 0, w0, h0, b);  // rows past the image are zero: they add nothing
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(128, 64) | (1u << 15) | (1u << 16);   // both operands MN-major
      // tap pairs (0,1) (2,3) (4,5) (6,7) (8,8): patch row of the first tap, distance to the second one (rows of 128 B)
      constexpr int first_row[5] = {0, 2, 11, 20, 22};
      constexpr int lbo_rows[5] = {1, 8, 1, 1, 0};
      uint64_t da_hi[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) da_hi[u] = make_sw128_mnmajor_desc(first_row[u] * 128, lbo_rows[u] * 128, 1280);
      const uint64_t db_hi = make_sw128_mnmajor_desc(0, 0, 1024);
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int t = ks; t < total_tiles; t += a.kslices) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES) & 0x3FFFFu;
        const uint64_t a0 = sA >> 4, b0 = (sA + Cfg::A_PITCH) >> 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {       // 16 pixels per MMA = two tile rows: 20 patch rows / 16 dy rows further on
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            umma_f16(tmem_base + u * 64, da_hi[u] + a0 + j * 160, db_hi + b0 + j * 128, idesc, (!first || j > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty[stage]);
        first = false;
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(done);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue: 4 warps, fp32 atomic accumulation
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const bool any = ks < total_tiles;     // (a CTA without tiles never commits MMAs: nothing to add)
    if (any) {
      mbar_wait(done, 0);
      tc_fence_after();
      // physical channel of cat(x0, x1) -> logical input channel of the weight tensor (or none: a zero-extended channel)
      int ci = cb * 64 + (m & 63);
      bool ci_live;
      if (cb < a.cb_split) {
        ci_live = ci < a.lc0;
      } else {
        const int c1 = ci - a.cb_split * 64;
        ci_live = c1 < a.lc1;
        ci = a.lc0 + c1;
      }
#pragma unroll 1
      for (int u = 0; u < 5; ++u) {
        const int tap = 2 * u + (m >> 6);
        const bool live = ci_live && tap < 9;
        const long long base = a.off + static_cast<long long>(tap) * a.s_tap + static_cast<long long>(ci) * a.s_ci;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + u * 64 + c * 32, v);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int co = c * 32 + j;
              if (co < a.lcout) grad_add(a.route, base + static_cast<long long>(co) * a.s_co, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace ub
