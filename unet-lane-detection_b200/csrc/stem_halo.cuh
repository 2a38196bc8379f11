// Stem fused into the second conv of the first encoder block (README.md:1451-1458 with in_channels <= 4):
//
//   x (NHWC4 bf16) -> [ Conv3x3(Cin -> 64) + BN + ReLU ] -> Conv3x3(64 -> 64) + BN + ReLU (+ 2x2 max-pool)
//
// as ONE kernel: the stem's output - 6.4 MB per 224x224 frame, written once by stem_umma_kernel and read once by
// conv_halo2_kernel<64>, the stem being bound by exactly that write - never goes to HBM. The first conv's output has one
// consumer and is not a skip connection, so nothing else needs it.
//
// The kernel is conv_halo2_kernel<64> (conv_halo.cuh: CTA pair, 16 x 8 pixel tiles, halo'd [18][10][64 ch] patch, shifted
// descriptors, resident weights) with a different source for the patch: instead of a TMA load, the 180 patch pixels are
// COMPUTED from the tile's [20][12] input pixels by a small tensor-core GEMM of their own:
//   * producer warps (3 groups x 4 warps, taking tiles in turn) stage the input pixels in shared memory and assemble the
//     im2col rows of the 180 patch pixels (K = 9 taps x 4 channels, padded to 48; two row blocks of 128) exactly as
//     stem_umma.cuh does for its output pixels;
//   * the MMA thread issues 2 x 3 tcgen05.mma.cta_group::2 (M = 256: both CTAs' row blocks, N = 64, K = 16) into a TMEM
//     region of its own (3 stages x 128 columns next to the 2 x 64 accumulator columns of the main conv);
//   * the same producer warps read the result back (tcgen05.ld), add the stem's bias, apply ReLU, ZERO the rows that lie
//     outside the image (they are the second conv's padding, not stem outputs), round to bf16 and write the patch into the
//     128B-swizzled layout the shifted descriptors of the main conv expect.
// The stem GEMM recomputes the halo (180 instead of 128 pixels per tile) on M = 256 blocks that are 70 % full: six extra
// N = 64 MMAs per 36 of the main conv. The values are the ones the two-kernel path produces (same MMAs, same rounding
// points), so the outputs are bit-identical (tests/test_gpu_layers.py).
//
// Barriers (pair protocol of ptx.cuh: "full" barriers live in the leader and count arrivals of both CTAs, commits are
// multicast to both):  w_full   weights landed (TMA, both CTAs' halves)
//   i_full[g]  im2col tile of producer group g complete in both CTAs (8 warp arrivals)      producers -> MMA
//   i_empty[g] the stem MMAs have read it (commit)                                          MMA -> producers
//   s_full[g]  stem result in TMEM (commit)                                                 MMA -> producers
//   s_empty[g] stem TMEM stage read out (2 x 128 thread arrivals)                           producers -> MMA
//   a_full[g]  patch stage g written in both CTAs (8 warp arrivals)                         producers -> MMA
//   a_empty[g] the main MMAs have read it (commit)                                          MMA -> producers
//   tfull / tempty[2] accumulator stages of the main conv, as in conv_halo.cuh
#pragma once
#include "conv_halo.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace ub {

struct StemHaloArgs {
  int B, H, W;              // images [b0, b0 + B)
  int b0;
  int tiles_w, tiles_h;     // 8-pixel / 16-row tiles per image
  const uint2* x;           // network input [.,H,W] x 4 bf16
  const float* stem_bias;   // [64] folded BN of the stem
  const float* bias;        // [64] folded BN of the second conv
  int relu;
  __nv_bfloat16* pool_out;  // optional: [.,H/2,W/2,64] = maxpool2x2(out)
};

struct StemHaloCfg {
  static constexpr int N = 64;
  static constexpr int G = 2;                           // producer groups = im2col / patch / stem-TMEM stages
  static constexpr int B_HALF = 32 * 128;               // one tap's weight rows held by ONE CTA (half of N = 64)
  static constexpr int W_BYTES = 9 * B_HALF;            // resident main weights
  static constexpr int WS_BYTES = B_HALF;               // stem weights (this CTA's 32 output channels x 64 K)
  // im2col tile: 180 rows of 128 B in two row blocks (rows 0..127 at +0, rows 128..179 at +16 KB). The second block's MMA
  // reads 128 rows: the 76 rows past the tile run into whatever follows in shared memory - rows of an MMA are independent
  // and those results are never read, so a stage only OWNS 23 KB.
  static constexpr int I_PITCH = HaloCfg::A_STAGE_PITCH;
  static constexpr int A_PITCH = HaloCfg::A_STAGE_PITCH;  // patch stage (180 rows x 128 B, padded to 23 KB)
  static constexpr int IN_PATCH = 20 * 12;              // input pixels of a tile (two halos)
  static constexpr int IN_BYTES = 2048;                 // their staging buffer (240 x 8 B), one per producer group
  static constexpr int STG_BYTES = 8 * 4096;
  static constexpr int BAR_BYTES = 256;
  // im2col stages per group: 1 = a group assembles a tile, waits for its stem result and converts it; 2 = it assembles its NEXT
  // tile before converting the current one. Measured on B200 (same box, 256 frames): 1.46-1.58 ms with 1, 1.89 ms with 2 -
  // the conversion is what the main MMAs wait for, and anything in front of it delays them.
  static constexpr int LOOKAHEAD = 0;
  static constexpr int IS = (1 + LOOKAHEAD) * G;
  static constexpr int SMEM_BYTES = IS * I_PITCH + G * A_PITCH + W_BYTES + WS_BYTES + STG_BYTES + G * IN_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;                 // 2 x 64 main accumulators + G x 128 stem results
  static constexpr int THREADS = (10 + 4 * G) * 32;     // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue, then G x 4 producer warps
  static_assert(SMEM_BYTES <= 232448, "shared memory");
  static_assert(128 + G * 128 <= TMEM_COLS, "tensor memory");
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(StemHaloCfg::THREADS, 1)
stem_halo2_kernel(const __grid_constant__ CUtensorMap tmW /* main weights, box 32 rows */,
                  const __grid_constant__ CUtensorMap tmWs /* stem weights [64][64], box 32 rows */,
                  const __grid_constant__ CUtensorMap tmOut, const StemHaloArgs a) {
  pdl_enter();
  using Cfg = StemHaloCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smI = smem;                                  // [IS] im2col tiles (1024-aligned; a tile's second row block reads past it)
  uint8_t* smA = smI + Cfg::IS * Cfg::I_PITCH;          // [G] patch stages
  uint8_t* smW = smA + Cfg::G * Cfg::A_PITCH;           // 9 taps x this CTA's half
  uint8_t* smWs = smW + Cfg::W_BYTES;
  uint8_t* smS = smWs + Cfg::WS_BYTES;                  // [8 warps][4 KB] output staging
  uint8_t* smP = smS + Cfg::STG_BYTES;                  // [G groups] input pixels
  uint64_t* bars = reinterpret_cast<uint64_t*>(smP + Cfg::G * Cfg::IN_BYTES);
  uint64_t* w_full = bars;
  uint64_t* i_full = bars + 1;                          // [IS]
  uint64_t* i_empty = i_full + Cfg::IS;                 // [IS]
  uint64_t* s_full = i_empty + Cfg::IS;                 // [G] each from here
  uint64_t* s_empty = s_full + Cfg::G;
  uint64_t* a_full = s_empty + Cfg::G;
  uint64_t* a_empty = a_full + Cfg::G;
  uint64_t* tfull = a_empty + Cfg::G;                   // [2]
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmWs);
    tma_prefetch_desc(&tmOut);
    mbar_init(w_full, 1);
    for (int s = 0; s < Cfg::IS; ++s) {
      mbar_init(&i_full[s], 8);        // 2 CTAs x 4 producer warps
      mbar_init(&i_empty[s], 1);
    }
    for (int s = 0; s < Cfg::G; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 256);     // 2 CTAs x 128 producer threads
      mbar_init(&a_full[s], 8);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 2 * 128);  // the epilogue threads of one warp group in both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_stem = tmem_base + 128;            // + g * 128 + row block * 64

  const int tiles_per_img = a.tiles_w * a.tiles_h;
  const int total_tiles = tiles_per_img * a.B;
  const int pairs = gridDim.x / 2;
  const int pair = blockIdx.x / 2;
  const int t_first = 2 * pair + static_cast<int>(rank);   // iteration i covers tiles 2*(pair + i*pairs) + rank
  const int t_step = 2 * pairs;

  if (warp == 0) {
    // ------------------------------------------------------------ weights: loaded once (both CTAs; the leader posts expect_tx)
    if (elect_one()) {
      if (leader) mbar_expect_tx(w_full, 2 * (Cfg::W_BYTES + Cfg::WS_BYTES));
      for (int tap = 0; tap < 9; ++tap) {
        tma2_load_2d(smW + tap * Cfg::B_HALF, &tmW, w_full, tap * 64, static_cast<int>(rank) * 32);
      }
      tma2_load_2d(smWs, &tmWs, w_full, 0, static_cast<int>(rank) * 32);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: one elected thread of the LEADER CTA
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(256, Cfg::N);
      const uint64_t da_hi = make_sw128_kmajor_desc(0, 1280, 0);   // patch: SBO = one patch row (10 pixels)
      const uint64_t dk_hi = make_sw128_kmajor_desc(0, 1024, 0);
      const uint64_t db_main = dk_hi + ((smem_u32(smW) & 0x3FFFFu) >> 4);
      const uint64_t db_stem = dk_hi + ((smem_u32(smWs) & 0x3FFFFu) >> 4);
      int n_it = 0;
      for (int t = t_first; t < total_tiles; t += t_step) ++n_it;   // (the leader's tile exists in every iteration)
      mbar_wait(w_full, 0);
      tc_fence_after();
      auto stem_issue = [&](int j) {
        const int g = j % Cfg::G;
        const uint32_t ph = (j / Cfg::G) & 1;
        const int is = j % Cfg::IS;                        // == g + G * ((j / G) & 1)
        mbar_wait(&s_empty[g], ph ^ 1);
        mbar_wait(&i_full[is], (j / Cfg::IS) & 1);
        tc_fence_after();
        const uint64_t di = dk_hi + ((smem_u32(smI + is * Cfg::I_PITCH) & 0x3FFFFu) >> 4);
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            umma2_f16(tmem_stem + g * 128 + rb * 64, di + rb * (16384 >> 4) + 2 * k, db_stem + 2 * k, idesc, k != 0);
          }
        }
        umma2_commit_both(&i_empty[is]);
        umma2_commit_both(&s_full[g]);
      };
      if (n_it > 0) stem_issue(0);
      for (int j = 0; j < n_it; ++j) {
        if (j + 1 < n_it) stem_issue(j + 1);
        const int g = j % Cfg::G;
        const int acc = j & 1;
        mbar_wait(&tempty[acc], ((j >> 1) & 1) ^ 1);
        mbar_wait(&a_full[g], (j / Cfg::G) & 1);
        tc_fence_after();
        const uint64_t da0 = da_hi + ((smem_u32(smA + g * Cfg::A_PITCH) & 0x3FFFFu) >> 4);
        const uint32_t d_tmem = tmem_base + acc * Cfg::N;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma2_f16(d_tmem, da0 + (((tap / 3) * 10 + (tap % 3)) * 8 + k * 2), db_main + (tap * (Cfg::B_HALF >> 4) + k * 2), idesc,
                      (tap | k) != 0);
          }
        }
        umma2_commit_both(&a_empty[g]);
        umma2_commit_both(&tfull[acc]);
      }
    }
    __syncwarp();
  } else if (warp >= 10) {
    // ------------------------------------------------------------ producers: group g builds the patches of iterations g, g+G, ...
    const int g = (warp - 10) >> 2;
    const int pt = ((warp - 10) & 3) * 32 + lane;   // thread of the group
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const uint32_t inp = smem_u32(smP + g * Cfg::IN_BYTES);
    uint8_t* patch = smA + g * Cfg::A_PITCH;
    const uint32_t bar_id = 1 + g;
    // input pixels of this thread: entries pt and pt + 128 of the [20][12] window
    const int r0 = pt / 12, c0 = pt % 12;
    const int r1 = (pt + 128) / 12, c1 = (pt + 128) % 12;
    const bool has1 = pt + 128 < Cfg::IN_PATCH;
    // im2col rows this thread assembles: patch pixels pt and pt + 128 (< 180); offset of their first tap in the input window
    const bool row1 = pt + 128 < 180;
    const uint32_t in0 = inp + ((pt / 10) * 12 + pt % 10) * 8;
    const uint32_t in1 = inp + (((pt + 128) / 10) * 12 + (pt + 128) % 10) * 8;
    // patch rows this thread converts: TMEM lane q*32 + lane of row block 0, and of row block 1 (rows 128..179: quarters 0, 1)
    const int crow0 = q * 32 + lane, crow1 = 128 + q * 32 + lane;
    const int cpr0 = crow0 / 10, cpc0 = crow0 % 10, cpr1 = crow1 / 10, cpc1 = crow1 % 10;
    auto fetch = [&](int t, uint2& v0, uint2& v1) {
      v0 = make_uint2(0u, 0u);
      v1 = make_uint2(0u, 0u);
      if (t >= total_tiles) return;                  // the pair's missing odd tile: zeros (its outputs are never stored)
      const int br = t / tiles_per_img;
      const int ti = t - br * tiles_per_img;
      const int w0 = (ti % a.tiles_w) * 8 - 2;
      const int h0 = (ti / a.tiles_w) * 16 - 2;
      const uint2* img = a.x + static_cast<size_t>(a.b0 + br) * a.H * a.W;
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
    auto lds64 = [](uint32_t addr, uint32_t& x, uint32_t& y) {
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(addr));
    };
    auto build_row = [&](uint8_t* imc, int row, uint32_t src) {
      uint32_t tx[9], ty[9];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) lds64(src + (r * 12 + c) * 8, tx[r * 3 + c], ty[r * 3 + c]);
      }
      const uint32_t dst = smem_u32(imc + row * 128);
      const int sw = row & 7;
#pragma unroll
      for (int j = 0; j < 4; ++j) st_shared_v4(dst + ((j ^ sw) << 4), tx[2 * j], ty[2 * j], tx[2 * j + 1], ty[2 * j + 1]);
      st_shared_v4(dst + ((4 ^ sw) << 4), tx[8], ty[8], 0u, 0u);
      st_shared_v4(dst + ((5 ^ sw) << 4), 0u, 0u, 0u, 0u);
    };
    // one row block of the stem result -> patch rows: + bias, ReLU, bf16; rows outside the image are the second conv's zero
    // padding. `border` is warp-uniform: interior tiles (80 % of them) skip the per-element selects.
    auto convert = [&](int rb, int row, bool inside, bool border) {
      uint32_t p[32];
      epi_load_unit(tmem_stem + (static_cast<uint32_t>(q * 32) << 16) + g * 128 + rb * 64, a.stem_bias, 1, p);
      if (border && !inside) {
#pragma unroll
        for (int i = 0; i < 32; ++i) p[i] = 0u;
      }
      if (row < 180) epi_stage_row(patch, row, p);
    };
    const int t_stepg = Cfg::G * t_step;
    // stage the input window held in (v0, v1), start the loads of the tile after it, assemble the im2col tile of iteration j
    // (tile index tn) into im2col stage j % IS and publish it
    uint2 v0 = make_uint2(0u, 0u), v1 = make_uint2(0u, 0u);
    auto build = [&](int j, int tn) {
      const int is = j % Cfg::IS;
      asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(inp + pt * 8), "r"(v0.x), "r"(v0.y) : "memory");
      if (has1) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(inp + (pt + 128) * 8), "r"(v1.x), "r"(v1.y) : "memory");
      named_bar_sync(bar_id, 128);                       // input window complete
      fetch(tn + t_stepg - static_cast<int>(rank) < total_tiles ? tn + t_stepg : total_tiles, v0, v1);   // next tile's loads in flight
      mbar_wait_parked(&i_empty[is], ((j / Cfg::IS) & 1) ^ 1, 2000);   // the stem MMAs that read this stage last have completed
      uint8_t* imc = smI + is * Cfg::I_PITCH;
      build_row(imc, pt, in0);
      if (row1) build_row(imc, pt + 128, in1);
      fence_proxy_async();                               // generic-proxy writes -> visible to the tensor core's async-proxy reads
      named_bar_sync(bar_id, 128);                       // (also: everyone has read the input window, it may be overwritten)
      if (lane == 0) mbar_arrive_leader(&i_full[is]);
    };
    int t = t_first + g * t_step;
    fetch(t, v0, v1);
    // (iteration j runs while the LEADER's tile exists: t - rank < total_tiles)
    if (Cfg::LOOKAHEAD && t - static_cast<int>(rank) < total_tiles) build(g, t);
    for (int j = g; t - static_cast<int>(rank) < total_tiles; t += t_stepg, j += Cfg::G) {
      const uint32_t ph = (j / Cfg::G) & 1;
      if (Cfg::LOOKAHEAD) {
        if (t + t_stepg - static_cast<int>(rank) < total_tiles) build(j + Cfg::G, t + t_stepg);
      } else {
        build(j, t);
      }
      // ---- stem result -> patch
      const bool live = t < total_tiles;
      const int br = live ? t / tiles_per_img : 0;
      const int ti = live ? t - br * tiles_per_img : 0;
      const int w0 = (ti % a.tiles_w) * 8 - 1;            // image coordinates of patch pixel (0, 0)
      const int h0 = (ti / a.tiles_w) * 16 - 1;
      const bool border = w0 < 0 || h0 < 0 || w0 + 10 > a.W || h0 + 18 > a.H;
      const bool in_a = static_cast<unsigned>(h0 + cpr0) < static_cast<unsigned>(a.H) &&
                        static_cast<unsigned>(w0 + cpc0) < static_cast<unsigned>(a.W);
      const bool in_b = static_cast<unsigned>(h0 + cpr1) < static_cast<unsigned>(a.H) &&
                        static_cast<unsigned>(w0 + cpc1) < static_cast<unsigned>(a.W);
      mbar_wait(&s_full[g], ph);
      tc_fence_after();
      mbar_wait_parked(&a_empty[g], ph ^ 1, 2000);       // the main MMAs of this group's previous tile have read the patch stage
      convert(0, crow0, in_a, border);
      if (q < 2) convert(1, crow1, in_b, border);        // rows 192.. do not exist (warp-uniform)
      tc_fence_before();
      mbar_arrive_leader(&s_empty[g]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&a_full[g]);
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9), as conv_halo.cuh HEPI_STORE (+ pool)
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int tw = m & 7;
    const int th = m >> 3;
    uint8_t* stg = smS + (warp - 2) * 4096;
    const bool pool_writer = ((tw | th) & 1) == 0;
    int it = 0;
    for (int t = t_first; t - static_cast<int>(rank) < total_tiles; t += t_step, ++it) {
      const int acc = it & 1;
      if (acc != cg) continue;
      const bool live = t < total_tiles;
      const int br = live ? t / tiles_per_img : 0;
      const int ti = live ? t - br * tiles_per_img : 0;
      const int b = a.b0 + br;
      const int w0 = (ti % a.tiles_w) * 8;
      const int h0 = (ti / a.tiles_w) * 16;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      uint32_t p[32];
      epi_load_unit(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::N, a.bias, a.relu, p);
      tc_fence_before();
      mbar_arrive_leader(&tempty[acc]);
      if (!live) continue;   // warp-uniform
      const int w = w0 + tw, h = h0 + th;
      const bool valid = (w < a.W) && (h < a.H);
      if (lane == 0) bulk_wait_group_read<0>();
      __syncwarp();
      epi_stage_row(stg, lane, p);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&tmOut, stg, 0, w0, h0 + 4 * q, b);
        bulk_commit_group();
      }
      if (a.pool_out != nullptr) {
        epi_pool2x2(p, 8);
        if (pool_writer && valid) {
          uint4* dst = reinterpret_cast<uint4*>(a.pool_out + ((static_cast<size_t>(b) * (a.H >> 1) + (h >> 1)) * (a.W >> 1) + (w >> 1)) * Cfg::N);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
        }
      }
    }
    if (lane == 0) bulk_wait_group_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the leader's MMAs read the peer's shared memory - nobody leaves before both are done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace ub
