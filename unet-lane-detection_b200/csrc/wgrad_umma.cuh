// Weight-gradient GEMM on tcgen05 tensor cores, directly on NHWC bf16 tensors (no transposes):
//
//   dW[co][tap][ci] += sum over pixels p of  dy[p][co] * x[p + shift(tap)][ci]        (README.md:2078 loss.backward())
//
// The reduction dimension (pixels) is the strided one in NHWC, so both operands are fed to the tensor core in
// MN-major form: a TMA box (64 ch, TW, TH, TB) lands in shared memory as [128 pixel rows][64 channels], which is
// exactly a 128B-swizzled MN-major UMMA block (8-row K groups 1024 B apart, 64-channel MN blocks LBO apart).
// tools/mn_probe.cu verified the descriptor convention on B200.
//
// One CTA owns one output tile D[M = 128 "ci rows"][N = BLOCK_N co] of one tap over a slice of the pixel tiles
// (split-K), then adds it to the fp32 gradient with red.global.add. The 128 rows are either 2 x 64 input channels
// (Cin >= 128, possibly from two source tensors for the decoder concat) or, for Cin == 64, the SAME 64 channels at
// two different taps (two shifted boxes), so M = 128 is always full.
// Two launch forms of the same body (PAIR, ptx.cuh "CTA pair"): wgrad_umma_kernel<N> = one CTA per output tile;
// wgrad_umma2_kernel<N> (N >= 128) = a CTA pair takes two consecutive row tiles (two ci blocks, or two taps when Cin <= 128)
// of the SAME column block and K slice: M = 256 per MMA, each CTA loads its own x boxes and only HALF of the dy blocks, so a
// third / a quarter less operand traffic is pulled through L2 (the 1-CTA form is L2-bound: profiles/r1_ncu_full_train_wgrad256.txt).
// Also used for ConvTranspose2d weight gradients: "taps" are the four (dy,dx) quads and the dy operand is read
// through the strided quad views of the upsampled gradient.
#pragma once
#include "ptx.cuh"

namespace ub {

struct WgradArgs {
  int B, H, W;                    // pixel grid of the reduction (conv: activation size; ConvT: input size)
  int TW, TH, TB;
  int tiles_w, tiles_h, tiles_b;
  int taps;                       // 9 (3x3 conv), 4 (ConvT quads: dy operand view selected by tap), 1
  int shift;                      // 1: tap -> spatial shift (tap/3-1, tap%3-1) of the x operand (3x3 conv)
  int pair_taps;                  // 1: rows 0..63 = tap 2*m_tile, rows 64..127 = tap 2*m_tile+1 (Cin == 64)
  int m_tiles;                    // Cin/128, or ceil(taps/2) when pair_taps
  int n_tiles;                    // Cout / BLOCK_N
  int ksplit;                     // CTAs sharing one output tile
  int c_split;                    // channels of x source 0 (the rest come from source 1)
  int Cin, Cout;                  // PHYSICAL channel counts (multiples of 64)
  int lc0, lc1, lcout;            // LOGICAL channels of x source 0 / 1 and of dy: physical channels beyond them are the exact
                                  //    zeros of a zero-extended tensor and have no gradient slot (dW is indexed logically)
  int dup_rows;                   // 1: Cin == 64 without tap pairing (ConvT): both 64-row blocks hold the same channels,
                                  //    rows 64..127 are dropped
  long long s_co, s_ci, s_tap;    // dW index = co*s_co + ci*s_ci + tap*s_tap (PyTorch layouts are written directly)
  int k_mmas;                     // 16-pixel MMAs per box = box rows / 16 (8 unless TB exceeds the batch)
  int a_bytes;                    // bytes one x / dy box load delivers (box rows * 128; fewer rows when TB exceeds the batch)
  int blk_bytes;                  // shared-memory pitch of one 64-channel box (full box rows * 128) = LBO of the descriptors
  int stages;                     // operand ring depth chosen by the host (<= WGRAD_MAX_STAGES)
  int rt_total;                   // row tiles = m_tiles * (pair_taps ? 1 : taps); a CTA pair takes two consecutive ones
  GradRoute route;                // where gradient atomics go (local buffer, or the owner rank's buffer over NVLink)
  long long off;                  // flat index of this tensor's first element
};

constexpr int WGRAD_MAX_STAGES = 8;
constexpr int WGRAD_RING_BYTES = 192 * 1024;

template <int BLOCK_N>
struct WgradCfg {
  static constexpr int NB = BLOCK_N / 64;
  static constexpr int SMEM_BYTES = WGRAD_RING_BYTES + 256 + 1024;
  // ring depth for a given box size (host + device)
  __host__ __device__ static constexpr int stages_for(int blk_bytes) {
    const int s = WGRAD_RING_BYTES / ((2 + NB) * blk_bytes);
    return s > WGRAD_MAX_STAGES ? WGRAD_MAX_STAGES : s;
  }
};

// MN-major, 128B-swizzle shared-memory descriptor: LBO = bytes between 64-wide MN blocks, SBO = bytes between 8-row K groups.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int BLOCK_N, bool PAIR>
__device__ __forceinline__ void wgrad_umma_body(const CUtensorMap& tmX0, const CUtensorMap& tmX1, const CUtensorMap& tmD0,
                                                const CUtensorMap& tmD1, const CUtensorMap& tmD2, const CUtensorMap& tmD3,
                                                const WgradArgs& a) {
  constexpr int P = PAIR ? 2 : 1;       // CTAs that share one MMA
  using Cfg = WgradCfg<BLOCK_N>;
  constexpr int NB = Cfg::NB / P;       // 64-channel dy blocks loaded by ONE CTA
  static_assert(!PAIR || Cfg::NB >= 2, "the CTA-pair form needs BLOCK_N >= 128");
  constexpr int TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  const int STAGES = a.stages;
  const int BLK = a.blk_bytes;
  const int STAGE_BYTES = (2 + NB) * BLK;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WGRAD_RING_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + WGRAD_MAX_STAGES;
  uint64_t* done = bars + 2 * WGRAD_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WGRAD_MAX_STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmD0);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_p<PAIR>(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  cta_sync_p<PAIR>();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item: (row-tile group, column block, K slice); row tile rt = m_tile + m_tiles * tap
  const uint32_t rank = pair_rank<PAIR>();
  const bool leader = rank == 0;
  int wi = blockIdx.x / P;
  const int ks = wi % a.ksplit;
  wi /= a.ksplit;
  const int n_tile = wi % a.n_tiles;
  wi /= a.n_tiles;
  const int rt = P * wi + static_cast<int>(rank);
  const bool dummy = rt >= a.rt_total;               // odd row-tile count: the pair's second CTA repeats the last tile,
  const int rtc = dummy ? a.rt_total - 1 : rt;       //   its result is dropped
  const int m_tile = rtc % a.m_tiles;
  const int tap_o = a.pair_taps ? 0 : rtc / a.m_tiles;  // tap of this tile (pair_taps mode: taps come from m_tile)
  const int ptiles = a.tiles_w * a.tiles_h * a.tiles_b;
  const int k_lo = static_cast<int>(static_cast<long long>(ptiles) * ks / a.ksplit);
  const int k_hi = static_cast<int>(static_cast<long long>(ptiles) * (ks + 1) / a.ksplit);

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = k_lo; kt < k_hi; ++kt) {
        const int w0 = (kt % a.tiles_w) * a.TW;
        const int h0 = ((kt / a.tiles_w) % a.tiles_h) * a.TH;
        const int b0 = (kt / (a.tiles_w * a.tiles_h)) * a.TB;
        mbar_wait_parked(&empty[stage], phase ^ 1);
        uint8_t* sA = smem + stage * STAGE_BYTES;
        uint8_t* sB = sA + 2 * BLK;
        if (leader) mbar_expect_tx(&full[stage], P * (2 + NB) * a.a_bytes);   // the loads of every CTA of the group
        // x operand: two 64-row blocks
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          int tap = tap_o, c = m_tile * 128 + hb * 64;
          if (a.dup_rows) c = 0;
          if (a.pair_taps) {
            tap = m_tile * 2 + hb;
            if (tap >= a.taps) tap = a.taps - 1;  // odd tap count: the last block is a duplicate whose rows are dropped
            c = 0;
          }
          int dy = 0, dx = 0;
          if (a.shift) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          if (c < a.c_split) {
            tma_load_4d_p<PAIR>(sA + hb * BLK, &tmX0, &full[stage], c, w0 + dx, h0 + dy, b0);
          } else {
            tma_load_4d_p<PAIR>(sA + hb * BLK, &tmX1, &full[stage], c - a.c_split, w0 + dx, h0 + dy, b0);
          }
        }
        // dy operand: NB 64-channel blocks (ConvT: through the quad view of this tile's tap)
        const CUtensorMap* md = &tmD0;
        if (a.taps == 4) md = tap_o == 0 ? &tmD0 : tap_o == 1 ? &tmD1 : tap_o == 2 ? &tmD2 : &tmD3;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          tma_load_4d_p<PAIR>(sB + nb * BLK, md, &full[stage], n_tile * BLOCK_N + (static_cast<int>(rank) * NB + nb) * 64, w0, h0, b0);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && elect_one()) {
      // both operands MN-major: bits 15 / 16 of the instruction descriptor
      constexpr uint32_t idesc = make_idesc_bf16_f32(128 * P, BLOCK_N) | (1u << 15) | (1u << 16);
      const uint64_t d_hi = make_sw128_mnmajor_desc(0, BLK, 1024);
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = k_lo; kt < k_hi; ++kt) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * STAGE_BYTES) & 0x3FFFFu;
        const uint64_t da = d_hi + (sA >> 4);
        const uint64_t db = d_hi + ((sA + 2 * BLK) >> 4);
#pragma unroll
        for (int j = 0; j < a.k_mmas; ++j) {
          // 16 pixel rows per MMA = two 8-row groups = 2048 B
          umma_f16_p<PAIR>(tmem_base, da + j * 128, db + j * 128, idesc, (kt > k_lo || j > 0) ? 1u : 0u);
        }
        umma_commit_p<PAIR>(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit_p<PAIR>(done);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue: 4 warps, fp32 atomic accumulation
    const int q = warp & 3;
    const int m = q * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after();
    int tap = tap_o, ci = m_tile * 128 + m;
    bool live = k_hi > k_lo && !dummy;
    if (a.pair_taps) {
      tap = m_tile * 2 + (m >> 6);
      ci = m & 63;
      live = live && tap < a.taps;
    }
    if (a.dup_rows) live = live && m < 64;
    // physical channel of cat(x0, x1) -> logical input channel of the weight tensor (or none: a zero-extended channel)
    if (ci < a.c_split) {
      live = live && ci < a.lc0;
    } else {
      live = live && (ci - a.c_split) < a.lc1;
      ci = a.lc0 + (ci - a.c_split);
    }
    const long long base = a.off + static_cast<long long>(tap) * a.s_tap + static_cast<long long>(ci) * a.s_ci;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int co = n_tile * BLOCK_N + c * 32 + j;
          if (co < a.lcout) grad_add(a.route, base + static_cast<long long>(co) * a.s_co, __uint_as_float(v[j]));
        }
      }
    }
  }

  tc_fence_before();
  cta_sync_p<PAIR>();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_p<PAIR>(tmem_base, TMEM_COLS);
  }
}


#define UB_WGRAD_PARAMS                                                                                                   \
  const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,                                    \
      const __grid_constant__ CUtensorMap tmD0, const __grid_constant__ CUtensorMap tmD1,                                \
      const __grid_constant__ CUtensorMap tmD2, const __grid_constant__ CUtensorMap tmD3, const WgradArgs a

template <int BLOCK_N>
__global__ void __launch_bounds__(192, 1) wgrad_umma_kernel(UB_WGRAD_PARAMS) {
  pdl_enter();
  wgrad_umma_body<BLOCK_N, false>(tmX0, tmX1, tmD0, tmD1, tmD2, tmD3, a);
}

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1) wgrad_umma2_kernel(UB_WGRAD_PARAMS) {
  pdl_enter();
  wgrad_umma_body<BLOCK_N, true>(tmX0, tmX1, tmD0, tmD1, tmD2, tmD3, a);
}

}  // namespace ub
