// Weight gradient of the stem convolution (Cin <= 4 -> 64, README.md:1452 under loss.backward()) on tensor cores:
//
//   dW[co][ci][tap] += sum_p dY[p][co] * x[p + shift(tap)][ci]
//
// Same idea as wgrad_umma.cuh (both operands MN-major, reduction over pixels), but the x operand has only 4 channels
// per pixel, so its [128 pixels][k = tap*4+ci, padded to 64] rows are assembled by producer warps exactly like the
// forward stem's im2col tile (stem_umma.cuh), while the dY operand arrives by TMA.
// One MMA group handles TWO 16x8-pixel tiles P1, P2 at once: A = [dY(P1) | dY(P2)] (M = 128), B = [im2col(P1) |
// im2col(P2)] (N = 128); the diagonal quadrants of D accumulate the wanted sums (the cross terms are discarded), so
// M = 128 stays full with a 64-channel gradient. The tensor work is negligible; the kernel is bound by reading dY once.
#pragma once
#include "ptx.cuh"
#include "wgrad_umma.cuh"

namespace ub {

struct StemWgradArgs {
  int B, H, W, Cin;
  int lCout;             // logical output channels (<= 64): gradients of zero-extended channels are not written
  int tiles_w, tiles_h;  // 8-pixel x 16-row tiles per image
  const uint2* x;        // [B,H,W] x 4 bf16
  GradRoute route;       // gradient destination (ptx.cuh)
  long long off;         // flat index of dW [64][Cin][3][3]
};

struct StemWgradCfg {
  static constexpr int STAGES = 3;
  static constexpr int STAGE_BYTES = 4 * 16384;  // dY(P1), dY(P2), im2col(P1), im2col(P2)
  static constexpr int THREADS = 10 * 32;        // warp 0 TMA, warp 1 MMA + TMEM, warps 2..9 im2col producers / epilogue
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
};

__global__ void __launch_bounds__(StemWgradCfg::THREADS, 1)
stem_wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmD, const StemWgradArgs a) {
  pdl_enter();
  using Cfg = StemWgradCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                 // [STAGES] TMA (1 arrival + tx bytes) + 256 producer threads -> MMA
  uint64_t* empty = bars + Cfg::STAGES;  // [STAGES] MMA -> TMA / producers
  uint64_t* done = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full[s], 257);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_img = a.tiles_w * a.tiles_h;
  const int total_tiles = tiles_per_img * a.B;
  const int pairs = (total_tiles + 1) / 2;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
        const int s = it % Cfg::STAGES;
        mbar_wait_parked(&empty[s], ((it / Cfg::STAGES) & 1) ^ 1);
        uint8_t* sD = smem + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[s], 2 * 16384);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int t = 2 * pr + h;  // t == total_tiles (odd count): image index B is out of bounds -> TMA zero-fills the box
          const int b = t / tiles_per_img;
          const int ti = t - b * tiles_per_img;
          tma_load_4d(sD + h * 16384, &tmD, &full[s], 0, (ti % a.tiles_w) * 8, (ti / a.tiles_w) * 16, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(128, 128) | (1u << 15) | (1u << 16);  // A and B MN-major
      const uint64_t d_hi = make_sw128_mnmajor_desc(0, 16384, 1024);
      int it = 0;
      for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
        const int s = it % Cfg::STAGES;
        mbar_wait(&full[s], (it / Cfg::STAGES) & 1);
        tc_fence_after();
        const uint32_t sD = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint64_t da = d_hi + (sD >> 4);
        const uint64_t db = d_hi + ((sD + 2 * 16384) >> 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) umma_f16(tmem_base, da + j * 128, db + j * 128, idesc, (it | j) != 0);
        umma_commit(&empty[s]);
      }
      umma_commit(done);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ im2col producers: thread = (tile half, pixel row)
    const int half = (warp - 2) >> 2;
    const int m = ((warp - 2) & 3) * 32 + lane;
    const int tw = m & 7, th = m >> 3;
    int it = 0;
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      const int s = it % Cfg::STAGES;
      const int t = 2 * pr + half;
      const int b = t / tiles_per_img;
      const int ti = t - b * tiles_per_img;
      const int w = (ti % a.tiles_w) * 8 + tw;
      const int h = (ti / a.tiles_w) * 16 + th;
      uint2 tap[9];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int hh = h + r - 1, ww = w + c - 1;
          uint2 v = make_uint2(0u, 0u);
          if (t < total_tiles && hh >= 0 && hh < a.H && ww >= 0 && ww < a.W) {
            v = __ldg(a.x + (static_cast<size_t>(b) * a.H + hh) * a.W + ww);
          }
          tap[r * 3 + c] = v;
        }
      }
      mbar_wait_parked(&empty[s], ((it / Cfg::STAGES) & 1) ^ 1, 2000);
      const uint32_t row = smem_u32(smem + s * Cfg::STAGE_BYTES + (2 + half) * 16384 + m * 128);
      const int sw = m & 7;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        st_shared_v4(row + ((j ^ sw) << 4), tap[2 * j].x, tap[2 * j].y, tap[2 * j + 1].x, tap[2 * j + 1].y);
      }
      st_shared_v4(row + ((4 ^ sw) << 4), tap[8].x, tap[8].y, 0u, 0u);
      st_shared_v4(row + ((5 ^ sw) << 4), 0u, 0u, 0u, 0u);
      st_shared_v4(row + ((6 ^ sw) << 4), 0u, 0u, 0u, 0u);
      st_shared_v4(row + ((7 ^ sw) << 4), 0u, 0u, 0u, 0u);
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }
    // ------------------------------------------------------------ epilogue (warps 2..5): diagonal quadrants -> dW
    if (warp < 6) {
      const int q = warp & 3;  // TMEM lane quarter this warp may read
      const int row = q * 32 + lane;
      mbar_wait(done, 0);
      tc_fence_after();
      const int co = row & 63;
      const int col0 = (row >> 6) * 64;  // rows 0..63 pair with columns 0..63, rows 64..127 with columns 64..127
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + col0, v);
      tmem_ld_wait();
      if (pairs > static_cast<int>(blockIdx.x) && co < a.lCout) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int tp = k >> 2, ci = k & 3;
          if (ci < a.Cin) grad_add(a.route, a.off + (static_cast<long long>(co) * a.Cin + ci) * 9 + tp, __uint_as_float(v[k]));
        }
      }
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + col0 + 32, v);
      tmem_ld_wait();
      if (pairs > static_cast<int>(blockIdx.x) && co < a.lCout) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // tap 8
          if (k < a.Cin) grad_add(a.route, a.off + (static_cast<long long>(co) * a.Cin + k) * 9 + 8, __uint_as_float(v[k]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace ub
