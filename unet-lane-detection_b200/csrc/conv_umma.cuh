// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), NHWC bf16 activations.
//
// Replaces, for the U-Net of the reference (README.md:1421-1481):
//   * Conv2d(3x3, pad 1, bias=False) + BatchNorm2d(eval, folded) + ReLU   (README.md:1451-1458)
//   * the 2x2/2 MaxPool that follows an encoder block (README.md:1429,1467) - fused in the epilogue
//   * torch.cat([skip, x], 1) (README.md:1478) - never materialised: the K loop walks two tensor maps
//   * ConvTranspose2d(2f, f, 2, 2) (README.md:1441-1443,1476) as a 1-tap GEMM with N = 4f and a
//     pixel-shuffle store
//
// GEMM view: D[M = 128 output pixels][N = BLOCK_N out channels] += A[M][K] * Wt[N][K],
// K = taps * Cin walked in 64-channel blocks (one 128-byte swizzled row per pixel / per out channel).
// The 128 pixels of a tile are a TB x TH x TW box of the [B,H,W] grid, so a TMA 4-D box load at
// (c0, w0+dx, h0+dy, b0) with zero OOB fill *is* the im2col tile for tap (dy,dx).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (one TMEM lane quarter each). Persistent over tiles; two TMEM accumulator
// stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "ptx.cuh"

namespace ub {

enum : int { EPI_STORE = 0, EPI_CONVT = 1 };

struct ConvArgs {
  int B, H, W;                    // spatial grid of the GEMM rows (conv: output == input size)
  int TW, TH, TB;                 // pixel box of one tile, TW*TH*TB == 128
  int tiles_w, tiles_h, tiles_b;  // number of boxes along each axis
  int n_tiles;                    // N / BLOCK_N
  int taps;                       // 9 (3x3, pad 1) or 1 (pointwise)
  int kc0, kc1;                   // 64-channel blocks taken from source 0 / source 1
  int epi;                        // EPI_*
  int relu;
  int a_bytes;                    // bytes one A box load delivers (128 rows x 128 B unless TB exceeds the batch dim)
  int Cout;                       // EPI_STORE: channels of out; EPI_CONVT: f (out channels of the ConvT)
  const float* bias;              // [Cout]
  __nv_bfloat16* out;             // EPI_STORE: [B,H,W,Cout]   EPI_CONVT: [B,2H,2W,Cout]
  __nv_bfloat16* pool_out;        // EPI_STORE only, optional: [B,H/2,W/2,Cout]
};

template <int BLOCK_N>
struct ConvCfg {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");
};

template <int BLOCK_N>
__global__ void __launch_bounds__(192, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmW, const ConvArgs a) {
  using Cfg = ConvCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                  // [STAGES] TMA -> MMA
  uint64_t* empty = bars + STAGES;        // [STAGES] MMA -> TMA
  uint64_t* tfull = bars + 2 * STAGES;    // [2] MMA -> epilogue
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
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

  const int kchunks = a.kc0 + a.kc1;
  const int num_kb = a.taps * kchunks;
  const int m_tiles = a.tiles_w * a.tiles_h * a.tiles_b;
  const int total_tiles = m_tiles * a.n_tiles;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int n_tile = t % a.n_tiles;
      const int m_tile = t / a.n_tiles;
      const int w0 = (m_tile % a.tiles_w) * a.TW;
      const int h0 = ((m_tile / a.tiles_w) % a.tiles_h) * a.TH;
      const int b0 = (m_tile / (a.tiles_w * a.tiles_h)) * a.TB;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / kchunks;
        const int ch = kb - tap * kchunks;
        int dy = 0, dx = 0;
        if (a.taps == 9) {
          dy = tap / 3 - 1;
          dx = tap % 3 - 1;
        }
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
        mbar_expect_tx(&full[stage], a.a_bytes + Cfg::B_BYTES);
        if (ch < a.kc0) {
          tma_load_4d(sA, &tmA0, &full[stage], ch * 64, w0 + dx, h0 + dy, b0);
        } else {
          tma_load_4d(sA, &tmA1, &full[stage], (ch - a.kc0) * 64, w0 + dx, h0 + dy, b0);
        }
        tma_load_2d(sB, &tmW, &full[stage], kb * 64, n_tile * BLOCK_N);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    constexpr uint32_t idesc = make_idesc_bf16_f32(128, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sB = sA + Cfg::A_BYTES;
        const uint64_t da = make_sw128_kmajor_desc(sA, 1024, 0);
        const uint64_t db = make_sw128_kmajor_desc(sB, 1024, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // +32 bytes (16 bf16) along K inside the 128-byte swizzled row == +2 in the >>4 address field
          umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(&tfull[acc]);
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ epilogue (4 warps, 128 rows)
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;
    const int tw = m % a.TW;
    const int th = (m / a.TW) % a.TH;
    const int tb = m / (a.TW * a.TH);
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n_tile = t % a.n_tiles;
      const int m_tile = t / a.n_tiles;
      const int w = (m_tile % a.tiles_w) * a.TW + tw;
      const int h = ((m_tile / a.tiles_w) % a.tiles_h) * a.TH + th;
      const int b = (m_tile / (a.tiles_w * a.tiles_h)) * a.TB + tb;
      const bool valid = (w < a.W) && (h < a.H) && (b < a.B);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
        const int n = n_tile * BLOCK_N + c * 32;  // first GEMM column of this chunk
        uint32_t p[16];
        if (a.epi == EPI_STORE) {
          const float4* bias4 = reinterpret_cast<const float4*>(a.bias + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(bias4 + j);
            float x0 = __uint_as_float(v[4 * j + 0]) + bb.x;
            float x1 = __uint_as_float(v[4 * j + 1]) + bb.y;
            float x2 = __uint_as_float(v[4 * j + 2]) + bb.z;
            float x3 = __uint_as_float(v[4 * j + 3]) + bb.w;
            if (a.relu) {
              x0 = fmaxf(x0, 0.f);
              x1 = fmaxf(x1, 0.f);
              x2 = fmaxf(x2, 0.f);
              x3 = fmaxf(x3, 0.f);
            }
            p[2 * j] = pack_bf16x2(x0, x1);
            p[2 * j + 1] = pack_bf16x2(x2, x3);
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(a.out + ((static_cast<size_t>(b) * a.H + h) * a.W + w) * a.Cout + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
          }
          if (a.pool_out != nullptr) {
            // 2x2 window partners are lane^1 (w) and lane^TW (h): both inside this warp for TW <= 16.
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              uint32_t x = p[j];
              x = bf16x2_max(x, __shfl_xor_sync(0xffffffffu, x, 1));
              x = bf16x2_max(x, __shfl_xor_sync(0xffffffffu, x, a.TW));
              p[j] = x;
            }
            if (valid && ((tw | th) & 1) == 0) {
              const int Hp = a.H >> 1, Wp = a.W >> 1;
              uint4* dst = reinterpret_cast<uint4*>(
                  a.pool_out + ((static_cast<size_t>(b) * Hp + (h >> 1)) * Wp + (w >> 1)) * a.Cout + n);
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[j] = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
            }
          }
        } else {  // EPI_CONVT: column n -> (dy, dx, co) with co fastest, f = a.Cout
          const int quad = n / a.Cout;
          const int co = n - quad * a.Cout;
          const int dy = quad >> 1, dx = quad & 1;
          const float4* bias4 = reinterpret_cast<const float4*>(a.bias + co);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(bias4 + j);
            p[2 * j] = pack_bf16x2(__uint_as_float(v[4 * j + 0]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y);
            p[2 * j + 1] = pack_bf16x2(__uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
          }
          if (valid) {
            const int Ho = a.H * 2, Wo = a.W * 2;
            uint4* dst = reinterpret_cast<uint4*>(
                a.out + ((static_cast<size_t>(b) * Ho + (2 * h + dy)) * Wo + (2 * w + dx)) * a.Cout + co);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
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
