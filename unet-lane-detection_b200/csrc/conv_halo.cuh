// 3x3 convolution for the wide, shallow layers (Cout = 64 or 128 at 224^2 / 112^2): one halo'd
// activation patch per 64-channel block feeds all nine taps through *shifted* UMMA descriptors, the
// packed weights stay resident in shared memory when they fit, and the output tile leaves through
// shared memory + one TMA store instead of 16-byte scattered stores.
//
// Why a second kernel: with one TMA box per tap (conv_umma.cuh) these layers need 128-192 bytes of
// operand traffic per tensor-pipe cycle per SM and a TMA + mbarrier round trip per 64-deep K block;
// profiles/round1 shows them at 30-60 % of the bf16 peak while the N = 256 layers sit at 87-100 %.
// Here a tile is 16 rows x 8 pixels (M = 128). The patch [18][10][64 ch] is loaded ONCE per channel
// block (23 KB instead of 9 x 16 KB); tap (r,s) is the descriptor
//     start = patch + (r*10 + s)*128 B,   SBO = 10*128 B   (8-row group = 8 consecutive pixels)
// which works because the 128B swizzle is a function of the shared-memory address itself
// (tools/halo_probe.cu verified this on B200, base_offset field = 0).
// Weights: one [BLOCK_N x 64] tile per (tap, 64-ch block); if all 9*KC tiles fit the ring they are
// loaded once per CTA (resident mode), otherwise they stream through it.
//
// Epilogues: STORE (+ fused 2x2 max-pool): TMEM -> registers -> bias/ReLU/bf16 -> swizzled staging
// tile in smem -> TMA store (the tensor map clips partial tiles); or HEAD: the 1x1 output conv +
// sigmoid + threshold (README.md:1481, src/unet.py:63-67) evaluated on the tile while it is still in
// registers, so the last 64-channel activation never goes to HBM.
#pragma once
#include "ptx.cuh"

namespace ub {

enum : int { HEPI_STORE = 0, HEPI_HEAD = 1 };

struct HaloArgs {
  int B, H, W;
  int tiles_w, tiles_h;  // 8-pixel / 16-row tiles per image
  int kc0, kc1;          // 64-channel blocks from source 0 / source 1
  int resident;          // 1: all weight tiles fit the ring and are loaded once
  int a_stages, b_stages, n_stg;  // shared-memory carve-up chosen by the host (see halo_smem_plan)
  int epi, relu, pool;
  int Cout;
  const float* bias;        // [Cout]
  const float* head_w;      // [Cout]                   (HEPI_HEAD)
  float head_b, thr;
  float* logits;            // [B,H,W] or null
  float* probs;             // [B,H,W] or null
  uint8_t* mask;            // [B,H,W] or null
};

struct HaloCfg {
  static constexpr int A_STAGE_BYTES = 18 * 10 * 128;  // 23040
  static constexpr int A_STAGE_PITCH = 23552;          // next multiple of 1024
  static constexpr int MAX_A = 4, MAX_B = 18;
  static constexpr int BAR_BYTES = 1024;  // 48 mbarriers + TMEM slot + 128 floats of head partial sums
  static constexpr int SMEM_LIMIT = 232448;
};

// Host + device: byte size of the dynamic shared memory for a given carve-up.
__host__ __device__ constexpr int halo_smem_bytes(int block_n, int a_stages, int b_stages, int n_stg, int pool) {
  return a_stages * HaloCfg::A_STAGE_PITCH + b_stages * block_n * 128 + n_stg * (block_n / 64) * (16384 + (pool ? 4096 : 0)) +
         HaloCfg::BAR_BYTES + 1024;
}

constexpr int HALO_THREADS = 320;  // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..9 epilogue (2 per TMEM lane quarter)

template <int BLOCK_N>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmPool, const HaloArgs a) {
  constexpr int B_TILE = BLOCK_N * 128;
  constexpr int HALVES = BLOCK_N / 64;
  constexpr int TMEM_COLS = 2 * BLOCK_N;
  const int AS = a.a_stages, BS = a.b_stages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smA + AS * HaloCfg::A_STAGE_PITCH;
  uint8_t* smS = smB + BS * B_TILE;                            // [n_stg][HALVES][16 KB] output staging
  uint8_t* smP = smS + a.n_stg * HALVES * 16384;               // [n_stg][HALVES][4 KB] pooled staging (if pool)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smP + (a.pool ? a.n_stg * HALVES * 4096 : 0));
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + HaloCfg::MAX_A;
  uint64_t* b_full = a_empty + HaloCfg::MAX_A;
  uint64_t* b_empty = b_full + HaloCfg::MAX_B;
  uint64_t* tfull = b_empty + HaloCfg::MAX_B;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* zpart = reinterpret_cast<float*>(tmem_slot + 2);      // [128] partial head sums of column group 1

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmOut);
    tma_prefetch_desc(&tmPool);
    for (int s = 0; s < AS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < BS; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 256);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int KC = a.kc0 + a.kc1;
  const int tiles_per_img = a.tiles_w * a.tiles_h;
  const int total_tiles = tiles_per_img * a.B;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: one elected thread runs the whole loop
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool first = true;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = t / tiles_per_img;
        const int ti = t - b * tiles_per_img;
        const int w0 = (ti % a.tiles_w) * 8;
        const int h0 = (ti / a.tiles_w) * 16;
        for (int c = 0; c < KC; ++c) {
          mbar_wait_parked(&a_empty[as], aph ^ 1);
          mbar_expect_tx(&a_full[as], HaloCfg::A_STAGE_BYTES);
          if (c < a.kc0) {
            tma_load_4d(smA + as * HaloCfg::A_STAGE_PITCH, &tmA0, &a_full[as], c * 64, w0 - 1, h0 - 1, b);
          } else {
            tma_load_4d(smA + as * HaloCfg::A_STAGE_PITCH, &tmA1, &a_full[as], (c - a.kc0) * 64, w0 - 1, h0 - 1, b);
          }
          if (++as == AS) {
            as = 0;
            aph ^= 1;
          }
          if (!a.resident || first) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait_parked(&b_empty[bs], bph ^ 1);
              mbar_expect_tx(&b_full[bs], B_TILE);
              tma_load_2d(smB + bs * B_TILE, &tmW, &b_full[bs], (tap * KC + c) * 64, 0);
              if (++bs == BS) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
        }
        first = false;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: one elected thread, straight-line MMA blocks
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(128, BLOCK_N);
      const uint64_t da_hi = make_sw128_kmajor_desc(0, 1280, 0);  // SBO = one patch row (10 pixels)
      const uint64_t db_hi = make_sw128_kmajor_desc(0, 1024, 0);
      const uint64_t db_base = db_hi + (smem_u32(smB) >> 4);
      int as = 0, bs = 0, it = 0;
      uint32_t aph = 0, bph = 0;
      bool first = true;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int c = 0; c < KC; ++c) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint64_t da0 = da_hi + (smem_u32(smA + as * HaloCfg::A_STAGE_PITCH) >> 4);
          if (a.resident && !first) {
            // steady state of the resident mode: 36 MMAs back to back, no barrier traffic
            const uint64_t dbc = db_base + static_cast<uint64_t>(c) * 9 * (B_TILE >> 4);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_f16(d_tmem, da0 + (((tap / 3) * 10 + (tap % 3)) * 8 + k * 2), dbc + (tap * (B_TILE >> 4) + k * 2), idesc,
                         (c | tap | k) != 0);
              }
            }
          } else {
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
              const int slot = a.resident ? (c * 9 + tap) : bs;
              mbar_wait(&b_full[slot], a.resident ? 0u : bph);
              tc_fence_after();
              const uint64_t db0 = db_base + static_cast<uint64_t>(slot) * (B_TILE >> 4);
              const uint64_t dat = da0 + (((tap / 3) * 10 + (tap % 3)) * 8);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // tap (r,s): +(r*10+s) patch rows of 128 B; k: +32 B inside the swizzled row (>>4 units)
                umma_f16(d_tmem, dat + k * 2, db0 + k * 2, idesc, (c | tap | k) != 0);
              }
              if (!a.resident) {
                umma_commit(&b_empty[bs]);
                if (++bs == BS) {
                  bs = 0;
                  bph ^= 1;
                }
              }
            }
          }
          umma_commit(&a_empty[as]);
          if (++as == AS) {
            as = 0;
            aph ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
        first = false;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue: 8 warps; warp pair (q, cg) owns rows 32q..32q+31
    // and the 32-column chunks c with (c & 1) == cg
    const int q = warp & 3;          // TMEM lane quarter this warp may read
    const int cg = (warp - 2) >> 2;  // column group
    const int m = q * 32 + lane;
    const int tw = m & 7;
    const int th = m >> 3;
    const bool store_thread = (threadIdx.x == 64);
    const int prow = (th >> 1) * 4 + (tw >> 1);  // row of this thread's 2x2 window in the pooled 8x4 tile
    const bool pool_writer = ((tw | th) & 1) == 0;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const int b = t / tiles_per_img;
      const int ti = t - b * tiles_per_img;
      const int w0 = (ti % a.tiles_w) * 8;
      const int h0 = (ti / a.tiles_w) * 16;
      const int sb = (a.n_stg == 2) ? (it & 1) : 0;
      uint8_t* stg = smS + sb * HALVES * 16384;
      uint8_t* pstg = smP + sb * HALVES * 4096;
      if (a.epi == HEPI_STORE) {
        // the staging buffer is free once the TMA store that last used it has finished reading it
        if (store_thread) {
          if (a.n_stg == 2) {
            bulk_wait_group_read<1>();
          } else {
            bulk_wait_group_read<0>();
          }
        }
        named_bar_sync(1, 256);
      }
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
      float z = 0.f;
#pragma unroll 1
      for (int c = cg; c < BLOCK_N / 32; c += 2) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        const int n = c * 32;
        const float4* bias4 = reinterpret_cast<const float4*>(a.bias + n);
        float4 bb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bb[j] = __ldg(bias4 + j);
        tmem_ld_wait();
        uint32_t p[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float x0 = __uint_as_float(v[4 * j + 0]) + bb[j].x;
          float x1 = __uint_as_float(v[4 * j + 1]) + bb[j].y;
          float x2 = __uint_as_float(v[4 * j + 2]) + bb[j].z;
          float x3 = __uint_as_float(v[4 * j + 3]) + bb[j].w;
          if (a.relu) {
            x0 = fmaxf(x0, 0.f);
            x1 = fmaxf(x1, 0.f);
            x2 = fmaxf(x2, 0.f);
            x3 = fmaxf(x3, 0.f);
          }
          p[2 * j] = pack_bf16x2(x0, x1);
          p[2 * j + 1] = pack_bf16x2(x2, x3);
        }
        if (a.epi == HEPI_STORE) {
          // staging tile: row m = 128 B (64 channels of one half), 16-byte chunks XOR-swizzled by (row & 7)
          const int half = c >> 1, j0 = (c & 1) * 4;
          const uint32_t row = smem_u32(stg + half * 16384 + m * 128);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            st_shared_v4(row + (((j0 + j) ^ (m & 7)) << 4), p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
          }
          if (a.pool) {
            // 2x2 partners: lane^1 (w) and lane^8 (h), both inside the warp (rows are [h][8 pixels])
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              uint32_t x = p[j];
              x = bf16x2_max(x, __shfl_xor_sync(0xffffffffu, x, 1));
              x = bf16x2_max(x, __shfl_xor_sync(0xffffffffu, x, 8));
              p[j] = x;
            }
            if (pool_writer) {
              const uint32_t prw = smem_u32(pstg + half * 4096 + prow * 128);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                st_shared_v4(prw + (((j0 + j) ^ (prow & 7)) << 4), p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
              }
            }
          }
        } else {
          // HEPI_HEAD: 1x1 conv over the bf16-rounded activations (same rounding point as the unfused path)
          const float4* hw4 = reinterpret_cast<const float4*>(a.head_w + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 ww = __ldg(hw4 + j);
            const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&p[2 * j]);
            const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&p[2 * j + 1]);
            z = fmaf(__low2float(lo), ww.x, z);
            z = fmaf(__high2float(lo), ww.y, z);
            z = fmaf(__low2float(hi), ww.z, z);
            z = fmaf(__high2float(hi), ww.w, z);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (a.epi == HEPI_STORE) {
        fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
        named_bar_sync(2, 256);
        if (store_thread) {
#pragma unroll
          for (int hf = 0; hf < HALVES; ++hf) {
            tma_store_4d(&tmOut, stg + hf * 16384, hf * 64, w0, h0, b);
            if (a.pool) tma_store_4d(&tmPool, pstg + hf * 4096, hf * 64, w0 >> 1, h0 >> 1, b);
          }
          bulk_commit_group();
        }
      } else {
        // combine the two column groups' partial sums: group 1 hands its half over through shared memory
        named_bar_sync(1, 256);  // previous tile's zpart has been consumed
        if (cg == 1) zpart[m] = z;
        named_bar_sync(2, 256);
        if (cg == 0) {
          const int w = w0 + tw, h = h0 + th;
          if (w < a.W && h < a.H) {
            const size_t pix = (static_cast<size_t>(b) * a.H + h) * a.W + w;
            z += zpart[m] + a.head_b;
            if (a.logits != nullptr) a.logits[pix] = z;
            if (a.probs != nullptr || a.mask != nullptr) {
              const float sg = 1.f / (1.f + expf(-z));
              if (a.probs != nullptr) a.probs[pix] = sg;
              if (a.mask != nullptr) a.mask[pix] = (sg > a.thr) ? 255 : 0;
            }
          }
        }
      }
    }
    if (store_thread) bulk_wait_group_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace ub
