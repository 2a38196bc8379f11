// 3x3 convolution for the wide, shallow layers (Cout = 64 or 128 at 224^2 / 112^2): one halo'd
// activation patch per 64-channel block feeds all nine taps through *shifted* UMMA descriptors, and
// the packed weights stay resident in shared memory when they fit.
//
// Why a second kernel: with one TMA box per tap (conv_umma.cuh) these layers need 128-192 bytes of
// operand traffic per tensor-pipe cycle per SM and a TMA + mbarrier round trip per 64-deep K block;
// profiles/round1 shows them at 30-60 % of the bf16 peak while the N = 256 layers sit at 87-100 %.
// Here a tile is 16 rows x 8 pixels (M = 128). The patch [18][10][64 ch] is loaded ONCE per channel
// block (23 KB instead of 9 x 16 KB); tap (r,s) is the descriptor
//     start = patch + (r*10 + s)*128 B,   SBO = 10*128 B   (8-row group = 8 consecutive pixels)
// which works because the 128B swizzle is a function of the shared-memory address itself
// (tools/halo_probe.cu verified this on B200, base_offset field = 0).
// Weights: [tap][64-ch block] tiles of BLOCK_N x 128 B, three taps per barrier stage; if all 9*KC
// tiles fit the ring they are loaded once per CTA (resident mode), otherwise they stream.
//
// Epilogues: STORE (+ fused 2x2 max-pool) as in conv_umma.cuh, or HEAD: the 1x1 output conv +
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
  int epi, relu;
  int Cout;
  const float* bias;        // [Cout]
  __nv_bfloat16* out;       // [B,H,W,Cout]            (HEPI_STORE)
  __nv_bfloat16* pool_out;  // [B,H/2,W/2,Cout] or null (HEPI_STORE)
  const float* head_w;      // [Cout]                   (HEPI_HEAD)
  float head_b, thr;
  float* logits;            // [B,H,W] or null
  float* probs;             // [B,H,W] or null
  uint8_t* mask;            // [B,H,W] or null
};

template <int BLOCK_N>
struct HaloCfg {
  static constexpr int A_STAGE_BYTES = 18 * 10 * 128;  // 23040
  static constexpr int A_STAGE_PITCH = 23552;          // next multiple of 1024
  static constexpr int A_STAGES = 3;
  static constexpr int B_TILE_BYTES = BLOCK_N * 128;
  static constexpr int B_STAGE_BYTES = 3 * B_TILE_BYTES;  // three taps (one kernel row) per stage
  static constexpr int B_STAGES = (BLOCK_N == 64) ? 6 : 3;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = A_STAGES * A_STAGE_PITCH + B_STAGES * B_STAGE_BYTES + BAR_BYTES + 1024;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int BLOCK_N>
__global__ void __launch_bounds__(192, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmW, const HaloArgs a) {
  using Cfg = HaloCfg<BLOCK_N>;
  constexpr int AS = Cfg::A_STAGES, BS = Cfg::B_STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + AS * Cfg::A_STAGE_PITCH;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + BS * Cfg::B_STAGE_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + AS;
  uint64_t* b_full = a_empty + AS;
  uint64_t* b_empty = b_full + BS;
  uint64_t* tfull = b_empty + BS;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW);
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

  const int KC = a.kc0 + a.kc1;
  const int tiles_per_img = a.tiles_w * a.tiles_h;
  const int total_tiles = tiles_per_img * a.B;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    bool first = true;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int b = t / tiles_per_img;
      const int ti = t - b * tiles_per_img;
      const int w0 = (ti % a.tiles_w) * 8;
      const int h0 = (ti / a.tiles_w) * 16;
      for (int c = 0; c < KC; ++c) {
        mbar_wait(&a_empty[as], aph ^ 1);
        mbar_expect_tx(&a_full[as], Cfg::A_STAGE_BYTES);
        if (c < a.kc0) {
          tma_load_4d(smA + as * Cfg::A_STAGE_PITCH, &tmA0, &a_full[as], c * 64, w0 - 1, h0 - 1, b);
        } else {
          tma_load_4d(smA + as * Cfg::A_STAGE_PITCH, &tmA1, &a_full[as], (c - a.kc0) * 64, w0 - 1, h0 - 1, b);
        }
        if (++as == AS) {
          as = 0;
          aph ^= 1;
        }
        if (!a.resident || first) {
          for (int r = 0; r < 3; ++r) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            mbar_expect_tx(&b_full[bs], Cfg::B_STAGE_BYTES);
            uint8_t* dst = smB + bs * Cfg::B_STAGE_BYTES;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              tma_load_2d(dst + s * Cfg::B_TILE_BYTES, &tmW, &b_full[bs], ((r * 3 + s) * KC + c) * 64, 0);
            }
            if (++bs == BS) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
      }
      first = false;
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16_f32(128, BLOCK_N);
    const uint64_t da_hi = make_sw128_kmajor_desc(0, 1280, 0);  // SBO = one patch row (10 pixels)
    const uint64_t db_hi = make_sw128_kmajor_desc(0, 1024, 0);
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
        const uint64_t da0 = da_hi + (smem_u32(smA + as * Cfg::A_STAGE_PITCH) >> 4);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int slot = a.resident ? (c * 3 + r) : bs;
          if (!a.resident || first) {
            mbar_wait(&b_full[slot], a.resident ? 0u : bph);
            tc_fence_after();
          }
          const uint64_t db0 = db_hi + (smem_u32(smB + slot * Cfg::B_STAGE_BYTES) >> 4);
#pragma unroll
          for (int s = 0; s < 3; ++s) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // tap (r,s): +(r*10+s) patch rows of 128 B; k: +32 B inside the swizzled row (>>4 units)
              umma_f16(d_tmem, da0 + ((r * 10 + s) * 8 + k * 2), db0 + (s * (Cfg::B_TILE_BYTES >> 4) + k * 2), idesc,
                       (c | r | s | k) != 0);
            }
          }
          if (!a.resident) {
            umma_commit(&b_empty[bs]);
            if (++bs == BS) {
              bs = 0;
              bph ^= 1;
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
  } else if (warp >= 2) {
    // ------------------------------------------------------------ epilogue (4 warps x 32 rows)
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int tw = m & 7;
    const int th = m >> 3;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const int b = t / tiles_per_img;
      const int ti = t - b * tiles_per_img;
      const int w = (ti % a.tiles_w) * 8 + tw;
      const int h = (ti / a.tiles_w) * 16 + th;
      const bool valid = (w < a.W) && (h < a.H);
      const size_t pix = (static_cast<size_t>(b) * a.H + h) * a.W + w;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
      float z = 0.f;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
        const int n = c * 32;
        const float4* bias4 = reinterpret_cast<const float4*>(a.bias + n);
        uint32_t p[16];
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
        if (a.epi == HEPI_STORE) {
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(a.out + pix * a.Cout + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
          }
          if (a.pool_out != nullptr) {
            // 2x2 partners: lane^1 (w) and lane^8 (h), both inside the warp (rows are [h][8 pixels])
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              uint32_t x = p[j];
              x = bf16x2_max(x, __shfl_xor_sync(0xffffffffu, x, 1));
              x = bf16x2_max(x, __shfl_xor_sync(0xffffffffu, x, 8));
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
      if (a.epi == HEPI_HEAD && valid) {
        z += a.head_b;
        if (a.logits != nullptr) a.logits[pix] = z;
        if (a.probs != nullptr || a.mask != nullptr) {
          const float sg = 1.f / (1.f + expf(-z));
          if (a.probs != nullptr) a.probs[pix] = sg;
          if (a.mask != nullptr) a.mask[pix] = (sg > a.thr) ? 255 : 0;
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
