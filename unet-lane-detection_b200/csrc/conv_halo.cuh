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
// Weights: one [BLOCK_N x 64] tile per (tap, 64-ch block), three taps (a kernel row) per barrier stage;
// if all 3*KC stages fit the ring they are loaded once per CTA (resident mode), otherwise they stream.
//
// Two launch forms of the same body (PAIR template parameter, ptx.cuh "CTA pair"):
//   conv_halo_kernel<N>   one CTA per 16x8 tile;
//   conv_halo2_kernel<N>  a CTA pair (cta_group::2): M = 256 per MMA, each SM multiplies its own tile and holds HALF of every
//                         weight tile. Why: the 1-CTA form is bound by the shared-memory operand reads of the tensor pipe (4 KB
//                         of A + BLOCK_N*32 B of B per 128xBLOCK_Nx16 MMA against 128 B/cycle: 48 cycles for N = 64 against a
//                         math floor of 32, 64 for N = 128 with no slack); with B split over the pair it is 40 / 48 cycles.
//                         Pair p works on tiles 2*(p + i*pairs) + rank; a missing odd tile is loaded fully out of bounds (zeros)
//                         and its epilogue stores nothing. Outputs are bit-identical to the 1-CTA form.
//
// Epilogues: STORE (+ fused 2x2 max-pool): warp-private units of 32 rows x 64 columns (epilogue.cuh):
// TMEM -> registers -> bias/ReLU/bf16 -> 4 KB swizzled staging tile -> TMA store of the 4-row sub-box
// (the tensor map clips partial tiles); or HEAD (Cout == 64): the 1x1 output conv +
// sigmoid + threshold (README.md:1481, src/unet.py:63-67) evaluated on the tile while it is still in
// registers, so the last 64-channel activation never goes to HBM.
#pragma once
#include "epilogue.cuh"
#include "ptx.cuh"

namespace ub {

enum : int { HEPI_STORE = 0, HEPI_HEAD = 1 };

struct HaloArgs {
  int B, H, W;           // images [b0, b0 + B) of the tensors behind the maps / output pointers
  int b0;
  int tiles_w, tiles_h;  // 8-pixel / 16-row tiles per image
  int kc0, kc1;          // 64-channel blocks from source 0 / source 1
  int resident;          // 1: all weight tiles fit the ring and are loaded once
  int a_stages, b_stages;   // shared-memory carve-up chosen by the host (halo_smem_plan); a weight stage = 3 taps
  int epi, relu;
  int Cout;
  const float* bias;        // [Cout]
  __nv_bfloat16* pool_out;  // HEPI_STORE, optional: [B,H/2,W/2,Cout] = maxpool2x2(out), written from registers
  const float* head_w;      // [Cout]                   (HEPI_HEAD)
  float head_b, thr;
  float* logits;            // [B,H,W] or null
  float* probs;             // [B,H,W] or null
  uint8_t* mask;            // [B,H,W] or null
  double* stat_sum;         // HEPI_STORE, optional (training): per-channel sum / sum of squares of the bf16 output
  double* stat_sumsq;
  EpiBnBwd bn;              // HEPI_STORE, optional (training backward): fused BatchNorm-backward sums (epilogue.cuh); bn.y null: off
};

struct HaloCfg {
  static constexpr int A_STAGE_BYTES = 18 * 10 * 128;  // 23040
  static constexpr int A_STAGE_PITCH = 23552;          // next multiple of 1024
  static constexpr int MAX_A = 4, MAX_B = 6;  // weight stages hold one kernel row (3 taps) each
  static constexpr int BAR_BYTES = 512;   // 48 mbarriers + TMEM slot
  static constexpr int STG_BYTES = 8 * 4096;  // one private 4 KB staging tile per epilogue warp
  static constexpr int SMEM_LIMIT = 232448;
};

// Host + device: byte size of the dynamic shared memory for a given carve-up (pair: a weight tile is BLOCK_N/2 rows per CTA).
// ystg: a second set of warp-private 4 KB tiles (the y sub-boxes of the fused BatchNorm-backward sums)
__host__ __device__ constexpr int halo_smem_bytes(int block_n, int a_stages, int b_stages, int head, int pair = 0, int ystg = 0) {
  return a_stages * HaloCfg::A_STAGE_PITCH + b_stages * 3 * (pair ? block_n / 2 : block_n) * 128 +
         (head ? 0 : HaloCfg::STG_BYTES) + (ystg ? HaloCfg::STG_BYTES : 0) + HaloCfg::BAR_BYTES + 1024;
}

constexpr int HALO_THREADS = 320;  // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..9 epilogue (2 per TMEM lane quarter)

template <int BLOCK_N, bool PAIR>
__device__ __forceinline__ void conv_halo_body(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmWh,
                                               const CUtensorMap& tmOut, const CUtensorMap& tmY, const HaloArgs& a) {
  constexpr int P = PAIR ? 2 : 1;                       // CTAs that share one MMA
  constexpr int B_HALF = (BLOCK_N / P) * 128;           // bytes of one (tap, 64-channel block) weight tile held by ONE CTA
  constexpr int HALVES = BLOCK_N / 64;
  constexpr int TMEM_COLS = 2 * BLOCK_N;
  const int AS = a.a_stages, BS = a.b_stages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smA + AS * HaloCfg::A_STAGE_PITCH;
  uint8_t* smS = smB + BS * 3 * B_HALF;
  const bool bnb = a.bn.y != nullptr && a.epi == HEPI_STORE;   // fused BatchNorm-backward sums: [8 warps][4 KB] y tiles
  uint8_t* smY = smS + (a.epi == HEPI_STORE ? HaloCfg::STG_BYTES : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smY + (bnb ? HaloCfg::STG_BYTES : 0));
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + HaloCfg::MAX_A;
  uint64_t* b_full = a_empty + HaloCfg::MAX_A;
  uint64_t* b_empty = b_full + HaloCfg::MAX_B;
  uint64_t* tfull = b_empty + HaloCfg::MAX_B;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* ybars = tempty + 3;   // [8] y tile of epilogue warp w has landed

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = pair_rank<PAIR>();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmWh);
    tma_prefetch_desc(&tmOut);
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
      mbar_init(&tempty[s], P * (HALVES == 1 ? 128 : 256));   // the epilogue threads of every CTA of the group
    }
    if (bnb) {
      tma_prefetch_desc(&tmY);
      for (int s = 0; s < 8; ++s) mbar_init(&ybars[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_p<PAIR>(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  cta_sync_p<PAIR>();   // (pair: the peer's barriers are initialised before anything signals them)
  tc_fence_after();
  pdl_wait();           // programmatic dependent launch: everything above overlapped the predecessor's tail
  const uint32_t tmem_base = *tmem_slot;

  const int KC = a.kc0 + a.kc1;
  const int tiles_per_img = a.tiles_w * a.tiles_h;
  const int total_tiles = tiles_per_img * a.B;
  const int pairs = gridDim.x / P;
  const int pair = blockIdx.x / P;
  // iteration i of this group covers tiles P*(pair + i*pairs) + rank; it runs while the first of them exists
  const int t_first = P * pair + static_cast<int>(rank);
  const int t_step = P * pairs;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs; the leader also posts expect_tx)
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool first = true;
      for (int t = t_first; t - static_cast<int>(rank) < total_tiles; t += t_step) {
        // a tile index past the end (odd tile count): image index == B is fully out of bounds -> the box is zero-filled
        const int br = t < total_tiles ? t / tiles_per_img : a.B;
        const int ti = t < total_tiles ? t - br * tiles_per_img : 0;
        const int b = a.b0 + br;
        const int w0 = (ti % a.tiles_w) * 8;
        const int h0 = (ti / a.tiles_w) * 16;
        for (int c = 0; c < KC; ++c) {
          mbar_wait_parked(&a_empty[as], aph ^ 1);
          if (leader) mbar_expect_tx(&a_full[as], P * HaloCfg::A_STAGE_BYTES);
          if (c < a.kc0) {
            tma_load_4d_p<PAIR>(smA + as * HaloCfg::A_STAGE_PITCH, &tmA0, &a_full[as], c * 64, w0 - 1, h0 - 1, b);
          } else {
            tma_load_4d_p<PAIR>(smA + as * HaloCfg::A_STAGE_PITCH, &tmA1, &a_full[as], (c - a.kc0) * 64, w0 - 1, h0 - 1, b);
          }
          if (++as == AS) {
            as = 0;
            aph ^= 1;
          }
          if (!a.resident || first) {
            for (int r = 0; r < 3; ++r) {  // one kernel row (3 taps) per weight stage; this CTA's half of the rows
              mbar_wait(&b_empty[bs], bph ^ 1);
              if (leader) mbar_expect_tx(&b_full[bs], P * 3 * B_HALF);
#pragma unroll
              for (int sx = 0; sx < 3; ++sx) {
                // the weight map's box is HALF a tile (BLOCK_N/2 rows): a pair CTA loads its half, a single CTA both
                if constexpr (PAIR) {
                  tma_load_2d_p<PAIR>(smB + (bs * 3 + sx) * B_HALF, &tmWh, &b_full[bs], ((r * 3 + sx) * KC + c) * 64,
                                      static_cast<int>(rank) * (BLOCK_N / 2));
                } else {
                  tma_load_2d(smB + (bs * 3 + sx) * B_HALF, &tmWh, &b_full[bs], ((r * 3 + sx) * KC + c) * 64, 0);
                  tma_load_2d(smB + (bs * 3 + sx) * B_HALF + B_HALF / 2, &tmWh, &b_full[bs], ((r * 3 + sx) * KC + c) * 64,
                              BLOCK_N / 2);
                }
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
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: one elected thread of the LEADER CTA
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(128 * P, BLOCK_N);
      const uint64_t da_hi = make_sw128_kmajor_desc(0, 1280, 0);  // SBO = one patch row (10 pixels)
      const uint64_t db_hi = make_sw128_kmajor_desc(0, 1024, 0);
      const uint64_t db_base = db_hi + ((smem_u32(smB) & 0x3FFFFu) >> 4);
      int as = 0, bs = 0, it = 0;
      uint32_t aph = 0, bph = 0;
      bool first = true;
      for (int t = t_first; t < total_tiles; t += t_step, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int c = 0; c < KC; ++c) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint64_t da0 = da_hi + ((smem_u32(smA + as * HaloCfg::A_STAGE_PITCH) & 0x3FFFFu) >> 4);
          if (a.resident && !first) {
            const uint64_t dbc = db_base + static_cast<uint64_t>(c) * 9 * (B_HALF >> 4);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_f16_p<PAIR>(d_tmem, da0 + (((tap / 3) * 10 + (tap % 3)) * 8 + k * 2), dbc + (tap * (B_HALF >> 4) + k * 2), idesc,
                          (c | tap | k) != 0);
              }
            }
          } else {
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
              const int slot = a.resident ? (c * 3 + r) : bs;
              mbar_wait(&b_full[slot], a.resident ? 0u : bph);
              tc_fence_after();
              const uint64_t db0 = db_base + static_cast<uint64_t>(slot) * 3 * (B_HALF >> 4);
#pragma unroll
              for (int sx = 0; sx < 3; ++sx) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_f16_p<PAIR>(d_tmem, da0 + ((r * 10 + sx) * 8 + k * 2), db0 + (sx * (B_HALF >> 4) + k * 2), idesc,
                            (c | r | sx | k) != 0);
                }
              }
              if (!a.resident) {
                umma_commit_p<PAIR>(&b_empty[bs]);
                if (++bs == BS) {
                  bs = 0;
                  bph ^= 1;
                }
              }
            }
          }
          umma_commit_p<PAIR>(&a_empty[as]);
          if (++as == AS) {
            as = 0;
            aph ^= 1;
          }
        }
        umma_commit_p<PAIR>(&tfull[acc]);
        first = false;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (both CTAs): as conv_halo.cuh on this CTA's tile
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int tw = m & 7;
    const int th = m >> 3;
    uint8_t* stg = smS + (warp - 2) * 4096;
    const bool pool_writer = ((tw | th) & 1) == 0;
    float st0[4] = {0.f, 0.f, 0.f, 0.f};
    // fused BatchNorm-backward sums: this warp's y tile + barrier, a cursor one unit ahead of the main loop, the constants of
    // the warp's 64 channels (fixed: one column group per warp) and its accumulators
    const int n = ((HALVES == 1) ? 0 : cg) * 64;
    uint8_t* ystg = smY + (warp - 2) * 4096;
    uint64_t* ybar = &ybars[warp - 2];
    uint32_t yph = 0;
    float bk[8], ba[4] = {0.f, 0.f, 0.f, 0.f};
    int yt = t_first + (HALVES == 1 ? cg * t_step : 0);
    const int y_step = (HALVES == 1 ? 2 : 1) * t_step;     // BLOCK_N == 64: the two warps of a quarter alternate tiles
    auto y_issue = [&]() {   // lane 0: request the y sub-box of tile yt (live tiles only), then advance the cursor
      if (yt >= total_tiles) return;
      const int yb = yt / tiles_per_img;
      const int yti = yt - yb * tiles_per_img;
      mbar_expect_tx(ybar, 4096);
      tma_load_4d(ystg, &tmY, ybar, n, (yti % a.tiles_w) * 8, (yti / a.tiles_w) * 16 + 4 * q, a.b0 + yb);
      yt += y_step;
    };
    if (bnb) {
      epi_bnbwd_consts(a.bn, n, lane, bk);
      if (lane == 0) y_issue();
    }
    int it = 0;
    for (int t = t_first; t - static_cast<int>(rank) < total_tiles; t += t_step, ++it) {
      const int acc = it & 1;
      if (HALVES == 1 && acc != cg) continue;
      const bool live = t < total_tiles;
      const int br = live ? t / tiles_per_img : 0;
      const int ti = live ? t - br * tiles_per_img : 0;
      const int b = a.b0 + br;
      const int w0 = (ti % a.tiles_w) * 8;
      const int h0 = (ti / a.tiles_w) * 16;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      uint32_t p[32];
      epi_load_unit(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + n, a.bias + n, a.relu, p);
      tc_fence_before();
      mbar_arrive_p<PAIR>(&tempty[acc]);
      if (!live) continue;   // warp-uniform
      const int w = w0 + tw, h = h0 + th;
      const bool valid = (w < a.W) && (h < a.H);
      if (a.epi == HEPI_STORE) {
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
        epi_stage_row(stg, lane, p);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmOut, stg, n, w0, h0 + 4 * q, b);
          bulk_commit_group();
        }
        if (a.pool_out != nullptr) {
          epi_pool2x2(p, 8);
          if (pool_writer && valid) {
            uint4* dst = reinterpret_cast<uint4*>(
                a.pool_out + ((static_cast<size_t>(b) * (a.H >> 1) + (h >> 1)) * (a.W >> 1) + (w >> 1)) * a.Cout + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
          }
        }
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        if (a.stat_sum != nullptr) epi_stats_accumulate(stg, lane, vmask, st0);
        if (bnb) {
          mbar_wait(ybar, yph);
          yph ^= 1;
          epi_bnbwd_accumulate(stg, ystg, lane, vmask, bk, ba);
          __syncwarp();             // every lane has finished reading the y tile: the next box may land in it
          if (lane == 0) y_issue();
        }
      } else {
        float z = a.head_b;
        const float4* hw4 = reinterpret_cast<const float4*>(a.head_w);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 ww = __ldg(hw4 + j);
          const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&p[2 * j]);
          const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&p[2 * j + 1]);
          z = fmaf(__low2float(lo), ww.x, z);
          z = fmaf(__high2float(lo), ww.y, z);
          z = fmaf(__low2float(hi), ww.z, z);
          z = fmaf(__high2float(hi), ww.w, z);
        }
        if (valid) {
          const size_t pix = (static_cast<size_t>(b) * a.H + h) * a.W + w;
          if (a.logits != nullptr) a.logits[pix] = z;
          if (a.probs != nullptr || a.mask != nullptr) {
            const float sg = 1.f / (1.f + expf(-z));
            if (a.probs != nullptr) a.probs[pix] = sg;
            if (a.mask != nullptr) a.mask[pix] = (sg > a.thr) ? 255 : 0;
          }
        }
      }
    }
    if (a.stat_sum != nullptr && a.epi == HEPI_STORE) epi_stats_flush(a.stat_sum, a.stat_sumsq, n, lane, st0);
    if (bnb) epi_bnbwd_flush(a.bn, n, lane, ba);
    if (lane == 0) bulk_wait_group_read<0>();
  }

  tc_fence_before();
  cta_sync_p<PAIR>();   // (pair: the leader's MMAs read the peer's shared memory - nobody leaves before both are done)
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_p<PAIR>(tmem_base, TMEM_COLS);
  }
}


#define UB_CONV_HALO_PARAMS                                                                                               \
  const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,                                     \
      const __grid_constant__ CUtensorMap tmW /* box = BLOCK_N/2 rows */, const __grid_constant__ CUtensorMap tmOut,      \
      const __grid_constant__ CUtensorMap tmY /* EpiBnBwd y boxes */, const HaloArgs a

template <int BLOCK_N>
__global__ void __launch_bounds__(HALO_THREADS, 1) conv_halo_kernel(UB_CONV_HALO_PARAMS) {
  pdl_launch();
  conv_halo_body<BLOCK_N, false>(tmA0, tmA1, tmW, tmOut, tmY, a);
}

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HALO_THREADS, 1) conv_halo2_kernel(UB_CONV_HALO_PARAMS) {
  pdl_launch();
  conv_halo_body<BLOCK_N, true>(tmA0, tmA1, tmW, tmOut, tmY, a);
}

}  // namespace ub
