// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), NHWC bf16 activations.
//
// Replaces, for the U-Net of the reference (README.md:1421-1481):
//   * Conv2d(3x3, pad 1, bias=False) + BatchNorm2d(eval, folded) + ReLU   (README.md:1451-1458)
//   * the 2x2/2 MaxPool that follows an encoder block (README.md:1429,1467) - fused in the epilogue
//   * torch.cat([skip, x], 1) (README.md:1478) - never materialised: the K loop walks two tensor maps
//   * ConvTranspose2d(2f, f, 2, 2) (README.md:1441-1443,1476) as a 1-tap GEMM with N = 4f whose four
//     (dy,dx) column groups are TMA-stored through four strided views of the 2x-upsampled output
//
// GEMM view: D[M = 128 output pixels][N = BLOCK_N out channels] += A[M][K] * Wt[N][K],
// K = taps * Cin walked in 64-channel blocks (one 128-byte swizzled row per pixel / per out channel).
// The 128 pixels of a tile are a TB x TH x TW box of the [B,H,W] grid, so a TMA 4-D box load at
// (c0, w0+dx, h0+dy, b0) with zero OOB fill *is* the im2col tile for tap (dy,dx).
//
// Two launch forms of the same body (PAIR template parameter, ptx.cuh "CTA pair"):
//   conv_umma_kernel<N>   one CTA per tile;
//   conv_umma2_kernel<N>  a CTA pair (cta_group::2) takes two adjacent pixel tiles of the SAME column block and issues one
//                         M = 256 MMA per step; each CTA loads its own A box and only HALF of the weight tile, so the B-operand
//                         shared-memory fill and reads per SM halve and the ring is deeper (32 KB instead of 48 KB per stage at
//                         BLOCK_N = 256). A missing odd tile is loaded at an out-of-range image index (zero fill) and its stores
//                         are clipped away by the tensor map. Outputs are bit-identical to the 1-CTA form.
//
// Warp roles (320 threads): warp 0 = TMA producer (one elected thread), warp 1 = TMEM owner + MMA
// issuer (one elected thread), warps 2..9 = epilogue: two warps per TMEM lane quarter. Persistent over
// tiles; two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
// Epilogue work is cut into warp-private units of 32 rows x 64 columns (epilogue.cuh): registers ->
// 4 KB swizzled staging tile -> TMA store of the quarter's sub-box (the tensor map clips partial tiles).
#pragma once
#include "epilogue.cuh"
#include "ptx.cuh"

namespace ub {

enum : int { EPI_STORE = 0, EPI_CONVT = 1 };

struct ConvArgs {
  int B, H, W;                    // spatial grid of the GEMM rows (conv: output == input size)
  int TW, TH, TB;                 // pixel box of one tile, TW*TH*TB == 128
  int tiles_w, tiles_h, tiles_b;  // number of boxes along each axis
  int n_tiles;                    // N / BLOCK_N
  int taps;                       // 9 (3x3, pad 1) or 1 (pointwise)
  int kc0, kc1, kc2, kc3;         // 64-channel blocks taken from activation sources 0..3 (concat / ConvT-dgrad quads)
  int epi;                        // EPI_*
  int relu;
  int stages;                     // operand ring depth (shared-memory split chosen by the host)
  int sub_h, sub_b;               // rows / images of the 32-row sub-box one TMEM lane quarter covers (sub_w == TW)
  __nv_bfloat16* pool_out;        // EPI_STORE, optional: [B,H/2,W/2,Cout] = maxpool2x2(out), written from registers
  int a_bytes;                    // bytes one A box load delivers (128 rows x 128 B unless TB exceeds the batch dim)
  int Cout;                       // EPI_STORE: channels of out; EPI_CONVT: f (out channels of the ConvT)
  const float* bias;              // [Cout]
  double* stat_sum;               // EPI_STORE, optional (training): per-channel sum / sum of squares of the bf16 output,
  double* stat_sumsq;             //   accumulated atomically ([Cout] each, zeroed by the caller)
  int split;                      // 1: fp32-class path - the output tensor has 2*Cout channels [hi | lo] (epilogue.cuh
                                  //    epi_load_unit_part); inputs are such tensors too (the host lists hi|lo and hi again as
                                  //    K sources against weights [w_hi | w_hi | w_lo]). No fused pool / statistics.
  EpiBnBwd bn;                    // EPI_STORE, optional (training backward): BatchNorm-backward sums of the layer whose incoming
                                  //    gradient this conv writes, taken in the epilogue (epilogue.cuh); bn.y == null: off
};

constexpr int CONV_THREADS = 320;

template <int BLOCK_N, bool PAIR = false>
struct ConvCfg {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = (PAIR ? BLOCK_N / 2 : BLOCK_N) * 128;   // weight rows held by ONE CTA
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int MAX_STAGES = 8;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_LIMIT = 232448;
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");
  static constexpr int STG_BYTES = 8 * 4096;  // one private 4 KB staging tile per epilogue warp
  // ystg: a second set of warp-private 4 KB tiles (the y sub-boxes of the fused BatchNorm-backward sums)
  __host__ __device__ static constexpr int smem_bytes(int stages, bool ystg = false) {
    return stages * STAGE_BYTES + (ystg ? 2 : 1) * STG_BYTES + BAR_BYTES + 1024;
  }
  static int plan_stages(bool ystg = false) {
    int s = (SMEM_LIMIT - BAR_BYTES - 1024 - (ystg ? 2 : 1) * STG_BYTES) / STAGE_BYTES;
    return s > MAX_STAGES ? MAX_STAGES : s;
  }
};

template <int BLOCK_N, bool PAIR>
__device__ __forceinline__ void conv_umma_body(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmA2,
                                               const CUtensorMap& tmA3, const CUtensorMap& tmW, const CUtensorMap& tmO0,
                                               const CUtensorMap& tmO1, const CUtensorMap& tmO2, const CUtensorMap& tmO3,
                                               const CUtensorMap& tmY, const ConvArgs& a) {
  using Cfg = ConvCfg<BLOCK_N, PAIR>;
  constexpr int P = PAIR ? 2 : 1;   // CTAs that share one MMA
  constexpr int MAXS = Cfg::MAX_STAGES;
  constexpr int HALVES = BLOCK_N / 64;
  const int STAGES = a.stages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smS = smem + STAGES * Cfg::STAGE_BYTES;  // [8 warps][4 KB] private output staging
  const bool bnb = a.bn.y != nullptr;   // fused BatchNorm-backward sums: [8 warps][4 KB] y tiles follow the staging tiles
  uint8_t* smY = smS + Cfg::STG_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smY + (bnb ? Cfg::STG_BYTES : 0));
  uint64_t* full = bars;                // [MAXS] TMA -> MMA
  uint64_t* empty = bars + MAXS;        // [MAXS] MMA -> TMA
  uint64_t* tfull = bars + 2 * MAXS;    // [2] MMA -> epilogue
  uint64_t* tempty = bars + 2 * MAXS + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAXS + 4);
  uint64_t* ybars = bars + 2 * MAXS + 5;   // [8] y tile of epilogue warp w has landed

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO0);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
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
    tmem_alloc_p<PAIR>(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  cta_sync_p<PAIR>();   // (pair: the peer's barriers are initialised before anything signals them)
  tc_fence_after();
  pdl_wait();           // programmatic dependent launch: everything above overlapped the predecessor's tail
  const uint32_t tmem_base = *tmem_slot;

  const int kchunks = a.kc0 + a.kc1 + a.kc2 + a.kc3;
  const int num_kb = a.taps * kchunks;
  const int m_tiles = a.tiles_w * a.tiles_h * a.tiles_b;
  const int m_groups = (m_tiles + P - 1) / P;
  const int total_tiles = m_groups * a.n_tiles;   // work items (n_tile, m_group); this CTA's m_tile = P*m_group + rank
  const uint32_t rank = pair_rank<PAIR>();
  const bool leader = rank == 0;
  const int pair0 = blockIdx.x / P, pair_step = gridDim.x / P;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: one elected thread runs the whole loop
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair0; t < total_tiles; t += pair_step) {
        const int n_tile = t % a.n_tiles;
        const int m_tile = P * (t / a.n_tiles) + static_cast<int>(rank);   // == m_tiles for the missing odd tile:
        const int w0 = (m_tile % a.tiles_w) * a.TW;                        //    b0 lands past the batch, the box is zero-filled
        const int h0 = ((m_tile / a.tiles_w) % a.tiles_h) * a.TH;
        const int b0 = (m_tile / (a.tiles_w * a.tiles_h)) * a.TB;
        int kb = 0;
        for (int tap = 0; tap < a.taps; ++tap) {
          int dy = 0, dx = 0;
          if (a.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          for (int ch = 0; ch < kchunks; ++ch, ++kb) {
            mbar_wait_parked(&empty[stage], phase ^ 1);
            uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sB = sA + Cfg::A_BYTES;
            if (leader) mbar_expect_tx(&full[stage], P * (a.a_bytes + Cfg::B_BYTES));   // the loads of every CTA of the group
            if (ch < a.kc0) {
              tma_load_4d_p<PAIR>(sA, &tmA0, &full[stage], ch * 64, w0 + dx, h0 + dy, b0);
            } else if (ch < a.kc0 + a.kc1) {
              tma_load_4d_p<PAIR>(sA, &tmA1, &full[stage], (ch - a.kc0) * 64, w0 + dx, h0 + dy, b0);
            } else if (ch < a.kc0 + a.kc1 + a.kc2) {
              tma_load_4d_p<PAIR>(sA, &tmA2, &full[stage], (ch - a.kc0 - a.kc1) * 64, w0 + dx, h0 + dy, b0);
            } else {
              tma_load_4d_p<PAIR>(sA, &tmA3, &full[stage], (ch - a.kc0 - a.kc1 - a.kc2) * 64, w0 + dx, h0 + dy, b0);
            }
            // the weight map's box is HALF a tile (BLOCK_N/2 rows): a pair CTA loads its half, a single CTA both
            if constexpr (PAIR) {
              tma_load_2d_p<PAIR>(sB, &tmW, &full[stage], kb * 64, n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2));
            } else {
              tma_load_2d(sB, &tmW, &full[stage], kb * 64, n_tile * BLOCK_N);
              tma_load_2d(sB + Cfg::B_BYTES / 2, &tmW, &full[stage], kb * 64, n_tile * BLOCK_N + BLOCK_N / 2);
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: one elected thread runs the whole loop
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(128 * P, BLOCK_N);
      const uint64_t d_hi = make_sw128_kmajor_desc(0, 1024, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair0; t < total_tiles; t += pair_step, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES) & 0x3FFFFu;
          const uint64_t da = d_hi + (sA >> 4);
          const uint64_t db = d_hi + ((sA + Cfg::A_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // +32 bytes (16 bf16) along K inside the 128-byte swizzled row == +2 in the >>4 address field
            umma_f16_p<PAIR>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit_p<PAIR>(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_p<PAIR>(&tfull[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue: 8 warps, warp-private units (epilogue.cuh)
    // warp (q, cg): TMEM lane quarter q; of a tile's 64-column halves it takes cg, cg+2, ... (BLOCK_N == 64: the two
    // warps of a quarter alternate tiles instead, i.e. warp group cg owns accumulator stage cg).
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int tw = m % a.TW;
    const int th = (m / a.TW) % a.TH;
    const int tb = m / (a.TW * a.TH);
    uint8_t* stg = smS + (warp - 2) * 4096;
    // the quarter's 32 rows form a (TW, sub_h, sub_b) sub-box of the tile at this offset
    int off_h = 0, off_b = 0;
    if (a.TB >= 4) {
      off_b = q * a.sub_b;
    } else if (a.TB == 2) {
      off_b = q >> 1;
      off_h = (q & 1) * a.sub_h;
    } else {
      off_h = q * a.sub_h;
    }
    const bool pool_writer = ((tw | th) & 1) == 0;
    // fused batch statistics: two accumulator sets (a warp serves at most two 64-column groups of a tile), flushed
    // when the tile's column block changes (never, when gridDim.x is a multiple of n_tiles) and at the end
    float st0[4] = {0.f, 0.f, 0.f, 0.f}, st1[4] = {0.f, 0.f, 0.f, 0.f};
    int st_ntile = -1;
    // fused BatchNorm-backward sums: this warp's y tile + barrier, a cursor (yt, yit, yhf) one unit ahead of the main loop
    // (the next unit's y box is requested as soon as the walk over the current one has finished), per-channel constants and
    // accumulators for the (at most two) 64-column groups the warp serves
    uint8_t* ystg = smY + (warp - 2) * 4096;
    uint64_t* ybar = &ybars[warp - 2];
    uint32_t yph = 0;
    float bk0[8], bk1[8], ba0[4] = {0.f, 0.f, 0.f, 0.f}, ba1[4] = {0.f, 0.f, 0.f, 0.f};
    const int hf_first = HALVES == 1 ? 0 : cg;
    int yt = pair0, yit = 0, yhf = hf_first;
    auto y_skip = [&]() {   // BLOCK_N == 64: the two warps of a quarter alternate tiles
      if (HALVES == 1) {
        while (yt < total_tiles && (yit & 1) != cg) {
          yt += pair_step;
          ++yit;
        }
      }
    };
    auto y_issue = [&]() {  // lane 0: request the y sub-box of unit (yt, yhf), then advance the cursor
      if (yt >= total_tiles) return;
      const int m_tile = P * (yt / a.n_tiles) + static_cast<int>(rank);
      const int yw0 = (m_tile % a.tiles_w) * a.TW;
      const int yh0 = ((m_tile / a.tiles_w) % a.tiles_h) * a.TH;
      const int yb0 = (m_tile / (a.tiles_w * a.tiles_h)) * a.TB;
      mbar_expect_tx(ybar, a.bn.bytes);
      tma_load_4d(ystg, &tmY, ybar, (yt % a.n_tiles) * BLOCK_N + yhf * 64, yw0, yh0 + off_h, yb0 + off_b);
      if (yhf + 2 < HALVES) {
        yhf += 2;
      } else {
        yhf = hf_first;
        yt += pair_step;
        ++yit;
        y_skip();
      }
    };
    if (bnb && lane == 0) {
      y_skip();
      y_issue();
    }
    int it = 0;
    for (int t = pair0; t < total_tiles; t += pair_step, ++it) {
      const int acc = it & 1;
      if (HALVES == 1 && acc != cg) continue;
      const int n_tile = t % a.n_tiles;
      const int m_tile = P * (t / a.n_tiles) + static_cast<int>(rank);
      const int w0 = (m_tile % a.tiles_w) * a.TW;
      const int h0 = ((m_tile / a.tiles_w) % a.tiles_h) * a.TH;
      const int b0 = (m_tile / (a.tiles_w * a.tiles_h)) * a.TB;
      uint32_t vmask = 0;
      if (a.stat_sum != nullptr || bnb) {
        vmask = __ballot_sync(0xffffffffu, (w0 + tw < a.W) && (h0 + th < a.H) && (b0 + tb < a.B));
        if (n_tile != st_ntile) {
          const int c_old = st_ntile * BLOCK_N + hf_first * 64, c_new = n_tile * BLOCK_N + hf_first * 64;
          if (st_ntile >= 0 && a.stat_sum != nullptr) {
            epi_stats_flush(a.stat_sum, a.stat_sumsq, c_old, lane, st0);
            if (HALVES == 4) epi_stats_flush(a.stat_sum, a.stat_sumsq, c_old + 128, lane, st1);
          }
          if (bnb) {
            if (st_ntile >= 0) {
              epi_bnbwd_flush(a.bn, c_old, lane, ba0);
              if (HALVES == 4) epi_bnbwd_flush(a.bn, c_old + 128, lane, ba1);
            }
            epi_bnbwd_consts(a.bn, c_new, lane, bk0);
            if (HALVES == 4) epi_bnbwd_consts(a.bn, c_new + 128, lane, bk1);
          }
          st_ntile = n_tile;
        }
      }
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int hf = (HALVES == 1 ? 0 : cg); hf < HALVES; hf += 2) {
        const int n = n_tile * BLOCK_N + hf * 64;  // first GEMM column of the unit
        int bias_off = n, quad = 0, co = n;
        if (a.epi == EPI_CONVT) {  // column n -> (quad = dy*2+dx, co): bias is per co, each quad is its own store view
          quad = n / a.Cout;
          co = n - quad * a.Cout;
          bias_off = co;
        }
        uint32_t p[32];
        if (a.split) {
          // hi part first (stored at channel co), the lo part below (stored at channel Cout + co)
          epi_load_unit_part(t_row + hf * 64, a.bias + bias_off, a.relu, 0, p);
          if (lane == 0) bulk_wait_group_read<0>();
          __syncwarp();
          epi_stage_row(stg, lane, p);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const CUtensorMap* mo = quad == 0 ? &tmO0 : quad == 1 ? &tmO1 : quad == 2 ? &tmO2 : &tmO3;
            tma_store_4d(mo, stg, co, w0, h0 + off_h, b0 + off_b);
            bulk_commit_group();
          }
          epi_load_unit_part(t_row + hf * 64, a.bias + bias_off, a.relu, 1, p);
          co += a.Cout;
        } else {
          epi_load_unit(t_row + hf * 64, a.bias + bias_off, a.relu, p);
        }
        if (hf + 2 >= HALVES) {
          // this warp's last TMEM read of the tile: hand the accumulator stage back to the MMA warp
          tc_fence_before();
          mbar_arrive_p<PAIR>(&tempty[acc]);
        }
        if (lane == 0) bulk_wait_group_read<0>();  // the previous unit's TMA store has finished reading the staging tile
        __syncwarp();
        epi_stage_row(stg, lane, p);
        fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          const CUtensorMap* mo = quad == 0 ? &tmO0 : quad == 1 ? &tmO1 : quad == 2 ? &tmO2 : &tmO3;
          tma_store_4d(mo, stg, co, w0, h0 + off_h, b0 + off_b);
          bulk_commit_group();
        }
        if (a.pool_out != nullptr) {
          // 2x2 window partners are lane^1 (w) and lane^TW (h): both inside this warp's quarter
          epi_pool2x2(p, a.TW);
          const int w = w0 + tw, h = h0 + th, b = b0 + tb;
          if (pool_writer && w < a.W && h < a.H && b < a.B) {
            uint4* dst = reinterpret_cast<uint4*>(
                a.pool_out + ((static_cast<size_t>(b) * (a.H >> 1) + (h >> 1)) * (a.W >> 1) + (w >> 1)) * a.Cout + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
          }
        }
        // (after the pool: the accumulator registers p[] are dead here, the staging tile still holds the unit)
        if (a.stat_sum != nullptr) {
          if (hf == hf_first) {
            epi_stats_accumulate(stg, lane, vmask, st0);
          } else {
            epi_stats_accumulate(stg, lane, vmask, st1);
          }
        }
        if (bnb) {
          mbar_wait(ybar, yph);
          yph ^= 1;
          if (hf == hf_first) {
            epi_bnbwd_accumulate(stg, ystg, lane, vmask, bk0, ba0);
          } else {
            epi_bnbwd_accumulate(stg, ystg, lane, vmask, bk1, ba1);
          }
          __syncwarp();             // every lane has finished reading the y tile: the next box may land in it
          if (lane == 0) y_issue();
        }
      }
    }
    if (st_ntile >= 0) {
      const int c0 = st_ntile * BLOCK_N + hf_first * 64;
      if (a.stat_sum != nullptr) {
        epi_stats_flush(a.stat_sum, a.stat_sumsq, c0, lane, st0);
        if (HALVES == 4) epi_stats_flush(a.stat_sum, a.stat_sumsq, c0 + 128, lane, st1);
      }
      if (bnb) {
        epi_bnbwd_flush(a.bn, c0, lane, ba0);
        if (HALVES == 4) epi_bnbwd_flush(a.bn, c0 + 128, lane, ba1);
      }
    }
    if (lane == 0) bulk_wait_group_read<0>();
  }

  tc_fence_before();
  cta_sync_p<PAIR>();   // (pair: the leader's MMAs read the peer's shared memory - nobody leaves before both are done)
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_p<PAIR>(tmem_base, Cfg::TMEM_COLS);
  }
}



#define UB_CONV_UMMA_PARAMS                                                                                              \
  const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,                                    \
      const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,                                \
      const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO0,                                 \
      const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmO2,                                \
      const __grid_constant__ CUtensorMap tmO3, const __grid_constant__ CUtensorMap tmY /* EpiBnBwd y boxes */,        \
      const ConvArgs a

template <int BLOCK_N>
__global__ void __launch_bounds__(CONV_THREADS, 1) conv_umma_kernel(UB_CONV_UMMA_PARAMS) {
  pdl_launch();
  conv_umma_body<BLOCK_N, false>(tmA0, tmA1, tmA2, tmA3, tmW, tmO0, tmO1, tmO2, tmO3, tmY, a);
}

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV_THREADS, 1) conv_umma2_kernel(UB_CONV_UMMA_PARAMS) {
  pdl_launch();
  conv_umma_body<BLOCK_N, true>(tmA0, tmA1, tmA2, tmA3, tmW, tmO0, tmO1, tmO2, tmO3, tmY, a);
}

}  // namespace ub
