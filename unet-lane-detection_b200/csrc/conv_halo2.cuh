// CTA-pair (cta_group::2) version of conv_halo.cuh: two CTAs of a cluster (two SMs of a TPC) run ONE
// tcgen05.mma of M = 256 per issue - each SM multiplies its own 128-pixel tile, and each SM holds only HALF of
// the weight tile (BLOCK_N/2 rows), which the pair shares.
//
// Why: profiles/r1_ncu_full_dec3c0.txt and tools/mma_rate_probe.cu show the 1-CTA halo kernel is bound by the
// shared-memory operand reads of the tensor pipe (4 KB of A + BLOCK_N*32 B of B per 128xBLOCK_Nx16 MMA against
// 128 B/cycle): 48 cycles for N = 64 (math floor 32), 64 for N = 128 (floor 64, no slack). With the B operand
// split over the pair each SM reads 4 KB + BLOCK_N*16 B: 40 cycles for N = 64, 48 (< 64) for N = 128.
//
// Protocol (same roles as conv_halo.cuh; warp 0 = TMA producer in BOTH CTAs, warp 1 = MMA issuer in the LEADER
// (cluster rank 0) only, warps 2..9 = epilogue in both):
//   a_full / b_full  live in the leader: both CTAs' TMA loads complete_tx on the leader's barrier
//                    (cp.async.bulk.tensor ... .cta_group::2, barrier address mapped to rank 0 with mapa),
//                    the leader's producer expects the bytes of both.
//   a_empty / b_empty / tfull  exist in both CTAs: the leader's tcgen05.commit multicasts the arrive to both.
//   tempty           lives in the leader: the epilogue threads of both CTAs arrive on it (remote mbarrier.arrive).
// Tiles: pair p works on tiles 2*(p + i*pairs) + rank; a missing odd tile is loaded fully out of bounds (zeros)
// and its epilogue stores nothing.
#pragma once
#include "conv_halo.cuh"

namespace ub {

// shared::cluster address of the same shared-memory offset in cluster rank 0 (the leader CTA of the pair)
__device__ __forceinline__ uint32_t leader_addr(const void* p) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(0));
  return r;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all earlier MMAs of this thread are done) on the barrier at this offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// TMA loads into this CTA's shared memory whose completion is signalled on the LEADER's barrier (same offset).
__device__ __forceinline__ void tma2_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_addr(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// Arrive on the leader's copy of `bar` from either CTA of the pair. No cluster-scope release: the only thing handed over is
// "my tcgen05.ld of this accumulator stage has completed" (tcgen05.wait::ld + tcgen05.fence::before_thread_sync precede it);
// a .release.cluster arrive compiles to ERRBAR + a membar that waits for every earlier global store of the warp (the pool
// stores) - 12 % of all stall samples in profiles/r1_ncu_full_halo2.txt, on the path that frees the TMEM stage.
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar)) : "memory");
}

// Shared memory of the pair kernel: a weight tile is BLOCK_N/2 rows here.
__host__ __device__ constexpr int halo2_smem_bytes(int block_n, int a_stages, int b_stages, int head) {
  return a_stages * HaloCfg::A_STAGE_PITCH + b_stages * 3 * (block_n / 2) * 128 + (head ? 0 : HaloCfg::STG_BYTES) +
         HaloCfg::BAR_BYTES + 1024;
}

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HALO_THREADS, 1)
conv_halo2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmWh /* box = BLOCK_N/2 rows */,
                  const __grid_constant__ CUtensorMap tmOut, const HaloArgs a) {
  constexpr int B_HALF = (BLOCK_N / 2) * 128;   // bytes of one (tap, 64-channel block) weight tile in ONE CTA
  constexpr int HALVES = BLOCK_N / 64;
  constexpr int TMEM_COLS = 2 * BLOCK_N;
  const int AS = a.a_stages, BS = a.b_stages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smA + AS * HaloCfg::A_STAGE_PITCH;
  uint8_t* smS = smB + BS * 3 * B_HALF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smS + (a.epi == HEPI_STORE ? HaloCfg::STG_BYTES : 0));
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + HaloCfg::MAX_A;
  uint64_t* b_full = a_empty + HaloCfg::MAX_A;
  uint64_t* b_empty = b_full + HaloCfg::MAX_B;
  uint64_t* tfull = b_empty + HaloCfg::MAX_B;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
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
      mbar_init(&tempty[s], 2 * (HALVES == 1 ? 128 : 256));   // the epilogue threads of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int KC = a.kc0 + a.kc1;
  const int tiles_per_img = a.tiles_w * a.tiles_h;
  const int total_tiles = tiles_per_img * a.B;
  const int pairs = gridDim.x >> 1;
  const int pair = blockIdx.x >> 1;
  // iteration i of this pair covers tiles 2*(pair + i*pairs) and +1; it runs while the first of the two exists
  const int t_first = 2 * pair + static_cast<int>(rank);
  const int t_step = 2 * pairs;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs; the leader also posts expect_tx)
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool first = true;
      for (int t = t_first; t - static_cast<int>(rank) < total_tiles; t += t_step) {
        // a tile index past the end (odd tile count): image index == B is fully out of bounds -> the box is zero-filled
        const int b = t < total_tiles ? t / tiles_per_img : a.B;
        const int ti = t < total_tiles ? t - b * tiles_per_img : 0;
        const int w0 = (ti % a.tiles_w) * 8;
        const int h0 = (ti / a.tiles_w) * 16;
        for (int c = 0; c < KC; ++c) {
          mbar_wait_parked(&a_empty[as], aph ^ 1);
          if (leader) mbar_expect_tx(&a_full[as], 2 * HaloCfg::A_STAGE_BYTES);
          if (c < a.kc0) {
            tma2_load_4d(smA + as * HaloCfg::A_STAGE_PITCH, &tmA0, &a_full[as], c * 64, w0 - 1, h0 - 1, b);
          } else {
            tma2_load_4d(smA + as * HaloCfg::A_STAGE_PITCH, &tmA1, &a_full[as], (c - a.kc0) * 64, w0 - 1, h0 - 1, b);
          }
          if (++as == AS) {
            as = 0;
            aph ^= 1;
          }
          if (!a.resident || first) {
            for (int r = 0; r < 3; ++r) {  // one kernel row (3 taps) per weight stage; this CTA's half of the rows
              mbar_wait(&b_empty[bs], bph ^ 1);
              if (leader) mbar_expect_tx(&b_full[bs], 2 * 3 * B_HALF);
#pragma unroll
              for (int sx = 0; sx < 3; ++sx) {
                tma2_load_2d(smB + (bs * 3 + sx) * B_HALF, &tmWh, &b_full[bs], ((r * 3 + sx) * KC + c) * 64,
                             static_cast<int>(rank) * (BLOCK_N / 2));
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
      constexpr uint32_t idesc = make_idesc_bf16_f32(256, BLOCK_N);
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
                umma2_f16(d_tmem, da0 + (((tap / 3) * 10 + (tap % 3)) * 8 + k * 2), dbc + (tap * (B_HALF >> 4) + k * 2), idesc,
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
                  umma2_f16(d_tmem, da0 + ((r * 10 + sx) * 8 + k * 2), db0 + (sx * (B_HALF >> 4) + k * 2), idesc,
                            (c | r | sx | k) != 0);
                }
              }
              if (!a.resident) {
                umma2_commit_both(&b_empty[bs]);
                if (++bs == BS) {
                  bs = 0;
                  bph ^= 1;
                }
              }
            }
          }
          umma2_commit_both(&a_empty[as]);
          if (++as == AS) {
            as = 0;
            aph ^= 1;
          }
        }
        umma2_commit_both(&tfull[acc]);
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
    int it = 0;
    for (int t = t_first; t - static_cast<int>(rank) < total_tiles; t += t_step, ++it) {
      const int acc = it & 1;
      if (HALVES == 1 && acc != cg) continue;
      const int hf = (HALVES == 1) ? 0 : cg;
      const int n = hf * 64;
      const bool live = t < total_tiles;
      const int b = live ? t / tiles_per_img : 0;
      const int ti = live ? t - b * tiles_per_img : 0;
      const int w0 = (ti % a.tiles_w) * 8;
      const int h0 = (ti / a.tiles_w) * 16;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      uint32_t p[32];
      epi_load_unit(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + n, a.bias + n, a.relu, p);
      tc_fence_before();
      mbar_arrive_leader(&tempty[acc]);
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
        if (a.stat_sum != nullptr) epi_stats_accumulate(stg, lane, __ballot_sync(0xffffffffu, valid), st0);
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
    if (a.stat_sum != nullptr && a.epi == HEPI_STORE) {
      epi_stats_flush(a.stat_sum, a.stat_sumsq, (HALVES == 1 ? 0 : cg) * 64, lane, st0);
    }
    if (lane == 0) bulk_wait_group_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the leader's MMAs read the peer's shared memory: nobody leaves before both are done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}

}  // namespace ub
