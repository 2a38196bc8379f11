// Warp-private epilogue shared by the tcgen05 kernels.
//
// Work unit = 32 accumulator rows (one TMEM lane quarter) x 64 columns. A unit is owned by ONE warp:
// tcgen05.ld -> +bias -> ReLU -> bf16 -> 4 KB 128B-swizzled staging tile private to the warp ->
// fence.proxy.async -> __syncwarp -> lane 0 issues the TMA store of that 32-row sub-box (and of the
// 2x2 max-pooled 8-row sub-box). No CTA-wide barrier is involved, so the eight epilogue warps drift
// freely and TMA stores of one unit overlap the TMEM reads of the next.
#pragma once
#include "ptx.cuh"

namespace ub {

struct EpiWarp {
  uint8_t* stg;    // 4 KB staging tile of this warp (1024-byte aligned)
  uint8_t* pstg;   // 1 KB pooled staging tile of this warp (1024-byte aligned), or nullptr
  int lane;
};

// Reads the 64 columns starting at `taddr` (this warp's lane quarter), applies bias (+ReLU) and packs to bf16.
// p[0..15] = columns 0..31, p[16..31] = columns 32..63 (two bf16 per register).
__device__ __forceinline__ void epi_load_unit(uint32_t taddr, const float* __restrict__ bias64, int relu, uint32_t (&p)[32]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(taddr + c * 32, v);
    const float4* bias4 = reinterpret_cast<const float4*>(bias64 + c * 32);
    float4 bb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bb[j] = __ldg(bias4 + j);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x0 = __uint_as_float(v[4 * j + 0]) + bb[j].x;
      const float x1 = __uint_as_float(v[4 * j + 1]) + bb[j].y;
      const float x2 = __uint_as_float(v[4 * j + 2]) + bb[j].z;
      const float x3 = __uint_as_float(v[4 * j + 3]) + bb[j].w;
      if (relu) {   // warp-uniform
        p[c * 16 + 2 * j] = pack_bf16x2_relu(x0, x1);
        p[c * 16 + 2 * j + 1] = pack_bf16x2_relu(x2, x3);
      } else {
        p[c * 16 + 2 * j] = pack_bf16x2(x0, x1);
        p[c * 16 + 2 * j + 1] = pack_bf16x2(x2, x3);
      }
    }
  }
}

// Split-precision variant (fp32-class path, conv_umma.cuh `split`): the fp32 value x after bias (+ReLU) leaves as TWO bf16
// numbers, hi = bf16(x) and lo = bf16(x - hi), i.e. 16 mantissa bits; the next layer multiplies hi and lo separately.
// Two TMEM reads of the same columns instead of 64 live fp32 registers. part 0 -> hi, part 1 -> lo.
__device__ __forceinline__ void epi_load_unit_part(uint32_t taddr, const float* __restrict__ bias64, int relu, int part,
                                                   uint32_t (&p)[32]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(taddr + c * 32, v);
    const float4* bias4 = reinterpret_cast<const float4*>(bias64 + c * 32);
    float4 bb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bb[j] = __ldg(bias4 + j);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x[4] = {__uint_as_float(v[4 * j + 0]) + bb[j].x, __uint_as_float(v[4 * j + 1]) + bb[j].y,
                    __uint_as_float(v[4 * j + 2]) + bb[j].z, __uint_as_float(v[4 * j + 3]) + bb[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (relu) x[k] = fmaxf(x[k], 0.f);
        if (part) x[k] -= __bfloat162float(__float2bfloat16_rn(x[k]));
      }
      p[c * 16 + 2 * j] = pack_bf16x2(x[0], x[1]);
      p[c * 16 + 2 * j + 1] = pack_bf16x2(x[2], x[3]);
    }
  }
}

// Row `lane` of the warp's staging tile <- 64 bf16 (128 B), 16-byte chunks XOR-swizzled by (row & 7) as TMA expects.
__device__ __forceinline__ void epi_stage_row(uint8_t* tile, int row, const uint32_t (&p)[32]) {
  const uint32_t base = smem_u32(tile + row * 128);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    st_shared_v4(base + ((j ^ (row & 7)) << 4), p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
  }
}

// ---- fused BatchNorm batch statistics (training forward) -------------------------------------------------------------
// Per-channel sum and sum of squares of the unit that was just staged, taken from the bf16-ROUNDED values in the staging
// tile (the statistics BatchNorm will normalise with are those of the tensor it actually reads). Lane L owns channels
// 2L, 2L+1 of the 64-column unit and walks the 32 rows; rows whose pixel lies outside the tensor (clipped by the TMA store)
// are skipped through the warp-uniform `valid_mask` (bit r = row r is a real pixel). acc = {s0, ss0, s1, ss1}.
__device__ __forceinline__ void epi_stats_accumulate(const uint8_t* tile, int lane, uint32_t valid_mask, float (&acc)[4]) {
  const uint32_t base = smem_u32(tile) + (lane & 3) * 4;
  const int chunk = lane >> 2;
  // eight rows per trip: the eight loads go out back to back and the arithmetic follows (one load per row behind a per-row
  // branch left every load's full shared-memory latency exposed: the walk cost ~2 k cycles per unit, ncu short-scoreboard
  // stalls on the unpack instructions); rows outside the tensor are masked AFTER the load (their staged values are finite)
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += 8) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[i]) : "r"(base + (r0 + i) * 128 + ((chunk ^ i) << 4)));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool valid = (valid_mask >> (r0 + i)) & 1u;
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      const float lo = valid ? __low2float(h) : 0.f, hi = valid ? __high2float(h) : 0.f;
      acc[0] += lo;
      acc[1] = fmaf(lo, lo, acc[1]);
      acc[2] += hi;
      acc[3] = fmaf(hi, hi, acc[3]);
    }
  }
}
// Adds the lane's partial sums for channels col + 2*lane, +1 to the global double accumulators and clears them.
__device__ __forceinline__ void epi_stats_flush(double* __restrict__ sum, double* __restrict__ sumsq, int col, int lane,
                                                float (&acc)[4]) {
  const int c = col + 2 * lane;
  atomicAdd(sum + c, static_cast<double>(acc[0]));
  atomicAdd(sumsq + c, static_cast<double>(acc[1]));
  atomicAdd(sum + c + 1, static_cast<double>(acc[2]));
  atomicAdd(sumsq + c + 1, static_cast<double>(acc[3]));
  acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
}

// ---- fused BatchNorm-backward sums (training backward, dgrad epilogues) ------------------------------------------------
// The dgrad conv of layer L+1 writes g = dL/da of layer L, a = relu(bn(y)). BatchNorm's backward needs, per channel,
// s1 = sum g*[a > 0] and s2 = sum g*[a > 0]*xhat before it can turn g into dL/dy (train_kernels.cuh bn_relu_bwd_apply_kernel).
// Instead of a separate pass over g and y (bn_relu_bwd_reduce_kernel) the epilogue that PRODUCES g takes the sums: the unit's
// y sub-box is TMA-loaded into a second warp-private 4 KB tile (same box and swizzle as the output staging tile) and the
// column walk of epi_stats_accumulate reads both. g is read back from the staging tile, i.e. bf16-rounded as stored.
struct EpiBnBwd {
  const void* y;        // raw conv output of layer L (null: no fusion); only tested for null here, the data comes through tmY
  const float* scale;   // per channel: gamma * invstd
  const float* shift;   // beta - mean * scale
  const float* mean;
  const float* invstd;
  float* s1;            // [C] += sum g       (fp32 atomics, one per lane and channel at the end of the kernel)
  float* s2;            // [C] += sum g * xhat
  int bytes;            // bytes one y box load delivers (4096 unless the box is clipped along the batch axis)
};
// k = {scale, shift, mean, invstd} of channel col + 2*lane, then of col + 2*lane + 1
__device__ __forceinline__ void epi_bnbwd_consts(const EpiBnBwd& b, int col, int lane, float (&k)[8]) {
  const int c = col + 2 * lane;
  const float2 sc = __ldg(reinterpret_cast<const float2*>(b.scale + c));
  const float2 sh = __ldg(reinterpret_cast<const float2*>(b.shift + c));
  const float2 mu = __ldg(reinterpret_cast<const float2*>(b.mean + c));
  const float2 is = __ldg(reinterpret_cast<const float2*>(b.invstd + c));
  k[0] = sc.x; k[1] = sh.x; k[2] = mu.x; k[3] = is.x;
  k[4] = sc.y; k[5] = sh.y; k[6] = mu.y; k[7] = is.y;
}
// acc = {s1, s2} of channel 2*lane, then of channel 2*lane + 1
__device__ __forceinline__ void epi_bnbwd_accumulate(const uint8_t* gtile, const uint8_t* ytile, int lane, uint32_t valid_mask,
                                                     const float (&k)[8], float (&acc)[4]) {
  const uint32_t off = (lane & 3) * 4;
  const uint32_t gbase = smem_u32(gtile) + off, ybase = smem_u32(ytile) + off;
  const int chunk = lane >> 2;
  // eight rows per trip, all sixteen loads first (see epi_stats_accumulate); invalid rows are masked after the load
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += 8) {
    uint32_t wg[8], wy[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t o = (r0 + i) * 128 + ((chunk ^ i) << 4);
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wg[i]) : "r"(gbase + o));
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wy[i]) : "r"(ybase + o));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool valid = (valid_mask >> (r0 + i)) & 1u;
      const __nv_bfloat162 hg = *reinterpret_cast<const __nv_bfloat162*>(&wg[i]);
      const __nv_bfloat162 hy = *reinterpret_cast<const __nv_bfloat162*>(&wy[i]);
      // (an invalid row's y is whatever the box load left there: keep it out of the arithmetic altogether - 0 * NaN is NaN)
      const float y0 = valid ? __low2float(hy) : 0.f, y1 = valid ? __high2float(hy) : 0.f;
      const float g0 = (valid && fmaf(y0, k[0], k[1]) > 0.f) ? __low2float(hg) : 0.f;
      const float g1 = (valid && fmaf(y1, k[4], k[5]) > 0.f) ? __high2float(hg) : 0.f;
      acc[0] += g0;
      acc[1] = fmaf(g0, (y0 - k[2]) * k[3], acc[1]);
      acc[2] += g1;
      acc[3] = fmaf(g1, (y1 - k[6]) * k[7], acc[3]);
    }
  }
}
__device__ __forceinline__ void epi_bnbwd_flush(const EpiBnBwd& b, int col, int lane, float (&acc)[4]) {
  const int c = col + 2 * lane;
  atomicAdd(b.s1 + c, acc[0]);
  atomicAdd(b.s2 + c, acc[1]);
  atomicAdd(b.s1 + c + 1, acc[2]);
  atomicAdd(b.s2 + c + 1, acc[3]);
  acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
}

// 2x2 max over the window partners lane^1 (w) and lane^xor_h (h); afterwards every lane of a window holds the max.
__device__ __forceinline__ void epi_pool2x2(uint32_t (&p)[32], int xor_h) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    uint32_t x = p[j];
    x = bf16x2_max(x, __shfl_xor_sync(0xffffffffu, x, 1));
    x = bf16x2_max(x, __shfl_xor_sync(0xffffffffu, x, xor_h));
    p[j] = x;
  }
}

}  // namespace ub
