// Bandwidth-bound and SIMT kernels around the tensor-core convolutions:
// weight packing (BN fold), layout conversion, uint8 preprocess (cv2-exact bilinear + normalise),
// the Cin<=4 stem convolution, the 1x1 head + sigmoid + threshold mask, stand-alone 2x2 max-pool.
#pragma once
#include "ptx.cuh"

namespace ub {

// ------------------------------------------------------------------------------------------------
// Weight packing. Reference layouts: Conv2d weight [Cout][Cin][3][3] (README.md:1452,1455),
// BatchNorm2d eval fold  w' = w * g/sqrt(var+eps),  b' = beta - mean * g/sqrt(var+eps),
// ConvTranspose2d weight [Cin][Cout][2][2] + bias (README.md:1441-1443).
// ------------------------------------------------------------------------------------------------

// -> wp[Cout][9][Cin] bf16 (K index = tap*Cin + ci), bias[Cout] fp32. gamma==nullptr: no BN (scale 1, bias 0).
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                    const float* __restrict__ var, float eps, int Cout, int Cin,
                                    __nv_bfloat16* __restrict__ wp, float* __restrict__ bias) {
  pdl_enter();
  const int total = Cout * 9 * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % Cin;
    const int tap = (i / Cin) % 9;
    const int co = i / (9 * Cin);
    float s = 1.f;
    if (gamma != nullptr) s = gamma[co] / sqrtf(var[co] + eps);
    wp[i] = __float2bfloat16_rn(w[(static_cast<size_t>(co) * Cin + ci) * 9 + tap] * s);
  }
  for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < Cout; co += gridDim.x * blockDim.x) {
    float bv = 0.f;
    if (gamma != nullptr) bv = beta[co] - mean[co] * (gamma[co] / sqrtf(var[co] + eps));
    bias[co] = bv;
  }
}

// Zero-padded variants used by the plan when a logical channel count is not a multiple of 64 (the deployed topology
// features [32,64,128]): the tensor is stored with 64-aligned channels whose extra entries are
// exactly zero (zero weights and zero bias produce relu(0) = 0), so results are unchanged.
// conv: logical w[Cout_l][C0_l + C1_l][3][3], sources 0 / 1 padded to C0_p / C1_p -> wp[Cout_p][9][C0_p + C1_p].
__global__ void pack_conv3x3_pad_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, const float* __restrict__ mean,
                                        const float* __restrict__ var, float eps, int Cout_l, int C0_l, int C1_l, int Cout_p,
                                        int C0_p, int C1_p, __nv_bfloat16* __restrict__ wp, float* __restrict__ bias) {
  pdl_enter();
  const int Cin_p = C0_p + C1_p, Cin_l = C0_l + C1_l;
  const int total = Cout_p * 9 * Cin_p;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cp = i % Cin_p;
    const int tap = (i / Cin_p) % 9;
    const int co = i / (9 * Cin_p);
    int ci = -1;
    if (cp < C0_p) {
      if (cp < C0_l) ci = cp;
    } else if (cp - C0_p < C1_l) {
      ci = C0_l + (cp - C0_p);
    }
    float v = 0.f;
    if (co < Cout_l && ci >= 0) {
      float s = 1.f;
      if (gamma != nullptr) s = gamma[co] / sqrtf(var[co] + eps);
      v = w[(static_cast<size_t>(co) * Cin_l + ci) * 9 + tap] * s;
    }
    wp[i] = __float2bfloat16_rn(v);
  }
  for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < Cout_p; co += gridDim.x * blockDim.x) {
    float bv = 0.f;
    if (gamma != nullptr && co < Cout_l) bv = beta[co] - mean[co] * (gamma[co] / sqrtf(var[co] + eps));
    bias[co] = bv;
  }
}
// ConvT: logical w[Cin_l][f_l][2][2] -> wp[(quad*f_p + co)][Cin_p], bias_p[f_p] (bias may be null: not written).
__global__ void pack_convT_pad_kernel(const float* __restrict__ w, const float* __restrict__ b, int Cin_l, int f_l, int Cin_p,
                                      int f_p, __nv_bfloat16* __restrict__ wp, float* __restrict__ bias_p) {
  pdl_enter();
  const int total = 4 * f_p * Cin_p;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % Cin_p;
    const int n = i / Cin_p;
    const int co = n % f_p;
    const int quad = n / f_p;
    float v = 0.f;
    if (ci < Cin_l && co < f_l) v = w[(static_cast<size_t>(ci) * f_l + co) * 4 + quad];
    wp[i] = __float2bfloat16_rn(v);
  }
  if (bias_p != nullptr) {
    for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < f_p; co += gridDim.x * blockDim.x) bias_p[co] = co < f_l ? b[co] : 0.f;
  }
}

// Stem (Cin <= 4): -> ws[9][4][Cout] fp32 (zero for ci >= Cin), bias[Cout].
__global__ void pack_stem_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const float* __restrict__ mean,
                                 const float* __restrict__ var, float eps, int Cout, int Cin,
                                 float* __restrict__ ws, float* __restrict__ bias, int keep_fp32, int Cout_l) {
  // Cout_l <= Cout: logical width of the reference tensors; the channels above it are stored as exact zeros
  pdl_enter();
  const int total = 9 * 4 * Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % Cout;
    const int ci = (i / Cout) % 4;
    const int tap = i / (4 * Cout);
    float s = 1.f;
    if (gamma != nullptr && co < Cout_l) s = gamma[co] / sqrtf(var[co] + eps);
    float v = 0.f;
    if (ci < Cin && co < Cout_l) {
      // round through bf16 so the stem uses the same weight precision as the tensor-core layers (keep_fp32: split path)
      v = w[(static_cast<size_t>(co) * Cin + ci) * 9 + tap] * s;
      if (!keep_fp32) v = __bfloat162float(__float2bfloat16_rn(v));
    }
    ws[i] = v;
  }
  for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < Cout; co += gridDim.x * blockDim.x) {
    float bv = 0.f;
    if (gamma != nullptr && co < Cout_l) bv = beta[co] - mean[co] * (gamma[co] / sqrtf(var[co] + eps));
    bias[co] = bv;
  }
}

// ConvTranspose2d(Cin, f, 2, 2): -> wp[(dy*2+dx)*f + co][Cin] bf16 (GEMM N = 4f, K = Cin).
__global__ void pack_convT_kernel(const float* __restrict__ w, int Cin, int f, __nv_bfloat16* __restrict__ wp) {
  pdl_enter();
  const int total = 4 * f * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % Cin;
    const int n = i / Cin;
    const int co = n % f;
    const int quad = n / f;
    wp[i] = __float2bfloat16_rn(w[(static_cast<size_t>(ci) * f + co) * 4 + quad]);
  }
}

// ------------------------------------------------------------------------------------------------
// Split-precision ("fp32-class") path: every activation and weight is carried as hi = bf16(v), lo = bf16(v - hi) (16 mantissa
// bits) and a product is evaluated as hi*hi + lo*hi + hi*lo in the fp32 TMEM accumulator (the lo*lo term, 2^-18 relative,
// is dropped). Activation tensors are [B,H,W,2C] = [hi C | lo C]; the conv kernel walks K sources (x[hi|lo] : 2C), (x[hi] : C)
// per input tensor, so the packed weights are, per tap,  [w_hi (C) | w_hi (C) | w_lo (C)]  for every source.
// conv: -> wp[Cout][9][3*(C0+C1)] bf16, bias fp32 (BN folded in fp32 BEFORE the split).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ __nv_bfloat16 split_part(float v, int part) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  return part < 2 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
}

__global__ void pack_conv3x3_split_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, const float* __restrict__ mean,
                                          const float* __restrict__ var, float eps, int Cout, int C0, int C1,
                                          __nv_bfloat16* __restrict__ wp, float* __restrict__ bias) {
  pdl_enter();
  const int Cin = C0 + C1, K3 = 3 * Cin;
  const int total = Cout * 9 * K3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % K3;
    const int tap = (i / K3) % 9;
    const int co = i / (9 * K3);
    int part, ci;
    if (k < 3 * C0) {
      part = k / C0;
      ci = k % C0;
    } else {
      part = (k - 3 * C0) / C1;
      ci = C0 + (k - 3 * C0) % C1;
    }
    float sc = 1.f;
    if (gamma != nullptr) sc = gamma[co] / sqrtf(var[co] + eps);
    wp[i] = split_part(w[(static_cast<size_t>(co) * Cin + ci) * 9 + tap] * sc, part);
  }
  for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < Cout; co += gridDim.x * blockDim.x) {
    float bv = 0.f;
    if (gamma != nullptr) bv = beta[co] - mean[co] * (gamma[co] / sqrtf(var[co] + eps));
    bias[co] = bv;
  }
}

// ConvT: w [Cin][f][2][2] -> wp[4f][3*Cin] (row n = quad*f + co; K = [w_hi | w_hi | w_lo]).
__global__ void pack_convT_split_kernel(const float* __restrict__ w, int Cin, int f, __nv_bfloat16* __restrict__ wp) {
  pdl_enter();
  const int K3 = 3 * Cin;
  const int total = 4 * f * K3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % K3;
    const int n = i / K3;
    const int co = n % f;
    const int quad = n / f;
    wp[i] = split_part(w[(static_cast<size_t>(k % Cin) * f + co) * 4 + quad], k / Cin);
  }
}

// 2x2/2 max-pool of a split tensor [B,H,W,2C]: the maximum is taken on hi + lo (fp32) and the winner's pair is copied.
__global__ void maxpool2x2_split_kernel(const uint4* __restrict__ x, int B, int H, int W, int C8, uint4* __restrict__ y) {
  pdl_enter();
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = i % C8;
    size_t r = i / C8;
    const int wo = r % Wo;
    r /= Wo;
    const int ho = r % Ho;
    const size_t b = r / Ho;
    const size_t pix0 = (b * H + 2 * ho) * W + 2 * wo;
    const size_t offs[4] = {pix0, pix0 + 1, pix0 + W, pix0 + W + 1};
    float best[8];
    __nv_bfloat16 bh[8], bl[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint4 vh = __ldg(x + offs[k] * (2 * C8) + c);
      const uint4 vl = __ldg(x + offs[k] * (2 * C8) + C8 + c);
      const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&vh);
      const __nv_bfloat16* l = reinterpret_cast<const __nv_bfloat16*>(&vl);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = __bfloat162float(h[j]) + __bfloat162float(l[j]);
        if (k == 0 || v > best[j]) {
          best[j] = v;
          bh[j] = h[j];
          bl[j] = l[j];
        }
      }
    }
    const size_t po = ((b * Ho + ho) * Wo + wo) * (2 * C8) + c;
    y[po] = *reinterpret_cast<const uint4*>(bh);
    y[po + C8] = *reinterpret_cast<const uint4*>(bl);
  }
}

// ------------------------------------------------------------------------------------------------
// NCHW fp32 [B,C<=4,H,W] -> NHWC4 bf16 [B,H,W,4] (missing channels zero). Module-boundary transform.
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc4_kernel(const float* __restrict__ x, int B, int C, int H, int W,
                                     uint2* __restrict__ y) {
  pdl_enter();
  const size_t hw = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * hw;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / hw, p = i - b * hw;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < C; ++k) c[k] = __ldg(x + (b * C + k) * hw + p);
    y[i] = make_uint2(pack_bf16x2(c[0], c[1]), pack_bf16x2(c[2], c[3]));
  }
}

// fp32-class path: the same transform without the bf16 rounding (NCHW fp32 -> NHWC4 fp32).
__global__ void nchw_to_nhwc4_f32_kernel(const float* __restrict__ x, int B, int C, int H, int W, float4* __restrict__ y) {
  pdl_enter();
  const size_t hw = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * hw;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / hw, p = i - b * hw;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < C; ++k) c[k] = __ldg(x + (b * C + k) * hw + p);
    y[i] = make_float4(c[0], c[1], c[2], c[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// Preprocess (north_star (d)): uint8 HWC frames -> bilinear resize to HxW exactly as cv2.resize
// INTER_LINEAR does on uint8 (11-bit fixed-point taps; src/unet.py:33) -> optional R<->B swap
// (BGR->RGB, src/unet_ros_node.py:310) -> (x - mean)/std (README.md:3110-3111) -> NHWC4 bf16.
// Source rows are staged in shared memory so global reads are fully coalesced 16-byte loads regardless of the
// horizontal scale.
// ------------------------------------------------------------------------------------------------
struct PreArgs {
  const uint8_t* src;  // [B, Hs, Ws, 3], row pitch in bytes
  size_t pitch;        // bytes between source rows
  size_t frame_stride; // bytes between source frames
  int B, Hs, Ws, H, W;
  int swap_rb;
  float mean[3], inv_std[3];  // in output channel order
  uint2* dst;          // [B,H,W] x (4 bf16); null when dst_f32 is given
  float4* dst_f32;     // fp32-class path: [B,H,W] x (4 fp32) instead - the normalised image without bf16 rounding
  uint8_t* dst_u8;     // optional [B,H,W,3] resized uint8 (pre-normalisation), for parity tests
};
__device__ __forceinline__ void pre_store(const PreArgs& a, size_t o, float f0, float f1, float f2) {
  if (a.dst_f32 != nullptr) {
    a.dst_f32[o] = make_float4(f0, f1, f2, 0.f);
  } else {
    a.dst[o] = make_uint2(pack_bf16x2(f0, f1), pack_bf16x2(f2, 0.f));
  }
}

// cv2 INTER_LINEAR (uint8) taps of one axis. Horizontal axis (clamp_weights): a tap outside the row is folded into its
// neighbour (fx = 0). Vertical axis: cv2 keeps the weights and only clips the ROW INDICES (resize.cpp), so border rows
// blend the same row twice with two truncations - modelled exactly.
__device__ __forceinline__ void resize_coef(int d, int dn, int sn, bool clamp_weights, int& s0, int& s1, int& a0, int& a1) {
  const double scale = static_cast<double>(sn) / dn;
  float f = static_cast<float>((d + 0.5) * scale - 0.5);
  int s = static_cast<int>(floorf(f));
  f -= s;
  if (clamp_weights) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= sn - 1) { f = 0.f; s = sn - 1; }
  }
  a1 = __float2int_rn(f * 2048.f);
  a0 = __float2int_rn((1.f - f) * 2048.f);
  s0 = min(max(s, 0), sn - 1);
  s1 = min(max(s + 1, 0), sn - 1);
}
// Vertical + horizontal blend of the four (already horizontally weighted) sums, as VResizeLinear does.
__device__ __forceinline__ int resize_blend(int h0, int h1, int by0, int by1) {
  const int v = (((by0 * (h0 >> 4)) >> 16) + ((by1 * (h1 >> 4)) >> 16) + 2) >> 2;
  return min(max(v, 0), 255);
}

// Persistent tile kernel: a CTA walks tiles of `rows_per_cta` consecutive output rows of one frame.
//   once per CTA: the horizontal taps of every output column and the vertical taps of every output row go to shared memory
//      (resize_coef costs a double division; the first version recomputed it per pixel, the second per tile);
//   per tile: 1. the source rows the tile needs are staged with 16-byte loads: the contiguous span [first, last] when it is
//                short (identity, up-scaling, down-scaling by <= 2: every row is read once), else the two rows per output row;
//             2. a thread owns output COLUMNS (its taps stay in registers) and walks the tile's rows, four at a time so that
//                48 shared-memory byte loads are in flight; a warp's stores of one row cover 256 contiguous bytes.
// profiles/: one CTA per output row with byte-wide staging reached 1.04 TB/s (r1); one CTA per 8-row tile with per-tile tables
// 0.84-1.45 TB/s; this form is the one measured in profiles/r2_pre_bench.json. Tried on top and dropped: three aligned 32-bit
// shared loads + two funnel shifts per row instead of six byte loads (fewer shared-memory wavefronts, but 137 instead of
// 119 us on 256 camera frames: the kernel is bound by instruction issue and dependent-latency chains, not by the LSU).
constexpr int PRE_ROWS = 8;
constexpr int PRE_THREADS = 256;
__host__ __device__ constexpr int pre_row_pitch(int Ws) { return ((Ws * 3 + 15) / 16) * 16 + 16; }   // + head misalignment
// dynamic shared memory: row slots + [W] x-tap table + [H] y-tap table (int4 each)
__host__ __device__ constexpr size_t pre_smem_bytes(int Ws, int W, int H, int rows) {
  return static_cast<size_t>(2 * rows) * pre_row_pitch(Ws) + static_cast<size_t>(W + H) * 16;
}

__global__ void __launch_bounds__(PRE_THREADS, 4) preprocess_u8_kernel(const PreArgs a, int rows_per_cta) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t pre_smem[];
  __shared__ int s_slot[2 * PRE_ROWS], s_src[2 * PRE_ROWS], s_off[2 * PRE_ROWS];
  __shared__ int s_nslots;
  const int R = rows_per_cta;
  const int pitch_s = pre_row_pitch(a.Ws);
  int4* xtab = reinterpret_cast<int4*>(pre_smem + static_cast<size_t>(2 * R) * pitch_s);
  int4* ytab = xtab + a.W;
  // cv::resize special cases: equal sizes copy (the taps below reduce to that), and an exact 2x2 decimation is computed
  // as INTER_AREA = (a + b + c + d + 2) >> 2
  const bool area2 = (a.Hs == 2 * a.H) && (a.Ws == 2 * a.W);
  for (int x = threadIdx.x; x < a.W; x += PRE_THREADS) {     // {3*sx0, 3*sx1, ax0, ax1}
    int sx0, sx1, ax0, ax1;
    resize_coef(x, a.W, a.Ws, true, sx0, sx1, ax0, ax1);
    if (area2) {
      sx0 = 2 * x;
      sx1 = 2 * x + 1;
    }
    xtab[x] = make_int4(3 * sx0, 3 * sx1, ax0, ax1);
  }
  for (int y = threadIdx.x; y < a.H; y += PRE_THREADS) {     // {sy0, sy1, by0, by1}
    int sy0, sy1, by0, by1;
    resize_coef(y, a.H, a.Hs, false, sy0, sy1, by0, by1);
    if (area2) {
      sy0 = 2 * y;
      sy1 = 2 * y + 1;
    }
    ytab[y] = make_int4(sy0, sy1, by0, by1);
  }
  __syncthreads();
  const int tiles_h = (a.H + R - 1) / R;
  const int total = a.B * tiles_h;
  const int row_bytes = a.Ws * 3;
  const uint8_t* src_end = a.src + static_cast<size_t>(a.B - 1) * a.frame_stride + static_cast<size_t>(a.Hs - 1) * a.pitch + row_bytes;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int b = t / tiles_h;
    const int y0 = (t - b * tiles_h) * R;
    const int nrows = min(R, a.H - y0);
    const uint8_t* frame = a.src + static_cast<size_t>(b) * a.frame_stride;
    if (threadIdx.x < 2 * PRE_ROWS) {    // slot tables: every one of these threads derives the (tiny) min / max itself
      int lo = 1 << 30, hi = -1;
      for (int r = 0; r < nrows; ++r) {
        const int4 yt = ytab[y0 + r];
        lo = min(lo, min(yt.x, yt.y));
        hi = max(hi, max(yt.x, yt.y));
      }
      const bool span = hi - lo + 1 <= 2 * R;   // span mode: slot i holds source row lo + i; sparse: two private slots per row
      const int n = span ? hi - lo + 1 : 2 * nrows;
      const int i = threadIdx.x;
      if (i < 2 * nrows) {
        const int4 yt = ytab[y0 + (i >> 1)];
        const int row = (i & 1) ? yt.y : yt.x;
        s_slot[i] = span ? row - lo : i;
        if (!span) s_src[i] = row;
      }
      if (span && i < n) s_src[i] = lo + i;
      if (i == 0) s_nslots = n;
    }
    __syncthreads();
    // stage the rows: 16-byte chunks from the aligned-down row address (never below the allocation: it is at least 16-byte
    // aligned); a chunk that would reach past the last byte of the last frame is read byte by byte instead. The (slot, chunk)
    // items are flattened and every thread issues EIGHT loads before it stores any of them: one load per row and thread,
    // each followed by its dependent shared-memory store, left too few bytes in flight to cover the HBM latency
    // (2.2 TB/s on camera frames; profiles/r2_pre_bench.json)
    const int nslots = s_nslots;
    const int chunks = (row_bytes + 30) >> 4;          // enough for any alignment of the row start; <= pitch_s / 16
    const int items = nslots * chunks;
    for (int base = 0; base < items; base += 8 * PRE_THREADS) {
      uint4 v[8];
      int dsto[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int it = base + k * PRE_THREADS + threadIdx.x;
        dsto[k] = -1;
        if (it < items) {
          const int i = it / chunks, c = it - i * chunks;
          const uint8_t* rowp = frame + static_cast<size_t>(s_src[i]) * a.pitch;
          const int off = static_cast<int>(reinterpret_cast<uintptr_t>(rowp) & 15);
          if (c == 0) s_off[i] = off;                  // where the row starts inside its slot
          const uint8_t* g = rowp - off + 16 * c;
          dsto[k] = i * pitch_s + 16 * c;
          if (g + 16 <= src_end) {
            v[k] = __ldg(reinterpret_cast<const uint4*>(g));
          } else {
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            for (int q = 0; q < 16; ++q) {
              if (g + q < src_end) w[q >> 2] |= static_cast<uint32_t>(__ldg(g + q)) << (8 * (q & 3));
            }
            v[k] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (dsto[k] >= 0) *reinterpret_cast<uint4*>(pre_smem + dsto[k]) = v[k];
      }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < a.W; x += PRE_THREADS) {
      const int4 xt = xtab[x];
#pragma unroll 4
      for (int r = 0; r < nrows; ++r) {
        const int4 yt = ytab[y0 + r];
        const int sl0 = s_slot[2 * r], sl1 = s_slot[2 * r + 1];
        const uint8_t* r0 = pre_smem + static_cast<size_t>(sl0) * pitch_s + s_off[sl0];
        const uint8_t* r1 = pre_smem + static_cast<size_t>(sl1) * pitch_s + s_off[sl1];
        int px[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (area2) {
            px[c] = (r0[xt.x + c] + r0[xt.y + c] + r1[xt.x + c] + r1[xt.y + c] + 2) >> 2;
          } else {
            const int h0 = r0[xt.x + c] * xt.z + r0[xt.y + c] * xt.w;
            const int h1 = r1[xt.x + c] * xt.z + r1[xt.y + c] * xt.w;
            px[c] = resize_blend(h0, h1, yt.z, yt.w);
          }
        }
        if (a.swap_rb) { const int tt = px[0]; px[0] = px[2]; px[2] = tt; }
        const size_t o = (static_cast<size_t>(b) * a.H + y0 + r) * a.W + x;
        if (a.dst_u8 != nullptr) {
          a.dst_u8[o * 3 + 0] = static_cast<uint8_t>(px[0]);
          a.dst_u8[o * 3 + 1] = static_cast<uint8_t>(px[1]);
          a.dst_u8[o * 3 + 2] = static_cast<uint8_t>(px[2]);
        }
        const float f0 = (px[0] - a.mean[0]) * a.inv_std[0];
        const float f1 = (px[1] - a.mean[1]) * a.inv_std[1];
        const float f2 = (px[2] - a.mean[2]) * a.inv_std[2];
        pre_store(a, o, f0, f1, f2);
      }
    }
    __syncthreads();   // the slot tables and the staged rows are free for the next tile
  }
}

// Bulk-copy form of the tile kernel (option pre_bulk, the default): the staging loop above costs about as many instructions
// as the blend (one 16-byte load, an index division and a dependent shared-memory store per chunk) and three CTA barriers
// per tile. Here one warp asks the copy engine for the tile's source rows (cp.async.bulk global -> shared, completion on
// an mbarrier; the lanes of warp 0 take one row each, or one request serves the whole span when the rows are contiguous
// in memory) and the rows of tile i+1 arrive while tile i is blended: two stages, one barrier per tile, no staging instructions at all.
// The copy engine moves 16-byte aligned chunks: a request starts at the aligned-down address of the row (never below the
// allocation, which is at least 16-byte aligned) and the row's offset inside its slot is kept in a small table; where the
// rounded-up end would pass the last byte of the last frame the requesting thread copies the (< 16) tail bytes itself.
// Same arithmetic as preprocess_u8_kernel (bit-equal to cv2, test_preprocess_bulk_kernel_*).
constexpr int PREB_THREADS = 256;
__host__ __device__ constexpr int preb_slot_pitch(int Ws) { return ((Ws * 3 + 15 + 15) >> 4) << 4; }
constexpr int PREB_MAX_STAGES = 4;
__host__ __device__ constexpr size_t preb_smem_bytes(int Ws, int W, int H, int rows, int stages = 2) {
  return static_cast<size_t>(stages) * (2 * rows) * preb_slot_pitch(Ws) + static_cast<size_t>(W + H) * 16;
}
__device__ __forceinline__ void bulk_load_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_add_tx(uint64_t* bar, uint32_t bytes) {   // expect-tx without an arrival
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// Template switches keep the per-pixel instruction count down (ncu of the first, all-runtime version: 130 instructions per
// output pixel, issue-bound at 59 % of the issue slots and 30 % occupancy; 18 % of the stall samples at the end-of-tile
// __syncthreads): AREA2 = the exact 2x2 decimation, SWAP = R<->B, GENERIC = optional outputs present (resized uint8 copy
// or fp32 NHWC4) - only then are the output pointers tested per pixel. The vertical blend's ((b * (h >> 4)) >> 16) is one
// multiply-high with b pre-shifted by 16 (exact: both factors are non-negative), and its result needs no clamp:
// (b0 + b1) <= 2049 and h >> 4 <= 32655 bound the sum by 1020, so (sum + 2) >> 2 <= 255.
// Stage hand-back without a CTA barrier: every warp arrives on the stage's `empty` mbarrier when it has finished a tile, and
// only warp 0 - before it requests the tile after next into that stage - waits for it.
template <bool AREA2, bool SWAP, bool GENERIC>
__global__ void __launch_bounds__(PREB_THREADS, 3) preprocess_bulk_u8_kernel(const PreArgs a, int rows_per_cta, int stages) {
  pdl_enter();
  extern __shared__ __align__(128) uint8_t preb_smem[];
  __shared__ __align__(8) uint64_t full[PREB_MAX_STAGES], empty[PREB_MAX_STAGES];
  __shared__ int s_off[PREB_MAX_STAGES][2 * PRE_ROWS];     // where slot i's row starts inside its stage
  const int R = rows_per_cta;
  const int S = stages;        // 2 .. PREB_MAX_STAGES: the rows of the next S - 1 tiles are in flight while one is blended
  const int row_bytes = a.Ws * 3;
  const int P = preb_slot_pitch(a.Ws);
  const int stage_bytes = 2 * R * P;
  int4* xtab = reinterpret_cast<int4*>(preb_smem + static_cast<size_t>(S) * stage_bytes);
  int4* ytab = xtab + a.W;
  for (int x = threadIdx.x; x < a.W; x += blockDim.x) {     // {3*sx0, 3*sx1, ax0, ax1}
    int sx0, sx1, ax0, ax1;
    resize_coef(x, a.W, a.Ws, true, sx0, sx1, ax0, ax1);
    if (AREA2) {
      sx0 = 2 * x;
      sx1 = 2 * x + 1;
    }
    xtab[x] = make_int4(3 * sx0, 3 * sx1, ax0, ax1);
  }
  for (int y = threadIdx.x; y < a.H; y += blockDim.x) {     // {sy0, sy1, by0 << 16, by1 << 16}
    int sy0, sy1, by0, by1;
    resize_coef(y, a.H, a.Hs, false, sy0, sy1, by0, by1);
    if (AREA2) {
      sy0 = 2 * y;
      sy1 = 2 * y + 1;
    }
    ytab[y] = make_int4(sy0, sy1, by0 << 16, by1 << 16);
  }
  const int nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 32);
      mbar_init(&empty[i], nwarps);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const int tiles_h = (a.H + R - 1) / R;
  const int total = a.B * tiles_h;
  const uint8_t* src_end = a.src + static_cast<size_t>(a.B - 1) * a.frame_stride + static_cast<size_t>(a.Hs - 1) * a.pitch + row_bytes;
  // warp 0: request the source rows of tile t into `stage`, one slot per lane (2R <= 16 slots). The row taps are
  // non-decreasing in y, so the tile needs the rows [sy0(first row), sy1(last row)]: slot i = row lo + i when that span fits
  // the stage, else two private slots per output row. Every lane arrives once on the stage's barrier (count 32) AFTER it
  // has announced its bytes, written its table entry and copied its tail bytes, so the phase ends when all of that is visible
  // and the copies have landed.
  auto request = [&](int t, int stage) {
    const int lane = threadIdx.x;
    const int b = t / tiles_h;
    const int y0 = (t - b * tiles_h) * R;
    const int nrows = min(R, a.H - y0);
    const uint8_t* frame = a.src + static_cast<size_t>(b) * a.frame_stride;
    uint8_t* sbase = preb_smem + static_cast<size_t>(stage) * stage_bytes;
    uint64_t* bar = &full[stage];
    // `nbytes` from global address g -> dst + (g & 15), dst 16-byte aligned
    auto span_in = [&](uint8_t* dst, const uint8_t* g, int nbytes) {
      const int off = static_cast<int>(reinterpret_cast<uintptr_t>(g) & 15);
      const uint8_t* g0 = g - off;
      const int want = (off + nbytes + 15) & ~15;
      int bulk = want;
      if (g0 + want > src_end) bulk = static_cast<int>((src_end - g0) & ~static_cast<ptrdiff_t>(15));
      if (bulk > 0) {
        mbar_add_tx(bar, static_cast<uint32_t>(bulk));
        bulk_load_g2s(dst, g0, static_cast<uint32_t>(bulk), bar);
      }
      for (int q = bulk; q < off + nbytes; ++q) dst[q] = __ldg(g0 + q);
    };
    const int lo = ytab[y0].x, n = ytab[y0 + nrows - 1].y - lo + 1;
    if (n <= 2 * R) {
      if (a.pitch == static_cast<size_t>(row_bytes)) {       // the rows follow each other in memory: one request
        const uint8_t* g = frame + static_cast<size_t>(lo) * a.pitch;
        if (lane == 0) span_in(sbase, g, n * row_bytes);
        if (lane < n) s_off[stage][lane] = static_cast<int>(reinterpret_cast<uintptr_t>(g) & 15) + lane * row_bytes;
      } else if (lane < n) {
        const uint8_t* g = frame + static_cast<size_t>(lo + lane) * a.pitch;
        span_in(sbase + lane * P, g, row_bytes);
        s_off[stage][lane] = lane * P + static_cast<int>(reinterpret_cast<uintptr_t>(g) & 15);
      }
    } else if (lane < 2 * nrows) {
      const int4 yt = ytab[y0 + (lane >> 1)];
      const uint8_t* g = frame + static_cast<size_t>((lane & 1) ? yt.y : yt.x) * a.pitch;
      span_in(sbase + lane * P, g, row_bytes);
      s_off[stage][lane] = lane * P + static_cast<int>(reinterpret_cast<uintptr_t>(g) & 15);
    }
    mbar_arrive(bar);
  };
  const bool producer = threadIdx.x < 32;
  const int G = gridDim.x;
  if (producer) {
    for (int i = 0; i < S - 1; ++i) {
      if (static_cast<int>(blockIdx.x) + i * G < total) request(blockIdx.x + i * G, i);
    }
  }
  uint32_t phase = 0;    // bit s: parity of the next completion of full[s]
  uint32_t ephase = 0;   // warp 0, bit s: parity of the completion of empty[s] that frees stage s for its next request
  int stage = 0;
  for (int t = blockIdx.x, it = 0; t < total; t += G, ++it, stage = (stage + 1 == S) ? 0 : stage + 1) {
    if (producer && t + (S - 1) * G < total) {
      // tile it + S - 1 goes to the stage tile it - 1 was read from (none before the first): wait until every warp has left it
      const int ps = (stage == 0) ? S - 1 : stage - 1;
      if (it > 0) {
        mbar_wait(&empty[ps], (ephase >> ps) & 1u);
        ephase ^= 1u << ps;
      }
      request(t + (S - 1) * G, ps);
    }
    const int b = t / tiles_h;
    const int y0 = (t - b * tiles_h) * R;
    const int nrows = min(R, a.H - y0);
    const int lo = ytab[y0].x;
    const bool span = ytab[y0 + nrows - 1].y - lo + 1 <= 2 * R;
    const uint8_t* base = preb_smem + static_cast<size_t>(stage) * stage_bytes;
    mbar_wait(&full[stage], (phase >> stage) & 1u);
    phase ^= 1u << stage;
    for (int x = threadIdx.x; x < a.W; x += blockDim.x) {
      const int4 xt = xtab[x];
      const uint32_t ax0 = static_cast<uint32_t>(xt.z), ax1 = static_cast<uint32_t>(xt.w);
      size_t o = (static_cast<size_t>(b) * a.H + y0) * a.W + x;
#pragma unroll 4
      for (int r = 0; r < nrows; ++r, o += a.W) {
        const int4 yt = ytab[y0 + r];
        const uint8_t* r0 = base + s_off[stage][span ? yt.x - lo : 2 * r];
        const uint8_t* r1 = base + s_off[stage][span ? yt.y - lo : 2 * r + 1];
        const uint8_t *p00 = r0 + xt.x, *p01 = r0 + xt.y, *p10 = r1 + xt.x, *p11 = r1 + xt.y;
        uint32_t px[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int cs = SWAP ? 2 - c : c;     // source channel of output channel c
          if (AREA2) {
            px[c] = (static_cast<uint32_t>(p00[cs]) + p01[cs] + p10[cs] + p11[cs] + 2u) >> 2;
          } else {
            const uint32_t h0 = p00[cs] * ax0 + p01[cs] * ax1;
            const uint32_t h1 = p10[cs] * ax0 + p11[cs] * ax1;
            px[c] = (__umulhi(static_cast<uint32_t>(yt.z), h0 >> 4) + __umulhi(static_cast<uint32_t>(yt.w), h1 >> 4) + 2u) >> 2;
          }
        }
        const float f0 = (static_cast<int>(px[0]) - a.mean[0]) * a.inv_std[0];
        const float f1 = (static_cast<int>(px[1]) - a.mean[1]) * a.inv_std[1];
        const float f2 = (static_cast<int>(px[2]) - a.mean[2]) * a.inv_std[2];
        if (GENERIC) {
          if (a.dst_u8 != nullptr) {
            a.dst_u8[o * 3 + 0] = static_cast<uint8_t>(px[0]);
            a.dst_u8[o * 3 + 1] = static_cast<uint8_t>(px[1]);
            a.dst_u8[o * 3 + 2] = static_cast<uint8_t>(px[2]);
          }
          pre_store(a, o, f0, f1, f2);
        } else {
          a.dst[o] = make_uint2(pack_bf16x2(f0, f1), pack_bf16x2(f2, 0.f));
        }
      }
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stage]);   // this warp has read the stage and its offset table
  }
}

// Same-size frames (Hs == H, Ws == W): cv2.resize returns a copy, so the preprocess is channel swap + normalise only.
// One thread per 4 pixels: three 32-bit loads (12 source bytes), two 16-byte stores (four NHWC4 bf16 pixels); a warp reads
// 384 and writes 1024 contiguous bytes; two groups per trip so that six loads are in flight per thread.
// Needs rows / frames that start on 4-byte boundaries and W % 4 == 0 (host checks).
__device__ __forceinline__ void pre_copy_group(const PreArgs& a, size_t o, uint32_t w0, uint32_t w1, uint32_t w2) {
  uint32_t px[4][3];   // 12 bytes = 4 pixels x 3 channels
  px[0][0] = w0 & 255u; px[0][1] = (w0 >> 8) & 255u; px[0][2] = (w0 >> 16) & 255u;
  px[1][0] = w0 >> 24;  px[1][1] = w1 & 255u;        px[1][2] = (w1 >> 8) & 255u;
  px[2][0] = (w1 >> 16) & 255u; px[2][1] = w1 >> 24; px[2][2] = w2 & 255u;
  px[3][0] = (w2 >> 8) & 255u;  px[3][1] = (w2 >> 16) & 255u; px[3][2] = w2 >> 24;
  uint32_t out[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (a.swap_rb) { const uint32_t t = px[k][0]; px[k][0] = px[k][2]; px[k][2] = t; }
    if (a.dst_u8 != nullptr) {
      a.dst_u8[(o + k) * 3 + 0] = static_cast<uint8_t>(px[k][0]);
      a.dst_u8[(o + k) * 3 + 1] = static_cast<uint8_t>(px[k][1]);
      a.dst_u8[(o + k) * 3 + 2] = static_cast<uint8_t>(px[k][2]);
    }
    const float f0 = (static_cast<int>(px[k][0]) - a.mean[0]) * a.inv_std[0];
    const float f1 = (static_cast<int>(px[k][1]) - a.mean[1]) * a.inv_std[1];
    const float f2 = (static_cast<int>(px[k][2]) - a.mean[2]) * a.inv_std[2];
    if (a.dst_f32 != nullptr) a.dst_f32[o + k] = make_float4(f0, f1, f2, 0.f);
    out[2 * k] = pack_bf16x2(f0, f1);
    out[2 * k + 1] = pack_bf16x2(f2, 0.f);
  }
  if (a.dst_f32 != nullptr) return;
  uint4* dp = reinterpret_cast<uint4*>(a.dst + o);
  dp[0] = make_uint4(out[0], out[1], out[2], out[3]);
  dp[1] = make_uint4(out[4], out[5], out[6], out[7]);
}

__global__ void __launch_bounds__(256) preprocess_copy_u8_kernel(const PreArgs a) {
  pdl_enter();
  const int wq = a.W >> 2;
  const size_t total = static_cast<size_t>(a.B) * a.H * wq;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  auto locate = [&](size_t i, const uint32_t*& sp, size_t& o) {
    const int xq = static_cast<int>(i % wq);
    const size_t r = i / wq;
    const int y = static_cast<int>(r % a.H);
    const size_t b = r / a.H;
    sp = reinterpret_cast<const uint32_t*>(a.src + b * a.frame_stride + static_cast<size_t>(y) * a.pitch) + 3 * xq;
    o = (b * a.H + y) * a.W + 4 * xq;
  };
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  for (; i + stride < total; i += 2 * stride) {
    const uint32_t *s0, *s1;
    size_t o0, o1;
    locate(i, s0, o0);
    locate(i + stride, s1, o1);
    const uint32_t a0 = __ldg(s0), a1 = __ldg(s0 + 1), a2 = __ldg(s0 + 2);
    const uint32_t b0 = __ldg(s1), b1 = __ldg(s1 + 1), b2 = __ldg(s1 + 2);
    pre_copy_group(a, o0, a0, a1, a2);
    pre_copy_group(a, o1, b0, b1, b2);
  }
  if (i < total) {
    const uint32_t* s0;
    size_t o0;
    locate(i, s0, o0);
    pre_copy_group(a, o0, __ldg(s0), __ldg(s0 + 1), __ldg(s0 + 2));
  }
}

// ------------------------------------------------------------------------------------------------
// IPM front end of the ROS node fused into the preprocess (src/unet_ros_node.py:297-313 + src/unet.py:24-42):
//   cv2.warpPerspective(bgr, M, (Ww, Hw))  ->  [cv2.resize to the same size = copy]  ->  BGR2RGB  ->
//   cv2.resize(rgb, (W, H))  ->  (x - mean)/std  ->  NHWC4 bf16
// The Hw x Ww bird's-eye image is never materialised: each network-input pixel needs 2x2 warped pixels, each of which
// is evaluated on the fly from 2x2 source pixels with cv2's exact arithmetic (imgwarp.cpp): destination coordinates in
// double per 64-column block, rounded to 1/32 pixel, 15-bit integer blend weights, +2^14 >> 15, BORDER_CONSTANT 0.
// m = inverse map (dst -> src) as cv::invert produces it from the matrix handed to warpPerspective.
// ------------------------------------------------------------------------------------------------
struct WarpPreArgs {
  const uint8_t* src;   // [B, Hs, Ws, 3] BGR, row pitch in bytes
  size_t pitch, frame_stride;
  int B, Hs, Ws;        // camera frame
  int Hw, Ww;           // warped (bird's-eye) image
  int H, W;             // network input
  int swap_rb;
  double m[9];
  float mean[3], inv_std[3];
  uint2* dst;           // [B,H,W] x (4 bf16)
  uint8_t* dst_u8;      // optional [B,H,W,3] resized uint8 frame (after the channel swap)
};

__device__ __forceinline__ void warp_pixel_u8(const uint8_t* __restrict__ frame, size_t pitch, int Hs, int Ws,
                                              const double* __restrict__ m, int xd, int yd, int (&out)[3]) {
  const int xb = (xd >> 6) << 6;
  const double x1 = static_cast<double>(xd - xb), xbd = static_cast<double>(xb), ydd = static_cast<double>(yd);
  // explicit round-to-nearest mul/add: no FMA contraction, the roundings must be cv2's
  const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(m[0], xbd), __dmul_rn(m[1], ydd)), m[2]);
  const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(m[3], xbd), __dmul_rn(m[4], ydd)), m[5]);
  const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(m[6], xbd), __dmul_rn(m[7], ydd)), m[8]);
  double Wd = __dadd_rn(W0, __dmul_rn(m[6], x1));
  Wd = (Wd != 0.0) ? __ddiv_rn(32.0, Wd) : 0.0;
  const double fX = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(X0, __dmul_rn(m[0], x1)), Wd)));
  const double fY = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Y0, __dmul_rn(m[3], x1)), Wd)));
  const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
  const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
  const int ax = X & 31, ay = Y & 31;
  const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
  const bool x0in = sx >= 0 && sx < Ws, x1in = sx + 1 >= 0 && sx + 1 < Ws;
  const bool y0in = sy >= 0 && sy < Hs, y1in = sy + 1 >= 0 && sy + 1 < Hs;
  const uint8_t* r0 = frame + static_cast<size_t>(y0in ? sy : 0) * pitch;
  const uint8_t* r1 = frame + static_cast<size_t>(y1in ? sy + 1 : 0) * pitch;
  const int c0 = (x0in ? sx : 0) * 3, c1 = (x1in ? sx + 1 : 0) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int v00 = (y0in && x0in) ? __ldg(r0 + c0 + c) : 0;
    const int v01 = (y0in && x1in) ? __ldg(r0 + c1 + c) : 0;
    const int v10 = (y1in && x0in) ? __ldg(r1 + c0 + c) : 0;
    const int v11 = (y1in && x1in) ? __ldg(r1 + c1 + c) : 0;
    const int acc = v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11;
    out[c] = min(max((acc + (1 << 14)) >> 15, 0), 255);
  }
}

__global__ void __launch_bounds__(256)
warp_preprocess_u8_kernel(const WarpPreArgs a) {
  pdl_enter();
  const size_t total = static_cast<size_t>(a.B) * a.H * a.W;
  const bool area2 = (a.Hw == 2 * a.H) && (a.Ww == 2 * a.W);
  double m[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) m[k] = a.m[k];
  for (size_t o = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; o < total;
       o += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(o % a.W);
    const int y = static_cast<int>((o / a.W) % a.H);
    const int b = static_cast<int>(o / (static_cast<size_t>(a.W) * a.H));
    const uint8_t* frame = a.src + b * a.frame_stride;
    int sy0, sy1, by0, by1, sx0, sx1, ax0, ax1;
    resize_coef(y, a.H, a.Hw, false, sy0, sy1, by0, by1);
    resize_coef(x, a.W, a.Ww, true, sx0, sx1, ax0, ax1);
    if (area2) {
      sy0 = 2 * y;
      sy1 = 2 * y + 1;
      sx0 = 2 * x;
      sx1 = 2 * x + 1;
    }
    int p00[3], p01[3], p10[3], p11[3], px[3];
    warp_pixel_u8(frame, a.pitch, a.Hs, a.Ws, m, sx0, sy0, p00);
    warp_pixel_u8(frame, a.pitch, a.Hs, a.Ws, m, sx1, sy0, p01);
    warp_pixel_u8(frame, a.pitch, a.Hs, a.Ws, m, sx0, sy1, p10);
    warp_pixel_u8(frame, a.pitch, a.Hs, a.Ws, m, sx1, sy1, p11);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (area2) {
        px[c] = (p00[c] + p01[c] + p10[c] + p11[c] + 2) >> 2;
      } else {
        px[c] = resize_blend(p00[c] * ax0 + p01[c] * ax1, p10[c] * ax0 + p11[c] * ax1, by0, by1);
      }
    }
    if (a.swap_rb) { const int t = px[0]; px[0] = px[2]; px[2] = t; }
    if (a.dst_u8 != nullptr) {
      a.dst_u8[o * 3 + 0] = static_cast<uint8_t>(px[0]);
      a.dst_u8[o * 3 + 1] = static_cast<uint8_t>(px[1]);
      a.dst_u8[o * 3 + 2] = static_cast<uint8_t>(px[2]);
    }
    const float f0 = (px[0] - a.mean[0]) * a.inv_std[0];
    const float f1 = (px[1] - a.mean[1]) * a.inv_std[1];
    const float f2 = (px[2] - a.mean[2]) * a.inv_std[2];
    a.dst[o] = make_uint2(pack_bf16x2(f0, f1), pack_bf16x2(f2, 0.f));
  }
}

// Stand-alone warp (the bird's-eye image itself, e.g. for display): dst [B,Hw,Ww,3] = cv2.warpPerspective(src, M).
__global__ void __launch_bounds__(256)
warp_perspective_u8_kernel(const WarpPreArgs a, uint8_t* __restrict__ dst) {
  pdl_enter();
  const size_t total = static_cast<size_t>(a.B) * a.Hw * a.Ww;
  double m[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) m[k] = a.m[k];
  for (size_t o = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; o < total;
       o += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(o % a.Ww);
    const int y = static_cast<int>((o / a.Ww) % a.Hw);
    const int b = static_cast<int>(o / (static_cast<size_t>(a.Ww) * a.Hw));
    int p[3];
    warp_pixel_u8(a.src + b * a.frame_stride, a.pitch, a.Hs, a.Ws, m, x, y, p);
    dst[o * 3 + 0] = static_cast<uint8_t>(p[0]);
    dst[o * 3 + 1] = static_cast<uint8_t>(p[1]);
    dst[o * 3 + 2] = static_cast<uint8_t>(p[2]);
  }
}

// Mask back to the source resolution (src/unet.py:70): cv2.resize(mask_u8 [Hs,Ws] -> [Hd,Wd]), INTER_LINEAR, exact.
// One thread per 4 consecutive output pixels (one 32-bit store); the 224x224 source stays in L1/L2.
__global__ void __launch_bounds__(256)
resize_gray_u8_kernel(const uint8_t* __restrict__ src, int B, int Hs, int Ws, uint8_t* __restrict__ dst, int Hd, int Wd) {
  pdl_enter();
  const int wq = (Wd + 3) / 4;
  const size_t total = static_cast<size_t>(B) * Hd * wq;
  const bool area2 = (Hs == 2 * Hd) && (Ws == 2 * Wd);
  const bool same = (Hs == Hd) && (Ws == Wd);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int xq = static_cast<int>(i % wq);
    const int y = static_cast<int>((i / wq) % Hd);
    const int b = static_cast<int>(i / (static_cast<size_t>(wq) * Hd));
    const uint8_t* s = src + static_cast<size_t>(b) * Hs * Ws;
    int sy0, sy1, by0, by1;
    resize_coef(y, Hd, Hs, false, sy0, sy1, by0, by1);
    uint8_t* drow = dst + (static_cast<size_t>(b) * Hd + y) * Wd;
    uint32_t pack = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = xq * 4 + k;
      int v = 0;
      if (x < Wd) {
        if (same) {
          v = __ldg(s + static_cast<size_t>(y) * Ws + x);
        } else if (area2) {
          const uint8_t* r0 = s + static_cast<size_t>(2 * y) * Ws + 2 * x;
          v = (__ldg(r0) + __ldg(r0 + 1) + __ldg(r0 + Ws) + __ldg(r0 + Ws + 1) + 2) >> 2;
        } else {
          int sx0, sx1, ax0, ax1;
          resize_coef(x, Wd, Ws, true, sx0, sx1, ax0, ax1);
          const uint8_t* r0 = s + static_cast<size_t>(sy0) * Ws;
          const uint8_t* r1 = s + static_cast<size_t>(sy1) * Ws;
          v = resize_blend(__ldg(r0 + sx0) * ax0 + __ldg(r0 + sx1) * ax1, __ldg(r1 + sx0) * ax0 + __ldg(r1 + sx1) * ax1, by0, by1);
        }
      }
      pack |= static_cast<uint32_t>(v) << (8 * k);
    }
    if ((Wd & 3) == 0) {
      *reinterpret_cast<uint32_t*>(drow + xq * 4) = pack;
    } else {
      for (int k = 0; k < 4 && xq * 4 + k < Wd; ++k) drow[xq * 4 + k] = static_cast<uint8_t>(pack >> (8 * k));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stem: Conv3x3(pad 1) + folded BN + ReLU for Cin <= 4 (README.md:1452 with in_channels=3).
// K = 27 is too thin for a 64-wide UMMA K block, and the layer is write-bandwidth bound
// (reads 8 B/pixel, writes 128 B/pixel), so it runs on the FP32 pipes: 16x16 pixel tile per block,
// each thread owns 2 horizontally adjacent pixels x 32 output channels, weights broadcast from smem.
// ------------------------------------------------------------------------------------------------
// MODE 0: bf16 path, x = NHWC4 bf16. MODE 1 / 2 (fp32-class path): x is fp32 - the module's own NCHW input (1) or the NHWC4
// fp32 tensor of the fp32 preprocess (2), i.e. no bf16 rounding of the image - and y is [B,H,W,2*Cout] = [hi | lo].
template <int MODE>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const void* __restrict__ xin, const float* __restrict__ ws, const float* __restrict__ bias,
                 int B, int H, int W, int Cin, int Cout, int relu, __nv_bfloat16* __restrict__ y) {
  pdl_enter();
  constexpr bool SPLIT = MODE != 0;
  const uint2* x = reinterpret_cast<const uint2*>(xin);
  extern __shared__ float sm[];
  float* sw = sm;                       // [9][4][Cout]
  float* sb = sw + 36 * Cout;           // [Cout]
  float4* st = reinterpret_cast<float4*>(sb + Cout);  // [18][18] input pixels (4 ch fp32)
  const int tiles_w = (W + 15) / 16, tiles_h = (H + 15) / 16;
  const int tile = blockIdx.x;
  const int w0 = (tile % tiles_w) * 16;
  const int h0 = ((tile / tiles_w) % tiles_h) * 16;
  const int b = tile / (tiles_w * tiles_h);
  for (int i = threadIdx.x; i < 36 * Cout; i += blockDim.x) sw[i] = ws[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias[i];
  for (int i = threadIdx.x; i < 18 * 18; i += blockDim.x) {
    const int hh = h0 + i / 18 - 1, ww = w0 + i % 18 - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
      if (MODE == 2) {
        v = __ldg(reinterpret_cast<const float4*>(xin) + (static_cast<size_t>(b) * H + hh) * W + ww);
      } else if (MODE == 1) {
        const float* xf = reinterpret_cast<const float*>(xin);
        float c4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < Cin; ++k) c4[k] = __ldg(xf + ((static_cast<size_t>(b) * Cin + k) * H + hh) * W + ww);
        v = make_float4(c4[0], c4[1], c4[2], c4[3]);
      } else {
        const uint2 r = __ldg(x + (static_cast<size_t>(b) * H + hh) * W + ww);
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&r.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
        v = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
      }
    }
    st[i] = v;
  }
  __syncthreads();
  const int cgroups = Cout / 32;
  for (int task = threadIdx.x; task < 128 * cgroups; task += blockDim.x) {
    const int cg = task / 128;
    const int pp = task % 128;
    const int px = (pp % 8) * 2, py = pp / 8;
    float acc0[32], acc1[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      acc0[j] = sb[cg * 32 + j];
      acc1[j] = acc0[j];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      // 4 input pixels of this row feed the 3 horizontal taps of both output pixels
      float4 in[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) in[k] = st[(py + r) * 18 + px + k];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const float* wt = sw + ((r * 3 + s) * 4) * Cout + cg * 32;
        const float i0[3] = {in[s].x, in[s].y, in[s].z};
        const float i1[3] = {in[s + 1].x, in[s + 1].y, in[s + 1].z};
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 wv = *reinterpret_cast<const float4*>(wt + ci * Cout + j4 * 4);
            acc0[j4 * 4 + 0] = fmaf(i0[ci], wv.x, acc0[j4 * 4 + 0]);
            acc0[j4 * 4 + 1] = fmaf(i0[ci], wv.y, acc0[j4 * 4 + 1]);
            acc0[j4 * 4 + 2] = fmaf(i0[ci], wv.z, acc0[j4 * 4 + 2]);
            acc0[j4 * 4 + 3] = fmaf(i0[ci], wv.w, acc0[j4 * 4 + 3]);
            acc1[j4 * 4 + 0] = fmaf(i1[ci], wv.x, acc1[j4 * 4 + 0]);
            acc1[j4 * 4 + 1] = fmaf(i1[ci], wv.y, acc1[j4 * 4 + 1]);
            acc1[j4 * 4 + 2] = fmaf(i1[ci], wv.z, acc1[j4 * 4 + 2]);
            acc1[j4 * 4 + 3] = fmaf(i1[ci], wv.w, acc1[j4 * 4 + 3]);
          }
        }
        // 4th input channel (only present when in_channels == 4)
        if (Cin > 3) {
          const float a0 = in[s].w, a1 = in[s + 1].w;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 wv = *reinterpret_cast<const float4*>(wt + 3 * Cout + j4 * 4);
            acc0[j4 * 4 + 0] = fmaf(a0, wv.x, acc0[j4 * 4 + 0]);
            acc0[j4 * 4 + 1] = fmaf(a0, wv.y, acc0[j4 * 4 + 1]);
            acc0[j4 * 4 + 2] = fmaf(a0, wv.z, acc0[j4 * 4 + 2]);
            acc0[j4 * 4 + 3] = fmaf(a0, wv.w, acc0[j4 * 4 + 3]);
            acc1[j4 * 4 + 0] = fmaf(a1, wv.x, acc1[j4 * 4 + 0]);
            acc1[j4 * 4 + 1] = fmaf(a1, wv.y, acc1[j4 * 4 + 1]);
            acc1[j4 * 4 + 2] = fmaf(a1, wv.z, acc1[j4 * 4 + 2]);
            acc1[j4 * 4 + 3] = fmaf(a1, wv.w, acc1[j4 * 4 + 3]);
          }
        }
      }
    }
    const int hh = h0 + py, ww = w0 + px;
    if (hh < H) {
#pragma unroll
      for (int pix = 0; pix < 2; ++pix) {
        if (ww + pix < W) {
          const float* acc = pix == 0 ? acc0 : acc1;
          uint4* dst = reinterpret_cast<uint4*>(y + ((static_cast<size_t>(b) * H + hh) * W + ww + pix) * (SPLIT ? 2 * Cout : Cout) +
                                                cg * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = relu ? fmaxf(acc[j * 8 + k], 0.f) : acc[j * 8 + k];
            dst[j] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                pack_bf16x2(v[6], v[7]));
            if (SPLIT) {
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] -= __bfloat162float(__float2bfloat16_rn(v[k]));
              dst[Cout / 8 + j] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                             pack_bf16x2(v[6], v[7]));
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Head (north_star (c)): Conv2d(f0, 1, 1)+bias (README.md:1447,1481) -> logits; optional sigmoid ->
// probabilities; optional mask = (sigmoid(z) > thr) * 255 uint8 (src/unet.py:63-67, strict '>').
// 8 lanes share one pixel (16 B = 8 channels each per step) so every warp load is contiguous.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, float bias, size_t npix, int C,
            float* __restrict__ logits, float* __restrict__ probs, uint8_t* __restrict__ mask, float thr) {
  pdl_enter();
  const int sub = threadIdx.x & 7;
  const size_t gstride = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 3;
  const size_t g0 = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 3;
  const size_t iters = (npix + gstride - 1) / gstride;  // uniform trip count: every lane joins the shuffles
  for (size_t it = 0; it < iters; ++it) {
    const size_t p = g0 + it * gstride;
    const bool live = p < npix;
    float acc = 0.f;
    if (live) {
      const uint4* row = reinterpret_cast<const uint4*>(x + p * C);
      for (int c8 = sub; c8 < C / 8; c8 += 8) {
        const uint4 r = __ldg(row + c8);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c8 * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + c8 * 8 + 4));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
        acc = fmaf(__low2float(h[0]), w0.x, acc);
        acc = fmaf(__high2float(h[0]), w0.y, acc);
        acc = fmaf(__low2float(h[1]), w0.z, acc);
        acc = fmaf(__high2float(h[1]), w0.w, acc);
        acc = fmaf(__low2float(h[2]), w1.x, acc);
        acc = fmaf(__high2float(h[2]), w1.y, acc);
        acc = fmaf(__low2float(h[3]), w1.z, acc);
        acc = fmaf(__high2float(h[3]), w1.w, acc);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (live && sub == 0) {
      const float z = acc + bias;
      if (logits != nullptr) logits[p] = z;
      if (probs != nullptr || mask != nullptr) {
        const float s = 1.f / (1.f + expf(-z));
        if (probs != nullptr) probs[p] = s;
        if (mask != nullptr) mask[p] = (s > thr) ? 255 : 0;
      }
    }
  }
}

// Head for out_channels > 1 (README.md:1447 builds nn.Conv2d(features[0], out_channels, 1) for any out_channels):
// x bf16 [B*hw][C], w fp32 [OC][C] (staged in shared memory), bias fp32 [OC]; outputs NCHW [B][OC][hw].
// One thread per pixel, eight output channels per sweep over the pixel's row.
__global__ void __launch_bounds__(256)
head_multi_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int B,
                  size_t hw, int C, int OC, float* __restrict__ logits, float* __restrict__ probs, uint8_t* __restrict__ mask,
                  float thr) {
  pdl_enter();
  extern __shared__ float hsm[];   // [OC][C] weights, [OC] bias
  for (int i = threadIdx.x; i < OC * C; i += blockDim.x) hsm[i] = w[i];
  for (int i = threadIdx.x; i < OC; i += blockDim.x) hsm[OC * C + i] = bias[i];
  __syncthreads();
  const size_t npix = static_cast<size_t>(B) * hw;
  for (size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; p < npix;
       p += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4* row = reinterpret_cast<const uint4*>(x + p * C);
    const size_t bi = p / hw, q = p - bi * hw;
    for (int oc0 = 0; oc0 < OC; oc0 += 8) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int c8 = 0; c8 < C / 8; ++c8) {
        const uint4 r = __ldg(row + c8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          f[2 * k] = __low2float(h[k]);
          f[2 * k + 1] = __high2float(h[k]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (oc0 + j < OC) {
            const float* wr = hsm + (oc0 + j) * C + c8 * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[j] = fmaf(f[k], wr[k], acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (oc0 + j < OC) {
          const float z = acc[j] + hsm[OC * C + oc0 + j];
          const size_t o = (bi * OC + oc0 + j) * hw + q;
          if (logits != nullptr) logits[o] = z;
          if (probs != nullptr || mask != nullptr) {
            const float sg = 1.f / (1.f + expf(-z));
            if (probs != nullptr) probs[o] = sg;
            if (mask != nullptr) mask[o] = (sg > thr) ? 255 : 0;
          }
        }
      }
    }
  }
}

// Stand-alone 2x2/2 max-pool on NHWC bf16 (used when the pool is not fused into a conv epilogue).
__global__ void maxpool2x2_kernel(const uint4* __restrict__ x, int B, int H, int W, int C8, uint4* __restrict__ y) {
  pdl_enter();
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = i % C8;
    size_t r = i / C8;
    const int wo = r % Wo;
    r /= Wo;
    const int ho = r % Ho;
    const size_t b = r / Ho;
    const size_t base = ((b * H + 2 * ho) * W + 2 * wo) * C8 + c;
    const uint4 a0 = __ldg(x + base), a1 = __ldg(x + base + C8);
    const uint4 a2 = __ldg(x + base + static_cast<size_t>(W) * C8), a3 = __ldg(x + base + static_cast<size_t>(W) * C8 + C8);
    uint4 o;
    o.x = bf16x2_max(bf16x2_max(a0.x, a1.x), bf16x2_max(a2.x, a3.x));
    o.y = bf16x2_max(bf16x2_max(a0.y, a1.y), bf16x2_max(a2.y, a3.y));
    o.z = bf16x2_max(bf16x2_max(a0.z, a1.z), bf16x2_max(a2.z, a3.z));
    o.w = bf16x2_max(bf16x2_max(a0.w, a1.w), bf16x2_max(a2.w, a3.w));
    y[i] = o;
  }
}

}  // namespace ub
