// Bandwidth-bound kernels of the training step (reference: README.md:2060-2084 train_one_epoch,
// 1855-1893 BCEDiceLoss, 2173-2174 AdamW). All tensors NHWC bf16 unless noted; per-channel statistics fp32.
// BatchNorm runs in training mode: batch statistics, biased variance for normalisation, unbiased for the
// running buffer, momentum 0.1 (PyTorch defaults used by nn.BatchNorm2d in README.md:1453,1456).
#pragma once
#include "ptx.cuh"

namespace ub {

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __low2float(h[i]);
    f[2 * i + 1] = __high2float(h[i]);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// Thread mapping shared by the per-channel reductions: C8 = C/8 threads cover one pixel (8 channels each), a block of
// 256 threads covers 256/C8 pixels per step; requires C8 to divide 256 (C in {64,128,256,512,1024,2048}).
// Per-block partial sums are combined through shared memory, then one atomicAdd per channel per block.

// sum[c] += sum_p y[p][c], sumsq[c] += sum_p y[p][c]^2
__global__ void __launch_bounds__(256)
chan_stats_kernel(const uint4* __restrict__ y, size_t npix, int C8, double* __restrict__ sum, double* __restrict__ sumsq) {
  pdl_enter();
  extern __shared__ float red[];  // [2][256][8]
  const int cl = threadIdx.x % C8;
  const int pl = threadIdx.x / C8;
  const int ppb = 256 / C8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ss[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t p = static_cast<size_t>(blockIdx.x) * ppb + pl; p < npix; p += static_cast<size_t>(gridDim.x) * ppb) {
    float f[8];
    unpack8(__ldg(y + p * C8 + cl), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i] += f[i];
      ss[i] = fmaf(f[i], f[i], ss[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x * 8 + i] = s[i];
    red[2048 + threadIdx.x * 8 + i] = ss[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C8 * 8; c += 256) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < ppb; ++j) {
      a += red[(j * C8 + c / 8) * 8 + (c & 7)];
      b += red[2048 + (j * C8 + c / 8) * 8 + (c & 7)];
    }
    atomicAdd(sum + c, static_cast<double>(a));
    atomicAdd(sumsq + c, static_cast<double>(b));
  }
}

// Per channel: batch mean / biased var -> scale = gamma*invstd, shift = beta - mean*scale; running-statistics update
// (momentum, unbiased variance). A separate tiny launch on purpose: folding it into the apply kernels (every thread deriving
// its eight channels' constants from the double sums) cost more than the launch - two fp64 divides per channel per thread made
// the 18 apply launches of a step 0.47 ms slower (measured, round 1).
struct BnFin {
  const double* sum;
  const double* sumsq;
  float count, eps, momentum;
  const float* gamma;
  const float* beta;
  float *mean, *invstd, *scale, *shift;
  float *running_mean, *running_var;   // may be null
  int C;     // physical channels of the tensor
  int lC;    // logical channels (<= C): gamma / beta / running statistics have lC entries; the zero-extended channels beyond
             // them get scale = shift = 0, so they stay exactly zero through BatchNorm + ReLU and its backward
};
__global__ void bn_finalize_kernel(const BnFin f) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= f.C) return;
  if (c >= f.lC) {
    f.mean[c] = 0.f;
    f.invstd[c] = 0.f;
    f.scale[c] = 0.f;
    f.shift[c] = 0.f;
    return;
  }
  const double md = f.sum[c] / static_cast<double>(f.count);
  const double vd = f.sumsq[c] / static_cast<double>(f.count) - md * md;
  const float m = static_cast<float>(md);
  const float var = fmaxf(static_cast<float>(vd), 0.f);
  const float is = rsqrtf(var + f.eps);
  f.mean[c] = m;
  f.invstd[c] = is;
  const float sc = f.gamma[c] * is;
  f.scale[c] = sc;
  f.shift[c] = f.beta[c] - m * sc;
  if (f.running_mean != nullptr) {
    f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * m;
    const float unbiased = f.count > 1.f ? var * f.count / (f.count - 1.f) : var;
    f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * unbiased;
  }
}

// The same finalisation inside the kernel that applies it (bn_relu_apply*_kernel with fin.sum != null): every thread works
// out scale / shift of its own eight channels from the batch sums - multiplications by 1/count in double, no division, a few
// dozen instructions per thread - and the first C/8 threads of the grid also publish mean / invstd / scale / shift (the
// backward reads them) and update the running statistics. Saves one launch per BatchNorm layer and step.
__device__ __forceinline__ void bn_fin_channels(const BnFin& f, int c0, bool publish, float (&sc)[8], float (&sh)[8]) {
  const double inv_n = 1.0 / static_cast<double>(f.count);   // one division per thread, of loop-invariant values
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + k;
    float m = 0.f, is = 0.f, var = 0.f;
    sc[k] = 0.f;
    sh[k] = 0.f;
    if (c < f.lC) {
      const double md = f.sum[c] * inv_n;
      const double vd = f.sumsq[c] * inv_n - md * md;
      m = static_cast<float>(md);
      var = fmaxf(static_cast<float>(vd), 0.f);
      is = rsqrtf(var + f.eps);
      sc[k] = __ldg(f.gamma + c) * is;
      sh[k] = __ldg(f.beta + c) - m * sc[k];
    }
    if (publish) {
      f.mean[c] = m;
      f.invstd[c] = is;
      f.scale[c] = sc[k];
      f.shift[c] = sh[k];
      if (c < f.lC && f.running_mean != nullptr) {
        f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * m;
        const float unbiased = f.count > 1.f ? var * f.count / (f.count - 1.f) : var;
        f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * unbiased;
      }
    }
  }
}

// a = relu(y*scale + shift). The grid stride is a multiple of C8 (C8 | 256), so a thread's channel group never changes
// and its eight scale/shift pairs live in registers. fin.sum != null: scale / shift come from the batch sums (above).
__global__ void __launch_bounds__(256)
bn_relu_apply_kernel(const uint4* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                     size_t n8, int C8, uint4* __restrict__ a, const BnFin fin) {
  pdl_enter();
  const size_t i0 = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const int c = static_cast<int>(i0 % C8) * 8;
  float sc[8], sh[8];
  if (fin.sum != nullptr) {
    bn_fin_channels(fin, c, i0 < static_cast<size_t>(C8), sc, sh);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = __ldg(scale + c + k);
      sh[k] = __ldg(shift + c + k);
    }
  }
  // four 16-byte loads in flight per thread (one per trip left the pass latency-bound: 5.0 TB/s at 224^2)
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = i0;
  for (; i + 3 * stride < n8; i += 4 * stride) {
    uint4 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) r[u] = __ldg(y + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(r[u], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaxf(fmaf(f[k], sc[k], sh[k]), 0.f);
      a[i + u * stride] = pack8(f);
    }
  }
  for (; i < n8; i += stride) {
    float f[8];
    unpack8(__ldg(y + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaxf(fmaf(f[k], sc[k], sh[k]), 0.f);
    a[i] = pack8(f);
  }
}

// a = relu(y*scale + shift) and p = maxpool2x2(a) in one pass (encoder conv1: README.md:1464-1467 in train mode).
// One thread = one 2x2 window x 8 channels.
__global__ void __launch_bounds__(256)
bn_relu_apply_pool_kernel(const uint4* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift, int B,
                          int H, int W, int C8, uint4* __restrict__ a, uint4* __restrict__ pl, const BnFin fin) {
  pdl_enter();
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * C8;
  const size_t i0 = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const int c8 = static_cast<int>(i0 % C8);  // loop-invariant: the grid stride is a multiple of C8
  float sc[8], sh[8];
  if (fin.sum != nullptr) {
    bn_fin_channels(fin, c8 * 8, i0 < static_cast<size_t>(C8), sc, sh);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = __ldg(scale + c8 * 8 + k);
      sh[k] = __ldg(shift + c8 * 8 + k);
    }
  }
  for (size_t i = i0; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t r = i / C8;
    const int wo = r % Wo;
    r /= Wo;
    const int ho = r % Ho;
    const size_t b = r / Ho;
    const size_t base = ((b * H + 2 * ho) * W + 2 * wo) * C8 + c8;
    const size_t off[4] = {0, (size_t)C8, (size_t)W * C8, (size_t)W * C8 + C8};
    float mx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) mx[k] = 0.f;  // post-ReLU values are >= 0
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[8];
      unpack8(__ldg(y + base + off[q]), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaxf(fmaf(f[k], sc[k], sh[k]), 0.f);
      const uint4 pk = pack8(f);
      a[base + off[q]] = pk;
      float fr[8];
      unpack8(pk, fr);  // pool the bf16-rounded values so p == maxpool(a) bit-exactly
#pragma unroll
      for (int k = 0; k < 8; ++k) mx[k] = fmaxf(mx[k], fr[k]);
    }
    pl[i] = pack8(mx);
  }
}

// Backward of (BN train + ReLU), pass 1: g = dA * [y*scale+shift > 0];  s1[c] += sum g, s2[c] += sum g * xhat.
__global__ void __launch_bounds__(256)
bn_relu_bwd_reduce_kernel(const uint4* __restrict__ dA, const uint4* __restrict__ y, const float* __restrict__ scale,
                          const float* __restrict__ shift, const float* __restrict__ mean,
                          const float* __restrict__ invstd, size_t npix, int C8, float* __restrict__ s1,
                          float* __restrict__ s2) {
  pdl_enter();
  extern __shared__ float red[];
  const int cl = threadIdx.x % C8;
  const int pl = threadIdx.x / C8;
  const int ppb = 256 / C8;
  float sc[8], sh[8], mu[8], is[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = scale[cl * 8 + i];
    sh[i] = shift[cl * 8 + i];
    mu[i] = mean[cl * 8 + i];
    is[i] = invstd[cl * 8 + i];
  }
  float a1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // four pixels per trip: eight independent 16-byte loads in flight per thread (the pass is HBM-latency bound otherwise)
  const size_t stride = static_cast<size_t>(gridDim.x) * ppb;
  size_t p = static_cast<size_t>(blockIdx.x) * ppb + pl;
  for (; p + 3 * stride < npix; p += 4 * stride) {
    uint4 ry[4], rd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ry[u] = __ldg(y + (p + u * stride) * C8 + cl);
      rd[u] = __ldg(dA + (p + u * stride) * C8 + cl);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float fy[8], fd[8];
      unpack8(ry[u], fy);
      unpack8(rd[u], fd);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float g = fmaf(fy[i], sc[i], sh[i]) > 0.f ? fd[i] : 0.f;
        a1[i] += g;
        a2[i] = fmaf(g, (fy[i] - mu[i]) * is[i], a2[i]);
      }
    }
  }
  for (; p < npix; p += stride) {
    float fy[8], fd[8];
    unpack8(__ldg(y + p * C8 + cl), fy);
    unpack8(__ldg(dA + p * C8 + cl), fd);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float g = fmaf(fy[i], sc[i], sh[i]) > 0.f ? fd[i] : 0.f;
      a1[i] += g;
      a2[i] = fmaf(g, (fy[i] - mu[i]) * is[i], a2[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x * 8 + i] = a1[i];
    red[2048 + threadIdx.x * 8 + i] = a2[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C8 * 8; c += 256) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < ppb; ++j) {
      a += red[(j * C8 + c / 8) * 8 + (c & 7)];
      b += red[2048 + (j * C8 + c / 8) * 8 + (c & 7)];
    }
    atomicAdd(s1 + c, a);
    atomicAdd(s2 + c, b);
  }
}

// pass 2: dY = gamma*invstd * (g - s1/N - xhat * s2/N) = sc*g + k1*y + k0 with per-channel
//   k1 = -sc*invstd*s2/N,  k0 = -sc*s1/N - k1*mean   (held in registers: the thread's channel group is loop-invariant)
__global__ void __launch_bounds__(256)
bn_relu_bwd_apply_kernel(const uint4* __restrict__ dA, const uint4* __restrict__ y, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ s1, const float* __restrict__ s2,
                         float inv_count, size_t n8, int C8, uint4* __restrict__ dY, GradRoute route, long long off_gamma,
                         long long off_beta, int lC) {
  pdl_enter();
  // block 0 also publishes d gamma = sum g*xhat (s2) and d beta = sum g (s1) into the (possibly remote) flat gradient
  // (logical channels only: gamma / beta have lC entries)
  if (blockIdx.x == 0 && route.local != nullptr) {
    for (int ch = threadIdx.x; ch < lC; ch += blockDim.x) {
      grad_add(route, off_gamma + ch, s2[ch]);
      grad_add(route, off_beta + ch, s1[ch]);
    }
  }
  const size_t i0 = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const int c = static_cast<int>(i0 % C8) * 8;
  float sc[8], sh[8], k1[8], k0[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = __ldg(scale + c + k);
    sh[k] = __ldg(shift + c + k);
    k1[k] = -sc[k] * __ldg(invstd + c + k) * __ldg(s2 + c + k) * inv_count;
    k0[k] = -sc[k] * __ldg(s1 + c + k) * inv_count - k1[k] * __ldg(mean + c + k);
  }
  for (size_t i = i0; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float fy[8], fd[8], o[8];
    unpack8(__ldg(y + i), fy);
    unpack8(__ldg(dA + i), fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float g = fmaf(fy[k], sc[k], sh[k]) > 0.f ? fd[k] : 0.f;
      o[k] = fmaf(sc[k], g, fmaf(k1[k], fy[k], k0[k]));
    }
    dY[i] = pack8(o);
  }
}

// BatchNorm-backward statistics fused into the pass that PRODUCES the incoming gradient dA of a conv layer (max-pool backward,
// head backward): that pass has dA in registers, so reading the layer's raw conv output y beside it yields
// s1 = sum dA*[relu on] and s2 = sum dA*[relu on]*xhat without bn_relu_bwd_reduce_kernel's extra read of dA (and its launch).
struct BnBwdStats {
  const uint4* y;       // raw conv output of the layer whose dA is being produced (null: no fusion)
  const float* scale;   // per channel: gamma * invstd
  const float* shift;   // beta - mean * scale
  const float* mean;
  const float* invstd;
  float* s1;            // [C] += sum g        (g = dA where relu(bn(y)) > 0)
  float* s2;            // [C] += sum g * xhat
};
struct BnBwdRegs {       // a thread's eight channels
  float sc[8], sh[8], mu[8], is[8], a1[8], a2[8];
};
__device__ __forceinline__ void bnbwd_init(const BnBwdStats& b, int c8, BnBwdRegs& r) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    r.sc[i] = __ldg(b.scale + c8 * 8 + i);
    r.sh[i] = __ldg(b.shift + c8 * 8 + i);
    r.mu[i] = __ldg(b.mean + c8 * 8 + i);
    r.is[i] = __ldg(b.invstd + c8 * 8 + i);
    r.a1[i] = 0.f;
    r.a2[i] = 0.f;
  }
}
// dA8: the eight gradient values AS STORED (bf16-rounded), so the sums match what the apply pass will read back
__device__ __forceinline__ void bnbwd_accumulate(BnBwdRegs& r, const uint4& y8, const uint4& dA8) {
  float fy[8], fd[8];
  unpack8(y8, fy);
  unpack8(dA8, fd);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float g = fmaf(fy[i], r.sc[i], r.sh[i]) > 0.f ? fd[i] : 0.f;
    r.a1[i] += g;
    r.a2[i] = fmaf(g, (fy[i] - r.mu[i]) * r.is[i], r.a2[i]);
  }
}
// block reduction (thread t holds channel group t % C8) + one atomicAdd per channel; red: 2 * 2048 floats of shared memory
__device__ __forceinline__ void bnbwd_flush(const BnBwdStats& b, const BnBwdRegs& r, int C8, float* red) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x * 8 + i] = r.a1[i];
    red[2048 + threadIdx.x * 8 + i] = r.a2[i];
  }
  __syncthreads();
  const int ppb = 256 / C8;
  for (int c = threadIdx.x; c < C8 * 8; c += 256) {
    float x = 0.f, z = 0.f;
    for (int j = 0; j < ppb; ++j) {
      x += red[(j * C8 + c / 8) * 8 + (c & 7)];
      z += red[2048 + (j * C8 + c / 8) * 8 + (c & 7)];
    }
    atomicAdd(b.s1 + c, x);
    atomicAdd(b.s2 + c, z);
  }
}

// Max-pool backward merged with the skip connection's other gradient:
// dA[b,h,w,c] = (d_skip ? d_skip[b,h,w,c] : 0) + (pixel is the FIRST maximum of its 2x2 window ? dP[b,h/2,w/2,c] : 0)
// BN = false keeps the BatchNorm registers (48 floats) out of the kernel: the first version carried them always and ran at one
// 256-thread CTA per SM (145 registers, 3.7 TB/s). The window's values stay PACKED (two bf16 per register) until they are used.
template <bool BN>
__global__ void __launch_bounds__(256)
maxpool_bwd_add_kernel(const uint4* __restrict__ a, const uint4* __restrict__ dP, const uint4* __restrict__ d_skip,
                       int skip_pitch8 /* uint4 per pixel of d_skip (>= C8: it may be the first half of a concat gradient) */,
                       int B, int H, int W, int C8, uint4* __restrict__ dA, const BnBwdStats bn) {
  pdl_enter();
  extern __shared__ float red[];   // 2 * 2048 floats when BN (C8 must then divide 256)
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * C8;
  BnBwdRegs br;
  if constexpr (BN) bnbwd_init(bn, threadIdx.x % C8, br);   // the grid stride is a multiple of C8: the channel group is fixed
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = i % C8;
    size_t r = i / C8;
    const int wo = r % Wo;
    r /= Wo;
    const int ho = r % Ho;
    const size_t b = r / Ho;
    const size_t pix0 = (b * H + 2 * ho) * W + 2 * wo;
    const size_t base = pix0 * C8 + c;
    const size_t off[4] = {0, (size_t)C8, (size_t)W * C8, (size_t)W * C8 + C8};
    const size_t poff[4] = {0, 1, (size_t)W, (size_t)W + 1};
    uint4 va[4], vs[4], out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) va[k] = __ldg(a + base + off[k]);
    const uint4 vg = __ldg(dP + i);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      vs[k] = d_skip != nullptr ? __ldg(d_skip + (pix0 + poff[k]) * skip_pitch8 + c) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {          // 32-bit word j = channels 2j, 2j+1
      float2 fa[4], fs[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        fa[k] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&(&va[k].x)[j]));
        fs[k] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&(&vs[k].x)[j]));
      }
      const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&(&vg.x)[j]));
      int bx = 0, by = 0;
      float mx = fa[0].x, my = fa[0].y;
#pragma unroll
      for (int k = 1; k < 4; ++k) {       // strict '>' keeps the first maximum in (row, col) scan order, as ATen does
        if (fa[k].x > mx) {
          mx = fa[k].x;
          bx = k;
        }
        if (fa[k].y > my) {
          my = fa[k].y;
          by = k;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        (&out[k].x)[j] = pack_bf16x2(fs[k].x + (k == bx ? g.x : 0.f), fs[k].y + (k == by ? g.y : 0.f));
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      dA[base + off[k]] = out[k];
      if constexpr (BN) bnbwd_accumulate(br, __ldg(bn.y + base + off[k]), out[k]);
    }
  }
  if constexpr (BN) bnbwd_flush(bn, br, C8, red);
}

// Head forward in training (README.md:1481): logits[p] = bias + sum_c a[p][c] * w[c]; 8 lanes share one pixel.
__global__ void __launch_bounds__(256)
head_fwd_train_kernel(const uint4* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias, size_t npix,
                      int C8, float* __restrict__ logits) {
  pdl_enter();
  const int sub = threadIdx.x & 7;
  const size_t stride = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 3;
  // the loop bound is warp-uniform (p0 is the warp's first pixel) so the shuffles always see all 32 lanes
  size_t p0 = ((blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5) << 2;
  if (C8 == 8) {
    // the usual case (64 channels: one 16-byte chunk per lane): weights in registers, four pixel groups per trip so four
    // loads are in flight per thread (one per trip ran at 4.4 TB/s)
    float wv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) wv[k] = __ldg(w + sub * 8 + k);
    const float bv = __ldg(bias);
    for (; p0 + 3 * stride < npix; p0 += 4 * stride) {
      uint4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t p = p0 + u * stride + ((threadIdx.x & 31) >> 3);
        r[u] = p < npix ? __ldg(a + p * 8 + sub) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t p = p0 + u * stride + ((threadIdx.x & 31) >> 3);
        float f[8], acc = 0.f;
        unpack8(r[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(f[k], wv[k], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (sub == 0 && p < npix) logits[p] = acc + bv;
      }
    }
  }
  for (; p0 < npix; p0 += stride) {
    const size_t p = p0 + ((threadIdx.x & 31) >> 3);
    float acc = 0.f;
    if (p < npix) {
      for (int c = sub; c < C8; c += 8) {
        float f[8];
        unpack8(__ldg(a + p * C8 + c), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(f[k], __ldg(w + c * 8 + k), acc);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (sub == 0 && p < npix) logits[p] = acc + __ldg(bias);
  }
}

// Head backward: dA[p][c] = dz[p] * w[c] (bf16);  dw[c] += sum_p dz[p] * a[p][c];  db += sum_p dz[p]
// Four pixels per trip (four independent 16-byte loads in flight per thread; one left the pass latency-bound at 3.3 TB/s);
// BN = false keeps the BatchNorm registers out of the kernel.
template <bool BN>
__global__ void __launch_bounds__(256)
head_bwd_kernel(const uint4* __restrict__ a, const float* __restrict__ dz, const float* __restrict__ w, size_t npix, int C8,
                uint4* __restrict__ dA, GradRoute route, long long off_w, long long off_b, const BnBwdStats bn, int lC) {
  pdl_enter();
  extern __shared__ float red[];   // (2048 + 256) floats; 2 * 2048 when BN
  const int cl = threadIdx.x % C8;
  const int pl = threadIdx.x / C8;
  const int ppb = 256 / C8;
  float wv[8], acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 8; ++i) wv[i] = w[cl * 8 + i];
  float bsum = 0.f;
  BnBwdRegs br;
  if constexpr (BN) bnbwd_init(bn, cl, br);
  auto one = [&](size_t p, const uint4& ra, float g) {
    float fa[8], o[8];
    unpack8(ra, fa);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] = fmaf(g, fa[i], acc[i]);
      o[i] = g * wv[i];
    }
    const uint4 pk = pack8(o);
    dA[p * C8 + cl] = pk;
    if constexpr (BN) bnbwd_accumulate(br, __ldg(bn.y + p * C8 + cl), pk);
    if (cl == 0) bsum += g;
  };
  const size_t stride = static_cast<size_t>(gridDim.x) * ppb;
  size_t p = static_cast<size_t>(blockIdx.x) * ppb + pl;
  for (; p + 3 * stride < npix; p += 4 * stride) {
    uint4 ra[4];
    float g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ra[u] = __ldg(a + (p + u * stride) * C8 + cl);
      g[u] = __ldg(dz + p + u * stride);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) one(p + u * stride, ra[u], g[u]);
  }
  for (; p < npix; p += stride) one(p, __ldg(a + p * C8 + cl), __ldg(dz + p));
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = acc[i];
  red[2048 + threadIdx.x] = bsum;
  __syncthreads();
  for (int c = threadIdx.x; c < C8 * 8; c += 256) {
    float s = 0.f;
    for (int j = 0; j < ppb; ++j) s += red[(j * C8 + c / 8) * 8 + (c & 7)];
    if (c < lC) grad_add(route, off_w + c, s);      // output.weight has lC (logical) entries
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int j = 0; j < 256; ++j) s += red[2048 + j];
    grad_add(route, off_b, s);
  }
  if constexpr (BN) {
    __syncthreads();   // red is reused
    bnbwd_flush(bn, br, C8, red);
  }
}

// out_channels > 1 (README.md:1447 builds any; the reference itself trains out_channels = 1): plain versions of the two head
// kernels. logits / dz are NCHW [B][OC][hw], w is the zero-extended weight [OC][C8*8], OC <= HEAD_MAX_OC.
constexpr int HEAD_MAX_OC = 8;
__global__ void __launch_bounds__(256)
head_fwd_train_multi_kernel(const uint4* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias, size_t npix,
                            size_t hw, int C8, int OC, float* __restrict__ logits) {
  pdl_enter();
  for (size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; p < npix;
       p += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float acc[HEAD_MAX_OC];
#pragma unroll
    for (int oc = 0; oc < HEAD_MAX_OC; ++oc) acc[oc] = 0.f;
    for (int c = 0; c < C8; ++c) {
      float f[8];
      unpack8(__ldg(a + p * C8 + c), f);
#pragma unroll
      for (int oc = 0; oc < HEAD_MAX_OC; ++oc) {
        if (oc < OC) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[oc] = fmaf(f[k], __ldg(w + (static_cast<size_t>(oc) * C8 + c) * 8 + k), acc[oc]);
        }
      }
    }
    const size_t b = p / hw, q = p - b * hw;
#pragma unroll
    for (int oc = 0; oc < HEAD_MAX_OC; ++oc) {
      if (oc < OC) logits[(b * OC + oc) * hw + q] = acc[oc] + __ldg(bias + oc);
    }
  }
}

// dA[p][c] = sum_oc dz[b][oc][q] * w[oc][c];  dw[oc][c] += sum_p dz * a[p][c] (c < lC);  db[oc] += sum_p dz
__global__ void __launch_bounds__(256)
head_bwd_multi_kernel(const uint4* __restrict__ a, const float* __restrict__ dz, const float* __restrict__ w, size_t npix, size_t hw,
                      int C8, int OC, uint4* __restrict__ dA, GradRoute route, long long off_w, long long off_b, int lC) {
  pdl_enter();
  extern __shared__ float red[];   // [256][8] + [256]
  const int cl = threadIdx.x % C8;
  const int pl = threadIdx.x / C8;
  const int ppb = 256 / C8;
  for (int oc = 0; oc < OC; ++oc) {          // one sweep per output channel for the parameter gradients (OC is small)
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float bsum = 0.f;
    for (size_t p = static_cast<size_t>(blockIdx.x) * ppb + pl; p < npix; p += static_cast<size_t>(gridDim.x) * ppb) {
      const size_t b = p / hw, q = p - b * hw;
      const float g = __ldg(dz + (b * OC + oc) * hw + q);
      float fa[8];
      unpack8(__ldg(a + p * C8 + cl), fa);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(g, fa[i], acc[i]);
      if (cl == 0) bsum += g;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = acc[i];
    red[2048 + threadIdx.x] = bsum;
    __syncthreads();
    for (int c = threadIdx.x; c < C8 * 8; c += 256) {
      float s_ = 0.f;
      for (int j = 0; j < ppb; ++j) s_ += red[(j * C8 + c / 8) * 8 + (c & 7)];
      if (c < lC) grad_add(route, off_w + static_cast<long long>(oc) * lC + c, s_);
    }
    if (threadIdx.x == 0) {
      float s_ = 0.f;
      for (int j = 0; j < 256; ++j) s_ += red[2048 + j];
      grad_add(route, off_b + oc, s_);
    }
    __syncthreads();
  }
  // activation gradient
  for (size_t p = static_cast<size_t>(blockIdx.x) * ppb + pl; p < npix; p += static_cast<size_t>(gridDim.x) * ppb) {
    const size_t b = p / hw, q = p - b * hw;
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int oc = 0; oc < OC; ++oc) {
      const float g = __ldg(dz + (b * OC + oc) * hw + q);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(g, __ldg(w + (static_cast<size_t>(oc) * C8 + cl) * 8 + i), o[i]);
    }
    dA[p * C8 + cl] = pack8(o);
  }
}

// BCEWithLogits(pos_weight) + Dice (README.md:1868-1893), pass 1: sums[0..3] += {sum bce_i, sum sigma*t, sum sigma, sum t}
__global__ void __launch_bounds__(256)
bce_dice_reduce_kernel(const float* __restrict__ z, const float* __restrict__ t, size_t n, float pos_weight,
                       double* __restrict__ sums) {
  pdl_enter();
  __shared__ double red[4][256];
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float x = z[i], y = t[i];
    const float lw = 1.f + (pos_weight - 1.f) * y;
    const float bce = (1.f - y) * x + lw * (log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f));
    const float sg = 1.f / (1.f + expf(-x));
    s0 += bce;
    s1 += sg * y;
    s2 += sg;
    s3 += y;
  }
  red[0][threadIdx.x] = s0;
  red[1][threadIdx.x] = s1;
  red[2][threadIdx.x] = s2;
  red[3][threadIdx.x] = s3;
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0;
    for (int j = 0; j < 256; ++j) s += red[threadIdx.x][j];
    atomicAdd(sums + threadIdx.x, s);
  }
}

// Validation metrics (README.md:2086-2120 validate + compute_dice) in one pass over the logits:
// sums[0..5] += {sum bce_i, sum sigma*t, sum sigma, sum t, sum [sigma > thr]*t, sum [sigma > thr]}
__global__ void __launch_bounds__(256)
val_metrics_reduce_kernel(const float* __restrict__ z, const float* __restrict__ t, size_t n, float pos_weight, float thr,
                          double* __restrict__ sums) {
  pdl_enter();
  __shared__ double red[6][256];
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float x = z[i], y = t[i];
    const float lw = 1.f + (pos_weight - 1.f) * y;
    s[0] += (1.f - y) * x + lw * (log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f));
    const float sg = 1.f / (1.f + expf(-x));
    s[1] += sg * y;
    s[2] += sg;
    s[3] += y;
    const float pm = sg > thr ? 1.f : 0.f;  // strict '>' as torch.sigmoid(outputs) > 0.5 (README.md:2103)
    s[4] += pm * y;
    s[5] += pm;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) red[k][threadIdx.x] = s[k];
  __syncthreads();
  if (threadIdx.x < 6) {
    double a = 0;
    for (int j = 0; j < 256; ++j) a += red[threadIdx.x][j];
    atomicAdd(sums + threadIdx.x, a);
  }
}
// out[0..3] = {total loss, bce, dice loss, dice score of the thresholded prediction}
__global__ void val_metrics_finalize_kernel(const double* __restrict__ sums, size_t n, float bce_w, float dice_w, float smooth,
                                            float* __restrict__ out) {
  pdl_enter();
  const double bce = sums[0] / static_cast<double>(n);
  const double dice = 1.0 - (2.0 * sums[1] + smooth) / (sums[2] + sums[3] + smooth);
  out[0] = static_cast<float>(bce_w * bce + dice_w * dice);
  out[1] = static_cast<float>(bce);
  out[2] = static_cast<float>(dice);
  out[3] = static_cast<float>((2.0 * sums[4] + smooth) / (sums[5] + sums[3] + smooth));
}

// pass 2: losses[0..2] = {total, bce, dice};  dz = d total / d z
__global__ void __launch_bounds__(256)
bce_dice_grad_kernel(const float* __restrict__ z, const float* __restrict__ t, size_t n, float pos_weight, float bce_w,
                     float dice_w, float smooth, const double* __restrict__ sums, float* __restrict__ dz,
                     float* __restrict__ losses) {
  pdl_enter();
  const double inter = sums[1], S = sums[2], T = sums[3];
  const double num = 2.0 * inter + smooth, den = S + T + smooth;
  const float inv_n = 1.f / static_cast<float>(n);
  const float k_den = static_cast<float>(1.0 / den), k_ratio = static_cast<float>(num / (den * den));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const float bce = static_cast<float>(sums[0] / static_cast<double>(n));
    const float dice = static_cast<float>(1.0 - num / den);
    losses[0] = bce_w * bce + dice_w * dice;
    losses[1] = bce;
    losses[2] = dice;
  }
  if (dz == nullptr) return;  // loss values only
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float x = z[i], y = t[i];
    const float sg = 1.f / (1.f + expf(-x));
    const float dbce = (sg * (1.f + (pos_weight - 1.f) * y) - pos_weight * y) * inv_n;
    // dice = 1 - num/den: d/dsigma = -(2t*den - num)/den^2 ; dsigma/dz = sigma(1-sigma)
    const float ddice = -(2.f * y * k_den - k_ratio) * sg * (1.f - sg);
    dz[i] = bce_w * dbce + dice_w * ddice;
  }
}

// AdamW (torch.optim.AdamW semantics, README.md:2173-2174) on flat fp32 arrays; grad is pre-multiplied by grad_scale (1/world).
// step_dev (optional): the step count lives in device memory (CUDA-graph replays cannot change kernel arguments), and the
// bias corrections are derived from it here; otherwise bc1/bc2 come from the host.
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
             float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2, float grad_scale,
             const int* __restrict__ step_dev, const float* __restrict__ lr_dev) {
  pdl_enter();
  if (lr_dev != nullptr) lr = *lr_dev;   // learning rate in device memory: a scheduler changes it without re-capturing the graph
  if (step_dev != nullptr) {
    const float st = static_cast<float>(*step_dev);
    bc1 = 1.f - powf(beta1, st);
    bc2 = 1.f - powf(beta2, st);
  }
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi;
  }
}

// Data-parallel AdamW over NVLink (ZeRO-1 style): this rank owns flat elements [lo, hi) (lo a multiple of 4). The step is
// applied to the owned shard only - optimizer state exists only for the shard - and the new parameter value is stored into
// EVERY rank's parameter buffer through the peer-mapped pointers (the all-gather is these stores). Two ways to get the
// summed gradient of the shard:
//  * push (grad_bases == null): the backward kernels already added every replica's contribution straight into the owner's
//    buffer (GradRoute); `grads` holds the sum and is cleared here for the next step;
//  * pull (grad_bases != null): every replica accumulated locally; the sum over ranks is formed here from coalesced 16-byte
//    loads of the peers' buffers over NVLink (the reduce-scatter is these loads). Nothing is cleared (the next backward
//    zeroes its local buffer after the closing barrier).
__device__ __forceinline__ float adamw_one(float& p, float g, float& m, float& v, float lr, float wd, float beta1, float beta2,
                                           float eps, float step_size, float inv_sqrt_bc2) {
  float pi = p * (1.f - lr * wd);
  m = beta1 * m + (1.f - beta1) * g;
  v = beta2 * v + (1.f - beta2) * g * g;
  pi -= step_size * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
  p = pi;
  return pi;
}

__global__ void __launch_bounds__(256)
adamw_shard_allgather_kernel(float* const* __restrict__ param_bases, float* const* __restrict__ grad_bases, int world, int rank,
                             float* __restrict__ grads, float* __restrict__ m, float* __restrict__ v, long long lo, long long hi,
                             float lr, float beta1, float beta2, float eps, float wd, float grad_scale,
                             const int* __restrict__ step_dev, const float* __restrict__ lr_dev) {
  pdl_enter();
  if (lr_dev != nullptr) lr = *lr_dev;
  const float st = static_cast<float>(*step_dev);
  const float bc1 = 1.f - powf(beta1, st), bc2 = 1.f - powf(beta2, st);
  const float step_size = lr / bc1, inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  float* plocal = param_bases[rank];
  const long long n4 = (hi - lo) / 4;   // full float4 groups; the (< 4 element) tail of the last shard is done by block 0
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < n4;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = lo + 4 * q;
    float4 g;
    if (grad_bases != nullptr) {
      g = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < world; ++r) {
        const float4 t = *reinterpret_cast<const float4*>(grad_bases[(rank + r) % world] + i);   // own buffer first
        g.x += t.x, g.y += t.y, g.z += t.z, g.w += t.w;
      }
    } else {
      g = *reinterpret_cast<const float4*>(grads + i);
      *reinterpret_cast<float4*>(grads + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 p4 = *reinterpret_cast<const float4*>(plocal + i);
    float4 m4 = *reinterpret_cast<const float4*>(m + 4 * q);
    float4 v4 = *reinterpret_cast<const float4*>(v + 4 * q);
    adamw_one(p4.x, g.x * grad_scale, m4.x, v4.x, lr, wd, beta1, beta2, eps, step_size, inv_sqrt_bc2);
    adamw_one(p4.y, g.y * grad_scale, m4.y, v4.y, lr, wd, beta1, beta2, eps, step_size, inv_sqrt_bc2);
    adamw_one(p4.z, g.z * grad_scale, m4.z, v4.z, lr, wd, beta1, beta2, eps, step_size, inv_sqrt_bc2);
    adamw_one(p4.w, g.w * grad_scale, m4.w, v4.w, lr, wd, beta1, beta2, eps, step_size, inv_sqrt_bc2);
    *reinterpret_cast<float4*>(m + 4 * q) = m4;
    *reinterpret_cast<float4*>(v + 4 * q) = v4;
    for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(param_bases[(rank + r) % world] + i) = p4;
  }
  if (blockIdx.x == 0) {
    for (long long i = lo + 4 * n4 + threadIdx.x; i < hi; i += blockDim.x) {
      float g = 0.f;
      if (grad_bases != nullptr) {
        for (int r = 0; r < world; ++r) g += grad_bases[r][i];
      } else {
        g = grads[i];
        grads[i] = 0.f;
      }
      float pi = plocal[i];
      adamw_one(pi, g * grad_scale, m[i - lo], v[i - lo], lr, wd, beta1, beta2, eps, step_size, inv_sqrt_bc2);
      for (int r = 0; r < world; ++r) param_bases[r][i] = pi;
    }
  }
}

// NVSwitch (NVLS) form of the same step: grads_mc / params_mc are MULTICAST addresses of the symmetric gradient / parameter
// buffers. multimem.ld_reduce makes the switch read the element from every replica and return the sum (the reduce-scatter
// costs one 16-byte response per float4 instead of world-1 peer loads); multimem.st writes the new parameter values into
// every replica with one store (the all-gather). Per GPU and step only 2 x 124/world MB cross its NVLink ports.
__device__ __forceinline__ float4 multimem_ld_reduce_add_v4(const float* mc_addr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc_addr)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st_v4(float* mc_addr, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float multimem_ld_reduce_add(const float* mc_addr) {
  float r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(r) : "l"(mc_addr) : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float* mc_addr, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc_addr), "f"(v) : "memory");
}

__global__ void __launch_bounds__(256)
adamw_shard_multimem_kernel(float* __restrict__ params_mc, const float* __restrict__ grads_mc, const float* __restrict__ plocal,
                            float* __restrict__ m, float* __restrict__ v, long long lo, long long hi, float lr, float beta1,
                            float beta2, float eps, float wd, float grad_scale, const int* __restrict__ step_dev,
                            const float* __restrict__ lr_dev) {
  pdl_enter();
  if (lr_dev != nullptr) lr = *lr_dev;
  const float st = static_cast<float>(*step_dev);
  const float bc1 = 1.f - powf(beta1, st), bc2 = 1.f - powf(beta2, st);
  const float step_size = lr / bc1, inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const long long n4 = (hi - lo) / 4;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < n4;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = lo + 4 * q;
    const float4 g = multimem_ld_reduce_add_v4(grads_mc + i);
    float4 p4 = *reinterpret_cast<const float4*>(plocal + i);
    float4 m4 = *reinterpret_cast<const float4*>(m + 4 * q);
    float4 v4 = *reinterpret_cast<const float4*>(v + 4 * q);
    adamw_one(p4.x, g.x * grad_scale, m4.x, v4.x, lr, wd, beta1, beta2, eps, step_size, inv_sqrt_bc2);
    adamw_one(p4.y, g.y * grad_scale, m4.y, v4.y, lr, wd, beta1, beta2, eps, step_size, inv_sqrt_bc2);
    adamw_one(p4.z, g.z * grad_scale, m4.z, v4.z, lr, wd, beta1, beta2, eps, step_size, inv_sqrt_bc2);
    adamw_one(p4.w, g.w * grad_scale, m4.w, v4.w, lr, wd, beta1, beta2, eps, step_size, inv_sqrt_bc2);
    *reinterpret_cast<float4*>(m + 4 * q) = m4;
    *reinterpret_cast<float4*>(v + 4 * q) = v4;
    multimem_st_v4(params_mc + i, p4);
  }
  if (blockIdx.x == 0) {
    for (long long i = lo + 4 * n4 + threadIdx.x; i < hi; i += blockDim.x) {
      const float g = multimem_ld_reduce_add(grads_mc + i);
      float pi = plocal[i];
      adamw_one(pi, g * grad_scale, m[i - lo], v[i - lo], lr, wd, beta1, beta2, eps, step_size, inv_sqrt_bc2);
      multimem_st(params_mc + i, pi);
    }
  }
}

// out[i] = sum over replicas of x[lo + i], formed by the switch (multimem.ld_reduce) - the gradient check of the NVLS exchange
__global__ void multimem_reduce_kernel(const float* __restrict__ x_mc, long long lo, long long n, float* __restrict__ out) {
  pdl_enter();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    out[i] = multimem_ld_reduce_add(x_mc + lo + i);
  }
}

// dgrad weights: wd[ci][tap'][co] = w[co][ci][8 - tap'] (180-degree rotated, in/out swapped), bf16; w fp32 [Cout][Cin][3][3]
__global__ void pack_conv3x3_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, __nv_bfloat16* __restrict__ wd) {
  pdl_enter();
  const int total = Cin * 9 * Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % Cout;
    const int tp = (i / Cout) % 9;
    const int ci = i / (9 * Cout);
    wd[i] = __float2bfloat16_rn(w[(static_cast<size_t>(co) * Cin + ci) * 9 + (8 - tp)]);
  }
}

// ConvT dgrad weights: wd[ci][(quad, co)] = w[ci][co][quad] (GEMM N = Cin rows, K = 4f); w fp32 [Cin][f][2][2]
__global__ void pack_convT_dgrad_kernel(const float* __restrict__ w, int Cin, int f, __nv_bfloat16* __restrict__ wd) {
  pdl_enter();
  const int total = Cin * 4 * f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % f;
    const int quad = (i / f) % 4;
    const int ci = i / (4 * f);
    wd[i] = __float2bfloat16_rn(w[(static_cast<size_t>(ci) * f + co) * 4 + quad]);
  }
}

// One launch builds every bf16 operand copy of a training step (43 launches before): a job per weight tensor, blocks
// assigned through the job table's running block count. kind 0: conv 3x3 -> forward layout wp[co][tap][ci] AND the rotated /
// transposed dgrad layout wd[ci][8-tap][co] from ONE read of w; kind 1: ConvT -> wp[(quad,co)][ci] and wd[ci][(quad,co)];
// kind 2: tensor-core stem -> wp[co][k = tap*4+ci] (zero padded to 64).
struct PackJob {
  int kind, block0;                // block0: first block of this job (jobs are sorted by it)
  int Cout, C0, C1;                // PHYSICAL channel counts of the operand copies (conv: Cin = C0 + C1 from two sources)
  int lCout, lC0, lC1;             // LOGICAL counts of the fp32 master tensor; physical channels beyond them are zero-filled
  long long w_off;                 // offset of the fp32 master tensor in the flat parameter buffer
  __nv_bfloat16* wp;               // forward layout (kind 3: a float* - the zero-extended fp32 copy)
  __nv_bfloat16* wd;               // null: no dgrad copy (first layer)
};
constexpr int PACK_ELEMS_PER_BLOCK = 256 * 8;

// kind 0: conv 3x3 -> forward layout wp[co][tap][c] AND the rotated / transposed dgrad layout wd[c][8-tap][co] from ONE read
// of w; kind 1: ConvT -> wp[(quad,co)][ci] and wd[ci][(quad,co)]; kind 2: tensor-core stem -> wp[co][k = tap*4+ci] (zero padded
// to 64); kind 3: fp32 vector (ConvT bias, head weights) zero-extended from lCout to Cout entries.
__device__ __forceinline__ void pack_one_block(const float* __restrict__ params, const PackJob* __restrict__ jobs, int njobs,
                                               int blk) {
  int j = 0;
  while (j + 1 < njobs && blk >= jobs[j + 1].block0) ++j;   // <= 64 jobs
  const PackJob jb = jobs[j];
  const float* w = params + jb.w_off;
  const int i0 = (blk - jb.block0) * PACK_ELEMS_PER_BLOCK + threadIdx.x;   // (kinds 1-3: a flat range of elements per block)
  if (jb.kind == 0) {
    // one block = a tile of 32 output channels x 32 (physical) input channels x 9 taps through shared memory: the fp32 master
    // is read in runs of 288 consecutive floats and both bf16 copies leave in 64-byte runs (the first version wrote the dgrad
    // copy as scattered 2-byte stores: 320 us per step for 124 MB)
    __shared__ float tile[32][289];
    const int Cin = jb.C0 + jb.C1, lCin = jb.lC0 + jb.lC1;
    const int cblocks = Cin / 32;
    const int tb = blk - jb.block0;
    const int co0 = (tb / cblocks) * 32, cp0 = (tb % cblocks) * 32;
    // the tile's 32 physical channels lie in ONE source (C0 is a multiple of 64): logical channel of cp0, valid count
    int ci0, nvalid;
    if (cp0 < jb.C0) {
      ci0 = cp0;
      nvalid = jb.lC0 - cp0;
    } else {
      ci0 = jb.lC0 + (cp0 - jb.C0);
      nvalid = jb.lC1 - (cp0 - jb.C0);
    }
    nvalid = nvalid < 0 ? 0 : (nvalid > 32 ? 32 : nvalid);
    __syncthreads();   // (a block may work several tiles: the previous one has been read out)
#pragma unroll 1
    for (int e0 = threadIdx.x; e0 < 32 * 288; e0 += 256 * 6) {   // 36 elements per thread, six loads in flight
      float f[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int e = e0 + u * 256;
        const int co_l = e / 288, k = e - co_l * 288;   // k = ci_l * 9 + tap
        f[u] = 0.f;
        if (co0 + co_l < jb.lCout && k < nvalid * 9) f[u] = __ldg(w + (static_cast<size_t>(co0 + co_l) * lCin + ci0) * 9 + k);
      }
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int e = e0 + u * 256;
        const int co_l = e / 288, k = e - co_l * 288;
        tile[co_l][k] = f[u];
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int e = threadIdx.x; e < 32 * 288; e += 256) {
      const int l = e & 31, tap = (e >> 5) % 9, o = e / 288;
      // forward layout wp[co][tap][cp]: 32 consecutive cp per (co, tap)
      jb.wp[(static_cast<size_t>(co0 + o) * 9 + tap) * Cin + cp0 + l] = __float2bfloat16_rn(tile[o][l * 9 + tap]);
      // dgrad layout wd[cp][8 - tap][co]: 32 consecutive co per (cp, tap)
      if (jb.wd != nullptr) {
        jb.wd[(static_cast<size_t>(cp0 + o) * 9 + (8 - tap)) * jb.Cout + co0 + l] = __float2bfloat16_rn(tile[l][o * 9 + tap]);
      }
    }
  } else if (jb.kind == 1) {
    const int f = jb.Cout, Cin = jb.C0, lf = jb.lCout, lCin = jb.lC0;
    const int total = 4 * f * Cin;
#pragma unroll 1
    for (int i = i0; i < total && i < i0 + PACK_ELEMS_PER_BLOCK; i += 256) {
      const int ci = i % Cin;
      const int n = i / Cin;          // quad*f + co
      const int co = n % f;
      const int quad = n / f;
      float x = 0.f;
      if (ci < lCin && co < lf) x = w[(static_cast<size_t>(ci) * lf + co) * 4 + quad];
      const __nv_bfloat16 v = __float2bfloat16_rn(x);
      jb.wp[i] = v;
      jb.wd[(static_cast<size_t>(ci) * 4 + quad) * f + co] = v;
    }
  } else if (jb.kind == 2) {
    const int total = jb.Cout * 64;
#pragma unroll 1
    for (int i = i0; i < total && i < i0 + PACK_ELEMS_PER_BLOCK; i += 256) {
      const int k = i % 64, co = i / 64;
      const int tap = k / 4, ci = k % 4;
      float v = 0.f;
      if (tap < 9 && ci < jb.C0 && co < jb.lCout) v = w[(static_cast<size_t>(co) * jb.C0 + ci) * 9 + tap];
      jb.wp[i] = __float2bfloat16_rn(v);
    }
  } else {
    float* dst = reinterpret_cast<float*>(jb.wp);
#pragma unroll 1
    for (int i = i0; i < jb.Cout && i < i0 + PACK_ELEMS_PER_BLOCK; i += 256) dst[i] = i < jb.lCout ? w[i] : 0.f;
  }
}

// Work blocks [block_lo, block_hi) of the job table, grid-stride: the table may be worked off in several launches, and a
// launch with a small grid (a few CTAs per SM, no shared memory) runs beside the tensor-core kernels of another stream
// instead of in front of them.
__global__ void __launch_bounds__(256)
pack_all_kernel(const float* __restrict__ params, const PackJob* __restrict__ jobs, int njobs, int block_lo, int block_hi) {
  pdl_enter();
  for (int blk = block_lo + static_cast<int>(blockIdx.x); blk < block_hi; blk += static_cast<int>(gridDim.x)) {
    pack_one_block(params, jobs, njobs, blk);
  }
}

// packed fp32 gradients -> PyTorch layouts: conv gp[co][tap][ci] -> g[co][ci][tap]; convT gp[quad][co][ci] -> g[ci][co][quad]
__global__ void unpack_conv_grad_kernel(const float* __restrict__ gp, int Cout, int Cin, float* __restrict__ g) {
  pdl_enter();
  const int total = Cout * Cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9;
    const int ci = (i / 9) % Cin;
    const int co = i / (9 * Cin);
    g[i] = gp[(static_cast<size_t>(co) * 9 + tap) * Cin + ci];
  }
}
__global__ void unpack_convT_grad_kernel(const float* __restrict__ gp, int Cin, int f, float* __restrict__ g) {
  pdl_enter();
  const int total = Cin * f * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int quad = i % 4;
    const int co = (i / 4) % f;
    const int ci = i / (4 * f);
    g[i] = gp[(static_cast<size_t>(quad) * f + co) * Cin + ci];
  }
}

// per-channel sum of a bf16 NHWC tensor (ConvT bias gradient): out[c] += sum_p x[p][c]
__global__ void __launch_bounds__(256)
chan_sum_kernel(const uint4* __restrict__ x, int pitch8 /* uint4 per pixel */, size_t npix, int C8, GradRoute route, long long off,
                int lC /* logical channels: the destination has lC entries */) {
  pdl_enter();
  extern __shared__ float red[];
  const int cl = threadIdx.x % C8;
  const int pl = threadIdx.x / C8;
  const int ppb = 256 / C8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t p = static_cast<size_t>(blockIdx.x) * ppb + pl; p < npix; p += static_cast<size_t>(gridDim.x) * ppb) {
    float f[8];
    unpack8(__ldg(x + p * pitch8 + cl), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += f[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = s[i];
  __syncthreads();
  for (int c = threadIdx.x; c < C8 * 8; c += 256) {
    float a = 0.f;
    for (int j = 0; j < ppb; ++j) a += red[(j * C8 + c / 8) * 8 + (c & 7)];
    if (c < lC) grad_add(route, off + c, a);
  }
}

// Stem weight gradient (Cin <= 4, Cout <= 128): dW[co][ci][tap] += sum_p dy[p][co] * x4[p+shift(tap)][ci]
// x4: NHWC4 bf16 network input. Persistent blocks walk 16x16 pixel tiles; a thread owns up to 5 (co, tap) items x 4 input
// channels in registers across ALL its tiles and flushes them with one atomicAdd each at the end.
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const uint2* __restrict__ x4, const __nv_bfloat16* __restrict__ dy, int B, int H, int W, int Cin, int Cout,
                  GradRoute route, long long off /* dW fp32 [Cout][Cin][3][3] */) {
  pdl_enter();
  extern __shared__ float sm[];
  float4* st = reinterpret_cast<float4*>(sm);                 // [18][18] input pixels (4 ch fp32)
  float* sd = sm + 18 * 18 * 4;                                // [256 px][Cout] dy tile (fp32)
  const int tiles_w = (W + 15) / 16, tiles_h = (H + 15) / 16;
  const int total_tiles = tiles_w * tiles_h * B;
  const int n_items = Cout * 9;
  float acc[5][4];
#pragma unroll
  for (int k = 0; k < 5; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int w0 = (tile % tiles_w) * 16;
    const int h0 = ((tile / tiles_w) % tiles_h) * 16;
    const int b = tile / (tiles_w * tiles_h);
    __syncthreads();  // previous tile's smem fully consumed
    for (int i = threadIdx.x; i < 18 * 18; i += blockDim.x) {
      const int hh = h0 + i / 18 - 1, ww = w0 + i % 18 - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const uint2 r = __ldg(x4 + (static_cast<size_t>(b) * H + hh) * W + ww);
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&r.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
        v = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
      }
      st[i] = v;
    }
    for (int i = threadIdx.x; i < 256 * Cout; i += blockDim.x) {
      const int px = i / Cout, co = i % Cout;
      const int hh = h0 + px / 16, ww = w0 + px % 16;
      float v = 0.f;
      if (hh < H && ww < W) v = __bfloat162float(dy[((static_cast<size_t>(b) * H + hh) * W + ww) * Cout + co]);
      sd[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int item = threadIdx.x + k * 256;
      if (item < n_items) {
        const int co = item % Cout, tap = item / Cout;
        const int r = tap / 3, s = tap % 3;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int px = 0; px < 256; ++px) {
          const float g = sd[px * Cout + co];
          const float4 xi = st[(px / 16 + r) * 18 + (px % 16) + s];
          a0 = fmaf(g, xi.x, a0);
          a1 = fmaf(g, xi.y, a1);
          a2 = fmaf(g, xi.z, a2);
          a3 = fmaf(g, xi.w, a3);
        }
        acc[k][0] += a0;
        acc[k][1] += a1;
        acc[k][2] += a2;
        acc[k][3] += a3;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int item = threadIdx.x + k * 256;
    if (item < n_items) {
      const int co = item % Cout, tap = item / Cout;
      for (int ci = 0; ci < Cin; ++ci) grad_add(route, off + (static_cast<long long>(co) * Cin + ci) * 9 + tap, acc[k][ci]);
    }
  }
}

}  // namespace ub
