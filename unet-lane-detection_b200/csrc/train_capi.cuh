// Training step of the U-Net on B200 (included at the end of capi.cu; uses its file-local helpers).
//
// Reference: README.md:2060-2084 train_one_epoch (model.train(): BatchNorm uses batch statistics, README.md:2062),
// README.md:1855-1893 BCEDiceLoss, README.md:2173-2174 AdamW. The trainer owns a fixed-batch workspace that keeps
// every layer's raw conv output y, its post-BN/ReLU activation a and one gradient buffer per activation.
//
// forward :  for each 3x3 conv   y = conv(x) [tcgen05 implicit GEMM, same kernels as inference, no bias]
//                                 batch stats of y -> scale/shift (+ running stats) -> a = relu(y*scale+shift) (+ 2x2 pool)
//            ConvT / concat / head as in inference (concat never materialised)
// backward:  head -> for each conv in reverse: BN+ReLU backward (two passes, in place) -> wgrad (MN-major tcgen05 GEMM,
//            fp32 atomics straight into the PyTorch-layout gradient) -> dgrad (the forward conv kernel on rotated weights);
//            ConvT backward = 4-source 1-tap GEMM over the quad views; max-pool backward merged with the skip gradient.
// Parameters and gradients are flat fp32 arrays in model.parameters() order (README.md:1427-1447 registration order).
#pragma once
#include "train_kernels.cuh"
#include "wgrad_umma.cuh"
#include "wgrad_halo.cuh"
#include "stem_wgrad_umma.cuh"

namespace {

struct TConv {
  int H, W, C0, C1, Cout;       // PHYSICAL channel counts (multiples of 64; a logical width such as 32 is stored zero-extended)
  int lC0, lC1, lCout;          // LOGICAL counts = shapes of the PyTorch parameters (0 in a memset TConv: same as physical)
  bool stem, pooled;
  uint8_t *x0, *x1;   // inputs (bf16 NHWC); stem: the trainer's copy of the network input (NHWC4)
  uint8_t *y, *a, *p; // raw conv output, relu(bn(y)), maxpool(a) (or null)
  uint8_t *g;         // gradient w.r.t. a, then (in place) w.r.t. y
  uint8_t *dx;        // dgrad target [B,H,W,C0+C1] (null: stem)
  uint8_t *wp, *wd;   // packed forward / dgrad weights (bf16)
  long long w_off, gamma_off, beta_off;
  double *sum, *sumsq;
  float *s1, *s2, *mean, *invstd, *scale, *shift;
  Layer fwd, dg;
  CUtensorMap wX0, wX1, wD;
  ub::WgradArgs wa;
  int w_bn, w_grid;
  bool w_pair;
  bool w_halo;        // Cout == 64: all nine taps from one halo'd patch (wgrad_halo.cuh) instead of one CTA per tap pair
  CUtensorMap wXh0, wXh1, wDh;
  ub::WgradHaloArgs wha;
  bool reduce_fused;  // the BatchNorm-backward sums of this layer come out of the kernel that produces its incoming gradient
  int dg_bn;          // >= 0: this layer's dgrad kernel writes the COMPLETE incoming gradient of conv #dg_bn and takes that
                      // layer's BatchNorm-backward sums in its epilogue (option dgrad_fuse); -1: no
};

struct TConvT {
  int H, W, Cin, f;   // input grid; PHYSICAL channels
  int lCin, lf;       // LOGICAL channels (weight [lCin][lf][2][2], bias [lf])
  float* bias_pad;    // fp32 [f]: the bias zero-extended to the physical width (pack job, every step)
  uint8_t *x, *y;     // input activation [B,H,W,Cin]; output up [B,2H,2W,f]
  uint8_t *dup;       // gradient w.r.t. up: channels [f,2f) of the concat gradient (pixel pitch 2f elements)
  uint8_t *dx;        // gradient w.r.t. x (= g of the producing conv)
  uint8_t *wp, *wd;
  long long w_off, b_off;
  Layer fwd, dg;
  CUtensorMap dgA[4];
  CUtensorMap wX, wDq[4];
  ub::WgradArgs wa;
  int w_bn, w_grid;
  bool w_pair;
  int dg_bn;          // as TConv::dg_bn: the conv whose incoming gradient the ConvT dgrad writes (u.dx == that conv's g)
};

}  // namespace

struct unet_b200_trainer {
  int B, H, W, in_ch, out_ch, levels;
  Opts opt;                    // the switches this trainer was created with
  int feat[UB_MAX_LEVELS];
  std::vector<TConv> convs;    // plan order: enc0.0, enc0.3, ..., bott.0, bott.3, dec0.0, dec0.3, ...
  std::vector<TConvT> ups;     // decoder order (deepest first)
  std::vector<int> fwd_order;  // >= 0: conv index; < 0: -(up index) - 1
  long long n_params;
  std::vector<long long> tensor_off;  // parameters() order, one entry per tensor (+ total at the end)
  long long head_w_off, head_b_off;
  float* head_w_pad;  // fp32 [out_ch][physical f0]: output.weight zero-extended (pack jobs, every step)
  size_t ws_bytes;
  uint8_t* ws;
  uint8_t* acc;       // accumulator region zeroed every step
  size_t acc_bytes;
  float* zero_bias;
  uint8_t* x_in;      // copy of the network input (NHWC4 bf16)
  ub::PackJob* jobs_dev;   // job table of pack_all_kernel (one launch builds every bf16 operand copy of a step)
  int n_jobs, pack_blocks;
  int pack_early_blocks;   // blocks of the jobs the first layers need (stem .. level 1): the rest is packed on the side stream
  int pack_join_conv;      // forward conv index that is the first reader of a late-packed operand
  cudaEvent_t ev_pack[2] = {nullptr, nullptr};
  // weight-gradient side stream (backward): the wgrad GEMM of layer L is forked off after the layer's dgrad and runs next to
  // the HBM-bound BatchNorm / pool backward passes of layer L-1; joined before the backward returns
  cudaStream_t s2 = nullptr;
  std::vector<cudaEvent_t> ev_fork;
  int ev_next = 0;
  int n_forks = 0;        // weight-gradient launches forked to s2 since stage 0
  int bwd_stage = -1;     // last backward stage enqueued
  bool fwd_done;
};

namespace {

struct Bump {
  uintptr_t base;
  size_t off;
  uint8_t* take(size_t bytes) {
    uint8_t* p = reinterpret_cast<uint8_t*>(base + off);
    off += align_up(bytes, 1024);
    return p;
  }
};

bool pow2_times_64(int c) { return c >= 64 && c <= 2048 && (c & (c - 1)) == 0; }

ub::BnBwdStats bn_stats_of(const TConv& c) {
  ub::BnBwdStats b;
  b.y = c.reduce_fused ? reinterpret_cast<const uint4*>(c.y) : nullptr;
  b.scale = c.scale;
  b.shift = c.shift;
  b.mean = c.mean;
  b.invstd = c.invstd;
  b.s1 = c.s1;
  b.s2 = c.s2;
  return b;
}
const ub::BnBwdStats kNoBnStats = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
// the same for a tcgen05 dgrad epilogue (epilogue.cuh EpiBnBwd); `bytes` is filled in by conv_layer_launch
ub::EpiBnBwd epi_bn_of(const TConv& c) {
  ub::EpiBnBwd b;
  b.y = c.y;
  b.scale = c.scale;
  b.shift = c.shift;
  b.mean = c.mean;
  b.invstd = c.invstd;
  b.s1 = c.s1;
  b.s2 = c.s2;
  b.bytes = 0;
  return b;
}

// One pass over the network assigning workspace addresses (base == 0: size computation only).
void trainer_layout(unet_b200_trainer* t, uintptr_t base) {
  Bump bp{base, 0};
  const size_t B = t->B;
  t->zero_bias = reinterpret_cast<float*>(bp.take(4096 * 4));
  t->jobs_dev = reinterpret_cast<ub::PackJob*>(bp.take(64 * sizeof(ub::PackJob)));
  t->head_w_pad = reinterpret_cast<float*>(bp.take((size_t)t->out_ch * t->convs.back().Cout * 4));
  for (TConvT& u : t->ups) u.bias_pad = reinterpret_cast<float*>(bp.take((size_t)u.f * 4));
  t->x_in = bp.take(B * t->H * t->W * 8);
  // accumulators (zeroed per step): per conv sum, sumsq (double) + s1, s2 (float)
  t->acc = reinterpret_cast<uint8_t*>(base + bp.off);
  for (TConv& c : t->convs) {
    c.sum = reinterpret_cast<double*>(bp.take((size_t)c.Cout * 8));
    c.sumsq = reinterpret_cast<double*>(bp.take((size_t)c.Cout * 8));
    c.s1 = reinterpret_cast<float*>(bp.take((size_t)c.Cout * 4));
    c.s2 = reinterpret_cast<float*>(bp.take((size_t)c.Cout * 4));
  }
  t->acc_bytes = base + bp.off - reinterpret_cast<uintptr_t>(t->acc);
  for (TConv& c : t->convs) {
    c.mean = reinterpret_cast<float*>(bp.take((size_t)c.Cout * 4));
    c.invstd = reinterpret_cast<float*>(bp.take((size_t)c.Cout * 4));
    c.scale = reinterpret_cast<float*>(bp.take((size_t)c.Cout * 4));
    c.shift = reinterpret_cast<float*>(bp.take((size_t)c.Cout * 4));
    const size_t cin = c.C0 + c.C1;
    c.wp = bp.take(c.stem ? (size_t)(c.Cout == 64 ? 64 * 64 * 2 : 36 * c.Cout * 4) : (size_t)c.Cout * 9 * cin * 2);
    c.wd = c.stem ? nullptr : bp.take((size_t)c.Cout * 9 * cin * 2);
    const size_t act = B * c.H * c.W * c.Cout * 2;
    c.y = bp.take(act);
    c.a = bp.take(act);
    c.g = bp.take(act);
    c.p = c.pooled ? bp.take(act / 4) : nullptr;
  }
  for (TConvT& u : t->ups) {
    u.wp = bp.take((size_t)4 * u.f * u.Cin * 2);
    u.wd = bp.take((size_t)4 * u.f * u.Cin * 2);
    u.y = bp.take(B * 4 * u.H * u.W * u.f * 2);
  }
  // wiring
  const int L = t->levels;
  uint8_t* cur = t->x_in;
  for (int i = 0; i < L; ++i) {
    TConv& c0 = t->convs[2 * i];
    TConv& c1 = t->convs[2 * i + 1];
    c0.x0 = cur;
    c0.x1 = nullptr;
    c1.x0 = c0.a;
    c1.x1 = nullptr;
    c1.dx = c0.g;
    c0.dx = (i > 0) ? bp.take(B * c0.H * c0.W * c0.C0 * 2) : nullptr;  // gradient w.r.t. the pooled input
    cur = c1.p;
  }
  {
    TConv& b0 = t->convs[2 * L];
    TConv& b1 = t->convs[2 * L + 1];
    b0.x0 = cur;
    b0.x1 = nullptr;
    b0.dx = bp.take(B * b0.H * b0.W * b0.C0 * 2);
    b1.x0 = b0.a;
    b1.x1 = nullptr;
    b1.dx = b0.g;
  }
  TConv* prev = &t->convs[2 * L + 1];
  for (int j = 0; j < L; ++j) {
    const int i = L - 1 - j;
    TConvT& u = t->ups[j];
    TConv& d0 = t->convs[2 * L + 2 + 2 * j];
    TConv& d1 = t->convs[2 * L + 3 + 2 * j];
    TConv& skip = t->convs[2 * i + 1];
    uint8_t* dcat = bp.take(B * d0.H * d0.W * 2 * u.f * 2);
    u.x = prev->a;
    u.dx = prev->g;
    u.dup = dcat + (size_t)u.f * 2;
    d0.x0 = skip.a;
    d0.x1 = u.y;
    d0.dx = dcat;
    d1.x0 = d0.a;
    d1.x1 = nullptr;
    d1.dx = d0.g;
    prev = &d1;
  }
  t->ws_bytes = bp.off;
}

template <int BN>
int launch_wgrad_t(const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap* d, const ub::WgradArgs& a, int grid,
                   int slot, bool pair, cudaStream_t st) {
  using Cfg = ub::WgradCfg<BN>;
  if constexpr (BN >= 128) {
    if (pair) {
      UB_CUDA(ensure_smem(ub::wgrad_umma2_kernel<BN>, AT_WGRAD2 + slot, Cfg::SMEM_BYTES));
      ub_launch(ub::wgrad_umma2_kernel<BN>, grid, 192, Cfg::SMEM_BYTES, st, x0, x1, d[0], d[1], d[2], d[3], a);
      UB_CUDA(cudaGetLastError());
      return UB_OK;
    }
  }
  UB_CUDA(ensure_smem(ub::wgrad_umma_kernel<BN>, AT_WGRAD + slot, Cfg::SMEM_BYTES));
  ub_launch(ub::wgrad_umma_kernel<BN>, grid, 192, Cfg::SMEM_BYTES, st, x0, x1, d[0], d[1], d[2], d[3], a);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

// `pair`: the geometry in `a` (ksplit, stages, grid) was planned for the CTA-pair kernel (wgrad_geometry)
int launch_wgrad(int bn, const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap* d, const ub::WgradArgs& a, int grid,
                 bool pair, cudaStream_t st) {
  switch (bn) {
    case 64: return launch_wgrad_t<64>(x0, x1, d, a, grid, 0, false, st);
    case 128: return launch_wgrad_t<128>(x0, x1, d, a, grid, 1, pair, st);
    case 256: return launch_wgrad_t<256>(x0, x1, d, a, grid, 2, pair, st);
  }
  return fail(UB_ERR_ARG, "unsupported wgrad BLOCK_N %d", bn);
}

// Common part of the weight-gradient launch geometry: pixel tiles, split-K, grid.
int wgrad_geometry(ub::WgradArgs& a, int B, int H, int W, int Cin, int Cout, int taps, int* bn, int* grid, bool* pair) {
  memset(&a, 0, sizeof(a));
  a.B = B;
  a.H = H;
  a.W = W;
  *bn = pick_block_n(Cout);
  // reduction tile = a box of 128 pixels, or 64 for the wide tiles: (2 + BLOCK_N/64) boxes per stage must leave room for a
  // ring deep enough to hide the L2 latency (BLOCK_N 256: 4 stages of 48 KB instead of 2 of 96 KB)
  const int rows_full = (*bn == 256 && tl_opts->wgrad_rows64) ? 64 : 128;
  a.TW = pow2_divisor(W, 16);
  a.TH = pow2_divisor(H, rows_full / a.TW);
  a.TB = rows_full / (a.TW * a.TH);
  a.blk_bytes = rows_full * 128;
  const int tb_eff = a.TB < B ? a.TB : B;
  const int rows = a.TW * a.TH * tb_eff;
  if (rows % 16 != 0) {
    return fail(UB_ERR_ARG, "wgrad: batch %d too small for the %dx%d level (pixel box of %d rows; need a multiple of 16)", B, H,
                W, rows);
  }
  a.k_mmas = rows / 16;
  a.a_bytes = rows * 128;
  a.tiles_w = (W + a.TW - 1) / a.TW;
  a.tiles_h = (H + a.TH - 1) / a.TH;
  a.tiles_b = (B + a.TB - 1) / a.TB;
  a.taps = taps;
  a.Cin = Cin;
  a.Cout = Cout;
  a.n_tiles = Cout / *bn;
  if (taps == 9 && Cin == 64) {
    a.pair_taps = 1;
    a.m_tiles = 5;
  } else if (Cin == 64) {
    // no taps to pair (ConvT quads read different dy views): the 128-row tile holds the 64 channels twice, the second copy
    // is dropped in the epilogue - half the tensor work is wasted on what is a tiny layer (the deployed topology's last ConvT)
    a.pair_taps = 0;
    a.dup_rows = 1;
    a.m_tiles = 1;
  } else {
    if (Cin % 128 != 0) return fail(UB_ERR_ARG, "wgrad: Cin=%d must be 64 or a multiple of 128", Cin);
    a.pair_taps = 0;
    a.m_tiles = Cin / 128;
  }
  a.lc0 = Cin;        // callers with zero-extended tensors overwrite these
  a.lc1 = 0;
  a.lcout = Cout;
  a.rt_total = (a.pair_taps ? 1 : taps) * a.m_tiles;
  // CTA pair: two consecutive row tiles share the dy operand. For a 3x3 conv any two row tiles do (the taps shift only x);
  // the ConvT quads read dy through different views, so there both tiles must belong to the same quad (m_tiles even).
  *pair = tl_opts->wgrad2 && *bn >= 128 && a.rt_total >= 2 && (taps != 4 || a.m_tiles % 2 == 0);
  const int P = *pair ? 2 : 1;
  a.stages = ub::WGRAD_RING_BYTES / ((2 + *bn / 64 / P) * a.blk_bytes);
  if (a.stages > ub::WGRAD_MAX_STAGES) a.stages = ub::WGRAD_MAX_STAGES;
  const int tiles = ((a.rt_total + P - 1) / P) * a.n_tiles;   // work items per K slice (each runs on P CTAs)
  const int ptiles = a.tiles_w * a.tiles_h * a.tiles_b;
  int ks = (cur_sms() / P) / tiles;
  if (ks < 1) ks = 1;
  if (ks > ptiles) ks = ptiles;
  a.ksplit = ks;
  *grid = P * tiles * ks;
  return UB_OK;
}

// weight gradient of a 3x3 conv: dW[co][ci][tap] += sum_p dY[p][co] * x[p + shift(tap)][ci]   (x = cat(x0, x1), dY = c.g)
int setup_conv_wgrad(TConv& c, int B) {
  const int cin = c.C0 + c.C1;
  int rc = wgrad_geometry(c.wa, B, c.H, c.W, cin, c.Cout, 9, &c.w_bn, &c.w_grid, &c.w_pair);
  if (rc != UB_OK) return rc;
  const int lc0 = c.lC0 > 0 ? c.lC0 : c.C0, lc1 = c.C1 > 0 ? (c.lC1 > 0 ? c.lC1 : c.C1) : 0, lcout = c.lCout > 0 ? c.lCout : c.Cout;
  c.wa.shift = 1;
  c.wa.c_split = c.C0;
  c.wa.lc0 = lc0;
  c.wa.lc1 = lc1;
  c.wa.lcout = lcout;
  c.wa.s_co = (long long)(lc0 + lc1) * 9;     // dW is the PyTorch tensor [lCout][lCin][3][3]
  c.wa.s_ci = 9;
  c.wa.s_tap = 1;
  rc = make_act_map(&c.wX0, c.x0, B, c.H, c.W, c.C0, c.wa.TW, c.wa.TH, c.wa.TB);
  if (rc != UB_OK) return rc;
  if (c.C1 > 0) {
    rc = make_act_map(&c.wX1, c.x1, B, c.H, c.W, c.C1, c.wa.TW, c.wa.TH, c.wa.TB);
    if (rc != UB_OK) return rc;
  } else {
    c.wX1 = c.wX0;
  }
  rc = make_act_map(&c.wD, c.g, B, c.H, c.W, c.Cout, c.wa.TW, c.wa.TH, c.wa.TB);
  if (rc != UB_OK) return rc;
  // Cout == 64 on a map with W % 8 == 0 (the 224^2 level): the halo-patch kernel
  c.w_halo = tl_opts->wgrad_halo && c.Cout == 64 && c.W % 8 == 0 && c.C0 % 64 == 0 && c.C1 % 64 == 0;
  if (c.w_halo) {
    ub::WgradHaloArgs& h = c.wha;
    memset(&h, 0, sizeof(h));
    h.B = B;
    h.H = c.H;
    h.W = c.W;
    h.tiles_w = c.W / 8;
    h.tiles_h = (c.H + 15) / 16;
    h.ncb = cin / 64;
    h.cb_split = c.C0 / 64;
    const int total = h.tiles_w * h.tiles_h * B;
    int ksl = cur_sms() / h.ncb;
    if (ksl < 1) ksl = 1;
    if (ksl > total) ksl = total;
    h.kslices = ksl;
    h.stages = ub::WgradHaloCfg::STAGES;
    h.lc0 = lc0;
    h.lc1 = lc1;
    h.lcout = lcout;
    h.s_co = c.wa.s_co;
    h.s_ci = c.wa.s_ci;
    h.s_tap = c.wa.s_tap;
    rc = make_halo_map(&c.wXh0, c.x0, B, c.H, c.W, c.C0);
    if (rc != UB_OK) return rc;
    if (c.C1 > 0) {
      rc = make_halo_map(&c.wXh1, c.x1, B, c.H, c.W, c.C1);
      if (rc != UB_OK) return rc;
    } else {
      c.wXh1 = c.wXh0;
    }
    rc = make_box_map(&c.wDh, c.g, B, c.H, c.W, c.Cout, 8, 16);
    if (rc != UB_OK) return rc;
  }
  return UB_OK;
}

// Backward of ConvTranspose2d(2x2, stride 2): maps for the 4-source dgrad GEMM and the 4-quad weight gradient.
// u.dup points at the gradient w.r.t. the upsampled tensor, whose pixel pitch is `pitch` elements (2f when it is the
// upper half of the concat gradient, f when it stands alone).
int setup_up_backward(TConvT& u, int B, size_t pitch = 0) {
  if (pitch == 0) pitch = 2 * (size_t)u.f;
  const size_t W2 = 2 * (size_t)u.W, H2 = 2 * (size_t)u.H;
  int rc;
  Layer& l = u.dg;
  memset(&l, 0, sizeof(l));
  l.kind = L_CONV;
  l.H = u.H;
  l.W = u.W;
  l.C0 = u.f;
  l.Cout = u.Cin;
  l.relu = 0;
  l.set = true;
  pick_tile(u.H, u.W, &l.TW, &l.TH, &l.TB);
  l.block_n = pick_block_n(u.Cin);
  for (int q = 0; q < 4; ++q) {
    const int dy = q >> 1, dx = q & 1;
    uint8_t* base = u.dup + ((size_t)dy * W2 + dx) * pitch * 2;
    rc = make_tile_store_map(&u.dgA[q], base, B, u.H, u.W, u.f, 2 * pitch, 2 * W2 * pitch, H2 * W2 * pitch, l.TW, l.TH, l.TB);
    if (rc != UB_OK) return rc;
  }
  if (u.wd != nullptr && u.dx != nullptr) {
    rc = make_w_map(&l.mW, u.wd, u.Cin, 4 * u.f, l.block_n);
    if (rc != UB_OK) return rc;
    rc = make_umma_store_maps(l, u.dx, nullptr, B);
    if (rc != UB_OK) return rc;
  }
  // weight gradient: dW[ci][co][quad] += sum_p x[p][ci] * dUp_quad[p][co]
  if (u.x != nullptr) {
    rc = wgrad_geometry(u.wa, B, u.H, u.W, u.Cin, u.f, 4, &u.w_bn, &u.w_grid, &u.w_pair);
    if (rc != UB_OK) return rc;
    const int lcin = u.lCin > 0 ? u.lCin : u.Cin, lf = u.lf > 0 ? u.lf : u.f;
    u.wa.shift = 0;
    u.wa.c_split = u.Cin;
    u.wa.lc0 = lcin;
    u.wa.lc1 = 0;
    u.wa.lcout = lf;
    u.wa.s_ci = 4 * (long long)lf;              // dW is the PyTorch tensor [lCin][lf][2][2]
    u.wa.s_co = 4;
    u.wa.s_tap = 1;
    rc = make_act_map(&u.wX, u.x, B, u.H, u.W, u.Cin, u.wa.TW, u.wa.TH, u.wa.TB);
    if (rc != UB_OK) return rc;
    for (int q = 0; q < 4; ++q) {
      const int dy = q >> 1, dx = q & 1;
      uint8_t* base = u.dup + ((size_t)dy * W2 + dx) * pitch * 2;
      rc = make_tile_store_map(&u.wDq[q], base, B, u.H, u.W, u.f, 2 * pitch, 2 * W2 * pitch, H2 * W2 * pitch, u.wa.TW, u.wa.TH,
                               u.wa.TB);
      if (rc != UB_OK) return rc;
    }
  }
  return UB_OK;
}

int trainer_build_maps(unet_b200_trainer* t) {
  const int B = t->B;
  int rc;
  for (TConv& c : t->convs) {
    const int cin = c.C0 + c.C1;
    if (c.stem) {
      memset(&c.fwd, 0, sizeof(c.fwd));
      if (c.Cout == 64) {
        rc = make_w_map_box(&c.fwd.mW, c.wp, 64, 64, 64);
        if (rc != UB_OK) return rc;
        c.fwd.stem_tw = stem_tile_w();
        rc = make_stem_out_map(&c.fwd.mOut, c.y, B, c.H, c.W, c.fwd.stem_tw);
        if (rc != UB_OK) return rc;
        rc = make_box_map(&c.wD, c.g, B, c.H, c.W, 64, 8, 16);  // dY tiles of the tensor-core stem weight gradient
        if (rc != UB_OK) return rc;
      }
      continue;
    }
    rc = conv_layer_setup(c.fwd, L_CONV, c.x0, c.C0, c.x1, c.C1, c.wp, c.y, nullptr, B, c.H, c.W, c.Cout, 0, true);
    if (rc != UB_OK) return rc;
    rc = conv_layer_setup(c.dg, L_CONV, c.g, c.Cout, nullptr, 0, c.wd, c.dx, nullptr, B, c.H, c.W, cin, 0, true);
    if (rc != UB_OK) return rc;
    if (c.dg_bn >= 0) {
      rc = make_epi_y_map(c.dg, t->convs[c.dg_bn].y, B);
      if (rc != UB_OK) return rc;
    }
    rc = setup_conv_wgrad(c, B);
    if (rc != UB_OK) return rc;
  }
  for (TConvT& u : t->ups) {
    rc = conv_layer_setup(u.fwd, L_CONVT, u.x, u.Cin, nullptr, 0, u.wp, u.y, nullptr, B, u.H, u.W, u.f, 0, false);
    if (rc != UB_OK) return rc;
    rc = setup_up_backward(u, B);
    if (rc != UB_OK) return rc;
    if (u.dg_bn >= 0) {
      rc = make_epi_y_map(u.dg, t->convs[u.dg_bn].y, B);
      if (rc != UB_OK) return rc;
    }
  }
  return UB_OK;
}

int chan_grid(size_t npix, int C8) {
  const int ppb = 256 / C8;
  size_t g = (npix + ppb - 1) / ppb;
  const size_t cap = (size_t)cur_sms() * 8;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

int trainer_conv_forward(unet_b200_trainer* t, TConv& c, int bn_idx, const float* params, float* const* running_mean,
                         float* const* running_var, float momentum, float eps, cudaStream_t st) {
  const int B = t->B;
  int rc;
  // the tensor-core kernels accumulate the per-channel batch statistics of their bf16 output in the epilogue
  bool stats_done = true;
  if (c.stem) {
    if (c.Cout == 64) {
      rc = launch_stem_umma(c.fwd.mW, c.fwd.mOut, c.x0, t->zero_bias, B, c.H, c.W, 0, st, c.sum, c.sumsq, c.fwd.stem_tw);
      if (rc != UB_OK) return rc;
    } else {
      const int tiles = ((c.W + 15) / 16) * ((c.H + 15) / 16) * B;
      const size_t smem = (size_t)(36 * c.Cout + c.Cout) * 4 + 18 * 18 * 16;
      ub_launch(ub::stem_conv_kernel<false>, tiles, 256, smem, st, reinterpret_cast<const uint2*>(c.x0), reinterpret_cast<const float*>(c.wp),
                                                      t->zero_bias, B, c.H, c.W, c.C0, c.Cout, 0,
                                                      reinterpret_cast<__nv_bfloat16*>(c.y));
      UB_CUDA(cudaGetLastError());
      stats_done = false;
    }
  } else {
    rc = conv_layer_launch(c.fwd, B, B, t->zero_bias, c.y, nullptr, st, c.sum, c.sumsq);
    if (rc != UB_OK) return rc;
  }
  const size_t npix = (size_t)B * c.H * c.W;
  const int C8 = c.Cout / 8;
  if (!stats_done) {
    ub_launch(ub::chan_stats_kernel, chan_grid(npix, C8), 256, 2 * 2048 * 4, st, reinterpret_cast<const uint4*>(c.y), npix, C8, c.sum,
                                                                           c.sumsq);
    UB_CUDA(cudaGetLastError());
  }
  ub::BnFin fin;
  fin.sum = c.sum;
  fin.sumsq = c.sumsq;
  fin.count = (float)npix;
  fin.eps = eps;
  fin.momentum = momentum;
  fin.gamma = params + c.gamma_off;
  fin.beta = params + c.beta_off;
  fin.mean = c.mean;
  fin.invstd = c.invstd;
  fin.scale = c.scale;
  fin.shift = c.shift;
  fin.running_mean = running_mean ? running_mean[bn_idx] : nullptr;
  fin.running_var = running_var ? running_var[bn_idx] : nullptr;
  fin.C = c.Cout;
  fin.lC = c.lCout > 0 ? c.lCout : c.Cout;
  // (no bn_finalize launch: the apply kernel finalises the statistics itself, train_kernels.cuh bn_fin_channels; the grid
  // always has at least C8 threads - one block is 256 >= C8 threads - so every channel gets published)
  if (c.pooled) {
    const size_t n = npix / 4 * C8;
    ub_launch(ub::bn_relu_apply_pool_kernel, grid_for(n, 256), 256, 0, st, reinterpret_cast<const uint4*>(c.y), c.scale, c.shift, B, c.H,
                                                                    c.W, C8, reinterpret_cast<uint4*>(c.a),
                                                                    reinterpret_cast<uint4*>(c.p), fin);
  } else {
    const size_t n8 = npix * C8;
    ub_launch(ub::bn_relu_apply_kernel, grid_for(n8, 256), 256, 0, st, reinterpret_cast<const uint4*>(c.y), c.scale, c.shift, n8, C8,
                                                                reinterpret_cast<uint4*>(c.a), fin);
  }
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

// BatchNorm(train)+ReLU backward on c.g in place (dA -> dY). The per-channel sums (sum g -> d beta, sum g*xhat -> d gamma)
// are accumulated locally in s1 / s2 (zeroed by the caller) - the apply pass needs THIS replica's sums - and then published
// into the flat gradient at off_gamma / off_beta through `route` (which may point at another GPU).
int conv_bn_backward(const TConv& c, int B, float* s1, float* s2, const ub::GradRoute& route, long long off_gamma,
                     long long off_beta, cudaStream_t st) {
  const size_t npix = (size_t)B * c.H * c.W;
  const int C8 = c.Cout / 8;
  const uint4* g4 = reinterpret_cast<const uint4*>(c.g);
  const uint4* y4 = reinterpret_cast<const uint4*>(c.y);
  if (!c.reduce_fused) {   // (fused: s1 / s2 were accumulated by the kernel that wrote c.g - max-pool or head backward)
    ub_launch(ub::bn_relu_bwd_reduce_kernel, chan_grid(npix, C8), 256, 2 * 2048 * 4, st, g4, y4, c.scale, c.shift, c.mean, c.invstd,
              npix, C8, s1, s2);
    UB_CUDA(cudaGetLastError());
  }
  const size_t n8 = npix * C8;
  ub_launch(ub::bn_relu_bwd_apply_kernel, grid_for(n8, 256), 256, 0, st, g4, y4, c.scale, c.shift, c.mean, c.invstd, s1, s2,
                                                                  1.f / (float)npix, n8, C8, reinterpret_cast<uint4*>(c.g), route,
                                                                  off_gamma, off_beta, c.lCout > 0 ? c.lCout : c.Cout);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

// Weight gradient of one conv (3x3 on tensor cores, stem on tensor cores for Cout == 64), accumulated into dw (PyTorch layout).
int conv_wgrad_launch(const TConv& c, int B, const ub::GradRoute& route, long long off, cudaStream_t st) {
  if (c.stem && c.Cout == 64) {
    UB_CUDA(ensure_smem(ub::stem_wgrad_umma_kernel, AT_STEM_WGRAD_TC, ub::StemWgradCfg::SMEM_BYTES));
    ub::StemWgradArgs sa;
    sa.B = B;
    sa.H = c.H;
    sa.W = c.W;
    sa.Cin = c.C0;
    sa.lCout = c.lCout > 0 ? c.lCout : c.Cout;
    sa.tiles_w = (c.W + 7) / 8;
    sa.tiles_h = (c.H + 15) / 16;
    sa.x = reinterpret_cast<const uint2*>(c.x0);
    sa.route = route;
    sa.off = off;
    const int pairs = (sa.tiles_w * sa.tiles_h * B + 1) / 2;
    const int grid = pairs < cur_sms() ? pairs : cur_sms();
    ub_launch(ub::stem_wgrad_umma_kernel, grid, ub::StemWgradCfg::THREADS, ub::StemWgradCfg::SMEM_BYTES, st, c.wD, sa);
    UB_CUDA(cudaGetLastError());
    return UB_OK;
  }
  if (c.stem) {
    if (c.Cout > 128) return fail(UB_ERR_ARG, "stem weight gradient supports Cout <= 128");
    const size_t smem = (size_t)(18 * 18 * 4 + 256 * c.Cout) * 4;
    UB_CUDA(ensure_smem(ub::stem_wgrad_kernel, AT_STEM_WGRAD, (18 * 18 * 4 + 256 * 128) * 4));
    const int tiles = ((c.W + 15) / 16) * ((c.H + 15) / 16) * B;
    const int grid = tiles < 2 * cur_sms() ? tiles : 2 * cur_sms();
    ub_launch(ub::stem_wgrad_kernel, grid, 256, smem, st, reinterpret_cast<const uint2*>(c.x0),
                                                    reinterpret_cast<const __nv_bfloat16*>(c.g), B, c.H, c.W, c.C0, c.Cout, route, off);
    UB_CUDA(cudaGetLastError());
    return UB_OK;
  }
  if (c.w_halo) {
    UB_CUDA(ensure_smem(ub::wgrad_halo_kernel, AT_WGRAD_HALO, ub::WgradHaloCfg::SMEM_BYTES));
    ub::WgradHaloArgs ha = c.wha;
    ha.route = route;
    ha.off = off;
    ub_launch(ub::wgrad_halo_kernel, ha.ncb * ha.kslices, ub::WgradHaloCfg::THREADS, ub::WgradHaloCfg::SMEM_BYTES, st, c.wXh0, c.wXh1,
              c.wDh, ha);
    UB_CUDA(cudaGetLastError());
    return UB_OK;
  }
  ub::WgradArgs wa = c.wa;
  wa.route = route;
  wa.off = off;
  const CUtensorMap d[4] = {c.wD, c.wD, c.wD, c.wD};
  return launch_wgrad(c.w_bn, c.wX0, c.wX1, d, wa, c.w_grid, c.w_pair, st);
}

// Stream the next weight-gradient launch goes to: the side stream, made to wait for everything enqueued on `st` so far
// (fork), or `st` itself when the overlap is switched off. Works the same eagerly and under stream capture.
int wgrad_stream(unet_b200_trainer* t, cudaStream_t st, cudaStream_t* out) {
  *out = st;
  if (!tl_opts->wgrad_stream || t->s2 == nullptr) return UB_OK;
  if (t->ev_next >= (int)t->ev_fork.size()) return fail(UB_ERR_STATE, "out of fork events");
  cudaEvent_t ev = t->ev_fork[t->ev_next++];
  UB_CUDA(cudaEventRecord(ev, st));
  UB_CUDA(cudaStreamWaitEvent(t->s2, ev, 0));
  ++t->n_forks;
  *out = t->s2;
  return UB_OK;
}

int trainer_conv_backward(unet_b200_trainer* t, TConv& c, const ub::GradRoute& route, cudaStream_t st) {
  const int B = t->B;
  int rc = conv_bn_backward(c, B, c.s1, c.s2, route, c.gamma_off, c.beta_off, st);
  if (rc != UB_OK) return rc;
  // dgrad first (the next layer's backward waits for it), then the wgrad is forked: it only feeds the gradient buffer, so it
  // runs concurrently with the bandwidth-bound BN / pool backward kernels that follow on `st`
  if (c.dx != nullptr) {
    if (c.dg_bn >= 0) {   // the epilogue also takes the BatchNorm-backward sums of the layer whose gradient it writes
      const ub::EpiBnBwd bn = epi_bn_of(t->convs[c.dg_bn]);
      rc = conv_layer_launch(c.dg, B, B, t->zero_bias, c.dx, nullptr, st, nullptr, nullptr, &bn);
    } else {
      rc = conv_layer_launch(c.dg, B, B, t->zero_bias, c.dx, nullptr, st);
    }
    if (rc != UB_OK) return rc;
  }
  cudaStream_t sw;
  rc = wgrad_stream(t, st, &sw);
  if (rc != UB_OK) return rc;
  return conv_wgrad_launch(c, B, route, c.w_off, sw);
}

// ConvT backward pieces on prepared maps: bias gradient + weight gradient, and the input gradient.
int up_wgrad_launch(const TConvT& u, int B, int pitch8, const ub::GradRoute& route, long long off_w, long long off_b,
                    cudaStream_t st) {
  const size_t npix_up = (size_t)B * 4 * u.H * u.W;
  const int C8 = u.f / 8;
  if (off_b >= 0) {
    ub_launch(ub::chan_sum_kernel, chan_grid(npix_up, C8), 256, 2048 * 4, st, reinterpret_cast<const uint4*>(u.dup), pitch8, npix_up, C8,
                                                                       route, off_b, u.lf > 0 ? u.lf : u.f);
    UB_CUDA(cudaGetLastError());
  }
  ub::WgradArgs wa = u.wa;
  wa.route = route;
  wa.off = off_w;
  return launch_wgrad(u.w_bn, u.wX, u.wX, u.wDq, wa, u.w_grid, u.w_pair, st);
}

// dX[b,h,w,ci] = sum_quad sum_co dUp[b,2h+dy,2w+dx,co] * w[ci][co][quad]: 1-tap GEMM, K walks the four quad views
int up_dgrad_launch(const TConvT& u, int B, const float* zero_bias, cudaStream_t st, const ub::EpiBnBwd* bn = nullptr) {
  ub::ConvArgs a = conv_args(u.dg, B, B, zero_bias, u.dx, nullptr);
  a.taps = 1;
  a.kc0 = a.kc1 = a.kc2 = a.kc3 = u.f / 64;
  if (bn != nullptr) {
    if (u.dg.y_bytes == 0) return fail(UB_ERR_STATE, "fused BN-backward sums without a y view");
    a.bn = *bn;
    a.bn.bytes = u.dg.y_bytes;
  }
  return launch_conv(u.dg.block_n, u.dgA, u.dg.mW, u.dg.mO, a, st, bn != nullptr ? &u.dg.mY : nullptr);
}

int trainer_up_backward(unet_b200_trainer* t, TConvT& u, const ub::GradRoute& route, cudaStream_t st) {
  int rc;
  if (u.dg_bn >= 0) {
    const ub::EpiBnBwd bn = epi_bn_of(t->convs[u.dg_bn]);
    rc = up_dgrad_launch(u, t->B, t->zero_bias, st, &bn);
  } else {
    rc = up_dgrad_launch(u, t->B, t->zero_bias, st);
  }
  if (rc != UB_OK) return rc;
  cudaStream_t sw;
  rc = wgrad_stream(t, st, &sw);
  if (rc != UB_OK) return rc;
  return up_wgrad_launch(u, t->B, 2 * (u.f / 8), route, u.w_off, u.b_off, sw);
}

// 4096 zero floats for the single-op entry points (allocated once per device)
int get_zero_bias(float** out) {
  DevState* d = cur_dev();
  float* zb = d->zero_bias.load(std::memory_order_acquire);
  if (zb == nullptr) {
    UB_CUDA(cudaMalloc(&zb, 4096 * 4));
    UB_CUDA(cudaMemset(zb, 0, 4096 * 4));
    float* expected = nullptr;
    if (!d->zero_bias.compare_exchange_strong(expected, zb)) {   // another thread got there first
      cudaFree(zb);
      zb = expected;
    }
  }
  *out = zb;
  return UB_OK;
}

}  // namespace

extern "C" {

int unet_b200_trainer_create(unet_b200_trainer** out, int batch, int H, int W, int in_channels, int out_channels,
                             const int* features, int levels) {
  if (out == nullptr || features == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (levels < 1 || levels > UB_MAX_LEVELS) return fail(UB_ERR_ARG, "levels must be in [1,%d]", UB_MAX_LEVELS);
  if (batch < 1) return fail(UB_ERR_ARG, "batch must be >= 1");
  if (in_channels < 1 || in_channels > 4) return fail(UB_ERR_ARG, "in_channels must be in [1,4] (got %d)", in_channels);
  if (out_channels < 1 || out_channels > ub::HEAD_MAX_OC) {
    return fail(UB_ERR_ARG, "training needs out_channels in [1,%d] (got %d)", ub::HEAD_MAX_OC, out_channels);
  }
  if (H % (1 << levels) != 0 || W % (1 << levels) != 0) {
    return fail(UB_ERR_ARG, "H=%d and W=%d must be divisible by 2^levels=%d", H, W, 1 << levels);
  }
  // Logical widths (the PyTorch parameter shapes) are multiples of 32; tensors are stored with 64-aligned PHYSICAL channel counts
  // whose extra channels are exact zeros (zero weights; BatchNorm scale = shift = 0 there), as in the inference plan. The
  // per-channel kernels need the physical widths to be powers of two.
  int fp[UB_MAX_LEVELS];
  for (int i = 0; i < levels; ++i) {
    fp[i] = (features[i] + 63) / 64 * 64;
    if (features[i] <= 0 || features[i] % 32 != 0 || !pow2_times_64(fp[i]) || fp[i] > 1024) {
      return fail(UB_ERR_ARG, "training needs features[%d]=%d to be a multiple of 32 whose 64-aligned width is in {64,128,256,512,1024}",
                  i, features[i]);
    }
  }
  if (fp[0] > 128) return fail(UB_ERR_ARG, "training needs features[0] <= 128 (stem weight gradient)");
  if (fp[0] == 128 && features[0] != 128) return fail(UB_ERR_ARG, "training needs features[0] in {32, 64, 128}");
  const int lbott = 2 * features[levels - 1], pbott = (lbott + 63) / 64 * 64;
  if (!pow2_times_64(pbott)) return fail(UB_ERR_ARG, "bottleneck width %d not supported", lbott);
  unet_b200_trainer* t = new (std::nothrow) unet_b200_trainer();
  if (t == nullptr) return fail(UB_ERR_ARG, "out of host memory");
  t->B = batch;
  t->H = H;
  t->W = W;
  t->in_ch = in_channels;
  t->out_ch = out_channels;
  t->levels = levels;
  t->ws = nullptr;
  t->fwd_done = false;
  t->opt = g_opts;
  for (int i = 0; i < levels; ++i) t->feat[i] = features[i];

  auto mk = [&](int h, int w, int c0, int c1, int cout, int lc0, int lc1, int lcout, bool stem, bool pooled) {
    TConv c;
    memset(&c, 0, sizeof(c));
    c.H = h;
    c.W = w;
    c.C0 = c0;
    c.C1 = c1;
    c.Cout = cout;
    c.lC0 = lc0;
    c.lC1 = lc1;
    c.lCout = lcout;
    c.stem = stem;
    c.pooled = pooled;
    t->convs.push_back(c);
  };
  int cin = in_channels, lcin = in_channels;
  for (int i = 0; i < levels; ++i) {
    const int h = H >> i, w = W >> i, f = fp[i], lf = features[i];
    mk(h, w, cin, 0, f, lcin, 0, lf, i == 0, false);
    mk(h, w, f, 0, f, lf, 0, lf, false, true);
    cin = f;
    lcin = lf;
  }
  {
    const int h = H >> levels, w = W >> levels;
    mk(h, w, cin, 0, pbott, lcin, 0, lbott, false, false);
    mk(h, w, pbott, 0, pbott, lbott, 0, lbott, false, false);
    cin = pbott;
    lcin = lbott;
  }
  for (int j = 0; j < levels; ++j) {
    const int i = levels - 1 - j;
    const int h = H >> i, w = W >> i, f = fp[i], lf = features[i];
    TConvT u;
    memset(&u, 0, sizeof(u));
    u.H = h / 2;
    u.W = w / 2;
    u.Cin = cin;
    u.f = f;
    u.lCin = lcin;
    u.lf = lf;
    if (lcin != 2 * lf) {
      delete t;
      return fail(UB_ERR_ARG, "decoder level %d: ConvT input channels %d != 2*%d", j, lcin, lf);
    }
    t->ups.push_back(u);
    mk(h, w, f, f, f, lf, lf, lf, false, false);
    mk(h, w, f, 0, f, lf, 0, lf, false, false);
    cin = f;
    lcin = lf;
  }
  // BatchNorm-backward sums fused into the producer of the incoming gradient where that producer is an elementwise kernel:
  // the encoder conv1 layers (max-pool backward) and the last conv (head backward)
  if (t->opt.bwd_fuse) {   // 1: all five layers, 2: the head's only
    if (t->opt.bwd_fuse == 1) {
      for (int i = 0; i < levels; ++i) t->convs[2 * i + 1].reduce_fused = true;
    }
    t->convs.back().reduce_fused = out_channels == 1;
  }
  // ... and into the tcgen05 dgrad epilogue where the producer is a dgrad GEMM that writes the complete gradient: the conv1 of
  // every block (its dgrad writes the block's conv0 gradient) and the ConvTs (their dgrad writes the gradient of the conv1
  // below them: the bottleneck's or the previous decoder level's). The stem's gradient comes from enc0.conv1's dgrad too.
  for (TConv& c : t->convs) c.dg_bn = -1;
  for (TConvT& u : t->ups) u.dg_bn = -1;
  if (t->opt.dgrad_fuse) {
    const int nblocks = 2 * levels + 1;
    for (int b = 0; b < nblocks; ++b) {
      t->convs[2 * b + 1].dg_bn = 2 * b;
      t->convs[2 * b].reduce_fused = true;
    }
    for (int j = 0; j < levels; ++j) {
      const int below = (j == 0) ? 2 * levels + 1 : 2 * levels + 3 + 2 * (j - 1);
      t->ups[j].dg_bn = below;
      t->convs[below].reduce_fused = true;
    }
  }
  // forward order and parameters() order (encoder, decoder, bottleneck, output: README.md:1427-1447)
  for (int i = 0; i < 2 * levels + 2; ++i) t->fwd_order.push_back(i);
  for (int j = 0; j < levels; ++j) {
    t->fwd_order.push_back(-(j + 1));
    t->fwd_order.push_back(2 * levels + 2 + 2 * j);
    t->fwd_order.push_back(2 * levels + 3 + 2 * j);
  }
  long long off = 0;
  auto conv_params = [&](TConv& c) {
    const long long cn = c.lC0 + c.lC1;     // parameter shapes are the LOGICAL ones
    c.w_off = off;
    t->tensor_off.push_back(off);
    off += (long long)c.lCout * cn * 9;
    c.gamma_off = off;
    t->tensor_off.push_back(off);
    off += c.lCout;
    c.beta_off = off;
    t->tensor_off.push_back(off);
    off += c.lCout;
  };
  for (int i = 0; i < 2 * levels; ++i) conv_params(t->convs[i]);
  for (int j = 0; j < levels; ++j) {
    TConvT& u = t->ups[j];
    u.w_off = off;
    t->tensor_off.push_back(off);
    off += (long long)u.lCin * u.lf * 4;
    u.b_off = off;
    t->tensor_off.push_back(off);
    off += u.lf;
    conv_params(t->convs[2 * levels + 2 + 2 * j]);
    conv_params(t->convs[2 * levels + 3 + 2 * j]);
  }
  conv_params(t->convs[2 * levels]);
  conv_params(t->convs[2 * levels + 1]);
  t->head_w_off = off;
  t->tensor_off.push_back(off);
  off += (long long)out_channels * features[0];
  t->head_b_off = off;
  t->tensor_off.push_back(off);
  off += out_channels;
  t->tensor_off.push_back(off);
  t->n_params = off;
  trainer_layout(t, 0);
  *out = t;
  return UB_OK;
}

void unet_b200_trainer_destroy(unet_b200_trainer* t) {
  if (t == nullptr) return;
  for (cudaEvent_t e : t->ev_fork) {
    if (e != nullptr) cudaEventDestroy(e);
  }
  for (cudaEvent_t e : t->ev_pack) {
    if (e != nullptr) cudaEventDestroy(e);
  }
  if (t->s2 != nullptr) cudaStreamDestroy(t->s2);
  delete t;
}
size_t unet_b200_trainer_workspace_bytes(const unet_b200_trainer* t) { return t ? t->ws_bytes : 0; }
long long unet_b200_trainer_num_params(const unet_b200_trainer* t) { return t ? t->n_params : 0; }
int unet_b200_trainer_num_tensors(const unet_b200_trainer* t) { return t ? (int)t->tensor_off.size() - 1 : 0; }
long long unet_b200_trainer_tensor_offset(const unet_b200_trainer* t, int idx) {
  if (t == nullptr || idx < 0 || idx >= (int)t->tensor_off.size()) return -1;
  return t->tensor_off[idx];
}

int unet_b200_trainer_bind(unet_b200_trainer* t, void* workspace_dev) {
  if (t == nullptr || workspace_dev == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (reinterpret_cast<uintptr_t>(workspace_dev) & 1023) return fail(UB_ERR_ARG, "workspace must be 1024-byte aligned");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  OptScope opt_scope(&t->opt);
  t->ws = static_cast<uint8_t*>(workspace_dev);
  trainer_layout(t, reinterpret_cast<uintptr_t>(workspace_dev));
  UB_CUDA(cudaMemset(t->zero_bias, 0, 4096 * 4));
  t->fwd_done = false;
  if (t->s2 == nullptr) {
    UB_CUDA(cudaStreamCreateWithFlags(&t->s2, cudaStreamNonBlocking));
    t->ev_fork.resize(t->convs.size() + t->ups.size() + 2 + 2 * (2 * t->levels + 3));
    for (cudaEvent_t& e : t->ev_fork) UB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaEvent_t& e : t->ev_pack) UB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  {
    std::vector<ub::PackJob> jobs;
    int blocks = 0;
    auto add = [&](int kind, int Cout, int C0, int C1, int lCout, int lC0, int lC1, long long w_off, void* wp, void* wd,
                   size_t elems) {
      ub::PackJob j;
      j.kind = kind;
      j.Cout = Cout;
      j.C0 = C0;
      j.C1 = C1;
      j.lCout = lCout;
      j.lC0 = lC0;
      j.lC1 = lC1;
      j.block0 = blocks;
      j.w_off = w_off;
      j.wp = reinterpret_cast<__nv_bfloat16*>(wp);
      j.wd = reinterpret_cast<__nv_bfloat16*>(wd);
      jobs.push_back(j);
      // kind 0: one block per 32 x 32-channel tile (pack_one_block); the others: flat ranges of PACK_ELEMS_PER_BLOCK elements
      blocks += kind == 0 ? (Cout / 32) * ((C0 + C1) / 32) : (int)((elems + ub::PACK_ELEMS_PER_BLOCK - 1) / ub::PACK_ELEMS_PER_BLOCK);
    };
    // the jobs are in first-use order; the operands of the first EARLY convs (stem, enc0.conv1, level 1) are packed on the
    // caller's stream, everything else on the side stream while those layers run (train_forward)
    const int EARLY = t->levels >= 2 ? 4 : 0;
    t->pack_early_blocks = 0;
    t->pack_join_conv = EARLY;
    int ci_idx = 0;
    for (TConv& c : t->convs) {
      if (c.stem && c.Cout == 64) {
        add(2, c.Cout, c.C0, 0, c.lCout, c.C0, 0, c.w_off, c.wp, nullptr, (size_t)c.Cout * 64);
      } else if (!c.stem) {
        add(0, c.Cout, c.C0, c.C1, c.lCout, c.lC0, c.lC1, c.w_off, c.wp, c.wd, (size_t)c.Cout * 9 * (c.C0 + c.C1));
      }
      if (++ci_idx == EARLY) t->pack_early_blocks = blocks;
    }
    for (TConvT& u : t->ups) {
      add(1, u.f, u.Cin, 0, u.lf, u.lCin, 0, u.w_off, u.wp, u.wd, (size_t)4 * u.f * u.Cin);
      add(3, u.f, 0, 0, u.lf, 0, 0, u.b_off, u.bias_pad, nullptr, (size_t)u.f);                 // bias, zero-extended
    }
    for (int oc = 0; oc < t->out_ch; ++oc) {      // one zero-extended row per output channel
      add(3, t->convs.back().Cout, 0, 0, t->feat[0], 0, 0, t->head_w_off + (long long)oc * t->feat[0],
          t->head_w_pad + (size_t)oc * t->convs.back().Cout, nullptr, (size_t)t->convs.back().Cout);
    }
    if (jobs.size() > 64) return fail(UB_ERR_ARG, "too many weight tensors for the pack job table");
    t->n_jobs = (int)jobs.size();
    t->pack_blocks = blocks;
    UB_CUDA(cudaMemcpy(t->jobs_dev, jobs.data(), jobs.size() * sizeof(ub::PackJob), cudaMemcpyHostToDevice));
  }
  return trainer_build_maps(t);
}

int unet_b200_train_forward(unet_b200_trainer* t, const void* x_nhwc4, const float* params, float* const* running_mean,
                            float* const* running_var, float momentum, float eps, float* logits, void* stream) {
  if (t == nullptr || x_nhwc4 == nullptr || params == nullptr || logits == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (t->ws == nullptr) return fail(UB_ERR_STATE, "trainer is not bound");
  OptScope opt_scope(&t->opt);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = t->B;
  UB_CUDA(cudaMemcpyAsync(t->x_in, x_nhwc4, (size_t)B * t->H * t->W * 8, cudaMemcpyDeviceToDevice, st));
  UB_CUDA(cudaMemsetAsync(t->acc, 0, t->acc_bytes, st));
  // bf16 operand copies of the current fp32 parameters (forward layout + the rotated / transposed dgrad layout): ONE launch
  // (the operands of the first layers on `st`; the bulk - 124 MB of fp32 masters - on the side stream under those layers)
  const bool pack_split = tl_opts->pack_split && tl_opts->wgrad_stream && t->s2 != nullptr && t->pack_early_blocks > 0 &&
                          t->pack_early_blocks < t->pack_blocks;
  if (pack_split) {
    // first the small launch the stem is waiting for, then the bulk as a SMALL persistent grid on the side stream: launched
    // with one CTA per work block it would sit in front of the stem in the block scheduler (measured: no overlap at all)
    ub_launch(ub::pack_all_kernel, t->pack_early_blocks, 256, 0, st, params, t->jobs_dev, t->n_jobs, 0, t->pack_early_blocks);
    UB_CUDA(cudaGetLastError());
    UB_CUDA(cudaEventRecord(t->ev_pack[0], st));
    UB_CUDA(cudaStreamWaitEvent(t->s2, t->ev_pack[0], 0));
    int late = t->pack_blocks - t->pack_early_blocks;
    const int cap = 4 * cur_sms();
    ub_launch(ub::pack_all_kernel, late < cap ? late : cap, 256, 0, t->s2, params, t->jobs_dev, t->n_jobs, t->pack_early_blocks,
              t->pack_blocks);
    UB_CUDA(cudaGetLastError());
    UB_CUDA(cudaEventRecord(t->ev_pack[1], t->s2));
  } else {
    ub_launch(ub::pack_all_kernel, t->pack_blocks, 256, 0, st, params, t->jobs_dev, t->n_jobs, 0, t->pack_blocks);
  }
  UB_CUDA(cudaGetLastError());
  for (TConv& c : t->convs) {
    if (c.stem && c.Cout != 64) {   // FP32-pipe stem (widths other than 64): fp32 weights, its own small kernel
      ub_launch(ub::pack_stem_kernel, grid_for(36 * c.Cout, 256), 256, 0, st, params + c.w_off, nullptr, nullptr, nullptr, nullptr, 0.f,
                                                                       c.Cout, c.C0, reinterpret_cast<float*>(c.wp), c.s1, 0, c.Cout);
      UB_CUDA(cudaGetLastError());
    }
  }
  // (no bias without BN: s1 / s2 stay zero - cleared with the accumulator region above - until the backward uses them)
  int rc;
  bool pack_joined = !pack_split;
  for (int id : t->fwd_order) {
    if (!pack_joined && (id < 0 || id >= t->pack_join_conv)) {   // first layer whose operands were packed on the side stream
      UB_CUDA(cudaStreamWaitEvent(st, t->ev_pack[1], 0));
      pack_joined = true;
    }
    if (id >= 0) {
      rc = trainer_conv_forward(t, t->convs[id], id, params, running_mean, running_var, momentum, eps, st);
    } else {
      TConvT& u = t->ups[-id - 1];
      rc = conv_layer_launch(u.fwd, B, B, u.bias_pad, u.y, nullptr, st);
    }
    if (rc != UB_OK) return rc;
  }
  const TConv& last = t->convs.back();
  const size_t npix = (size_t)B * last.H * last.W;
  if (t->out_ch == 1) {
    ub_launch(ub::head_fwd_train_kernel, grid_for(npix * 8, 256), 256, 0, st, reinterpret_cast<const uint4*>(last.a),
                                                                       t->head_w_pad, params + t->head_b_off, npix,
                                                                       last.Cout / 8, logits);
  } else {
    ub_launch(ub::head_fwd_train_multi_kernel, grid_for(npix, 256), 256, 0, st, reinterpret_cast<const uint4*>(last.a), t->head_w_pad,
              params + t->head_b_off, npix, (size_t)last.H * last.W, last.Cout / 8, t->out_ch, logits);
  }
  UB_CUDA(cudaGetLastError());
  t->fwd_done = true;
  return UB_OK;
}

// ---- backward in stages ------------------------------------------------------------------------------------------------
// The backward pass is cut where a contiguous range of the flat gradient becomes final, so that a data-parallel caller can
// start exchanging that range while the rest of the backward still runs (decoder gradients are final first):
//   stage 0            head (output.weight / output.bias); also clears the gradient buffer
//   stage 1 .. L       decoder level j = L - stage  (shallowest level first: ConvT j + its double conv)
//   stage L + 1        bottleneck
//   stage L + 2 .. 2L+1  encoder level i = 2L + 1 - stage
// A stage only ENQUEUES work (on `st` and on the trainer's weight-gradient side stream); unet_b200_trainer_join makes a
// stream wait for all of it.
static int trainer_num_stages(const unet_b200_trainer* t) { return 2 * t->levels + 2; }

static void stage_range(const unet_b200_trainer* t, int stage, long long* lo, long long* hi) {
  const int L = t->levels;
  if (stage == 0) {
    *lo = t->head_w_off;
    *hi = t->n_params;
  } else if (stage <= L) {
    const int j = L - stage;
    *lo = t->ups[j].w_off;
    *hi = (j + 1 < L) ? t->ups[j + 1].w_off : t->convs[2 * L].w_off;   // next decoder level, or the bottleneck
  } else if (stage == L + 1) {
    *lo = t->convs[2 * L].w_off;
    *hi = t->head_w_off;
  } else {
    const int i = 2 * L + 1 - stage;
    *lo = t->convs[2 * i].w_off;
    *hi = (i + 1 < L) ? t->convs[2 * i + 2].w_off : t->ups[0].w_off;
  }
}

static int train_backward_stage_impl(unet_b200_trainer* t, int stage, const float* dlogits, const float* params,
                                     const ub::GradRoute& route, bool zero_local, cudaStream_t st) {
  if (t == nullptr || params == nullptr || route.local == nullptr) return fail(UB_ERR_ARG, "null argument");
  const int B = t->B, L = t->levels;
  if (stage < 0 || stage >= trainer_num_stages(t)) return fail(UB_ERR_ARG, "stage %d outside [0,%d)", stage, trainer_num_stages(t));
  OptScope opt_scope(&t->opt);
  int rc;
  if (stage == 0) {
    if (dlogits == nullptr) return fail(UB_ERR_ARG, "null argument");
    if (!t->fwd_done) return fail(UB_ERR_STATE, "train_backward needs a preceding train_forward");
    t->ev_next = 0;
    t->n_forks = 0;
    t->bwd_stage = 0;
    if (zero_local) UB_CUDA(cudaMemsetAsync(route.local, 0, (size_t)t->n_params * 4, st));
    // s1 received the pack kernels' (all-zero) bias in the forward and s2 was cleared with the accumulator region: both are
    // zero here, once per forward/backward pair
    TConv& last = t->convs.back();
    const size_t npix = (size_t)B * last.H * last.W;
    const int C8 = last.Cout / 8;
    if (t->out_ch == 1) {
      if (last.reduce_fused) {
        ub_launch(ub::head_bwd_kernel<true>, chan_grid(npix, C8), 256, 2 * 2048 * 4, st,
            reinterpret_cast<const uint4*>(last.a), dlogits, t->head_w_pad, npix, C8, reinterpret_cast<uint4*>(last.g), route,
            t->head_w_off, t->head_b_off, bn_stats_of(last), t->feat[0]);
      } else {
        ub_launch(ub::head_bwd_kernel<false>, chan_grid(npix, C8), 256, 2 * 2048 * 4, st,
            reinterpret_cast<const uint4*>(last.a), dlogits, t->head_w_pad, npix, C8, reinterpret_cast<uint4*>(last.g), route,
            t->head_w_off, t->head_b_off, kNoBnStats, t->feat[0]);
      }
    } else {
      ub_launch(ub::head_bwd_multi_kernel, chan_grid(npix, C8), 256, (2048 + 256) * 4, st,
          reinterpret_cast<const uint4*>(last.a), dlogits, t->head_w_pad, npix, (size_t)last.H * last.W, C8, t->out_ch,
          reinterpret_cast<uint4*>(last.g), route, t->head_w_off, t->head_b_off, t->feat[0]);
    }
    UB_CUDA(cudaGetLastError());
  } else {
    if (t->bwd_stage != stage - 1) return fail(UB_ERR_STATE, "backward stages must run in order (got %d after %d)", stage, t->bwd_stage);
    if (stage <= L) {
      const int j = L - stage;
      rc = trainer_conv_backward(t, t->convs[2 * L + 3 + 2 * j], route, st);
      if (rc != UB_OK) return rc;
      rc = trainer_conv_backward(t, t->convs[2 * L + 2 + 2 * j], route, st);
      if (rc != UB_OK) return rc;
      rc = trainer_up_backward(t, t->ups[j], route, st);
      if (rc != UB_OK) return rc;
    } else if (stage == L + 1) {
      rc = trainer_conv_backward(t, t->convs[2 * L + 1], route, st);
      if (rc != UB_OK) return rc;
      rc = trainer_conv_backward(t, t->convs[2 * L], route, st);
      if (rc != UB_OK) return rc;
    } else {
      const int i = 2 * L + 1 - stage;
      TConv& c1 = t->convs[2 * i + 1];
      const TConv& next0 = t->convs[2 * i + 2];                 // consumer of the pooled tensor (next encoder level / bottleneck)
      const TConv& d0 = t->convs[2 * L + 2 + 2 * (L - 1 - i)];  // decoder conv that consumed the skip
      const int C8 = c1.Cout / 8;
      const size_t n = (size_t)B * (c1.H / 2) * (c1.W / 2) * C8;
      if (c1.reduce_fused) {
        ub_launch(ub::maxpool_bwd_add_kernel<true>, grid_for(n, 256), 256, 2 * 2048 * 4, st,
            reinterpret_cast<const uint4*>(c1.a), reinterpret_cast<const uint4*>(next0.dx), reinterpret_cast<const uint4*>(d0.dx),
            2 * C8, B, c1.H, c1.W, C8, reinterpret_cast<uint4*>(c1.g), bn_stats_of(c1));
      } else {
        ub_launch(ub::maxpool_bwd_add_kernel<false>, grid_for(n, 256), 256, 0, st,
            reinterpret_cast<const uint4*>(c1.a), reinterpret_cast<const uint4*>(next0.dx), reinterpret_cast<const uint4*>(d0.dx),
            2 * C8, B, c1.H, c1.W, C8, reinterpret_cast<uint4*>(c1.g), kNoBnStats);
      }
      UB_CUDA(cudaGetLastError());
      rc = trainer_conv_backward(t, c1, route, st);
      if (rc != UB_OK) return rc;
      rc = trainer_conv_backward(t, t->convs[2 * i], route, st);
      if (rc != UB_OK) return rc;
    }
  }
  t->bwd_stage = stage;
  if (stage == trainer_num_stages(t) - 1) t->fwd_done = false;
  return UB_OK;
}

// `waiter` waits for everything the backward stages have enqueued so far: the work on `st` and the forked weight-gradient GEMMs.
static int trainer_join_impl(unet_b200_trainer* t, cudaStream_t st, cudaStream_t waiter) {
  if (t->ev_next + 2 > (int)t->ev_fork.size()) return fail(UB_ERR_STATE, "out of join events");
  if (t->s2 != nullptr && t->n_forks > 0) {
    cudaEvent_t e2 = t->ev_fork[t->ev_next++];
    UB_CUDA(cudaEventRecord(e2, t->s2));
    UB_CUDA(cudaStreamWaitEvent(waiter, e2, 0));
  }
  if (waiter != st) {
    cudaEvent_t e1 = t->ev_fork[t->ev_next++];
    UB_CUDA(cudaEventRecord(e1, st));
    UB_CUDA(cudaStreamWaitEvent(waiter, e1, 0));
  }
  return UB_OK;
}

static int train_backward_impl(unet_b200_trainer* t, const float* dlogits, const float* params, const ub::GradRoute& route,
                               bool zero_local, cudaStream_t st) {
  if (t == nullptr) return fail(UB_ERR_ARG, "null argument");
  for (int s = 0; s < trainer_num_stages(t); ++s) {
    int rc = train_backward_stage_impl(t, s, dlogits, params, route, zero_local, st);
    if (rc != UB_OK) return rc;
  }
  return trainer_join_impl(t, st, st);   // the gradient is complete on `st` when every forked wgrad has finished
}

int unet_b200_train_backward(unet_b200_trainer* t, const float* dlogits, const float* params, float* grads, void* stream) {
  ub::GradRoute route{nullptr, grads, 0u};
  return train_backward_impl(t, dlogits, params, route, true, static_cast<cudaStream_t>(stream));
}

int unet_b200_trainer_num_stages(const unet_b200_trainer* t) { return t ? trainer_num_stages(t) : 0; }

int unet_b200_trainer_stage_range(const unet_b200_trainer* t, int stage, long long* lo, long long* hi) {
  if (t == nullptr || lo == nullptr || hi == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (stage < 0 || stage >= trainer_num_stages(t)) return fail(UB_ERR_ARG, "stage %d outside [0,%d)", stage, trainer_num_stages(t));
  stage_range(t, stage, lo, hi);
  return UB_OK;
}

int unet_b200_train_backward_stage(unet_b200_trainer* t, int stage, const float* dlogits, const float* params, float* grads,
                                   void* stream) {
  ub::GradRoute route{nullptr, grads, 0u};
  return train_backward_stage_impl(t, stage, dlogits, params, route, true, static_cast<cudaStream_t>(stream));
}

int unet_b200_trainer_join(unet_b200_trainer* t, void* stream, void* waiter_stream) {
  if (t == nullptr) return fail(UB_ERR_ARG, "null argument");
  return trainer_join_impl(t, static_cast<cudaStream_t>(stream), static_cast<cudaStream_t>(waiter_stream));
}

int unet_b200_train_backward_p2p(unet_b200_trainer* t, const float* dlogits, const float* params, float* grads_local,
                                 float* const* grad_bases_dev, int world, void* stream) {
  if (grad_bases_dev == nullptr || world < 1) return fail(UB_ERR_ARG, "bad peer table");
  if (t == nullptr) return fail(UB_ERR_ARG, "null argument");
  const unsigned shard = (unsigned)(((t->n_params + world - 1) / world + 3) / 4 * 4);   // as in adamw_step_p2p
  ub::GradRoute route{grad_bases_dev, grads_local, shard};
  return train_backward_impl(t, dlogits, params, route, false, static_cast<cudaStream_t>(stream));
}

// AdamW over NVLink on the flat range [lo, hi) (lo a multiple of 4): sum of the range over every rank's gradient buffer
// (peer loads; grad_bases == NULL: the local buffer already holds the sum and is cleared), update, store to every rank's
// parameters. m / v hold hi - lo elements: the caller's optimizer state for exactly this range.
int unet_b200_adamw_range_p2p(float* const* param_bases_dev, float* const* grad_bases_dev, int world, int rank, float* grads_local,
                              long long lo, long long hi, float* m, float* v, float lr, const float* lr_dev, float beta1,
                              float beta2, float eps, float weight_decay, const int* step_dev, float grad_scale, void* stream) {
  if (param_bases_dev == nullptr || grads_local == nullptr || m == nullptr || v == nullptr || step_dev == nullptr) {
    return fail(UB_ERR_ARG, "null argument");
  }
  if (world < 1 || rank < 0 || rank >= world) return fail(UB_ERR_ARG, "bad rank / world");
  if (lo < 0 || (lo & 3)) return fail(UB_ERR_ARG, "range start %lld must be a non-negative multiple of 4", lo);
  int rc = device_check();
  if (rc != UB_OK) return rc;
  if (hi <= lo) return UB_OK;
  ub_launch(ub::adamw_shard_allgather_kernel, grid_for((size_t)(hi - lo + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream),
      param_bases_dev, grad_bases_dev, world, rank, grads_local, m, v, lo, hi, lr, beta1, beta2, eps, weight_decay, grad_scale,
      step_dev, lr_dev);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_adamw_step_p2p(float* const* param_bases_dev, float* const* grad_bases_dev, int world, int rank,
                             float* grads_local, float* exp_avg_shard,
                             float* exp_avg_sq_shard, long long n, float lr, float beta1, float beta2, float eps,
                             float weight_decay, const int* step_dev, float grad_scale, void* stream) {
  if (world < 1 || rank < 0 || rank >= world) return fail(UB_ERR_ARG, "bad rank / world");
  const long long shard = ((n + world - 1) / world + 3) / 4 * 4;   // multiple of 4: 16-byte accesses stay aligned
  const long long lo = shard * rank;
  long long hi = lo + shard;
  if (hi > n) hi = n;
  return unet_b200_adamw_range_p2p(param_bases_dev, grad_bases_dev, world, rank, grads_local, lo, hi, exp_avg_shard,
                                   exp_avg_sq_shard, lr, nullptr, beta1, beta2, eps, weight_decay, step_dev, grad_scale, stream);
}

// NVSwitch form on the flat range [lo, hi): multimem.ld_reduce of the gradients, multimem.st of the new parameters.
int unet_b200_adamw_range_multimem(float* params_mc, const float* grads_mc, const float* params_local, long long lo, long long hi,
                                   float* m, float* v, float lr, const float* lr_dev, float beta1, float beta2, float eps,
                                   float weight_decay, const int* step_dev, float grad_scale, void* stream) {
  if (params_mc == nullptr || grads_mc == nullptr || params_local == nullptr || m == nullptr || v == nullptr || step_dev == nullptr) {
    return fail(UB_ERR_ARG, "null argument");
  }
  if (lo < 0 || (lo & 3)) return fail(UB_ERR_ARG, "range start %lld must be a non-negative multiple of 4", lo);
  int rc = device_check();
  if (rc != UB_OK) return rc;
  if (hi <= lo) return UB_OK;
  ub_launch(ub::adamw_shard_multimem_kernel, grid_for((size_t)(hi - lo + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream),
      params_mc, grads_mc, params_local, m, v, lo, hi, lr, beta1, beta2, eps, weight_decay, grad_scale, step_dev, lr_dev);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_adamw_step_multimem(float* params_mc, const float* grads_mc, const float* params_local, int world, int rank,
                                  float* exp_avg_shard, float* exp_avg_sq_shard, long long n, float lr, float beta1, float beta2,
                                  float eps, float weight_decay, const int* step_dev, float grad_scale, void* stream) {
  if (world < 1 || rank < 0 || rank >= world) return fail(UB_ERR_ARG, "bad rank / world");
  const long long shard = ((n + world - 1) / world + 3) / 4 * 4;
  const long long lo = shard * rank;
  long long hi = lo + shard;
  if (hi > n) hi = n;
  return unet_b200_adamw_range_multimem(params_mc, grads_mc, params_local, lo, hi, exp_avg_shard, exp_avg_sq_shard, lr, nullptr,
                                        beta1, beta2, eps, weight_decay, step_dev, grad_scale, stream);
}

int unet_b200_multimem_reduce(const float* x_mc, long long lo, long long n, float* out, void* stream) {
  if (x_mc == nullptr || out == nullptr || n < 0) return fail(UB_ERR_ARG, "bad argument");
  if (n == 0) return UB_OK;
  ub_launch(ub::multimem_reduce_kernel, grid_for((size_t)n, 256), 256, 0, static_cast<cudaStream_t>(stream), x_mc, lo, n, out);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_bce_dice_loss(const float* logits, const float* target, size_t n, float pos_weight, float bce_weight,
                            float dice_weight, float smooth, double* scratch4, float* losses3, float* dlogits, void* stream) {
  if (logits == nullptr || target == nullptr || scratch4 == nullptr || losses3 == nullptr) return fail(UB_ERR_ARG, "null argument");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UB_CUDA(cudaMemsetAsync(scratch4, 0, 4 * sizeof(double), st));
  ub_launch(ub::bce_dice_reduce_kernel, grid_for(n, 256), 256, 0, st, logits, target, n, pos_weight, scratch4);
  UB_CUDA(cudaGetLastError());
  ub_launch(ub::bce_dice_grad_kernel, grid_for(n, 256), 256, 0, st, logits, target, n, pos_weight, bce_weight, dice_weight,
                                                             smooth, scratch4, dlogits, losses3);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_validation_metrics(const float* logits, const float* target, size_t n, float pos_weight, float bce_weight,
                                 float dice_weight, float smooth, float threshold, double* scratch6, float* out4, void* stream) {
  if (logits == nullptr || target == nullptr || scratch6 == nullptr || out4 == nullptr) return fail(UB_ERR_ARG, "null argument");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UB_CUDA(cudaMemsetAsync(scratch6, 0, 6 * sizeof(double), st));
  ub_launch(ub::val_metrics_reduce_kernel, grid_for(n, 256), 256, 0, st, logits, target, n, pos_weight, threshold, scratch6);
  UB_CUDA(cudaGetLastError());
  ub_launch(ub::val_metrics_finalize_kernel, 1, 1, 0, st, scratch6, n, bce_weight, dice_weight, smooth, out4);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                         float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
  if (params == nullptr || grads == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (step < 1) return fail(UB_ERR_ARG, "step counts from 1");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  ub_launch(ub::adamw_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), params, grads, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                                   beta2, eps, weight_decay, bc1, bc2, grad_scale,
                                                                                   nullptr, nullptr);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_adamw_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                             const float* lr_dev, float beta1, float beta2, float eps, float weight_decay, const int* step_dev,
                             float grad_scale, void* stream) {
  if (params == nullptr || grads == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr || step_dev == nullptr) {
    return fail(UB_ERR_ARG, "null argument");
  }
  int rc = device_check();
  if (rc != UB_OK) return rc;
  ub_launch(ub::adamw_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), params, grads, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                                   beta2, eps, weight_decay, 1.f, 1.f, grad_scale,
                                                                                   step_dev, lr_dev);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

// ---- single training ops (parity tests; the trainer above runs the same code on prebuilt maps) --------------------
int unet_b200_pack_conv3x3_dgrad(const float* w, int Cout, int Cin, void* wd, void* stream) {
  if (w == nullptr || wd == nullptr) return fail(UB_ERR_ARG, "null argument");
  ub_launch(ub::pack_conv3x3_dgrad_kernel, grid_for((size_t)Cout * 9 * Cin, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      w, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(wd));
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_pack_convT2x2_dgrad(const float* w, int Cin, int f, void* wd, void* stream) {
  if (w == nullptr || wd == nullptr) return fail(UB_ERR_ARG, "null argument");
  ub_launch(ub::pack_convT_dgrad_kernel, grid_for((size_t)4 * f * Cin, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      w, Cin, f, reinterpret_cast<__nv_bfloat16*>(wd));
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_conv3x3_wgrad(const void* x0, int C0, const void* x1, int C1, const void* dy, int B, int H, int W, int Cout,
                            float* dw, void* stream) {
  if (x0 == nullptr || dy == nullptr || dw == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (C0 % 64 != 0 || C1 % 64 != 0 || Cout % 64 != 0 || C0 <= 0 || C1 < 0) return fail(UB_ERR_ARG, "channels must be multiples of 64");
  if (C1 > 0 && x1 == nullptr) return fail(UB_ERR_ARG, "x1 is null but C1 > 0");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  TConv c;
  memset(&c, 0, sizeof(c));
  c.H = H;
  c.W = W;
  c.C0 = C0;
  c.C1 = C1;
  c.Cout = Cout;
  c.x0 = static_cast<uint8_t*>(const_cast<void*>(x0));
  c.x1 = static_cast<uint8_t*>(const_cast<void*>(x1));
  c.g = static_cast<uint8_t*>(const_cast<void*>(dy));
  rc = setup_conv_wgrad(c, B);
  if (rc != UB_OK) return rc;
  return conv_wgrad_launch(c, B, ub::GradRoute{nullptr, dw, 0u}, 0, static_cast<cudaStream_t>(stream));
}

int unet_b200_stem_wgrad(const void* x4, const void* dy, int B, int H, int W, int Cin, int Cout, float* dw, void* stream) {
  if (x4 == nullptr || dy == nullptr || dw == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (Cin < 1 || Cin > 4) return fail(UB_ERR_ARG, "stem Cin must be in [1,4]");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  TConv c;
  memset(&c, 0, sizeof(c));
  c.stem = true;
  c.H = H;
  c.W = W;
  c.C0 = Cin;
  c.Cout = Cout;
  c.x0 = static_cast<uint8_t*>(const_cast<void*>(x4));
  c.g = static_cast<uint8_t*>(const_cast<void*>(dy));
  if (Cout == 64) {
    rc = make_box_map(&c.wD, c.g, B, H, W, 64, 8, 16);
    if (rc != UB_OK) return rc;
  }
  return conv_wgrad_launch(c, B, ub::GradRoute{nullptr, dw, 0u}, 0, static_cast<cudaStream_t>(stream));
}

int unet_b200_convT2x2_wgrad(const void* x, int Cin, const void* dup, int dup_pitch, int B, int H, int W, int f, float* dw,
                             float* dbias, void* stream) {
  if (x == nullptr || dup == nullptr || dw == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (Cin % 128 != 0 || f % 64 != 0 || dup_pitch < f || dup_pitch % 8 != 0) return fail(UB_ERR_ARG, "bad channel counts / pitch");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  TConvT u;
  memset(&u, 0, sizeof(u));
  u.H = H;
  u.W = W;
  u.Cin = Cin;
  u.f = f;
  u.x = static_cast<uint8_t*>(const_cast<void*>(x));
  u.dup = static_cast<uint8_t*>(const_cast<void*>(dup));
  rc = setup_up_backward(u, B, (size_t)dup_pitch);
  if (rc != UB_OK) return rc;
  // dw and dbias are separate caller buffers: two local routes
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dbias != nullptr) {
    const size_t npix_up = (size_t)B * 4 * H * W;
    ub_launch(ub::chan_sum_kernel, chan_grid(npix_up, f / 8), 256, 2048 * 4, st, reinterpret_cast<const uint4*>(u.dup), dup_pitch / 8,
                                                                          npix_up, f / 8, ub::GradRoute{nullptr, dbias, 0u}, 0, f);
    UB_CUDA(cudaGetLastError());
  }
  return up_wgrad_launch(u, B, dup_pitch / 8, ub::GradRoute{nullptr, dw, 0u}, 0, -1, st);
}

int unet_b200_convT2x2_dgrad(const void* dup, int dup_pitch, const void* wd, int B, int H, int W, int Cin, int f, void* dx,
                             void* stream) {
  if (dup == nullptr || wd == nullptr || dx == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (Cin % 64 != 0 || f % 64 != 0 || dup_pitch < f || dup_pitch % 8 != 0) return fail(UB_ERR_ARG, "bad channel counts / pitch");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  float* zb;
  rc = get_zero_bias(&zb);
  if (rc != UB_OK) return rc;
  TConvT u;
  memset(&u, 0, sizeof(u));
  u.H = H;
  u.W = W;
  u.Cin = Cin;
  u.f = f;
  u.dup = static_cast<uint8_t*>(const_cast<void*>(dup));
  u.wd = static_cast<uint8_t*>(const_cast<void*>(wd));
  u.dx = static_cast<uint8_t*>(dx);
  rc = setup_up_backward(u, B, (size_t)dup_pitch);
  if (rc != UB_OK) return rc;
  return up_dgrad_launch(u, B, zb, static_cast<cudaStream_t>(stream));
}

int unet_b200_bn_relu_train_fwd(const void* y, const float* gamma, const float* beta, int B, int H, int W, int C, float eps,
                                float momentum, float* running_mean, float* running_var, void* a, void* pool, float* stats4,
                                double* scratch2, void* stream) {
  if (y == nullptr || gamma == nullptr || beta == nullptr || a == nullptr || stats4 == nullptr || scratch2 == nullptr) {
    return fail(UB_ERR_ARG, "null argument");
  }
  if (!pow2_times_64(C)) return fail(UB_ERR_ARG, "C=%d must be a power of two in [64,2048]", C);
  if (pool != nullptr && ((H | W) & 1)) return fail(UB_ERR_ARG, "pooling needs even H and W");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t npix = (size_t)B * H * W;
  const int C8 = C / 8;
  UB_CUDA(cudaMemsetAsync(scratch2, 0, (size_t)2 * C * sizeof(double), st));
  ub_launch(ub::chan_stats_kernel, chan_grid(npix, C8), 256, 2 * 2048 * 4, st, reinterpret_cast<const uint4*>(y), npix, C8, scratch2,
                                                                         scratch2 + C);
  UB_CUDA(cudaGetLastError());
  ub::BnFin fin;
  fin.sum = scratch2;
  fin.sumsq = scratch2 + C;
  fin.count = (float)npix;
  fin.eps = eps;
  fin.momentum = momentum;
  fin.gamma = gamma;
  fin.beta = beta;
  fin.mean = stats4;
  fin.invstd = stats4 + C;
  fin.scale = stats4 + 2 * C;
  fin.shift = stats4 + 3 * C;
  fin.running_mean = running_mean;
  fin.running_var = running_var;
  fin.C = C;
  fin.lC = C;
  ub_launch(ub::bn_finalize_kernel, (C + 127) / 128, 128, 0, st, fin);
  UB_CUDA(cudaGetLastError());
  ub::BnFin no_fin;   // this entry point keeps the two-kernel form (finalise, then apply from the published scale / shift)
  memset(&no_fin, 0, sizeof(no_fin));
  if (pool != nullptr) {
    ub_launch(ub::bn_relu_apply_pool_kernel, grid_for(npix / 4 * C8, 256), 256, 0, st, reinterpret_cast<const uint4*>(y), stats4 + 2 * C,
                                                                                stats4 + 3 * C, B, H, W, C8,
                                                                                reinterpret_cast<uint4*>(a),
                                                                                reinterpret_cast<uint4*>(pool), no_fin);
  } else {
    ub_launch(ub::bn_relu_apply_kernel, grid_for(npix * C8, 256), 256, 0, st, reinterpret_cast<const uint4*>(y), stats4 + 2 * C,
                                                                       stats4 + 3 * C, npix * C8, C8, reinterpret_cast<uint4*>(a), no_fin);
  }
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

int unet_b200_bn_relu_bwd(void* g, const void* y, const float* stats4, int B, int H, int W, int C, float* dgamma, float* dbeta,
                          void* stream) {
  if (g == nullptr || y == nullptr || stats4 == nullptr || dgamma == nullptr || dbeta == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (!pow2_times_64(C)) return fail(UB_ERR_ARG, "C=%d must be a power of two in [64,2048]", C);
  int rc = device_check();
  if (rc != UB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TConv c;
  memset(&c, 0, sizeof(c));
  c.H = H;
  c.W = W;
  c.Cout = C;
  c.g = static_cast<uint8_t*>(g);
  c.y = static_cast<uint8_t*>(const_cast<void*>(y));
  float* s4 = const_cast<float*>(stats4);
  c.mean = s4;
  c.invstd = s4 + C;
  c.scale = s4 + 2 * C;
  c.shift = s4 + 3 * C;
  UB_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)C * 4, st));
  UB_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)C * 4, st));
  // the outputs double as the local accumulators; nothing to publish
  return conv_bn_backward(c, B, dbeta, dgamma, ub::GradRoute{nullptr, nullptr, 0u}, 0, 0, st);
}

int unet_b200_maxpool2x2_bwd(const void* a, const void* dP, const void* dskip, int skip_pitch, int B, int H, int W, int C,
                             void* dA, void* stream) {
  if (a == nullptr || dP == nullptr || dA == nullptr) return fail(UB_ERR_ARG, "null argument");
  if (C % 8 != 0 || ((H | W) & 1) || (dskip != nullptr && (skip_pitch < C || skip_pitch % 8 != 0))) return fail(UB_ERR_ARG, "bad shape");
  int rc = device_check();
  if (rc != UB_OK) return rc;
  const int C8 = C / 8;
  const size_t n = (size_t)B * (H / 2) * (W / 2) * C8;
  ub_launch(ub::maxpool_bwd_add_kernel<false>, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream),
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(dP), reinterpret_cast<const uint4*>(dskip),
      dskip ? skip_pitch / 8 : C8, B, H, W, C8, reinterpret_cast<uint4*>(dA), kNoBnStats);
  UB_CUDA(cudaGetLastError());
  return UB_OK;
}

}  // extern "C"
