// Experiment (not product code): can one halo'd activation patch in shared memory feed all nine 3x3
// taps through shifted UMMA shared-memory descriptors?  Patch = [18 rows][pitch pixels][64 ch] bf16,
// 128B-swizzled by TMA; tap (r,s) reads rows (h+r, w+s) for an output tile of 16 rows x 8 pixels:
//   start = patch + (r*pitch + s)*128 B,  SBO = pitch*128 B,  8-row group = 8 consecutive pixels.
// Variants: pitch 10 (dense halo, SBO not a multiple of 1024) or 16 (padded), base_offset field 0 or
// (start>>7)&7.  Single CTA, no pipelining; the result tile is dumped as fp32 [128][64].
#include <cstdio>
#include "../unet-lane-detection_b200/csrc/ptx.cuh"

using namespace ub;

__global__ void __launch_bounds__(128, 1)
halo_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int h0, int w0,
                  int pitch, int use_bo, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // up to 18*16*128 = 36864 B
  uint8_t* sB = smem + 36864;               // 9 * 8192 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 36864 + 73728);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 18 * pitch * 128 + 9 * 8192);
    tma_load_4d(sA, &tmA, &bars[0], 0, w0 - 1, h0 - 1, 0);
    for (int t = 0; t < 9; ++t) tma_load_2d(sB + t * 8192, &tmW, &bars[0], t * 64, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16_f32(128, 64);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int t = 0; t < 9; ++t) {
      const int r = t / 3, s = t % 3;
      const uint32_t start = a0 + (r * pitch + s) * 128;
      const uint32_t bo = use_bo ? ((start >> 7) & 7) : 0;
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = make_sw128_kmajor_desc(start + k * 32, pitch * 128, bo);
        const uint64_t db = make_sw128_kmajor_desc(b0 + t * 8192 + k * 32, 1024, 0);
        umma_f16(tmem, da, db, idesc, (t | k) != 0);
      }
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    float* dst = out + (warp * 32 + lane) * 64 + c * 32;
    for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int halo_probe(const void* x, int H, int W, const void* wp, int h0, int w0, int pitch, int use_bo,
                          float* out) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess) return -1;
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  CUtensorMap ma, mw;
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, 1};
    cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    cuuint32_t box[4] = {64, (cuuint32_t)pitch, 18, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -2;
  }
  {
    cuuint64_t dims[2] = {576, 64};
    cuuint64_t strides[1] = {576 * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t es[2] = {1, 1};
    if (enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -3;
  }
  const int smem = 36864 + 73728 + 64 + 1024;
  if (cudaFuncSetAttribute(halo_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -4;
  halo_probe_kernel<<<1, 128, smem>>>(ma, mw, h0, w0, pitch, use_bo, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    fprintf(stderr, "halo_probe: %s\n", cudaGetErrorString(e));
    return -5;
  }
  return 0;
}
