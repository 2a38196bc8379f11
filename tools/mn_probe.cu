// Experiment: UMMA with MN-major operands (needed for wgrad on NHWC tensors, where the reduction dim = pixels is the
// strided one). D[128 x 64] = A^T B with A global [K=128][M=128], B global [K=128][N=64], bf16, both loaded by TMA as
// [k rows][64 elements] 128B-swizzled blocks. Descriptor: MN-major, SWIZZLE_128B, LBO = bytes between 64-wide MN blocks,
// SBO = bytes between 8-row K groups (variant 1 swaps the two).
#include <cstdio>
#include "../unet-lane-detection_b200/csrc/ptx.cuh"
using namespace ub;

__device__ __forceinline__ uint64_t mn_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
mn_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int variant, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // 2 x 16 KB
  uint8_t* sB = smem + 32768;    // 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    if (elect_one()) {
      mbar_expect_tx(&bars[0], 49152);
      tma_load_2d(sA, &tmA, &bars[0], 0, 0);
      tma_load_2d(sA + 16384, &tmA, &bars[0], 64, 0);
      tma_load_2d(sB, &tmB, &bars[0], 0, 0);
      mbar_wait(&bars[0], 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16_f32(128, 64) | (1u << 15) | (1u << 16);  // A and B MN-major
      const uint32_t lbo = variant == 0 ? 16384 : 1024, sbo = variant == 0 ? 1024 : 16384;
      for (int j = 0; j < 8; ++j) {
        const uint64_t da = mn_desc(smem_u32(sA) + j * 2048, lbo, sbo);
        const uint64_t db = mn_desc(smem_u32(sB) + j * 2048, lbo, sbo);
        umma_f16(tmem, da, db, idesc, j != 0);
      }
      umma_commit(&bars[1]);
    }
    __syncwarp();
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    float* dst = out + (warp * 32 + lane) * 64 + c * 32;
    for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make2d(EncodeTiledFn enc, CUtensorMap* m, const void* p, int cols, int rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -1;
}

extern "C" int mn_probe(const void* A /*[128][128]*/, const void* B /*[128][64]*/, int variant, float* out /*[128][64]*/) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess) return -1;
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  CUtensorMap ma, mb;
  if (make2d(enc, &ma, A, 128, 128) || make2d(enc, &mb, B, 64, 128)) return -2;
  const int smem = 49152 + 64 + 1024;
  cudaFuncSetAttribute(mn_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mn_probe_kernel<<<1, 128, smem>>>(ma, mb, variant, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "mn_probe: %s\n", cudaGetErrorString(e)); return -5; }
  return 0;
}
