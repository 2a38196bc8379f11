// Experiment: tcgen05.mma issue/throughput rate for M=128, N in {64,128,256}, K=16 (bf16), operands in smem.
// One CTA per SM (grid = #SMs) to see the rate under full-chip power conditions; data is garbage (zeros).
#include <cstdio>
#include "../unet-lane-detection_b200/csrc/ptx.cuh"
using namespace ub;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16_f32(128, N);
    const uint64_t da = make_sw128_kmajor_desc(smem_u32(smem), 1024, 0);
    const uint64_t db = make_sw128_kmajor_desc(smem_u32(smem + 16384), 1024, 0);
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem, da + 2 * k, db + 2 * k, idesc, 1);
      }
      umma_commit(bar);
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(bar, 0);
    long long t2 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int N> void run(int iters) {
  long long* d; cudaMalloc(&d, 16); long long h[2];
  const int smem = 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    mma_rate_kernel<N><<<148, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d error %s\n", N, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)\n", N, double(h[0]) / (4.0 * iters), double(h[1]) / (4.0 * iters), N / 2);
  cudaFree(d);
}
int main() { run<64>(2000); run<128>(2000); run<256>(2000); run<32>(2000); return 0; }
