/* The drop-in boundary from plain C: no Python, no torch - include/unet_b200.h + libunet_b200.so + the CUDA runtime for
 * device memory. Builds the reference's default network (README.md:1424: UNet(3, 1, [64,128,256,512])) with seeded random
 * weights, pushes a few uint8 frames through unet_b200_infer_u8_host_stream and checks what a C caller can check without an
 * oracle: status codes, the mask being exactly (prob > threshold) * 255 of the returned probabilities (src/unet.py:63-67),
 * run-to-run determinism of inference. Without an sm_100 device it stops after the host-side part (plan geometry, argument
 * validation), which is what the CPU test suite runs.
 *   gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/c_abi_demo.c -o c_abi_demo \
 *       -Lunet-lane-detection_b200 -lunet_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/unet-lane-detection_b200 -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "unet_b200.h"

#define CHECK_UB(call)                                                                          \
  do {                                                                                          \
    int rc_ = (call);                                                                           \
    if (rc_ != UB_OK) {                                                                         \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, unet_b200_last_error());                    \
      return 1;                                                                                 \
    }                                                                                           \
  } while (0)
#define CHECK_CUDA(call)                                                                        \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "%s -> %s\n", #call, cudaGetErrorString(e_));                             \
      return 1;                                                                                 \
    }                                                                                           \
  } while (0)

static unsigned long long rng_state = 0x2545F4914F6CDD1DULL;
static float frand(void) { /* xorshift64*, uniform in [-1, 1) */
  rng_state ^= rng_state >> 12;
  rng_state ^= rng_state << 25;
  rng_state ^= rng_state >> 27;
  return (float)((rng_state * 0x2545F4914F6CDD1DULL) >> 40) / 8388608.0f - 1.0f;
}

/* fp32 host array -> fresh device array */
static float* upload(const float* h, size_t n) {
  float* d = NULL;
  if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess) return NULL;
  if (cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
  return d;
}
static float* random_device(size_t n, float scale) {
  float* h = (float*)malloc(n * sizeof(float));
  float* d;
  size_t i;
  if (h == NULL) return NULL;
  for (i = 0; i < n; ++i) h[i] = frand() * scale;
  d = upload(h, n);
  free(h);
  return d;
}
static float* const_device(size_t n, float v) {
  float* h = (float*)malloc(n * sizeof(float));
  float* d;
  size_t i;
  if (h == NULL) return NULL;
  for (i = 0; i < n; ++i) h[i] = v;
  d = upload(h, n);
  free(h);
  return d;
}

int main(void) {
  const int feats[4] = {64, 128, 256, 512};
  const int levels = 4, H = 224, W = 224, cap = 8, frames_n = 5;
  const float mean3[3] = {123.675f, 116.28f, 103.53f}, std3[3] = {58.395f, 57.12f, 57.375f}; /* README.md:3110-3111 */
  unet_b200_plan* plan = NULL;
  unet_b200_plan* bad = NULL;
  const int bad_feats[2] = {64, 0};
  int i, j;

  printf("libunet_b200 version %d\n", unet_b200_version());
  CHECK_UB(unet_b200_plan_create(&plan, cap, H, W, 3, 1, feats, levels));
  printf("plan: %d convs, %d layers, %d kernels per pass, workspace %.1f MB (%.1f MB unshared), weights %.1f MB\n",
         unet_b200_plan_num_convs(plan), unet_b200_plan_num_layers(plan), unet_b200_forward_launches(plan),
         unet_b200_plan_workspace_bytes(plan) / 1e6, unet_b200_plan_workspace_unshared_bytes(plan) / 1e6,
         unet_b200_plan_weight_bytes(plan) / 1e6);
  if (unet_b200_plan_num_convs(plan) != 18 || unet_b200_plan_num_layers(plan) != 23) {
    fprintf(stderr, "unexpected plan geometry\n");
    return 1;
  }
  if (unet_b200_plan_create(&bad, cap, H, W, 3, 1, bad_feats, 2) != UB_ERR_ARG) {
    fprintf(stderr, "a zero feature width must be refused\n");
    return 1;
  }
  printf("argument validation: \"%s\"\n", unet_b200_last_error());
  if (unet_b200_forward(plan, (const void*)8, 1, NULL, NULL, NULL, 0.5f, NULL) != UB_ERR_STATE) {
    fprintf(stderr, "forward on an unbound plan must be a state error\n");
    return 1;
  }
  if (unet_b200_device_ok() != UB_OK) {
    printf("no sm_100 device (%s): host-side checks only\n", unet_b200_last_error());
    unet_b200_plan_destroy(plan);
    return 0;
  }

  {
    void *ws = NULL, *wt = NULL, *staging = NULL;
    unsigned char* frames = NULL;
    float *probs = NULL, *probs2 = NULL;
    unsigned char *mask = NULL, *mask2 = NULL;
    const size_t npix = (size_t)frames_n * H * W;
    size_t k, on = 0;
    double psum = 0.0;
    int conv = 0;
    CHECK_CUDA(cudaMalloc(&ws, unet_b200_plan_workspace_bytes(plan)));
    CHECK_CUDA(cudaMalloc(&wt, unet_b200_plan_weight_bytes(plan)));
    CHECK_CUDA(cudaMemset(wt, 0, unet_b200_plan_weight_bytes(plan)));
    CHECK_UB(unet_b200_plan_bind(plan, ws, wt));
    /* 3x3 convs in plan order (include/unet_b200.h): encoder blocks, bottleneck, decoder blocks deepest first */
    for (i = 0; i < 2 * levels + 1; ++i) {
      int cin, cout;
      if (i < levels) {                 /* encoder level i */
        cin = i == 0 ? 3 : feats[i - 1];
        cout = feats[i];
      } else if (i == levels) {         /* bottleneck */
        cin = feats[levels - 1];
        cout = 2 * feats[levels - 1];
      } else {                          /* decoder level j (0 = deepest): conv over cat(skip, up) */
        j = i - levels - 1;
        cout = feats[levels - 1 - j];
        cin = 2 * cout;
      }
      for (j = 0; j < 2; ++j) {
        const int ci = j == 0 ? cin : cout;
        float* w = random_device((size_t)cout * ci * 9, sqrtf(6.0f / (9.0f * ci)));   /* He-uniform: activations keep their scale */
        float* gamma = const_device(cout, 1.0f);
        float* beta = random_device(cout, 0.1f);
        float* mu = random_device(cout, 0.1f);
        float* var = const_device(cout, 1.0f);
        if (!w || !gamma || !beta || !mu || !var) return 1;
        CHECK_UB(unet_b200_plan_set_conv(plan, conv++, w, gamma, beta, mu, var, 1e-5f, NULL));
        CHECK_CUDA(cudaDeviceSynchronize());
        cudaFree(w); cudaFree(gamma); cudaFree(beta); cudaFree(mu); cudaFree(var);
      }
    }
    for (j = 0; j < levels; ++j) {      /* ConvTranspose2d(2f, f, 2, 2), deepest first */
      const int f = feats[levels - 1 - j];
      float* w = random_device((size_t)2 * f * f * 4, sqrtf(3.0f / (2.0f * f)));
      float* b = random_device(f, 0.05f);
      if (!w || !b) return 1;
      CHECK_UB(unet_b200_plan_set_convT(plan, j, w, b, NULL));
      CHECK_CUDA(cudaDeviceSynchronize());
      cudaFree(w); cudaFree(b);
    }
    {
      float* w = random_device(feats[0], 1.0f);
      float* b = const_device(1, -0.25f);
      if (!w || !b) return 1;
      CHECK_UB(unet_b200_plan_set_head(plan, w, b, NULL));
      cudaFree(w); cudaFree(b);
    }
    CHECK_CUDA(cudaMalloc(&staging, unet_b200_infer_stream_staging_bytes(plan, H, W)));
    CHECK_CUDA(cudaMallocHost((void**)&frames, npix * 3));
    CHECK_CUDA(cudaMallocHost((void**)&probs, npix * sizeof(float)));
    CHECK_CUDA(cudaMallocHost((void**)&probs2, npix * sizeof(float)));
    CHECK_CUDA(cudaMallocHost((void**)&mask, npix));
    CHECK_CUDA(cudaMallocHost((void**)&mask2, npix));
    for (k = 0; k < npix * 3; ++k) frames[k] = (unsigned char)((k * 2654435761ULL >> 13) & 255);
    CHECK_UB(unet_b200_infer_u8_host_stream(plan, staging, frames, frames_n, H, W, 1, mean3, std3, 0.5f, NULL, probs, mask, NULL));
    CHECK_UB(unet_b200_infer_u8_host_stream(plan, staging, frames, frames_n, H, W, 1, mean3, std3, 0.5f, NULL, probs2, mask2, NULL));
    for (k = 0; k < npix; ++k) {
      const unsigned char want = probs[k] > 0.5f ? 255 : 0;       /* strict '>' (src/unet.py:67) */
      if (!(probs[k] >= 0.0f && probs[k] <= 1.0f) || mask[k] != want) {
        fprintf(stderr, "pixel %zu: prob %g mask %d\n", k, probs[k], mask[k]);
        return 1;
      }
      psum += probs[k];
      on += mask[k] != 0;
    }
    if (memcmp(probs, probs2, npix * sizeof(float)) != 0 || memcmp(mask, mask2, npix) != 0) {
      fprintf(stderr, "two runs of the same frames differ\n");
      return 1;
    }
    printf("%d frames (%d kernels): mean probability %.4f, %.1f %% of the pixels above 0.5, mask == (prob > 0.5) * 255 on all "
           "%zu pixels, second run bit-identical\n", frames_n, unet_b200_infer_stream_launches(plan, frames_n, H, W),
           psum / (double)npix, 100.0 * (double)on / (double)npix, npix);
    if (on == 0 || on == npix) printf("(note: the mask is constant for these random weights)\n");
    cudaFreeHost(frames); cudaFreeHost(probs); cudaFreeHost(probs2); cudaFreeHost(mask); cudaFreeHost(mask2);
    cudaFree(staging); cudaFree(ws); cudaFree(wt);
  }
  unet_b200_plan_destroy(plan);
  printf("OK\n");
  return 0;
}
