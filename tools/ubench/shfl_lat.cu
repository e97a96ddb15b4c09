// micro-benchmark: latency of the loop-carried chain of the Gauss-Seidel strip step (one warp)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, long long* cyc, int n, float a, float b, float c, float d, float r) {
  float W = threadIdx.x * 0.001f, x = 0.5f;
  long long t0, t1;
  // 1: shfl.up + fadd
  t0 = clock64();
  for (int i = 0; i < n; i++) { float S = __shfl_up_sync(0xffffffffu, W, 1); W = S + a; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // 2: fadd chain only (x4)
  t0 = clock64();
  for (int i = 0; i < n; i++) { W = W + a; W = W * b; W = W + c; W = W * d; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // 3: full step: W*cxW + E*cx + S*cy + N*cz - r) * w, S by shuffle, select on lane 0
  float cxW = a;
  t0 = clock64();
  for (int i = 0; i < n; i++) {
    float S = __shfl_up_sync(0xffffffffu, W, 1);
    if (threadIdx.x == 0) S = x;
    float res = (W * cxW + x * b + S * c + x * d - r) * a;
    W = res;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // 4: same with the shuffle source = previous lane via shfl.idx
  t0 = clock64();
  for (int i = 0; i < n; i++) {
    float S = __shfl_sync(0xffffffffu, W, (threadIdx.x + 31) & 31);
    float res = (W * cxW + x * b + S * c + x * d - r) * a;
    W = res;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // 5: through shared memory: st.shared / ld.shared of the neighbour's value
  __shared__ float sh[2][64];
  t0 = clock64();
  for (int i = 0; i < n; i++) {
    sh[i & 1][threadIdx.x + 1] = W;
    __syncwarp();
    float S = sh[i & 1][threadIdx.x];
    float res = (W * cxW + x * b + S * c + x * d - r) * a;
    W = res;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  out[threadIdx.x] = W;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 128); cudaMalloc(&cyc, 64);
  const int n = 4096;
  for (int rep = 0; rep < 2; rep++) k<<<1, 32>>>(out, cyc, n, 0.9f, 0.1f, 0.2f, 0.3f, 0.01f);
  long long h[5];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("shfl.up+fadd %.1f cyc/iter | 4 dependent fp ops %.1f | full step (shfl.up) %.1f | full step (shfl.idx) %.1f | full step (smem) %.1f\n",
         h[0] / (double)n, h[1] / (double)n, h[2] / (double)n, h[3] / (double)n, h[4] / (double)n);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
