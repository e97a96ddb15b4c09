// standalone check of the bulk-copy + mbarrier loader pattern used by smooth_rows.cuh
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>
namespace rlfc { struct RowTab { const float4* T; int C, K, entries; }; struct SkewLevel { const float4* A; const float2* nd; int nstrips, Tsk; };
struct DevLevel { SkewLevel sk; RowTab rt; int n, m, P; size_t stride; const float *lx, *ly, *inv, *diag; float *r, *r2, *x, *d; }; }
#define RLFC_NO_SOLVER_H
#include "../../rlfluidcontrol_b200/csrc/smooth_rows.cuh"
using namespace rlfc;
using namespace rlfc::rows_detail;
__global__ void k(const float* src, float* out, int n_iter) {
  extern __shared__ __align__(16) unsigned char sm[];
  float* buf = (float*)sm;                       // 16 slots x 1024 floats
  unsigned long long* bars = (unsigned long long*)(sm + 16 * 4096);
  if (threadIdx.x == 0) { for (int k = 0; k < 16; k++) mbar_init(bars + k, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto issue = [&](int q) {
    unsigned long long* bar = bars + ((q - 1) & 15);
    mbar_expect_tx(bar, 4096);
    bulk_g2s(buf + (q & 15) * 1024, src + (size_t)q * 1024, 3584, bar);
    bulk_g2s(buf + (q & 15) * 1024 + 896, src + (size_t)q * 1024 + 896, 512, bar);
  };
  if (warp == 1) { if (lane == 0) { fence_proxy_async(); for (int q = 1; q <= 5; q++) issue(q); } mbar_wait(bars + 0, 0); }
  __syncthreads();
  float acc = 0;
  for (int t = 1; t <= n_iter; t++) {
    if (warp == 1) {
      if (lane == 0) { fence_proxy_async(); issue(t + 5); }
      mbar_wait(bars + (t & 15), ((unsigned)t >> 4) & 1u);
    } else {
      acc += buf[(t & 15) * 1024 + threadIdx.x];
    }
    __syncthreads();
  }
  if (warp == 1) for (int q = n_iter + 2; q <= n_iter + 5; q++) mbar_wait(bars + ((q - 1) & 15), ((unsigned)(q - 1) >> 4) & 1u);
  __syncthreads();
  out[threadIdx.x] = acc;
}
int main() {
  const int n_iter = 100;
  float *src, *out;
  cudaMalloc(&src, (n_iter + 8) * 4096); cudaMalloc(&out, 64 * 4);
  float* h = new float[(n_iter + 8) * 1024];
  for (int i = 0; i < (n_iter + 8) * 1024; i++) h[i] = (float)(i / 1024);
  cudaMemcpy(src, h, (n_iter + 8) * 4096, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 4096 + 256);
  k<<<1, 64, 16 * 4096 + 256>>>(src, out, n_iter);
  cudaError_t e = cudaDeviceSynchronize();
  float ho[64]; cudaMemcpy(ho, out, 256, cudaMemcpyDeviceToHost);
  printf("err=%s out[0]=%f expect %f\n", cudaGetErrorString(e), ho[0], (float)(n_iter * (n_iter + 1) / 2));
  return 0;
}
