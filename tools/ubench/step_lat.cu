// micro-benchmark: what a Gauss-Seidel strip step costs as operands, shuffles and stores are added (one warp, one SM)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 lds128v(unsigned a) {
  float4 v;
  asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds128(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void st64v(uint2* p, unsigned v, unsigned tag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v), "r"(tag) : "memory");
}
constexpr int NB = 64;   // batches of 8 steps
template <int MODE>
__global__ void k(float* outf, uint2* outg, long long* cyc) {
  __shared__ __align__(16) float4 coef[16][32];
  __shared__ float rr[16][32], ee[16][32], ed[16];
  const int lane = threadIdx.x;
  for (int k = 0; k < 16; k++) { coef[k][lane] = make_float4(0.1f, 0.2f, 0.3f, -0.9f); rr[k][lane] = 0.01f * lane; ee[k][lane] = 0.5f; ed[k] = 0.25f; }
  __syncwarp();
  const unsigned a_c = (unsigned)__cvta_generic_to_shared(&coef[0][lane]);
  const unsigned a_r = (unsigned)__cvta_generic_to_shared(&rr[0][lane]);
  const unsigned a_e = (unsigned)__cvta_generic_to_shared(&ee[0][lane]);
  const unsigned a_x = (unsigned)__cvta_generic_to_shared(&ed[0]);
  const bool l0 = lane == 0, l31 = lane == 31;
  float W = lane * 0.001f, cxW = 0.1f;
  uint2* out = outg + lane;
  long long t0 = clock64();
#pragma unroll 1
  for (int b = 0; b < NB; b++) {
    const unsigned off = (b & 1) * 8;
    float4 cn = (MODE & 1) ? lds128v(a_c + 512u * off) : lds128(a_c + 512u * off);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float4 c = cn;
      if (k + 1 < 8) cn = (MODE & 1) ? lds128v(a_c + 512u * (off + k + 1)) : lds128(a_c + 512u * (off + k + 1));
      const float rv = lds32(a_r + 128u * (off + k));
      const float E = lds32(a_e + 128u * (off + k));
      const float axv = lds32(a_x + 4u * (off + k));
      float S = __shfl_up_sync(0xffffffffu, W, 1);
      float N = (MODE & 2) ? __shfl_down_sync(0xffffffffu, E, 1) : E;
      if (l31) N = axv;
      if (l0) S = axv;
      const float res = (W * cxW + E * c.x + S * c.y + N * c.z - rv) * c.w;
      if (MODE & 4) {
        if (l0 || l31) st64v(out + 4096 + k, __float_as_uint(res), 7u);
        st64v(out + k * 32, __float_as_uint(res), 7u);
      }
      W = res;
      cxW = c.x;
    }
    if (MODE & 4) out += 8 * 32;
  }
  long long t1 = clock64();
  if (lane == 0) cyc[0] = t1 - t0;
  outf[lane] = W;
}
template <int MODE> void run(const char* name, float* outf, uint2* outg, long long* cyc) {
  for (int rep = 0; rep < 2; rep++) k<MODE><<<1, 32>>>(outf, outg, cyc);
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-40s %.1f cyc/step\n", name, h / (double)(NB * 8));
}
int main() {
  float* outf; uint2* outg; long long* cyc;
  cudaMalloc(&outf, 128); cudaMalloc(&outg, 64 << 20); cudaMalloc(&cyc, 64);
  run<0>("operands from smem", outf, outg, cyc);
  run<1>("+ volatile coefficient loads", outf, outg, cyc);
  run<2>("+ shfl.down of E", outf, outg, cyc);
  run<3>("volatile + shfl.down", outf, outg, cyc);
  run<4>("stores (volatile), no shfl.down", outf, outg, cyc);
  run<7>("everything", outf, outg, cyc);
  run<6>("stores + shfl.down, plain coef loads", outf, outg, cyc);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
