// Issue-rate probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), alone and interleaved with ALU work.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ uint32_t alu1(uint32_t a, uint32_t b) { uint32_t d; asm volatile("lop3.b32 %0, %1, %2, %1, 0x96;" : "=r"(d) : "r"(a), "r"(b)); return d; }
template <int MODE> __global__ void probe(float* out, int iters, float s) {
  float a[8]; uint64_t p[8]; uint32_t u[8];
  for (int i = 0; i < 8; i++) { a[i] = s + i; p[i] = pk(s + i, s - i); u[i] = threadIdx.x + i; }
  uint64_t ps = pk(s, s);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MODE == 0) a[i] = fma1(a[i], s, s);                       // 64 FFMA
        if (MODE == 1) p[i] = fma2(p[i], ps, ps);                     // 64 FFMA2 (128 fp32 fma)
        if (MODE == 2) { a[i] = fma1(a[i], s, s); u[i] = alu1(u[i], u[(i + 1) & 7]); }   // 64 FFMA + 64 LOP3
        if (MODE == 3) { p[i] = fma2(p[i], ps, ps); u[i] = alu1(u[i], u[(i + 1) & 7]); } // 64 FFMA2 + 64 LOP3
      }
    }
  }
  float acc = 0; for (int i = 0; i < 8; i++) { acc += a[i] + (float)(p[i] & 0xff) + u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, int per_iter_issue) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<148 * 8, 256>>>(out, 16, 1.0001f);
  cudaEventRecord(e0); probe<MODE><<<148 * 8, 256>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_inst = 148.0 * 8 * 8 * iters * per_iter_issue;
  printf("%-18s %.3f ms  %.1f G warp-inst/s  (%.2f per SMSP-clk at 1.965 GHz)\n", name, ms, warp_inst / ms / 1e6, warp_inst / ms / 1e6 / (148 * 4 * 1.965));
  cudaFree(out);
}
int main() { run<0>("FFMA", 64); run<1>("FFMA2", 64); run<2>("FFMA+LOP3", 128); run<3>("FFMA2+LOP3", 128); return 0; }
