// Pipe-throughput microbenchmark for the softmax inner loop: MUFU.EX2, F2FP (bf16x2 pack), FFMA2, FMNMX3 and
// mixes, at 8 warps per SM (2 per sub-partition) like the attention kernel.  Prints cycles per warp-instruction
// per sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = -0.001f * (threadIdx.x + i);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {  // ex2 only
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 1) {  // pack only
        uint32_t u;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a[i]), "f"(a[(i + 1) & 15]));
        acc ^= u;
      } else if (MODE == 2) {  // ex2 + pack (1 pack per 2 ex2)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (i & 1) {
          uint32_t u;
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a[i]), "f"(a[i - 1]));
          acc ^= u;
        }
      } else if (MODE == 3) {  // fma only
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(-0.001f));
      } else if (MODE == 4) {  // ex2 + 4 fma (poly-ish)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[(i + 8) & 15]) : "f"(0.999f), "f"(-0.001f));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[(i + 9) & 15]) : "f"(0.999f), "f"(-0.001f));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<MODE><<<148, 256>>>(out, cyc, iters);
  k<MODE><<<148, 256>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  // 2 warps per sub-partition
  printf("%-28s %8.2f cycles per warp-instruction per sub-partition (%d instr/iter/warp)\n", name,
         avg / (double(iters) * instr_per_iter * 2), instr_per_iter);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("MUFU.EX2", 16);
  run<1>("F2FP.BF16 pack", 16);
  run<2>("EX2 + 0.5 F2FP (per ex2)", 16);
  run<3>("FFMA", 16);
  run<4>("EX2 + 2 FFMA (per ex2)", 16);
  return 0;
}
