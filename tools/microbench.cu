// Developer microbenchmark: issue throughput (cycles per warp-instruction per SM sub-partition) of the
// instructions the softmax warps are made of, on the real part. nvcc -arch=sm_100a -O3 -o mb microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#define ITERS 4096
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[8]; uint64_t p[8]; uint32_t h[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(a[i] + 1.f)); h[i] = 0; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.5f));
      if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(p[(i + 1) & 7]), "l"(p[(i + 2) & 7]));
      if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(p[(i + 1) & 7]));
      if (OP == 4) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) & 7])); asm volatile("" :: "r"(h[i])); }
      if (OP == 9) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) & 7])); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 4) & 7])); }
      if (OP == 10) asm volatile("shl.b32 %0, %0, 3;" : "+r"(h[i]));
      if (OP == 11) asm volatile("mad.lo.s32 %0, %0, 8388608, %1;" : "+r"(h[i]) : "r"(h[(i + 1) & 7]));
      if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      if (OP == 6) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(1.5f));
      if (OP == 7) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[(i + 4) & 7]) : "f"(1.0001f), "f"(0.5f)); }
      if (OP == 8) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += a[i] + x + y + __uint_as_float(h[i]); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int threads : {128, 256, 512}) {
    k<OP><<<148, threads>>>(out, cyc, 0.5f); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    int warps_per_smsp = threads / 128;
    double n_inst = double(ITERS) * 8 * ((OP == 7 || OP == 9) ? 2 : 1);
    printf("%-28s warps/SMSP=%d  cycles/warp-inst (per warp)=%.2f  per SMSP=%.2f\n", name, warps_per_smsp, avg / n_inst, avg / n_inst / warps_per_smsp);
  }
}
int main() {
  run<0>("MUFU.EX2"); run<1>("FFMA"); run<2>("FFMA2 (f32x2)"); run<3>("FADD2 (f32x2)"); run<4>("F2FP.F16.F32.PACK");
  run<9>("F2FP + MUFU.EX2 interleaved"); run<10>("SHL"); run<11>("IMAD (x*2^23+y)"); run<5>("FMNMX"); run<8>("FMNMX3"); run<6>("FADD"); run<7>("MUFU.EX2 + FFMA interleaved");
  return 0;
}
