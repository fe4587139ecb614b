// Pipe-throughput micro-benchmark for the softmax inner loop: how many warp-instructions per cycle per
// SM for MUFU.EX2 (f32 / f16x2), FFMA, FMNMX, cvt.f16x2, HFMA2, and a polynomial exp2 on the FMA pipe.
#include <cuda_fp16.h>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

template <int MODE>
__global__ void k(float* out, float seed) {
  float a[UNROLL];
  unsigned h[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) { a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; h[i] = 0x3c003c00u + i; }
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(seed), "f"(a[(i + 1) % UNROLL]));
      if (MODE == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) % UNROLL]));
      if (MODE == 4) {   // keep the conversion live: its result feeds the next iteration's input
        unsigned t;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(a[i]), "f"(a[(i + 1) % UNROLL]));
        a[i] = __uint_as_float(t | 0x3f000000u);
      }
      if (MODE == 10) asm volatile("ex2.approx.f32 %0, %0;" : "+f"(a[i]));          // non-ftz MUFU
      if (MODE == 11) {  // integer fp32->fp16 pack (round-half-up) : 2 IADD + 2 SHF + 1 LOP3 per pair
        unsigned x0 = __float_as_uint(a[i]) + 0x1000u, x1 = __float_as_uint(a[(i + 1) % UNROLL]) + 0x1000u;
        unsigned t = ((x1 << 3) & 0xFFFF0000u) | (x0 >> 13);
        a[i] = __uint_as_float(t | 0x3f000000u);
      }
      if (MODE == 12) {  // scalar cvt.rn.f16.f32 (F2F)
        unsigned short t;
        asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(t) : "f"(a[i]));
        a[i] = __uint_as_float((unsigned)t | 0x3f000000u);
      }
      if (MODE == 5) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(h[(i + 1) % UNROLL]), "r"(h[(i + 2) % UNROLL]));
      if (MODE == 6) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(h[(i + 1) % UNROLL]));
      if (MODE == 7) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) % UNROLL]));
      if (MODE == 8) {   // exp2 emulation: round-to-int via magic add, degree-3 polynomial, exponent insert
        float x = a[i];
        float fl = __fadd_rn(x, 12582912.f);
        float r = x - (fl - 12582912.f);
        float p = fmaf(fmaf(fmaf(0.0555f, r, 0.2402f), r, 0.6931f), r, 1.0f);
        a[i] = __int_as_float(__float_as_int(p) + (__float_as_int(fl) << 23)) * 1e-30f + x * 0.999f;
      }
      if (MODE == 9) {  // mixed: 1 MUFU f32 + 1 FFMA + 1 FMNMX + 1 FADD  (the v1 per-element mix)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(seed), "f"(a[(i + 1) % UNROLL]));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 2) % UNROLL]));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 3) % UNROLL]));
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) s += a[i] + __uint_as_float(h[i]);
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int warps_per_sm, int ops_per_iter) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* d; cudaMalloc(&d, 4);
  dim3 grid(sms), block(32 * warps_per_sm);
  k<MODE><<<grid, block>>>(d, 0.5f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<MODE><<<grid, block>>>(d, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double warp_instr = (double)ITERS * UNROLL * ops_per_iter * warps_per_sm;      // per SM
  double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-28s warps/SM=%2d  ms=%.3f  warp-instr/cycle/SM=%.3f  (cycles per warp-instr per SMSP=%.2f)\n", name,
         warps_per_sm, ms, warp_instr / cycles, 4.0 * cycles / warp_instr);
  cudaFree(d);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("MUFU.EX2 f32", w, 1);
    run<1>("MUFU.EX2 f16x2", w, 1);
    run<2>("FFMA", w, 1);
    run<3>("FMNMX", w, 1);
    run<4>("cvt.rn.f16x2.f32", w, 1);
    run<5>("HFMA2", w, 1);
    run<6>("HMNMX2", w, 1);
    run<7>("FADD", w, 1);
    run<8>("poly exp2 (per exp)", w, 1);
    run<9>("mix ex2+ffma+max+add (x4)", w, 4);
    run<10>("MUFU.EX2 f32 non-ftz", w, 1);
    run<11>("int pack f32x2->f16x2 (5 ALU)", w, 1);
    run<12>("cvt.rn.f16.f32 scalar", w, 1);
  }
  return 0;
}
