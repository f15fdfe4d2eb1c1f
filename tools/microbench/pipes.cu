// Microbenchmark 2: per-pipe issue rates on sm_100a that the blend kernel's instruction budget is
// planned against: packed FMUL2 / FADD2 / FFMA2, scalar FMUL, the ALU-pipe ops the kernel uses
// (FSETP+SEL, FMNMX, LOP3/IADD) and FMA+ALU co-issue.  Prints warp-instructions per clock per
// SM sub-partition (SMSP).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define DEVINL __device__ __forceinline__
DEVINL uint64_t pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
DEVINL void unpack(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
DEVINL uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
DEVINL uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
DEVINL uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int ITERS = 2048, CH = 8;
struct Params { uint64_t negzero2; float m, a; int one; };

template <int MODE>
__global__ void __launch_bounds__(256) bench(float *out, const __grid_constant__ Params P) {
  float x[2 * CH];
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) x[i] = 1.0f + (float)(threadIdx.x + i) * 1e-3f;
  uint64_t v[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) v[i] = pack(x[2 * i], x[2 * i + 1]);
  const uint64_t mm = pack(P.m, P.m), aa = pack(P.a, P.a);
  uint32_t u[2 * CH];
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) u[i] = threadIdx.x * 7 + i;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) v[i] = mul2(v[i], mm);                                   // 8 FMUL2
      if (MODE == 1) v[i] = add2(v[i], aa);                                   // 8 FADD2
      if (MODE == 2) v[i] = fma2(v[i], mm, aa);                               // 8 FFMA2
      if (MODE == 3) v[i] = add2(fma2(v[i], mm, P.negzero2), aa);             // 8 x (unfusable mul) + 8 FADD2
      if (MODE == 4) { x[2 * i] = __fmul_rn(x[2 * i], P.m); x[2 * i + 1] = __fmul_rn(x[2 * i + 1], P.m); }   // 16 FMUL
      if (MODE == 5) { x[2 * i] = fminf(x[2 * i], P.m + x[2 * i + 1]); x[2 * i + 1] = fmaxf(x[2 * i + 1], x[2 * i]); }  // FMNMX-ish
      if (MODE == 6) { u[2 * i] = (u[2 * i] ^ u[2 * i + 1]) + 0x9E3779B9u; u[2 * i + 1] = (u[2 * i + 1] & u[2 * i]) | 0x55u; } // LOP3/IADD
      if (MODE == 7) { v[i] = fma2(v[i], mm, aa); u[2 * i] = (u[2 * i] ^ u[2 * i + 1]) + 0x9E3779B9u; u[2 * i + 1] = (u[2 * i + 1] & u[2 * i]) | 0x55u; }
      if (MODE == 8) { v[i] = fma2(v[i], mm, aa); x[2 * i] = fminf(x[2 * i], P.a); x[2 * i + 1] = fmaxf(x[2 * i + 1], x[2 * i]); }
      if (MODE == 9) {   // FSETP + SEL pairs
        x[2 * i] = (x[2 * i] > x[2 * i + 1]) ? P.m : x[2 * i];
        x[2 * i + 1] = (x[2 * i + 1] >= x[2 * i]) ? P.a : x[2 * i + 1];
      }
      if (MODE == 10) { v[i] = fma2(v[i], mm, aa); x[2 * i] = __fmul_rn(x[2 * i], P.m); x[2 * i + 1] = __fmul_rn(x[2 * i + 1], P.m); }  // FFMA2 + 2 FMUL
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) { float l, h; unpack(v[i], l, h); s += l + h + x[2 * i] + x[2 * i + 1] + (float)(u[2 * i] ^ u[2 * i + 1]); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, float *d, double instr_per_iter) {
  const int blocks = 148 * 8, threads = 256;   // 8 CTAs x 8 warps per SM = 16 warps per SMSP
  Params P{0x8000000080000000ull, 0.999f, 1e-4f, 1};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<MODE><<<blocks, threads>>>(d, P);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) bench<MODE><<<blocks, threads>>>(d, P);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double warp_instr = (double)blocks * threads / 32 * ITERS * instr_per_iter;
  const double smsp_clk = (double)ms * 1e-3 * 1.965e9 * 148 * 4;   // assumes 1965 MHz
  printf("%-46s %8.3f ms  %6.3f warp-instr/clk/SMSP (listed instrs only)\n", name, ms, warp_instr / smsp_clk);
}

int main() {
  float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  run<0>("FMUL2 x8", d, 8);
  run<1>("FADD2 x8", d, 8);
  run<2>("FFMA2 x8", d, 8);
  run<3>("FFMA2(+runtime -0) x8 + FADD2 x8", d, 16);
  run<4>("FMUL x16", d, 16);
  run<5>("FADD+FMNMX, FMNMX x8 (24 instr)", d, 24);
  run<6>("LOP3/IADD x~32", d, 32);
  run<7>("FFMA2 x8 + LOP3/IADD x~32", d, 40);
  run<8>("FFMA2 x8 + FMNMX x16", d, 24);
  run<9>("FSETP+SEL x16 (32 instr)", d, 32);
  run<10>("FFMA2 x8 + FMUL x16", d, 24);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
