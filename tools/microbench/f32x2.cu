// Microbenchmark: issue/pipe throughput of scalar vs packed (f32x2) FP32 ops on sm_100a, and a
// bit-exactness check of the packed ops (rn/rz) against the scalar ones.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define DEVINL __device__ __forceinline__
DEVINL uint64_t pack(float lo, float hi) {
  uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
DEVINL void unpack(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
DEVINL uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
DEVINL uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
DEVINL uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
DEVINL uint64_t add2rz(uint64_t a, uint64_t b) {
  uint64_t d; asm volatile("add.rz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
DEVINL uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

constexpr int ITERS = 4096;
constexpr int CH = 8;   // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(256) bench(float *out, float seed) {
  float x[CH * 2];
#pragma unroll
  for (int i = 0; i < CH * 2; ++i) x[i] = seed + (float)(threadIdx.x + i) * 1e-3f;
  const float m = 0.999f, a = 1e-4f;
  if (MODE == 0) {          // scalar FFMA, 2*CH chains (same flop count as packed with CH chains)
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int i = 0; i < CH * 2; ++i) x[i] = __fmaf_rn(x[i], m, a);
  } else if (MODE == 1) {   // packed FFMA2, CH chains
    uint64_t v[CH]; const uint64_t mm = pack(m, m), aa = pack(a, a);
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = pack(x[2 * i], x[2 * i + 1]);
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int i = 0; i < CH; ++i) v[i] = fma2(v[i], mm, aa);
#pragma unroll
    for (int i = 0; i < CH; ++i) unpack(v[i], x[2 * i], x[2 * i + 1]);
  } else if (MODE == 2) {   // scalar FMUL + FADD (no fma), 2*CH chains
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int i = 0; i < CH * 2; ++i) x[i] = __fadd_rn(__fmul_rn(x[i], m), a);
  } else if (MODE == 3) {   // packed FMUL2 + FADD2
    uint64_t v[CH]; const uint64_t mm = pack(m, m), aa = pack(a, a);
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = pack(x[2 * i], x[2 * i + 1]);
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int i = 0; i < CH; ++i) v[i] = add2(mul2(v[i], mm), aa);
#pragma unroll
    for (int i = 0; i < CH; ++i) unpack(v[i], x[2 * i], x[2 * i + 1]);
  } else if (MODE == 4) {   // packed FMUL2 + FADD2.RZ mixed with ALU work (FMNMX) to see co-issue
    uint64_t v[CH]; const uint64_t mm = pack(m, m), aa = pack(a, a);
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = pack(x[2 * i], x[2 * i + 1]);
    float y[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) y[i] = x[i];
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int i = 0; i < CH; ++i) { v[i] = add2rz(mul2(v[i], mm), aa); y[i] = fminf(fmaxf(y[i], a), m); y[i] = fmaxf(fminf(y[i], x[i]), a); }
#pragma unroll
    for (int i = 0; i < CH; ++i) { unpack(v[i], x[2 * i], x[2 * i + 1]); x[i] += y[i]; }
  } else if (MODE == 5) {   // scalar equivalent of mode 4
    float y[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) y[i] = x[i];
    float x0[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x0[i] = x[i];
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < CH * 2; ++i) x[i] = __fadd_rz(__fmul_rn(x[i], m), a);
#pragma unroll
      for (int i = 0; i < CH; ++i) { y[i] = fminf(fmaxf(y[i], a), m); y[i] = fmaxf(fminf(y[i], x0[i]), a); }
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] += y[i];
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH * 2; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// exactness: packed result bits == scalar result bits over pseudo-random operands
__global__ void exact(uint32_t *bad, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = i * 2654435761u + 12345u;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return __uint_as_float((s & 0x3FFFFFFFu) | 0x30000000u) - 1.0f; };
  float a0 = rnd(), a1 = rnd(), b0 = rnd(), b1 = rnd(), c0 = rnd() * 8388608.f, c1 = rnd();
  float l, h; uint32_t nb = 0;
  unpack(fma2(pack(a0, a1), pack(b0, b1), pack(c0, c1)), l, h);
  nb += (__float_as_uint(l) != __float_as_uint(__fmaf_rn(a0, b0, c0))) + (__float_as_uint(h) != __float_as_uint(__fmaf_rn(a1, b1, c1)));
  unpack(mul2(pack(a0, a1), pack(b0, b1)), l, h);
  nb += (__float_as_uint(l) != __float_as_uint(__fmul_rn(a0, b0))) + (__float_as_uint(h) != __float_as_uint(__fmul_rn(a1, b1)));
  unpack(add2(pack(a0, a1), pack(c0, c1)), l, h);
  nb += (__float_as_uint(l) != __float_as_uint(__fadd_rn(a0, c0))) + (__float_as_uint(h) != __float_as_uint(__fadd_rn(a1, c1)));
  unpack(add2rz(pack(a0 * 255.f, a1 * 255.f), pack(8388608.f, 8388608.f)), l, h);
  nb += (__float_as_uint(l) != __float_as_uint(__fadd_rz(a0 * 255.f, 8388608.f))) + (__float_as_uint(h) != __float_as_uint(__fadd_rz(a1 * 255.f, 8388608.f)));
  unpack(sub2(pack(a0, a1), pack(c0, c1)), l, h);
  nb += (__float_as_uint(l) != __float_as_uint(__fsub_rn(a0, c0))) + (__float_as_uint(h) != __float_as_uint(__fsub_rn(a1, c1)));
  if (nb) atomicAdd(bad, nb);
}

template <int MODE>
void run(const char *name, float *d, double flop_per_iter_thread) {
  const int blocks = 148 * 8, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<MODE><<<blocks, threads>>>(d, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) bench<MODE><<<blocks, threads>>>(d, 1.0f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double ops = (double)blocks * threads * ITERS * flop_per_iter_thread;
  printf("%-42s %8.3f ms  %8.2f T lane-ops/s\n", name, ms, ops / ms / 1e9);
}

int main() {
  float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  run<0>("scalar FFMA x16 chains", d, 16);
  run<1>("packed FFMA2 x8 chains (16 fma)", d, 16);
  run<2>("scalar FMUL+FADD x16", d, 32);
  run<3>("packed FMUL2+FADD2 x8", d, 32);
  run<4>("packed MUL2+ADD2.RZ x8 + 4 FMNMX x8", d, 32 + 32);
  run<5>("scalar MUL+ADD.RZ x16 + 4 FMNMX x8", d, 32 + 32);
  uint32_t *bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
  exact<<<4096, 256>>>(bad, 4096 * 256);
  uint32_t hb = 1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  printf("packed-vs-scalar bit mismatches: %u (err %s)\n", hb, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
