// pipe_bench.cu -- issue-rate microbenchmark of the SASS instructions the BVH node test is made of
// (sm_100a).  Each test runs 8 independent dependency chains per thread, 8 warps per SMSP, and reports
// warp-instructions per clock per SMSP (1.0 = one instruction issued every cycle).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_bench pipe_bench.cu && ./pipe_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define CH 8

template <int T>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t r[CH];
  uint32_t s[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) {
    r[i] = seed * (threadIdx.x + 1) + i * 0x01010101u + 0x3f800000u;
    s[i] = r[i] ^ 0x00123400u;
  }
  uint32_t c0 = seed | 0x3f800001u, c1 = seed + 0x3f000000u;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (T == 0) asm volatile("prmt.b32 %0, %1, %0, 0x3240;" : "+r"(r[i]) : "r"(c0));
      if (T == 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c0));
      if (T == 2) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c0), "f"(*(float*)&c1));
      if (T == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c0), "f"(*(float*)&c1));
      if (T == 4) {
        if (i % 2 == 0) {
          asm volatile("{ .reg .b64 a, b, c; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; mov.b64 c, {%3, %2};\n"
                       "fma.rn.f32x2 a, a, b, c; mov.b64 {%0, %1}, a; }"
                       : "+r"(r[i]), "+r"(r[i + 1]) : "r"(c0), "r"(c1));
        }
      }
      if (T == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(c0), "r"(c1));
      if (T == 6) asm volatile("{ .reg .pred p; setp.gt.f32 p, %0, %1; selp.b32 %0, %2, %0, p; }" : "+r"(r[i]) : "f"(*(float*)&c0), "r"(c1));
      if (T == 7) asm volatile("{ .reg .b16 h; mov.b32 {h, _}, %0; cvt.f32.f16 %0, h; }" : "+r"(r[i]));
      if (T == 8) asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(r[i]));
      if (T == 9) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c0), "r"(c1));
      if (T == 10) {  // PRMT + FFMA pairs (different pipes): can they issue back to back?
        asm volatile("prmt.b32 %0, %1, %0, 0x3240;" : "+r"(r[i]) : "r"(c0));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&s[i]) : "f"(*(float*)&c0), "f"(*(float*)&c1));
      }
      if (T == 11) {  // PRMT + FMNMX (same pipe?)
        asm volatile("prmt.b32 %0, %1, %0, 0x3240;" : "+r"(r[i]) : "r"(c0));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(*(float*)&s[i]) : "f"(*(float*)&c0));
      }
      if (T == 12) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c0), "r"(c1));
      if (T == 13) asm volatile("fma.rn.f32 %0, %0, 0f3F800347, %1;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c1));
      if (T == 14) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c1));
      if (T == 15) asm volatile("{ .reg .pred p; setp.gt.f32 p, %0, %1; @p or.b32 %0, %0, %2; }" : "+r"(r[i]) : "f"(*(float*)&c0), "r"(c1));
      if (T == 16) {  // FFMA + FMNMX + PRMT triple
        asm volatile("prmt.b32 %0, %1, %0, 0x3240;" : "+r"(r[i]) : "r"(c0));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&s[i]) : "f"(*(float*)&c0), "f"(*(float*)&c1));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&s[i]) : "f"(*(float*)&c1), "f"(*(float*)&c0));
      }
      if (T >= 20 && T < 30) {
        asm volatile("prmt.b32 %0, %1, %0, 0x3240;" : "+r"(r[i]) : "r"(c0));
        if (T == 20) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(c0), "r"(c1));
        if (T == 21) asm volatile("{ .reg .b16 h; mov.b32 {h, _}, %0; cvt.f32.f16 %0, h; }" : "+r"(s[i]));
        if (T == 22) asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(s[i]));
        if (T == 23) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float*)&s[i]) : "f"(*(float*)&c1));
        if (T == 24) asm volatile("{ .reg .pred p; setp.gt.f32 p, %0, %1; @p bra L%=; L%=: }" :: "f"(*(float*)&s[i]), "f"(*(float*)&c0));
        if (T == 25) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(s[i]) : "r"((s[i] & 0x3fcu)));
        if (T == 26) asm volatile("shl.b32 %0, %0, 3;" : "+r"(s[i]));
        if (T == 27) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(c0), "r"(c1));
        if (T == 28) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(c0));
      }
      if (T >= 30 && T < 40) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&r[i]) : "f"(*(float*)&c0), "f"(*(float*)&c1));
        if (T == 30) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(c0), "r"(c1));
        if (T == 31) asm volatile("{ .reg .b16 h; mov.b32 {h, _}, %0; cvt.f32.f16 %0, h; }" : "+r"(s[i]));
        if (T == 32) asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(s[i]));
        if (T == 33) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(c0), "r"(c1));
      }
      if (T == 40) {  // 1 PRMT + 2 HADD2.F32 + 2 FFMA: the fp16-pair dequantisation pattern
        asm volatile("prmt.b32 %0, %1, %0, 0x3240;" : "+r"(r[i]) : "r"(c0));
        uint32_t a, b;
        asm volatile("{ .reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }" : "=r"(a), "=r"(b) : "r"(r[i]));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(*(float*)&s[i]) : "f"(*(float*)&a), "f"(*(float*)&c1));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(*(float*)&s[i]) : "f"(*(float*)&b), "f"(*(float*)&c0));
      }
      if (T == 41) {  // 2 PRMT + 2 FFMA: today's pattern
        uint32_t a, b;
        asm volatile("prmt.b32 %0, %1, %2, 0x3240;" : "=r"(a) : "r"(c0), "r"(r[i]));
        asm volatile("prmt.b32 %0, %1, %2, 0x3250;" : "=r"(b) : "r"(c0), "r"(r[i]));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(*(float*)&s[i]) : "f"(*(float*)&a), "f"(*(float*)&c1));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(*(float*)&s[i]) : "f"(*(float*)&b), "f"(*(float*)&c0));
        r[i] += s[i];
      }
      if (T == 17) asm volatile("{ .reg .b32 t; add.f16x2 t, %0, %1; mov.b32 %0, t; }" : "+r"(r[i]) : "r"(c0));
      if (T == 18) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(c0), "r"(c1));
      if (T == 19) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(c0));
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) acc ^= r[i] ^ s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int T>
void run(const char* name, double perIter, uint32_t* out, long long* cyc) {
  int nb = 148;
  k<T><<<nb, 1024>>>(out, cyc, 3);
  k<T><<<nb, 1024>>>(out, cyc, 3);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < nb; i++) avg += double(h[i]);
  avg /= nb;
  // per SMSP: 8 warps x ITERS x perIter instructions
  double inst = 8.0 * ITERS * perIter;
  printf("%-34s %6.3f warp-inst/clk/SMSP  (%.0f cycles)\n", name, inst / avg, avg);
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("PRMT", CH, out, cyc);
  run<1>("FMNMX (2-input)", CH, out, cyc);
  run<2>("FMNMX3 (3-input)", CH, out, cyc);
  run<3>("FFMA (3 regs)", CH, out, cyc);
  run<13>("FFMA (immediate)", CH, out, cyc);
  run<4>("FFMA2 (f32x2), per instruction", CH / 2, out, cyc);
  run<14>("FADD", CH, out, cyc);
  run<12>("IMAD", CH, out, cyc);
  run<5>("LOP3", CH, out, cyc);
  run<9>("SHF", CH, out, cyc);
  run<6>("FSETP+SEL (2 inst)", 2 * CH, out, cyc);
  run<15>("FSETP+@p LOP (2 inst)", 2 * CH, out, cyc);
  run<7>("HADD2.F32 (cvt f16->f32)", CH, out, cyc);
  run<8>("I2F(P) u32->f32", CH, out, cyc);
  run<17>("HADD2", CH, out, cyc);
  run<18>("HFMA2", CH, out, cyc);
  run<19>("HMNMX2", CH, out, cyc);
  run<10>("PRMT + FFMA (2 inst)", 2 * CH, out, cyc);
  run<11>("PRMT + FMNMX (2 inst)", 2 * CH, out, cyc);
  run<16>("PRMT + 2 FFMA (3 inst)", 3 * CH, out, cyc);
  run<20>("PRMT + IMAD (2 inst)", 2 * CH, out, cyc);
  run<21>("PRMT + HADD2.F32 (2 inst)", 2 * CH, out, cyc);
  run<22>("PRMT + I2FP (2 inst)", 2 * CH, out, cyc);
  run<23>("PRMT + FADD (2 inst)", 2 * CH, out, cyc);
  run<25>("PRMT + LDS (2 inst)", 2 * CH, out, cyc);
  run<26>("PRMT + SHL (2 inst)", 2 * CH, out, cyc);
  run<27>("PRMT + HFMA2 (2 inst)", 2 * CH, out, cyc);
  run<28>("PRMT + IMAD.HI (2 inst)", 2 * CH, out, cyc);
  run<30>("FFMA + IMAD (2 inst)", 2 * CH, out, cyc);
  run<31>("FFMA + HADD2.F32 (2 inst)", 2 * CH, out, cyc);
  run<32>("FFMA + I2FP (2 inst)", 2 * CH, out, cyc);
  run<33>("FFMA + HFMA2 (2 inst)", 2 * CH, out, cyc);
  run<40>("PRMT + 2 HADD2.F32 + 2 FFMA (5)", 5 * CH, out, cyc);
  run<41>("2 PRMT + 2 FFMA (4 inst + add)", 5 * CH, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
