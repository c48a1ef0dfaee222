// TEST INFRASTRUCTURE.  Checks the division used by csrc/actquant.cu (div_exact): with r = RN(1/a),
// q0 = RN(t*r), e = fma(-a, q0, t), q = fma(e, r, q0) equals IEEE t / a for the operand classes the
// Sym/Asym activation quantizers produce: t = integer codes (|t| <= 32767) or dtype-valued
// differences with |t| >= 1e-30 or t == 0, a = any normal float in [1e-30, 1e30] (fp32 / bf16 /
// fp16 valued).  Everything else (tiny t: the residual underflows; non-finite quotients) takes the
// IEEE fallback in the kernel and is skipped here.
//   gcc -O2 -fopenmp -o /tmp/div_check_sym oracle/div_check_sym.c -lm && /tmp/div_check_sym
#include <math.h>
#include <stdlib.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static inline uint64_t rng(uint64_t* s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }
static inline float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t asu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
int main(int argc, char** argv) {
  long n = argc > 1 ? atol(argv[1]) : 100000000L;
  long bad = 0, total = 0;
#pragma omp parallel reduction(+ : bad, total)
  {
    uint64_t s = 0x9E3779B97F4A7C15ULL + 104729ULL * omp_get_thread_num();
    for (long i = 0; i < n; i++) {
      uint64_t r1 = rng(&s), r2 = rng(&s);
      int e = 127 - 99 + (int)((r1 >> 40) % 199);              // 2^-99 .. 2^99
      uint32_t ma = (uint32_t)r1 & 0x7FFFFF;
      int mode = (r2 >> 60) & 3;
      if (mode == 1) ma &= 0x7F0000;
      if (mode == 2) ma &= 0x7FE000;
      float a = asf(((uint32_t)e << 23) | ma);
      if (a < 1e-30f || a > 1e30f) continue;
      float t;
      if ((r2 >> 58) & 1) {
        t = (float)((int)((r2 >> 8) % 65535) - 32767);         // integer code
      } else {
        uint32_t mt = (uint32_t)r2 & 0x7FFFFF;
        if (mode == 1) mt &= 0x7F0000;
        if (mode == 2) mt &= 0x7FE000;
        int et = e - 30 + (int)((r2 >> 32) % 60);
        if (et < 1) et = 1;
        if (et > 254) et = 254;
        t = asf(((uint32_t)et << 23) | mt | ((uint32_t)(r2 >> 57) << 31));
      }
      if (t != 0.0f && fabsf(t) < 1e-30f) continue;            // kernel: IEEE fallback
      float ref = t / a;
      float r = 1.0f / a;
      float q0 = t * r;
      float e0 = fmaf(-a, q0, t);
      float q1 = fmaf(e0, r, q0);
      if (!(fabsf(q1) <= 3.0e38f)) continue;                   // kernel: IEEE fallback
      if (ref != 0.0f && fabsf(ref) < 1.2e-38f) continue;      // subnormal quotient: not produced
      if (asu(q1) != asu(ref)) bad++;
      total++;
    }
  }
  printf("total %ld bad %ld\n", total, bad);
  return bad != 0;
}
