/* Test infrastructure (not product code): is  q = fma(fma(-b, q0, a), r, q0),  q0 = RN(a * r),
 * r = RN(1 / b)  bit-identical to IEEE a / b  (1) for the constant divisors 3 and 15 over EVERY
 * positive normal float a whose quotient stays normal, and (2) for random normal (a, b) with
 * quotients in [2^-20, 2^20]?  csrc/ptq.cu uses the sequence for the group parameters of
 * MXQGPT.fasterquant (quantizer.py:94,99,115-121) only where this program reports zero mismatches.
 *   gcc -O2 -fopenmp -o /tmp/div_check_general oracle/div_check_general.c -lm && /tmp/div_check_general */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static inline uint64_t rng(uint64_t* s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }
static inline float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t asu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float mdiv(float a, float b, float r) {
  float q0 = a * r;
  float e0 = fmaf(-b, q0, a);
  return fmaf(e0, r, q0);
}
int main(void) {
  const float ds[2] = {3.0f, 15.0f};
  for (int k = 0; k < 2; ++k) {
    const float d = ds[k], r = 1.0f / d;
    long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long u = 0x03800000L; u < 0x7F000000L; ++u) {      /* a in [2^-120, 2^127): quotient normal */
      const float a = asf((uint32_t)u);
      if (asu(mdiv(a, d, r)) != asu(a / d)) bad++;
    }
    printf("divisor %g: exhaustive over positive normals, mismatches %ld\n", d, bad);
  }
  long bad = 0, total = 0;
#pragma omp parallel reduction(+ : bad, total)
  {
    uint64_t s = 88172645463325252ULL + 104729ULL * (uint64_t)omp_get_thread_num();
    for (long i = 0; i < 1500000000L; ++i) {
      const uint64_t r1 = rng(&s), r2 = rng(&s);
      const int eb = 127 - 30 + (int)((r1 >> 40) % 40);       /* b in [2^-30, 2^10) */
      const int ea = eb - 20 + (int)((r2 >> 40) % 41);        /* quotient in [2^-21, 2^21) */
      uint32_t mb = (uint32_t)r1 & 0x7FFFFF, ma = (uint32_t)r2 & 0x7FFFFF;
      const int mode = (int)((r2 >> 60) & 3);
      if (mode == 1) { mb |= 0x7FFF00; }                       /* mantissa of b close to all ones */
      if (mode == 2) { ma &= 0x7FE000; mb &= 0x7FE000; }       /* fp16-valued operands */
      const float a = asf(((uint32_t)ea << 23) | ma), b = asf(((uint32_t)eb << 23) | mb);
      if (asu(mdiv(a, b, 1.0f / b)) != asu(a / b)) bad++;
      total++;
    }
  }
  printf("random normal pairs: %ld, mismatches %ld\n", total, bad);
  return 0;
}
