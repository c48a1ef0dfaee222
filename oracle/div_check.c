#include <stdio.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <omp.h>
static inline uint64_t rng(uint64_t *s){ *s ^= *s<<13; *s ^= *s>>7; *s ^= *s<<17; return *s; }
static inline float asf(uint32_t u){ float f; memcpy(&f,&u,4); return f; }
static inline uint32_t asu(float f){ uint32_t u; memcpy(&u,&f,4); return u; }
int main(){
  long bad1=0,bad2=0,total=0;
  #pragma omp parallel reduction(+:bad1,bad2,total)
  {
    uint64_t s = 88172645463325252ULL + 7919*omp_get_thread_num();
    for(long i=0;i<400000000L;i++){
      uint64_t r1=rng(&s), r2=rng(&s);
      // a: random mantissa, exponent in [2^-27, 2^4]
      int e = 127 - 27 + (int)(r1>>40)%32;
      uint32_t ma = (uint32_t)r1 & 0x7FFFFF;
      int mode = (r2>>60)&3;
      if(mode==1) ma &= 0x7F0000; // bf16-like
      if(mode==2) ma &= 0x7FE000; // fp16-like
      float a = asf(((uint32_t)e<<23)|ma);
      // t in [0, a]: random float <= a
      uint32_t mt = (uint32_t)r2 & 0x7FFFFF; int et = e - (int)((r2>>32)%12);
      if(mode==1) mt &= 0x7F0000;
      if(mode==2) mt &= 0x7FE000;
      if(et<1) et=1;
      float t = asf(((uint32_t)et<<23)|mt);
      if(t>a) t=a;
      float ref = t/a;
      float r = 1.0f/a;
      float q0 = t*r;
      float e0 = fmaf(-a,q0,t);
      float q1 = fmaf(e0,r,q0);
      float e1 = fmaf(-a,q1,t);
      float q2 = fmaf(e1,r,q1);
      if(asu(q1)!=asu(ref)) bad1++;
      if(asu(q2)!=asu(ref)) bad2++;
      total++;
    }
  }
  printf("total %ld bad1 %ld bad2 %ld\n",total,bad1,bad2);
  // exhaustive small ints
  int badi=0;
  for(int sb=1;sb<=8;sb++){ int sv=(1<<sb)-1; float r=1.0f/sv; for(int q=0;q<=sv;q++){ float q0=q*r; float e0=fmaf(-(float)sv,q0,(float)q); float q1=fmaf(e0,r,q0); if(asu(q1)!=asu((float)q/(float)sv)) badi++; } }
  printf("int bad %d\n",badi);
  return 0;
}
