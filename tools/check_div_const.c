/* Exhaustive fp32 proof obligation for lb::div_const (2d-lb_b200/csrc/lb_device.cuh):
 * for every finite float x, compare the IEEE quotient x / c with the FMA-corrected sequence
 * q0 = x*rc; r = fma(-q0, c, x); q = fma(r, rc, q0), rc = RN(1/c), for the three lattice constants
 * float32(cs2), float32(2 cs2), float32(2 cs^4) (opencl_dim.py:26-30, :305).
 *   gcc -O2 -mfma -fopenmp -ffp-contract=off tools/check_div_const.c -o /tmp/chk -lm && /tmp/chk
 * Expected (about a minute on 8 cores): mismatches(|x|>=1e-30, finite quotient)=0 for all three. */
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
int main(){
  double csd = 1.0/sqrt(3.0);
  double c2 = pow(csd,2), c22 = 2*c2, c4 = 2*pow(csd,4);
  float cs[3] = {(float)c2,(float)c22,(float)c4};
  for(int k=0;k<3;k++){
    float c=cs[k]; float rc=(float)(1.0/(double)c);
    uint64_t bad=0, bad_normal=0; uint32_t first=0; float minbad=1e30f;
    #pragma omp parallel for reduction(+:bad,bad_normal)
    for(uint64_t i=0;i<(1ull<<32);i++){
      uint32_t b=(uint32_t)i; float x; memcpy(&x,&b,4);
      if(!isfinite(x)) continue;
      float ref = x / c;
      float q = x*rc; float r = fmaf(-q,c,x); float q2 = fmaf(r,rc,q);
      uint32_t a1,a2; memcpy(&a1,&ref,4); memcpy(&a2,&q2,4);
      if(a1!=a2 && !(ref==0.0f && q2==0.0f)){ bad++; if(fabsf(x)>=1e-30f && isfinite(ref)) bad_normal++; }
    }
    printf("c=%.9g rc=%.9g mismatches=%llu mismatches(|x|>=1e-30, finite quotient)=%llu\n",c,rc,(unsigned long long)bad,(unsigned long long)bad_normal);
  }
  return 0;
}
