#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
static inline uint64_t rng(uint64_t* s){ uint64_t x=*s; x^=x<<13; x^=x>>7; x^=x<<17; *s=x; return x; }
int main(){
  double rcp[129]; for(int n=1;n<=128;++n) rcp[n]=1.0/(double)n;
  uint64_t s=88172645463325252ull; long bad1=0,bad2=0,tot=0;
  for(int n=1;n<=128;++n){
    double b=(double)n, y=rcp[n];
    for(long i=0;i<6000000;++i){
      uint64_t r=rng(&s); double a;
      int mode=i%4;
      if(mode==0){ a=-(double)(r>>11)*(1.0/9007199254740992.0)*4000.0; }        // typical sums of log probs
      else if(mode==1){ uint64_t bits=(r&0x800fffffffffffffull)|((uint64_t)(1023-40+(r>>52)%80)<<52); memcpy(&a,&bits,8);} // random mantissa, exponent +-40
      else if(mode==2){ double q=(double)(r>>11)*(1.0/9007199254740992.0)*100.0; a=nextafter(q*b, (r&1)?1e300:-1e300);} // near exact multiples
      else { a=-3.4e35*(double)((r>>40)+1)/16777216.0; }
      double ref=a/b;
      double q0=a*y; double r0=fma(-q0,b,a); double q1=fma(r0,y,q0);
      double r1=fma(-q1,b,a); double q2=fma(r1,y,q1);
      if(q1!=ref) ++bad1; if(q2!=ref) ++bad2; ++tot;
    }
  }
  printf("total %ld, one-correction mismatches %ld, two-correction mismatches %ld\n",tot,bad1,bad2);
  return 0;
}
