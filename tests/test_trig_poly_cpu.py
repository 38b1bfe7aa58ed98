"""CPU check of the wide sine kernels of numbacs_b200/csrc/fastmath.cuh: the constants are parsed
out of the CUDA header, a C replica of the device arithmetic (fma, no contraction) is compiled
with gcc and compared with long-double references.  Gates: sin(pi u) 12-instruction form <= 3.5 ulp
and <= 4e-16 absolute; split form <= 3 ulp; sin_wide <= 2.5 ulp (tools/fit_trig_poly.py measures
2.98 / 2.33 / 1.98 on 2e7 samples)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_SRC = r"""
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
static const double K[] = {%s};
#define magic K[0]
#define pi_hi K[1]
#define pi_lo K[2]
#define inv_pi K[3]
static const double *cp = K + 4, *cs = K + 12;
static double flip(double d, int bit){ uint64_t u; memcpy(&u,&d,8); u ^= ((uint64_t)(bit&1))<<63; memcpy(&d,&u,8); return d; }
static int loint(double t){ uint64_t u; memcpy(&u,&t,8); return (int)(uint32_t)u; }
static double sinpi14(double u){ double t=u+magic; int q=loint(t); double r=u-(t-magic), z=r*r, p=cp[7];
  for(int k=6;k>=0;--k) p=fma(p,z,cp[k]); p=fma(p,z,pi_lo); return flip(fma(r,pi_hi,r*p),q&1); }
static double sinpi12(double u){ double t=u+magic; int q=loint(t); double r=u-(t-magic), z=r*r, p=cp[7];
  for(int k=6;k>=0;--k) p=fma(p,z,cp[k]); return flip(r*fma(p,z,pi_hi),q&1); }
static double sin_wide(double x){ double t=fma(x,inv_pi,magic); int q=loint(t); double k=t-magic;
  double r=fma(-k,pi_lo,fma(-k,pi_hi,x)), z=r*r, p=cs[7]; for(int j=6;j>=0;--j) p=fma(p,z,cs[j]); return flip(fma(r*z,p,r),q&1); }
int main(void){ srand48(1); double m14=0,m12=0,a12=0,ms=0;
  for(long i=0;i<2000000;i++){ double u=(drand48()-0.5)*((i%%3==0)?8.0:(i%%3==1?2.0e5:1.0));
    long double kk=roundl((long double)u), rr=(long double)u-kk, tr=sinl(M_PIl*rr); if(((long long)kk)&1) tr=-tr;
    double ulp=fabs(nextafter((double)tr,INFINITY)-(double)tr);
    double e=fabs((double)((long double)sinpi14(u)-tr)); if(e/ulp>m14) m14=e/ulp;
    e=fabs((double)((long double)sinpi12(u)-tr)); if(e/ulp>m12) m12=e/ulp; if(e>a12) a12=e;
    double x=(drand48()-0.5)*((i%%2)?20.0:1.9e5); long double t2=sinl((long double)x);
    ulp=fabs(nextafter((double)t2,INFINITY)-(double)t2); e=fabs((double)((long double)sin_wide(x)-t2)); if(e/ulp>ms) ms=e/ulp; }
  /* exact zeros / extrema on the lattice the double-gyre walls hit */
  int ok = sinpi12(0.0)==0.0 && sinpi12(1.0)==0.0 && sinpi12(2.0)==0.0 && sinpi12(-3.0)==0.0
        && sinpi12(1e15)==0.0 && sinpi12(0.75)==-sinpi12(-0.75) && sinpi12(0.25)==sinpi12(0.75)
        && fabs(sinpi12(0.5)-1.0)<=2.3e-16 && fabs(sinpi12(1.5)+1.0)<=2.3e-16;
  printf("%%.4f %%.4f %%.4e %%.4f %%d\n", m14, m12, a12, ms, ok);
  return 0; }
"""


def test_wide_sine_kernels_accuracy(tmp_path):
    src = open(os.path.join(ROOT, "numbacs_b200", "csrc", "fastmath.cuh")).read()
    m = re.search(r"static __constant__ WideTrigConsts kWide = \{(.*?)\};", src, re.S)
    assert m, "kWide not found in fastmath.cuh"
    nums = re.findall(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?", m.group(1))
    assert len(nums) == 4 + 8 + 8 + 8, nums      # magic, pi_hi, pi_lo, 1/pi, cp[8], cs[8], cc[8]
    c = tmp_path / "t.c"
    c.write_text(C_SRC % ", ".join(nums))
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(c), "-lm"])
    m14, m12, a12, ms, ok = subprocess.check_output([str(exe)], text=True).split()
    assert float(m14) <= 3.0, m14
    assert float(m12) <= 3.5 and float(a12) <= 4e-16, (m12, a12)
    assert float(ms) <= 2.5, ms
    assert int(ok) == 1       # sin(pi k) == 0 exactly, odd, symmetric about 1/2, +-1 within an ulp


C_SRC2 = r"""
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
static const double W[] = {%s};
static const double E[] = {%s};
#define magic W[0]
#define pi_hi W[1]
#define pi_lo W[2]
#define inv_pi W[3]
static const double *cs = W + 12, *cc = W + 20;
static double flip(double d, int bit){ uint64_t u; memcpy(&u,&d,8); u ^= ((uint64_t)(bit&1))<<63; memcpy(&d,&u,8); return d; }
static int loint(double t){ uint64_t u; memcpy(&u,&t,8); return (int)(uint32_t)u; }
static void sincos_wide(double x, double *s, double *c){ double t=fma(x,inv_pi,magic); int q=loint(t); double k=t-magic;
  double r=fma(-k,pi_lo,fma(-k,pi_hi,x)), z=r*r, ps=cs[7], pc=cc[7];
  for(int j=6;j>=0;--j){ ps=fma(ps,z,cs[j]); pc=fma(pc,z,cc[j]); }
  *s=flip(fma(r*z,ps,r),q&1); *c=flip(fma(z,pc,1.0),q&1); }
static double expm1_neg(double x){ x=fmax(x,-64.0); double t=fma(x,E[0],magic); int n=loint(t); double k=t-magic;
  double r=fma(-k,E[2],fma(-k,E[1],x)), q=E[4+11]; for(int j=10;j>=0;--j) q=fma(q,r,E[4+j]);
  double p=fma(r*r,q,r); uint64_t sb=((uint64_t)(uint32_t)(n+1023))<<52; double s; memcpy(&s,&sb,8); return fma(p,s,s-1.0); }
int main(void){ srand48(3); double ms=0, ac=0, me=0;
  for(long i=0;i<2000000;i++){ double x=(drand48()-0.5)*((i%%2)?40.0:1.9e5), s, c; sincos_wide(x,&s,&c);
    long double ts=sinl((long double)x), tc=cosl((long double)x);
    double ulp=fabs(nextafter((double)ts,INFINITY)-(double)ts), e=fabs((double)((long double)s-ts)); if(e/ulp>ms) ms=e/ulp;
    e=fabs((double)((long double)c-tc)); if(e>ac) ac=e;
    int k=i%%4; double y = k==0 ? -drand48()*64 : k==1 ? -drand48()*2 : k==2 ? -drand48()*1e-3 : -pow(10.0,-drand48()*300);
    long double te=expm1l((long double)y); ulp=fabs(nextafter((double)te,-INFINITY)-(double)te); e=fabs((double)((long double)expm1_neg(y)-te)); if(e/ulp>me) me=e/ulp; }
  int ok = expm1_neg(0.0)==0.0 && expm1_neg(-700.0)==-1.0 && expm1_neg(-1e-300)==-1e-300;
  printf("%%.4f %%.4e %%.4f %%d\n", ms, ac, me, ok); return 0; }
"""


def _consts(src, name):
    m = re.search(r"static __constant__ \w+ " + name + r" = \{(.*?)\};", src, re.S)
    assert m, name
    return re.findall(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?", m.group(1))


def test_wide_sincos_and_expm1_accuracy(tmp_path):
    """sincos_wide (sine <= 2.5 ulp, cosine <= 4e-16 ABSOLUTE: it is evaluated as 1 + z P(z)) and
    expm1_neg (<= 1.5 ulp on [-64, 0] down to denormals) of the Bickley-jet RHS."""
    src = open(os.path.join(ROOT, "numbacs_b200", "csrc", "fastmath.cuh")).read()
    w, e = _consts(src, "kWide"), _consts(src, "kExpm1")
    assert len(w) == 4 + 8 + 8 + 8 and len(e) == 4 + 12, (len(w), len(e))
    c = tmp_path / "t.c"
    c.write_text(C_SRC2 % (", ".join(w), ", ".join(e)))
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(c), "-lm"])
    ms, ac, me, ok = subprocess.check_output([str(exe)], text=True).split()
    assert float(ms) <= 2.5, ms
    assert float(ac) <= 4e-16, ac
    assert float(me) <= 1.5, me
    assert int(ok) == 1
