#!/usr/bin/env python
"""Coefficients and accuracy of the "wide" sine kernels of numbacs_b200/csrc/fastmath.cuh.

    python tools/fit_trig_poly.py            # prints the coefficients, then compiles and runs a C
                                             # replica of the device arithmetic (fma, no contraction)
                                             # against long-double references on 2e7 random arguments

Polynomials: Chebyshev-node interpolation (near-minimax) with 60-digit arithmetic of
    (sin(pi r)/r - pi)/r^2   on r^2 in [0, 1/4]        -> kWide.cp[8]   (sinpi kernels)
    (sin r / r - 1)/r^2      on r^2 in [0, (pi/2)^2]   -> kWide.cs[8]   (sin_wide_v)
Measured here (gcc, glibc): sinpi split form 2.33 ulp, 12-instruction form 2.96 ulp (3.3e-16
absolute), sin_wide 1.98 ulp; the reference's own sin(pi*x), |x| < 4: 1.4e-15 absolute."""
import os
import subprocess
import tempfile

import mpmath as mp

mp.mp.dps = 60


def fit(n, zmax, g):
    nodes = [(mp.mpf(zmax) / 2) * (1 + mp.cos(mp.pi * (2 * k + 1) / (2 * n))) for k in range(n)]
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, z in enumerate(nodes):
        for j in range(n):
            A[i, j] = z ** j
        b[i] = g(z)
    c = mp.lu_solve(A, b)
    return [float(c[i]) for i in range(n)]


def g_sinpi(z):
    r = mp.sqrt(z)
    return -(mp.pi ** 3) / 6 if z == 0 else (mp.sin(mp.pi * r) / r - mp.pi) / z


def g_sin(z):
    r = mp.sqrt(z)
    return -mp.mpf(1) / 6 if z == 0 else (mp.sin(r) / r - 1) / z


cp = fit(8, mp.mpf(1) / 4, g_sinpi)
cs = fit(8, (mp.pi / 2) ** 2 * mp.mpf("1.0001"), g_sin)
pi_hi = float(mp.pi)
pi_lo = float(mp.pi - mp.mpf(pi_hi))
print("cp =", ", ".join(repr(c) for c in cp))
print("cs =", ", ".join(repr(c) for c in cs))
print("pi_hi =", repr(pi_hi), " pi_lo =", repr(pi_lo), " 1/pi =", repr(float(1 / mp.pi)))

C_SRC = r"""
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
static const double magic = 6755399441055744.0, pi_hi = %(pi_hi)s, pi_lo = %(pi_lo)s, inv_pi = %(inv_pi)s;
static const double cp[8] = {%(cp)s};
static const double cs[8] = {%(cs)s};
static double flip(double d, int bit){ uint64_t u; memcpy(&u,&d,8); u ^= ((uint64_t)(bit&1))<<63; memcpy(&d,&u,8); return d; }
static int loint(double t){ uint64_t u; memcpy(&u,&t,8); return (int)(uint32_t)u; }
static double sinpi14(double u){ double t=u+magic; int q=loint(t); double r=u-(t-magic), z=r*r, p=cp[7];
  for(int k=6;k>=0;--k) p=fma(p,z,cp[k]); p=fma(p,z,pi_lo); return flip(fma(r,pi_hi,r*p),q&1); }
static double sinpi12(double u){ double t=u+magic; int q=loint(t); double r=u-(t-magic), z=r*r, p=cp[7];
  for(int k=6;k>=0;--k) p=fma(p,z,cp[k]); return flip(r*fma(p,z,pi_hi),q&1); }
static double sin_wide(double x){ double t=fma(x,inv_pi,magic); int q=loint(t); double k=t-magic;
  double r=fma(-k,pi_lo,fma(-k,pi_hi,x)), z=r*r, p=cs[7]; for(int j=6;j>=0;--j) p=fma(p,z,cs[j]); return flip(fma(r*z,p,r),q&1); }
int main(void){ srand48(1); double m14=0,m12=0,a12=0,ms=0,aref=0;
  for(long i=0;i<20000000;i++){ double u=(drand48()-0.5)*((i%%3==0)?8.0:(i%%3==1?200.0:1.0));
    long double kk=roundl((long double)u), rr=(long double)u-kk, tr=sinl(M_PIl*rr); if(((long long)kk)&1) tr=-tr;
    double ulp=fabs(nextafter((double)tr,INFINITY)-(double)tr);
    double e=fabs((double)((long double)sinpi14(u)-tr)); if(e/ulp>m14) m14=e/ulp;
    e=fabs((double)((long double)sinpi12(u)-tr)); if(e/ulp>m12) m12=e/ulp; if(e>a12) a12=e;
    if(fabs(u)<4){ e=fabs((double)((long double)sin(M_PI*u)-tr)); if(e>aref) aref=e; }
    double x=(drand48()-0.5)*((i%%2)?20.0:2000.0); long double t2=sinl((long double)x);
    ulp=fabs(nextafter((double)t2,INFINITY)-(double)t2); e=fabs((double)((long double)sin_wide(x)-t2)); if(e/ulp>ms) ms=e/ulp; }
  printf("sinpi split form %%.2f ulp | sinpi 12-instruction form %%.2f ulp (%%.2e abs) | sin_wide %%.2f ulp | "
         "reference-style sin(pi*x), |x|<4: %%.2e abs\n", m14, m12, a12, ms, aref);
  return 0; }
"""
with tempfile.TemporaryDirectory() as d:
    src = os.path.join(d, "t.c")
    open(src, "w").write(C_SRC % {"pi_hi": repr(pi_hi), "pi_lo": repr(pi_lo), "inv_pi": repr(float(1 / mp.pi)),
                                  "cp": ", ".join(repr(c) for c in cp), "cs": ", ".join(repr(c) for c in cs)})
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", os.path.join(d, "t"), src, "-lm"])
    subprocess.check_call([os.path.join(d, "t")])
