// pow2_glibc.cuh -- bit-exact glibc pow(x, 2.0) on the device.
//
// FastMarching3D.py:68-71 squares NumPy *scalars* with `**2` ((Tmax-Tarray[a])**2, C**2, (sumlist(Tarray))**2):
// that is libm pow(x, 2.0), and glibc's pow is not correctly rounded -- pow(x, 2.0) != x*x for ~0.06 % of the
// inputs.  The difference is one ulp, far inside the 1e-9 tolerance of the field, but it decides exact TIES of the
// pop order, and with them the accepted set of the reference's early exit on uniform-cost volumes.  This is the
// algorithm of glibc 2.39 sysdeps/ieee754/dbl-64/e_pow.c (log_inline / exp_inline, ARM optimized routines) for
// y = 2 and finite x > 0, as the x86_64 FMA build executes it: __builtin_fma where the source says so AND where GCC
// contracts a*b+c (the multiarch variant is compiled with -mfma and the default -ffp-contract=fast); each contraction
// below was checked against libm bit for bit (tests/test_kernel_logic_emu.py: 1e7 inputs; 6e7 while porting).
// Tables: pow2_glibc_tables.inc, generated from the image's own libm by tools/gen_pow2_tables.py.
// Outside [e^-255, e^255] (where exp_inline leaves its main path) the plain product is returned.
#pragma once
#include "fm_common.cuh"

namespace fmb {

#ifdef FMB_HOST_EMU
#define POW2_TABLE static const
#else
#define POW2_TABLE __device__ const
#endif
#include "pow2_glibc_tables.inc"
#undef POW2_TABLE

__device__ __forceinline__ double pow2_glibc(double x) {
    if (!(x > 0.0) || !(x < __longlong_as_double(0x7ff0000000000000LL))) return __dmul_rn(x, x);   // 0, inf, NaN
    unsigned long long ix = (unsigned long long)__double_as_longlong(x);
    if ((ix >> 52) == 0) {                                   // subnormal: normalise so that the exponent becomes negative
        ix = (unsigned long long)__double_as_longlong(__dmul_rn(x, 4503599627370496.0));
        ix -= 52ULL << 52;
    }
    // ---- log_inline: x = 2^k z, z in [OFF, 2 OFF); log(x) = k ln2 + log(c) + log1p(z/c - 1)
    const unsigned long long tmp = ix - 0x3fe6955500000000ULL;
    const int i = (int)((tmp >> (52 - 7)) % 128);
    const int k = (int)((long long)tmp >> 52);
    const unsigned long long iz = ix - (tmp & (0xfffULL << 52));
    const double z = __longlong_as_double((long long)iz), kd = (double)k;
    const double invc = POW2_LOGT[3 * i], logc = POW2_LOGT[3 * i + 1], logctail = POW2_LOGT[3 * i + 2];
    const double r = __fma_rn(z, invc, -1.0);
    const double t1 = __fma_rn(kd, POW2_LN2HI, logc);
    const double t2 = __dadd_rn(t1, r);
    const double lo1 = __fma_rn(kd, POW2_LN2LO, logctail);
    const double lo2 = __dadd_rn(__dsub_rn(t1, t2), r);
    const double ar = __dmul_rn(POW2_A[0], r), ar2 = __dmul_rn(r, ar), ar3 = __dmul_rn(r, ar2);
    const double hi = __dadd_rn(t2, ar2);
    const double lo3 = __fma_rn(ar, r, -ar2);
    const double lo4 = __dadd_rn(__dsub_rn(t2, hi), ar2);
    const double q1 = __fma_rn(r, POW2_A[6], POW2_A[5]);
    const double q2 = __fma_rn(ar2, q1, __fma_rn(r, POW2_A[4], POW2_A[3]));
    const double q3 = __fma_rn(ar2, q2, __fma_rn(r, POW2_A[2], POW2_A[1]));
    const double lo = __fma_rn(ar3, q3, __dadd_rn(__dadd_rn(__dadd_rn(lo1, lo2), lo3), lo4));
    const double y = __dadd_rn(hi, lo);
    const double tail = __dadd_rn(__dsub_rn(hi, y), lo);
    // ---- pow: ehi + elo = 2 log(x)
    const double ehi = __dmul_rn(2.0, y);
    const double elo = __fma_rn(2.0, tail, __fma_rn(2.0, y, -ehi));
    // ---- exp_inline(ehi, elo)
    const unsigned abstop = (unsigned)(((unsigned long long)__double_as_longlong(ehi) >> 52) & 0x7ff);
    if (abstop - 0x3c9u >= 0x408u - 0x3c9u) {
        if (abstop - 0x3c9u >= 0x80000000u) return __dadd_rn(1.0, ehi);      // |2 log x| < 2^-54
        return __dmul_rn(x, x);                                              // outside the supported range
    }
    double kd2 = __fma_rn(POW2_INVLN2N, ehi, POW2_SHIFT);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd2);
    kd2 = __dsub_rn(kd2, POW2_SHIFT);
    double rr = __fma_rn(kd2, POW2_NEGLN2LON, __fma_rn(kd2, POW2_NEGLN2HIN, ehi));
    rr = __dadd_rn(rr, elo);
    const unsigned idx = 2u * (unsigned)(ki % 128);
    const unsigned long long top = ki << (52 - 7);
    const double tl = __longlong_as_double((long long)POW2_EXPT[idx]);
    const unsigned long long sbits = POW2_EXPT[idx + 1] + top;
    const double r2 = __dmul_rn(rr, rr);
    const double e2 = __fma_rn(r2, __fma_rn(rr, POW2_C[1], POW2_C[0]), __dadd_rn(tl, rr));
    const double tm = __fma_rn(__dmul_rn(r2, r2), __fma_rn(rr, POW2_C[3], POW2_C[2]), e2);
    const double scale = __longlong_as_double((long long)sbits);
    return __fma_rn(scale, tm, scale);
}

}  // namespace fmb
