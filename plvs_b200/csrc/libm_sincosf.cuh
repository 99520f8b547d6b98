// Bit-exact device port of glibc's float cosf()/sinf() for |x| < 120.
//
// Why: the reference's descriptor steering is `cos(angle)`/`sin(angle)` on floats inside
// `using namespace std` (src/ORBextractor.cc:99,146-147), i.e. glibc cosf/sinf.  CUDA's cosf/sinf
// differ from glibc in the last ulp on ~0.4-0.8 % of angles (SURVEY.md §8c' item 5), which can flip
// a cvRound in the rBRIEF sampling.  glibc >= 2.28 evaluates both functions in DOUBLE precision:
// reduce by pi/2 (x - n*hpi with n = round(x * 2/pi)), then one of two minimax polynomials, and
// rounds once to float.  Restating that algorithm with the same published coefficients and
// non-fused double arithmetic reproduces libm for every float in [0, 6.4] (1 087 163 598 values
// swept exhaustively on the host by tests/test_sincosf_port.py's generator; 0 mismatches for
// either function, with and without host FMA contraction).  Angles here are deg*pi/180 in [0, 2pi].
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PLVS_HD __host__ __device__ __forceinline__
#else
#define PLVS_HD static inline
#endif

namespace plvs {

#if defined(__CUDA_ARCH__)
#define PLVS_DMUL(a, b) __dmul_rn((a), (b))
#define PLVS_DADD(a, b) __dadd_rn((a), (b))
#else
#define PLVS_DMUL(a, b) ((a) * (b))
#define PLVS_DADD(a, b) ((a) + (b))
#endif

PLVS_HD uint32_t f32_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } v; v.f = f; return v.u;
#endif
}
PLVS_HD uint32_t f32_abstop12(float f) { return (f32_bits(f) >> 20) & 0x7ffu; }

// odd == 0: sine-type polynomial in x (x2 = x*x); odd == 1: cosine-type polynomial.  `neg` flips
// the sign of the cosine coefficients (second half-period).
PLVS_HD float sincosf_poly(double x, double x2, int odd, bool neg)
{
    if (!odd) {
        const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
        double x3 = PLVS_DMUL(x, x2);
        double t = PLVS_DADD(s2, PLVS_DMUL(x2, s3));
        double x7 = PLVS_DMUL(x3, x2);
        double s = PLVS_DADD(x, PLVS_DMUL(x3, s1));
        return (float)PLVS_DADD(s, PLVS_DMUL(x7, t));
    }
    const double sg = neg ? -1.0 : 1.0;
    const double c0 = sg, c1 = sg * -0x1.ffffffd0c621cp-2, c2 = sg * 0x1.55553e1068f19p-5;
    const double c3 = sg * -0x1.6c087e89a359dp-10, c4 = sg * 0x1.99343027bf8c3p-16;
    double x4 = PLVS_DMUL(x2, x2);
    double t2 = PLVS_DADD(c3, PLVS_DMUL(x2, c4));
    double t1 = PLVS_DADD(c0, PLVS_DMUL(x2, c1));
    double x6 = PLVS_DMUL(x4, x2);
    double c = PLVS_DADD(t1, PLVS_DMUL(x4, c2));
    return (float)PLVS_DADD(c, PLVS_DMUL(x6, t2));
}

// (cosf(y), sinf(y)) exactly as glibc returns them, for 0 <= y < 120.
PLVS_HD void libm_sincosf(float y, float* cos_out, float* sin_out)
{
    double x = (double)y;
    if (f32_abstop12(y) < f32_abstop12(0x1.921fb6p-1f)) {          // |y| < pi/4
        if (f32_abstop12(y) < f32_abstop12(0x1p-12f)) { *sin_out = y; *cos_out = 1.0f; return; }
        double x2 = PLVS_DMUL(x, x);
        *sin_out = sincosf_poly(x, x2, 0, false);
        *cos_out = sincosf_poly(x, x2, 1, false);
        return;
    }
    double r = PLVS_DMUL(x, 0x1.45F306DC9C883p+23);                // x * 2/pi * 2^24
    int n = ((int32_t)r + 0x800000) >> 24;
    x = PLVS_DADD(x, -PLVS_DMUL((double)n, 0x1.921FB54442D18p0));  // x - n*pi/2
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const bool neg = (n & 2) != 0;
    double xs = PLVS_DMUL(x, sgn), x2 = PLVS_DMUL(x, x);
    *sin_out = sincosf_poly(xs, x2, n & 1, neg);
    *cos_out = sincosf_poly(xs, x2, (n ^ 1) & 1, neg);
}

}  // namespace plvs
