// rko_math.hpp — elementary functions used by the parity oracle.
//
// TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the shipped
// product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
// leg may build or call it.
//
// The reference (rsruckig, lib/src/rsruckig/roots.rs:189-266,301) calls the
// platform libm for cbrt / atan2 / cos / sin / acos. Two builds of the oracle
// exist:
//   * RKO_MATH_GLIBC    (default): calls glibc libm, i.e. what a Linux build of
//     the Rust reference links against — this is the reference-faithful oracle.
//   * RKO_MATH_PORTABLE: uses the self-contained routines below, which consist
//     only of IEEE +,-,*,/,sqrt and integer bit operations, so the CUDA side
//     (rsruckig_b200/csrc/rk_math.cuh restates the same published fdlibm
//     algorithms) can reproduce them bit for bit. tests/ compares the two
//     builds against each other to bound what the substitution can change.
//
// The portable routines restate the public-domain fdlibm algorithms
// (s_cbrt.c, k_sin.c, k_cos.c, e_acos.c, s_atan.c, e_atan2.c; Sun
// Microsystems 1993, FreeBSD msun revisions) — third-party published
// algorithms, not part of /root/reference.
#pragma once
// Phase markers for the op-counting build (rko_count.cpp defines RKO_PHASE before it includes the oracle): a scope that
// attributes the operations counted inside it to one part of the algorithm. They expand to nothing everywhere else.
#ifndef RKO_PHASE
#define RKO_PHASE(id) ((void)0)
#endif
enum {
    RKO_PH_OTHER = 0,      // validation, brake, block, synchronisation, dispatch
    RKO_PH_S1_QUARTIC,     // position_third_step1.rs:329-601 (NONE / ACC0 / ACC1)
    RKO_PH_S1_ACC0_ACC1,   // :241-327
    RKO_PH_S1_VEL,         // :97-240
    RKO_PH_S1_RESCUE,      // :603-895
    RKO_PH_S2_ACC0_ACC1_VEL,
    RKO_PH_S2_VEL,
    RKO_PH_S2_ACC0_VEL,
    RKO_PH_S2_ACC1_VEL,
    RKO_PH_S2_TAIL,        // ACC0_ACC1, ACC1, ACC0, NONE of Step 2
    RKO_PH_CHECK,          // Profile::check (profile.rs:323-482), whoever calls it
    RKO_PH_ROOTS,          // solve_quart_monic / solve_cub (roots.rs:132-370), whoever calls it
    RKO_PH_NEWTON,         // shrink_interval (roots.rs:425-481)
    RKO_PH_COUNT
};
#include <cmath>
#include <cstdint>
#include <cstring>

namespace rkm {

static inline uint64_t d2u(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
static inline double u2d(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }
static inline uint32_t hi32(double x) { return (uint32_t)(d2u(x) >> 32); }
static inline uint32_t lo32(double x) { return (uint32_t)(d2u(x) & 0xffffffffu); }

// ---- cbrt (fdlibm s_cbrt.c, FreeBSD polynomial variant; error < 0.667 ulp)
static inline double p_cbrt(double x) {
    const uint32_t B1 = 715094163u, B2 = 696219795u;
    const double P0 = 1.87595182427177009643, P1 = -1.88497979543377169875,
                 P2 = 1.621429720105354466140, P3 = -0.758397934778766047437,
                 P4 = 0.145996192886612446982;
    uint32_t hx = hi32(x), low = lo32(x);
    uint32_t sign = hx & 0x80000000u;
    hx ^= sign;
    if (hx >= 0x7ff00000u) return x + x;  // NaN, Inf
    double t;
    if (hx < 0x00100000u) {  // zero or subnormal
        if ((hx | low) == 0) return x;
        t = u2d((uint64_t)0x43500000u << 32);  // 2**54
        t *= x;
        uint32_t high = hi32(t);
        t = u2d((uint64_t)(sign | ((high & 0x7fffffffu) / 3 + B2)) << 32);
    } else {
        t = u2d((uint64_t)(sign | (hx / 3 + B1)) << 32);
    }
    double r = (t * t) * (t / x);
    t = t * ((P0 + r * (P1 + r * P2)) + ((r * r) * r) * (P3 + r * P4));
    t = u2d((d2u(t) + 0x80000000ull) & 0xffffffffc0000000ull);
    double s = t * t;
    r = x / s;
    double w = t + t;
    r = (r - t) / (w + r);
    t = t + t * r;
    return t;
}

// ---- sin / cos kernels on [-pi/4, pi/4] (fdlibm k_sin.c / k_cos.c with y = 0)
static inline double k_sin(double x) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = x * x;
    double v = z * x;
    double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return x + v * (S1 + z * r);
}
static inline double k_cos(double x) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = x * x;
    double w = z * z;
    double r = z * (C1 + z * (C2 + z * C3)) + (w * w) * (C4 + z * (C5 + z * C6));
    double hz = 0.5 * z;
    w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + z * r);
}

// Quadrant reduction with a two-part pi/2 (fdlibm e_rem_pio2.c "medium" first
// step). The reference only ever passes theta in [0, pi/3]; the reduction is
// accurate far beyond that range (|x| < ~1e5) but is not a general one.
static inline void rem_pio2(double x, int& n, double& r) {
    const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
                 pio2_1t = 6.07710050650619224932e-11;
    double fn = x * invpio2 + (x >= 0.0 ? 0.5 : -0.5);
    n = (int)fn;  // truncation toward zero == round-half-away of x*invpio2
    double dn = (double)n;
    r = (x - dn * pio2_1) - dn * pio2_1t;
}
static inline double p_sin(double x) {
    if (!(x == x) || x - x != 0.0) return x - x;  // NaN / Inf -> NaN
    if (x >= -0.78539816339744830962 && x <= 0.78539816339744830962) return k_sin(x);
    int n; double r; rem_pio2(x, n, r);
    switch (n & 3) {
        case 0: return k_sin(r);
        case 1: return k_cos(r);
        case 2: return -k_sin(r);
        default: return -k_cos(r);
    }
}
static inline double p_cos(double x) {
    if (!(x == x) || x - x != 0.0) return x - x;
    if (x >= -0.78539816339744830962 && x <= 0.78539816339744830962) return k_cos(x);
    int n; double r; rem_pio2(x, n, r);
    switch (n & 3) {
        case 0: return k_cos(r);
        case 1: return -k_sin(r);
        case 2: return -k_cos(r);
        default: return k_sin(r);
    }
}

// ---- acos (fdlibm e_acos.c)
static inline double p_acos(double x) {
    const double pi = 3.14159265358979311600e+00, pio2_hi = 1.57079632679489655800e+00,
                 pio2_lo = 6.12323399573676603587e-17,
                 pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    uint32_t hx = hi32(x), lx = lo32(x);
    uint32_t ix = hx & 0x7fffffffu;
    if (ix >= 0x3ff00000u) {  // |x| >= 1
        if (((ix - 0x3ff00000u) | lx) == 0) {
            if ((hx >> 31) == 0) return 0.0;
            return pi + 2.0 * pio2_lo;
        }
        return (x - x) / (x - x);  // NaN
    }
    if (ix < 0x3fe00000u) {  // |x| < 0.5
        if (ix <= 0x3c600000u) return pio2_hi + pio2_lo;
        double z = x * x;
        double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        double r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    } else if ((hx >> 31) != 0) {  // x < -0.5
        double z = (1.0 + x) * 0.5;
        double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        double s = std::sqrt(z);
        double r = p / q;
        double w = r * s - pio2_lo;
        return pi - 2.0 * (s + w);
    } else {  // x > 0.5
        double z = (1.0 - x) * 0.5;
        double s = std::sqrt(z);
        double df = u2d(d2u(s) & 0xffffffff00000000ull);
        double c = (z - df * df) / (s + df);
        double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        double r = p / q;
        double w = r * s + c;
        return 2.0 * (df + w);
    }
}

// ---- atan (fdlibm s_atan.c)
static inline double p_atan(double x) {
    const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01,
                              9.82793723247329054082e-01, 1.57079632679489655800e+00};
    const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17,
                              1.39033110312309984516e-17, 6.12323399573676603587e-17};
    const double aT[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01,
                           1.42857142725034663711e-01,  -1.11111104054623557880e-01,
                           9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                           6.66107313738753120669e-02,  -5.83357013379057348645e-02,
                           4.97687799461593236017e-02,  -3.65315727442169155270e-02,
                           1.62858201153657823623e-02};
    uint32_t hx = hi32(x);
    uint32_t ix = hx & 0x7fffffffu;
    bool neg = (hx >> 31) != 0;
    int id;
    if (ix >= 0x44100000u) {  // |x| >= 2^66
        if (x != x) return x + x;
        return neg ? -(atanhi[3] + atanlo[3]) : (atanhi[3] + atanlo[3]);
    }
    if (ix < 0x3fdc0000u) {  // |x| < 0.4375
        if (ix < 0x3e200000u) return x;
        id = -1;
    } else {
        x = std::fabs(x);
        if (ix < 0x3ff30000u) {      // |x| < 1.1875
            if (ix < 0x3fe60000u) {  // 7/16 <= |x| < 11/16
                id = 0; x = (2.0 * x - 1.0) / (2.0 + x);
            } else {
                id = 1; x = (x - 1.0) / (x + 1.0);
            }
        } else {
            if (ix < 0x40038000u) {  // |x| < 2.4375
                id = 2; x = (x - 1.5) / (1.0 + 1.5 * x);
            } else {
                id = 3; x = -1.0 / x;
            }
        }
    }
    double z = x * x;
    double w = z * z;
    double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return neg ? -z : z;
}

// ---- atan2 (fdlibm e_atan2.c), finite and special arguments
static inline double p_atan2(double y, double x) {
    const double tiny = 1.0e-300, pi_o_2 = 1.5707963267948965580E+00,
                 pi = 3.1415926535897931160E+00, pi_lo = 1.2246467991473531772E-16,
                 pi_o_4 = 7.8539816339744827900E-01;
    if (x != x || y != y) return x + y;
    uint32_t hx = hi32(x), lx = lo32(x), hy = hi32(y), ly = lo32(y);
    uint32_t ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (((hx - 0x3ff00000u) | lx) == 0) return p_atan(y);  // x == 1.0
    int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);
    if ((iy | ly) == 0) {  // y == 0
        switch (m) {
            case 0: case 1: return y;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if ((ix | lx) == 0) return ((hy >> 31) != 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7ff00000u) {
        if (iy == 0x7ff00000u) {
            switch (m) {
                case 0: return pi_o_4 + tiny;
                case 1: return -pi_o_4 - tiny;
                case 2: return 3.0 * pi_o_4 + tiny;
                default: return -3.0 * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
                case 0: return 0.0;
                case 1: return -0.0;
                case 2: return pi + tiny;
                default: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7ff00000u) return ((hy >> 31) != 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    int k = ((int)iy - (int)ix) >> 20;
    double z;
    if (k > 60) { z = pi_o_2 + 0.5 * pi_lo; m &= 1; }
    else if ((hx >> 31) != 0 && k < -60) z = 0.0;
    else z = p_atan(std::fabs(y / x));
    switch (m) {
        case 0: return z;
        case 1: return -z;
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

}  // namespace rkm

namespace rko {
#if defined(RKO_MATH_PORTABLE)
static inline double m_cbrt(double x) { return rkm::p_cbrt(x); }
static inline double m_sin(double x) { return rkm::p_sin(x); }
static inline double m_cos(double x) { return rkm::p_cos(x); }
static inline double m_acos(double x) { return rkm::p_acos(x); }
static inline double m_atan2(double y, double x) { return rkm::p_atan2(y, x); }
#else
static inline double m_cbrt(double x) { return ::cbrt(x); }
static inline double m_sin(double x) { return ::sin(x); }
static inline double m_cos(double x) { return ::cos(x); }
static inline double m_acos(double x) { return ::acos(x); }
static inline double m_atan2(double y, double x) { return ::atan2(y, x); }
#endif

// Rust f64::min / f64::max: return the non-NaN operand (brake.rs:71,91,98-104;
// position_third_step2.rs:278,376,455,539,616-617,1671; roots.rs:247).
static inline double fmin_(double a, double b) { return (a != a) ? b : ((b != b) ? a : (b < a ? b : a)); }
static inline double fmax_(double a, double b) { return (a != a) ? b : ((b != b) ? a : (b > a ? b : a)); }
}  // namespace rko
