// rk_step2.cuh — Step 2 on the device, third-order position interface: find a
// profile of one DoF that lasts exactly tf (position_third_step2.rs).
//
// Each family is a standalone device function over (S2 inputs, direction-
// resolved limits) that either returns false or fills t[7], jf and the two
// codes of the first candidate that passes check_pos3 — the same short-circuit
// order as the reference inside a family. The chain over the 16 (family,
// direction) stages (position_third_step2.rs:2449-2482) lives in the kernel
// file, where stages are executed warp-uniformly over compacted work lists.
// Included once per pass: rk_kernels.cu compiles these headers twice — into namespace rk (exact divisions with their fall-back
// branch) and, with RK_PASS2 defined and `rk` renamed, into namespace rk_lazy (rk_math.cuh, RK_LAZY_DDIV).
#ifndef RK_PASS2
#ifndef RK_H_STEP2_1
#define RK_H_STEP2_1
#define RK_H_STEP2_GO
#endif
#else
#ifndef RK_H_STEP2_2
#define RK_H_STEP2_2
#define RK_H_STEP2_GO
#endif
#endif
#ifdef RK_H_STEP2_GO
#undef RK_H_STEP2_GO
#include "rk_core.cuh"

namespace rk {

struct S2 {
    Bound b;
    real tf;
    // products precomputed by the reference (position_third_step2.rs:51-128)
    real pd, tf_tf, tf_p3, tf_p4, vd, vd_vd, vf_vf, ad, ad_ad, a0_a0, af_af, a0_p3, a0_p4, a0_p5, a0_p6, af_p3, af_p4, af_p5, af_p6, g1, g2;
    Den jj;

    __device__ __forceinline__ void init(const Bound& bb, real tf_, double j_max) {
        b = bb;
        tf = tf_;
        pd = b.pf - b.p0;
        tf_tf = tf * tf; tf_p3 = tf_tf * tf; tf_p4 = tf_tf * tf_tf;
        vd = b.vf - b.v0; vd_vd = vd * vd; vf_vf = b.vf * b.vf;
        ad = b.af - b.a0; ad_ad = ad * ad;
        a0_a0 = b.a0 * b.a0; af_af = b.af * b.af;
        a0_p3 = b.a0 * a0_a0; a0_p4 = a0_a0 * a0_a0; a0_p5 = a0_p3 * a0_a0; a0_p6 = a0_p4 * a0_a0;
        af_p3 = b.af * af_af; af_p4 = af_af * af_af; af_p5 = af_p3 * af_af; af_p6 = af_p4 * af_af;
        jj = Den(j_max * j_max);
        g1 = -pd + tf * b.v0;
        g2 = -2.0 * pd + tf * (b.v0 + b.vf);
    }
};

// Result slot of a family.
struct Hit {
    double t[7];
    double jf;
    int lim, cs, dir;
};

#define RK_S2_LOCALS                                                                                                        \
    const real v0 = s.b.v0, a0 = s.b.a0, vf = s.b.vf, af = s.b.af, tf = s.tf, pd = s.pd, vd = s.vd, ad = s.ad;            \
    const real a0_a0 = s.a0_a0, af_af = s.af_af, a0_p3 = s.a0_p3, af_p3 = s.af_p3, a0_p4 = s.a0_p4, af_p4 = s.af_p4;      \
    const real vd_vd = s.vd_vd, ad_ad = s.ad_ad, tf_tf = s.tf_tf, g1 = s.g1, g2 = s.g2;                                     \
    const Den j_max_j_max = s.jj;                                                                                            \
    const Den v_max = L.v_max, v_min = L.v_min, a_max = L.a_max, a_min = L.a_min, j_max = L.j_max;                         \
    (void)v0; (void)vf; (void)pd; (void)vd; (void)ad; (void)a0_p3; (void)af_p3; (void)a0_p4; (void)af_p4; (void)vd_vd;     \
    (void)ad_ad; (void)tf_tf; (void)g1; (void)g2; (void)j_max_j_max; (void)v_max; (void)v_min; (void)a_min; (void)a0_a0;   \
    (void)af_af; (void)a0; (void)af; (void)tf; (void)a_max; (void)j_max;                                                    \
    double (&t)[7] = h.t;

#define RK_TRY_INLINE(CS, LIM, JF)                               \
    if (check_pos3_inline(t, (CS), (LIM), (JF), L, s.b)) {       \
        h.jf = (double)(JF); h.lim = (LIM); h.cs = (CS);         \
        h.dir = (L.v_max > 0.0) ? DIR_UP : DIR_DOWN;             \
        return true;                                             \
    }

#define RK_TRY(CS, LIM, JF)                                      \
    if (check_pos3(t, (CS), (LIM), (JF), L, s.b)) {              \
        h.jf = (double)(JF); h.lim = (LIM); h.cs = (CS);         \
        h.dir = (L.v_max > 0.0) ? DIR_UP : DIR_DOWN;             \
        return true;                                             \
    }

// position_third_step2.rs:130-245
__device__ RK_S2_FN bool s2_acc0_acc1_vel(const S2& s, const Lim& L, Hit& h) {
    RK_S2_LOCALS
    // Profile UDDU, Solution 1
    if ((2.0 * (a_max - a_min) + ad) / j_max < tf) {
        const real h1 = sqrt((a0_p4 + af_p4 - 4.0 * a0_p3 * (2.0 * a_max + a_min) / THREE - 4.0 * af_p3 * (a_max + 2.0 * a_min) / THREE
                                + 2.0 * (a0_a0 - af_af) * a_max * a_max
                                + (4.0 * a0 * a_max - 2.0 * a0_a0) * (af_af - 2.0 * af * a_min + (a_min - a_max) * a_min + 2.0 * j_max * (a_min * tf - vd))
                                + 2.0 * af_af * (a_min * a_min + 2.0 * j_max * (a_max * tf - vd))
                                + 4.0 * j_max * (2.0 * a_min * (af * vd + j_max * g1) + (a_max * a_max - a_min * a_min) * vd + j_max * vd_vd)
                                + 8.0 * a_max * j_max_j_max * (pd - tf * vf))
                                   / (a_max * a_min)
                               + 4.0 * af_af + 2.0 * a0_a0 + (4.0 * af + a_max - a_min) * (a_max - a_min)
                               + 4.0 * j_max * (a_min - a_max + j_max * tf - 2.0 * af) * tf)
                          * fabs(j_max) / j_max;
        t[0] = (-a0 + a_max) / j_max;
        t[1] = (-(af_af - a0_a0 + 2.0 * a_max * a_max + a_min * (a_min - 2.0 * ad - 3.0 * a_max) + 2.0 * j_max * (a_min * tf - vd)) + a_min * h1)
               / (2.0 * (a_max - a_min) * j_max);
        t[2] = a_max / j_max;
        t[3] = (a_min - a_max + h1) / twice(j_max);
        t[4] = -a_min / j_max;
        t[5] = tf - (t[0] + t[1] + t[2] + t[3] + 2.0 * t[4] + af / j_max);
        t[6] = t[4] + af / j_max;
        RK_TRY_INLINE(UDDU, L_ACC0_ACC1_VEL, j_max)
    }
    // Profile UDUD
    if ((-a0 + 4.0 * a_max - af) / j_max < tf) {
        t[0] = (-a0 + a_max) / j_max;
        t[1] = (3.0 * (a0_p4 + af_p4) - 4.0 * (a0_p3 + af_p3) * a_max - 4.0 * af_p3 * a_max + 24.0 * (a0 + af) * a_max * a_max * a_max
                - 6.0 * (af_af + a0_a0) * (a_max * a_max - 2.0 * j_max * vd)
                + 6.0 * a0_a0 * (af_af - 2.0 * af * a_max - 2.0 * a_max * j_max * tf)
                - 12.0 * a_max * a_max * (2.0 * a_max * a_max - 2.0 * a_max * j_max * tf + j_max * vd)
                - 24.0 * af * a_max * j_max * vd + 12.0 * j_max_j_max * (2.0 * a_max * g1 + vd_vd))
               / (12.0 * a_max * j_max * (a0_a0 + af_af - 2.0 * (a0 + af) * a_max + 2.0 * (a_max * a_max - a_max * j_max * tf + j_max * vd)));
        t[2] = a_max / j_max;
        t[3] = (-a0_a0 - af_af + 2.0 * a_max * (a0 + af - 2.0 * a_max) - 2.0 * j_max * vd) / (2.0 * a_max * j_max) + tf;
        t[4] = t[2];
        t[5] = tf - (t[0] + t[1] + t[2] + t[3] + 2.0 * t[4] - af / j_max);
        t[6] = t[4] - af / j_max;
        RK_TRY_INLINE(UDUD, L_ACC0_ACC1_VEL, j_max)
    }
    return false;
}

// position_third_step2.rs:247-414
__device__ RK_S2_FN bool s2_acc1_vel(const S2& s, const Lim& L, Hit& h) {
    RK_S2_LOCALS
    // Profile UDDU
    {
        const real ph1 = a0_a0 + af_af - a_min * (a0 + 2.0 * af - a_min) - 2.0 * j_max * (vd - a_min * tf);
        const real ph2 = 2.0 * a_min * (j_max * g1 + af * vd) - a_min * a_min * vd + j_max * vd_vd;
        const real ph3 = af_af + a_min * (a_min - 2.0 * af) - 2.0 * j_max * (vd - a_min * tf);
        const real c0 = (2.0 * (2.0 * a0 - a_min)) / j_max;
        const real c1 = (4.0 * a0_a0 + ph1 - 3.0 * a0 * a_min) / j_max_j_max;
        const real c2 = (2.0 * a0 * ph1) / (j_max_j_max * j_max);
        const real c3 = (3.0 * (a0_p4 + af_p4) - 4.0 * (a0_p3 + 2.0 * af_p3) * a_min + 6.0 * af_af * (a_min * a_min - 2.0 * j_max * vd)
                           + 12.0 * j_max * ph2 + 6.0 * a0_a0 * ph3)
                          / (12.0 * j_max_j_max * j_max_j_max);
        const real t_min = -a0 / j_max;
        const real t_max = rmin((tf + 2.0 * a_min / j_max - (a0 + af) / j_max) * 0.5, (a_max - a0) / j_max);
        const Roots roots = solve_quart_monic(c0, c1, c2, c3);
#pragma unroll 1
        for (int ri = 0; ri < roots.n; ++ri) {
            real tt = roots.get(ri);
            if (tt < t_min || tt > t_max) continue;
            // Single Newton step (regarding pd)
            if (fabs(a0 + j_max * tt) > 16.0 * EPS) {
                const real h0 = j_max * tt * tt;
                const real orig = -pd
                                    + (3.0 * (a0_p4 + af_p4) - 8.0 * af_p3 * a_min - 4.0 * a0_p3 * a_min
                                       + 6.0 * af_af * (a_min * a_min + 2.0 * j_max * (h0 - vd))
                                       + 6.0 * a0_a0 * (af_af - 2.0 * af * a_min + a_min * a_min + 2.0 * a_min * j_max * (-2.0 * tt + tf) + 2.0 * j_max * (5.0 * h0 - vd))
                                       + 24.0 * a0 * j_max * tt * (a0_a0 + af_af - 2.0 * af * a_min + a_min * a_min + 2.0 * j_max * (a_min * (-tt + tf) + h0 - vd))
                                       - 24.0 * af * a_min * j_max * (h0 - vd)
                                       + 12.0 * j_max * (a_min * a_min * (h0 - vd) + j_max * (h0 - vd) * (h0 - vd)))
                                          / (24.0 * a_min * j_max_j_max)
                                    + h0 * (tf - tt) + tf * v0;
                const real deriv = (a0 + j_max * tt)
                                     * ((a0_a0 + af_af) / (a_min * j_max) + (a_min - a0 - 2.0 * af) / j_max + (4.0 * a0 * tt + 2.0 * h0 - 2.0 * vd) / a_min + 2.0 * tf - 3.0 * tt);
                tt -= orig / deriv;
            }
            const real h1 = -((a0_a0 + af_af) * 0.5 + j_max * (-vd + 2.0 * a0 * tt + j_max * tt * tt)) / a_min;
            t[0] = tt;
            t[1] = 0.0;
            t[2] = a0 / j_max + tt;
            t[3] = tf - (h1 - a_min + a0 + af) / j_max - 2.0 * tt;
            t[4] = -a_min / j_max;
            t[5] = (h1 + a_min) / j_max;
            t[6] = t[4] + af / j_max;
            RK_TRY(UDDU, L_ACC1_VEL, j_max)
        }
    }
    // Profile UDUD
    {
        const real ph1 = a0_a0 - af_af + (2.0 * af - a0) * a_max - a_max * a_max - 2.0 * j_max * (vd - a_max * tf);
        const real ph2 = a_max * a_max + 2.0 * j_max * vd;
        const real ph3 = af_af + ph2 - 2.0 * a_max * (af + j_max * tf);
        const real ph4 = 2.0 * a_max * j_max * g1 + a_max * a_max * vd + j_max * vd_vd;
        const real c0 = (4.0 * a0 - 2.0 * a_max) / j_max;
        const real c1 = (4.0 * a0_a0 - 3.0 * a0 * a_max + ph1) / j_max_j_max;
        const real c2 = (2.0 * a0 * ph1) / (j_max_j_max * j_max);
        const real c3 = (3.0 * (a0_p4 + af_p4) - 4.0 * (a0_p3 + 2.0 * af_p3) * a_max - 24.0 * af * a_max * j_max * vd + 12.0 * j_max * ph4
                           - 6.0 * a0_a0 * ph3 + 6.0 * af_af * ph2)
                          / (12.0 * j_max_j_max * j_max_j_max);
        const real t_min = -a0 / j_max;
        const real t_max = rmin((tf + ad / j_max - 2.0 * a_max / j_max) * 0.5, (a_max - a0) / j_max);
        const Roots roots = solve_quart_monic(c0, c1, c2, c3);
#pragma unroll 1
        for (int ri = 0; ri < roots.n; ++ri) {
            const real tt = roots.get(ri);
            if (tt > t_max || tt < t_min) continue;
            const real h1 = ((a0_a0 - af_af) * 0.5 + j_max_j_max * tt * tt - j_max * (vd - 2.0 * a0 * tt)) / a_max;
            t[0] = tt;
            t[1] = 0.0;
            t[2] = tt + a0 / j_max;
            t[3] = tf + (h1 + ad - a_max) / j_max - 2.0 * tt;
            t[4] = a_max / j_max;
            t[5] = -(h1 + a_max) / j_max;
            t[6] = t[4] - af / j_max;
            RK_TRY(UDUD, L_ACC1_VEL, j_max)
        }
    }
    return false;
}

// position_third_step2.rs:416-605
__device__ RK_S2_FN bool s2_acc0_vel(const S2& s, const Lim& L, Hit& h) {
    RK_S2_LOCALS
    if (tf < rmax((-a0 + a_max) / j_max, 0.0) + rmax(a_max / j_max, 0.0)) return false;

    const real ph1 = 12.0 * j_max * (-a_max * a_max * vd - j_max * vd_vd + 2.0 * a_max * j_max * (-pd + tf * vf));
    // Profile UDDU
    {
        const real c0 = (2.0 * a_max) / j_max;
        const real c1 = (a0_a0 - af_af + 2.0 * ad * a_max + a_max * a_max + 2.0 * j_max * (vd - a_max * tf)) / j_max_j_max;
        const real c2 = 0.0;
        const real c3 = -(-3.0 * (a0_p4 + af_p4) + 4.0 * (af_p3 + 2.0 * a0_p3) * a_max - 12.0 * a0 * a_max * (af_af - 2.0 * j_max * vd)
                            + 6.0 * a0_a0 * (af_af - a_max * a_max - 2.0 * j_max * vd)
                            + 6.0 * af_af * (a_max * a_max - 2.0 * a_max * j_max * tf + 2.0 * j_max * vd) + ph1)
                          / (12.0 * j_max_j_max * j_max_j_max);
        const real t_min = -af / j_max;
        const real t_max = rmin(tf - (2.0 * a_max - a0) / j_max, -a_min / j_max);
        const Roots roots = solve_quart_monic(c0, c1, c2, c3);
#pragma unroll 1
        for (int ri = 0; ri < roots.n; ++ri) {
            real tt = roots.get(ri);
            if (tt < t_min || tt > t_max) continue;
            // Single Newton step (regarding pd)
            if (tt > EPS) {
                const real h1 = j_max * tt * tt + vd;
                const real orig = (-3.0 * (a0_p4 + af_p4) + 4.0 * (af_p3 + 2.0 * a0_p3) * a_max - 24.0 * af * a_max * j_max_j_max * tt * tt
                                     - 12.0 * a0 * a_max * (af_af - 2.0 * j_max * h1)
                                     + 6.0 * a0_a0 * (af_af - a_max * a_max - 2.0 * j_max * h1)
                                     + 6.0 * af_af * (a_max * a_max - 2.0 * a_max * j_max * tf + 2.0 * j_max * h1)
                                     - 12.0 * j_max * (a_max * a_max * h1 + j_max * h1 * h1 + 2.0 * a_max * j_max * (pd + j_max * tt * tt * (tt - tf) - tf * vf)))
                                    / (24.0 * a_max * j_max_j_max);
                const real deriv = -tt * (a0_a0 - af_af + 2.0 * a_max * (ad - j_max * tf) + a_max * a_max + 3.0 * a_max * j_max * tt + 2.0 * j_max * h1) / a_max;
                tt -= orig / deriv;
            }
            const real h1 = ((a0_a0 - af_af) * 0.5 + j_max * (j_max * tt * tt + vd)) / a_max;
            t[0] = (-a0 + a_max) / j_max;
            t[1] = (h1 - a_max) / j_max;
            t[2] = a_max / j_max;
            t[3] = tf - (h1 + ad + a_max) / j_max - 2.0 * tt;
            t[4] = tt;
            t[5] = 0.0;
            t[6] = af / j_max + tt;
            RK_TRY(UDDU, L_ACC0_VEL, j_max)
        }
    }
    // Profile UDUD
    {
        const real c0 = (-2.0 * a_max) / j_max;
        const real c1 = -(a0_a0 + af_af - 2.0 * (a0 + af) * a_max + a_max * a_max + 2.0 * j_max * (vd - a_max * tf)) / j_max_j_max;
        const real c2 = 0.0;
        const real c3 = (3.0 * (a0_p4 + af_p4) - 4.0 * (af_p3 + 2.0 * a0_p3) * a_max + 6.0 * a0_a0 * (af_af + a_max * a_max + 2.0 * j_max * vd)
                           - 12.0 * a0 * a_max * (af_af + 2.0 * j_max * vd)
                           + 6.0 * af_af * (a_max * a_max - 2.0 * a_max * j_max * tf + 2.0 * j_max * vd) - ph1)
                          / (12.0 * j_max_j_max * j_max_j_max);
        const real t_min = af / j_max;
        const real t_max = rmin(tf - a_max / j_max, a_max / j_max);
        const Roots roots = solve_quart_monic(c0, c1, c2, c3);
#pragma unroll 1
        for (int ri = 0; ri < roots.n; ++ri) {
            real tt = roots.get(ri);
            if (tt < t_min || tt > t_max) continue;
            // Single Newton step (regarding pd)
            {
                const real h1 = j_max * tt * tt - vd;
                const real orig = -(3.0 * (a0_p4 + af_p4) - 4.0 * (2.0 * a0_p3 + af_p3) * a_max + 24.0 * af * a_max * j_max_j_max * tt * tt
                                      - 12.0 * a0 * a_max * (af_af - 2.0 * j_max * h1)
                                      + 6.0 * a0_a0 * (af_af + a_max * a_max - 2.0 * j_max * h1)
                                      + 6.0 * af_af * (a_max * a_max - 2.0 * j_max * (tf * a_max + h1))
                                      + 12.0 * j_max * (-a_max * a_max * h1 + j_max * h1 * h1 - 2.0 * a_max * j_max * (-pd + j_max * tt * tt * (tt - tf) + tf * vf)))
                                    / (24.0 * a_max * j_max_j_max);
                const real deriv = tt * (a0_a0 + af_af - 2.0 * j_max * h1 - 2.0 * (a0 + af + j_max * tf) * a_max + a_max * a_max + 3.0 * a_max * j_max * tt) / a_max;
                tt -= orig / deriv;
            }
            const real h1 = ((a0_a0 + af_af) * 0.5 + j_max * (vd - j_max * tt * tt)) / a_max;
            t[0] = (-a0 + a_max) / j_max;
            t[1] = (h1 - a_max) / j_max;
            t[2] = a_max / j_max;
            t[3] = tf - (h1 - a0 - af + a_max) / j_max - 2.0 * tt;
            t[4] = tt;
            t[5] = 0.0;
            t[6] = -(af / j_max) + tt;
            RK_TRY(UDUD, L_ACC0_VEL, j_max)
        }
    }
    return false;
}

}  // namespace rk
#endif  // RK_H_STEP2_GO
