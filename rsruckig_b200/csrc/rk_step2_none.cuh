// rk_step2_none.cuh — Step 2 family NONE (position_third_step2.rs:1439-2202):
// no limit is reached. The tail of the Step-2 chain: four quartics, two cubics
// and several closed-form three-step profiles, tried in the reference's order.
// Rare on random inputs (< 0.1 % of Step-2 results) but the heaviest family.
// Included once per pass: rk_kernels.cu compiles these headers twice — into namespace rk (exact divisions with their fall-back
// branch) and, with RK_PASS2 defined and `rk` renamed, into namespace rk_lazy (rk_math.cuh, RK_LAZY_DDIV).
#ifndef RK_PASS2
#ifndef RK_H_STEP2_NONE_1
#define RK_H_STEP2_NONE_1
#define RK_H_STEP2_NONE_GO
#endif
#else
#ifndef RK_H_STEP2_NONE_2
#define RK_H_STEP2_NONE_2
#define RK_H_STEP2_NONE_GO
#endif
#endif
#ifdef RK_H_STEP2_NONE_GO
#undef RK_H_STEP2_NONE_GO
#include "rk_step2_acc.cuh"

namespace rk {

__device__ RK_S2_FN bool s2_none(const S2& s, const Lim& L, Hit& h) {
    RK_S2_LOCALS
    const real vf_vf = s.vf_vf, tf_p3 = s.tf_p3, tf_p4 = s.tf_p4, a0_p5 = s.a0_p5, a0_p6 = s.a0_p6, af_p5 = s.af_p5, af_p6 = s.af_p6;

    if (fabs(v0) < EPS && fabs(a0) < EPS && fabs(af) < EPS) {
        const real h1 = sqrt(tf_tf * vf_vf + sq(4.0 * pd - tf * vf));
        const real jf = 4.0 * (4.0 * pd - 2.0 * tf * vf + h1) / tf_p3;
        t[0] = tf * 0.25;
        t[1] = 0.0;
        t[2] = 2.0 * t[0];
        t[3] = 0.0;
        t[4] = 0.0;
        t[5] = 0.0;
        t[6] = t[0];
        RK_TRY(UDDU, L_NONE, jf)
    }

    if (fabs(a0) < EPS && fabs(af) < EPS) {
        // a3 != 0, UDDU: first acc, then constant
        const real c0 = -2.0 * tf;
        const real c1 = 2.0 * vd / j_max + tf_tf;
        const real c2 = 4.0 * (pd - tf * vf) / j_max;
        const real c3 = (vd_vd + j_max * tf * g2) / (j_max_j_max);
        const Roots roots = solve_quart_monic(c0, c1, c2, c3);
#pragma unroll 1
        for (int ri = 0; ri < roots.n; ++ri) {
            real tt = roots.get(ri);
            if (tt > tf * 0.5 || tt > (a_max - a0) / j_max) continue;
            // Single Newton step (regarding pd)
            {
                const real h1 = (j_max * tt * (tt - tf) + vd) / (j_max * (2.0 * tt - tf));
                const real h2 = (2.0 * j_max * tt * (tt - tf) + j_max * tf_tf - 2.0 * vd) / (j_max * (2.0 * tt - tf) * (2.0 * tt - tf));
                const real orig = (-2.0 * pd + 2.0 * tf * v0 + h1 * h1 * j_max * (tf - 2.0 * tt) + j_max * tf * (2.0 * h1 * tt - tt * tt - (h1 - tt) * tf)) * 0.5;
                const real deriv = (j_max * tf * (2.0 * tt - tf) * (h2 - 1.0)) * 0.5 + h1 * j_max * (tf - (2.0 * tt - tf) * h2 - h1);
                tt -= orig / deriv;
            }
            t[0] = tt;
            t[1] = 0.0;
            t[2] = (j_max * tt * (tt - tf) + vd) / (j_max * (2.0 * tt - tf));
            t[3] = tf - 2.0 * tt;
            t[4] = tt - t[2];
            t[5] = 0.0;
            t[6] = 0.0;
            RK_TRY(UDDU, L_NONE, j_max)
        }
    }

    // UDUD T 0246
    {
        const real h0 = sqrt(2.0 * j_max_j_max
                               * (2.0 * sq(a0_p3 - af_p3 - 3.0 * af_af * j_max * tf + 9.0 * af * j_max_j_max * tf_tf - 3.0 * a0_a0 * (af + j_max * tf)
                                           + 3.0 * a0 * sq(af + j_max * tf) + 3.0 * j_max_j_max * (8.0 * pd + j_max * tf_tf * tf - 8.0 * tf * vf))
                                  - 3.0 * (a0_a0 + af_af - 2.0 * af * j_max * tf - 2.0 * a0 * (af + j_max * tf) - j_max * (j_max * tf_tf + 4.0 * v0 - 4.0 * vf))
                                        * (a0_p4 + af_p4 + 4.0 * af_p3 * j_max * tf + 6.0 * af_af * j_max_j_max * tf_tf
                                           - 3.0 * j_max_j_max * j_max_j_max * tf_tf * tf_tf - 4.0 * a0_p3 * (af + j_max * tf) + 6.0 * a0_a0 * sq(af + j_max * tf)
                                           - 12.0 * af * j_max_j_max * (8.0 * pd + j_max * tf_tf * tf - 8.0 * tf * v0) + 48.0 * j_max_j_max * vd_vd
                                           + 48.0 * j_max_j_max * j_max * tf * g2
                                           - 4.0 * a0 * (af_p3 + 3.0 * af_af * j_max * tf - 9.0 * af * j_max_j_max * tf_tf
                                                         - 3.0 * j_max_j_max * (8.0 * pd + j_max * tf_tf * tf - 8.0 * tf * vf)))))
                          / j_max;
        const real h1 = 12.0 * j_max * (-a0_a0 - af_af + 2.0 * af * j_max * tf + 2.0 * a0 * (af + j_max * tf) + j_max * (j_max * tf_tf + 4.0 * v0 - 4.0 * vf));
        const real h2 = -4.0 * a0_p3 + 4.0 * af_p3 + 12.0 * a0_a0 * af - 12.0 * a0 * af_af + 48.0 * j_max_j_max * pd + 12.0 * (a0_a0 - af_af) * j_max * tf
                          - 24.0 * j_max_j_max * tf * (v0 + vf) + 24.0 * ad * j_max * vd;
        const real h3 = 2.0 * a0_p3 - 2.0 * af_p3 - 6.0 * a0_a0 * af + 6.0 * a0 * af_af;
        t[0] = (h3 - 48.0 * j_max_j_max * (tf * vf - pd) - 6.0 * (a0_a0 + af_af) * j_max * tf + 12.0 * a0 * af * j_max * tf
                + 6.0 * (a0 + 3.0 * af + j_max * tf) * tf_tf * j_max_j_max - h0)
               / h1;
        t[1] = 0.0;
        t[2] = (h2 + h0) / h1;
        t[3] = 0.0;
        t[4] = (-h2 + h0) / h1;
        t[5] = 0.0;
        t[6] = (-h3 + 48.0 * j_max_j_max * (tf * v0 - pd) - 6.0 * (a0_a0 + af_af) * j_max * tf + 12.0 * a0 * af * j_max * tf
                + 6.0 * (af + 3.0 * a0 + j_max * tf) * tf_tf * j_max_j_max - h0)
               / h1;
        RK_TRY(UDUD, L_NONE, j_max)
    }

    // a3 != 0, UDDU — T 0234
    {
        const real ph1 = af + j_max * tf;
        const real c0 = -2.0 * (ad + j_max * tf) / j_max;
        const real c1 = 2.0 * (a0_a0 + af_af + j_max * (af * tf + vd) - 2.0 * a0 * ph1) / j_max_j_max + tf_tf;
        const real c2 = 2.0 * (a0_p3 - af_p3 - 3.0 * af_af * j_max * tf + 3.0 * a0 * ph1 * (ph1 - a0) - 6.0 * j_max_j_max * (-pd + tf * vf)) / (3.0 * j_max_j_max * j_max);
        const real c3 = (a0_p4 + af_p4 + 4.0 * af_p3 * j_max * tf - 4.0 * a0_p3 * ph1 + 6.0 * a0_a0 * ph1 * ph1 + 24.0 * j_max_j_max * af * g1
                           - 4.0 * a0 * (af_p3 + 3.0 * af_af * j_max * tf + 6.0 * j_max_j_max * (-pd + tf * vf)) + 6.0 * j_max_j_max * af_af * tf_tf
                           + 12.0 * j_max_j_max * (vd_vd + j_max * tf * g2))
                          / (12.0 * j_max_j_max * j_max_j_max);
        const real t_min = ad / j_max;
        const real t_max = rmin((a_max - a0) / j_max, (ad / j_max + tf) * 0.5);
        const Roots roots = solve_quart_monic(c0, c1, c2, c3);
#pragma unroll 1
        for (int ri = 0; ri < roots.n; ++ri) {
            real tt = roots.get(ri);
            if (tt < t_min || tt > t_max) continue;
            // Single Newton step (regarding pd)
            {
                const real h0 = j_max * (2.0 * tt - tf) - ad;
                const real h1 = (ad_ad - 2.0 * af * j_max * tt + 2.0 * a0 * j_max * (tt - tf) + 2.0 * j_max * (j_max * tt * (tt - tf) + vd)) / (2.0 * j_max * h0);
                const real h2 = (-ad_ad + 2.0 * j_max_j_max * (tf_tf + tt * (tt - tf)) + (a0 + af) * j_max * tf - ad * h0 - 2.0 * j_max * vd) / (h0 * h0);
                const real orig = (-a0_p3 + af_p3 + 3.0 * ad_ad * j_max * (h1 - tt) + 3.0 * ad * j_max_j_max * (h1 - tt) * (h1 - tt) - 3.0 * a0 * af * ad
                                     + 3.0 * j_max_j_max
                                           * (a0 * tf_tf - 2.0 * pd + 2.0 * tf * v0 + h1 * h1 * j_max * (tf - 2.0 * tt)
                                              + j_max * tf * (2.0 * h1 * tt - tt * tt - (h1 - tt) * tf)))
                                    / (6.0 * j_max_j_max);
                const real deriv = (h0 * (-ad + j_max * tf) * (h2 - 1.0)) / twice(j_max) + h1 * (-ad + j_max * (tf - h1) - h0 * h2);
                tt -= orig / deriv;
            }
            t[0] = tt;
            t[1] = 0.0;
            t[2] = (ad_ad + 2.0 * j_max * (-a0 * tf - ad * tt + j_max * tt * (tt - tf) + vd)) / (2.0 * j_max * (-ad + j_max * (2.0 * tt - tf)));
            t[3] = ad / j_max + tf - 2.0 * tt;
            t[4] = tf - (tt + t[2] + t[3]);
            t[5] = 0.0;
            t[6] = 0.0;
            RK_TRY(UDDU, L_NONE, j_max)
        }
    }

    // T 3456
    {
        const real h1 = 3.0 * j_max * (ad_ad + 2.0 * j_max * (a0 * tf - vd));
        const real h2 = ad_ad + 2.0 * j_max * (a0 * tf - vd);
        const real h0 = sqrt(4.0 * sq(2.0 * (a0_p3 - af_p3) - 6.0 * a0_a0 * (af - j_max * tf) + 6.0 * j_max_j_max * g1
                                        + 3.0 * a0 * (2.0 * af_af - 2.0 * j_max * af * tf + j_max_j_max * tf_tf) + 6.0 * ad * j_max * vd)
                               - 18.0 * h2 * h2 * h2)
                          / h1 * fabs(j_max) / j_max;
        t[0] = 0.0;
        t[1] = 0.0;
        t[2] = 0.0;
        t[3] = (af_p3 - a0_p3 + 3.0 * (af_af - a0_a0) * j_max * tf - 3.0 * ad * (a0 * af + 2.0 * j_max * vd) - 6.0 * j_max_j_max * g2) / h1;
        t[4] = (tf - t[3] - h0) * 0.5 - ad / twice(j_max);
        t[5] = h0;
        t[6] = (tf - t[3] + ad / j_max - h0) * 0.5;
        RK_TRY(UDDU, L_NONE, j_max)
    }

    // T 2346
    {
        const real ph1 = ad_ad + 2.0 * (af + a0) * j_max * tf - j_max * (j_max * tf_tf + 4.0 * vd);
        const real ph2 = j_max * tf_tf * g1 - vd * (-2.0 * pd - tf * v0 + 3.0 * tf * vf);
        const real ph3 = 5.0 * af_af - 8.0 * af * j_max * tf + 2.0 * j_max * (2.0 * j_max * tf_tf - vd);
        const real ph4 = j_max_j_max * tf_p4 - 2.0 * vd_vd + 8.0 * j_max * tf * (-pd + tf * vf);
        const real ph5 = 5.0 * af_p4 - 8.0 * af_p3 * j_max * tf - 12.0 * af_af * j_max * (j_max * tf_tf + vd)
                           + 24.0 * af * j_max_j_max * (-2.0 * pd + j_max * tf_p3 + 2.0 * tf * vf) - 6.0 * j_max_j_max * ph4;
        const real ph6 = -vd_vd + j_max * tf * (-2.0 * pd + 3.0 * tf * v0 - tf * vf) - af * g2;

        const real c0 = -(4.0 * (a0_p3 - af_p3) - 12.0 * a0_a0 * (af - j_max * tf) + 6.0 * a0 * (2.0 * af_af - 2.0 * af * j_max * tf + j_max * (j_max * tf_tf - 2.0 * vd))
                            + 6.0 * af * j_max * (3.0 * j_max * tf_tf + 2.0 * vd) - 6.0 * j_max_j_max * (-4.0 * pd + j_max * tf_p3 - 2.0 * tf * v0 + 6.0 * tf * vf))
                          / (3.0 * j_max * ph1);
        const real c1 = -(-a0_p4 - af_p4 + 4.0 * a0_p3 * (af - j_max * tf) + a0_a0 * (-6.0 * af_af + 8.0 * af * j_max * tf - 4.0 * j_max * (j_max * tf_tf - vd))
                            + 2.0 * af_af * j_max * (j_max * tf_tf + 2.0 * vd) - 4.0 * af * j_max_j_max * (-3.0 * pd + j_max * tf_p3 + 2.0 * tf * v0 + tf * vf)
                            + j_max_j_max * (j_max_j_max * tf_p4 - 8.0 * vd_vd + 4.0 * j_max * tf * (-3.0 * pd + tf * v0 + 2.0 * tf * vf))
                            + 2.0 * a0 * (2.0 * af_p3 - 2.0 * af_af * j_max * tf + af * j_max * (-3.0 * j_max * tf_tf - 4.0 * vd)
                                          + j_max_j_max * (-6.0 * pd + j_max * tf_p3 - 4.0 * tf * v0 + 10.0 * tf * vf)))
                          / (j_max_j_max * ph1);
        const real c2 = -(a0_p5 - af_p5 + af_p4 * j_max * tf - 5.0 * a0_p4 * (af - j_max * tf) + 2.0 * a0_p3 * ph3 + 4.0 * af_p3 * j_max * (j_max * tf_tf + vd)
                            + 12.0 * j_max_j_max * af * ph6
                            - 2.0 * a0_a0 * (5.0 * af_p3 - 9.0 * af_af * j_max * tf - 6.0 * af * j_max * vd + 6.0 * j_max_j_max * (-2.0 * pd - tf * v0 + 3.0 * tf * vf))
                            - 12.0 * j_max_j_max * j_max * ph2 + a0 * ph5)
                          / (3.0 * j_max_j_max * j_max * ph1);
        const real c3 = -(-a0_p6 - af_p6 + 6.0 * a0_p5 * (af - j_max * tf) - 48.0 * af_p3 * j_max_j_max * g1
                            + 72.0 * j_max_j_max * j_max * (j_max * g1 * g1 + vd_vd * vd + 2.0 * af * g1 * vd) - 3.0 * a0_p4 * ph3
                            - 36.0 * af_af * j_max_j_max * vd_vd + 6.0 * af_p4 * j_max * vd
                            + 4.0 * a0_p3 * (5.0 * af_p3 - 9.0 * af_af * j_max * tf - 6.0 * af * j_max * vd + 6.0 * j_max_j_max * (-2.0 * pd - tf * v0 + 3.0 * tf * vf))
                            - 3.0 * a0_a0 * ph5
                            + 6.0 * a0 * (af_p5 - af_p4 * j_max * tf - 4.0 * af_p3 * j_max * (j_max * tf_tf + vd) + 12.0 * j_max_j_max * (-af * ph6 + j_max * ph2)))
                          / (18.0 * j_max_j_max * j_max_j_max * ph1);

        const real t_max = (a0 - a_min) / j_max;
        const Roots roots = solve_quart_monic(c0, c1, c2, c3);
#pragma unroll 1
        for (int ri = 0; ri < roots.n; ++ri) {
            real tt = roots.get(ri);
            if (tt > t_max) continue;
            // Single Newton step (regarding pd)
            {
                const real h1 = ad_ad * 0.5 + j_max * (af * tt + (j_max * tt - a0) * (tt - tf) - vd);
                const real h2 = -ad + j_max * (tf - 2.0 * tt);
                const real h3 = sqrt(h1);
                const real orig = (af_p3 - a0_p3 + 3.0 * af * j_max * tt * (af + j_max * tt) + 3.0 * a0_a0 * (af + j_max * tt)
                                     - 3.0 * a0 * (af_af + 2.0 * af * j_max * tt + j_max_j_max * (tt * tt - tf_tf))
                                     + 3.0 * j_max_j_max * (-2.0 * pd + j_max * tt * (tt - tf) * tf + 2.0 * tf * v0))
                                        / (6.0 * j_max_j_max)
                                    - h3 * h3 * h3 / (j_max * fabs(j_max)) + ((-ad - j_max * tt) * h1) / (j_max_j_max);
                const real deriv = (6.0 * j_max * h2 * h3 / fabs(j_max) + 2.0 * (-ad - j_max * tf) * h2
                                      - 2.0 * (3.0 * ad_ad + af * j_max * (8.0 * tt - 2.0 * tf) + 4.0 * a0 * j_max * (-2.0 * tt + tf)
                                               + 2.0 * j_max * (j_max * tt * (3.0 * tt - 2.0 * tf) - vd)))
                                     / four_times(j_max);
                tt -= orig / deriv;
            }
            const real h1 = sqrt(2.0 * ad_ad + 4.0 * j_max * (ad * tt + a0 * tf + j_max * tt * (tt - tf) - vd)) / fabs(j_max);
            // Solution 2 with aPlat
            t[0] = 0.0;
            t[1] = 0.0;
            t[2] = tt;
            t[3] = tf - 2.0 * tt - ad / j_max - h1;
            t[4] = h1 * 0.5;
            t[5] = 0.0;
            t[6] = tf - (tt + t[3] + t[4]);
            RK_TRY(UDDU, L_NONE, j_max)
        }
    }

    // a3 != 0, UDUD — T 0124
    {
        const real ph0 = -2.0 * pd - tf * v0 + 3.0 * tf * vf;
        const real ph1 = -ad + j_max * tf;
        const real ph2 = j_max * tf_tf * g1 - vd * ph0;
        const real ph3 = 5.0 * af_af + 2.0 * j_max * (2.0 * j_max * tf_tf - vd - 4.0 * af * tf);
        const real ph4 = j_max_j_max * tf_p4 - 2.0 * vd_vd + 8.0 * j_max * tf * (-pd + tf * vf);
        const real ph5 = 5.0 * af_p4 - 8.0 * af_p3 * j_max * tf - 12.0 * af_af * j_max * (j_max * tf_tf + vd)
                           + 24.0 * af * j_max_j_max * (-2.0 * pd + j_max * tf_p3 + 2.0 * tf * vf) - 6.0 * j_max_j_max * ph4;
        const real ph6 = -vd_vd + j_max * tf * (-2.0 * pd + 3.0 * tf * v0 - tf * vf);
        const real ph7 = 3.0 * j_max_j_max * ph1 * ph1;

        const real c0 = (4.0 * af * tf - 2.0 * j_max * tf_tf - 4.0 * vd) / ph1;
        const real c1 = (-2.0 * (a0_p4 + af_p4) + 8.0 * af_p3 * j_max * tf + 6.0 * af_af * j_max_j_max * tf_tf + 8.0 * a0_p3 * (af - j_max * tf)
                           - 12.0 * a0_a0 * (af - j_max * tf) * (af - j_max * tf)
                           - 12.0 * af * j_max_j_max * (-pd + j_max * tf_p3 - 2.0 * tf * v0 + 3.0 * tf * vf)
                           + 2.0 * a0 * (4.0 * af_p3 - 12.0 * af_af * j_max * tf + 9.0 * af * j_max_j_max * tf_tf - 3.0 * j_max_j_max * (2.0 * pd + j_max * tf_p3 - 2.0 * tf * vf))
                           + 3.0 * j_max_j_max * (j_max_j_max * tf_p4 + 4.0 * vd_vd - 4.0 * j_max * tf * (pd + tf * v0 - 2.0 * tf * vf)))
                          / ph7;
        const real c2 = (-a0_p5 + af_p5 - af_p4 * j_max * tf + 5.0 * a0_p4 * (af - j_max * tf) - 2.0 * a0_p3 * ph3 - 4.0 * af_p3 * j_max * (j_max * tf_tf + vd)
                           + 12.0 * af_af * j_max_j_max * g2 - 12.0 * af * j_max_j_max * ph6
                           + 2.0 * a0_a0 * (5.0 * af_p3 - 9.0 * af_af * j_max * tf - 6.0 * af * j_max * vd + 6.0 * j_max_j_max * ph0)
                           + 12.0 * j_max_j_max * j_max * ph2
                           + a0 * (-5.0 * af_p4 + 8.0 * af_p3 * j_max * tf + 12.0 * af_af * j_max * (j_max * tf_tf + vd)
                                   - 24.0 * af * j_max_j_max * (-2.0 * pd + j_max * tf_p3 + 2.0 * tf * vf) + 6.0 * j_max_j_max * ph4))
                          / (j_max * ph7);
        const real c3 = -(a0_p6 + af_p6 - 6.0 * a0_p5 * (af - j_max * tf) + 48.0 * af_p3 * j_max_j_max * g1
                            - 72.0 * j_max_j_max * j_max * (j_max * g1 * g1 + vd_vd * vd + 2.0 * af * g1 * vd) + 3.0 * a0_p4 * ph3 - 6.0 * af_p4 * j_max * vd
                            + 36.0 * af_af * j_max_j_max * vd_vd
                            - 4.0 * a0_p3 * (5.0 * af_p3 - 9.0 * af_af * j_max * tf - 6.0 * af * j_max * vd + 6.0 * j_max_j_max * ph0) + 3.0 * a0_a0 * ph5
                            - 6.0 * a0 * (af_p5 - af_p4 * j_max * tf - 4.0 * af_p3 * j_max * (j_max * tf_tf + vd) + 12.0 * j_max_j_max * (af_af * g2 - af * ph6 + j_max * ph2)))
                          / (6.0 * j_max_j_max * ph7);
        (void)ph5;

        const Roots roots = solve_quart_monic(c0, c1, c2, c3);
#pragma unroll 1
        for (int ri = 0; ri < roots.n; ++ri) {
            const real tt = roots.get(ri);
            if (tt > tf || tt > (a_max - a0) / j_max) continue;
            const real h1 = sqrt(ad_ad / twice(j_max_j_max) + (a0 * (tt + tf) - af * tt + j_max * tt * tf - vd) / j_max);
            t[0] = tt;
            t[1] = tf - ad / j_max - 2.0 * h1;
            t[2] = h1;
            t[3] = 0.0;
            t[4] = ad / j_max + h1 - tt;
            t[5] = 0.0;
            t[6] = 0.0;
            RK_TRY(UDUD, L_NONE, j_max)
        }
    }

    // 3 step profile (UZD), T 012
    {
        const real h1 = sqrt(-ad_ad + j_max * (2.0 * (a0 + af) * tf - 4.0 * vd + j_max * tf_tf)) / fabs(j_max);
        t[0] = (tf - h1 + ad / j_max) * 0.5;
        t[1] = h1;
        t[2] = (tf - h1 - ad / j_max) * 0.5;
        t[3] = 0.0;
        t[4] = 0.0;
        t[5] = 0.0;
        t[6] = 0.0;
        RK_TRY(UDDU, L_NONE, j_max)
    }

    // 3 step profile (UZU); validated with j_max, not with its own jf (:2164-2171)
    {
        const real c0 = ad_ad;
        const real c1 = ad_ad * tf;
        const real c2 = (a0_a0 + af_af + 10.0 * a0 * af) * tf_tf + 24.0 * (tf * (af * v0 - a0 * vf) - pd * ad) + 12.0 * vd_vd;
        const real c3 = -3.0 * tf * ((a0_a0 + af_af + 2.0 * a0 * af) * tf_tf - 4.0 * vd * (a0 + af) * tf + 4.0 * vd_vd);
        const Roots roots = solve_cub(c0, c1, c2, c3);
#pragma unroll 1
        for (int ri = 0; ri < roots.n; ++ri) {
            const real tt = roots.get(ri);
            if (tt > tf) continue;
            const real jf = ad / (tf - tt);
            t[0] = (2.0 * (vd - a0 * tf) + ad * (tt - tf)) / (2.0 * jf * tt);
            t[1] = tt;
            t[2] = 0.0;
            t[3] = 0.0;
            t[4] = 0.0;
            t[5] = 0.0;
            t[6] = tf - (t[0] + t[1]);
            RK_TRY(UDDU, L_NONE, j_max)
        }
    }

    // 3 step profile (UDU)
    {
        t[0] = (ad_ad / j_max + 2.0 * (a0 + af) * tf - j_max * tf_tf - 4.0 * vd) / (4.0 * (ad - j_max * tf));
        t[1] = 0.0;
        t[2] = -ad / twice(j_max) + tf * 0.5;
        t[3] = 0.0;
        t[4] = 0.0;
        t[5] = 0.0;
        t[6] = tf - (t[0] + t[2]);
        RK_TRY(UDDU, L_NONE, j_max)
    }
    return false;
}

}  // namespace rk
#endif  // RK_H_STEP2_NONE_GO
