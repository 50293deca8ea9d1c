// rk_step2_acc.cuh — Step 2 families that do not reach the velocity limit and
// are closed form (no polynomial solve): ACC0_ACC1, ACC1, ACC0
// (position_third_step2.rs:976-1437). Some of them use a reduced jerk jf.
// Included once per pass: rk_kernels.cu compiles these headers twice — into namespace rk (exact divisions with their fall-back
// branch) and, with RK_PASS2 defined and `rk` renamed, into namespace rk_lazy (rk_math.cuh, RK_LAZY_DDIV).
#ifndef RK_PASS2
#ifndef RK_H_STEP2_ACC_1
#define RK_H_STEP2_ACC_1
#define RK_H_STEP2_ACC_GO
#endif
#else
#ifndef RK_H_STEP2_ACC_2
#define RK_H_STEP2_ACC_2
#define RK_H_STEP2_ACC_GO
#endif
#endif
#ifdef RK_H_STEP2_ACC_GO
#undef RK_H_STEP2_ACC_GO
#include "rk_step2.cuh"

namespace rk {

__device__ __forceinline__ double sq(double x) { return x * x; }  // roots.rs:14 pow2

// position_third_step2.rs:976-1079
__device__ RK_S2_FN bool s2_acc0_acc1(const S2& s, const Lim& L, Hit& h) {
    RK_S2_LOCALS
    if (fabs(a0) < EPS && fabs(af) < EPS) {
        const real h1 = 2.0 * a_min * g1 + vd_vd + a_max * (2.0 * pd + a_min * tf_tf - 2.0 * tf * vf);
        const real h2 = (a_max - a_min) * (-a_min * vd + a_max * (a_min * tf - vd));
        const real jf = h2 / h1;
        t[0] = a_max / jf;
        t[1] = (-2.0 * a_max * h1 + a_min * a_min * g2) / h2;
        t[2] = t[0];
        t[3] = 0.0;
        t[4] = -a_min / jf;
        t[5] = tf - (2.0 * t[0] + t[1] + 2.0 * t[4]);
        t[6] = t[4];
        RK_TRY(UDDU, L_ACC0_ACC1, jf)
        return false;
    }
    // UDDU
    {
        const real h1 = sqrt(144.0 * sq((a_max - a_min) * (-a_min * vd + a_max * (a_min * tf - vd)) - af_af * (a_max * tf - vd)
                                          + 2.0 * af * a_min * (a_max * tf - vd) + a0_a0 * (a_min * tf + v0 - vf) - 2.0 * a0 * a_max * (a_min * tf - vd))
                               + 48.0 * ad
                                     * (3.0 * a0_p3 - 3.0 * af_p3 + 12.0 * a_max * a_min * (-a_max + a_min) + 4.0 * af_af * (a_max + 2.0 * a_min)
                                        + a0 * (-3.0 * af_af + 8.0 * af * (a_min - a_max) + 6.0 * (a_max * a_max + 2.0 * a_max * a_min - a_min * a_min))
                                        + 6.0 * af * (a_max * a_max - 2.0 * a_max * a_min - a_min * a_min)
                                        + a0_a0 * (3.0 * af - 4.0 * (2.0 * a_max + a_min)))
                                     * (2.0 * a_min * g1 + vd * vd + a_max * (2.0 * pd + a_min * tf * tf - 2.0 * tf * vf)));
        const real jf = -(3.0 * af_af * a_max * tf - 3.0 * a0_a0 * a_min * tf - 6.0 * ad * a_max * a_min * tf + 3.0 * a_max * a_min * (a_min - a_max) * tf
                            + 3.0 * (a0_a0 - af_af) * vd + 6.0 * vd * (af * a_min - a0 * a_max) + 3.0 * (a_max * a_max - a_min * a_min) * vd + h1 * 0.25)
                          / (6.0 * (2.0 * a_min * g1 + vd * vd + a_max * (2.0 * pd + a_min * tf_tf - 2.0 * tf * vf)));
        t[0] = (a_max - a0) / jf;
        t[1] = (a0_a0 - af_af + 2.0 * ad * a_min - 2.0 * (a_max * a_max - 2.0 * a_max * a_min + a_min * a_min + a_min * jf * tf - jf * vd))
               / (2.0 * (a_max - a_min) * jf);
        t[2] = a_max / jf;
        t[3] = 0.0;
        t[4] = -a_min / jf;
        t[5] = tf - (t[0] + t[1] + t[2] + 2.0 * t[4] + af / jf);
        t[6] = t[4] + af / jf;
        RK_TRY(UDDU, L_ACC0_ACC1, jf)
    }
    return false;
}

// position_third_step2.rs:1081-1310
__device__ RK_S2_FN bool s2_acc1(const S2& s, const Lim& L, Hit& h) {
    RK_S2_LOCALS
    // a3 != 0, case UDDU
    {
        const real h0 = sqrt(j_max_j_max
                               * (a0_p4 + af_p4 - 4.0 * af_p3 * j_max * tf + 6.0 * af_af * j_max_j_max * tf_tf - 4.0 * a0_p3 * (af - j_max * tf)
                                  + 6.0 * a0_a0 * (af - j_max * tf) * (af - j_max * tf) + 24.0 * af * j_max_j_max * g1
                                  - 4.0 * a0 * (af_p3 - 3.0 * af_af * j_max * tf + 6.0 * j_max_j_max * (-pd + tf * vf))
                                  - 12.0 * j_max_j_max * (-vd_vd + j_max * tf * g2))
                               / THREE)
                          / j_max;
        const real h1 = sqrt((a0_a0 + af_af - 2.0 * a0 * af - 2.0 * ad * j_max * tf + 2.0 * h0) / j_max_j_max + tf_tf);
        t[0] = -(a0_a0 + af_af + 2.0 * a0 * (j_max * tf - af) - 2.0 * j_max * vd + h0) / (2.0 * j_max * (-ad + j_max * tf));
        t[1] = 0.0;
        t[2] = (tf - h1) * 0.5 - ad / twice(j_max);
        t[3] = 0.0;
        t[4] = 0.0;
        t[5] = h1;
        t[6] = tf - (t[0] + t[2] + t[5]);
        RK_TRY(UDDU, L_ACC1, j_max)
    }
    // case UDUD
    {
        const real h0 = sqrt(j_max_j_max
                               * (a0_p4 + af_p4 + 4.0 * (af_p3 - a0_p3) * j_max * tf + 6.0 * af_af * j_max_j_max * tf_tf
                                  + 6.0 * a0_a0 * (af + j_max * tf) * (af + j_max * tf) + 24.0 * af * j_max_j_max * g1
                                  - 4.0 * a0 * (a0_a0 * af + af_p3 + 3.0 * af_af * j_max * tf + 6.0 * j_max_j_max * (-pd + tf * vf))
                                  + 12.0 * j_max_j_max * (vd_vd + j_max * tf * g2))
                               / THREE)
                          / j_max;
        const real h1 = sqrt((a0_a0 + af_af - 2.0 * a0 * af + 2.0 * ad * j_max * tf + 2.0 * h0) / j_max_j_max + tf_tf);
        t[0] = 0.0;
        t[1] = 0.0;
        t[2] = -(a0_a0 + af_af - 2.0 * a0 * af + 2.0 * j_max * (vd - a0 * tf) + h0) / (2.0 * j_max * (ad + j_max * tf));
        t[3] = 0.0;
        t[4] = ad / twice(j_max) + (tf - h1) * 0.5;
        t[5] = h1;
        t[6] = tf - (t[5] + t[4] + t[2]);
        RK_TRY(UDUD, L_ACC1, j_max)
    }
    // case UDDU, solution 2
    {
        const real h0a = a0_p3 - af_p3 - 3.0 * a0_a0 * a_min + 3.0 * a_min * a_min * (a0 + j_max * tf) + 3.0 * af * a_min * (-a_min - 2.0 * j_max * tf)
                           - 3.0 * af_af * (-a_min - j_max * tf) - 3.0 * j_max_j_max * (-2.0 * pd - a_min * tf_tf + 2.0 * tf * vf);
        const real h0b = a0_a0 + af_af - 2.0 * (a0 + af) * a_min + 2.0 * (a_min * a_min - j_max * (-a_min * tf + vd));
        const real h0c = a0_p4 + 3.0 * af_p4 - 4.0 * (a0_p3 + 2.0 * af_p3) * a_min + 6.0 * a0_a0 * a_min * a_min
                           + 6.0 * af_af * (a_min * a_min - 2.0 * j_max * vd)
                           + 12.0 * j_max * (2.0 * a_min * j_max * g1 - a_min * a_min * vd + j_max * vd_vd) + 24.0 * af * a_min * j_max * vd
                           - 4.0 * a0
                                 * (af_p3 - 3.0 * af * a_min * (-a_min - 2.0 * j_max * tf) + 3.0 * af_af * (-a_min - j_max * tf)
                                    + 3.0 * j_max * (-a_min * a_min * tf + j_max * (-2.0 * pd - a_min * tf_tf + 2.0 * tf * vf)));
        const real h1 = fabs(j_max) / j_max * sqrt(4.0 * h0a * h0a - 6.0 * h0b * h0c);
        const real h2 = 6.0 * j_max * h0b;
        t[0] = 0.0;
        t[1] = 0.0;
        t[2] = (2.0 * h0a + h1) / h2;
        t[3] = -(a0_a0 + af_af - 2.0 * (a0 + af) * a_min + 2.0 * (a_min * a_min + a_min * j_max * tf - j_max * vd)) / (2.0 * j_max * (a0 - a_min - j_max * t[2]));
        t[4] = (a0 - a_min) / j_max - t[2];
        t[5] = tf - (t[2] + t[3] + t[4] + (af - a_min) / j_max);
        t[6] = (af - a_min) / j_max;
        RK_TRY(UDDU, L_ACC1, j_max)
    }
    // case UDUD, solution 1
    {
        const real h0a = -a0_p3 + af_p3 + 3.0 * (a0_a0 - af_af) * a_max - 3.0 * ad * a_max * a_max - 6.0 * af * a_max * j_max * tf + 3.0 * af_af * j_max * tf
                           + 3.0 * j_max * (a_max * a_max * tf + j_max * (-2.0 * pd - a_max * tf_tf + 2.0 * tf * vf));
        const real h0b = a0_a0 - af_af + 2.0 * ad * a_max + 2.0 * j_max * (a_max * tf - vd);
        const real h0c = a0_p4 + 3.0 * af_p4 - 4.0 * (a0_p3 + 2.0 * af_p3) * a_max + 6.0 * a0_a0 * a_max * a_max - 24.0 * af * a_max * j_max * vd
                           + 12.0 * j_max * (2.0 * a_max * j_max * g1 + j_max * vd_vd + a_max * a_max * vd) + 6.0 * af_af * (a_max * a_max + 2.0 * j_max * vd)
                           - 4.0 * a0
                                 * (af_p3 + 3.0 * af * a_max * (a_max - 2.0 * j_max * tf) - 3.0 * af_af * (a_max - j_max * tf)
                                    + 3.0 * j_max * (a_max * a_max * tf + j_max * (-2.0 * pd - a_max * tf_tf + 2.0 * tf * vf)));
        const real h1 = fabs(j_max) / j_max * sqrt(4.0 * h0a * h0a - 6.0 * h0b * h0c);
        const real h2 = 6.0 * j_max * h0b;
        t[0] = 0.0;
        t[1] = 0.0;
        t[2] = -(2.0 * h0a + h1) / h2;
        t[3] = 2.0 * h1 / h2;
        t[4] = (a_max - a0) / j_max + t[2];
        t[5] = tf - (t[2] + t[3] + t[4] + (-af + a_max) / j_max);
        t[6] = (-af + a_max) / j_max;
        RK_TRY(UDUD, L_ACC1, j_max)
    }
    return false;
}

// position_third_step2.rs:1312-1437
__device__ RK_S2_FN bool s2_acc0(const S2& s, const Lim& L, Hit& h) {
    RK_S2_LOCALS
    // UDUD
    {
        const real h1 = sqrt(ad_ad / twice(j_max_j_max) - ad * (a_max - a0) / (j_max_j_max) + (a_max * tf - vd) / j_max);
        t[0] = (a_max - a0) / j_max;
        t[1] = tf - ad / j_max - 2.0 * h1;
        t[2] = h1;
        t[3] = 0.0;
        t[4] = (af - a_max) / j_max + h1;
        t[5] = 0.0;
        t[6] = 0.0;
        RK_TRY(UDUD, L_NONE, j_max)
    }
    // UDUD
    {
        const real h0a = -a0_a0 + af_af - 2.0 * ad * a_max + 2.0 * j_max * (a_max * tf - vd);
        const real h0b = a0_p3 + 2.0 * af_p3 - 6.0 * af_af * a_max - 3.0 * a0_a0 * (af - j_max * tf) - 3.0 * a0 * a_max * (a_max - 2.0 * af + 2.0 * j_max * tf)
                           - 3.0 * j_max * (j_max * (-2.0 * pd + a_max * tf_tf + 2.0 * tf * v0) + a_max * (a_max * tf - 2.0 * vd))
                           + 3.0 * af * (a_max * a_max + 2.0 * a_max * j_max * tf - 2.0 * j_max * vd);
        const real h0 = fabs(j_max) * sqrt(4.0 * h0b * h0b - 18.0 * h0a * h0a * h0a);
        const real h1 = 3.0 * j_max * h0a;
        t[0] = (-a0 + a_max) / j_max;
        t[1] = (-a0_p3 + af_p3 + af_af * (-6.0 * a_max + 3.0 * j_max * tf) + a0_a0 * (-3.0 * af + 6.0 * a_max + 3.0 * j_max * tf)
                + 6.0 * af * (a_max * a_max - j_max * vd) + 3.0 * a0 * (af_af - 2.0 * (a_max * a_max + j_max * vd))
                - 6.0 * j_max * (a_max * (a_max * tf - 2.0 * vd) + j_max * g2))
               / h1;
        t[2] = -(ad + h0 / h1) / twice(j_max) + tf * 0.5 - t[1] * 0.5;
        t[3] = h0 / (j_max * h1);
        t[4] = 0.0;
        t[5] = 0.0;
        t[6] = tf - (t[0] + t[1] + t[2] + t[3]);
        RK_TRY(UDDU, L_NONE, j_max)
    }
    // a3 != 0, UDDU solution 1
    {
        const real h0a = a0_p3 + 2.0 * af_p3 - 6.0 * (af_af + a_max * a_max) * a_max - 6.0 * (a0 + af) * a_max * j_max * tf + 9.0 * a_max * a_max * (af + j_max * tf)
                           + 3.0 * a0 * a_max * (-2.0 * af + 3.0 * a_max) + 3.0 * a0_a0 * (af - 2.0 * a_max + j_max * tf) - 6.0 * j_max_j_max * g1
                           + 6.0 * (af - a_max) * j_max * vd - 3.0 * a_max * j_max_j_max * tf_tf;
        const real h0b = a0_a0 + af_af + 2.0 * (a_max * a_max - (a0 + af) * a_max + j_max * (vd - a_max * tf));
        const real h1 = fabs(j_max) / j_max * sqrt(4.0 * h0a * h0a - 18.0 * h0b * h0b * h0b);
        const real h2 = 6.0 * j_max * h0b;
        t[0] = (-a0 + a_max) / j_max;
        t[1] = ad / j_max - 2.0 * t[0] - (2.0 * h0a - h1) / h2 + tf;
        t[2] = -(2.0 * h0a + h1) / h2;
        t[3] = (2.0 * h0a - h1) / h2;
        t[4] = tf - (t[0] + t[1] + t[2] + t[3]);
        t[5] = 0.0;
        t[6] = 0.0;
        RK_TRY(UDDU, L_ACC0, j_max)
    }
    return false;
}

}  // namespace rk
#endif  // RK_H_STEP2_ACC_GO
