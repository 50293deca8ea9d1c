// rko_step2_pos3.hpp — parity oracle: Step 2 (fixed duration) for the
// third-order position interface.
//
// TEST INFRASTRUCTURE ONLY (see rko_math.hpp). Restates rsruckig v2.1.2
// lib/src/rsruckig/position_third_step2.rs; the family order in get_profile
// (:2449-2482) is semantics: the first candidate that passes Profile::check wins.
// time_none_smooth (:2204-2447) is dead in the reference (minimize_jerk is
// always false, :126) and is not restated.
#pragma once
#include "rko_core.hpp"

namespace rko {

struct PositionThirdOrderStep2 {
    double v0, a0, tf, vf, af, _v_max, _v_min, _a_max, _a_min, _j_max;
    double pd, tf_tf, tf_p3, tf_p4, vd, vd_vd, vf_vf, ad, ad_ad, a0_a0, af_af;
    double a0_p3, a0_p4, a0_p5, a0_p6, af_p3, af_p4, af_p5, af_p6, j_max_j_max, g1, g2;

    // :51-128
    PositionThirdOrderStep2(double tf_, double p0, double v0_, double a0_, double pf, double vf_, double af_, double v_max, double v_min, double a_max, double a_min, double j_max)
        : v0(v0_), a0(a0_), tf(tf_), vf(vf_), af(af_), _v_max(v_max), _v_min(v_min), _a_max(a_max), _a_min(a_min), _j_max(j_max) {
        pd = pf - p0;
        tf_tf = tf * tf;
        tf_p3 = tf_tf * tf;
        tf_p4 = tf_tf * tf_tf;

        vd = vf - v0;
        vd_vd = vd * vd;
        vf_vf = vf * vf;

        ad = af - a0;
        ad_ad = ad * ad;
        a0_a0 = a0 * a0;
        af_af = af * af;

        a0_p3 = a0 * a0_a0;
        a0_p4 = a0_a0 * a0_a0;
        a0_p5 = a0_p3 * a0_a0;
        a0_p6 = a0_p4 * a0_a0;
        af_p3 = af * af_af;
        af_p4 = af_af * af_af;
        af_p5 = af_p3 * af_af;
        af_p6 = af_p4 * af_af;

        j_max_j_max = j_max * j_max;

        g1 = -pd + tf * v0;
        g2 = -2.0 * pd + tf * (v0 + vf);
    }

    static double pow2(double x) { return x * x; }

    // :130-245
    bool time_acc0_acc1_vel(Profile& profile, double v_max, double v_min, double a_max, double a_min, double j_max) {
        RKO_PHASE(RKO_PH_S2_ACC0_ACC1_VEL);
        // Profile UDDU, Solution 1
        if ((2.0 * (a_max - a_min) + ad) / j_max < tf) {
            const double h1 = std::sqrt(
                (a0_p4 + af_p4
                    - 4.0 * a0_p3 * (2.0 * a_max + a_min) / 3.0
                    - 4.0 * af_p3 * (a_max + 2.0 * a_min) / 3.0
                    + 2.0 * (a0_a0 - af_af) * a_max * a_max
                    + (4.0 * a0 * a_max - 2.0 * a0_a0)
                    * (af_af - 2.0 * af * a_min
                    + (a_min - a_max) * a_min
                    + 2.0 * j_max * (a_min * tf - vd))
                    + 2.0
                    * af_af
                    * (a_min * a_min + 2.0 * j_max * (a_max * tf - vd))
                    + 4.0
                    * j_max
                    * (2.0 * a_min * (af * vd + j_max * g1)
                    + (a_max * a_max - a_min * a_min) * vd
                    + j_max * vd_vd)
                    + 8.0 * a_max * j_max_j_max * (pd - tf * vf))
                    / (a_max * a_min)
                    + 4.0 * af_af
                    + 2.0 * a0_a0
                    + (4.0 * af + a_max - a_min) * (a_max - a_min)
                    + 4.0 * j_max * (a_min - a_max + j_max * tf - 2.0 * af) * tf
            ) * std::fabs(j_max)
                / j_max;
            profile.t[0] = (-a0 + a_max) / j_max;
            profile.t[1] = (-(af_af - a0_a0
                + 2.0 * a_max * a_max
                + a_min * (a_min - 2.0 * ad - 3.0 * a_max)
                + 2.0 * j_max * (a_min * tf - vd))
                + a_min * h1)
                / (2.0 * (a_max - a_min) * j_max);
            profile.t[2] = a_max / j_max;
            profile.t[3] = (a_min - a_max + h1) / (2.0 * j_max);
            profile.t[4] = -a_min / j_max;
            profile.t[5] = tf
                - (profile.t[0]
                + profile.t[1]
                + profile.t[2]
                + profile.t[3]
                + 2.0 * profile.t[4]
                + af / j_max);
            profile.t[6] = profile.t[4] + af / j_max;

            if (profile.check_with_timing(UDDU, ACC0_ACC1_VEL, j_max, v_max, v_min, a_max, a_min)) return true;
        }

        // Profile UDUD
        if ((-a0 + 4.0 * a_max - af) / j_max < tf) {
            profile.t[0] = (-a0 + a_max) / j_max;
            profile.t[1] = (3.0 * (a0_p4 + af_p4)
                - 4.0 * (a0_p3 + af_p3) * a_max
                - 4.0 * af_p3 * a_max
                + 24.0 * (a0 + af) * a_max * a_max * a_max
                - 6.0 * (af_af + a0_a0) * (a_max * a_max - 2.0 * j_max * vd)
                + 6.0
                * a0_a0
                * (af_af - 2.0 * af * a_max - 2.0 * a_max * j_max * tf)
                - 12.0
                * a_max
                * a_max
                * (2.0 * a_max * a_max - 2.0 * a_max * j_max * tf + j_max * vd)
                - 24.0 * af * a_max * j_max * vd
                + 12.0 * j_max_j_max * (2.0 * a_max * g1 + vd_vd))
                / (12.0
                * a_max
                * j_max
                * (a0_a0 + af_af - 2.0 * (a0 + af) * a_max
                + 2.0 * (a_max * a_max - a_max * j_max * tf + j_max * vd)));
            profile.t[2] = a_max / j_max;
            profile.t[3] = (-a0_a0 - af_af
                + 2.0 * a_max * (a0 + af - 2.0 * a_max)
                - 2.0 * j_max * vd)
                / (2.0 * a_max * j_max)
                + tf;
            profile.t[4] = profile.t[2];
            profile.t[5] = tf
                - (profile.t[0] + profile.t[1] + profile.t[2] + profile.t[3] + 2.0 * profile.t[4]
                - af / j_max);
            profile.t[6] = profile.t[4] - af / j_max;

            if (profile.check_with_timing(UDUD, ACC0_ACC1_VEL, j_max, v_max, v_min, a_max, a_min)) return true;
        }
        return false;
    }

    // :247-414
    bool time_acc1_vel(Profile& profile, double v_max, double v_min, double a_max, double a_min, double j_max) {
        RKO_PHASE(RKO_PH_S2_ACC1_VEL);
        // Profile UDDU
        {
            const double ph1 = a0_a0 + af_af
                - a_min * (a0 + 2.0 * af - a_min)
                - 2.0 * j_max * (vd - a_min * tf);
            const double ph2 = 2.0 * a_min * (j_max * g1 + af * vd) - a_min * a_min * vd
                + j_max * vd_vd;
            const double ph3 = af_af + a_min * (a_min - 2.0 * af)
                - 2.0 * j_max * (vd - a_min * tf);
            double polynom[4];

            polynom[0] = (2.0 * (2.0 * a0 - a_min)) / j_max;
            polynom[1] = (4.0 * a0_a0 + ph1 - 3.0 * a0 * a_min) / j_max_j_max;
            polynom[2] = (2.0 * a0 * ph1) / (j_max_j_max * j_max);
            polynom[3] = (3.0 * (a0_p4 + af_p4)
                - 4.0 * (a0_p3 + 2.0 * af_p3) * a_min
                + 6.0 * af_af * (a_min * a_min - 2.0 * j_max * vd)
                + 12.0 * j_max * ph2
                + 6.0 * a0_a0 * ph3)
                / (12.0 * j_max_j_max * j_max_j_max);

            const double t_min = -a0 / j_max;
            const double t_max = fmin_(
                (tf + 2.0 * a_min / j_max - (a0 + af) / j_max) / 2.0,
                (a_max - a0) / j_max
            );

            roots::PositiveSet<4> rts = roots::solve_quart_monic(polynom[0], polynom[1], polynom[2], polynom[3]);
            rts.sort();
            for (int ri = 0; ri < rts.size; ++ri) {
                double t = rts.data[ri];
                if (t < t_min || t > t_max) continue;

                // Single Newton step (regarding pd)
                if (std::fabs(a0 + j_max * t) > 16.0 * EPS) {
                    const double h0 = j_max * t * t;
                    const double orig = -pd
                        + (3.0 * (a0_p4 + af_p4)
                        - 8.0 * af_p3 * a_min
                        - 4.0 * a0_p3 * a_min
                        + 6.0 * af_af * (a_min * a_min + 2.0 * j_max * (h0 - vd))
                        + 6.0
                        * a0_a0
                        * (af_af - 2.0 * af * a_min
                        + a_min * a_min
                        + 2.0 * a_min * j_max * (-2.0 * t + tf)
                        + 2.0 * j_max * (5.0 * h0 - vd))
                        + 24.0
                        * a0
                        * j_max
                        * t
                        * (a0_a0 + af_af - 2.0 * af * a_min
                        + a_min * a_min
                        + 2.0 * j_max * (a_min * (-t + tf) + h0 - vd))
                        - 24.0 * af * a_min * j_max * (h0 - vd)
                        + 12.0
                        * j_max
                        * (a_min * a_min * (h0 - vd)
                        + j_max * (h0 - vd) * (h0 - vd)))
                        / (24.0 * a_min * j_max_j_max)
                        + h0 * (tf - t)
                        + tf * v0;
                    const double deriv = (a0 + j_max * t)
                        * ((a0_a0 + af_af) / (a_min * j_max)
                        + (a_min - a0 - 2.0 * af) / j_max
                        + (4.0 * a0 * t + 2.0 * h0 - 2.0 * vd) / a_min
                        + 2.0 * tf
                        - 3.0 * t);

                    t -= orig / deriv;
                }

                const double h1 = -((a0_a0 + af_af) / 2.0
                    + j_max * (-vd + 2.0 * a0 * t + j_max * t * t))
                    / a_min;

                profile.t[0] = t;
                profile.t[1] = 0.0;
                profile.t[2] = a0 / j_max + t;
                profile.t[3] = tf - (h1 - a_min + a0 + af) / j_max - 2.0 * t;
                profile.t[4] = -a_min / j_max;
                profile.t[5] = (h1 + a_min) / j_max;
                profile.t[6] = profile.t[4] + af / j_max;

                if (profile.check_with_timing(UDDU, ACC1_VEL, j_max, v_max, v_min, a_max, a_min)) return true;
            }
        }

        // Profile UDUD
        {
            const double ph1 = a0_a0 - af_af + (2.0 * af - a0) * a_max
                - a_max * a_max
                - 2.0 * j_max * (vd - a_max * tf);
            const double ph2 = a_max * a_max + 2.0 * j_max * vd;
            const double ph3 = af_af + ph2 - 2.0 * a_max * (af + j_max * tf);
            const double ph4 = 2.0 * a_max * j_max * g1 + a_max * a_max * vd + j_max * vd_vd;

            double polynom[4];
            polynom[0] = (4.0 * a0 - 2.0 * a_max) / j_max;
            polynom[1] = (4.0 * a0_a0 - 3.0 * a0 * a_max + ph1) / j_max_j_max;
            polynom[2] = (2.0 * a0 * ph1) / (j_max_j_max * j_max);
            polynom[3] = (3.0 * (a0_p4 + af_p4)
                - 4.0 * (a0_p3 + 2.0 * af_p3) * a_max
                - 24.0 * af * a_max * j_max * vd
                + 12.0 * j_max * ph4
                - 6.0 * a0_a0 * ph3
                + 6.0 * af_af * ph2)
                / (12.0 * j_max_j_max * j_max_j_max);

            const double t_min = -a0 / j_max;
            const double t_max = fmin_(
                (tf + ad / j_max - 2.0 * a_max / j_max) / 2.0,
                (a_max - a0) / j_max
            );

            roots::PositiveSet<4> rts = roots::solve_quart_monic(polynom[0], polynom[1], polynom[2], polynom[3]);
            rts.sort();
            for (int ri = 0; ri < rts.size; ++ri) {
                const double t = rts.data[ri];
                if (t > t_max || t < t_min) continue;

                const double h1 = ((a0_a0 - af_af) / 2.0 + j_max_j_max * t * t
                    - j_max * (vd - 2.0 * a0 * t))
                    / a_max;

                profile.t[0] = t;
                profile.t[1] = 0.0;
                profile.t[2] = t + a0 / j_max;
                profile.t[3] = tf + (h1 + ad - a_max) / j_max - 2.0 * t;
                profile.t[4] = a_max / j_max;
                profile.t[5] = -(h1 + a_max) / j_max;
                profile.t[6] = profile.t[4] - af / j_max;

                if (profile.check_with_timing(UDUD, ACC1_VEL, j_max, v_max, v_min, a_max, a_min)) return true;
            }
        }

        return false;
    }

    // :416-605
    bool time_acc0_vel(Profile& profile, double v_max, double v_min, double a_max, double a_min, double j_max) {
        RKO_PHASE(RKO_PH_S2_ACC0_VEL);
        if (tf < fmax_((-a0 + a_max) / j_max, 0.0) + fmax_(a_max / j_max, 0.0)) return false;

        const double ph1 = 12.0
            * j_max
            * (-a_max * a_max * vd - j_max * vd_vd
            + 2.0 * a_max * j_max * (-pd + tf * vf));

        // Profile UDDU
        {
            double polynom[4];
            polynom[0] = (2.0 * a_max) / j_max;
            polynom[1] = (a0_a0 - af_af
                + 2.0 * ad * a_max
                + a_max * a_max
                + 2.0 * j_max * (vd - a_max * tf))
                / j_max_j_max;
            polynom[2] = 0.0;
            polynom[3] = -(-3.0 * (a0_p4 + af_p4)
                + 4.0 * (af_p3 + 2.0 * a0_p3) * a_max
                - 12.0 * a0 * a_max * (af_af - 2.0 * j_max * vd)
                + 6.0 * a0_a0 * (af_af - a_max * a_max - 2.0 * j_max * vd)
                + 6.0
                * af_af
                * (a_max * a_max - 2.0 * a_max * j_max * tf + 2.0 * j_max * vd)
                + ph1)
                / (12.0 * j_max_j_max * j_max_j_max);

            const double t_min = -af / j_max;
            const double t_max = fmin_(tf - (2.0 * a_max - a0) / j_max, -a_min / j_max);
            roots::PositiveSet<4> rts = roots::solve_quart_monic(polynom[0], polynom[1], polynom[2], polynom[3]);
            rts.sort();
            for (int ri = 0; ri < rts.size; ++ri) {
                double t = rts.data[ri];
                if (t < t_min || t > t_max) continue;

                // Single Newton step (regarding pd)
                if (t > EPS) {
                    const double h1 = j_max * t * t + vd;
                    const double orig = (-3.0 * (a0_p4 + af_p4)
                        + 4.0 * (af_p3 + 2.0 * a0_p3) * a_max
                        - 24.0 * af * a_max * j_max_j_max * t * t
                        - 12.0 * a0 * a_max * (af_af - 2.0 * j_max * h1)
                        + 6.0 * a0_a0 * (af_af - a_max * a_max - 2.0 * j_max * h1)
                        + 6.0
                        * af_af
                        * (a_max * a_max - 2.0 * a_max * j_max * tf + 2.0 * j_max * h1)
                        - 12.0
                        * j_max
                        * (a_max * a_max * h1
                        + j_max * h1 * h1
                        + 2.0
                        * a_max
                        * j_max
                        * (pd + j_max * t * t * (t - tf)
                        - tf * vf)))
                        / (24.0 * a_max * j_max_j_max);
                    const double deriv = -t
                        * (a0_a0 - af_af
                        + 2.0 * a_max * (ad - j_max * tf)
                        + a_max * a_max
                        + 3.0 * a_max * j_max * t
                        + 2.0 * j_max * h1)
                        / a_max;

                    t -= orig / deriv;
                }

                const double h1 = ((a0_a0 - af_af) / 2.0 + j_max * (j_max * t * t + vd)) / a_max;

                profile.t[0] = (-a0 + a_max) / j_max;
                profile.t[1] = (h1 - a_max) / j_max;
                profile.t[2] = a_max / j_max;
                profile.t[3] = tf - (h1 + ad + a_max) / j_max - 2.0 * t;
                profile.t[4] = t;
                profile.t[5] = 0.0;
                profile.t[6] = af / j_max + t;

                if (profile.check_with_timing(UDDU, ACC0_VEL, j_max, v_max, v_min, a_max, a_min)) return true;
            }
        }

        // Profile UDUD
        {
            double polynom[4];
            polynom[0] = (-2.0 * a_max) / j_max;
            polynom[1] = -(a0_a0 + af_af - 2.0 * (a0 + af) * a_max
                + a_max * a_max
                + 2.0 * j_max * (vd - a_max * tf))
                / j_max_j_max;
            polynom[2] = 0.0;
            polynom[3] = (3.0 * (a0_p4 + af_p4)
                - 4.0 * (af_p3 + 2.0 * a0_p3) * a_max
                + 6.0 * a0_a0 * (af_af + a_max * a_max + 2.0 * j_max * vd)
                - 12.0 * a0 * a_max * (af_af + 2.0 * j_max * vd)
                + 6.0
                * af_af
                * (a_max * a_max - 2.0 * a_max * j_max * tf + 2.0 * j_max * vd)
                - ph1)
                / (12.0 * j_max_j_max * j_max_j_max);

            const double t_min = af / j_max;
            const double t_max = fmin_(tf - a_max / j_max, a_max / j_max);

            roots::PositiveSet<4> rts = roots::solve_quart_monic(polynom[0], polynom[1], polynom[2], polynom[3]);
            rts.sort();
            for (int ri = 0; ri < rts.size; ++ri) {
                double t = rts.data[ri];
                if (t < t_min || t > t_max) continue;

                // Single Newton step (regarding pd)
                {
                    const double h1 = j_max * t * t - vd;
                    const double orig = -(3.0 * (a0_p4 + af_p4)
                        - 4.0 * (2.0 * a0_p3 + af_p3) * a_max
                        + 24.0 * af * a_max * j_max_j_max * t * t
                        - 12.0 * a0 * a_max * (af_af - 2.0 * j_max * h1)
                        + 6.0 * a0_a0 * (af_af + a_max * a_max - 2.0 * j_max * h1)
                        + 6.0
                        * af_af
                        * (a_max * a_max - 2.0 * j_max * (tf * a_max + h1))
                        + 12.0
                        * j_max
                        * (-a_max * a_max * h1 + j_max * h1 * h1
                        - 2.0
                        * a_max
                        * j_max
                        * (-pd
                        + j_max * t * t * (t - tf)
                        + tf * vf)))
                        / (24.0 * a_max * j_max_j_max);
                    const double deriv = t
                        * (a0_a0 + af_af
                        - 2.0 * j_max * h1
                        - 2.0 * (a0 + af + j_max * tf) * a_max
                        + a_max * a_max
                        + 3.0 * a_max * j_max * t)
                        / a_max;

                    t -= orig / deriv;
                }

                const double h1 = ((a0_a0 + af_af) / 2.0 + j_max * (vd - j_max * t * t)) / a_max;

                profile.t[0] = (-a0 + a_max) / j_max;
                profile.t[1] = (h1 - a_max) / j_max;
                profile.t[2] = a_max / j_max;
                profile.t[3] = tf - (h1 - a0 - af + a_max) / j_max - 2.0 * t;
                profile.t[4] = t;
                profile.t[5] = 0.0;
                profile.t[6] = -(af / j_max) + t;

                if (profile.check_with_timing(UDUD, ACC0_VEL, j_max, v_max, v_min, a_max, a_min)) return true;
            }
        }

        return false;
    }

    // :723-778 (closure check_root of the UDDU branch of time_vel)
    bool vel_uddu_check_root(Profile& profile, double t, double v_max, double v_min, double a_max, double a_min, double j_max) {
        // Single Newton step (regarding pd)
        {
            const double h1 = std::sqrt(
                (a0_a0 + af_af) / (2.0 * j_max_j_max)
                    + (2.0 * a0 * t + j_max * t * t - vd) / j_max
            );
            const double orig = -pd
                - (2.0 * a0_p3
                + 4.0 * af_p3
                + 24.0 * a0 * j_max * t * (af + j_max * (h1 + t - tf))
                + 6.0 * a0_a0 * (af + j_max * (2.0 * t - tf))
                + 6.0 * (a0_a0 + af_af) * j_max * h1
                + 12.0 * af * j_max * (j_max * t * t - vd)
                + 12.0
                * j_max_j_max
                * (j_max * t * t * (h1 + t - tf)
                - tf * v0
                - h1 * vd))
                / (12.0 * j_max_j_max);
            const double deriv_newton = -(a0 + j_max * t)
                * (3.0 * (h1 + t) - 2.0 * tf + (a0 + 2.0 * af) / j_max);
            if (!(orig != orig) && !(deriv_newton != deriv_newton) && std::fabs(deriv_newton) > EPS) {
                t -= orig / deriv_newton;
            }
        }

        if (t > tf || t != t) return false;

        const double h1 = std::sqrt(
            (a0_a0 + af_af) / (2.0 * j_max_j_max)
                + (t * (2.0 * a0 + j_max * t) - vd) / j_max
        );
        profile.t[0] = t;
        profile.t[1] = 0.0;
        profile.t[2] = t + a0 / j_max;
        profile.t[3] = tf - 2.0 * (t + h1) - (a0 + af) / j_max;
        profile.t[4] = h1;
        profile.t[5] = 0.0;
        profile.t[6] = h1 + af / j_max;

        return profile.check_with_timing(UDDU, VEL, j_max, v_max, v_min, a_max, a_min);
    }

    // :886-949 (closure check_root of the UDUD branch of time_vel)
    bool vel_udud_check_root(Profile& profile, double t, double v_max, double v_min, double a_max, double a_min, double j_max) {
        // Double Newton step (regarding pd)
        {
            double h1 = std::sqrt(
                (af_af - a0_a0) / (2.0 * j_max_j_max)
                    - ((2.0 * a0 + j_max * t) * t - vd) / j_max
            );
            double orig = -pd
                + (af_p3 - a0_p3
                + 3.0 * a0_a0 * j_max * (tf - 2.0 * t))
                / (6.0 * j_max_j_max)
                + (2.0 * a0 + j_max * t) * t * (tf - t)
                + (j_max * h1 - af) * h1 * h1
                + tf * v0;
            double deriv_newton = (a0 + j_max * t)
                * (2.0 * (af + j_max * tf) - 3.0 * j_max * (h1 + t) - a0)
                / j_max;

            t -= orig / deriv_newton;

            h1 = std::sqrt(
                (af_af - a0_a0) / (2.0 * j_max_j_max)
                    - ((2.0 * a0 + j_max * t) * t - vd) / j_max
            );
            orig = -pd
                + (af_p3 - a0_p3
                + 3.0 * a0_a0 * j_max * (tf - 2.0 * t))
                / (6.0 * j_max_j_max)
                + (2.0 * a0 + j_max * t) * t * (tf - t)
                + (j_max * h1 - af) * h1 * h1
                + tf * v0;
            if (std::fabs(orig) > 1e-9) {
                deriv_newton = (a0 + j_max * t)
                    * (2.0 * (af + j_max * tf)
                    - 3.0 * j_max * (h1 + t)
                    - a0)
                    / j_max;

                t -= orig / deriv_newton;
            }
        }

        const double h1 = std::sqrt(
            (af_af - a0_a0) / (2.0 * j_max_j_max)
                - ((2.0 * a0 + j_max * t) * t - vd) / j_max
        );
        profile.t[0] = t;
        profile.t[1] = 0.0;
        profile.t[2] = t + a0 / j_max;
        profile.t[3] = tf - 2.0 * (t + h1) + ad / j_max;
        profile.t[4] = h1;
        profile.t[5] = 0.0;
        profile.t[6] = h1 - af / j_max;

        return profile.check_with_timing(UDUD, VEL, j_max, v_max, v_min, a_max, a_min);
    }

    // :607-974
    bool time_vel(Profile& profile, double v_max, double v_min, double a_max, double a_min, double j_max) {
        RKO_PHASE(RKO_PH_S2_VEL);
        using roots::Poly; using roots::poly_eval; using roots::poly_deri; using roots::poly_monic_deri;
        using roots::shrink_interval; using roots::TOLERANCE;
        const double tz_min = fmax_(0.0, -a0 / j_max);
        const double tz_max = fmin_((tf - a0 / j_max) / 2.0, (a_max - a0) / j_max);

        // Profile UDDU
        if (std::fabs(v0) < EPS && std::fabs(a0) < EPS && std::fabs(vf) < EPS && std::fabs(af) < EPS) {
            double polynom[4];
            polynom[0] = 1.0;
            polynom[1] = -tf / 2.0;
            polynom[2] = 0.0;
            polynom[3] = pd / (2.0 * j_max);

            roots::PositiveSet<3> rts = roots::solve_cub(polynom[0], polynom[1], polynom[2], polynom[3]);
            rts.sort();
            for (int ri = 0; ri < rts.size; ++ri) {
                double t = rts.data[ri];
                if (t > tf / 4.0) continue;

                // Single Newton step (regarding pd)
                if (t > EPS) {
                    const double orig = -pd + j_max * t * t * (tf - 2.0 * t);
                    const double deriv = 2.0 * j_max * t * (tf - 3.0 * t);
                    t -= orig / deriv;
                }

                profile.t[0] = t;
                profile.t[1] = 0.0;
                profile.t[2] = t;
                profile.t[3] = tf - 4.0 * t;
                profile.t[4] = t;
                profile.t[5] = 0.0;
                profile.t[6] = t;

                if (profile.check_with_timing(UDDU, VEL, j_max, v_max, v_min, a_max, a_min)) return true;
            }
        } else {
            const double p1 = af_af
                - 2.0 * j_max * (-2.0 * af * tf + j_max * tf_tf + 3.0 * vd);
            const double ph1 = af_p3 - 3.0 * j_max_j_max * g1 - 3.0 * af * j_max * vd;
            const double ph2 = af_p4
                + 8.0 * af_p3 * j_max * tf
                + 12.0
                * j_max
                * (3.0 * j_max * vd_vd - af_af * vd
                + 2.0 * af * j_max * (g1 - tf * vd)
                - 2.0 * j_max_j_max * tf * g1);
            const double ph3 = a0 * (af - j_max * tf);
            const double ph4 = j_max * (-ad + j_max * tf);

            // Find root of 5th order polynom
            Poly polynom;
            polynom.push(1.0);
            polynom.push((15.0 * a0_a0 + af_af + 4.0 * af * j_max * tf
                - 16.0 * ph3
                - 2.0 * j_max * (j_max * tf_tf + 3.0 * vd))
                / (4.0 * ph4));
            polynom.push((29.0 * a0_p3 - 2.0 * af_p3 - 33.0 * a0 * ph3
                + 6.0 * j_max_j_max * g1
                + 6.0 * af * j_max * vd
                + 6.0 * a0 * p1)
                / (6.0 * j_max * ph4));
            polynom.push((61.0 * a0_p4 - 76.0 * a0_a0 * ph3 - 16.0 * a0 * ph1
                + 30.0 * a0_a0 * p1
                + ph2)
                / (24.0 * j_max_j_max * ph4));
            polynom.push((a0
                * (7.0 * a0_p4 - 10.0 * a0_a0 * ph3 - 4.0 * a0 * ph1
                + 6.0 * a0_a0 * p1
                + ph2))
                / (12.0 * j_max_j_max * j_max * ph4));
            polynom.push((7.0 * a0_p6 + af_p6 - 12.0 * a0_p4 * ph3
                + 48.0 * af_p3 * j_max_j_max * g1
                - 8.0 * a0_p3 * ph1
                - 72.0
                * j_max_j_max
                * j_max
                * (j_max * g1 * g1
                + vd_vd * vd
                + 2.0 * af * g1 * vd)
                - 6.0 * af_p4 * j_max * vd
                + 36.0 * af_af * j_max_j_max * vd_vd
                + 9.0 * a0_p4 * p1
                + 3.0 * a0_a0 * ph2)
                / (144.0 * j_max_j_max * j_max_j_max * ph4));

            const Poly deriv = poly_monic_deri(polynom);
            const Poly dderiv = poly_deri(deriv);

            // Solve 4th order derivative analytically
            roots::PositiveSet<4> d_extremas = roots::solve_quart_monic(deriv[1], deriv[2], deriv[3], deriv[4]);

            double tz_current = tz_min;

            d_extremas.sort();
            for (int ri = 0; ri < d_extremas.size; ++ri) {
                double tz = d_extremas.data[ri];
                if (tz >= tz_max) continue;

                const double orig = poly_eval(deriv, tz);
                if (std::fabs(orig) > TOLERANCE) {
                    tz -= orig / poly_eval(dderiv, tz);
                }

                const double val_new = poly_eval(polynom, tz);
                if (std::fabs(val_new) < 64.0 * std::fabs(poly_eval(dderiv, tz)) * TOLERANCE) {
                    if (vel_uddu_check_root(profile, tz, v_max, v_min, a_max, a_min, j_max)) return true;
                } else if (poly_eval(polynom, tz_current) * val_new < 0.0
                           && vel_uddu_check_root(profile, shrink_interval(polynom, tz_current, tz), v_max, v_min, a_max, a_min, j_max)) {
                    return true;
                }
                tz_current = tz;
            }
            const double val_max = poly_eval(polynom, tz_max);
            if (poly_eval(polynom, tz_current) * val_max < 0.0) {
                if (vel_uddu_check_root(profile, shrink_interval(polynom, tz_current, tz_max), v_max, v_min, a_max, a_min, j_max)) return true;
            } else if (std::fabs(val_max) < 8.0 * EPS && vel_uddu_check_root(profile, tz_max, v_max, v_min, a_max, a_min, j_max)) {
                return true;
            }
        }

        // Profile UDUD
        {
            const double ph1 = af_af
                - 2.0 * j_max * (2.0 * af * tf + j_max * tf_tf - 3.0 * vd);
            const double ph2 = af_p3 - 3.0 * j_max_j_max * g1 + 3.0 * af * j_max * vd;
            const double ph3 = 2.0 * j_max * tf * g1 + 3.0 * vd_vd;
            const double ph4 = af_p4 - 8.0 * af_p3 * j_max * tf
                + 12.0
                * j_max
                * (j_max * ph3
                + af_af * vd
                + 2.0 * af * j_max * (g1 - tf * vd));
            const double ph5 = af + j_max * tf;

            // Find root of 6th order polynom
            Poly polynom;
            polynom.push(1.0);
            polynom.push((5.0 * a0 - ph5) / j_max);
            polynom.push((39.0 * a0_a0 - ph1 - 16.0 * a0 * ph5) / (4.0 * j_max_j_max));
            polynom.push((55.0 * a0_p3 - 33.0 * a0_a0 * ph5 - 6.0 * a0 * ph1
                + 2.0 * ph2)
                / (6.0 * j_max_j_max * j_max));
            polynom.push((101.0 * a0_p4 + ph4 - 76.0 * a0_p3 * ph5 - 30.0 * a0_a0 * ph1
                + 16.0 * a0 * ph2)
                / (24.0 * j_max_j_max * j_max_j_max));
            polynom.push((a0
                * (11.0 * a0_p4 + ph4 - 10.0 * a0_p3 * ph5 - 6.0 * a0_a0 * ph1
                + 4.0 * a0 * ph2))
                / (12.0 * j_max_j_max * j_max_j_max * j_max));
            polynom.push((11.0 * a0_p6
                - af_p6
                - 12.0 * a0_p5 * ph5
                - 48.0 * af_p3 * j_max_j_max * g1
                - 9.0 * a0_p4 * ph1
                + 72.0
                * j_max_j_max
                * j_max
                * (j_max * g1 * g1
                - vd_vd * vd
                - 2.0 * af * g1 * vd)
                - 6.0 * af_p4 * j_max * vd
                - 36.0 * af_af * j_max_j_max * vd_vd
                + 8.0 * a0_p3 * ph2
                + 3.0 * a0_a0 * ph4)
                / (144.0 * j_max_j_max * j_max_j_max * j_max_j_max));

            const Poly deriv = poly_monic_deri(polynom);
            const Poly dderiv = poly_monic_deri(deriv);

            double dd_tz_current = tz_min;
            // Set<(f64, f64), 6>: insertion order, exact duplicates dropped (:861, quirk Q9)
            double iv_first[6], iv_second[6];
            int n_iv = 0;
            auto iv_insert = [&](double f, double s) {
                for (int i = 0; i < n_iv; ++i) if (iv_first[i] == f && iv_second[i] == s) return;
                iv_first[n_iv] = f; iv_second[n_iv] = s; ++n_iv;
            };

            roots::PositiveSet<4> dd_extremas = roots::solve_quart_monic(dderiv[1], dderiv[2], dderiv[3], dderiv[4]);
            dd_extremas.sort();
            for (int ri = 0; ri < dd_extremas.size; ++ri) {
                double tz = dd_extremas.data[ri];
                if (tz >= tz_max) continue;

                const double orig = poly_eval(dderiv, tz);
                if (std::fabs(orig) > TOLERANCE) {
                    tz -= orig / poly_eval(poly_deri(dderiv), tz);
                }

                if (poly_eval(deriv, dd_tz_current) * poly_eval(deriv, tz) < 0.0) {
                    iv_insert(dd_tz_current, tz);
                }
                dd_tz_current = tz;
            }
            if (poly_eval(deriv, dd_tz_current) * poly_eval(deriv, tz_max) < 0.0) {
                iv_insert(dd_tz_current, tz_max);
            }

            double tz_current = tz_min;

            for (int k = 0; k < n_iv; ++k) {
                const double tz = shrink_interval(deriv, iv_first[k], iv_second[k]);

                if (tz >= tz_max) continue;

                const double p_val = poly_eval(polynom, tz);
                if (std::fabs(p_val) < 64.0 * std::fabs(poly_eval(dderiv, tz)) * TOLERANCE) {
                    if (vel_udud_check_root(profile, tz, v_max, v_min, a_max, a_min, j_max)) return true;
                } else if (poly_eval(polynom, tz_current) * p_val < 0.0
                           && vel_udud_check_root(profile, shrink_interval(polynom, tz_current, tz), v_max, v_min, a_max, a_min, j_max)) {
                    return true;
                }
                tz_current = tz;
            }
            if (poly_eval(polynom, tz_current) * poly_eval(polynom, tz_max) < 0.0
                && vel_udud_check_root(profile, shrink_interval(polynom, tz_current, tz_max), v_max, v_min, a_max, a_min, j_max)) {
                return true;
            }
        }

        return false;
    }

    // :976-1079
    bool time_acc0_acc1(Profile& profile, double v_max, double v_min, double a_max, double a_min, double) {
        RKO_PHASE(RKO_PH_S2_TAIL);
        if (std::fabs(a0) < EPS && std::fabs(af) < EPS) {
            const double h1 = 2.0 * a_min * g1
                + vd_vd
                + a_max * (2.0 * pd + a_min * tf_tf - 2.0 * tf * vf);
            const double h2 = (a_max - a_min) * (-a_min * vd + a_max * (a_min * tf - vd));

            const double jf = h2 / h1;
            profile.t[0] = a_max / jf;
            profile.t[1] = (-2.0 * a_max * h1 + a_min * a_min * g2) / h2;
            profile.t[2] = profile.t[0];
            profile.t[3] = 0.0;
            profile.t[4] = -a_min / jf;
            profile.t[5] = tf - (2.0 * profile.t[0] + profile.t[1] + 2.0 * profile.t[4]);
            profile.t[6] = profile.t[4];

            return profile.check_with_timing(UDDU, ACC0_ACC1, jf, v_max, v_min, a_max, a_min);
        }
        // UDDU
        {
            const double h1 = std::sqrt(
                144.0
                    * pow2(
                    (a_max - a_min) * (-a_min * vd + a_max * (a_min * tf - vd))
                        - af_af * (a_max * tf - vd)
                        + 2.0 * af * a_min * (a_max * tf - vd)
                        + a0_a0 * (a_min * tf + v0 - vf)
                        - 2.0 * a0 * a_max * (a_min * tf - vd)
                )
                    + 48.0
                    * ad
                    * (3.0 * a0_p3 - 3.0 * af_p3
                    + 12.0 * a_max * a_min * (-a_max + a_min)
                    + 4.0 * af_af * (a_max + 2.0 * a_min)
                    + a0
                    * (-3.0 * af_af
                    + 8.0 * af * (a_min - a_max)
                    + 6.0 * (a_max * a_max + 2.0 * a_max * a_min - a_min * a_min))
                    + 6.0
                    * af
                    * (a_max * a_max - 2.0 * a_max * a_min - a_min * a_min)
                    + a0_a0 * (3.0 * af - 4.0 * (2.0 * a_max + a_min)))
                    * (2.0 * a_min * g1
                    + vd * vd
                    + a_max
                    * (2.0 * pd + a_min * tf * tf
                    - 2.0 * tf * vf))
            );
            const double jf = -(3.0 * af_af * a_max * tf
                - 3.0 * a0_a0 * a_min * tf
                - 6.0 * ad * a_max * a_min * tf
                + 3.0 * a_max * a_min * (a_min - a_max) * tf
                + 3.0 * (a0_a0 - af_af) * vd
                + 6.0 * vd * (af * a_min - a0 * a_max)
                + 3.0 * (a_max * a_max - a_min * a_min) * vd
                + h1 / 4.0)
                / (6.0
                * (2.0 * a_min * g1
                + vd * vd
                + a_max * (2.0 * pd + a_min * tf_tf - 2.0 * tf * vf)));
            profile.t[0] = (a_max - a0) / jf;
            profile.t[1] = (a0_a0 - af_af + 2.0 * ad * a_min
                - 2.0
                * (a_max * a_max - 2.0 * a_max * a_min + a_min * a_min + a_min * jf * tf
                - jf * vd))
                / (2.0 * (a_max - a_min) * jf);
            profile.t[2] = a_max / jf;
            profile.t[3] = 0.0;
            profile.t[4] = -a_min / jf;
            profile.t[5] = tf
                - (profile.t[0] + profile.t[1] + profile.t[2] + 2.0 * profile.t[4] + af / jf);
            profile.t[6] = profile.t[4] + af / jf;

            if (profile.check_with_timing(UDDU, ACC0_ACC1, jf, v_max, v_min, a_max, a_min)) return true;
        }

        return false;
    }

    // :1081-1310
    bool time_acc1(Profile& profile, double v_max, double v_min, double a_max, double a_min, double j_max) {
        RKO_PHASE(RKO_PH_S2_TAIL);
        // a3 != 0.0
        // Case UDDU
        {
            const double h0 = std::sqrt(
                j_max_j_max
                    * (a0_p4 + af_p4 - 4.0 * af_p3 * j_max * tf
                    + 6.0 * af_af * j_max_j_max * tf_tf
                    - 4.0 * a0_p3 * (af - j_max * tf)
                    + 6.0
                    * a0_a0
                    * (af - j_max * tf)
                    * (af - j_max * tf)
                    + 24.0 * af * j_max_j_max * g1
                    - 4.0
                    * a0
                    * (af_p3 - 3.0 * af_af * j_max * tf
                    + 6.0 * j_max_j_max * (-pd + tf * vf))
                    - 12.0 * j_max_j_max * (-vd_vd + j_max * tf * g2))
                    / 3.0
            ) / j_max;
            const double h1 = std::sqrt(
                (a0_a0 + af_af
                    - 2.0 * a0 * af
                    - 2.0 * ad * j_max * tf
                    + 2.0 * h0)
                    / j_max_j_max
                    + tf_tf
            );

            profile.t[0] = -(a0_a0 + af_af + 2.0 * a0 * (j_max * tf - af)
                - 2.0 * j_max * vd
                + h0)
                / (2.0 * j_max * (-ad + j_max * tf));
            profile.t[1] = 0.0;
            profile.t[2] = (tf - h1) / 2.0 - ad / (2.0 * j_max);
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = h1;
            profile.t[6] = tf - (profile.t[0] + profile.t[2] + profile.t[5]);

            if (profile.check_with_timing(UDDU, ACC1, j_max, v_max, v_min, a_max, a_min)) return true;
        }

        // Case UDUD
        {
            const double h0 = std::sqrt(
                j_max_j_max
                    * (a0_p4
                    + af_p4
                    + 4.0 * (af_p3 - a0_p3) * j_max * tf
                    + 6.0 * af_af * j_max_j_max * tf_tf
                    + 6.0
                    * a0_a0
                    * (af + j_max * tf)
                    * (af + j_max * tf)
                    + 24.0 * af * j_max_j_max * g1
                    - 4.0
                    * a0
                    * (a0_a0 * af
                    + af_p3
                    + 3.0 * af_af * j_max * tf
                    + 6.0 * j_max_j_max * (-pd + tf * vf))
                    + 12.0 * j_max_j_max * (vd_vd + j_max * tf * g2))
                    / 3.0
            ) / j_max;
            const double h1 = std::sqrt(
                (a0_a0 + af_af - 2.0 * a0 * af
                    + 2.0 * ad * j_max * tf
                    + 2.0 * h0)
                    / j_max_j_max
                    + tf_tf
            );

            profile.t[0] = 0.0;
            profile.t[1] = 0.0;
            profile.t[2] = -(a0_a0 + af_af - 2.0 * a0 * af
                + 2.0 * j_max * (vd - a0 * tf)
                + h0)
                / (2.0 * j_max * (ad + j_max * tf));
            profile.t[3] = 0.0;
            profile.t[4] = ad / (2.0 * j_max) + (tf - h1) / 2.0;
            profile.t[5] = h1;
            profile.t[6] = tf - (profile.t[5] + profile.t[4] + profile.t[2]);

            if (profile.check_with_timing(UDUD, ACC1, j_max, v_max, v_min, a_max, a_min)) return true;
        }

        // Case UDDU, Solution 2
        {
            const double h0a = a0_p3 - af_p3 - 3.0 * a0_a0 * a_min
                + 3.0 * a_min * a_min * (a0 + j_max * tf)
                + 3.0 * af * a_min * (-a_min - 2.0 * j_max * tf)
                - 3.0 * af_af * (-a_min - j_max * tf)
                - 3.0
                * j_max_j_max
                * (-2.0 * pd - a_min * tf_tf + 2.0 * tf * vf);
            const double h0b = a0_a0 + af_af - 2.0 * (a0 + af) * a_min
                + 2.0 * (a_min * a_min - j_max * (-a_min * tf + vd));
            const double h0c = a0_p4 + 3.0 * af_p4 - 4.0 * (a0_p3 + 2.0 * af_p3) * a_min
                + 6.0 * a0_a0 * a_min * a_min
                + 6.0 * af_af * (a_min * a_min - 2.0 * j_max * vd)
                + 12.0
                * j_max
                * (2.0 * a_min * j_max * g1 - a_min * a_min * vd
                + j_max * vd_vd)
                + 24.0 * af * a_min * j_max * vd
                - 4.0
                * a0
                * (af_p3 - 3.0 * af * a_min * (-a_min - 2.0 * j_max * tf)
                + 3.0 * af_af * (-a_min - j_max * tf)
                + 3.0
                * j_max
                * (-a_min * a_min * tf
                + j_max
                * (-2.0 * pd - a_min * tf_tf
                + 2.0 * tf * vf)));
            const double h1 = std::fabs(j_max) / j_max * std::sqrt(4.0 * h0a * h0a - 6.0 * h0b * h0c);
            const double h2 = 6.0 * j_max * h0b;

            profile.t[0] = 0.0;
            profile.t[1] = 0.0;
            profile.t[2] = (2.0 * h0a + h1) / h2;
            profile.t[3] = -(a0_a0 + af_af - 2.0 * (a0 + af) * a_min
                + 2.0 * (a_min * a_min + a_min * j_max * tf - j_max * vd))
                / (2.0 * j_max * (a0 - a_min - j_max * profile.t[2]));
            profile.t[4] = (a0 - a_min) / j_max - profile.t[2];
            profile.t[5] = tf - (profile.t[2] + profile.t[3] + profile.t[4] + (af - a_min) / j_max);
            profile.t[6] = (af - a_min) / j_max;

            if (profile.check_with_timing(UDDU, ACC1, j_max, v_max, v_min, a_max, a_min)) return true;
        }

        // Case UDUD, Solution 1
        {
            const double h0a = -a0_p3 + af_p3 + 3.0 * (a0_a0 - af_af) * a_max
                - 3.0 * ad * a_max * a_max
                - 6.0 * af * a_max * j_max * tf
                + 3.0 * af_af * j_max * tf
                + 3.0
                * j_max
                * (a_max * a_max * tf
                + j_max * (-2.0 * pd - a_max * tf_tf + 2.0 * tf * vf));
            const double h0b = a0_a0 - af_af
                + 2.0 * ad * a_max
                + 2.0 * j_max * (a_max * tf - vd);
            const double h0c = a0_p4 + 3.0 * af_p4 - 4.0 * (a0_p3 + 2.0 * af_p3) * a_max
                + 6.0 * a0_a0 * a_max * a_max
                - 24.0 * af * a_max * j_max * vd
                + 12.0
                * j_max
                * (2.0 * a_max * j_max * g1
                + j_max * vd_vd
                + a_max * a_max * vd)
                + 6.0 * af_af * (a_max * a_max + 2.0 * j_max * vd)
                - 4.0
                * a0
                * (af_p3 + 3.0 * af * a_max * (a_max - 2.0 * j_max * tf)
                - 3.0 * af_af * (a_max - j_max * tf)
                + 3.0
                * j_max
                * (a_max * a_max * tf
                + j_max
                * (-2.0 * pd - a_max * tf_tf
                + 2.0 * tf * vf)));
            const double h1 = std::fabs(j_max) / j_max * std::sqrt(4.0 * h0a * h0a - 6.0 * h0b * h0c);
            const double h2 = 6.0 * j_max * h0b;

            profile.t[0] = 0.0;
            profile.t[1] = 0.0;
            profile.t[2] = -(2.0 * h0a + h1) / h2;
            profile.t[3] = 2.0 * h1 / h2;
            profile.t[4] = (a_max - a0) / j_max + profile.t[2];
            profile.t[5] = tf - (profile.t[2] + profile.t[3] + profile.t[4] + (-af + a_max) / j_max);
            profile.t[6] = (-af + a_max) / j_max;

            if (profile.check_with_timing(UDUD, ACC1, j_max, v_max, v_min, a_max, a_min)) return true;
        }
        return false;
    }

    // :1312-1437
    bool time_acc0(Profile& profile, double v_max, double v_min, double a_max, double a_min, double j_max) {
        RKO_PHASE(RKO_PH_S2_TAIL);
        // UDUD
        {
            const double h1 = std::sqrt(
                ad_ad / (2.0 * j_max_j_max)
                    - ad * (a_max - a0) / (j_max_j_max)
                    + (a_max * tf - vd) / j_max
            );

            profile.t[0] = (a_max - a0) / j_max;
            profile.t[1] = tf - ad / j_max - 2.0 * h1;
            profile.t[2] = h1;
            profile.t[3] = 0.0;
            profile.t[4] = (af - a_max) / j_max + h1;
            profile.t[5] = 0.0;
            profile.t[6] = 0.0;

            if (profile.check_with_timing(UDUD, NONE, j_max, v_max, v_min, a_max, a_min)) return true;
        }

        // UDUD
        {
            const double h0a = -a0_a0 + af_af - 2.0 * ad * a_max
                + 2.0 * j_max * (a_max * tf - vd);
            const double h0b = a0_p3 + 2.0 * af_p3
                - 6.0 * af_af * a_max
                - 3.0 * a0_a0 * (af - j_max * tf)
                - 3.0 * a0 * a_max * (a_max - 2.0 * af + 2.0 * j_max * tf)
                - 3.0
                * j_max
                * (j_max * (-2.0 * pd + a_max * tf_tf + 2.0 * tf * v0)
                + a_max * (a_max * tf - 2.0 * vd))
                + 3.0
                * af
                * (a_max * a_max + 2.0 * a_max * j_max * tf - 2.0 * j_max * vd);
            const double h0 = std::fabs(j_max) * std::sqrt(4.0 * h0b * h0b - 18.0 * h0a * h0a * h0a);
            const double h1 = 3.0 * j_max * h0a;

            profile.t[0] = (-a0 + a_max) / j_max;
            profile.t[1] = (-a0_p3
                + af_p3
                + af_af * (-6.0 * a_max + 3.0 * j_max * tf)
                + a0_a0 * (-3.0 * af + 6.0 * a_max + 3.0 * j_max * tf)
                + 6.0 * af * (a_max * a_max - j_max * vd)
                + 3.0 * a0 * (af_af - 2.0 * (a_max * a_max + j_max * vd))
                - 6.0 * j_max * (a_max * (a_max * tf - 2.0 * vd) + j_max * g2))
                / h1;
            profile.t[2] = -(ad + h0 / h1) / (2.0 * j_max) + tf / 2.0 - profile.t[1] / 2.0;
            profile.t[3] = h0 / (j_max * h1);
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = tf - (profile.t[0] + profile.t[1] + profile.t[2] + profile.t[3]);

            if (profile.check_with_timing(UDDU, NONE, j_max, v_max, v_min, a_max, a_min)) return true;
        }
        // a3 != 0.0
        // UDDU Solution 1
        {
            const double h0a = a0_p3 + 2.0 * af_p3
                - 6.0 * (af_af + a_max * a_max) * a_max
                - 6.0 * (a0 + af) * a_max * j_max * tf
                + 9.0 * a_max * a_max * (af + j_max * tf)
                + 3.0 * a0 * a_max * (-2.0 * af + 3.0 * a_max)
                + 3.0 * a0_a0 * (af - 2.0 * a_max + j_max * tf)
                - 6.0 * j_max_j_max * g1
                + 6.0 * (af - a_max) * j_max * vd
                - 3.0 * a_max * j_max_j_max * tf_tf;
            const double h0b = a0_a0
                + af_af
                + 2.0
                * (a_max * a_max - (a0 + af) * a_max
                + j_max * (vd - a_max * tf));
            const double h1 = std::fabs(j_max) / j_max * std::sqrt(4.0 * h0a * h0a - 18.0 * h0b * h0b * h0b);
            const double h2 = 6.0 * j_max * h0b;

            profile.t[0] = (-a0 + a_max) / j_max;
            profile.t[1] = ad / j_max - 2.0 * profile.t[0] - (2.0 * h0a - h1) / h2 + tf;
            profile.t[2] = -(2.0 * h0a + h1) / h2;
            profile.t[3] = (2.0 * h0a - h1) / h2;
            profile.t[4] = tf - (profile.t[0] + profile.t[1] + profile.t[2] + profile.t[3]);
            profile.t[5] = 0.0;
            profile.t[6] = 0.0;

            if (profile.check_with_timing(UDDU, ACC0, j_max, v_max, v_min, a_max, a_min)) return true;
        }
        return false;
    }

    // :1439-2202
    bool time_none(Profile& profile, double v_max, double v_min, double a_max, double a_min, double j_max) {
        RKO_PHASE(RKO_PH_S2_TAIL);
        if (std::fabs(v0) < EPS && std::fabs(a0) < EPS && std::fabs(af) < EPS) {
            const double h1 = std::sqrt(tf_tf * vf_vf + pow2(4.0 * pd - tf * vf));
            const double jf = 4.0 * (4.0 * pd - 2.0 * tf * vf + h1) / tf_p3;

            profile.t[0] = tf / 4.0;
            profile.t[1] = 0.0;
            profile.t[2] = 2.0 * profile.t[0];
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = profile.t[0];

            if (profile.check_with_timing(UDDU, NONE, jf, v_max, v_min, a_max, a_min)) return true;
        }

        if (std::fabs(a0) < EPS && std::fabs(af) < EPS) {
            // Profiles with a3 != 0, Solution UDDU: first acc, then constant
            double polynom[4];
            polynom[0] = -2.0 * tf;
            polynom[1] = 2.0 * vd / j_max + tf_tf;
            polynom[2] = 4.0 * (pd - tf * vf) / j_max;
            polynom[3] = (vd_vd + j_max * tf * g2) / (j_max_j_max);

            roots::PositiveSet<4> rts = roots::solve_quart_monic(polynom[0], polynom[1], polynom[2], polynom[3]);
            rts.sort();
            for (int ri = 0; ri < rts.size; ++ri) {
                double t = rts.data[ri];
                if (t > tf / 2.0 || t > (a_max - a0) / j_max) continue;

                // Single Newton step (regarding pd)
                {
                    const double h1 = (j_max * t * (t - tf) + vd)
                        / (j_max * (2.0 * t - tf));
                    const double h2 = (2.0 * j_max * t * (t - tf) + j_max * tf_tf
                        - 2.0 * vd)
                        / (j_max * (2.0 * t - tf) * (2.0 * t - tf));
                    const double orig = (-2.0 * pd
                        + 2.0 * tf * v0
                        + h1 * h1 * j_max * (tf - 2.0 * t)
                        + j_max
                        * tf
                        * (2.0 * h1 * t - t * t - (h1 - t) * tf))
                        / 2.0;
                    const double deriv = (j_max * tf * (2.0 * t - tf) * (h2 - 1.0)) / 2.0
                        + h1 * j_max * (tf - (2.0 * t - tf) * h2 - h1);

                    t -= orig / deriv;
                }

                profile.t[0] = t;
                profile.t[1] = 0.0;
                profile.t[2] = (j_max * t * (t - tf) + vd)
                    / (j_max * (2.0 * t - tf));
                profile.t[3] = tf - 2.0 * t;
                profile.t[4] = t - profile.t[2];
                profile.t[5] = 0.0;
                profile.t[6] = 0.0;

                if (profile.check_with_timing(UDDU, NONE, j_max, v_max, v_min, a_max, a_min)) return true;
            }
        }

        // UDUD T 0246
        {
            const double h0 = std::sqrt(
                2.0 * j_max_j_max
                    * (2.0
                    * pow2(
                    a0_p3 - af_p3 - 3.0 * af_af * j_max * tf
                        + 9.0 * af * j_max_j_max * tf_tf
                        - 3.0 * a0_a0 * (af + j_max * tf)
                        + 3.0 * a0 * pow2(af + j_max * tf)
                        + 3.0
                        * j_max_j_max
                        * (8.0 * pd + j_max * tf_tf * tf
                        - 8.0 * tf * vf)
                )
                    - 3.0
                    * (a0_a0 + af_af
                    - 2.0 * af * j_max * tf
                    - 2.0 * a0 * (af + j_max * tf)
                    - j_max * (j_max * tf_tf + 4.0 * v0 - 4.0 * vf))
                    * (a0_p4
                    + af_p4
                    + 4.0 * af_p3 * j_max * tf
                    + 6.0 * af_af * j_max_j_max * tf_tf
                    - 3.0
                    * j_max_j_max
                    * j_max_j_max
                    * tf_tf
                    * tf_tf
                    - 4.0 * a0_p3 * (af + j_max * tf)
                    + 6.0 * a0_a0 * pow2(af + j_max * tf)
                    - 12.0
                    * af
                    * j_max_j_max
                    * (8.0 * pd + j_max * tf_tf * tf
                    - 8.0 * tf * v0)
                    + 48.0 * j_max_j_max * vd_vd
                    + 48.0 * j_max_j_max * j_max * tf * g2
                    - 4.0
                    * a0
                    * (af_p3 + 3.0 * af_af * j_max * tf
                    - 9.0 * af * j_max_j_max * tf_tf
                    - 3.0
                    * j_max_j_max
                    * (8.0 * pd + j_max * tf_tf * tf
                    - 8.0 * tf * vf))))
            ) / j_max;
            const double h1 = 12.0
                * j_max
                * (-a0_a0 - af_af
                + 2.0 * af * j_max * tf
                + 2.0 * a0 * (af + j_max * tf)
                + j_max * (j_max * tf_tf + 4.0 * v0 - 4.0 * vf));
            const double h2 = -4.0 * a0_p3 + 4.0 * af_p3 + 12.0 * a0_a0 * af
                - 12.0 * a0 * af_af
                + 48.0 * j_max_j_max * pd
                + 12.0 * (a0_a0 - af_af) * j_max * tf
                - 24.0 * j_max_j_max * tf * (v0 + vf)
                + 24.0 * ad * j_max * vd;
            const double h3 = 2.0 * a0_p3 - 2.0 * af_p3 - 6.0 * a0_a0 * af
                + 6.0 * a0 * af_af;

            profile.t[0] = (h3
                - 48.0 * j_max_j_max * (tf * vf - pd)
                - 6.0 * (a0_a0 + af_af) * j_max * tf
                + 12.0 * a0 * af * j_max * tf
                + 6.0
                * (a0 + 3.0 * af + j_max * tf)
                * tf_tf
                * j_max_j_max
                - h0)
                / h1;
            profile.t[1] = 0.0;
            profile.t[2] = (h2 + h0) / h1;
            profile.t[3] = 0.0;
            profile.t[4] = (-h2 + h0) / h1;
            profile.t[5] = 0.0;
            profile.t[6] = (-h3 + 48.0 * j_max_j_max * (tf * v0 - pd)
                - 6.0 * (a0_a0 + af_af) * j_max * tf
                + 12.0 * a0 * af * j_max * tf
                + 6.0
                * (af + 3.0 * a0 + j_max * tf)
                * tf_tf
                * j_max_j_max
                - h0)
                / h1;

            if (profile.check_with_timing(UDUD, NONE, j_max, v_max, v_min, a_max, a_min)) return true;
        }

        // Profiles with a3 != 0.0, Solution UDDU
        {
            // T 0234
            {
                const double ph1 = af + j_max * tf;

                double polynom[4];
                polynom[0] = -2.0 * (ad + j_max * tf) / j_max;
                polynom[1] = 2.0
                    * (a0_a0 + af_af + j_max * (af * tf + vd)
                    - 2.0 * a0 * ph1)
                    / j_max_j_max
                    + tf_tf;
                polynom[2] = 2.0
                    * (a0_p3 - af_p3 - 3.0 * af_af * j_max * tf
                    + 3.0 * a0 * ph1 * (ph1 - a0)
                    - 6.0 * j_max_j_max * (-pd + tf * vf))
                    / (3.0 * j_max_j_max * j_max);
                polynom[3] = (a0_p4 + af_p4 + 4.0 * af_p3 * j_max * tf
                    - 4.0 * a0_p3 * ph1
                    + 6.0 * a0_a0 * ph1 * ph1
                    + 24.0 * j_max_j_max * af * g1
                    - 4.0
                    * a0
                    * (af_p3
                    + 3.0 * af_af * j_max * tf
                    + 6.0 * j_max_j_max * (-pd + tf * vf))
                    + 6.0 * j_max_j_max * af_af * tf_tf
                    + 12.0 * j_max_j_max * (vd_vd + j_max * tf * g2))
                    / (12.0 * j_max_j_max * j_max_j_max);

                const double t_min = ad / j_max;
                const double t_max = fmin_((a_max - a0) / j_max, (ad / j_max + tf) / 2.0);

                roots::PositiveSet<4> rts = roots::solve_quart_monic(polynom[0], polynom[1], polynom[2], polynom[3]);
                rts.sort();
                for (int ri = 0; ri < rts.size; ++ri) {
                    double t = rts.data[ri];
                    if (t < t_min || t > t_max) continue;

                    // Single Newton step (regarding pd)
                    {
                        const double h0 = j_max * (2.0 * t - tf) - ad;
                        const double h1 = (ad_ad - 2.0 * af * j_max * t
                            + 2.0 * a0 * j_max * (t - tf)
                            + 2.0 * j_max * (j_max * t * (t - tf) + vd))
                            / (2.0 * j_max * h0);
                        const double h2 = (-ad_ad
                            + 2.0 * j_max_j_max * (tf_tf + t * (t - tf))
                            + (a0 + af) * j_max * tf
                            - ad * h0
                            - 2.0 * j_max * vd)
                            / (h0 * h0);
                        const double orig = (-a0_p3
                            + af_p3
                            + 3.0 * ad_ad * j_max * (h1 - t)
                            + 3.0 * ad * j_max_j_max * (h1 - t) * (h1 - t)
                            - 3.0 * a0 * af * ad
                            + 3.0
                            * j_max_j_max
                            * (a0 * tf_tf - 2.0 * pd
                            + 2.0 * tf * v0
                            + h1 * h1 * j_max * (tf - 2.0 * t)
                            + j_max
                            * tf
                            * (2.0 * h1 * t - t * t - (h1 - t) * tf)))
                            / (6.0 * j_max_j_max);
                        const double deriv = (h0 * (-ad + j_max * tf) * (h2 - 1.0))
                            / (2.0 * j_max)
                            + h1 * (-ad + j_max * (tf - h1) - h0 * h2);

                        t -= orig / deriv;
                    }

                    profile.t[0] = t;
                    profile.t[1] = 0.0;
                    profile.t[2] = (ad_ad
                        + 2.0
                        * j_max
                        * (-a0 * tf - ad * t
                        + j_max * t * (t - tf)
                        + vd))
                        / (2.0 * j_max * (-ad + j_max * (2.0 * t - tf)));
                    profile.t[3] = ad / j_max + tf - 2.0 * t;
                    profile.t[4] = tf - (t + profile.t[2] + profile.t[3]);
                    profile.t[5] = 0.0;
                    profile.t[6] = 0.0;

                    if (profile.check_with_timing(UDDU, NONE, j_max, v_max, v_min, a_max, a_min)) return true;
                }
            }

            // T 3456
            {
                const double h1 = 3.0 * j_max * (ad_ad + 2.0 * j_max * (a0 * tf - vd));
                const double h2 = ad_ad + 2.0 * j_max * (a0 * tf - vd);
                const double h0 = std::sqrt(
                    4.0 * pow2(
                        2.0 * (a0_p3 - af_p3)
                            - 6.0 * a0_a0 * (af - j_max * tf)
                            + 6.0 * j_max_j_max * g1
                            + 3.0
                            * a0
                            * (2.0 * af_af - 2.0 * j_max * af * tf
                            + j_max_j_max * tf_tf)
                            + 6.0 * ad * j_max * vd
                    ) - 18.0 * h2 * h2 * h2
                ) / h1
                    * std::fabs(j_max)
                    / j_max;

                profile.t[0] = 0.0;
                profile.t[1] = 0.0;
                profile.t[2] = 0.0;
                profile.t[3] = (af_p3 - a0_p3
                    + 3.0 * (af_af - a0_a0) * j_max * tf
                    - 3.0 * ad * (a0 * af + 2.0 * j_max * vd)
                    - 6.0 * j_max_j_max * g2)
                    / h1;
                profile.t[4] = (tf - profile.t[3] - h0) / 2.0 - ad / (2.0 * j_max);
                profile.t[5] = h0;
                profile.t[6] = (tf - profile.t[3] + ad / j_max - h0) / 2.0;

                if (profile.check_with_timing(UDDU, NONE, j_max, v_max, v_min, a_max, a_min)) return true;
            }

            // T 2346
            {
                const double ph1 = ad_ad + 2.0 * (af + a0) * j_max * tf
                    - j_max * (j_max * tf_tf + 4.0 * vd);
                const double ph2 = j_max * tf_tf * g1
                    - vd * (-2.0 * pd - tf * v0 + 3.0 * tf * vf);
                const double ph3 = 5.0 * af_af - 8.0 * af * j_max * tf
                    + 2.0 * j_max * (2.0 * j_max * tf_tf - vd);
                const double ph4 = j_max_j_max * tf_p4 - 2.0 * vd_vd
                    + 8.0 * j_max * tf * (-pd + tf * vf);
                const double ph5 = 5.0 * af_p4
                    - 8.0 * af_p3 * j_max * tf
                    - 12.0 * af_af * j_max * (j_max * tf_tf + vd)
                    + 24.0
                    * af
                    * j_max_j_max
                    * (-2.0 * pd + j_max * tf_p3 + 2.0 * tf * vf)
                    - 6.0 * j_max_j_max * ph4;
                const double ph6 = -vd_vd
                    + j_max
                    * tf
                    * (-2.0 * pd + 3.0 * tf * v0 - tf * vf)
                    - af * g2;

                double polynom[4];
                polynom[0] = -(4.0 * (a0_p3 - af_p3)
                    - 12.0 * a0_a0 * (af - j_max * tf)
                    + 6.0
                    * a0
                    * (2.0 * af_af - 2.0 * af * j_max * tf
                    + j_max * (j_max * tf_tf - 2.0 * vd))
                    + 6.0 * af * j_max * (3.0 * j_max * tf_tf + 2.0 * vd)
                    - 6.0
                    * j_max_j_max
                    * (-4.0 * pd + j_max * tf_p3 - 2.0 * tf * v0
                    + 6.0 * tf * vf))
                    / (3.0 * j_max * ph1);
                polynom[1] = -(-a0_p4 - af_p4
                    + 4.0 * a0_p3 * (af - j_max * tf)
                    + a0_a0
                    * (-6.0 * af_af + 8.0 * af * j_max * tf
                    - 4.0 * j_max * (j_max * tf_tf - vd))
                    + 2.0 * af_af * j_max * (j_max * tf_tf + 2.0 * vd)
                    - 4.0
                    * af
                    * j_max_j_max
                    * (-3.0 * pd
                    + j_max * tf_p3
                    + 2.0 * tf * v0
                    + tf * vf)
                    + j_max_j_max
                    * (j_max_j_max * tf_p4 - 8.0 * vd_vd
                    + 4.0
                    * j_max
                    * tf
                    * (-3.0 * pd + tf * v0 + 2.0 * tf * vf))
                    + 2.0
                    * a0
                    * (2.0 * af_p3 - 2.0 * af_af * j_max * tf
                    + af * j_max * (-3.0 * j_max * tf_tf - 4.0 * vd)
                    + j_max_j_max
                    * (-6.0 * pd + j_max * tf_p3 - 4.0 * tf * v0
                    + 10.0 * tf * vf)))
                    / (j_max_j_max * ph1);
                polynom[2] = -(a0_p5 - af_p5 + af_p4 * j_max * tf
                    - 5.0 * a0_p4 * (af - j_max * tf)
                    + 2.0 * a0_p3 * ph3
                    + 4.0 * af_p3 * j_max * (j_max * tf_tf + vd)
                    + 12.0 * j_max_j_max * af * ph6
                    - 2.0
                    * a0_a0
                    * (5.0 * af_p3
                    - 9.0 * af_af * j_max * tf
                    - 6.0 * af * j_max * vd
                    + 6.0
                    * j_max_j_max
                    * (-2.0 * pd - tf * v0 + 3.0 * tf * vf))
                    - 12.0 * j_max_j_max * j_max * ph2
                    + a0 * ph5)
                    / (3.0 * j_max_j_max * j_max * ph1);
                polynom[3] = -(-a0_p6 - af_p6
                    + 6.0 * a0_p5 * (af - j_max * tf)
                    - 48.0 * af_p3 * j_max_j_max * g1
                    + 72.0
                    * j_max_j_max
                    * j_max
                    * (j_max * g1 * g1
                    + vd_vd * vd
                    + 2.0 * af * g1 * vd)
                    - 3.0 * a0_p4 * ph3
                    - 36.0 * af_af * j_max_j_max * vd_vd
                    + 6.0 * af_p4 * j_max * vd
                    + 4.0
                    * a0_p3
                    * (5.0 * af_p3
                    - 9.0 * af_af * j_max * tf
                    - 6.0 * af * j_max * vd
                    + 6.0
                    * j_max_j_max
                    * (-2.0 * pd - tf * v0 + 3.0 * tf * vf))
                    - 3.0 * a0_a0 * ph5
                    + 6.0
                    * a0
                    * (af_p5
                    - af_p4 * j_max * tf
                    - 4.0 * af_p3 * j_max * (j_max * tf_tf + vd)
                    + 12.0 * j_max_j_max * (-af * ph6 + j_max * ph2)))
                    / (18.0 * j_max_j_max * j_max_j_max * ph1);

                const double t_max = (a0 - a_min) / j_max;

                roots::PositiveSet<4> rts = roots::solve_quart_monic(polynom[0], polynom[1], polynom[2], polynom[3]);
                rts.sort();
                for (int ri = 0; ri < rts.size; ++ri) {
                    double t = rts.data[ri];
                    if (t > t_max) continue;

                    // Single Newton step (regarding pd)
                    {
                        const double h1 = ad_ad / 2.0
                            + j_max
                            * (af * t + (j_max * t - a0) * (t - tf)
                            - vd);
                        const double h2 = -ad + j_max * (tf - 2.0 * t);
                        const double h3 = std::sqrt(h1);
                        const double orig = (af_p3 - a0_p3
                            + 3.0 * af * j_max * t * (af + j_max * t)
                            + 3.0 * a0_a0 * (af + j_max * t)
                            - 3.0
                            * a0
                            * (af_af
                            + 2.0 * af * j_max * t
                            + j_max_j_max * (t * t - tf_tf))
                            + 3.0
                            * j_max_j_max
                            * (-2.0 * pd
                            + j_max * t * (t - tf) * tf
                            + 2.0 * tf * v0))
                            / (6.0 * j_max_j_max)
                            - h3 * h3 * h3 / (j_max * std::fabs(j_max))
                            + ((-ad - j_max * t) * h1) / (j_max_j_max);
                        const double deriv = (6.0 * j_max * h2 * h3 / std::fabs(j_max)
                            + 2.0 * (-ad - j_max * tf) * h2
                            - 2.0
                            * (3.0 * ad_ad
                            + af * j_max * (8.0 * t - 2.0 * tf)
                            + 4.0 * a0 * j_max * (-2.0 * t + tf)
                            + 2.0
                            * j_max
                            * (j_max * t * (3.0 * t - 2.0 * tf) - vd)))
                            / (4.0 * j_max);

                        t -= orig / deriv;
                    }

                    const double h1 = std::sqrt(
                        2.0 * ad_ad
                            + 4.0
                            * j_max
                            * (ad * t + a0 * tf + j_max * t * (t - tf)
                            - vd)
                    ) / std::fabs(j_max);

                    // Solution 2 with aPlat
                    profile.t[0] = 0.0;
                    profile.t[1] = 0.0;
                    profile.t[2] = t;
                    profile.t[3] = tf - 2.0 * t - ad / j_max - h1;
                    profile.t[4] = h1 / 2.0;
                    profile.t[5] = 0.0;
                    profile.t[6] = tf - (t + profile.t[3] + profile.t[4]);

                    if (profile.check_with_timing(UDDU, NONE, j_max, v_max, v_min, a_max, a_min)) return true;
                }
            }
        }

        // Profiles with a3 != 0.0, Solution UDUD
        {
            // T 0124
            {
                const double ph0 = -2.0 * pd - tf * v0 + 3.0 * tf * vf;
                const double ph1 = -ad + j_max * tf;
                const double ph2 = j_max * tf_tf * g1 - vd * ph0;
                const double ph3 = 5.0 * af_af
                    + 2.0 * j_max * (2.0 * j_max * tf_tf - vd - 4.0 * af * tf);
                const double ph4 = j_max_j_max * tf_p4 - 2.0 * vd_vd
                    + 8.0 * j_max * tf * (-pd + tf * vf);
                const double ph5 = 5.0 * af_p4
                    - 8.0 * af_p3 * j_max * tf
                    - 12.0 * af_af * j_max * (j_max * tf_tf + vd)
                    + 24.0
                    * af
                    * j_max_j_max
                    * (-2.0 * pd + j_max * tf_p3 + 2.0 * tf * vf)
                    - 6.0 * j_max_j_max * ph4;
                const double ph6 = -vd_vd
                    + j_max
                    * tf
                    * (-2.0 * pd + 3.0 * tf * v0 - tf * vf);
                const double ph7 = 3.0 * j_max_j_max * ph1 * ph1;

                double polynom[4];
                polynom[0] = (4.0 * af * tf - 2.0 * j_max * tf_tf - 4.0 * vd) / ph1;
                polynom[1] = (-2.0 * (a0_p4 + af_p4)
                    + 8.0 * af_p3 * j_max * tf
                    + 6.0 * af_af * j_max_j_max * tf_tf
                    + 8.0 * a0_p3 * (af - j_max * tf)
                    - 12.0
                    * a0_a0
                    * (af - j_max * tf)
                    * (af - j_max * tf)
                    - 12.0
                    * af
                    * j_max_j_max
                    * (-pd + j_max * tf_p3 - 2.0 * tf * v0
                    + 3.0 * tf * vf)
                    + 2.0
                    * a0
                    * (4.0 * af_p3 - 12.0 * af_af * j_max * tf
                    + 9.0 * af * j_max_j_max * tf_tf
                    - 3.0
                    * j_max_j_max
                    * (2.0 * pd + j_max * tf_p3 - 2.0 * tf * vf))
                    + 3.0
                    * j_max_j_max
                    * (j_max_j_max * tf_p4 + 4.0 * vd_vd
                    - 4.0
                    * j_max
                    * tf
                    * (pd + tf * v0 - 2.0 * tf * vf)))
                    / ph7;
                polynom[2] = (-a0_p5 + af_p5 - af_p4 * j_max * tf
                    + 5.0 * a0_p4 * (af - j_max * tf)
                    - 2.0 * a0_p3 * ph3
                    - 4.0 * af_p3 * j_max * (j_max * tf_tf + vd)
                    + 12.0 * af_af * j_max_j_max * g2
                    - 12.0 * af * j_max_j_max * ph6
                    + 2.0
                    * a0_a0
                    * (5.0 * af_p3
                    - 9.0 * af_af * j_max * tf
                    - 6.0 * af * j_max * vd
                    + 6.0 * j_max_j_max * ph0)
                    + 12.0 * j_max_j_max * j_max * ph2
                    + a0
                    * (-5.0 * af_p4
                    + 8.0 * af_p3 * j_max * tf
                    + 12.0 * af_af * j_max * (j_max * tf_tf + vd)
                    - 24.0
                    * af
                    * j_max_j_max
                    * (-2.0 * pd + j_max * tf_p3 + 2.0 * tf * vf)
                    + 6.0 * j_max_j_max * ph4))
                    / (j_max * ph7);
                polynom[3] = -(a0_p6 + af_p6
                    - 6.0 * a0_p5 * (af - j_max * tf)
                    + 48.0 * af_p3 * j_max_j_max * g1
                    - 72.0
                    * j_max_j_max
                    * j_max
                    * (j_max * g1 * g1
                    + vd_vd * vd
                    + 2.0 * af * g1 * vd)
                    + 3.0 * a0_p4 * ph3
                    - 6.0 * af_p4 * j_max * vd
                    + 36.0 * af_af * j_max_j_max * vd_vd
                    - 4.0
                    * a0_p3
                    * (5.0 * af_p3
                    - 9.0 * af_af * j_max * tf
                    - 6.0 * af * j_max * vd
                    + 6.0 * j_max_j_max * ph0)
                    + 3.0 * a0_a0 * ph5
                    - 6.0
                    * a0
                    * (af_p5
                    - af_p4 * j_max * tf
                    - 4.0 * af_p3 * j_max * (j_max * tf_tf + vd)
                    + 12.0
                    * j_max_j_max
                    * (af_af * g2 - af * ph6 + j_max * ph2)))
                    / (6.0 * j_max_j_max * ph7);

                roots::PositiveSet<4> rts = roots::solve_quart_monic(polynom[0], polynom[1], polynom[2], polynom[3]);
                rts.sort();
                for (int ri = 0; ri < rts.size; ++ri) {
                    const double t = rts.data[ri];
                    if (t > tf || t > (a_max - a0) / j_max) continue;

                    const double h1 = std::sqrt(
                        ad_ad / (2.0 * j_max_j_max)
                            + (a0 * (t + tf) - af * t + j_max * t * tf
                            - vd)
                            / j_max
                    );

                    profile.t[0] = t;
                    profile.t[1] = tf - ad / j_max - 2.0 * h1;
                    profile.t[2] = h1;
                    profile.t[3] = 0.0;
                    profile.t[4] = ad / j_max + h1 - t;
                    profile.t[5] = 0.0;
                    profile.t[6] = 0.0;

                    if (profile.check_with_timing(UDUD, NONE, j_max, v_max, v_min, a_max, a_min)) return true;
                }
            }
        }

        // 3 step profile (ak. UZD), sometimes missed because of numerical errors T 012
        {
            const double h1 = std::sqrt(-ad_ad + j_max * (2.0 * (a0 + af) * tf - 4.0 * vd + j_max * tf_tf)) / std::fabs(j_max);

            profile.t[0] = (tf - h1 + ad / j_max) / 2.0;
            profile.t[1] = h1;
            profile.t[2] = (tf - h1 - ad / j_max) / 2.0;
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = 0.0;

            if (profile.check_with_timing(UDDU, NONE, j_max, v_max, v_min, a_max, a_min)) return true;
        }

        // 3 step profile (ak. UZU), sometimes missed because of numerical errors
        {
            double polynom[4];
            polynom[0] = ad_ad;
            polynom[1] = ad_ad * tf;
            polynom[2] = (a0_a0 + af_af + 10.0 * a0 * af) * tf_tf + 24.0 * (tf * (af * v0 - a0 * vf) - pd * ad) + 12.0 * vd_vd;
            polynom[3] = -3.0 * tf * ((a0_a0 + af_af + 2.0 * a0 * af) * tf_tf - 4.0 * vd * (a0 + af) * tf + 4.0 * vd_vd);

            roots::PositiveSet<3> rts = roots::solve_cub(polynom[0], polynom[1], polynom[2], polynom[3]);
            rts.sort();
            for (int ri = 0; ri < rts.size; ++ri) {
                const double t = rts.data[ri];
                if (t > tf) continue;
                const double jf = ad / (tf - t);

                profile.t[0] = (2.0 * (vd - a0 * tf) + ad * (t - tf)) / (2.0 * jf * t);
                profile.t[1] = t;
                profile.t[2] = 0.0;
                profile.t[3] = 0.0;
                profile.t[4] = 0.0;
                profile.t[5] = 0.0;
                profile.t[6] = tf - (profile.t[0] + profile.t[1]);

                // note: checked with j_max, not jf (:2164-2171)
                if (profile.check_with_timing(UDDU, NONE, j_max, v_max, v_min, a_max, a_min)) return true;
            }
        }

        // 3 step profile (ak. UDU), sometimes missed because of numerical errors
        {
            profile.t[0] = (ad_ad / j_max + 2.0 * (a0 + af) * tf - j_max * tf_tf - 4.0 * vd) / (4.0 * (ad - j_max * tf));
            profile.t[1] = 0.0;
            profile.t[2] = -ad / (2.0 * j_max) + tf / 2.0;
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = tf - (profile.t[0] + profile.t[2]);

            if (profile.check_with_timing(UDDU, NONE, j_max, v_max, v_min, a_max, a_min)) return true;
        }

        return false;
    }

    // :2449-2482
    bool get_profile(Profile& profile) {
        const bool up_first = pd > tf * v0;
        const double v_max = up_first ? _v_max : _v_min;
        const double v_min = up_first ? _v_min : _v_max;
        const double a_max = up_first ? _a_max : _a_min;
        const double a_min = up_first ? _a_min : _a_max;
        const double j_max = up_first ? _j_max : -_j_max;

        return time_acc0_acc1_vel(profile, v_max, v_min, a_max, a_min, j_max)
            || time_vel(profile, v_max, v_min, a_max, a_min, j_max)
            || time_acc0_vel(profile, v_max, v_min, a_max, a_min, j_max)
            || time_acc1_vel(profile, v_max, v_min, a_max, a_min, j_max)
            || time_acc0_acc1_vel(profile, v_min, v_max, a_min, a_max, -j_max)
            || time_vel(profile, v_min, v_max, a_min, a_max, -j_max)
            || time_acc0_vel(profile, v_min, v_max, a_min, a_max, -j_max)
            || time_acc1_vel(profile, v_min, v_max, a_min, a_max, -j_max)
            || time_acc0_acc1(profile, v_max, v_min, a_max, a_min, j_max)
            || time_acc0(profile, v_max, v_min, a_max, a_min, j_max)
            || time_acc1(profile, v_max, v_min, a_max, a_min, j_max)
            || time_none(profile, v_max, v_min, a_max, a_min, j_max)
            || time_acc0_acc1(profile, v_min, v_max, a_min, a_max, -j_max)
            || time_acc0(profile, v_min, v_max, a_min, a_max, -j_max)
            || time_acc1(profile, v_min, v_max, a_min, a_max, -j_max)
            || time_none(profile, v_min, v_max, a_min, a_max, -j_max);
    }
};

}  // namespace rko
