// rko_steps.hpp — parity oracle: Step 1 for every interface/order, Step 2 for
// all but the third-order position interface (that one is rko_step2_pos3.hpp).
//
// TEST INFRASTRUCTURE ONLY (see rko_math.hpp). Restates rsruckig v2.1.2:
//   position_third_step1.rs, velocity_third_step{1,2}.rs,
//   position_second_step{1,2}.rs, velocity_second_step{1,2}.rs,
//   position_first_step{1,2}.rs. Candidate order IS semantics.
#pragma once
#include "rko_core.hpp"

namespace rko {

// ============================================== position_third_step1.rs
struct PositionThirdOrderStep1 {
    double v0, a0, vf, af, _v_max, _v_min, _a_max, _a_min, _j_max;
    double pd, v0_v0, vf_vf, a0_a0, a0_p3, a0_p4, af_af, af_p3, af_p4, j_max_j_max;
    Profile valid_profiles[6];
    int current_index = 0;

    // position_third_step1.rs:39-86
    PositionThirdOrderStep1(double p0, double v0_, double a0_, double pf, double vf_, double af_, double v_max, double v_min, double a_max, double a_min, double j_max)
        : v0(v0_), a0(a0_), vf(vf_), af(af_), _v_max(v_max), _v_min(v_min), _a_max(a_max), _a_min(a_min), _j_max(j_max) {
        pd = pf - p0;
        v0_v0 = v0 * v0;
        vf_vf = vf * vf;
        a0_a0 = a0 * a0;
        af_af = af * af;
        a0_p3 = a0 * a0_a0;
        a0_p4 = a0_a0 * a0_a0;
        af_p3 = af * af_af;
        af_p4 = af_af * af_af;
        j_max_j_max = j_max * j_max;
    }

    // :89-95 (count saturates at 5, quirk Q11)
    void add_profile() {
        if (current_index < 5) {
            current_index += 1;
            valid_profiles[current_index].set_boundary_from_profile(valid_profiles[current_index - 1]);
        }
    }

    // :97-240 (all four variants are checked as ReachedLimits::Vel, quirk Q6)
    void time_all_vel(double v_max, double v_min, double a_max, double a_min, double j_max, bool) {
        RKO_PHASE(RKO_PH_S1_VEL);
        Profile* profile = &valid_profiles[current_index];
        // ACC0_ACC1_VEL
        profile->t[0] = (-a0 + a_max) / j_max;
        profile->t[1] = (a0_a0 / 2.0 - a_max * a_max - j_max * (v0 - v_max)) / (a_max * j_max);
        profile->t[2] = a_max / j_max;
        profile->t[3] = (3.0 * (a0_p4 * a_min - af_p4 * a_max)
            + 8.0 * a_max * a_min * (af_p3 - a0_p3 + 3.0 * j_max * (a0 * v0 - af * vf))
            + 6.0 * a0_a0 * a_min * (a_max * a_max - 2.0 * j_max * v0)
            - 6.0 * af_af * a_max * (a_min * a_min - 2.0 * j_max * vf)
            - 12.0 * j_max * (a_max * a_min * (a_max * (v0 + v_max) - a_min * (vf + v_max) - 2.0 * j_max * pd)
                + (a_min - a_max) * j_max * v_max * v_max
                + j_max * (a_max * vf_vf - a_min * v0_v0)))
            / (24.0 * a_max * a_min * j_max_j_max * v_max);
        profile->t[4] = -a_min / j_max;
        profile->t[5] = -(af_af / 2.0 - a_min * a_min - j_max * (vf - v_max)) / (a_min * j_max);
        profile->t[6] = profile->t[4] + af / j_max;

        if (profile->check_with_timing(UDDU, VEL, j_max, v_max, v_min, a_max, a_min)) {
            add_profile();
            return;
        }

        // ACC1_VEL
        profile = &valid_profiles[current_index];
        const double t_acc0 = std::sqrt(a0_a0 / (2.0 * j_max_j_max) + (v_max - v0) / j_max);

        profile->t[0] = t_acc0 - a0 / j_max;
        profile->t[1] = 0.0;
        profile->t[2] = t_acc0;
        profile->t[3] = -(3.0 * af_p4
            - 8.0 * a_min * (af_p3 - a0_p3)
            - 24.0 * a_min * j_max * (a0 * v0 - af * vf)
            + 6.0 * af_af * (a_min * a_min - 2.0 * j_max * vf)
            - 12.0 * j_max * (2.0 * a_min * j_max * pd
                + a_min * a_min * (vf + v_max)
                + j_max * (v_max * v_max - vf_vf)
                + a_min * t_acc0 * (a0_a0 - 2.0 * j_max * (v0 + v_max))))
            / (24.0 * a_min * j_max_j_max * v_max);

        if (profile->check_with_timing(UDDU, VEL, j_max, v_max, v_min, a_max, a_min)) {
            add_profile();
            return;
        }

        // ACC0_VEL
        profile = &valid_profiles[current_index];
        const double t_acc1 = std::sqrt(af_af / (2.0 * j_max_j_max) + (v_max - vf) / j_max);

        profile->t[0] = (-a0 + a_max) / j_max;
        profile->t[1] = (a0_a0 / 2.0 - a_max * a_max - j_max * (v0 - v_max)) / (a_max * j_max);
        profile->t[2] = a_max / j_max;
        profile->t[3] = (3.0 * a0_p4
            + 8.0 * a_max * (af_p3 - a0_p3)
            + 24.0 * a_max * j_max * (a0 * v0 - af * vf)
            + 6.0 * a0_a0 * (a_max * a_max - 2.0 * j_max * v0)
            - 12.0 * j_max * (-2.0 * a_max * j_max * pd
                + a_max * a_max * (v0 + v_max)
                + j_max * (v_max * v_max - v0_v0)
                + a_max * t_acc1 * (-af_af + 2.0 * (vf + v_max) * j_max)))
            / (24.0 * a_max * j_max_j_max * v_max);
        profile->t[4] = t_acc1;
        profile->t[5] = 0.0;
        profile->t[6] = t_acc1 + af / j_max;

        if (profile->check_with_timing(UDDU, VEL, j_max, v_max, v_min, a_max, a_min)) {
            add_profile();
            return;
        }

        // VEL — Solution 3/4
        profile = &valid_profiles[current_index];
        profile->t[0] = t_acc0 - a0 / j_max;
        profile->t[1] = 0.0;
        profile->t[2] = t_acc0;
        profile->t[3] = (af_p3 - a0_p3) / (3.0 * j_max_j_max * v_max)
            + (a0 * v0 - af * vf + (af_af * t_acc1 + a0_a0 * t_acc0) / 2.0) / (j_max * v_max)
            - (v0 / v_max + 1.0) * t_acc0
            - (vf / v_max + 1.0) * t_acc1
            + pd / v_max;
        if (profile->check_with_timing(UDDU, VEL, j_max, v_max, v_min, a_max, a_min)) {
            add_profile();
        }
    }

    // :241-327
    void time_acc0_acc1(double v_max, double v_min, double a_max, double a_min, double j_max, bool return_after_found) {
        RKO_PHASE(RKO_PH_S1_ACC0_ACC1);
        double h1 = (3. * (af_p4 * a_max - a0_p4 * a_min)
            + a_max * a_min * (8. * (a0_p3 - af_p3)
                + 3. * a_max * a_min * (a_max - a_min)
                + 6. * a_min * af_af
                - 6. * a_max * a0_a0)
            + 12. * j_max * (a_max * a_min * ((a_max - 2. * a0) * v0 - (a_min - 2. * af) * vf)
                + a_min * a0_a0 * v0
                - a_max * af_af * vf))
            / (3. * (a_max - a_min) * j_max_j_max)
            + 4. * (a_max * vf_vf - a_min * v0_v0 - 2. * a_min * a_max * pd) / (a_max - a_min);

        if (h1 >= 0.) {
            h1 = std::sqrt(h1) / 2.;
            const double h2 = a0_a0 / (2. * a_max * j_max) + (a_min - 2. * a_max) / (2. * j_max) - v0 / a_max;
            const double h3 = -af_af / (2. * a_min * j_max) - (a_max - 2. * a_min) / (2. * j_max) + vf / a_min;

            // UDDU: Solution 2
            if (h2 > h1 / a_max && h3 > -h1 / a_min) {
                Profile* profile = &valid_profiles[current_index];
                profile->t[0] = (-a0 + a_max) / j_max;
                profile->t[1] = h2 - h1 / a_max;
                profile->t[2] = a_max / j_max;
                profile->t[3] = 0.;
                profile->t[4] = -a_min / j_max;
                profile->t[5] = h3 + h1 / a_min;
                profile->t[6] = profile->t[4] + af / j_max;

                if (profile->check_with_timing(UDDU, ACC0_ACC1, j_max, v_max, v_min, a_max, a_min)) {
                    add_profile();
                    if (return_after_found) return;
                }
            }

            // UDDU: Solution 1
            if (h2 > -h1 / a_max && h3 > h1 / a_min) {
                Profile* profile = &valid_profiles[current_index];
                profile->t[0] = (-a0 + a_max) / j_max;
                profile->t[1] = h2 + h1 / a_max;
                profile->t[2] = a_max / j_max;
                profile->t[3] = 0.;
                profile->t[4] = -a_min / j_max;
                profile->t[5] = h3 - h1 / a_min;
                profile->t[6] = profile->t[4] + af / j_max;

                if (profile->check_with_timing(UDDU, ACC0_ACC1, j_max, v_max, v_min, a_max, a_min)) {
                    add_profile();
                }
            }
        }
    }

    // :329-601
    void time_all_none_acc0_acc1(double v_max, double v_min, double a_max, double a_min, double j_max, bool return_after_found) {
        RKO_PHASE(RKO_PH_S1_QUARTIC);
        const double j_max_j_max_l = j_max * j_max;  // the local that shadows self.j_max_j_max in part of the reference
        // NONE UDDU / UDUD Strategy: t7 == 0 (equals UDDU)
        const double h2_none = (a0_a0 - af_af) / (2.0 * j_max) + (vf - v0);
        const double h2_h2 = h2_none * h2_none;
        const double t_min_none = (a0 - af) / j_max;
        const double t_max_none = (a_max - a_min) / j_max;

        double polynom_none[4];
        polynom_none[0] = 0.0;
        polynom_none[1] = -2.0 * (a0_a0 + af_af - 2.0 * j_max * (v0 + vf)) / (j_max_j_max_l);
        polynom_none[2] = 4.0 * (a0_p3 - af_p3 + 3.0 * j_max * (af * vf - a0 * v0)) / (3.0 * j_max_j_max_l * j_max) - 4.0 * pd / j_max;
        polynom_none[3] = -h2_h2 / (j_max_j_max_l);

        // ACC0
        const double h3_acc0 = (a0_a0 - af_af) / (2.0 * a_max * j_max) + (vf - v0) / a_max;
        const double t_min_acc0 = (a_max - af) / j_max;
        const double t_max_acc0 = (a_max - a_min) / j_max;

        const double h0_acc0 = 3.0 * (af_p4 - a0_p4)
            + 8.0 * (a0_p3 - af_p3) * a_max
            + 24.0 * a_max * j_max * (af * vf - a0 * v0)
            - 6.0 * a0_a0 * (a_max * a_max - 2.0 * j_max * v0)
            + 6.0 * af_af * (a_max * a_max - 2.0 * j_max * vf)
            + 12.0 * j_max * (j_max * (vf_vf - v0_v0 - 2.0 * a_max * pd) - a_max * a_max * (vf - v0));
        const double h2_acc0 = -af_af + a_max * a_max + 2.0 * j_max * vf;

        double polynom_acc0[4];
        polynom_acc0[0] = -2.0 * a_max / j_max;
        polynom_acc0[1] = h2_acc0 / (j_max_j_max_l);
        polynom_acc0[2] = 0.0;
        polynom_acc0[3] = h0_acc0 / (12.0 * j_max_j_max * j_max_j_max);

        // ACC1
        const double h3_acc1 = -(a0_a0 + af_af) / (2.0 * j_max * a_min) + a_min / j_max + (vf - v0) / a_min;
        const double t_min_acc1 = (a_min - a0) / j_max;
        const double t_max_acc1 = (a_max - a0) / j_max;

        const double h0_acc1 = (a0_p4 - af_p4) / 4.0
            + 2.0 * (af_p3 - a0_p3) * a_min / 3.0
            + (a0_a0 - af_af) * a_min * a_min / 2.0
            + j_max * (af_af * vf
                + a0_a0 * v0
                + 2.0 * a_min * (j_max * pd - a0 * v0 - af * vf)
                + a_min * a_min * (v0 + vf)
                + j_max * (v0_v0 - vf_vf));
        const double h2_acc1 = a0_a0 - a0 * a_min + 2.0 * j_max * v0;

        double polynom_acc1[4];
        polynom_acc1[0] = 2.0 * (2.0 * a0 - a_min) / j_max;
        polynom_acc1[1] = (5.0 * a0_a0 + a_min * (a_min - 6.0 * a0) + 2.0 * j_max * v0) / (j_max_j_max_l);
        polynom_acc1[2] = 2.0 * (a0 - a_min) * h2_acc1 / (j_max_j_max_l * j_max);
        polynom_acc1[3] = h0_acc1 / (j_max_j_max * j_max_j_max);

        double polynom_acc0_min[4] = {polynom_acc0[0], polynom_acc0[1], polynom_acc0[2], polynom_acc0[3]};
        polynom_acc0_min[0] += 4.0 * t_min_acc0;
        polynom_acc0_min[1] += (3.0 * polynom_acc0[0] + 6.0 * t_min_acc0) * t_min_acc0;
        polynom_acc0_min[2] += (2.0 * polynom_acc0[1] + (3.0 * polynom_acc0[0] + 4.0 * t_min_acc0) * t_min_acc0) * t_min_acc0;
        polynom_acc0_min[3] += (polynom_acc0[2] + (polynom_acc0[1] + (polynom_acc0[0] + t_min_acc0) * t_min_acc0) * t_min_acc0) * t_min_acc0;

        const bool polynom_acc0_has_solution = (polynom_acc0_min[0] < 0.0) || (polynom_acc0_min[1] < 0.0) || (polynom_acc0_min[2] < 0.0) || (polynom_acc0_min[3] <= 0.0);
        const bool polynom_acc1_has_solution = (polynom_acc1[0] < 0.0) || (polynom_acc1[1] < 0.0) || (polynom_acc1[2] < 0.0) || (polynom_acc1[3] <= 0.0);

        roots::PositiveSet<4> roots_none = roots::solve_quart_monic(polynom_none[0], polynom_none[1], polynom_none[2], polynom_none[3]);
        roots::PositiveSet<4> roots_acc0;
        if (polynom_acc0_has_solution) roots_acc0 = roots::solve_quart_monic(polynom_acc0[0], polynom_acc0[1], polynom_acc0[2], polynom_acc0[3]);
        roots::PositiveSet<4> roots_acc1;
        if (polynom_acc1_has_solution) roots_acc1 = roots::solve_quart_monic(polynom_acc1[0], polynom_acc1[1], polynom_acc1[2], polynom_acc1[3]);

        roots_none.sort();
        for (int ri = 0; ri < roots_none.size; ++ri) {
            double t = roots_none.data[ri];
            if (t < t_min_none || t > t_max_none) continue;

            // Single Newton-step (regarding pd)
            if (t > EPS) {
                const double h1 = j_max * t * t;
                const double orig = -h2_h2 / (4.0 * j_max * t)
                    + h2_none * (af / j_max + t)
                    + (4.0 * a0_p3 + 2.0 * af_p3
                        - 6.0 * a0_a0 * (af + 2.0 * j_max * t)
                        + 12.0 * (af - a0) * j_max * v0
                        + 3.0 * j_max_j_max * (-4.0 * pd + (h1 + 8.0 * v0) * t))
                        / (12.0 * j_max_j_max);
                const double deriv = h2_none + 2.0 * v0 - a0_a0 / j_max + h2_h2 / (4.0 * h1) + (3.0 * h1) / 4.0;

                const double delta_t = orig / deriv;
                t -= delta_t;
            }
            Profile* profile = &valid_profiles[current_index];
            const double h0 = h2_none / (2.0 * j_max * t);
            profile->t[0] = h0 + t / 2.0 - a0 / j_max;
            profile->t[1] = 0.0;
            profile->t[2] = t;
            profile->t[3] = 0.0;
            profile->t[4] = 0.0;
            profile->t[5] = 0.0;
            profile->t[6] = -h0 + t / 2.0 + af / j_max;

            if (profile->check_with_timing(UDDU, NONE, j_max, v_max, v_min, a_max, a_min)) {
                add_profile();
                if (return_after_found) return;
            }
        }

        roots_acc0.sort();
        for (int ri = 0; ri < roots_acc0.size; ++ri) {
            double t = roots_acc0.data[ri];
            if (t < t_min_acc0 || t > t_max_acc0) continue;

            // Single Newton step (regarding pd)
            if (t > EPS) {
                const double h1 = j_max * t;
                const double orig = h0_acc0 / (12.0 * j_max_j_max * t) + t * (h2_acc0 + h1 * (h1 - 2.0 * a_max));
                const double deriv = 2.0 * (h2_acc0 + h1 * (2.0 * h1 - 3.0 * a_max));

                const double delta_t = orig / deriv;
                t -= delta_t;
            }
            Profile* profile = &valid_profiles[current_index];
            profile->t[0] = (-a0 + a_max) / j_max;
            profile->t[1] = h3_acc0 - 2.0 * t + (j_max / a_max) * t * t;
            profile->t[2] = t;
            profile->t[3] = 0.0;
            profile->t[4] = 0.0;
            profile->t[5] = 0.0;
            profile->t[6] = (af - a_max) / j_max + t;

            if (profile->check_with_timing(UDDU, ACC0, j_max, v_max, v_min, a_max, a_min)) {
                add_profile();
                if (return_after_found) return;
            }
        }

        roots_acc1.sort();
        for (int ri = 0; ri < roots_acc1.size; ++ri) {
            double t = roots_acc1.data[ri];
            if (t < t_min_acc1 || t > t_max_acc1) continue;

            // Double Newton step (regarding pd)
            if (t > EPS) {
                const double h5 = a0_p3 + 2.0 * j_max * a0 * v0;
                double h1 = j_max * t;
                double orig = -(h0_acc1 / 2.0
                    + h1 * (h5
                        + a0 * (a_min - 2.0 * h1) * (a_min - h1)
                        + a0_a0 * (5.0 * h1 / 2.0 - 2.0 * a_min)
                        + a_min * a_min * h1 / 2.0
                        + j_max * (h1 / 2.0 - a_min) * (h1 * t + 2.0 * v0)))
                    / j_max;
                double deriv = (a_min - a0 - h1) * (h2_acc1 + h1 * (4.0 * a0 - a_min + 2.0 * h1));
                double delta_t = fmin_(orig / deriv, t);
                t -= delta_t;

                h1 = j_max * t;
                orig = -(h0_acc1 / 2.0
                    + h1 * (h5
                        + a0 * (a_min - 2.0 * h1) * (a_min - h1)
                        + a0_a0 * (5.0 * h1 / 2.0 - 2.0 * a_min)
                        + a_min * a_min * h1 / 2.0
                        + j_max * (h1 / 2.0 - a_min) * (h1 * t + 2.0 * v0)))
                    / j_max;

                if (std::fabs(orig) > 1e-9) {
                    deriv = (a_min - a0 - h1) * (h2_acc1 + h1 * (4.0 * a0 - a_min + 2.0 * h1));
                    delta_t = orig / deriv;
                    t -= delta_t;

                    h1 = j_max * t;
                    orig = -(h0_acc1 / 2.0
                        + h1 * (h5
                            + a0 * (a_min - 2.0 * h1) * (a_min - h1)
                            + a0_a0 * (5.0 * h1 / 2.0 - 2.0 * a_min)
                            + a_min * a_min * h1 / 2.0
                            + j_max * (h1 / 2.0 - a_min) * (h1 * t + 2.0 * v0)))
                        / j_max;

                    if (std::fabs(orig) > 1e-9) {
                        deriv = (a_min - a0 - h1) * (h2_acc1 + h1 * (4.0 * a0 - a_min + 2.0 * h1));
                        delta_t = orig / deriv;
                        t -= delta_t;
                    }
                }
            }
            Profile* profile = &valid_profiles[current_index];
            profile->t[0] = t;
            profile->t[1] = 0.0;
            profile->t[2] = (a0 - a_min) / j_max + t;
            profile->t[3] = 0.0;
            profile->t[4] = 0.0;
            profile->t[5] = h3_acc1 - (2.0 * a0 + j_max * t) * t / a_min;
            profile->t[6] = (af - a_min) / j_max;

            if (profile->check_with_timing(UDDU, ACC1, j_max, v_max, v_min, a_max, a_min)) {
                add_profile();
                if (return_after_found) return;
            }
        }
    }

    // :603-642
    void time_acc1_vel_two_step(double v_max, double v_min, double a_max, double a_min, double j_max) {
        RKO_PHASE(RKO_PH_S1_RESCUE);
        Profile* profile = &valid_profiles[current_index];
        profile->t[0] = 0.0;
        profile->t[1] = 0.0;
        profile->t[2] = a0 / j_max;
        profile->t[3] = -(3.0 * af_p4
            - 8.0 * a_min * (af_p3 - a0_p3)
            - 24.0 * a_min * j_max * (a0 * v0 - af * vf)
            + 6.0 * af_af * (a_min * a_min - 2.0 * j_max * vf)
            - 12.0 * j_max * (2. * a_min * j_max * pd
                + a_min * a_min * (vf + v_max)
                + j_max * (v_max * v_max - vf_vf)
                + a_min * a0 * (a0_a0 - 2.0 * j_max * (v0 + v_max)) / j_max))
            / (24.0 * a_min * j_max_j_max * v_max);
        profile->t[4] = -a_min / j_max;
        profile->t[5] = -(af_af / 2.0 - a_min * a_min + j_max * (v_max - vf)) / (a_min * j_max);
        profile->t[6] = profile->t[4] + af / j_max;

        if (profile->check_with_timing(UDDU, ACC1_VEL, j_max, v_max, v_min, a_max, a_min)) {
            add_profile();
        }
    }

    // :644-778
    void time_acc0_two_step(double v_max, double v_min, double a_max, double a_min, double j_max) {
        RKO_PHASE(RKO_PH_S1_RESCUE);
        // Two step
        {
            Profile* profile = &valid_profiles[current_index];
            profile->t[0] = 0.0;
            profile->t[1] = (af_af - a0_a0 + 2.0 * j_max * (vf - v0)) / (2.0 * a0 * j_max);
            profile->t[2] = (a0 - af) / j_max;
            profile->t[3] = 0.0;
            profile->t[4] = 0.0;
            profile->t[5] = 0.0;
            profile->t[6] = 0.0;

            if (profile->check_with_timing(UDDU, ACC0, j_max, v_max, v_min, a_max, a_min)) {
                add_profile();
                return;
            }
        }

        // Three step - Removed pf
        {
            Profile* profile = &valid_profiles[current_index];
            profile->t[0] = (-a0 + a_max) / j_max;
            profile->t[1] = (a0_a0 + af_af - 2.0 * a_max * a_max + 2.0 * j_max * (vf - v0)) / (2.0 * a_max * j_max);
            profile->t[2] = (-af + a_max) / j_max;
            profile->t[3] = 0.0;
            profile->t[4] = 0.0;
            profile->t[5] = 0.0;
            profile->t[6] = 0.0;

            if (profile->check_with_timing(UDDU, ACC0, j_max, v_max, v_min, a_max, a_min)) {
                add_profile();
                return;
            }
        }

        // Three step - Removed a_max
        {
            Profile* profile = &valid_profiles[current_index];
            const double h0 = 3.0 * (af_af - a0_a0 + 2.0 * j_max * (v0 + vf));
            const double h2 = a0_p3
                + 2.0 * af_p3
                + 6.0 * j_max_j_max * pd
                + 6.0 * (af - a0) * j_max * vf
                - 3.0 * a0 * af_af;
            const double h1 = std::sqrt(
                2.0 * (2.0 * h2 * h2
                    + h0 * (a0_p4 - 6.0 * a0_a0 * (af_af + 2.0 * j_max * vf)
                        + 8.0 * a0 * (af_p3 + 3.0 * j_max_j_max * pd + 3.0 * af * j_max * vf)
                        - 3.0 * (af_p4 + 4.0 * af_af * j_max * vf + 4.0 * j_max_j_max * (vf_vf - v0_v0))))
            ) * std::fabs(j_max) / j_max;
            profile->t[0] = (4.0 * af_p3 + 2.0 * a0_p3 - 6.0 * a0 * af_af
                + 12.0 * j_max_j_max * pd
                + 12.0 * (af - a0) * j_max * vf
                + h1)
                / (2.0 * j_max * h0);
            profile->t[1] = -h1 / (j_max * h0);
            profile->t[2] = (-4.0 * a0_p3 - 2.0 * af_p3
                + 6.0 * a0_a0 * af
                + 12.0 * j_max_j_max * pd
                - 12.0 * (af - a0) * j_max * v0
                + h1)
                / (2.0 * j_max * h0);
            profile->t[3] = 0.0;
            profile->t[4] = 0.0;
            profile->t[5] = 0.0;
            profile->t[6] = 0.0;

            if (profile->check_with_timing(UDDU, ACC0, j_max, v_max, v_min, a_max, a_min)) {
                add_profile();
                return;
            }
        }

        // Three step - t=(a_max - a_min)/j_max
        {
            const double t = (a_max - a_min) / j_max;
            Profile* profile = &valid_profiles[current_index];
            profile->t[0] = (-a0 + a_max) / j_max;
            profile->t[1] = (a0_a0 - af_af) / (2.0 * a_max * j_max) + (vf - v0 + j_max * t * t) / a_max - 2.0 * t;
            profile->t[2] = t;
            profile->t[3] = 0.0;
            profile->t[4] = 0.0;
            profile->t[5] = 0.0;
            profile->t[6] = (af - a_min) / j_max;

            if (profile->check_with_timing(UDDU, ACC0, j_max, v_max, v_min, a_max, a_min)) {
                add_profile();
            }
        }
    }

    // :780-841
    void time_vel_two_step(double v_max, double v_min, double a_max, double a_min, double j_max) {
        RKO_PHASE(RKO_PH_S1_RESCUE);
        const double h1 = std::sqrt(af_af / (2.0 * j_max_j_max) + (v_max - vf) / j_max);
        // Four step
        {
            // Solution 3/4
            Profile* profile = &valid_profiles[current_index];
            profile->t[0] = -a0 / j_max;
            profile->t[1] = 0.0;
            profile->t[2] = 0.0;
            profile->t[3] = (af_p3 - a0_p3) / (3.0 * j_max_j_max * v_max)
                + (a0 * v0 - af * vf + (af_af * h1) / 2.0) / (j_max * v_max)
                - (vf / v_max + 1.0) * h1
                + pd / v_max;
            profile->t[4] = h1;
            profile->t[5] = 0.0;
            profile->t[6] = h1 + af / j_max;

            if (profile->check_with_timing(UDDU, VEL, j_max, v_max, v_min, a_max, a_min)) {
                add_profile();
                return;
            }
        }

        // Four step
        {
            Profile* profile = &valid_profiles[current_index];
            profile->t[0] = 0.0;
            profile->t[1] = 0.0;
            profile->t[2] = a0 / j_max;
            profile->t[3] = (af_p3 - a0_p3) / (3.0 * j_max_j_max * v_max)
                + (a0 * v0 - af * vf + (af_af * h1 + a0_p3 / j_max) / 2.0) / (j_max * v_max)
                - (v0 / v_max + 1.0) * a0 / j_max
                - (vf / v_max + 1.0) * h1
                + pd / v_max;
            profile->t[4] = h1;
            profile->t[5] = 0.0;
            profile->t[6] = h1 + af / j_max;

            if (profile->check_with_timing(UDDU, VEL, j_max, v_max, v_min, a_max, a_min)) {
                add_profile();
            }
        }
    }

    // :843-895
    void time_none_two_step(double v_max, double v_min, double a_max, double a_min, double j_max) {
        RKO_PHASE(RKO_PH_S1_RESCUE);
        // Two step
        {
            Profile* profile = &valid_profiles[current_index];
            const double h0 = std::sqrt((a0_a0 + af_af) / 2.0 + j_max * (vf - v0)) * std::fabs(j_max) / j_max;
            profile->t[0] = (h0 - a0) / j_max;
            profile->t[1] = 0.0;
            profile->t[2] = (h0 - af) / j_max;
            profile->t[3] = 0.0;
            profile->t[4] = 0.0;
            profile->t[5] = 0.0;
            profile->t[6] = 0.0;

            if (profile->check_with_timing(UDDU, NONE, j_max, v_max, v_min, a_max, a_min)) {
                add_profile();
                return;
            }
        }

        // Single step
        {
            Profile* profile = &valid_profiles[current_index];
            profile->t[0] = (af - a0) / j_max;
            profile->t[1] = 0.0;
            profile->t[2] = 0.0;
            profile->t[3] = 0.0;
            profile->t[4] = 0.0;
            profile->t[5] = 0.0;
            profile->t[6] = 0.0;

            if (profile->check_with_timing(UDDU, NONE, j_max, v_max, v_min, a_max, a_min)) {
                add_profile();
            }
        }
    }

    // :897-980
    bool time_all_single_step(Profile* profile, double v_max, double v_min, double a_max, double a_min, double) {
        if (std::fabs(af - a0) > EPS) return false;

        for (int i = 0; i < 7; ++i) profile->t[i] = 0.0;

        if (std::fabs(a0) > EPS) {
            const double q = std::sqrt(2.0 * a0 * pd + v0_v0);

            // Solution 1
            profile->t[3] = (-v0 + q) / a0;
            if (profile->t[3] >= 0.0 && profile->check_with_timing(UDDU, NONE, 0.0, v_max, v_min, a_max, a_min)) return true;

            // Solution 2
            profile->t[3] = -(v0 + q) / a0;
            if (profile->t[3] >= 0.0 && profile->check_with_timing(UDDU, NONE, 0.0, v_max, v_min, a_max, a_min)) return true;
        } else if (std::fabs(v0) > EPS) {
            profile->t[3] = pd / v0;
            if (profile->check_with_timing(UDDU, NONE, 0.0, v_max, v_min, a_max, a_min)) return true;
        } else if (std::fabs(pd) < EPS && profile->check_with_timing(UDDU, NONE, 0.0, v_max, v_min, a_max, a_min)) {
            return true;
        }

        return false;
    }

    // :982-1263
    bool get_profile(const Profile& input, Block& block) {
        // Zero-limits special case (block.a / block.b are NOT cleared here, quirk Q14)
        if (_j_max == 0.0 || _a_max == 0.0 || _a_min == 0.0) {
            Profile& p = block.p_min;
            p.set_boundary_from_profile(input);

            if (time_all_single_step(&p, _v_max, _v_min, _a_max, _a_min, _j_max)) {
                block.t_min = p.t_sum[6] + p.brake.duration + p.accel.duration;
                if (std::fabs(v0) > EPS || std::fabs(a0) > EPS) {
                    block.a = Interval();
                    block.a.left = block.t_min;
                    block.a.right = INF;
                    block.has_a = true;
                }
                return true;
            }
            return false;
        }

        valid_profiles[0].set_boundary_from_profile(input);
        current_index = 0;

        if (std::fabs(vf) < EPS && std::fabs(af) < EPS) {
            const double v_max = (pd >= 0.0) ? _v_max : _v_min;
            const double v_min = (pd >= 0.0) ? _v_min : _v_max;
            const double a_max = (pd >= 0.0) ? _a_max : _a_min;
            const double a_min = (pd >= 0.0) ? _a_min : _a_max;
            const double j_max = (pd >= 0.0) ? _j_max : -_j_max;

            if (std::fabs(v0) < EPS && std::fabs(a0) < EPS && std::fabs(pd) < EPS) {
                time_all_none_acc0_acc1(v_max, v_min, a_max, a_min, j_max, true);
            } else {
                // There is no blocked interval when vf==0 && af==0, so return after first found profile
                time_all_vel(v_max, v_min, a_max, a_min, j_max, true);
                if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
                time_all_none_acc0_acc1(v_max, v_min, a_max, a_min, j_max, true);
                if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
                time_acc0_acc1(v_max, v_min, a_max, a_min, j_max, true);
                if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);

                time_all_vel(v_min, v_max, a_min, a_max, -j_max, true);
                if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
                time_all_none_acc0_acc1(v_min, v_max, a_min, a_max, -j_max, true);
                if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
                time_acc0_acc1(v_min, v_max, a_min, a_max, -j_max, true);
            }
        } else {
            time_all_none_acc0_acc1(_v_max, _v_min, _a_max, _a_min, _j_max, false);
            time_all_none_acc0_acc1(_v_min, _v_max, _a_min, _a_max, -_j_max, false);
            time_acc0_acc1(_v_max, _v_min, _a_max, _a_min, _j_max, false);
            time_acc0_acc1(_v_min, _v_max, _a_min, _a_max, -_j_max, false);
            time_all_vel(_v_max, _v_min, _a_max, _a_min, _j_max, false);
            time_all_vel(_v_min, _v_max, _a_min, _a_max, -_j_max, false);
        }

        if (current_index == 0) {
            time_none_two_step(_v_max, _v_min, _a_max, _a_min, _j_max);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
            time_none_two_step(_v_min, _v_max, _a_min, _a_max, -_j_max);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
            time_acc0_two_step(_v_max, _v_min, _a_max, _a_min, _j_max);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
            time_acc0_two_step(_v_min, _v_max, _a_min, _a_max, -_j_max);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
            time_vel_two_step(_v_max, _v_min, _a_max, _a_min, _j_max);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
            time_vel_two_step(_v_min, _v_max, _a_min, _a_max, -_j_max);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
            time_acc1_vel_two_step(_v_max, _v_min, _a_max, _a_min, _j_max);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
            time_acc1_vel_two_step(_v_min, _v_max, _a_min, _a_max, -_j_max);
        }

        return Block::calculate_block(block, valid_profiles, current_index);
    }
};

// ============================================== velocity_third_step1.rs
struct VelocityThirdOrderStep1 {
    double a0, af, _a_max, _a_min, _j_max, vd;
    Profile valid_profiles[6];
    int current_index = 0;

    // :24-35
    VelocityThirdOrderStep1(double v0, double a0_, double vf, double af_, double a_max, double a_min, double j_max)
        : a0(a0_), af(af_), _a_max(a_max), _a_min(a_min), _j_max(j_max), vd(vf - v0) {}

    // :38-44
    void add_profile() {
        if (current_index < 5) {
            current_index += 1;
            valid_profiles[current_index].set_boundary_from_profile(valid_profiles[current_index - 1]);
        }
    }

    // :46-62
    void time_acc0(double a_max, double a_min, double j_max, bool) {
        Profile* profile = &valid_profiles[current_index];
        profile->t[0] = (-a0 + a_max) / j_max;
        profile->t[1] = (a0 * a0 + af * af) / (2.0 * a_max * j_max) - a_max / j_max + vd / a_max;
        profile->t[2] = (-af + a_max) / j_max;
        profile->t[3] = 0.0;
        profile->t[4] = 0.0;
        profile->t[5] = 0.0;
        profile->t[6] = 0.0;

        if (profile->check_for_velocity(UDDU, ACC0, j_max, a_max, a_min)) add_profile();
    }

    // :64-117
    void time_none(double a_max, double a_min, double j_max, bool return_after_found) {
        double h1 = (a0 * a0 + af * af) / 2.0 + j_max * vd;
        if (h1 >= 0.0) {
            h1 = std::sqrt(h1);

            // Solution 1
            {
                Profile* profile = &valid_profiles[current_index];
                profile->t[0] = -(a0 + h1) / j_max;
                profile->t[1] = 0.0;
                profile->t[2] = -(af + h1) / j_max;
                profile->t[3] = 0.0;
                profile->t[4] = 0.0;
                profile->t[5] = 0.0;
                profile->t[6] = 0.0;

                if (profile->check_for_velocity(UDDU, NONE, j_max, a_max, a_min)) {
                    add_profile();
                    if (return_after_found) return;
                }
            }

            // Solution 2
            {
                Profile* profile = &valid_profiles[current_index];
                profile->t[0] = (-a0 + h1) / j_max;
                profile->t[1] = 0.0;
                profile->t[2] = (-af + h1) / j_max;
                profile->t[3] = 0.0;
                profile->t[4] = 0.0;
                profile->t[5] = 0.0;
                profile->t[6] = 0.0;

                if (profile->check_for_velocity(UDDU, NONE, j_max, a_max, a_min)) add_profile();
            }
        }
    }

    // :119-161
    bool time_all_single_step(Profile* profile, double a_max, double a_min, double) {
        if (std::fabs(af - a0) > EPS) return false;

        for (int i = 0; i < 7; ++i) profile->t[i] = 0.0;

        if (std::fabs(a0) > EPS) {
            profile->t[3] = vd / a0;
            if (profile->check_for_velocity(UDDU, NONE, 0.0, a_max, a_min)) return true;
        } else if (std::fabs(vd) < EPS && profile->check_for_velocity(UDDU, NONE, 0.0, a_max, a_min)) {
            return true;
        }
        return false;
    }

    // :163-242
    bool get_profile(const Profile& input, Block& block) {
        // Zero-limits special case
        if (_j_max == 0.0) {
            Profile& p = block.p_min;
            p.set_boundary_from_profile(input);

            if (time_all_single_step(&p, _a_max, _a_min, _j_max)) {
                block.t_min = p.t_sum[6] + p.brake.duration + p.accel.duration;
                if (std::fabs(a0) > EPS) {
                    block.a = Interval();
                    block.a.left = block.t_min;
                    block.a.right = INF;
                    block.has_a = true;
                }
                return true;
            }
            return false;
        }

        valid_profiles[0].set_boundary_from_profile(input);
        current_index = 0;

        if (std::fabs(af) < EPS) {
            // There is no blocked interval when af==0, so return after first found profile
            const double a_max = (vd >= 0.0) ? _a_max : _a_min;
            const double a_min = (vd >= 0.0) ? _a_min : _a_max;
            const double j_max = (vd >= 0.0) ? _j_max : -_j_max;

            time_none(a_max, a_min, j_max, true);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
            time_acc0(a_max, a_min, j_max, true);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);

            time_none(a_min, a_max, -j_max, true);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
            time_acc0(a_min, a_max, -j_max, true);
        } else {
            time_none(_a_max, _a_min, _j_max, false);
            time_none(_a_min, _a_max, -_j_max, false);
            time_acc0(_a_max, _a_min, _j_max, false);
            time_acc0(_a_min, _a_max, -_j_max, false);
        }

        return Block::calculate_block(block, valid_profiles, current_index);
    }
};

// ============================================== velocity_third_step2.rs
struct VelocityThirdOrderStep2 {
    double a0, tf, af, _a_max, _a_min, _j_max, vd, ad;

    // :15-36
    VelocityThirdOrderStep2(double tf_, double v0, double a0_, double vf, double af_, double a_max, double a_min, double j_max)
        : a0(a0_), tf(tf_), af(af_), _a_max(a_max), _a_min(a_min), _j_max(j_max), vd(vf - v0), ad(af_ - a0_) {}

    // :42-120
    bool time_acc0(Profile& profile, double a_max, double a_min, double j_max) {
        // UD Solution 1/2
        {
            const double h1 = std::sqrt((-ad * ad + 2.0 * j_max * ((a0 + af) * tf - 2.0 * vd)) / (j_max * j_max) + tf * tf);

            profile.t[0] = ad / (2.0 * j_max) + (tf - h1) / 2.0;
            profile.t[1] = h1;
            profile.t[2] = tf - (profile.t[0] + h1);
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = 0.0;

            if (profile.check_for_velocity(UDDU, ACC0, j_max, a_max, a_min)) {
                profile.pf = profile.p[7];
                return true;
            }
        }

        // UU Solution
        {
            const double h1 = -ad + j_max * tf;

            profile.t[0] = -ad * ad / (2.0 * j_max * h1) + (vd - a0 * tf) / h1;
            profile.t[1] = -ad / j_max + tf;
            profile.t[2] = 0.0;
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = tf - (profile.t[0] + profile.t[1]);

            if (profile.check_for_velocity(UDDU, ACC0, j_max, a_max, a_min)) {
                profile.pf = profile.p[7];
                return true;
            }
        }

        // UU Solution - 2 step
        {
            profile.t[0] = 0.0;
            profile.t[1] = -ad / j_max + tf;
            profile.t[2] = 0.0;
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = ad / j_max;

            if (profile.check_for_velocity(UDDU, ACC0, j_max, a_max, a_min)) {
                profile.pf = profile.p[7];
                return true;
            }
        }

        return false;
    }

    // :122-174 (jf is not checked against j_max, quirk Q10)
    bool time_none(Profile& profile, double a_max, double a_min, double j_max) {
        if (std::fabs(a0) < EPS && std::fabs(af) < EPS && std::fabs(vd) < EPS) {
            profile.t[0] = 0.0;
            profile.t[1] = tf;
            profile.t[2] = 0.0;
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = 0.0;

            if (profile.check_for_velocity(UDDU, NONE, j_max, a_max, a_min)) {
                profile.pf = profile.p[7];
                return true;
            }
        }

        // UD Solution 1/2
        {
            const double h1 = 2.0 * (af * tf - vd);

            profile.t[0] = h1 / ad;
            profile.t[1] = tf - profile.t[0];
            profile.t[2] = 0.0;
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = 0.0;

            const double jf = ad * ad / h1;

            if (profile.check_for_velocity(UDDU, NONE, jf, a_max, a_min)) {
                profile.pf = profile.p[7];
                return true;
            }
        }

        return false;
    }

    // :176-179
    bool check_all(Profile& profile, double a_max, double a_min, double j_max) {
        return time_acc0(profile, a_max, a_min, j_max) || time_none(profile, a_max, a_min, j_max);
    }

    // :181-191
    bool get_profile(Profile& profile) {
        if (vd > 0.0) {
            return check_all(profile, _a_max, _a_min, _j_max) || check_all(profile, _a_min, _a_max, -_j_max);
        }
        return check_all(profile, _a_min, _a_max, -_j_max) || check_all(profile, _a_max, _a_min, _j_max);
    }
};

// ============================================== position_second_step1.rs
struct PositionSecondOrderStep1 {
    double v0, vf, _v_max, _v_min, _a_max, _a_min, pd;
    Profile valid_profiles[6];
    int current_index = 0;

    // :25-46
    PositionSecondOrderStep1(double p0, double v0_, double pf, double vf_, double v_max, double v_min, double a_max, double a_min)
        : v0(v0_), vf(vf_), _v_max(v_max), _v_min(v_min), _a_max(a_max), _a_min(a_min), pd(pf - p0) {}

    // :49-55 (stores at most 3, quirk Q12)
    void add_profile(const Profile& profile) {
        if (current_index < 3) {
            valid_profiles[current_index] = profile;
            current_index += 1;
            valid_profiles[current_index].set_boundary_from_profile(profile);
        }
    }

    // :57-88
    void time_acc0(Profile& profile, double v_max, double v_min, double a_max, double a_min, bool) {
        profile.t[0] = (-v0 + v_max) / a_max;
        profile.t[1] = (a_min * v0 * v0 - a_max * vf * vf) / (2.0 * a_max * a_min * v_max)
            + v_max * (a_max - a_min) / (2.0 * a_max * a_min)
            + pd / v_max;
        profile.t[2] = (vf - v_max) / a_min;
        profile.t[3] = 0.0;
        profile.t[4] = 0.0;
        profile.t[5] = 0.0;
        profile.t[6] = 0.0;

        if (profile.check_for_second_order(UDDU, ACC0, a_max, a_min, v_max, v_min)) add_profile(profile);
    }

    // :90-154
    void time_none(double v_max, double v_min, double a_max, double a_min, bool return_after_found) {
        double h1 = (a_max * vf * vf - a_min * v0 * v0 - 2.0 * a_max * a_min * pd) / (a_max - a_min);
        if (h1 >= 0.0) {
            h1 = std::sqrt(h1);

            // Solution 1
            {
                Profile profile = valid_profiles[current_index];
                profile.t[0] = -(v0 + h1) / a_max;
                profile.t[1] = 0.0;
                profile.t[2] = (vf + h1) / a_min;
                profile.t[3] = 0.0;
                profile.t[4] = 0.0;
                profile.t[5] = 0.0;
                profile.t[6] = 0.0;

                if (profile.check_for_second_order(UDDU, NONE, a_max, a_min, v_max, v_min)) {
                    add_profile(profile);
                    if (return_after_found) return;
                }
            }

            // Solution 2
            {
                Profile profile = valid_profiles[current_index];
                profile.t[0] = (-v0 + h1) / a_max;
                profile.t[1] = 0.0;
                profile.t[2] = (vf - h1) / a_min;
                profile.t[3] = 0.0;
                profile.t[4] = 0.0;
                profile.t[5] = 0.0;
                profile.t[6] = 0.0;

                if (profile.check_for_second_order(UDDU, NONE, a_max, a_min, v_max, v_min)) add_profile(profile);
            }
        }
    }

    // :155-210
    bool time_all_single_step(Profile* profile, double v_max, double v_min, double, double) {
        if (std::fabs(vf - v0) > EPS) return false;

        for (int i = 0; i < 7; ++i) profile->t[i] = 0.0;

        if (profile->check_for_second_order(UDDU, NONE, 0.0, 0.0, v_max, v_min)) return true;

        if (std::fabs(v0) > EPS) {
            profile->t[3] = pd / v0;
            if (profile->check_for_second_order(UDDU, NONE, 0.0, 0.0, v_max, v_min)) return true;
        } else if (std::fabs(pd) < EPS && profile->check_for_second_order(UDDU, NONE, 0.0, 0.0, v_max, v_min)) {
            return true;
        }
        return false;
    }

    // :212-312
    bool get_profile(const Profile& input, Block& block) {
        // Zero-limits special case
        if (_v_max == 0.0 && _v_min == 0.0) {
            Profile& p = block.p_min;
            p.set_boundary_from_profile(input);

            if (time_all_single_step(&p, _v_max, _v_min, _a_max, _a_min)) {
                block.t_min = p.t_sum[6] + p.brake.duration + p.accel.duration;
                if (std::fabs(v0) > EPS) {
                    block.a = Interval();
                    block.a.left = block.t_min;
                    block.a.right = INF;
                    block.has_a = true;
                }
                return true;
            }
            return false;
        }

        valid_profiles[0].set_boundary_from_profile(input);
        current_index = 0;
        Profile profile = valid_profiles[0];

        if (std::fabs(vf) < EPS) {
            // There is no blocked interval when vf==0, so return after first found profile
            const double v_max = (pd >= 0.0) ? _v_max : _v_min;
            const double v_min = (pd >= 0.0) ? _v_min : _v_max;
            const double a_max = (pd >= 0.0) ? _a_max : _a_min;
            const double a_min = (pd >= 0.0) ? _a_min : _a_max;

            time_none(v_max, v_min, a_max, a_min, true);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
            time_acc0(profile, v_max, v_min, a_max, a_min, true);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);

            time_none(v_min, v_max, a_min, a_max, true);
            if (current_index > 0) return Block::calculate_block(block, valid_profiles, current_index);
            time_acc0(profile, v_min, v_max, a_min, a_max, true);
        } else {
            time_none(_v_max, _v_min, _a_max, _a_min, false);
            time_none(_v_min, _v_max, _a_min, _a_max, false);
            time_acc0(profile, _v_max, _v_min, _a_max, _a_min, false);
            time_acc0(profile, _v_min, _v_max, _a_min, _a_max, false);
        }

        return Block::calculate_block(block, valid_profiles, current_index);
    }
};

// ============================================== position_second_step2.rs
struct PositionSecondOrderStep2 {
    double v0, tf, vf, _v_max, _v_min, _a_max, _a_min, pd, vd;

    // :24-46
    PositionSecondOrderStep2(double tf_, double p0, double v0_, double pf, double vf_, double v_max, double v_min, double a_max, double a_min)
        : v0(v0_), tf(tf_), vf(vf_), _v_max(v_max), _v_min(v_min), _a_max(a_max), _a_min(a_min), pd(pf - p0), vd(vf_ - v0_) {}

    // :48-139
    bool time_acc0(Profile& profile, double v_max, double v_min, double a_max, double a_min, bool) {
        // UD Solution 1/2
        {
            const double h1 = std::sqrt((2.0 * a_max * (pd - tf * vf) - 2.0 * a_min * (pd - tf * v0) + vd * vd) / (a_max * a_min) + tf * tf);

            profile.t[0] = (a_max * vd - a_max * a_min * (tf - h1)) / (a_max * (a_max - a_min));
            profile.t[1] = h1;
            profile.t[2] = tf - (profile.t[0] + h1);
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = 0.0;

            if (profile.check_for_second_order(UDDU, ACC0, a_max, a_min, v_max, v_min)) {
                profile.pf = profile.p[7];
                return true;
            }
        }

        // UU Solution
        {
            const double h1 = -vd + a_max * tf;

            profile.t[0] = -vd * vd / (2.0 * a_max * h1) + (pd - v0 * tf) / h1;
            profile.t[1] = -vd / a_max + tf;
            profile.t[2] = 0.0;
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = tf - (profile.t[0] + profile.t[1]);

            if (profile.check_for_second_order(UDDU, ACC0, a_max, a_min, v_max, v_min)) {
                profile.pf = profile.p[7];
                return true;
            }
        }

        // UU Solution - 2 step
        {
            profile.t[0] = 0.0;
            profile.t[1] = -vd / a_max + tf;
            profile.t[2] = 0.0;
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = vd / a_max;

            if (profile.check_for_second_order(UDDU, ACC0, a_max, a_min, v_max, v_min)) {
                profile.pf = profile.p[7];
                return true;
            }
        }

        return false;
    }

    // :141-203
    bool time_none(Profile& profile, double v_max, double v_min, double a_max, double a_min, bool) {
        if (std::fabs(v0) < EPS && std::fabs(vf) < EPS && std::fabs(pd) < EPS) {
            profile.t[0] = 0.0;
            profile.t[1] = tf;
            profile.t[2] = 0.0;
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = 0.0;

            if (profile.check_for_second_order(UDDU, NONE, a_max, a_min, v_max, v_min)) {
                profile.pf = profile.p[7];
                return true;
            }
        }

        // UD Solution 1/2
        {
            const double h1 = 2.0 * (vf * tf - pd);

            profile.t[0] = h1 / vd;
            profile.t[1] = tf - profile.t[0];
            profile.t[2] = 0.0;
            profile.t[3] = 0.0;
            profile.t[4] = 0.0;
            profile.t[5] = 0.0;
            profile.t[6] = 0.0;

            const double af = vd * vd / h1;

            if ((a_min - 1e-12 < af) && (af < a_max + 1e-12) && profile.check_for_second_order(UDDU, NONE, af, -af, v_max, v_min)) {
                profile.pf = profile.p[7];
                return true;
            }
        }

        return false;
    }

    // :206-216
    bool check_all(Profile& profile, double v_max, double v_min, double a_max, double a_min) {
        return time_acc0(profile, v_max, v_min, a_max, a_min, false) || time_none(profile, v_max, v_min, a_max, a_min, false);
    }

    // :219-229
    bool get_profile(Profile& profile) {
        if (pd > 0.0) {
            return check_all(profile, _v_max, _v_min, _a_max, _a_min) || check_all(profile, _v_min, _v_max, _a_min, _a_max);
        }
        return check_all(profile, _v_min, _v_max, _a_min, _a_max) || check_all(profile, _v_max, _v_min, _a_max, _a_min);
    }
};

// ============================================== velocity_second_step1.rs:23-46
struct VelocitySecondOrderStep1 {
    double _a_max, _a_min, vd;
    VelocitySecondOrderStep1(double v0, double vf, double a_max, double a_min) : _a_max(a_max), _a_min(a_min), vd(vf - v0) {}

    // block.a / block.b are not cleared here (quirk Q14)
    bool get_profile(const Profile& input, Block& block) {
        Profile& p = block.p_min;
        p.set_boundary_from_profile(input);

        const double af = (vd > 0.0) ? _a_max : _a_min;
        p.t[0] = 0.0;
        p.t[1] = vd / af;
        p.t[2] = 0.0;
        p.t[3] = 0.0;
        p.t[4] = 0.0;
        p.t[5] = 0.0;
        p.t[6] = 0.0;

        if (p.check_for_second_order_velocity(UDDU, ACC0, af)) {
            block.t_min = p.t_sum[6] + p.brake.duration + p.accel.duration;
            return true;
        }
        return false;
    }
};

// ============================================== velocity_second_step2.rs:23-46
struct VelocitySecondOrderStep2 {
    double tf, _a_max, _a_min, vd;
    VelocitySecondOrderStep2(double tf_, double v0, double vf, double a_max, double a_min) : tf(tf_), _a_max(a_max), _a_min(a_min), vd(vf - v0) {}

    bool get_profile(Profile& profile) {
        const double af = vd / tf;
        profile.t[0] = 0.0;
        profile.t[1] = tf;
        profile.t[2] = 0.0;
        profile.t[3] = 0.0;
        profile.t[4] = 0.0;
        profile.t[5] = 0.0;
        profile.t[6] = 0.0;

        if (profile.check_for_second_order_velocity_with_timing_a_limits(UDDU, ACC0, tf, af, _a_max, _a_min)) {
            profile.pf = profile.p[7];
            return true;
        }
        return false;
    }
};

// ============================================== position_first_step1.rs:20-42
struct PositionFirstOrderStep1 {
    double _v_max, _v_min, pd;
    PositionFirstOrderStep1(double p0, double pf, double v_max, double v_min) : _v_max(v_max), _v_min(v_min), pd(pf - p0) {}

    bool get_profile(const Profile& input, Block& block) {
        Profile& p = block.p_min;
        p.set_boundary_from_profile(input);

        const double vf = (pd > 0.0) ? _v_max : _v_min;
        p.t[0] = 0.0;
        p.t[1] = 0.0;
        p.t[2] = 0.0;
        p.t[3] = pd / vf;
        p.t[4] = 0.0;
        p.t[5] = 0.0;
        p.t[6] = 0.0;

        if (p.check_for_first_order(vf, UDDU, VEL)) {
            block.t_min = p.t_sum[6] + p.brake.duration + p.accel.duration;
            return true;
        }
        return false;
    }
};

// ============================================== position_first_step2.rs:22-34
struct PositionFirstOrderStep2 {
    double tf, _v_max, _v_min, pd;
    PositionFirstOrderStep2(double tf_, double p0, double pf, double v_max, double v_min) : tf(tf_), _v_max(v_max), _v_min(v_min), pd(pf - p0) {}

    bool get_profile(Profile& profile) {
        const double vf = pd / tf;

        profile.t[0] = 0.0;
        profile.t[1] = 0.0;
        profile.t[2] = 0.0;
        profile.t[3] = tf;
        profile.t[4] = 0.0;
        profile.t[5] = 0.0;
        profile.t[6] = 0.0;

        return profile.check_for_first_order(vf, UDDU, NONE);
    }
};

}  // namespace rko
