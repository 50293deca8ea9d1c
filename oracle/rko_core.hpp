// rko_core.hpp — parity oracle, numeric core: Profile, BrakeProfile, roots, Block.
//
// TEST INFRASTRUCTURE ONLY (see rko_math.hpp). CPU restatement of rsruckig
// v2.1.2; every function cites the reference file:line it follows. Arithmetic
// keeps the reference's association order; compile with -ffp-contract=off.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <limits>
#include <utility>
#include <vector>

#include "rko_math.hpp"

namespace rko {

constexpr double EPS = DBL_EPSILON;  // f64::EPSILON
constexpr double INF = std::numeric_limits<double>::infinity();

// profile.rs:14-22
constexpr double V_EPS = 1e-12, A_EPS = 1e-12, J_EPS = 1e-12;
constexpr double P_PRECISION = 1e-8, V_PRECISION = 1e-8, A_PRECISION = 1e-10;
constexpr double T_MAX = 1e12;

// profile.rs:25-49 (discriminant order kept)
enum ReachedLimits : uint8_t { ACC0_ACC1_VEL = 0, VEL, ACC0, ACC1, ACC0_ACC1, ACC0_VEL, ACC1_VEL, NONE };
enum Direction : uint8_t { UP = 0, DOWN };
enum ControlSigns : uint8_t { UDDU = 0, UDUD };

// util.rs:27-33
static inline void integrate(double t, double p0, double v0, double a0, double j, double& p, double& v, double& a) {
    p = p0 + t * (v0 + t * (a0 / 2.0 + t * j / 6.0));
    v = v0 + t * (a0 + t * j / 2.0);
    a = a0 + t * j;
}

// ------------------------------------------------------------------ brake.rs
struct BrakeProfile {
    double duration = 0.0;
    double t[2] = {0, 0}, j[2] = {0, 0}, a[2] = {0, 0}, v[2] = {0, 0}, p[2] = {0, 0};

    static constexpr double BEPS = 2.2e-14;  // brake.rs:12

    static double v_at_t(double v0, double a0, double j, double t) { return v0 + t * (a0 + j * t / 2.0); }  // :15
    static double v_at_a_zero(double v0, double a0, double j) { return v0 + (a0 * a0) / (2.0 * j); }        // :20

    // brake.rs:46-75
    void acceleration_brake(double v0, double a0, double v_max, double v_min, double a_max, double a_min, double j_max) {
        j[0] = -j_max;

        const double t_to_a_max = (a0 - a_max) / j_max;
        const double t_to_a_zero = a0 / j_max;

        const double v_at_a_max = v_at_t(v0, a0, -j_max, t_to_a_max);
        const double v_at_a_zero_ = v_at_t(v0, a0, -j_max, t_to_a_zero);

        if ((v_at_a_zero_ > v_max && j_max > 0.0) || (v_at_a_zero_ < v_max && j_max < 0.0)) {
            velocity_brake(v0, a0, v_max, v_min, a_max, a_min, j_max);
        } else if ((v_at_a_max < v_min && j_max > 0.0) || (v_at_a_max > v_min && j_max < 0.0)) {
            const double t_to_v_min = -(v_at_a_max - v_min) / a_max;
            const double t_to_v_max = -a_max / (2.0 * j_max) - (v_at_a_max - v_max) / a_max;

            t[0] = t_to_a_max + BEPS;
            t[1] = fmin_(t_to_v_min, fmax_(t_to_v_max - BEPS, 0.0));
        } else {
            t[0] = t_to_a_max + BEPS;
        }
    }

    // brake.rs:77-105
    void velocity_brake(double v0, double a0, double v_max, double v_min, double /*a_max*/, double a_min, double j_max) {
        j[0] = -j_max;
        const double t_to_a_min = (a0 - a_min) / j_max;
        const double t_to_v_max = a0 / j_max + (std::sqrt(a0 * a0 + 2.0 * j_max * (v0 - v_max))) / std::fabs(j_max);
        const double t_to_v_min = a0 / j_max + (std::sqrt(a0 * a0 / 2.0 + j_max * (v0 - v_min))) / std::fabs(j_max);
        const double t_min_to_v_max = fmin_(t_to_v_max, t_to_v_min);

        if (t_to_a_min < t_min_to_v_max) {
            const double v_at_a_min = v_at_t(v0, a0, -j_max, t_to_a_min);
            const double t_to_v_max_with_constant = -(v_at_a_min - v_max) / a_min;
            const double t_to_v_min_with_constant = a_min / (2.0 * j_max) - (v_at_a_min - v_min) / a_min;

            t[0] = fmax_(t_to_a_min - BEPS, 0.0);
            t[1] = fmax_(fmin_(t_to_v_max_with_constant, t_to_v_min_with_constant), 0.0);
        } else {
            t[0] = fmax_(t_min_to_v_max - BEPS, 0.0);
        }
    }

    // brake.rs:107-139
    void get_position_brake_trajectory(double v0, double a0, double v_max, double v_min, double a_max, double a_min, double j_max) {
        t[0] = 0.0; t[1] = 0.0; j[0] = 0.0; j[1] = 0.0;

        if (j_max == 0.0 || a_max == 0.0 || a_min == 0.0) return;

        if (a0 > a_max) {
            acceleration_brake(v0, a0, v_max, v_min, a_max, a_min, j_max);
        } else if (a0 < a_min) {
            acceleration_brake(v0, a0, v_min, v_max, a_min, a_max, -j_max);
        } else if ((v0 > v_max && v_at_a_zero(v0, a0, -j_max) > v_min) || (a0 > 0.0 && v_at_a_zero(v0, a0, j_max) > v_max)) {
            velocity_brake(v0, a0, v_max, v_min, a_max, a_min, j_max);
        } else if ((v0 < v_min && v_at_a_zero(v0, a0, j_max) < v_max) || (a0 < 0.0 && v_at_a_zero(v0, a0, -j_max) < v_min)) {
            velocity_brake(v0, a0, v_min, v_max, a_min, a_max, -j_max);
        }
    }

    // brake.rs:141-167
    void get_second_order_position_brake_trajectory(double v0, double v_max, double v_min, double a_max, double a_min) {
        t[0] = 0.0; t[1] = 0.0; j[0] = 0.0; j[1] = 0.0; a[0] = 0.0; a[1] = 0.0;

        if (a_max == 0.0 || a_min == 0.0) return;

        if (v0 > v_max) {
            a[0] = a_min;
            t[0] = (v_max - v0) / a_min + EPS;
        } else if (v0 < v_min) {
            a[0] = a_max;
            t[0] = (v_min - v0) / a_max + EPS;
        }
    }

    // brake.rs:169-186
    void get_velocity_brake_trajectory(double a0, double a_max, double a_min, double j_max) {
        t[0] = 0.0; t[1] = 0.0; j[0] = 0.0; j[1] = 0.0;

        if (j_max == 0.0) return;

        if (a0 > a_max) {
            j[0] = -j_max;
            t[0] = (a0 - a_max) / j_max + EPS;
        } else if (a0 < a_min) {
            j[0] = j_max;
            t[0] = -(a0 - a_min) / j_max + EPS;
        }
    }

    // brake.rs:188-193
    void get_second_order_velocity_brake_trajectory() { t[0] = 0.0; t[1] = 0.0; j[0] = 0.0; j[1] = 0.0; }

    // brake.rs:195-220
    void finalize(double& ps, double& vs, double& as) {
        if (t[0] <= 0.0 && t[1] <= 0.0) {
            duration = 0.0;
            return;
        }

        duration = t[0];
        p[0] = ps; v[0] = vs; a[0] = as;
        integrate(t[0], ps, vs, as, j[0], ps, vs, as);

        if (t[1] > 0.0) {
            duration += t[1];
            p[1] = ps; v[1] = vs; a[1] = as;
            integrate(t[1], ps, vs, as, j[1], ps, vs, as);
        }
    }

    // brake.rs:222-235
    void finalize_second_order(double& ps, double& vs, double& as) {
        if (t[0] <= 0.0) {
            duration = 0.0;
            return;
        }

        duration = t[0];
        p[0] = ps; v[0] = vs;
        integrate(t[0], ps, vs, a[0], 0.0, ps, vs, as);
    }
};

// ---------------------------------------------------------------- profile.rs
struct Profile {
    double t[7] = {0}, t_sum[7] = {0}, j[7] = {0};
    double a[8] = {0}, v[8] = {0}, p[8] = {0};
    BrakeProfile brake, accel;
    double pf = 0.0, vf = 0.0, af = 0.0;
    ReachedLimits limits = NONE;
    Direction direction = UP;
    ControlSigns control_signs = UDDU;

    // profile.rs:323-482 (set_limits is always false in the reference, :498)
    bool check(ControlSigns control_signs_, ReachedLimits limits_, double jf, double v_max, double v_min, double a_max, double a_min) {
        RKO_PHASE(RKO_PH_CHECK);
        if (t[0] < 0.0) return false;

        t_sum[0] = t[0];
        for (int i = 0; i < 6; ++i) {
            if (t[i + 1] < 0.0) return false;
            t_sum[i + 1] = t_sum[i] + t[i + 1];
        }

        if ((limits_ == ACC0_ACC1_VEL || limits_ == ACC0_VEL || limits_ == ACC1_VEL || limits_ == VEL) && t[3] < EPS) return false;
        if ((limits_ == ACC0 || limits_ == ACC0_ACC1) && t[1] < EPS) return false;
        if ((limits_ == ACC1 || limits_ == ACC0_ACC1) && t[5] < EPS) return false;
        if (t_sum[6] > T_MAX) return false;

        if (control_signs_ == UDDU) {
            j[0] = (t[0] > 0.0 ? jf : 0.0); j[1] = 0.0; j[2] = (t[2] > 0.0 ? -jf : 0.0); j[3] = 0.0;
            j[4] = (t[4] > 0.0 ? -jf : 0.0); j[5] = 0.0; j[6] = (t[6] > 0.0 ? jf : 0.0);
        } else {
            j[0] = (t[0] > 0.0 ? jf : 0.0); j[1] = 0.0; j[2] = (t[2] > 0.0 ? -jf : 0.0); j[3] = 0.0;
            j[4] = (t[4] > 0.0 ? jf : 0.0); j[5] = 0.0; j[6] = (t[6] > 0.0 ? -jf : 0.0);
        }

        direction = (v_max > 0.0) ? UP : DOWN;

        const double v_upp_lim = (direction == UP ? v_max : v_min) + V_EPS;
        const double v_low_lim = (direction == UP ? v_min : v_max) - V_EPS;

        for (int i = 0; i < 7; ++i) {
            a[i + 1] = a[i] + t[i] * j[i];
            v[i + 1] = v[i] + t[i] * (a[i] + t[i] * j[i] / 2.0);
            p[i + 1] = p[i] + t[i] * (v[i] + t[i] * (a[i] / 2.0 + t[i] * j[i] / 6.0));

            if ((limits_ == ACC0_ACC1_VEL || limits_ == ACC0_ACC1 || limits_ == ACC0_VEL || limits_ == ACC1_VEL || limits_ == VEL) && i == 2) {
                a[3] = 0.0;
            }

            if (i > 1 && a[i + 1] * a[i] < -EPS) {
                const double v_a_zero = v[i] - (a[i] * a[i]) / (2.0 * j[i]);
                if (v_a_zero > v_upp_lim || v_a_zero < v_low_lim) return false;
            }
        }

        control_signs = control_signs_;
        limits = limits_;

        const double a_upp_lim = (direction == UP ? a_max : a_min) + A_EPS;
        const double a_low_lim = (direction == UP ? a_min : a_max) - A_EPS;

        return std::fabs(p[7] - pf) < P_PRECISION && std::fabs(v[7] - vf) < V_PRECISION && std::fabs(a[7] - af) < A_PRECISION
            && a[1] >= a_low_lim && a[1] <= a_upp_lim && a[3] >= a_low_lim && a[3] <= a_upp_lim && a[5] >= a_low_lim && a[5] <= a_upp_lim
            && v[3] <= v_upp_lim && v[3] >= v_low_lim && v[4] <= v_upp_lim && v[4] >= v_low_lim
            && v[5] <= v_upp_lim && v[5] >= v_low_lim && v[6] <= v_upp_lim && v[6] >= v_low_lim;
    }

    // profile.rs:485-499
    bool check_with_timing(ControlSigns cs, ReachedLimits lim, double jf, double v_max, double v_min, double a_max, double a_min) {
        return check(cs, lim, jf, v_max, v_min, a_max, a_min);
    }

    // profile.rs:502-516
    bool check_with_timing_full(ControlSigns cs, ReachedLimits lim, double /*tf*/, double jf, double v_max, double v_min, double a_max, double a_min, double j_max) {
        return (std::fabs(jf) < std::fabs(j_max) + J_EPS) && check_with_timing(cs, lim, jf, v_max, v_min, a_max, a_min);
    }

    // profile.rs:121-207 (note the UDUD pattern ends with +jf here, quirk Q7)
    bool check_for_velocity(ControlSigns control_signs_, ReachedLimits limits_, double jf, double a_max, double a_min) {
        if (t[0] < 0.0) return false;

        t_sum[0] = t[0];
        for (int i = 0; i < 6; ++i) {
            if (t[i + 1] < 0.0) return false;
            t_sum[i + 1] = t_sum[i] + t[i + 1];
        }

        if (limits_ == ACC0 && t[1] < EPS) return false;
        if (t_sum[6] > T_MAX) return false;

        if (control_signs_ == UDDU) {
            j[0] = (t[0] > 0.0 ? jf : 0.0); j[1] = 0.0; j[2] = (t[2] > 0.0 ? -jf : 0.0); j[3] = 0.0;
            j[4] = (t[4] > 0.0 ? -jf : 0.0); j[5] = 0.0; j[6] = (t[6] > 0.0 ? jf : 0.0);
        } else {
            j[0] = (t[0] > 0.0 ? jf : 0.0); j[1] = 0.0; j[2] = (t[2] > 0.0 ? -jf : 0.0); j[3] = 0.0;
            j[4] = (t[4] > 0.0 ? jf : 0.0); j[5] = 0.0; j[6] = (t[6] > 0.0 ? jf : 0.0);
        }

        for (int i = 0; i < 7; ++i) {
            a[i + 1] = a[i] + t[i] * j[i];
            v[i + 1] = v[i] + t[i] * (a[i] + t[i] * j[i] / 2.0);
            p[i + 1] = p[i] + t[i] * (v[i] + t[i] * (a[i] / 2.0 + t[i] * j[i] / 6.0));
        }

        control_signs = control_signs_;
        limits = limits_;

        direction = (a_max > 0.0) ? UP : DOWN;
        const double a_upp_lim = (direction == UP ? a_max : a_min) + A_EPS;
        const double a_low_lim = (direction == UP ? a_min : a_max) - A_EPS;

        return std::fabs(v[7] - vf) < V_PRECISION && std::fabs(a[7] - af) < A_PRECISION
            && a[1] >= a_low_lim && a[3] >= a_low_lim && a[5] >= a_low_lim
            && a[1] <= a_upp_lim && a[3] <= a_upp_lim && a[5] <= a_upp_lim;
    }

    // profile.rs:210-221
    bool check_for_velocity_with_timing(double /*tf*/, ControlSigns cs, ReachedLimits lim, double jf, double a_max, double a_min) {
        return check_for_velocity(cs, lim, jf, a_max, a_min);
    }

    // profile.rs:224-236
    bool check_for_velocity_with_timing_full(double tf, ControlSigns cs, ReachedLimits lim, double jf, double a_max, double a_min, double j_max) {
        return std::fabs(jf) < std::fabs(j_max) + J_EPS && check_for_velocity_with_timing(tf, cs, lim, jf, a_max, a_min);
    }

    // profile.rs:239-252
    void set_boundary_for_velocity(double p0_new, double v0_new, double a0_new, double vf_new, double af_new) {
        a[0] = a0_new; v[0] = v0_new; p[0] = p0_new; af = af_new; vf = vf_new;
    }

    // profile.rs:256-294
    bool check_for_second_order_velocity(ControlSigns control_signs_, ReachedLimits limits_, double a_up) {
        if (t[1] < 0.0) return false;

        t_sum[0] = 0.0;
        for (int i = 1; i < 7; ++i) t_sum[i] = t[1];
        if (t_sum[6] > T_MAX) return false;

        for (int i = 0; i < 7; ++i) j[i] = 0.0;
        for (int i = 0; i < 8; ++i) a[i] = 0.0;
        a[1] = (t[1] > 0.0) ? a_up : 0.0;
        a[7] = af;

        for (int i = 0; i < 7; ++i) {
            v[i + 1] = v[i] + t[i] * a[i];
            p[i + 1] = p[i] + t[i] * (v[i] + t[i] * a[i] / 2.0);
        }

        control_signs = control_signs_;
        limits = limits_;
        direction = (a_up > 0.0) ? UP : DOWN;

        return std::fabs(v[7] - vf) < V_PRECISION;
    }

    // profile.rs:297-305
    bool check_for_second_order_velocity_with_timing(double /*tf*/, ControlSigns cs, ReachedLimits lim, double a_up) {
        return check_for_second_order_velocity(cs, lim, a_up);
    }

    // profile.rs:308-320
    bool check_for_second_order_velocity_with_timing_a_limits(ControlSigns cs, ReachedLimits lim, double tf, double a_up, double a_max, double a_min) {
        return a_min - A_EPS < a_up && a_up < a_max + A_EPS && check_for_second_order_velocity_with_timing(tf, cs, lim, a_up);
    }

    // profile.rs:519-528
    void set_boundary_from_profile(const Profile& o) {
        a[0] = o.a[0]; v[0] = o.v[0]; p[0] = o.p[0];
        af = o.af; vf = o.vf; pf = o.pf;
        brake = o.brake; accel = o.accel;
    }

    // profile.rs:531-546
    void set_boundary(double p0_new, double v0_new, double a0_new, double pf_new, double vf_new, double af_new) {
        a[0] = a0_new; v[0] = v0_new; p[0] = p0_new; af = af_new; vf = vf_new; pf = pf_new;
    }

    // profile.rs:549-637 (|v7-vf| uses P_PRECISION, quirk Q8)
    bool check_for_second_order(ControlSigns control_signs_, ReachedLimits limits_, double a_up, double a_down, double v_max, double v_min) {
        if (t[0] < 0.0) return false;

        t_sum[0] = t[0];
        for (int i = 0; i < 6; ++i) {
            if (t[i + 1] < 0.0) return false;
            t_sum[i + 1] = t_sum[i] + t[i + 1];
        }

        if (t_sum[6] > T_MAX) return false;

        for (int i = 0; i < 7; ++i) j[i] = 0.0;

        if (control_signs_ == UDDU) {
            a[0] = (t[0] > 0.0 ? a_up : 0.0); a[1] = 0.0; a[2] = (t[2] > 0.0 ? a_down : 0.0); a[3] = 0.0;
            a[4] = (t[4] > 0.0 ? a_down : 0.0); a[5] = 0.0; a[6] = (t[6] > 0.0 ? a_up : 0.0); a[7] = af;
        } else {
            a[0] = (t[0] > 0.0 ? a_up : 0.0); a[1] = 0.0; a[2] = (t[2] > 0.0 ? a_down : 0.0); a[3] = 0.0;
            a[4] = (t[4] > 0.0 ? a_up : 0.0); a[5] = 0.0; a[6] = (t[6] > 0.0 ? a_down : 0.0); a[7] = af;
        }

        direction = (v_max > 0.0) ? UP : DOWN;

        const double v_upp_lim = (direction == UP) ? v_max + V_EPS : v_min + V_EPS;
        const double v_low_lim = (direction == UP) ? v_min - V_EPS : v_max - V_EPS;

        for (int i = 0; i < 7; ++i) {
            v[i + 1] = v[i] + t[i] * a[i];
            p[i + 1] = p[i] + t[i] * (v[i] + t[i] * a[i] / 2.0);
        }

        control_signs = control_signs_;
        limits = limits_;

        return std::fabs(p[7] - pf) < P_PRECISION && std::fabs(v[7] - vf) < P_PRECISION
            && v[2] <= v_upp_lim && v[2] >= v_low_lim && v[3] <= v_upp_lim && v[3] >= v_low_lim
            && v[4] <= v_upp_lim && v[4] >= v_low_lim && v[5] <= v_upp_lim && v[5] >= v_low_lim
            && v[6] <= v_upp_lim && v[6] >= v_low_lim && v[7] <= v_upp_lim && v[7] >= v_low_lim;
    }

    // profile.rs:640-653
    bool check_for_second_order_with_timing(ControlSigns cs, ReachedLimits lim, double /*tf*/, double a_up, double a_down, double v_max, double v_min) {
        return check_for_second_order(cs, lim, a_up, a_down, v_max, v_min);
    }

    // profile.rs:656-681
    bool check_for_second_order_with_timing_full(ControlSigns cs, ReachedLimits lim, double tf, double a_up, double a_down, double v_max, double v_min, double a_max, double a_min) {
        return (a_min - A_EPS < a_up) && (a_up < a_max + A_EPS) && (a_min - A_EPS < a_down) && (a_down < a_max + A_EPS)
            && check_for_second_order_with_timing(cs, lim, tf, a_up, a_down, v_max, v_min);
    }

    // profile.rs:685-726
    bool check_for_first_order(double v_up, ControlSigns control_signs_, ReachedLimits limits_) {
        if (t[3] < 0.0) return false;

        t_sum[0] = 0.0; t_sum[1] = 0.0; t_sum[2] = 0.0;
        t_sum[3] = t[3]; t_sum[4] = t[3]; t_sum[5] = t[3]; t_sum[6] = t[3];
        if (t_sum[6] > T_MAX) return false;

        for (int i = 0; i < 7; ++i) j[i] = 0.0;
        for (int i = 0; i < 7; ++i) a[i] = 0.0;
        a[7] = af;
        for (int i = 0; i < 7; ++i) v[i] = 0.0;
        v[3] = (t[3] > 0.0) ? v_up : 0.0;
        v[7] = vf;

        for (int i = 0; i < 7; ++i) {
            p[i + 1] = p[i] + t[i] * (v[i] + t[i] * a[i] / 2.0);
        }

        control_signs = control_signs_;
        limits = limits_;
        direction = (v_up > 0.0) ? UP : DOWN;

        return std::fabs(p[7] - pf) < P_PRECISION;
    }

    // profile.rs:729-737
    bool check_for_first_order_with_timing(ControlSigns cs, ReachedLimits lim, double /*tf*/, double v_up) {
        return check_for_first_order(v_up, cs, lim);
    }

    // profile.rs:740-752
    bool check_for_first_order_with_timing_full(ControlSigns cs, ReachedLimits lim, double tf, double v_up, double v_max, double v_min) {
        return (v_min - V_EPS < v_up) && (v_up < v_max + V_EPS) && check_for_first_order_with_timing(cs, lim, tf, v_up);
    }
};

// ------------------------------------------------------------------ roots.rs
namespace roots {

constexpr double COS_120 = -0.50;
constexpr double SIN_120 = 0.866025403784438600;  // roots.rs:12
constexpr double TOLERANCE = 1e-14;

static inline double pow2(double v) { return v * v; }

// roots.rs:20-122 — Set drops exact duplicates, PositiveSet keeps value >= 0,
// sorted_* gives the ascending order the `&mut set.into_iter()` loops see.
template <int N>
struct PositiveSet {
    double data[N];
    int size = 0;
    void insert(double value) {
        if (value >= 0.0) {
            for (int i = 0; i < size; ++i) if (data[i] == value) return;
            data[size++] = value;
        }
    }
    void sort() {  // stable insertion sort == slice::sort_by for these sizes
        for (int i = 1; i < size; ++i) {
            double x = data[i];
            int k = i - 1;
            while (k >= 0 && data[k] > x) { data[k + 1] = data[k]; --k; }
            data[k + 1] = x;
        }
    }
};

// roots.rs:132-231
static inline PositiveSet<3> solve_cub(double a, double b, double c, double d) {
    RKO_PHASE(RKO_PH_ROOTS);
    PositiveSet<3> roots;

    if (std::fabs(d) < EPS) {
        roots.insert(0.0);

        const double tmp = d;
        const double c2 = b;
        const double b2 = a;
        // (the reference shadows: d_=c, c=b, b=a)
        const double cc = c2, bb = b2;
        (void)tmp;
        // Converting to a quadratic equation: b*x^2 + c*x + tmp with b=a(old), c=b(old)
        if (std::fabs(bb) < EPS) {
            if (std::fabs(cc) > EPS) {
                roots.insert(-tmp / cc);
            }
        } else {
            const double discriminant = cc * cc - 4.0 * bb * tmp;
            if (discriminant >= 0.0) {
                const double inv2b = 1.0 / (2.0 * bb);
                const double y = std::sqrt(discriminant);
                roots.insert((-cc + y) * inv2b);
                roots.insert((-cc - y) * inv2b);
            }
        }
    } else if (std::fabs(a) < EPS) {
        if (std::fabs(b) < EPS) {
            if (std::fabs(c) > EPS) {
                roots.insert(-d / c);
            }
        } else {
            const double discriminant = c * c - 4.0 * b * d;
            if (discriminant >= 0.0) {
                const double inv2b = 1.0 / (2.0 * b);
                const double y = std::sqrt(discriminant);
                roots.insert((-c + y) * inv2b);
                roots.insert((-c - y) * inv2b);
            }
        }
    } else {
        const double inva = 1.0 / a;
        const double invaa = inva * inva;
        const double bb = b * b;
        const double bover3a = b * inva / 3.0;
        const double p = (a * c - bb / 3.0) * invaa;
        const double halfq = (2.0 * bb * b - 9.0 * a * b * c + 27.0 * a * a * d) / 54.0 * invaa * inva;
        const double yy = p * p * p / 27.0 + halfq * halfq;

        if (yy > EPS) {
            const double y = std::sqrt(yy);
            const double uuu = -halfq + y;
            const double vvv = -halfq - y;
            const double www = std::fabs(uuu) > std::fabs(vvv) ? uuu : vvv;
            const double w = m_cbrt(www);
            roots.insert(w - p / (3.0 * w) - bover3a);
        } else if (yy < -EPS) {
            const double x = -halfq;
            const double y = std::sqrt(-yy);
            double theta, r;

            if (std::fabs(x) > EPS) {
                theta = m_atan2(y, x);
                r = std::sqrt(x * x - yy);
            } else {
                theta = 3.14159265358979323846 / 2.0;
                r = y;
            }
            theta /= 3.0;
            r = 2.0 * m_cbrt(r);
            const double ux = m_cos(theta) * r;
            const double uyi = m_sin(theta) * r;

            roots.insert(ux - bover3a);
            roots.insert(ux * COS_120 - uyi * SIN_120 - bover3a);
            roots.insert(ux * COS_120 + uyi * SIN_120 - bover3a);
        } else {
            const double www = -halfq;
            const double w = 2.0 * m_cbrt(www);

            roots.insert(w - bover3a);
            roots.insert(w * COS_120 - bover3a);
        }
    }
    return roots;
}

// roots.rs:237-274
static inline int solve_resolvent(double x[3], double a, double b, double c) {
    a = a / 3.0;
    const double a2 = a * a;
    double q = a2 - b / 3.0;
    const double r = (a * (2.0 * a2 - b) + c) / 2.0;
    const double r2 = r * r;
    const double q3 = q * q * q;

    if (r2 < q3) {
        const double q_sqrt = std::sqrt(q);
        const double t = fmax_(fmin_(r / (q * q_sqrt), 1.0), -1.0);
        q = -2.0 * q_sqrt;

        const double theta = m_acos(t) / 3.0;
        const double ux = m_cos(theta) * q;
        const double uyi = m_sin(theta) * q;
        x[0] = ux - a;
        x[1] = ux * COS_120 - uyi * SIN_120 - a;
        x[2] = ux * COS_120 + uyi * SIN_120 - a;
        return 3;
    } else {
        double a_ = m_cbrt(-std::fabs(r) - std::sqrt(r2 - q3));
        if (r < 0.0) a_ = -a_;
        const double b_ = (a_ == 0.0) ? 0.0 : q / a_;

        x[0] = (a_ + b_) - a;
        x[1] = -(a_ + b_) / 2.0 - a;
        x[2] = std::sqrt(3.0) * (a_ - b_) / 2.0;
        if (std::fabs(x[2]) < EPS) {
            x[2] = x[1];
            return 2;
        }
        return 1;
    }
}

// roots.rs:278-370
static inline PositiveSet<4> solve_quart_monic(double a, double b, double c, double d) {
    RKO_PHASE(RKO_PH_ROOTS);
    PositiveSet<4> roots;

    const double a_squared = a * a;
    const double four_b = 4.0 * b;

    if (std::fabs(d) < EPS) {
        if (std::fabs(c) < EPS) {
            roots.insert(0.0);

            const double d_ = a_squared - four_b;
            if (std::fabs(d_) < EPS) {
                roots.insert(-a / 2.0);
            } else if (d_ > 0.0) {
                const double sqrt_d = std::sqrt(d_);
                roots.insert((-a - sqrt_d) / 2.0);
                roots.insert((-a + sqrt_d) / 2.0);
            }
            return roots;
        }

        if (std::fabs(a) < EPS && std::fabs(b) < EPS) {
            roots.insert(0.0);
            roots.insert(-m_cbrt(c));
            return roots;
        }
    }

    const double a3 = -b;
    const double b3 = a * c - 4.0 * d;
    const double c3 = -a_squared * d - c * c + four_b * d;

    double x3[3] = {0.0, 0.0, 0.0};
    const int number_zeroes = solve_resolvent(x3, a3, b3, c3);

    double y = x3[0];
    if (number_zeroes != 1) {
        if (std::fabs(x3[1]) > std::fabs(y)) y = x3[1];
        if (std::fabs(x3[2]) > std::fabs(y)) y = x3[2];
    }

    double q1, q2, p1, p2;

    double d_ = y * y - 4.0 * d;
    if (std::fabs(d_) < EPS) {
        q2 = y / 2.0;
        q1 = q2;
        d_ = a_squared - 4.0 * (b - y);
        if (std::fabs(d_) < EPS) {
            p2 = a / 2.0;
            p1 = p2;
        } else {
            const double sqrt_d = std::sqrt(d_);
            p1 = (a + sqrt_d) / 2.0;
            p2 = (a - sqrt_d) / 2.0;
        }
    } else {
        const double sqrt_d = std::sqrt(d_);
        q1 = (y + sqrt_d) / 2.0;
        q2 = (y - sqrt_d) / 2.0;
        p1 = (a * q1 - c) / (q1 - q2);
        p2 = (c - a * q2) / (q1 - q2);
    }

    constexpr double EPS_M_BY_16 = 16.0 * EPS;
    d_ = p1 * p1 - 4.0 * q1;
    if (std::fabs(d_) < EPS_M_BY_16) {
        roots.insert(-p1 / 2.0);
    } else if (d_ > 0.0) {
        const double sqrt_d = std::sqrt(d_);
        roots.insert((-p1 - sqrt_d) / 2.0);
        roots.insert((-p1 + sqrt_d) / 2.0);
    }

    d_ = p2 * p2 - 4.0 * q2;
    if (std::fabs(d_) < EPS_M_BY_16) {
        roots.insert(-p2 / 2.0);
    } else if (d_ > 0.0) {
        const double sqrt_d = std::sqrt(d_);
        roots.insert((-p2 - sqrt_d) / 2.0);
        roots.insert((-p2 + sqrt_d) / 2.0);
    }

    return roots;
}

// Small polynomial container standing in for ArrayVec<f64, N> (roots.rs:377-415)
struct Poly {
    double c[7];
    int len = 0;
    void push(double x) { c[len++] = x; }
    double operator[](int i) const { return c[i]; }
};

// roots.rs:379-386
static inline Poly poly_deri(const Poly& coeffs) {
    Poly deriv;
    const int len = coeffs.len;
    for (int i = 0; i < len - 1; ++i) deriv.push((double)(len - 1 - i) * coeffs[i]);
    return deriv;
}

// roots.rs:389-397
static inline Poly poly_monic_deri(const Poly& monic_coeffs) {
    Poly deriv;
    const int len = monic_coeffs.len;
    deriv.push(1.0);
    for (int i = 1; i < len - 1; ++i) deriv.push((double)(len - 1 - i) * monic_coeffs[i] / (double)(len - 1));
    return deriv;
}

// roots.rs:400-415 (sums from the constant term upward)
static inline double poly_eval(const Poly& p, double x) {
    double result = 0.0;
    const int n = p.len;
    if (std::fabs(x) < EPS) {
        result = p[n - 1];
    } else if (std::fabs(x - 1.0) < EPS) {
        for (int i = 0; i < n; ++i) result += p[i];
    } else {
        double xn = 1.0;
        for (int i = n - 1; i >= 0; --i) {
            result += p[i] * xn;
            xn *= x;
        }
    }
    return result;
}

// roots.rs:424-481 (MAX_ITS = 128, :420)
static inline double shrink_interval(const Poly& p, double l, double h) {
    RKO_PHASE(RKO_PH_NEWTON);
    const Poly deriv = poly_deri(p);
    const double fl = poly_eval(p, l);
    const double fh = poly_eval(p, h);
    if (fl == 0.0) return l;
    if (fh == 0.0) return h;
    if (fl > 0.0) std::swap(l, h);

    double rts = (l + h) / 2.0;
    double dxold = std::fabs(h - l);
    double dx = dxold;
    double f = poly_eval(p, rts);
    double df = poly_eval(deriv, rts);
    for (int it = 0; it < 128; ++it) {
        if ((((rts - h) * df - f) * ((rts - l) * df - f) > 0.0) || std::fabs(2.0 * f) > std::fabs(dxold) * std::fabs(df)) {
            dxold = dx;
            dx = (h - l) / 2.0;
            rts = l + dx;
            if (std::fabs(l - rts) < EPS) break;
        } else {
            dxold = dx;
            dx = f / df;
            const double temp = rts;
            rts -= dx;
            if (std::fabs(temp - rts) < EPS) break;
        }

        if (std::fabs(dx) < TOLERANCE) break;

        f = poly_eval(p, rts);
        df = poly_eval(deriv, rts);
        if (f < 0.0) l = rts; else h = rts;
    }
    return rts;
}

}  // namespace roots

// ------------------------------------------------------------------ profile.rs:754-895 (trajectory query helpers)
struct PositionExtrema { double min, max, t_min, t_max; };

// profile.rs:754-776
static inline void check_position_extremum(double t_ext, double t_sum, double t, double p, double v, double a, double j, PositionExtrema& ext) {
    if (0.0 < t_ext && t_ext < t) {
        double p_ext, v_ext, a_ext;
        integrate(t_ext, p, v, a, j, p_ext, v_ext, a_ext);
        if (a_ext > 0.0 && p_ext < ext.min) {
            ext.min = p_ext;
            ext.t_min = t_sum + t_ext;
        } else if (a_ext < 0.0 && p_ext > ext.max) {
            ext.max = p_ext;
            ext.t_max = t_sum + t_ext;
        }
    }
}

// profile.rs:778-806
static inline void check_step_for_position_extremum(double t_sum, double t, double p, double v, double a, double j, PositionExtrema& ext) {
    if (p < ext.min) { ext.min = p; ext.t_min = t_sum; }
    if (p > ext.max) { ext.max = p; ext.t_max = t_sum; }
    if (j != 0.0) {
        const double d = a * a - 2.0 * j * v;
        if (std::fabs(d) < EPS) {
            check_position_extremum(-a / j, t_sum, t, p, v, a, j, ext);
        } else if (d > 0.0) {
            const double d_sqrt = std::sqrt(d);
            check_position_extremum((-a - d_sqrt) / j, t_sum, t, p, v, a, j, ext);
            check_position_extremum((-a + d_sqrt) / j, t_sum, t, p, v, a, j, ext);
        }
    }
}

// profile.rs:808-863
static inline PositionExtrema get_position_extrema(const Profile& q) {
    PositionExtrema ext{INF, -INF, 0.0, 0.0};
    if (q.brake.duration > 0.0 && q.brake.t[0] > 0.0) {
        check_step_for_position_extremum(0.0, q.brake.t[0], q.brake.p[0], q.brake.v[0], q.brake.a[0], q.brake.j[0], ext);
        if (q.brake.t[1] > 0.0)
            check_step_for_position_extremum(q.brake.t[0], q.brake.t[1], q.brake.p[1], q.brake.v[1], q.brake.a[1], q.brake.j[1], ext);
    }
    double t_current_sum = 0.0;
    for (int i = 0; i < 7; ++i) {
        if (i > 0) t_current_sum = q.t_sum[i - 1];
        check_step_for_position_extremum(t_current_sum + q.brake.duration, q.t[i], q.p[i], q.v[i], q.a[i], q.j[i], ext);
    }
    if (q.pf < ext.min) { ext.min = q.pf; ext.t_min = q.t_sum[6] + q.brake.duration; }
    if (q.pf > ext.max) { ext.max = q.pf; ext.t_max = q.t_sum[6] + q.brake.duration; }
    return ext;
}

// profile.rs:865-895. The roots of the cubic are visited in INSERTION order (get_data(), Q9), not ascending.
static inline bool get_first_state_at_position(const Profile& q, double pt, double offset, double& time, double& vt, double& at) {
    for (int i = 0; i < 7; ++i) {
        if (std::fabs(q.p[i] - pt) < EPS) {
            time = offset + (i > 0 ? q.t_sum[i - 1] : 0.0);
            vt = q.v[i]; at = q.a[i];
            return true;
        }
        if (q.t[i] == 0.0) continue;
        const roots::PositiveSet<3> r = roots::solve_cub(q.j[i] / 6.0, q.a[i] / 2.0, q.v[i], q.p[i] - pt);
        for (int k = 0; k < r.size; ++k) {
            const double t_ = r.data[k];
            if (0.0 < t_ && t_ <= q.t[i]) {
                time = offset + t_ + (i > 0 ? q.t_sum[i - 1] : 0.0);
                double pp;
                integrate(t_, q.p[i], q.v[i], q.a[i], q.j[i], pp, vt, at);
                return true;
            }
        }
    }
    if (std::fabs(q.pf - pt) < 1e-9) {
        time = offset + q.t_sum[6];
        vt = q.vf; at = q.af;
        return true;
    }
    return false;
}

// ------------------------------------------------------------------ block.rs
struct Interval {
    double left = 0.0, right = 0.0;
    Profile profile;

    // block.rs:220-239
    static Interval from_profiles(const Profile& profile_left, const Profile& profile_right) {
        const double left_duration = profile_left.t_sum[6] + profile_left.brake.duration + profile_left.accel.duration;
        const double right_duration = profile_right.t_sum[6] + profile_right.brake.duration + profile_right.accel.duration;
        Interval iv;
        if (left_duration < right_duration) {
            iv.left = left_duration; iv.right = right_duration; iv.profile = profile_right;
        } else {
            iv.left = right_duration; iv.right = left_duration; iv.profile = profile_left;
        }
        return iv;
    }
};

struct Block {
    Profile p_min;
    double t_min = 0.0;
    bool has_a = false, has_b = false;
    Interval a, b;

    // block.rs:30-37
    void set_min_profile(const Profile& profile) {
        p_min = profile;
        t_min = p_min.t_sum[6] + p_min.brake.duration + p_min.accel.duration;
        has_a = false;
        has_b = false;
    }

    // block.rs:17-26
    static void remove_profile(Profile* valid_profiles, int& valid_profile_counter, int index) {
        for (int i = index; i < valid_profile_counter - 1; ++i) valid_profiles[i] = valid_profiles[i + 1];
        valid_profile_counter -= 1;
    }

    // block.rs:40-154 (numerical_robust is always None -> true)
    static bool calculate_block(Block& block, Profile* valid_profiles, int& valid_profile_counter) {
        if (valid_profile_counter == 1) {
            block.set_min_profile(valid_profiles[0]);
            return true;
        } else if (valid_profile_counter == 2) {
            if (std::fabs(valid_profiles[0].t_sum[6] - valid_profiles[1].t_sum[6]) < 8.0 * EPS) {
                block.set_min_profile(valid_profiles[0]);
                return true;
            }

            const int idx_min = (valid_profiles[0].t_sum[6] < valid_profiles[1].t_sum[6]) ? 0 : 1;
            const int idx_else_1 = (idx_min + 1) % 2;

            block.set_min_profile(valid_profiles[idx_min]);
            block.a = Interval::from_profiles(valid_profiles[idx_min], valid_profiles[idx_else_1]);
            block.has_a = true;
            return true;
        } else if (valid_profile_counter == 4) {
            if (std::fabs(valid_profiles[0].t_sum[6] - valid_profiles[1].t_sum[6]) < 32.0 * EPS
                && valid_profiles[0].direction != valid_profiles[1].direction) {
                remove_profile(valid_profiles, valid_profile_counter, 1);
            } else if ((std::fabs(valid_profiles[2].t_sum[6] - valid_profiles[3].t_sum[6]) < 256.0 * EPS
                        && valid_profiles[2].direction != valid_profiles[3].direction)
                       || (std::fabs(valid_profiles[0].t_sum[6] - valid_profiles[3].t_sum[6]) < 256.0 * EPS
                           && valid_profiles[0].direction != valid_profiles[3].direction)) {
                remove_profile(valid_profiles, valid_profile_counter, 3);
            } else {
                return false;
            }
        } else if (valid_profile_counter % 2 == 0) {
            return false;
        }

        // first minimum of t_sum[6] (Iterator::min_by returns the first of equal minima)
        int idx_min = 0;
        for (int i = 1; i < valid_profile_counter; ++i) {
            if (valid_profiles[i].t_sum[6] < valid_profiles[idx_min].t_sum[6]) idx_min = i;
        }

        block.set_min_profile(valid_profiles[idx_min]);

        if (valid_profile_counter == 3) {
            const int idx_else_1 = (idx_min + 1) % 3;
            const int idx_else_2 = (idx_min + 2) % 3;

            block.a = Interval::from_profiles(valid_profiles[idx_else_1], valid_profiles[idx_else_2]);
            block.has_a = true;
            return true;
        } else if (valid_profile_counter == 5) {
            const int idx_else_1 = (idx_min + 1) % 5;
            const int idx_else_2 = (idx_min + 2) % 5;
            const int idx_else_3 = (idx_min + 3) % 5;
            const int idx_else_4 = (idx_min + 4) % 5;

            if (valid_profiles[idx_else_1].direction == valid_profiles[idx_else_2].direction) {
                block.a = Interval::from_profiles(valid_profiles[idx_else_1], valid_profiles[idx_else_2]);
                block.b = Interval::from_profiles(valid_profiles[idx_else_3], valid_profiles[idx_else_4]);
            } else {
                block.a = Interval::from_profiles(valid_profiles[idx_else_1], valid_profiles[idx_else_4]);
                block.b = Interval::from_profiles(valid_profiles[idx_else_2], valid_profiles[idx_else_3]);
            }
            block.has_a = true;
            block.has_b = true;
            return true;
        }

        return false;
    }

    // block.rs:157-172
    bool is_blocked(double t) const {
        if (t < t_min) return true;
        if (has_a && t > a.left && t < a.right) return true;
        if (has_b && t > b.left && t < b.right) return true;
        return false;
    }
};

}  // namespace rko
