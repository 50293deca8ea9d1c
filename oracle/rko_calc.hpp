// rko_calc.hpp — parity oracle: InputParameter / validation, TargetCalculator,
// Trajectory::at_time and the Ruckig façade (calculate / update).
//
// TEST INFRASTRUCTURE ONLY (see rko_math.hpp). Restates rsruckig v2.1.2:
//   input_parameter.rs:147-420, calculator_target.rs, trajectory.rs:52-189,
//   ruckig.rs:110-321, output_parameter.rs:110-120, result.rs:34-61.
// Objects keep state between calls exactly like the reference (blocks, profiles
// and output.time are never reset), so the reference tests can be replayed.
#pragma once
#include <string>
#include <vector>

#include "rko_core.hpp"
#include "rko_steps.hpp"
#include "rko_step2_pos3.hpp"

namespace rko {

// input_parameter.rs:35-45, :72-85, :110-117 (discriminant order kept)
enum ControlInterface : int { CI_POSITION = 0, CI_VELOCITY = 1, CI_ACCELERATION = 2 };
enum Synchronization : int { SYNC_TIME = 0, SYNC_TIME_IF_NECESSARY = 1, SYNC_PHASE = 2, SYNC_NONE = 3 };
enum DurationDiscretization : int { DD_CONTINUOUS = 0, DD_DISCRETE = 1 };

// result.rs:34-61
enum RuckigResult : int {
    R_WORKING = 0, R_FINISHED = 1, R_ERROR = -1, R_ERROR_INVALID_INPUT = -100, R_ERROR_TRAJECTORY_DURATION = -101,
    R_ERROR_POSITIONAL_LIMITS = -102, R_ERROR_ZERO_LIMITS = -104, R_ERROR_EXECUTION_TIME_CALCULATION = -110,
    R_ERROR_SYNCHRONIZATION_CALCULATION = -111
};

// Outcome of a fallible reference call: Ok(code) or Err(kind, message).
enum ErrKind : int { ERR_NONE = 0, ERR_VALIDATION = 1, ERR_CALCULATOR = 2 };
struct Outcome {
    int code = R_WORKING;
    int err = ERR_NONE;
    std::string msg;
    bool is_err() const { return err != ERR_NONE; }
};

// input_parameter.rs:147-243
struct InputParameter {
    int degrees_of_freedom = 0;
    int control_interface = CI_POSITION;
    int synchronization = SYNC_TIME;
    int duration_discretization = DD_CONTINUOUS;
    std::vector<double> current_position, current_velocity, current_acceleration;
    std::vector<double> target_position, target_velocity, target_acceleration;
    std::vector<double> max_velocity, max_acceleration, max_jerk;
    bool has_min_velocity = false, has_min_acceleration = false;
    std::vector<double> min_velocity, min_acceleration;
    std::vector<uint8_t> enabled;
    bool has_per_dof_control_interface = false, has_per_dof_synchronization = false;
    std::vector<int> per_dof_control_interface, per_dof_synchronization;
    bool has_minimum_duration = false;
    double minimum_duration = 0.0;

    InputParameter() {}
    explicit InputParameter(int dofs) { resize(dofs); }

    // input_parameter.rs:220-243 defaults
    void resize(int dofs) {
        degrees_of_freedom = dofs;
        current_position.assign(dofs, 0.0); current_velocity.assign(dofs, 0.0); current_acceleration.assign(dofs, 0.0);
        target_position.assign(dofs, 0.0); target_velocity.assign(dofs, 0.0); target_acceleration.assign(dofs, 0.0);
        max_velocity.assign(dofs, 0.0); max_acceleration.assign(dofs, INF); max_jerk.assign(dofs, INF);
        min_velocity.assign(dofs, 0.0); min_acceleration.assign(dofs, 0.0);
        enabled.assign(dofs, 1);
        per_dof_control_interface.assign(dofs, CI_POSITION); per_dof_synchronization.assign(dofs, SYNC_TIME);
    }

    // input_parameter.rs:190-211
    bool equals(const InputParameter& o) const {
        auto opt_eq = [](bool ha, const std::vector<double>& a, bool hb, const std::vector<double>& b) {
            if (ha != hb) return false;
            return !ha || a == b;
        };
        auto opt_eq_i = [](bool ha, const std::vector<int>& a, bool hb, const std::vector<int>& b) {
            if (ha != hb) return false;
            return !ha || a == b;
        };
        return current_position == o.current_position && current_velocity == o.current_velocity
            && current_acceleration == o.current_acceleration && target_position == o.target_position
            && target_velocity == o.target_velocity && target_acceleration == o.target_acceleration
            && max_velocity == o.max_velocity && max_acceleration == o.max_acceleration && max_jerk == o.max_jerk
            && enabled == o.enabled
            && has_minimum_duration == o.has_minimum_duration && (!has_minimum_duration || minimum_duration == o.minimum_duration)
            && opt_eq(has_min_velocity, min_velocity, o.has_min_velocity, o.min_velocity)
            && opt_eq(has_min_acceleration, min_acceleration, o.has_min_acceleration, o.min_acceleration)
            && control_interface == o.control_interface && synchronization == o.synchronization
            && duration_discretization == o.duration_discretization
            && opt_eq_i(has_per_dof_control_interface, per_dof_control_interface, o.has_per_dof_control_interface, o.per_dof_control_interface)
            && opt_eq_i(has_per_dof_synchronization, per_dof_synchronization, o.has_per_dof_synchronization, o.per_dof_synchronization);
    }

    static double v_at_a_zero(double v0, double a0, double j) { return v0 + (a0 * a0) / (2.0 * j); }  // :245-248

    // input_parameter.rs:251-420. Returns "" when valid, otherwise the (abridged)
    // message of the FIRST failing check in DoF order; `which` gets a stable id.
    std::string validate(bool check_current, bool check_target, int* which_dof = nullptr) const {
        auto fail = [&](int dof, const char* m) { if (which_dof) *which_dof = dof; return std::string(m) + " (DoF " + std::to_string(dof) + ")"; };
        for (int dof = 0; dof < degrees_of_freedom; ++dof) {
            const double j_max = max_jerk[dof];
            if (j_max != j_max || j_max < 0.0) return fail(dof, "Maximum jerk limit should be larger than or equal to zero.");

            const double a_max = max_acceleration[dof];
            if (a_max != a_max || a_max < 0.0) return fail(dof, "maximum acceleration limit should be larger than or equal to zero.");

            const double a_min = has_min_acceleration ? min_acceleration[dof] : -max_acceleration[dof];
            if (a_min != a_min || a_min > 0.0) return fail(dof, "minimum acceleration limit should be smaller than or equal to zero.");

            const double a0 = current_acceleration[dof];
            if (a0 != a0) return fail(dof, "current acceleration should be a valid number.");
            const double af = target_acceleration[dof];
            if (af != af) return fail(dof, "target acceleration should be a valid number.");

            if (check_current) {
                if (a0 > a_max) return fail(dof, "current acceleration exceeds its maximum acceleration limit.");
                if (a0 < a_min) return fail(dof, "current acceleration undercuts its minimum acceleration limit.");
            }
            if (check_target) {
                if (af > a_max) return fail(dof, "target acceleration exceeds its maximum acceleration limit.");
                if (af < a_min) return fail(dof, "target acceleration undercuts its minimum acceleration limit.");
            }

            const double v0 = current_velocity[dof];
            if (v0 != v0) return fail(dof, "current velocity should be a valid number.");
            const double vf = target_velocity[dof];
            if (vf != vf) return fail(dof, "target velocity should be a valid number.");

            const int ci = has_per_dof_control_interface ? per_dof_control_interface[dof] : control_interface;
            if (ci == CI_POSITION) {
                const double p0 = current_position[dof];
                if (p0 != p0) return fail(dof, "current position should be a valid number.");
                const double pf = target_position[dof];
                if (pf != pf) return fail(dof, "target position should be a valid number.");

                const double v_max = max_velocity[dof];
                if (v_max != v_max || v_max < 0.0) return fail(dof, "maximum velocity limit should be larger than or equal to zero.");
                const double v_min = has_min_velocity ? min_velocity[dof] : -v_max;
                if (v_min != v_min || v_min > 0.0) return fail(dof, "minimum velocity limit should be smaller than or equal to zero.");

                if (check_current) {
                    if (v0 > v_max) return fail(dof, "current velocity exceeds its maximum velocity limit.");
                    if (v0 < v_min) return fail(dof, "current velocity undercuts its minimum velocity limit.");
                }
                if (check_target) {
                    if (vf > v_max) return fail(dof, "target velocity exceeds its maximum velocity limit.");
                    if (vf < v_min) return fail(dof, "target velocity undercuts its minimum velocity limit.");
                }
                if (check_current) {
                    if (a0 > 0.0 && j_max > 0.0 && v_at_a_zero(v0, a0, j_max) > v_max)
                        return fail(dof, "will inevitably reach a velocity from the current kinematic state that will exceed its maximum velocity limit.");
                    if (a0 < 0.0 && j_max > 0.0 && v_at_a_zero(v0, a0, -j_max) < v_min)
                        return fail(dof, "will inevitably reach a velocity from the current kinematic state that will undercut its minimum velocity limit.");
                }
                if (check_target) {
                    if (af < 0.0 && j_max > 0.0 && v_at_a_zero(vf, af, j_max) > v_max)
                        return fail(dof, "will inevitably have reached a velocity from the target kinematic state that will exceed its maximum velocity limit.");
                    if (af > 0.0 && j_max > 0.0 && v_at_a_zero(vf, af, -j_max) < v_min)
                        return fail(dof, "will inevitably have reached a velocity from the target kinematic state that will undercut its minimum velocity limit.");
                }
            }
        }
        return std::string();
    }
};

// trajectory.rs:14-50
struct Trajectory {
    std::vector<Profile> profiles;  // one section
    double duration = 0.0;
    std::vector<double> cumulative_times, independent_min_durations;
    int degrees_of_freedom = 0;

    Trajectory() {}
    explicit Trajectory(int dofs) { resize(dofs); }
    void resize(int dofs) {
        degrees_of_freedom = dofs;
        profiles.assign(dofs, Profile());
        duration = 0.0;
        cumulative_times.assign(dofs, 0.0);
        independent_min_durations.assign(dofs, 0.0);
    }

    // trajectory.rs:52-189 (one section: profiles.len() == 1)
    void at_time(double time, double* new_position, double* new_velocity, double* new_acceleration, double* new_jerk, int* new_section) const {
        const int dofs = (int)profiles.size();
        auto set_integrate = [&](int dof, double t, double p, double v, double a, double j) {
            double pos, vel, acc;
            integrate(t, p, v, a, j, pos, vel, acc);
            if (new_position) new_position[dof] = pos;
            if (new_velocity) new_velocity[dof] = vel;
            if (new_acceleration) new_acceleration[dof] = acc;
            if (new_jerk) new_jerk[dof] = j;
        };

        if (time >= duration) {
            if (new_section) *new_section = 1;
            for (int dof = 0; dof < dofs; ++dof) {
                const Profile& p = profiles[dof];
                const double t_pre = p.brake.duration;
                const double t_diff = time - (t_pre + p.t_sum[6]);
                set_integrate(dof, t_diff, p.p[7], p.v[7], p.a[7], 0.0);
            }
            return;
        }

        // cumulative_times has DOF entries of which only [0] is written (quirk Q13)
        int new_section_index = (int)cumulative_times.size();
        for (int i = 0; i < (int)cumulative_times.size(); ++i) {
            if (cumulative_times[i] > time) { new_section_index = i; break; }
        }
        if (new_section) *new_section = new_section_index;
        double t_diff = time;
        if (new_section_index > 0) t_diff -= cumulative_times[new_section_index - 1];
        // Only section 0 exists; the reference would index out of bounds otherwise.
        for (int dof = 0; dof < dofs; ++dof) {
            const Profile& p = profiles[dof];
            double t_diff_dof = t_diff;

            if (new_section_index == 0 && p.brake.duration > 0.0) {
                if (t_diff_dof < p.brake.duration) {
                    const int index = (t_diff_dof < p.brake.t[0]) ? 0 : 1;
                    if (index > 0) t_diff_dof -= p.brake.t[index - 1];
                    set_integrate(dof, t_diff_dof, p.brake.p[index], p.brake.v[index], p.brake.a[index], p.brake.j[index]);
                    continue;
                } else {
                    t_diff_dof -= p.brake.duration;
                }
            }
            if (t_diff_dof >= p.t_sum[6]) {
                set_integrate(dof, t_diff_dof - p.t_sum[6], p.p[7], p.v[7], p.a[7], 0.0);
                continue;
            }

            int index_dof = 6;
            for (int i = 0; i < 7; ++i) {
                if (p.t_sum[i] > t_diff_dof) { index_dof = i; break; }
            }
            if (index_dof > 0) t_diff_dof -= p.t_sum[index_dof - 1];

            set_integrate(dof, t_diff_dof, p.p[index_dof], p.v[index_dof], p.a[index_dof], p.j[index_dof]);
        }
    }
};

// calculator_target.rs:25-56
struct TargetCalculator {
    double eps = EPS;
    bool return_error_at_maximal_duration = true;
    std::vector<double> new_phase_control, pd, possible_t_syncs;
    std::vector<size_t> idx;
    std::vector<Block> blocks;
    std::vector<double> inp_min_velocity, inp_min_acceleration;
    std::vector<int> inp_per_dof_control_interface, inp_per_dof_synchronization;
    int degrees_of_freedom = 0;
    // Instrumentation (not in the reference): limiting DoF of the last call, -1 = none.
    int last_limiting_dof = -1;
    // Instrumentation: which path produced each DoF's final profile
    // (0 = step1 p_min, 1 = a.profile, 2 = b.profile, 3 = phase, 4 = step2, 5 = disabled/untouched)
    std::vector<int> profile_source;

    explicit TargetCalculator(int dofs) : degrees_of_freedom(dofs) {
        blocks.assign(dofs, Block());
        inp_min_velocity.assign(dofs, 0.0); inp_min_acceleration.assign(dofs, 0.0);
        inp_per_dof_control_interface.assign(dofs, CI_POSITION); inp_per_dof_synchronization.assign(dofs, SYNC_TIME);
        new_phase_control.assign(dofs, 0.0); pd.assign(dofs, 0.0);
        possible_t_syncs.assign(3 * dofs + 1, 0.0);
        idx.assign(3 * dofs + 1, 0);
        profile_source.assign(dofs, 5);
    }

    // calculator_target.rs:60-148
    bool is_input_collinear(const InputParameter& inp, Direction limiting_direction, int limiting_dof) {
        for (int dof = 0; dof < degrees_of_freedom; ++dof) pd[dof] = inp.target_position[dof] - inp.current_position[dof];

        const std::vector<double>* scale_vector = nullptr;
        int scale_dof = -1;
        for (int dof = 0; dof < degrees_of_freedom; ++dof) {
            if (inp_per_dof_synchronization[dof] != SYNC_PHASE) continue;

            if (inp_per_dof_control_interface[dof] == CI_POSITION && std::fabs(pd[dof]) > eps) { scale_vector = &pd; scale_dof = dof; break; }
            else if (std::fabs(inp.current_velocity[dof]) > eps) { scale_vector = &inp.current_velocity; scale_dof = dof; break; }
            else if (std::fabs(inp.current_acceleration[dof]) > eps) { scale_vector = &inp.current_acceleration; scale_dof = dof; break; }
            else if (std::fabs(inp.target_velocity[dof]) > eps) { scale_vector = &inp.target_velocity; scale_dof = dof; break; }
            else if (std::fabs(inp.target_acceleration[dof]) > eps) { scale_vector = &inp.target_acceleration; scale_dof = dof; break; }
        }

        if (scale_dof < 0) return false;

        const double scale = (*scale_vector)[scale_dof];
        const double pd_scale = pd[scale_dof] / scale;
        const double v0_scale = inp.current_velocity[scale_dof] / scale;
        const double vf_scale = inp.target_velocity[scale_dof] / scale;
        const double a0_scale = inp.current_acceleration[scale_dof] / scale;
        const double af_scale = inp.target_acceleration[scale_dof] / scale;

        const double scale_limiting = (*scale_vector)[limiting_dof];
        double control_limiting = (limiting_direction == UP) ? inp.max_jerk[limiting_dof] : -inp.max_jerk[limiting_dof];
        if (std::isinf(inp.max_jerk[limiting_dof])) {
            control_limiting = (limiting_direction == UP) ? inp.max_acceleration[limiting_dof] : inp_min_acceleration[limiting_dof];
        }

        for (int dof = 0; dof < degrees_of_freedom; ++dof) {
            if (inp_per_dof_synchronization[dof] != SYNC_PHASE) continue;

            const double current_scale = (*scale_vector)[dof];
            if ((inp_per_dof_control_interface[dof] == CI_POSITION && std::fabs(pd[dof] - pd_scale * current_scale) > eps)
                || std::fabs(inp.current_velocity[dof] - v0_scale * current_scale) > eps
                || std::fabs(inp.current_acceleration[dof] - a0_scale * current_scale) > eps
                || std::fabs(inp.target_velocity[dof] - vf_scale * current_scale) > eps
                || std::fabs(inp.target_acceleration[dof] - af_scale * current_scale) > eps) {
                return false;
            }

            new_phase_control[dof] = control_limiting * current_scale / scale_limiting;
        }
        return true;
    }

    // calculator_target.rs:150-275
    bool synchronize(bool has_t_min, double t_min, double& t_sync, int& limiting_dof, std::vector<Profile>& profiles, bool discrete_duration, double delta_time) {
        const int D = degrees_of_freedom;
        bool any_interval = false;
        for (int dof = 0; dof < D; ++dof) {
            if (inp_per_dof_synchronization[dof] == SYNC_NONE) {
                possible_t_syncs[dof] = 0.0;
                possible_t_syncs[D + dof] = INF;
                possible_t_syncs[2 * D + dof] = INF;
                continue;
            }

            possible_t_syncs[dof] = blocks[dof].t_min;
            possible_t_syncs[D + dof] = blocks[dof].has_a ? blocks[dof].a.right : INF;
            possible_t_syncs[2 * D + dof] = blocks[dof].has_b ? blocks[dof].b.right : INF;
            any_interval |= blocks[dof].has_a || blocks[dof].has_b;
        }
        possible_t_syncs[3 * D] = has_t_min ? t_min : INF;
        any_interval |= has_t_min;

        if (discrete_duration) {
            for (double& possible_t_sync : possible_t_syncs) {
                if (std::isinf(possible_t_sync)) continue;

                const double remainder = std::fmod(possible_t_sync, delta_time);
                if (remainder > eps) possible_t_sync += delta_time - remainder;
            }
        }

        const size_t idx_end = any_interval ? idx.size() : (size_t)D;
        for (size_t i = 0; i < idx_end; ++i) idx[i] = i;

        // slice::sort_by is stable; NaN panics in the reference (partial_cmp().unwrap())
        std::stable_sort(idx.begin(), idx.begin() + idx_end, [&](size_t i, size_t j) { return possible_t_syncs[i] < possible_t_syncs[j]; });

        // Start at last tmin (or worse); runs to the end of idx, stale tail included (quirk Q4)
        for (size_t k = (size_t)(D - 1); k < idx.size(); ++k) {
            const size_t i = idx[k];
            const double possible_t_sync = possible_t_syncs[i];
            bool is_blocked = false;
            for (int dof = 0; dof < D; ++dof) {
                if (inp_per_dof_synchronization[dof] == SYNC_NONE) continue;
                if (blocks[dof].is_blocked(possible_t_sync)) { is_blocked = true; break; }
            }
            if (is_blocked || possible_t_sync < (has_t_min ? t_min : 0.0) || std::isinf(possible_t_sync)) continue;

            t_sync = possible_t_sync;
            if (i == (size_t)(3 * D)) {
                limiting_dof = -1;
                return true;
            }

            const size_t div = i / D;
            limiting_dof = (int)(i % D);
            switch (div) {
                case 0: profiles[limiting_dof] = blocks[limiting_dof].p_min; profile_source[limiting_dof] = 0; break;
                case 1: profiles[limiting_dof] = blocks[limiting_dof].a.profile; profile_source[limiting_dof] = 1; break;
                case 2: profiles[limiting_dof] = blocks[limiting_dof].b.profile; profile_source[limiting_dof] = 2; break;
                default: break;
            }
            return true;
        }
        return false;
    }

    // Handler semantics (error.rs:98-122): ThrowErrorHandler turns a handler call
    // into Err(..); IgnoreErrorHandler returns Ok(()) and the caller carries on.
    static bool handle(bool throw_mode, int kind, const std::string& m, Outcome& out) {
        if (throw_mode) { out.err = kind; out.msg = m; return true; }
        return false;
    }

    // calculator_target.rs:278-833
    Outcome calculate(const InputParameter& inp, Trajectory& traj, double delta_time, bool throw_mode) {
        Outcome out;
        const int D = degrees_of_freedom;
        last_limiting_dof = -1;
        for (int i = 0; i < D; ++i) {
            inp_per_dof_control_interface[i] = inp.control_interface;
            inp_per_dof_synchronization[i] = inp.synchronization;
            profile_source[i] = 5;
        }
        if (inp.has_per_dof_control_interface) for (int i = 0; i < D; ++i) inp_per_dof_control_interface[i] = inp.per_dof_control_interface[i];
        if (inp.has_per_dof_synchronization) for (int i = 0; i < D; ++i) inp_per_dof_synchronization[i] = inp.per_dof_synchronization[i];

        for (int dof = 0; dof < D; ++dof) {
            Profile& p = traj.profiles[dof];

            inp_min_velocity[dof] = inp.has_min_velocity ? inp.min_velocity[dof] : -inp.max_velocity[dof];
            inp_min_acceleration[dof] = inp.has_min_acceleration ? inp.min_acceleration[dof] : -inp.max_acceleration[dof];

            if (!inp.enabled[dof]) {
                p.p[7] = inp.current_position[dof];
                p.v[7] = inp.current_velocity[dof];
                p.a[7] = inp.current_acceleration[dof];
                p.t_sum[6] = 0.0;

                blocks[dof].t_min = 0.0;
                blocks[dof].has_a = false;
                blocks[dof].has_b = false;
                continue;
            }

            const double min_vel = inp_min_velocity[dof];
            const double min_acc = inp_min_acceleration[dof];
            const bool jerk_inf = std::isinf(inp.max_jerk[dof]);
            const bool acc_inf = std::isinf(inp.max_acceleration[dof]);

            switch (inp_per_dof_control_interface[dof]) {
                case CI_POSITION:
                    if (!jerk_inf) {
                        p.brake.get_position_brake_trajectory(inp.current_velocity[dof], inp.current_acceleration[dof], inp.max_velocity[dof], min_vel, inp.max_acceleration[dof], min_acc, inp.max_jerk[dof]);
                    } else if (!acc_inf) {
                        p.brake.get_second_order_position_brake_trajectory(inp.current_velocity[dof], inp.max_velocity[dof], min_vel, inp.max_acceleration[dof], min_acc);
                    }
                    p.set_boundary(inp.current_position[dof], inp.current_velocity[dof], inp.current_acceleration[dof], inp.target_position[dof], inp.target_velocity[dof], inp.target_acceleration[dof]);
                    break;
                case CI_VELOCITY:
                    if (!jerk_inf) {
                        p.brake.get_velocity_brake_trajectory(inp.current_acceleration[dof], inp.max_acceleration[dof], min_acc, inp.max_jerk[dof]);
                    } else {
                        p.brake.get_second_order_velocity_brake_trajectory();
                    }
                    p.set_boundary_for_velocity(inp.current_position[dof], inp.current_velocity[dof], inp.current_acceleration[dof], inp.target_velocity[dof], inp.target_acceleration[dof]);
                    break;
                default: break;
            }

            if (!jerk_inf) {
                p.brake.finalize(p.p[0], p.v[0], p.a[0]);
            } else if (!acc_inf) {
                p.brake.finalize_second_order(p.p[0], p.v[0], p.a[0]);
            }

            bool found_profile = false;
            switch (inp_per_dof_control_interface[dof]) {
                case CI_POSITION:
                    if (!jerk_inf) {
                        PositionThirdOrderStep1 step1(p.p[0], p.v[0], p.a[0], p.pf, p.vf, p.af, inp.max_velocity[dof], min_vel, inp.max_acceleration[dof], min_acc, inp.max_jerk[dof]);
                        found_profile = step1.get_profile(p, blocks[dof]);
                    } else if (!acc_inf) {
                        PositionSecondOrderStep1 step1(p.p[0], p.v[0], p.pf, p.vf, inp.max_velocity[dof], min_vel, inp.max_acceleration[dof], min_acc);
                        found_profile = step1.get_profile(p, blocks[dof]);
                    } else {
                        PositionFirstOrderStep1 step1(p.p[0], p.pf, inp.max_velocity[dof], min_vel);
                        found_profile = step1.get_profile(p, blocks[dof]);
                    }
                    break;
                case CI_VELOCITY:
                    if (!jerk_inf) {
                        VelocityThirdOrderStep1 step1(p.v[0], p.a[0], p.vf, p.af, inp.max_acceleration[dof], min_acc, inp.max_jerk[dof]);
                        found_profile = step1.get_profile(p, blocks[dof]);
                    } else {
                        VelocitySecondOrderStep1 step1(p.v[0], p.vf, inp.max_acceleration[dof], min_acc);
                        found_profile = step1.get_profile(p, blocks[dof]);
                    }
                    break;
                default: break;  // Acceleration interface: unimplemented (quirk Q20)
            }

            if (!found_profile) {
                const bool has_zero_limits = inp.max_acceleration[dof] == 0.0 || min_acc == 0.0 || inp.max_jerk[dof] == 0.0;
                if (has_zero_limits) {
                    if (handle(throw_mode, ERR_CALCULATOR, "zero limits conflict in step 1, dof: " + std::to_string(dof), out)) { out.code = R_ERROR_ZERO_LIMITS; return out; }
                    out.code = R_ERROR_ZERO_LIMITS;
                    return out;
                }
                if (handle(throw_mode, ERR_CALCULATOR, "error in step 1, dof: " + std::to_string(dof), out)) { out.code = R_ERROR_EXECUTION_TIME_CALCULATION; return out; }
                out.code = R_ERROR_EXECUTION_TIME_CALCULATION;
                return out;
            }

            traj.independent_min_durations[dof] = blocks[dof].t_min;
        }

        const bool discrete_duration = inp.duration_discretization == DD_DISCRETE;
        if (D == 1 && !inp.has_minimum_duration && !discrete_duration) {
            traj.duration = blocks[0].t_min;
            traj.profiles[0] = blocks[0].p_min;
            profile_source[0] = 0;
            traj.cumulative_times[0] = traj.duration;
            out.code = R_WORKING;
            return out;
        }

        int limiting_dof = -1;
        const bool found_synchronization = synchronize(inp.has_minimum_duration, inp.minimum_duration, traj.duration, limiting_dof, traj.profiles, discrete_duration, delta_time);
        if (!found_synchronization) {
            bool has_zero_limits = false;
            for (int dof = 0; dof < D; ++dof) {
                if (inp.max_acceleration[dof] == 0.0 || (inp.has_min_acceleration ? inp.min_acceleration[dof] : -inp.max_acceleration[dof]) == 0.0 || inp.max_jerk[dof] == 0.0) {
                    has_zero_limits = true;
                    break;
                }
            }

            if (has_zero_limits) {
                handle(throw_mode, ERR_CALCULATOR, "zero limits conflict with other degrees of freedom in time synchronization " + std::to_string(traj.duration), out);
                out.code = R_ERROR_ZERO_LIMITS;
                return out;
            }
            handle(throw_mode, ERR_CALCULATOR, "error in time synchronization: " + std::to_string(traj.duration), out);
            out.code = R_ERROR_SYNCHRONIZATION_CALCULATION;
            return out;
        }

        // None Synchronization
        for (int dof = 0; dof < D; ++dof) {
            if (inp.enabled[dof] && inp_per_dof_synchronization[dof] == SYNC_NONE) {
                traj.profiles[dof] = blocks[dof].p_min;
                profile_source[dof] = 0;
                if (blocks[dof].t_min > traj.duration) {
                    traj.duration = blocks[dof].t_min;
                    limiting_dof = dof;
                }
            }
        }
        traj.cumulative_times[0] = traj.duration;
        last_limiting_dof = limiting_dof;

        if (return_error_at_maximal_duration && traj.duration > 7.6e3) {
            out.code = R_ERROR_TRAJECTORY_DURATION;
            return out;
        }

        if (std::fabs(traj.duration - 0.0) < EPS) {
            for (int dof = 0; dof < D; ++dof) { traj.profiles[dof] = blocks[dof].p_min; profile_source[dof] = 0; }
            out.code = R_WORKING;
            return out;
        }

        bool all_none = true, any_phase = false, all_phase_or_none = true;
        for (int dof = 0; dof < D; ++dof) {
            const int s = inp_per_dof_synchronization[dof];
            all_none &= (s == SYNC_NONE);
            any_phase |= (s == SYNC_PHASE);
            all_phase_or_none &= (s == SYNC_PHASE || s == SYNC_NONE);
        }

        if (!discrete_duration && all_none) {
            out.code = R_WORKING;
            return out;
        }

        // Phase Synchronization
        if (limiting_dof >= 0 && any_phase) {
            const Profile p_limiting = traj.profiles[limiting_dof];
            if (is_input_collinear(inp, p_limiting.direction, limiting_dof)) {
                bool found_time_synchronization = true;
                for (int dof = 0; dof < D; ++dof) {
                    if (!inp.enabled[dof] || dof == limiting_dof || inp_per_dof_synchronization[dof] != SYNC_PHASE) continue;

                    Profile& p = traj.profiles[dof];
                    const double t_profile = traj.duration - p.brake.duration - p.accel.duration;

                    for (int i = 0; i < 7; ++i) p.t[i] = p_limiting.t[i];
                    p.control_signs = p_limiting.control_signs;
                    const bool jerk_inf = std::isinf(inp.max_jerk[dof]);
                    const bool acc_inf = std::isinf(inp.max_acceleration[dof]);
                    const double npc = new_phase_control[dof];

                    switch (inp_per_dof_control_interface[dof]) {
                        case CI_POSITION:
                            if (p.control_signs == UDDU) {
                                if (!jerk_inf) {
                                    found_time_synchronization &= p.check_with_timing_full(UDDU, NONE, t_profile, npc, inp.max_velocity[dof], inp_min_velocity[dof], inp.max_acceleration[dof], inp_min_acceleration[dof], inp.max_jerk[dof]);
                                } else if (!acc_inf) {
                                    found_time_synchronization &= p.check_for_second_order_with_timing_full(UDDU, NONE, t_profile, npc, -npc, inp.max_velocity[dof], inp_min_velocity[dof], inp.max_acceleration[dof], inp_min_acceleration[dof]);
                                } else {
                                    found_time_synchronization &= p.check_for_first_order_with_timing_full(UDDU, NONE, t_profile, npc, inp.max_velocity[dof], inp_min_velocity[dof]);
                                }
                            } else {
                                if (!jerk_inf) {
                                    found_time_synchronization &= p.check_with_timing_full(UDUD, NONE, t_profile, npc, inp.max_velocity[dof], inp_min_velocity[dof], inp.max_acceleration[dof], inp_min_acceleration[dof], inp.max_jerk[dof]);
                                } else {
                                    found_time_synchronization &= p.check_for_second_order_with_timing_full(UDUD, NONE, t_profile, npc, -npc, inp.max_velocity[dof], inp_min_velocity[dof], inp.max_acceleration[dof], inp_min_acceleration[dof]);
                                }
                            }
                            break;
                        case CI_VELOCITY:
                            if (p.control_signs == UDDU) {
                                if (!jerk_inf) {
                                    found_time_synchronization &= p.check_for_velocity_with_timing_full(t_profile, UDDU, NONE, npc, inp.max_acceleration[dof], inp_min_acceleration[dof], inp.max_jerk[dof]);
                                } else {
                                    found_time_synchronization &= p.check_for_second_order_velocity_with_timing_a_limits(UDDU, NONE, t_profile, npc, inp.max_acceleration[dof], inp_min_acceleration[dof]);
                                }
                            } else {
                                if (!jerk_inf) {
                                    found_time_synchronization &= p.check_for_velocity_with_timing_full(t_profile, UDUD, NONE, npc, inp.max_acceleration[dof], inp_min_acceleration[dof], inp.max_jerk[dof]);
                                } else {
                                    found_time_synchronization &= p.check_for_second_order_velocity_with_timing_a_limits(UDUD, NONE, t_profile, npc, inp.max_acceleration[dof], inp_min_acceleration[dof]);
                                }
                            }
                            break;
                        default: break;
                    }

                    p.limits = p_limiting.limits;  // after the check (quirk Q25)
                    profile_source[dof] = 3;
                }

                if (found_time_synchronization && all_phase_or_none) {
                    out.code = R_WORKING;
                    return out;
                }
            }
        }

        // Time Synchronization
        for (int dof = 0; dof < D; ++dof) {
            const bool skip_synchronization = (dof == limiting_dof || inp_per_dof_synchronization[dof] == SYNC_NONE) && !discrete_duration;
            if (!inp.enabled[dof] || skip_synchronization) continue;

            Profile& p = traj.profiles[dof];
            const double t_profile = traj.duration - p.brake.duration - p.accel.duration;

            if (inp_per_dof_synchronization[dof] == SYNC_TIME_IF_NECESSARY && std::fabs(inp.target_velocity[dof]) < eps && std::fabs(inp.target_acceleration[dof]) < eps) {
                traj.profiles[dof] = blocks[dof].p_min;
                profile_source[dof] = 0;
                continue;
            }

            // Extremal profile re-use; b is only looked at when a is absent (quirk Q3)
            if (std::fabs(t_profile - blocks[dof].t_min) < 2.0 * eps) {
                traj.profiles[dof] = blocks[dof].p_min;
                profile_source[dof] = 0;
                continue;
            } else if (blocks[dof].has_a) {
                if (std::fabs(t_profile - blocks[dof].a.right) < 2.0 * eps) {
                    traj.profiles[dof] = blocks[dof].a.profile;
                    profile_source[dof] = 1;
                    continue;
                }
            } else if (blocks[dof].has_b) {
                if (std::fabs(t_profile - blocks[dof].b.right) < 2.0 * eps) {
                    traj.profiles[dof] = blocks[dof].b.profile;
                    profile_source[dof] = 2;
                    continue;
                }
            }

            bool found_time_synchronization = false;
            const bool jerk_inf = std::isinf(inp.max_jerk[dof]);
            const bool acc_inf = std::isinf(inp.max_acceleration[dof]);
            switch (inp_per_dof_control_interface[dof]) {
                case CI_POSITION:
                    if (!jerk_inf) {
                        PositionThirdOrderStep2 step2(t_profile, p.p[0], p.v[0], p.a[0], p.pf, p.vf, p.af, inp.max_velocity[dof], inp_min_velocity[dof], inp.max_acceleration[dof], inp_min_acceleration[dof], inp.max_jerk[dof]);
                        found_time_synchronization = step2.get_profile(p);
                    } else if (!acc_inf) {
                        PositionSecondOrderStep2 step2(t_profile, p.p[0], p.v[0], p.pf, p.vf, inp.max_velocity[dof], inp_min_velocity[dof], inp.max_acceleration[dof], inp_min_acceleration[dof]);
                        found_time_synchronization = step2.get_profile(p);
                    } else {
                        PositionFirstOrderStep2 step2(t_profile, p.p[0], p.pf, inp.max_velocity[dof], inp_min_velocity[dof]);
                        found_time_synchronization = step2.get_profile(p);
                    }
                    break;
                case CI_VELOCITY:
                    if (!jerk_inf) {
                        VelocityThirdOrderStep2 step2(t_profile, p.v[0], p.a[0], p.vf, p.af, inp.max_acceleration[dof], inp_min_acceleration[dof], inp.max_jerk[dof]);
                        found_time_synchronization = step2.get_profile(p);
                    } else {
                        VelocitySecondOrderStep2 step2(t_profile, p.v[0], p.vf, inp.max_acceleration[dof], inp_min_acceleration[dof]);
                        found_time_synchronization = step2.get_profile(p);
                    }
                    break;
                default: break;
            }
            profile_source[dof] = 4;

            if (!found_time_synchronization) {
                handle(throw_mode, ERR_CALCULATOR, "error in step 2, dof: " + std::to_string(dof), out);
                out.code = R_ERROR_EXECUTION_TIME_CALCULATION;
                return out;
            }
        }

        out.code = R_WORKING;
        return out;
    }
};

// output_parameter.rs:36-120
struct OutputParameter {
    int degrees_of_freedom = 0;
    Trajectory trajectory;
    std::vector<double> new_position, new_velocity, new_acceleration, new_jerk;
    double time = 0.0;
    int new_section = 0;
    bool did_section_change = false, new_calculation = false, was_calculation_interrupted = false;

    explicit OutputParameter(int dofs) : degrees_of_freedom(dofs), trajectory(dofs) {
        new_position.assign(dofs, 0.0); new_velocity.assign(dofs, 0.0); new_acceleration.assign(dofs, 0.0); new_jerk.assign(dofs, 0.0);
    }

    void pass_to_input(InputParameter& input) const {
        input.current_position = new_position;
        input.current_velocity = new_velocity;
        input.current_acceleration = new_acceleration;
    }
};

// ruckig.rs:76-321
struct Ruckig {
    InputParameter current_input;
    bool current_input_initialized = false;
    TargetCalculator calculator;
    int degrees_of_freedom;
    double delta_time;
    bool throw_mode;  // ThrowErrorHandler (true) / IgnoreErrorHandler (false)

    Ruckig(int dofs, double delta_time_, bool throw_mode_)
        : current_input(dofs), calculator(dofs), degrees_of_freedom(dofs), delta_time(delta_time_), throw_mode(throw_mode_) {}

    void reset() { current_input_initialized = false; }

    // ruckig.rs:150-171
    Outcome validate_input(const InputParameter& input, bool check_current, bool check_target) const {
        Outcome out;
        const std::string m = input.validate(check_current, check_target);
        if (!m.empty() && throw_mode) { out.err = ERR_VALIDATION; out.code = R_ERROR_INVALID_INPUT; out.msg = m; return out; }
        if (delta_time <= 0.0 && input.duration_discretization != DD_CONTINUOUS) {
            if (throw_mode) { out.err = ERR_VALIDATION; out.code = R_ERROR_INVALID_INPUT; out.msg = "delta time (control rate) parameter should be larger than zero."; return out; }
        }
        return out;
    }

    // ruckig.rs:210-218
    Outcome calculate(const InputParameter& input, Trajectory& traj) {
        Outcome v = validate_input(input, false, true);
        if (v.is_err()) return v;
        return calculator.calculate(input, traj, delta_time, throw_mode);
    }

    // ruckig.rs:266-321 (output.time is never reset, quirk Q1; Ok(Error*) codes of calculate are dropped)
    Outcome update(const InputParameter& input, OutputParameter& output) {
        Outcome out;
        output.new_calculation = false;

        if (!current_input_initialized || !input.equals(current_input)) {
            Outcome c = calculate(input, output.trajectory);
            if (c.is_err()) return c;

            output.new_calculation = true;
            current_input = input;
            current_input_initialized = true;
        }

        const int old_section = output.new_section;
        output.time += delta_time;
        int section_copy = output.new_section;  // passed as a temporary (quirk Q2)
        output.trajectory.at_time(output.time, output.new_position.data(), output.new_velocity.data(), output.new_acceleration.data(), output.new_jerk.data(), &section_copy);
        output.did_section_change = output.new_section > old_section;

        output.pass_to_input(current_input);

        if (output.time > output.trajectory.duration) { out.code = R_FINISHED; return out; }
        out.code = R_WORKING;
        return out;
    }
};

}  // namespace rko
