// host_api_test.cpp — the reference's first integration tests (rsruckig test_suite/tests/tests.rs) written against the C++
// host-side mirror include/ruckig_b200.hpp, so that they read like the originals. Built and run by tests/test_cpp_host_api.py.
//
//   host_api_test            on a B200: runs the ports, prints "OK <n> checks", exit 0
//   host_api_test --no-gpu   without a device: the library must refuse to work (there is no CPU path), exit 0 if it does
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/ruckig_b200.hpp"

using namespace ruckig_b200;

static int checks = 0, failures = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        ++checks;                                                                \
        if (!(cond)) { ++failures; std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)
static bool approx(double a, double b, double eps = 1e-4) { return std::fabs(a - b) <= eps; }

// tests.rs:33-77
static void test_at_time() {
    Ruckig<1, ThrowErrorHandler> otg(0.005);
    InputParameter<1> input;
    input.current_position = {0.0};
    input.target_position = {1.0};
    input.max_velocity = {1.0};
    input.max_acceleration = {1.0};
    input.max_jerk = {1.0};

    Trajectory<1> traj;
    CHECK(otg.calculate(input, traj) == RuckigResult::Working);
    OutputParameter<1> output;
    CHECK(otg.update(input, output) == RuckigResult::Working);
    CHECK(approx(traj.get_duration(), 3.1748));

    std::vector<double> p, v, a;
    traj.at_time(0.0, &p, &v, &a);
    CHECK(approx(p[0], input.current_position[0]));
    traj.at_time(3.1748 / 2.0, &p, &v, &a);
    CHECK(approx(p[0], 0.5));
}

// tests.rs:81-303 (the parts that go through calculate / update / at_time / the query helpers)
static void test_secondary() {
    Ruckig<3, ThrowErrorHandler> otg(0.005);
    InputParameter<3> input;
    OutputParameter<3> output;
    input.current_position = {0.0, -2.0, 0.0};
    input.target_position = {1.0, -3.0, 2.0};
    input.target_velocity = {0.0, 0.3, 0.0};
    input.max_velocity = {1.0, 1.0, 1.0};
    input.max_acceleration = {1.0, 1.0, 1.0};
    input.max_jerk = {1.0, 1.0, 1.0};

    Trajectory<3> traj;
    otg.calculate(input, traj);
    CHECK(approx(traj.get_duration(), 4.0));

    otg.update(input, output);
    CHECK(output.new_calculation);
    CHECK(approx(output.trajectory.get_duration(), 4.0));

    std::vector<double> p, v, a, j;
    size_t section = 99;
    output.trajectory.at_time(0.0, &p, &v, &a);
    CHECK(p == input.current_position && v == input.current_velocity && a == input.current_acceleration);
    output.trajectory.at_time(output.trajectory.get_duration(), &p, &v, &a);
    for (int d = 0; d < 3; ++d) CHECK(approx(p[d], input.target_position[d]) && approx(v[d], input.target_velocity[d]) && approx(a[d], input.target_acceleration[d]));
    output.trajectory.at_time(2.0, &p, &v, &a, &j, &section);
    CHECK(approx(p[0], 0.5) && approx(p[1], -2.6871268303003437) && approx(p[2], 1.0));
    CHECK(j[0] == 0.0 && j[1] == 0.0 && j[2] == -1.0);
    CHECK(section == 0);
    output.trajectory.at_time(5.0, &p, &v, &a, &j, &section);
    CHECK(j[0] == 0.0 && j[1] == 0.0 && j[2] == 0.0);
    CHECK(section == 1);

    const auto& imd = output.trajectory.get_independent_min_durations();
    CHECK(approx(imd[0], 3.1748021039) && approx(imd[1], 3.6860977315) && approx(imd[2], output.trajectory.get_duration()));

    // a target velocity above its limit: ValidationError under the throwing handler, the cycle is not advanced
    input.target_velocity = {2.0, 0.3, 0.0};
    bool thrown = false;
    const double time_before = output.time;
    try {
        otg.update(input, output);
    } catch (const RuckigError& e) {
        thrown = e.kind() == RuckigError::Kind::ValidationError && std::string(e.what()).find("exceeds its maximum velocity limit") != std::string::npos;
    }
    CHECK(thrown);
    CHECK(!output.new_calculation);
    CHECK(output.time == time_before);

    input.target_velocity = {0.2, -0.3, 0.8};
    otg.update(input, output);
    CHECK(output.new_calculation);

    // the same input again: no new calculation, time advances by delta_time
    const double t1 = output.time;
    output.pass_to_input(input);
    CHECK(otg.update(input, output) == RuckigResult::Working);
    CHECK(!output.new_calculation);
    CHECK(output.time == t1 + 0.005);

    // query helpers on the first trajectory
    const auto extrema = traj.get_position_extrema();
    CHECK(extrema.size() == 3 && approx(extrema[0].min, 0.0) && approx(extrema[0].max, 1.0));
    const auto hit = traj.get_first_time_at_position(0, 0.5);
    CHECK(hit.has_value() && approx(*hit, 2.0));
    CHECK(!traj.get_first_time_at_position(0, 5.0).has_value());

    const auto profiles = traj.get_profiles();
    CHECK(profiles.size() == 1 && profiles[0].size() == 3);
    CHECK(approx(profiles[0][2].t_sum[6] + profiles[0][2].brake_duration, traj.get_duration()));
}

// validate_input under both handlers (ruckig.rs:150-171, error.rs:98-122)
static void test_error_handlers() {
    InputParameter<0> input(2);  // DOF = 0: the number of DoFs is given at run time (Ruckig::new(Some(2), ..))
    input.max_velocity = {1.0, 1.0};
    input.max_acceleration = {1.0, 1.0};
    input.max_jerk = {1.0, -1.0};
    Ruckig<0, IgnoreErrorHandler> lenient(0.01, 2);
    CHECK(!lenient.validate_input(input, false, true));
    Ruckig<0, ThrowErrorHandler> strict(0.01, 2);
    bool thrown = false;
    try {
        strict.validate_input(input, false, true);
    } catch (const RuckigError& e) {
        thrown = e.kind() == RuckigError::Kind::ValidationError && std::string(e.what()) == "Maximum jerk limit of DoF 1 should be larger than or equal to zero.";
    }
    CHECK(thrown);
}

// the batched entry point against the single-problem one
static void test_batch_matches_single() {
    const int64_t n = 5;
    const size_t D = 2;
    std::vector<double> p0(D * n), v0(D * n, 0.0), a0(D * n, 0.0), pf(D * n), vf(D * n, 0.0), af(D * n, 0.0), vm(D * n, 1.0), am(D * n, 2.0), jm(D * n, 3.0);
    for (int64_t i = 0; i < n; ++i)
        for (size_t d = 0; d < D; ++d) { p0[d * n + i] = 0.1 * (double)i - 0.3 * (double)d; pf[d * n + i] = 1.0 + 0.25 * (double)i * ((double)d + 1.0); }
    rk_batch_in in{};
    in.current_position = p0.data(); in.current_velocity = v0.data(); in.current_acceleration = a0.data();
    in.target_position = pf.data(); in.target_velocity = vf.data(); in.target_acceleration = af.data();
    in.max_velocity = vm.data(); in.max_acceleration = am.data(); in.max_jerk = jm.data();
    BatchRuckig<2, ThrowErrorHandler> batch(0.005);
    std::vector<int32_t> status(n, -1);
    std::vector<double> duration(n, 0.0);
    batch.calculate(n, in, {}, status.data(), duration.data());
    std::vector<double> pos(3 * D * n);
    batch.at_time(0.0, 0.1, 3, false, pos.data());

    Ruckig<2, ThrowErrorHandler> otg(0.005);
    for (int64_t i = 0; i < n; ++i) {
        InputParameter<2> input;
        for (size_t d = 0; d < D; ++d) { input.current_position[d] = p0[d * n + i]; input.target_position[d] = pf[d * n + i]; }
        input.max_velocity = {1.0, 1.0}; input.max_acceleration = {2.0, 2.0}; input.max_jerk = {3.0, 3.0};
        Trajectory<2> traj;
        CHECK((int32_t)otg.calculate(input, traj) == status[i]);
        CHECK(traj.get_duration() == duration[i]);  // the same kernels: bit-identical
        std::vector<double> p;
        double t = 0.0;
        for (int k = 0; k < 3; ++k) {
            t += 0.1;
            traj.at_time(t, &p);
            for (size_t d = 0; d < D; ++d) CHECK(p[d] == pos[((size_t)k * D + d) * n + i]);
        }
    }
    CHECK(batch.error_message(0).empty());

    // the compact result layout and the multi-handle entry (two handles on device 0) return the same bits
    BatchRuckig<2, ThrowErrorHandler> compact(0.005, std::nullopt, 0, RK_RESULT_COMPACT);
    std::vector<int32_t> status_c(n, -1);
    std::vector<double> duration_c(n, 0.0), pos_c(3 * D * n), t_full(7 * D * n), t_compact(7 * D * n);
    compact.calculate(n, in, {}, status_c.data(), duration_c.data());
    compact.at_time(0.0, 0.1, 3, false, pos_c.data());
    batch.read(RK_FIELD_T, t_full.data());
    compact.read(RK_FIELD_T, t_compact.data());
    CHECK(status_c == status && duration_c == duration && pos_c == pos && t_compact == t_full);
    MultiBatchRuckig<2, ThrowErrorHandler> multi(0.005, {0, 0});
    std::vector<int32_t> status_m(n, -1);
    std::vector<double> duration_m(n, 0.0);
    multi.calculate(n, in, {}, status_m.data(), duration_m.data());
    CHECK(multi.parts() == 2 && status_m == status && duration_m == duration);

    // RK_LIMITS_PER_DOF: one limit per DoF for the whole batch (the limits above are the same for every trajectory)
    const std::vector<double> vm_d(D, 1.0), am_d(D, 2.0), jm_d(D, 3.0);
    rk_batch_in shared = in;
    shared.max_velocity = vm_d.data(); shared.max_acceleration = am_d.data(); shared.max_jerk = jm_d.data();
    BatchSettings fleet;
    fleet.limits_per_dof = true;
    std::vector<int32_t> status_s(n, -1);
    std::vector<double> duration_s(n, 0.0), t_shared(7 * D * n);
    compact.calculate(n, shared, fleet, status_s.data(), duration_s.data());
    compact.read(RK_FIELD_T, t_shared.data());
    CHECK(status_s == status && duration_s == duration && t_shared == t_full);
}

// Never called: instantiates the device-buffer classes so that the whole header is compiled under -Wall -Wextra -Werror (their
// behaviour is covered through the same C calls by tests/test_rollout.py).
[[maybe_unused]] static int64_t instantiate_device_buffer_classes(const rk_batch_in& in, const rk_update_out& out) {
    BatchRollout<7, IgnoreErrorHandler> rollout(4, 0.005);
    rollout.reset();
    const int64_t recalculated = rollout.update(in, out);
    rollout.sync();
    return recalculated + (rollout.trajectories() != nullptr) + (rollout.stream() != nullptr);
}

int main(int argc, char** argv) {
    if (argc > 1 && std::strcmp(argv[1], "--no-gpu") == 0) {
        try {
            Ruckig<3> otg(0.005);
        } catch (const RuckigError& e) {
            std::printf("refused as expected: %s\n", e.what());
            return e.kind() == RuckigError::Kind::Library ? 0 : 1;
        }
        std::printf("a CUDA device is present: --no-gpu has nothing to check\n");
        return 0;
    }
    try {
        test_at_time();
        test_secondary();
        test_error_handlers();
        test_batch_matches_single();
    } catch (const std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 2;
    }
    if (failures) { std::printf("%d of %d checks failed\n", failures, checks); return 1; }
    std::printf("OK %d checks\n", checks);
    return 0;
}
