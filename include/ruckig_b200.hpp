// ruckig_b200.hpp — C++17 host-side mirror of rsruckig's public API on top of the C ABI (ruckig_b200.h).
//
// The reference is a compiled (Rust) crate and no Rust toolchain exists where this repository is built, so next to the Python
// mirror the tests go through (rsruckig_b200/api.py) this header gives compiled host code the reference's own names,
// argument meaning and error behaviour (rsruckig v2.1.2; citations relative to lib/src/rsruckig/):
//
//   Ruckig<DOF, E>::{Ruckig, reset, validate_input, calculate, update}        ruckig.rs:110-321
//   InputParameter<DOF>                                                        input_parameter.rs:147-243
//   OutputParameter<DOF>::{pass_to_input}                                      output_parameter.rs:36-120
//   Trajectory<DOF>::{at_time, get_duration, get_independent_min_durations,
//                     get_profiles, get_position_extrema, get_first_time_at_position}   trajectory.rs:158-246
//   RuckigResult, RuckigError, ThrowErrorHandler / IgnoreErrorHandler          result.rs:34-61, error.rs:15-122
//   BatchRuckig<DOF, E>: the batched entry point (n problems in SoA arrays [dof][n], host or device memory)
//   BatchRollout<DOF, E>: n closed control loops on the device (the batched Ruckig::update)
//
// As in the reference DOF = 0 means "number of DoFs given at run time" (Ruckig::new(Some(n), ..)), any other DOF fixes it.
// Every calculation runs on the GPU — a single problem is a batch of one; there is no CPU path behind these classes:
// without a CUDA device the first object that needs the library throws RuckigError (kind Library, RK_ERR_NO_DEVICE).
// Header-only; link against libruckig_b200.so.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ruckig_b200.h"

namespace ruckig_b200 {

enum class ControlInterface : int32_t { Position = RK_POSITION, Velocity = RK_VELOCITY, Acceleration = RK_ACCELERATION };  // input_parameter.rs:35-45
enum class Synchronization : int32_t { Time = RK_SYNC_TIME, TimeIfNecessary = RK_SYNC_TIME_IF_NECESSARY, Phase = RK_SYNC_PHASE, None = RK_SYNC_NONE };  // :72-85
enum class DurationDiscretization : int32_t { Continuous = RK_CONTINUOUS, Discrete = RK_DISCRETE };  // :110-117

enum class RuckigResult : int32_t {  // result.rs:34-61
    Working = 0,
    Finished = 1,
    Error = -1,
    ErrorInvalidInput = -100,
    ErrorTrajectoryDuration = -101,
    ErrorPositionalLimits = -102,
    ErrorZeroLimits = -104,
    ErrorExecutionTimeCalculation = -110,
    ErrorSynchronizationCalculation = -111
};

enum class ReachedLimits : int32_t { Acc0Acc1Vel = 0, Vel, Acc0, Acc1, Acc0Acc1, Acc0Vel, Acc1Vel, None };  // profile.rs:25-34
enum class Direction : int32_t { Up = 0, Down = 1 };                                                         // profile.rs:37-41
enum class ControlSigns : int32_t { UDDU = 0, UDUD = 1 };                                                    // profile.rs:44-49

// error.rs:15-34. `Library` is not a reference variant: it reports a failure of the CUDA library itself (no device, out of
// memory, bad argument), which has no counterpart in a CPU crate.
class RuckigError : public std::runtime_error {
public:
    enum class Kind { ValidationError, CalculatorError, Library };
    RuckigError(Kind kind, const std::string& what) : std::runtime_error(what), kind_(kind) {}
    Kind kind() const { return kind_; }

private:
    Kind kind_;
};

// error.rs:98-122: what happens to a validation / calculator error. Under ThrowErrorHandler it is thrown (the reference
// returns Err); under IgnoreErrorHandler a validation failure is ignored and the calculation proceeds, a calculator failure
// comes back as its result code.
struct ThrowErrorHandler { static constexpr bool throws = true; };
struct IgnoreErrorHandler { static constexpr bool throws = false; };

namespace detail {

inline void check(rk_status st) {
    if (st != RK_OK) throw RuckigError(RuckigError::Kind::Library, std::string(rk_status_string(st)) + ": " + rk_last_error());
}

// One rk_handle per device and process, shared by every object of this header (calls on one handle are stream ordered).
// It lives as long as the process: destroying it from a static destructor would run after the CUDA runtime's own teardown.
class Engine {
public:
    static Engine& get(int32_t device = 0) {
        static std::array<Engine*, 64> engines{};
        if (device < 0 || device >= (int32_t)engines.size()) throw RuckigError(RuckigError::Kind::Library, "device index out of range");
        if (!engines[device]) {
            std::unique_ptr<Engine> e(new Engine());
            check(rk_create(device, &e->h));
            engines[device] = e.release();
        }
        return *engines[device];
    }
    rk_handle* h = nullptr;

private:
    Engine() = default;
};

// texts of InputParameter::validate (input_parameter.rs:251-420), keyed by the RK_INVALID_* id of the first failing check
inline std::string validation_message(int32_t reason) {
    static const char* const text[] = {
        "invalid input",
        "Maximum jerk limit of DoF {} should be larger than or equal to zero.",
        "maximum acceleration limit of DoF {} should be larger than or equal to zero.",
        "minimum acceleration limit of DoF {} should be smaller than or equal to zero.",
        "current acceleration of DoF {} should be a valid number.",
        "target acceleration of DoF {} should be a valid number.",
        "current acceleration of DoF {} exceeds its maximum acceleration limit.",
        "current acceleration of DoF {} undercuts its minimum acceleration limit.",
        "target acceleration of DoF {} exceeds its maximum acceleration limit.",
        "target acceleration of DoF {} undercuts its minimum acceleration limit.",
        "current velocity of DoF {} should be a valid number.",
        "target velocity of DoF {} should be a valid number.",
        "current position of DoF {} should be a valid number.",
        "target position of DoF {} should be a valid number.",
        "maximum velocity limit of DoF {} should be larger than or equal to zero.",
        "minimum velocity limit of DoF {} should be smaller than or equal to zero.",
        "current velocity of DoF {} exceeds its maximum velocity limit.",
        "current velocity of DoF {} undercuts its minimum velocity limit.",
        "target velocity of DoF {} exceeds its maximum velocity limit.",
        "target velocity of DoF {} undercuts its minimum velocity limit.",
        "DoF {} will inevitably reach a velocity from the current kinematic state that will exceed its maximum velocity limit.",
        "DoF {} will inevitably reach a velocity from the current kinematic state that will undercut its minimum velocity limit.",
        "DoF {} will inevitably have reached a velocity from the target kinematic state that will exceed its maximum velocity limit.",
        "DoF {} will inevitably have reached a velocity from the target kinematic state that will undercut its minimum velocity limit.",
        "delta time (control rate) parameter should be larger than zero."};
    const int32_t id = reason >> 8, dof = reason & 0xFF;
    std::string s = text[(id >= 1 && id <= RK_INVALID_DELTA_TIME) ? id : 0];
    const size_t at = s.find("{}");
    if (at != std::string::npos) s.replace(at, 2, std::to_string(dof));
    return s;
}

// texts of calculator_target.rs:459-470, :509-517, :820-825 rebuilt from the status and the error-detail word
inline std::string calculator_message(RuckigResult code, int32_t detail) {
    const int32_t stage = detail >> 8, dof = detail & 0xFF;
    const bool zero = code == RuckigResult::ErrorZeroLimits;
    if (stage == 1) return (zero ? "zero limits conflict in step 1, dof: " : "error in step 1, dof: ") + std::to_string(dof);
    if (stage == 2) return zero ? "zero limits conflict with other degrees of freedom in time synchronization" : "error in time synchronization";
    return "error in step 2, dof: " + std::to_string(dof);
}

inline bool is_calculator_error(RuckigResult r) {
    return r == RuckigResult::ErrorZeroLimits || r == RuckigResult::ErrorExecutionTimeCalculation || r == RuckigResult::ErrorSynchronizationCalculation;
}

inline size_t resolve_dofs(size_t DOF, std::optional<size_t> dofs) {
    const size_t d = DOF != 0 ? DOF : dofs.value_or(0);
    if (d < 1 || d > RK_MAX_DOF) throw RuckigError(RuckigError::Kind::Library, "degrees of freedom must be 1.." + std::to_string(RK_MAX_DOF));
    return d;
}

}  // namespace detail

// ---------------------------------------------------------------------------------------------------------- InputParameter
// input_parameter.rs:147-243, same defaults (max_acceleration = max_jerk = +inf, all DoFs enabled).
template <size_t DOF>
struct InputParameter {
    using Vector = std::vector<double>;
    size_t degrees_of_freedom;
    ControlInterface control_interface = ControlInterface::Position;
    Synchronization synchronization = Synchronization::Time;
    DurationDiscretization duration_discretization = DurationDiscretization::Continuous;
    Vector current_position, current_velocity, current_acceleration;
    Vector target_position, target_velocity, target_acceleration;
    Vector max_velocity, max_acceleration, max_jerk;
    std::optional<Vector> min_velocity, min_acceleration;
    std::vector<bool> enabled;
    std::optional<std::vector<ControlInterface>> per_dof_control_interface;
    std::optional<std::vector<Synchronization>> per_dof_synchronization;
    std::optional<double> minimum_duration;
    std::optional<double> interrupt_calculation_duration;  // never read by the reference either

    explicit InputParameter(std::optional<size_t> dofs = std::nullopt)
        : degrees_of_freedom(detail::resolve_dofs(DOF, dofs)),
          current_position(degrees_of_freedom, 0.0), current_velocity(degrees_of_freedom, 0.0), current_acceleration(degrees_of_freedom, 0.0),
          target_position(degrees_of_freedom, 0.0), target_velocity(degrees_of_freedom, 0.0), target_acceleration(degrees_of_freedom, 0.0),
          max_velocity(degrees_of_freedom, 0.0), max_acceleration(degrees_of_freedom, std::numeric_limits<double>::infinity()),
          max_jerk(degrees_of_freedom, std::numeric_limits<double>::infinity()), enabled(degrees_of_freedom, true) {}

    // input_parameter.rs:190-211 (interrupt_calculation_duration does not take part); f64 ==, so NaN != NaN
    bool operator==(const InputParameter& o) const {
        return current_position == o.current_position && current_velocity == o.current_velocity && current_acceleration == o.current_acceleration &&
               target_position == o.target_position && target_velocity == o.target_velocity && target_acceleration == o.target_acceleration &&
               max_velocity == o.max_velocity && max_acceleration == o.max_acceleration && max_jerk == o.max_jerk && enabled == o.enabled &&
               minimum_duration == o.minimum_duration && min_velocity == o.min_velocity && min_acceleration == o.min_acceleration &&
               control_interface == o.control_interface && synchronization == o.synchronization &&
               duration_discretization == o.duration_discretization && per_dof_control_interface == o.per_dof_control_interface &&
               per_dof_synchronization == o.per_dof_synchronization;
    }
    bool operator!=(const InputParameter& o) const { return !(*this == o); }
};

namespace detail {

// The n = 1 view of an InputParameter for the C ABI (the buffers it points into live as long as this object).
template <size_t DOF>
struct SingleProblem {
    rk_batch_desc desc{};
    rk_batch_in in{};
    std::vector<uint8_t> enabled;
    std::vector<int32_t> per_dof_ci, per_dof_sync;
    double minimum_duration = 0.0;

    SingleProblem(const InputParameter<DOF>& ip, double delta_time, bool throws) {
        const size_t D = ip.degrees_of_freedom;
        auto sized = [D](const std::vector<double>& v, const char* name) -> const double* {
            if (v.size() != D) throw RuckigError(RuckigError::Kind::Library, std::string(name) + " has the wrong number of elements");
            return v.data();
        };
        desc.n = 1;
        desc.dof = (int32_t)D;
        desc.control_interface = (int32_t)ip.control_interface;
        desc.synchronization = (int32_t)ip.synchronization;
        desc.duration_discretization = (int32_t)ip.duration_discretization;
        desc.error_mode = throws ? RK_ERRORS_THROW : RK_ERRORS_IGNORE;
        desc.memory_space = RK_MEM_HOST;
        desc.delta_time = delta_time;
        if (ip.per_dof_control_interface) {
            for (ControlInterface c : *ip.per_dof_control_interface) per_dof_ci.push_back((int32_t)c);
            if (per_dof_ci.size() != D) throw RuckigError(RuckigError::Kind::Library, "per_dof_control_interface has the wrong number of elements");
            desc.per_dof_control_interface = per_dof_ci.data();
        }
        if (ip.per_dof_synchronization) {
            for (Synchronization s : *ip.per_dof_synchronization) per_dof_sync.push_back((int32_t)s);
            if (per_dof_sync.size() != D) throw RuckigError(RuckigError::Kind::Library, "per_dof_synchronization has the wrong number of elements");
            desc.per_dof_synchronization = per_dof_sync.data();
        }
        in.current_position = sized(ip.current_position, "current_position");
        in.current_velocity = sized(ip.current_velocity, "current_velocity");
        in.current_acceleration = sized(ip.current_acceleration, "current_acceleration");
        in.target_position = sized(ip.target_position, "target_position");
        in.target_velocity = sized(ip.target_velocity, "target_velocity");
        in.target_acceleration = sized(ip.target_acceleration, "target_acceleration");
        in.max_velocity = sized(ip.max_velocity, "max_velocity");
        in.max_acceleration = sized(ip.max_acceleration, "max_acceleration");
        in.max_jerk = sized(ip.max_jerk, "max_jerk");
        if (ip.min_velocity) in.min_velocity = sized(*ip.min_velocity, "min_velocity");
        if (ip.min_acceleration) in.min_acceleration = sized(*ip.min_acceleration, "min_acceleration");
        if (ip.enabled.size() != D) throw RuckigError(RuckigError::Kind::Library, "enabled has the wrong number of elements");
        enabled.resize(D);
        for (size_t d = 0; d < D; ++d) enabled[d] = ip.enabled[d] ? 1 : 0;
        in.enabled = enabled.data();
        if (ip.minimum_duration) {
            minimum_duration = *ip.minimum_duration;
            in.minimum_duration = &minimum_duration;
        }
    }
    SingleProblem(const SingleProblem&) = delete;
    SingleProblem& operator=(const SingleProblem&) = delete;
};

}  // namespace detail

template <size_t DOF, class E>
class Ruckig;

// ------------------------------------------------------------------------------------------------------ Profile / Trajectory
// Host copy of one DoF's Profile (profile.rs:76-118).
struct Profile {
    std::array<double, 7> t{}, t_sum{}, j{};
    std::array<double, 8> a{}, v{}, p{};
    double brake_duration = 0.0;
    std::array<double, 2> brake_t{};
    ReachedLimits limits = ReachedLimits::None;
    Direction direction = Direction::Up;
    ControlSigns control_signs = ControlSigns::UDDU;
};

// profile.rs:52-62: extrema of the position over the whole motion
struct Bound {
    double min = 0.0, max = 0.0, t_min = 0.0, t_max = 0.0;
};

// trajectory.rs:14-21. The profiles live on the GPU (an rk_trajectories of one); they are read on demand.
template <size_t DOF>
class Trajectory {
public:
    using Vector = std::vector<double>;
    size_t degrees_of_freedom;

    explicit Trajectory(std::optional<size_t> dofs = std::nullopt)
        : degrees_of_freedom(detail::resolve_dofs(DOF, dofs)), independent_min_durations_(degrees_of_freedom, 0.0) {}
    ~Trajectory() { if (dev_) rk_trajectories_destroy(dev_); }
    Trajectory(Trajectory&& o) noexcept { *this = std::move(o); }
    Trajectory& operator=(Trajectory&& o) noexcept {
        if (this != &o) {
            if (dev_) rk_trajectories_destroy(dev_);
            degrees_of_freedom = o.degrees_of_freedom; duration_ = o.duration_;
            independent_min_durations_ = std::move(o.independent_min_durations_);
            dev_ = o.dev_; handle_ = o.handle_; o.dev_ = nullptr;
        }
        return *this;
    }
    Trajectory(const Trajectory&) = delete;
    Trajectory& operator=(const Trajectory&) = delete;

    // trajectory.rs:158-189; every output is optional (nullptr = not wanted), as the reference's Option<&mut ..> arguments
    void at_time(double time, Vector* new_position, Vector* new_velocity = nullptr, Vector* new_acceleration = nullptr,
                 Vector* new_jerk = nullptr, size_t* new_section = nullptr) const {
        require_calculated();
        const size_t D = degrees_of_freedom;
        // sample k = 0 is taken after one addition of delta_time to time_start: time + 0.0 == time
        rk_sample_desc sd{time, 0.0, 1, RK_MEM_HOST};
        auto ptr = [D](Vector* v) -> double* { if (!v) return nullptr; v->resize(D); return v->data(); };
        int32_t section = 0;
        rk_sample_out so{ptr(new_position), ptr(new_velocity), ptr(new_acceleration), ptr(new_jerk), new_section ? &section : nullptr};
        detail::check(rk_at_time_batch(handle_, dev_, &sd, &so));
        if (new_section) *new_section = (size_t)section;
    }

    double get_duration() const { return duration_; }                                                   // trajectory.rs:197-199
    const Vector& get_independent_min_durations() const { return independent_min_durations_; }          // :207-209
    Vector get_intermediate_durations() const { return Vector{duration_}; }                             // :202-204 (one section)

    // trajectory.rs:192-194: one section of `dof` profiles
    std::vector<std::vector<Profile>> get_profiles() const {
        require_calculated();
        const size_t D = degrees_of_freedom;
        std::vector<Profile> out(D);
        std::vector<double> buf(8 * D);
        auto rd = [&](int32_t field, size_t slots) { buf.assign(slots * D, 0.0); detail::check(rk_trajectories_read(handle_, dev_, field, buf.data(), RK_MEM_HOST)); };
        rd(RK_FIELD_T, 7);     for (size_t d = 0; d < D; ++d) for (size_t k = 0; k < 7; ++k) out[d].t[k] = buf[k * D + d];
        rd(RK_FIELD_T_SUM, 7); for (size_t d = 0; d < D; ++d) for (size_t k = 0; k < 7; ++k) out[d].t_sum[k] = buf[k * D + d];
        rd(RK_FIELD_J, 7);     for (size_t d = 0; d < D; ++d) for (size_t k = 0; k < 7; ++k) out[d].j[k] = buf[k * D + d];
        rd(RK_FIELD_A, 8);     for (size_t d = 0; d < D; ++d) for (size_t k = 0; k < 8; ++k) out[d].a[k] = buf[k * D + d];
        rd(RK_FIELD_V, 8);     for (size_t d = 0; d < D; ++d) for (size_t k = 0; k < 8; ++k) out[d].v[k] = buf[k * D + d];
        rd(RK_FIELD_P, 8);     for (size_t d = 0; d < D; ++d) for (size_t k = 0; k < 8; ++k) out[d].p[k] = buf[k * D + d];
        rd(RK_FIELD_BRAKE_DURATION, 1); for (size_t d = 0; d < D; ++d) out[d].brake_duration = buf[d];
        rd(RK_FIELD_BRAKE_T, 2);        for (size_t d = 0; d < D; ++d) for (size_t k = 0; k < 2; ++k) out[d].brake_t[k] = buf[k * D + d];
        std::vector<uint32_t> codes(D);
        detail::check(rk_trajectories_read(handle_, dev_, RK_FIELD_CODES, codes.data(), RK_MEM_HOST));
        for (size_t d = 0; d < D; ++d) {
            out[d].limits = (ReachedLimits)(codes[d] & 0xFF);
            out[d].control_signs = (ControlSigns)((codes[d] >> 8) & 0xFF);
            out[d].direction = (Direction)((codes[d] >> 16) & 0xFF);
        }
        return {out};
    }

    // trajectory.rs:211-231
    std::vector<Bound> get_position_extrema() const {
        require_calculated();
        const size_t D = degrees_of_freedom;
        std::vector<double> lo(D), hi(D), tlo(D), thi(D);
        detail::check(rk_position_extrema_batch(handle_, dev_, lo.data(), hi.data(), tlo.data(), thi.data(), RK_MEM_HOST));
        std::vector<Bound> out(D);
        for (size_t d = 0; d < D; ++d) out[d] = Bound{lo[d], hi[d], tlo[d], thi[d]};
        return out;
    }

    // trajectory.rs:233-246
    std::optional<double> get_first_time_at_position(size_t dof, double position) const {
        require_calculated();
        const size_t D = degrees_of_freedom;
        if (dof >= D) return std::nullopt;
        std::vector<double> query(D, position), time(D, 0.0);
        detail::check(rk_first_time_at_position_batch(handle_, dev_, query.data(), time.data(), RK_MEM_HOST));
        if (std::isnan(time[dof])) return std::nullopt;
        return time[dof];
    }

private:
    template <size_t, class> friend class Ruckig;
    void require_calculated() const {
        if (!dev_) throw RuckigError(RuckigError::Kind::Library, "Trajectory: calculate() has not filled this trajectory");
    }
    double duration_ = 0.0;
    Vector independent_min_durations_;
    rk_trajectories* dev_ = nullptr;
    rk_handle* handle_ = nullptr;
};

// --------------------------------------------------------------------------------------------------------- OutputParameter
// output_parameter.rs:36-120
template <size_t DOF>
struct OutputParameter {
    using Vector = std::vector<double>;
    size_t degrees_of_freedom;
    Trajectory<DOF> trajectory;
    Vector new_position, new_velocity, new_acceleration, new_jerk;
    double time = 0.0;
    size_t new_section = 0;
    bool did_section_change = false;
    bool new_calculation = false;
    bool was_calculation_interrupted = false;
    double calculation_duration = 0.0;

    explicit OutputParameter(std::optional<size_t> dofs = std::nullopt)
        : degrees_of_freedom(detail::resolve_dofs(DOF, dofs)), trajectory(dofs), new_position(degrees_of_freedom, 0.0),
          new_velocity(degrees_of_freedom, 0.0), new_acceleration(degrees_of_freedom, 0.0), new_jerk(degrees_of_freedom, 0.0) {}

    void pass_to_input(InputParameter<DOF>& input) const {  // output_parameter.rs:110-120
        input.current_position = new_position;
        input.current_velocity = new_velocity;
        input.current_acceleration = new_acceleration;
    }
};

// ------------------------------------------------------------------------------------------------------------------ Ruckig
// ruckig.rs:76-321 for a single problem; every calculation runs on the GPU as a batch of one.
template <size_t DOF, class E = ThrowErrorHandler>
class Ruckig {
public:
    size_t degrees_of_freedom;
    double delta_time;

    // Ruckig::new(dofs, delta_time), ruckig.rs:110-119
    explicit Ruckig(double delta_time_, std::optional<size_t> dofs = std::nullopt, int32_t device = 0)
        : degrees_of_freedom(detail::resolve_dofs(DOF, dofs)), delta_time(delta_time_), current_input_(dofs),
          handle_(detail::Engine::get(device).h) {}

    void reset() { current_input_initialized_ = false; }  // ruckig.rs:125-127

    // ruckig.rs:150-171: under ThrowErrorHandler an invalid input throws RuckigError(ValidationError); otherwise returns validity
    bool validate_input(const InputParameter<DOF>& input, bool check_current_state_within_limits, bool check_target_state_within_limits) const {
        detail::SingleProblem<DOF> sp(input, delta_time, true);
        uint8_t valid = 0;
        int32_t reason = 0;
        detail::check(rk_validate_batch(handle_, &sp.desc, &sp.in, check_current_state_within_limits ? 1 : 0,
                                        check_target_state_within_limits ? 1 : 0, &valid, &reason));
        if (!valid && E::throws) throw RuckigError(RuckigError::Kind::ValidationError, detail::validation_message(reason));
        return valid != 0;
    }

    // ruckig.rs:210-218
    RuckigResult calculate(const InputParameter<DOF>& input, Trajectory<DOF>& traj) {
        const size_t D = degrees_of_freedom;
        detail::SingleProblem<DOF> sp(input, delta_time, E::throws);
        if (!traj.dev_ || traj.handle_ != handle_) {
            if (traj.dev_) rk_trajectories_destroy(traj.dev_);
            traj.dev_ = nullptr;
            traj.handle_ = handle_;
            detail::check(rk_trajectories_create(handle_, 1, (int32_t)D, &traj.dev_));
        }
        int32_t status = 0;
        double duration = 0.0;
        std::vector<double> imd(D, 0.0);
        rk_batch_out out{&status, &duration, imd.data()};
        detail::check(rk_calculate_batch(handle_, &sp.desc, &sp.in, traj.dev_, &out));
        const RuckigResult code = (RuckigResult)status;
        if (E::throws && code == RuckigResult::ErrorInvalidInput)
            throw RuckigError(RuckigError::Kind::ValidationError, detail::validation_message(error_detail(traj)));
        traj.duration_ = duration;
        traj.independent_min_durations_ = imd;
        if (E::throws && detail::is_calculator_error(code))
            throw RuckigError(RuckigError::Kind::CalculatorError, detail::calculator_message(code, error_detail(traj)));
        return code;
    }

    // ruckig.rs:266-321 (output.time is not reset by a new calculation, as in the reference: SURVEY Appendix A, Q1)
    RuckigResult update(const InputParameter<DOF>& input, OutputParameter<DOF>& output) {
        output.new_calculation = false;
        if (!current_input_initialized_ || input != current_input_) {
            calculate(input, output.trajectory);  // Ok(Error*) codes are dropped (:285); a thrown error leaves the cycle untouched
            output.new_calculation = true;
            current_input_ = input;
            current_input_initialized_ = true;
        }
        output.time += delta_time;
        output.trajectory.at_time(output.time, &output.new_position, &output.new_velocity, &output.new_acceleration, &output.new_jerk, nullptr);
        output.did_section_change = false;  // :300-302: new_section is passed by value
        output.pass_to_input(current_input_);
        return output.time > output.trajectory.get_duration() ? RuckigResult::Finished : RuckigResult::Working;
    }

private:
    int32_t error_detail(const Trajectory<DOF>& traj) const {
        int32_t detail_word = 0;
        detail::check(rk_trajectories_read(handle_, traj.dev_, RK_FIELD_ERROR_DETAIL, &detail_word, RK_MEM_HOST));
        return detail_word;
    }
    InputParameter<DOF> current_input_;
    bool current_input_initialized_ = false;
    rk_handle* handle_;
};

// ------------------------------------------------------------------------------------------------------------- BatchRuckig
// The batched entry point: Ruckig::calculate and Trajectory::at_time for n independent problems in SoA arrays [dof][n]
// (element (d, i) at ptr[d * n + i]), host or device memory. Batch-level settings are those of InputParameter.
struct BatchSettings {
    ControlInterface control_interface = ControlInterface::Position;
    Synchronization synchronization = Synchronization::Time;
    DurationDiscretization duration_discretization = DurationDiscretization::Continuous;
    std::optional<std::vector<ControlInterface>> per_dof_control_interface;
    std::optional<std::vector<Synchronization>> per_dof_synchronization;
    bool device_memory = false;  // rk_batch_in / outputs are device pointers (asynchronous on stream()) instead of host pointers
    bool limits_per_dof = false; // the five limit arrays are [dof], shared by the whole batch (RK_LIMITS_PER_DOF), instead of [dof][n]
};

template <size_t DOF, class E = ThrowErrorHandler>
class BatchRuckig {
public:
    size_t degrees_of_freedom;
    double delta_time;

    // result_layout: RK_RESULT_FULL keeps the reference's Profile per DoF on the device (59 doubles), RK_RESULT_COMPACT the 20
    // doubles the calculation decided; every reader (at_time, read, the query helpers) returns identical values either way
    explicit BatchRuckig(double delta_time_, std::optional<size_t> dofs = std::nullopt, int32_t device = 0, int32_t result_layout = RK_RESULT_FULL)
        : degrees_of_freedom(detail::resolve_dofs(DOF, dofs)), delta_time(delta_time_), handle_(detail::Engine::get(device).h), layout_(result_layout) {}
    ~BatchRuckig() { if (traj_) rk_trajectories_destroy(traj_); }
    BatchRuckig(const BatchRuckig&) = delete;
    BatchRuckig& operator=(const BatchRuckig&) = delete;

    // Ruckig::calculate for trajectories 0..n-1. status[i] is the RuckigResult value the reference returns for problem i
    // (under ThrowErrorHandler -100 / -104 / -110 / -111 are what the reference returns as Err; error_message(i) rebuilds the
    // text). status / duration / independent_min_durations may be null.
    void calculate(int64_t n, const rk_batch_in& in, const BatchSettings& settings = {}, int32_t* status = nullptr, double* duration = nullptr,
                   double* independent_min_durations = nullptr) {
        ensure(n);
        std::vector<int32_t> ci, sy;
        const rk_batch_desc desc = make_desc(n, settings, ci, sy);
        rk_batch_out out{status, duration, independent_min_durations};
        detail::check(rk_calculate_batch(handle_, &desc, &in, traj_, &out));
    }

    // Ruckig::validate_input for every problem: valid[i] = 1 if it passes; reason (nullable) = (RK_INVALID_* << 8) | dof
    void validate_input(int64_t n, const rk_batch_in& in, bool check_current_state_within_limits, bool check_target_state_within_limits,
                        uint8_t* valid, int32_t* reason = nullptr, const BatchSettings& settings = {}) const {
        std::vector<int32_t> ci, sy;
        const rk_batch_desc desc = make_desc(n, settings, ci, sy);
        detail::check(rk_validate_batch(handle_, &desc, &in, check_current_state_within_limits ? 1 : 0, check_target_state_within_limits ? 1 : 0, valid, reason));
    }

    // Trajectory::at_time for `steps` control cycles of every trajectory: outputs [steps][dof][n], each nullable
    void at_time(double time_start, double cycle_time, int32_t steps, bool device_memory, double* position, double* velocity = nullptr,
                 double* acceleration = nullptr, double* jerk = nullptr, int32_t* section = nullptr) const {
        if (!traj_) throw RuckigError(RuckigError::Kind::Library, "BatchRuckig: calculate() has not been called");
        rk_sample_desc sd{time_start, cycle_time, steps, device_memory ? RK_MEM_DEVICE : RK_MEM_HOST};
        rk_sample_out so{position, velocity, acceleration, jerk, section};
        detail::check(rk_at_time_batch(handle_, traj_, &sd, &so));
    }

    // the text the reference would have put into RuckigError for trajectory i of the last batch ("" if it has none)
    std::string error_message(int64_t i) const {
        if (!traj_ || i < 0 || i >= n_) return "";
        std::vector<int32_t> status((size_t)n_), word((size_t)n_);
        detail::check(rk_trajectories_read(handle_, traj_, RK_FIELD_STATUS, status.data(), RK_MEM_HOST));
        detail::check(rk_trajectories_read(handle_, traj_, RK_FIELD_ERROR_DETAIL, word.data(), RK_MEM_HOST));
        const RuckigResult code = (RuckigResult)status[(size_t)i];
        if (code == RuckigResult::ErrorInvalidInput) return detail::validation_message(word[(size_t)i]);
        if (detail::is_calculator_error(code)) return detail::calculator_message(code, word[(size_t)i]);
        return "";
    }

    void read(int32_t field, void* dst, bool device_memory = false) const {
        if (!traj_) throw RuckigError(RuckigError::Kind::Library, "BatchRuckig: calculate() has not been called");
        detail::check(rk_trajectories_read(handle_, traj_, field, dst, device_memory ? RK_MEM_DEVICE : RK_MEM_HOST));
    }
    void sync() const { detail::check(rk_sync(handle_)); }
    void* stream() const { return rk_stream(handle_); }
    rk_handle* handle() const { return handle_; }
    rk_trajectories* trajectories() const { return traj_; }

private:
    void ensure(int64_t n) {
        if (traj_ && n_ == n) return;
        if (traj_) rk_trajectories_destroy(traj_);
        traj_ = nullptr;
        detail::check(rk_trajectories_create_ex(handle_, n, (int32_t)degrees_of_freedom, layout_, &traj_));
        n_ = n;
    }
    rk_batch_desc make_desc(int64_t n, const BatchSettings& s, std::vector<int32_t>& ci, std::vector<int32_t>& sy) const {
        rk_batch_desc d{};
        d.n = n;
        d.dof = (int32_t)degrees_of_freedom;
        d.control_interface = (int32_t)s.control_interface;
        d.synchronization = (int32_t)s.synchronization;
        d.duration_discretization = (int32_t)s.duration_discretization;
        d.error_mode = E::throws ? RK_ERRORS_THROW : RK_ERRORS_IGNORE;
        d.memory_space = s.device_memory ? RK_MEM_DEVICE : RK_MEM_HOST;
        d.limits_layout = s.limits_per_dof ? RK_LIMITS_PER_DOF : RK_LIMITS_PER_TRAJECTORY;
        d.delta_time = delta_time;
        if (s.per_dof_control_interface) {
            for (ControlInterface c : *s.per_dof_control_interface) ci.push_back((int32_t)c);
            d.per_dof_control_interface = ci.data();
        }
        if (s.per_dof_synchronization) {
            for (Synchronization c : *s.per_dof_synchronization) sy.push_back((int32_t)c);
            d.per_dof_synchronization = sy.data();
        }
        return d;
    }
    rk_handle* handle_;
    int32_t layout_ = RK_RESULT_FULL;
    rk_trajectories* traj_ = nullptr;
    int64_t n_ = 0;
};

// ------------------------------------------------------------------------------------------------------------ MultiBatchRuckig
// One HOST batch over several devices (rk_calculate_batch_multi): part g = the contiguous index range rk_shard_range(n, G, g) on
// its own handle, all parts concurrently (one host thread each inside the library), status / duration / independent minimum
// durations gathered into the caller's arrays. `devices` may name a device more than once. north_star: "sharded trivially
// across the 8 GPUs of one box with no collectives beyond a final host gather".
template <size_t DOF, class E = ThrowErrorHandler>
class MultiBatchRuckig {
public:
    size_t degrees_of_freedom;
    double delta_time;

    MultiBatchRuckig(double delta_time_, const std::vector<int32_t>& devices, std::optional<size_t> dofs = std::nullopt, int32_t result_layout = RK_RESULT_COMPACT)
        : degrees_of_freedom(detail::resolve_dofs(DOF, dofs)), delta_time(delta_time_), layout_(result_layout) {
        for (int32_t d : devices) {
            rk_handle* h = nullptr;
            detail::check(rk_create(d, &h));
            handles_.push_back(h);
        }
    }
    ~MultiBatchRuckig() {
        for (rk_trajectories* t : trajs_) rk_trajectories_destroy(t);
        for (rk_handle* h : handles_) rk_destroy(h);
    }
    MultiBatchRuckig(const MultiBatchRuckig&) = delete;
    MultiBatchRuckig& operator=(const MultiBatchRuckig&) = delete;

    size_t parts() const { return handles_.size(); }
    rk_handle* handle(size_t g) const { return handles_[g]; }
    rk_trajectories* trajectories(size_t g) const { return trajs_[g]; }  // part g holds trajectories rk_shard_range(n, parts(), g)

    void calculate(int64_t n, const rk_batch_in& in, BatchSettings settings = {}, int32_t* status = nullptr, double* duration = nullptr,
                   double* independent_min_durations = nullptr) {
        if (n != n_) {
            for (rk_trajectories* t : trajs_) rk_trajectories_destroy(t);
            trajs_.clear();
            for (size_t g = 0; g < handles_.size(); ++g) {
                int64_t first = 0, count = 0;
                rk_shard_range(n, (int32_t)handles_.size(), (int32_t)g, &first, &count);
                rk_trajectories* t = nullptr;
                detail::check(rk_trajectories_create_ex(handles_[g], count, (int32_t)degrees_of_freedom, layout_, &t));
                trajs_.push_back(t);
            }
            n_ = n;
        }
        rk_batch_desc d{};
        d.n = n;
        d.dof = (int32_t)degrees_of_freedom;
        d.control_interface = (int32_t)settings.control_interface;
        d.synchronization = (int32_t)settings.synchronization;
        d.duration_discretization = (int32_t)settings.duration_discretization;
        d.error_mode = E::throws ? RK_ERRORS_THROW : RK_ERRORS_IGNORE;
        d.memory_space = RK_MEM_HOST;
        d.delta_time = delta_time;
        std::vector<int32_t> ci, sy;
        if (settings.per_dof_control_interface) {
            for (ControlInterface c : *settings.per_dof_control_interface) ci.push_back((int32_t)c);
            d.per_dof_control_interface = ci.data();
        }
        if (settings.per_dof_synchronization) {
            for (Synchronization c : *settings.per_dof_synchronization) sy.push_back((int32_t)c);
            d.per_dof_synchronization = sy.data();
        }
        rk_batch_out out{status, duration, independent_min_durations};
        detail::check(rk_calculate_batch_multi(handles_.data(), trajs_.data(), (int32_t)handles_.size(), &d, &in, &out));
    }

private:
    int32_t layout_;
    std::vector<rk_handle*> handles_;
    std::vector<rk_trajectories*> trajs_;
    int64_t n_ = -1;
};

// ------------------------------------------------------------------------------------------------------------ BatchRollout
// n independent closed control loops on the device: the batched Ruckig::update (ruckig.rs:266-321). update() is one control
// cycle for every environment: those whose input changed (or that were never calculated) are re-solved, all advance their time
// by delta_time, sample their trajectory and — with pass_to_input — get the new state written into in.current_*, ready for the
// next cycle (OutputParameter::pass_to_input). All buffers are DEVICE arrays [dof][n] / [n] (rk_update_batch takes no host
// buffers). Same calls as rsruckig_b200.api.BatchRollout, which tests/test_rollout.py holds to one stateful oracle calculator
// per environment.
template <size_t DOF, class E = ThrowErrorHandler>
class BatchRollout {
public:
    int64_t n;
    size_t degrees_of_freedom;
    double delta_time;

    BatchRollout(int64_t n_, double delta_time_, std::optional<size_t> dofs = std::nullopt, int32_t device = 0)
        : n(n_), degrees_of_freedom(detail::resolve_dofs(DOF, dofs)), delta_time(delta_time_), handle_(detail::Engine::get(device).h) {
        detail::check(rk_rollout_create(handle_, n, (int32_t)degrees_of_freedom, &rollout_));
    }
    ~BatchRollout() { if (rollout_) rk_rollout_destroy(rollout_); }
    BatchRollout(const BatchRollout&) = delete;
    BatchRollout& operator=(const BatchRollout&) = delete;

    void reset() { detail::check(rk_rollout_reset(handle_, rollout_)); }  // Ruckig::reset + a fresh OutputParameter per environment

    // One control cycle. `out` names the device arrays that receive OutputParameter's fields (each nullable) and carries the
    // pass_to_input / reset_time_on_new_calculation switches. Returns the number of environments that were re-solved.
    int64_t update(const rk_batch_in& in, const rk_update_out& out, BatchSettings settings = {}) {
        settings.device_memory = true;
        rk_batch_desc d{};
        d.n = n;
        d.dof = (int32_t)degrees_of_freedom;
        d.control_interface = (int32_t)settings.control_interface;
        d.synchronization = (int32_t)settings.synchronization;
        d.duration_discretization = (int32_t)settings.duration_discretization;
        d.error_mode = E::throws ? RK_ERRORS_THROW : RK_ERRORS_IGNORE;
        d.memory_space = RK_MEM_DEVICE;
        d.delta_time = delta_time;
        std::vector<int32_t> ci, sy;
        if (settings.per_dof_control_interface) {
            for (ControlInterface c : *settings.per_dof_control_interface) ci.push_back((int32_t)c);
            d.per_dof_control_interface = ci.data();
        }
        if (settings.per_dof_synchronization) {
            for (Synchronization c : *settings.per_dof_synchronization) sy.push_back((int32_t)c);
            d.per_dof_synchronization = sy.data();
        }
        int64_t recalculated = 0;
        detail::check(rk_update_batch(handle_, &d, &in, rollout_, &out, &recalculated));
        return recalculated;
    }

    rk_trajectories* trajectories() const { return rk_rollout_trajectories(rollout_); }  // owned by the rollout
    void sync() const { detail::check(rk_sync(handle_)); }
    void* stream() const { return rk_stream(handle_); }

private:
    rk_handle* handle_;
    rk_rollout* rollout_ = nullptr;
};

}  // namespace ruckig_b200
