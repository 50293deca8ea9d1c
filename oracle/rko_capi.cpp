// rko_capi.cpp — C entry points of the parity oracle (liboracle_*.so).
//
// TEST INFRASTRUCTURE ONLY (see rko_math.hpp): loaded by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
// legs. The batch functions take the same rk_batch_desc / rk_batch_in structs
// as the product's C ABI (include/ruckig_b200.h) so one set of buffers feeds
// both sides.
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "../include/ruckig_b200.h"
#include "rko_calc.hpp"

using namespace rko;

namespace {

struct Batch {
    int64_t n = 0;
    int dof = 0;
    std::vector<Trajectory> traj;
    std::vector<int32_t> status, limiting;
    std::vector<uint8_t> source;  // [dof][n]
};

void fill_input(InputParameter& ip, const rk_batch_desc* d, const rk_batch_in* in, int64_t i) {
    const int D = d->dof;
    const int64_t n = d->n;
    ip.control_interface = d->control_interface;
    ip.synchronization = d->synchronization;
    ip.duration_discretization = d->duration_discretization;
    ip.has_per_dof_control_interface = d->per_dof_control_interface != nullptr;
    ip.has_per_dof_synchronization = d->per_dof_synchronization != nullptr;
    ip.has_min_velocity = in->min_velocity != nullptr;
    ip.has_min_acceleration = in->min_acceleration != nullptr;
    for (int k = 0; k < D; ++k) {
        const int64_t o = (int64_t)k * n + i;
        ip.current_position[k] = in->current_position[o];
        ip.current_velocity[k] = in->current_velocity[o];
        ip.current_acceleration[k] = in->current_acceleration[o];
        ip.target_position[k] = in->target_position[o];
        ip.target_velocity[k] = in->target_velocity[o];
        ip.target_acceleration[k] = in->target_acceleration[o];
        ip.max_velocity[k] = in->max_velocity[o];
        ip.max_acceleration[k] = in->max_acceleration[o];
        ip.max_jerk[k] = in->max_jerk[o];
        if (ip.has_min_velocity) ip.min_velocity[k] = in->min_velocity[o];
        if (ip.has_min_acceleration) ip.min_acceleration[k] = in->min_acceleration[o];
        ip.enabled[k] = in->enabled ? in->enabled[o] : 1;
        if (ip.has_per_dof_control_interface) ip.per_dof_control_interface[k] = d->per_dof_control_interface[k];
        if (ip.has_per_dof_synchronization) ip.per_dof_synchronization[k] = d->per_dof_synchronization[k];
    }
    ip.has_minimum_duration = false;
    if (in->minimum_duration) {
        const double md = in->minimum_duration[i];
        if (md == md) { ip.has_minimum_duration = true; ip.minimum_duration = md; }
    }
}

}  // namespace

extern "C" {

const char* rko_math_flavour() {
#if defined(RKO_MATH_PORTABLE)
    return "portable";
#else
    return "glibc";
#endif
}

void* rko_batch_new(int64_t n, int32_t dof) {
    Batch* b = new Batch();
    b->n = n; b->dof = dof;
    b->traj.assign(n, Trajectory(dof));
    b->status.assign(n, 0);
    b->limiting.assign(n, -1);
    b->source.assign((size_t)n * dof, 5);
    return b;
}
void rko_batch_free(void* p) { delete (Batch*)p; }

// Ruckig::calculate on a FRESH Ruckig + Trajectory per trajectory (the batched
// semantics: no state carried from one problem to the next). Returns seconds.
double rko_batch_calculate(void* pb, const rk_batch_desc* d, const rk_batch_in* in, int32_t threads) {
    Batch* b = (Batch*)pb;
    const int64_t n = d->n;
    const int D = d->dof;
    if (threads < 1) threads = 1;
    const bool throw_mode = d->error_mode == RK_ERRORS_THROW;
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int64_t lo, int64_t hi) {
        InputParameter ip(D);
        for (int64_t i = lo; i < hi; ++i) {
            fill_input(ip, d, in, i);
            Ruckig otg(D, d->delta_time, throw_mode);
            Trajectory& tr = b->traj[i];
            tr.resize(D);
            Outcome o = otg.calculate(ip, tr);
            b->status[i] = o.code;
            b->limiting[i] = otg.calculator.last_limiting_dof;
            for (int k = 0; k < D; ++k) b->source[(size_t)k * n + i] = (uint8_t)otg.calculator.profile_source[k];
        }
    };
    if (threads == 1) {
        work(0, n);
    } else {
        std::vector<std::thread> th;
        const int64_t per = (n + threads - 1) / threads;
        for (int t = 0; t < threads; ++t) {
            const int64_t lo = t * per, hi = std::min<int64_t>(n, lo + per);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& x : th) x.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Throughput-only variant: one reused Ruckig + Trajectory per thread, as the
// reference bench does (bench/src/benchmark_target.rs:122-163); keeps only
// status and duration. Returns seconds spent inside calculate().
double rko_bench_calculate(const rk_batch_desc* d, const rk_batch_in* in, int32_t threads, int32_t* status, double* duration) {
    const int64_t n = d->n;
    const int D = d->dof;
    if (threads < 1) threads = 1;
    const bool throw_mode = d->error_mode == RK_ERRORS_THROW;
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int64_t lo, int64_t hi) {
        InputParameter ip(D);
        Ruckig otg(D, d->delta_time, throw_mode);
        for (int64_t i = lo; i < hi; ++i) {
            fill_input(ip, d, in, i);
            Trajectory tr(D);
            Outcome o = otg.calculate(ip, tr);
            if (status) status[i] = o.code;
            if (duration) duration[i] = tr.duration;
        }
    };
    if (threads == 1) {
        work(0, n);
    } else {
        std::vector<std::thread> th;
        const int64_t per = (n + threads - 1) / threads;
        for (int t = 0; t < threads; ++t) {
            const int64_t lo = t * per, hi = std::min<int64_t>(n, lo + per);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& x : th) x.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// valid[i] = 1 when InputParameter::validate(check_current, check_target) and the
// delta_time rule of Ruckig::validate_input pass (ruckig.rs:150-171).
void rko_batch_validate(const rk_batch_desc* d, const rk_batch_in* in, int32_t check_current, int32_t check_target, uint8_t* valid) {
    InputParameter ip(d->dof);
    for (int64_t i = 0; i < d->n; ++i) {
        fill_input(ip, d, in, i);
        bool ok = ip.validate(check_current != 0, check_target != 0).empty();
        if (d->delta_time <= 0.0 && d->duration_discretization != DD_CONTINUOUS) ok = false;
        valid[i] = ok ? 1 : 0;
    }
}

// As rko_batch_validate, and the text of the first failing check per trajectory ("" when valid or when only the delta_time
// rule fails) into msgs[i * msg_stride ...] (NUL-terminated, truncated to the stride).
void rko_batch_validate_messages(const rk_batch_desc* d, const rk_batch_in* in, int32_t check_current, int32_t check_target, uint8_t* valid,
                                 char* msgs, int32_t msg_stride) {
    InputParameter ip(d->dof);
    for (int64_t i = 0; i < d->n; ++i) {
        fill_input(ip, d, in, i);
        const std::string m = ip.validate(check_current != 0, check_target != 0);
        bool ok = m.empty();
        if (d->delta_time <= 0.0 && d->duration_discretization != DD_CONTINUOUS) ok = false;
        valid[i] = ok ? 1 : 0;
        char* o = msgs + i * (int64_t)msg_stride;
        const size_t k = m.size() < (size_t)msg_stride - 1 ? m.size() : (size_t)msg_stride - 1;
        std::memcpy(o, m.data(), k);
        o[k] = 0;
    }
}

int32_t rko_batch_read(void* pb, int32_t field, void* dst) {
    Batch* b = (Batch*)pb;
    const int64_t n = b->n;
    const int D = b->dof;
    auto put = [&](int slots, auto getter) {
        double* o = (double*)dst;
        for (int s = 0; s < slots; ++s)
            for (int k = 0; k < D; ++k)
                for (int64_t i = 0; i < n; ++i) o[((int64_t)s * D + k) * n + i] = getter(b->traj[i].profiles[k], s);
    };
    switch (field) {
        case RK_FIELD_STATUS: std::memcpy(dst, b->status.data(), n * sizeof(int32_t)); return 0;
        case RK_FIELD_DURATION: for (int64_t i = 0; i < n; ++i) ((double*)dst)[i] = b->traj[i].duration; return 0;
        case RK_FIELD_INDEPENDENT_MIN_DURATIONS:
            for (int k = 0; k < D; ++k) for (int64_t i = 0; i < n; ++i) ((double*)dst)[(int64_t)k * n + i] = b->traj[i].independent_min_durations[k];
            return 0;
        case RK_FIELD_LIMITING_DOF: std::memcpy(dst, b->limiting.data(), n * sizeof(int32_t)); return 0;
        case RK_FIELD_CODES:
            for (int k = 0; k < D; ++k) for (int64_t i = 0; i < n; ++i) {
                const Profile& p = b->traj[i].profiles[k];
                ((uint32_t*)dst)[(int64_t)k * n + i] = (uint32_t)p.limits | ((uint32_t)p.control_signs << 8) | ((uint32_t)p.direction << 16) | ((uint32_t)b->source[(size_t)k * n + i] << 24);
            }
            return 0;
        case RK_FIELD_T: put(7, [](const Profile& p, int s) { return p.t[s]; }); return 0;
        case RK_FIELD_T_SUM: put(7, [](const Profile& p, int s) { return p.t_sum[s]; }); return 0;
        case RK_FIELD_J: put(7, [](const Profile& p, int s) { return p.j[s]; }); return 0;
        case RK_FIELD_A: put(8, [](const Profile& p, int s) { return p.a[s]; }); return 0;
        case RK_FIELD_V: put(8, [](const Profile& p, int s) { return p.v[s]; }); return 0;
        case RK_FIELD_P: put(8, [](const Profile& p, int s) { return p.p[s]; }); return 0;
        case RK_FIELD_BRAKE_DURATION: put(1, [](const Profile& p, int) { return p.brake.duration; }); return 0;
        case RK_FIELD_BRAKE_T: put(2, [](const Profile& p, int s) { return p.brake.t[s]; }); return 0;
        case RK_FIELD_BRAKE_J: put(2, [](const Profile& p, int s) { return p.brake.j[s]; }); return 0;
        case RK_FIELD_BRAKE_A: put(2, [](const Profile& p, int s) { return p.brake.a[s]; }); return 0;
        case RK_FIELD_BRAKE_V: put(2, [](const Profile& p, int s) { return p.brake.v[s]; }); return 0;
        case RK_FIELD_BRAKE_P: put(2, [](const Profile& p, int s) { return p.brake.p[s]; }); return 0;
        case RK_FIELD_PF_VF_AF: put(3, [](const Profile& p, int s) { return s == 0 ? p.pf : (s == 1 ? p.vf : p.af); }); return 0;
        default: return -1;
    }
}

// Trajectory::at_time for K samples per trajectory; sample times accumulate by
// repeated addition like Ruckig::update does (ruckig.rs:293). Returns seconds.
// Trajectory::get_position_extrema (trajectory.rs:211-231, one section): out = [4][dof][n] = min, max, t_min, t_max
void rko_batch_position_extrema(void* pb, double* out) {
    Batch* b = (Batch*)pb;
    const int64_t n = b->n, stride = (int64_t)b->dof * n;
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < b->dof; ++d) {
            const PositionExtrema e = get_position_extrema(b->traj[i].profiles[d]);
            const int64_t o = (int64_t)d * n + i;
            out[o] = e.min; out[stride + o] = e.max; out[2 * stride + o] = e.t_min; out[3 * stride + o] = e.t_max;
        }
}

// Trajectory::get_first_time_at_position (trajectory.rs:233-246) for one query position per (dof, trajectory); NaN = None
void rko_batch_first_time_at_position(void* pb, const double* positions, double* times) {
    Batch* b = (Batch*)pb;
    const int64_t n = b->n;
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < b->dof; ++d) {
            const int64_t o = (int64_t)d * n + i;
            double t, v, a;
            times[o] = get_first_state_at_position(b->traj[i].profiles[d], positions[o], 0.0, t, v, a) ? t : std::nan("");
        }
}

double rko_batch_at_time(void* pb, const rk_sample_desc* sd, const rk_sample_out* out, int32_t threads) {
    Batch* b = (Batch*)pb;
    const int64_t n = b->n;
    const int D = b->dof;
    const int K = sd->steps;
    if (threads < 1) threads = 1;
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int64_t lo, int64_t hi) {
        std::vector<double> p(D), v(D), a(D), j(D);
        for (int64_t i = lo; i < hi; ++i) {
            double time = sd->time_start;
            for (int k = 0; k < K; ++k) {
                time += sd->delta_time;
                int section = 0;
                b->traj[i].at_time(time, p.data(), v.data(), a.data(), j.data(), &section);
                for (int dd = 0; dd < D; ++dd) {
                    const int64_t o = ((int64_t)k * D + dd) * n + i;
                    if (out->position) out->position[o] = p[dd];
                    if (out->velocity) out->velocity[o] = v[dd];
                    if (out->acceleration) out->acceleration[o] = a[dd];
                    if (out->jerk) out->jerk[o] = j[dd];
                }
                if (out->section) out->section[(int64_t)k * n + i] = section;
            }
        }
    };
    if (threads == 1) {
        work(0, n);
    } else {
        std::vector<std::thread> th;
        const int64_t per = (n + threads - 1) / threads;
        for (int t = 0; t < threads; ++t) {
            const int64_t lo = t * per, hi = std::min<int64_t>(n, lo + per);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& x : th) x.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- stateful single-problem API: replays the reference's own tests ---------
struct Otg {
    Ruckig otg;
    InputParameter input;
    OutputParameter output;
    Outcome last;
    Otg(int dof, double dt, bool throw_mode) : otg(dof, dt, throw_mode), input(dof), output(dof) {}
};

void* rko_otg_new(int32_t dof, double delta_time, int32_t throw_mode) { return new Otg(dof, delta_time, throw_mode != 0); }
void rko_otg_free(void* p) { delete (Otg*)p; }
void rko_otg_reset(void* p) { ((Otg*)p)->otg.reset(); }
void rko_otg_new_output(void* p) { Otg* o = (Otg*)p; o->output = OutputParameter(o->otg.degrees_of_freedom); }

// Input fields are passed as a flat struct of n = 1 SoA pointers.
void rko_otg_set_input(void* p, const rk_batch_desc* d, const rk_batch_in* in) {
    Otg* o = (Otg*)p;
    fill_input(o->input, d, in, 0);
}
static int32_t finish(Otg* o, const Outcome& r) { o->last = r; return r.is_err() ? (r.err == ERR_VALIDATION ? -1000 : -2000) + 0 : r.code; }
// returns the RuckigResult code, or -1000 (Err(ValidationError)) / -2000 (Err(CalculatorError))
int32_t rko_otg_update(void* p) { Otg* o = (Otg*)p; return finish(o, o->otg.update(o->input, o->output)); }
int32_t rko_otg_calculate(void* p) { Otg* o = (Otg*)p; return finish(o, o->otg.calculate(o->input, o->output.trajectory)); }
int32_t rko_otg_validate(void* p, int32_t cc, int32_t ct) { Otg* o = (Otg*)p; return finish(o, o->otg.validate_input(o->input, cc != 0, ct != 0)); }
const char* rko_otg_last_message(void* p) { return ((Otg*)p)->last.msg.c_str(); }
int32_t rko_otg_last_code(void* p) { return ((Otg*)p)->last.code; }
double rko_otg_duration(void* p) { return ((Otg*)p)->output.trajectory.duration; }
double rko_otg_time(void* p) { return ((Otg*)p)->output.time; }
int32_t rko_otg_new_calculation(void* p) { return ((Otg*)p)->output.new_calculation ? 1 : 0; }
// which: 0 new_position, 1 new_velocity, 2 new_acceleration, 3 new_jerk, 4 independent_min_durations
void rko_otg_get_output(void* p, int32_t which, double* dst) {
    Otg* o = (Otg*)p;
    const std::vector<double>* v = nullptr;
    switch (which) {
        case 0: v = &o->output.new_position; break;
        case 1: v = &o->output.new_velocity; break;
        case 2: v = &o->output.new_acceleration; break;
        case 3: v = &o->output.new_jerk; break;
        default: v = &o->output.trajectory.independent_min_durations; break;
    }
    for (size_t i = 0; i < v->size(); ++i) dst[i] = (*v)[i];
}
void rko_otg_at_time(void* p, double time, double* pos, double* vel, double* acc, double* jerk, int32_t* section) {
    Otg* o = (Otg*)p;
    int s = section ? *section : 0;
    o->output.trajectory.at_time(time, pos, vel, acc, jerk, &s);
    if (section) *section = s;
}
// profile field of one DoF: field ids as RK_FIELD_T .. RK_FIELD_PF_VF_AF; returns slot count
int32_t rko_otg_profile(void* p, int32_t dof, int32_t field, double* dst) {
    const Profile& q = ((Otg*)p)->output.trajectory.profiles[dof];
    switch (field) {
        case RK_FIELD_T: for (int i = 0; i < 7; ++i) dst[i] = q.t[i]; return 7;
        case RK_FIELD_T_SUM: for (int i = 0; i < 7; ++i) dst[i] = q.t_sum[i]; return 7;
        case RK_FIELD_J: for (int i = 0; i < 7; ++i) dst[i] = q.j[i]; return 7;
        case RK_FIELD_A: for (int i = 0; i < 8; ++i) dst[i] = q.a[i]; return 8;
        case RK_FIELD_V: for (int i = 0; i < 8; ++i) dst[i] = q.v[i]; return 8;
        case RK_FIELD_P: for (int i = 0; i < 8; ++i) dst[i] = q.p[i]; return 8;
        case RK_FIELD_BRAKE_DURATION: dst[0] = q.brake.duration; return 1;
        case RK_FIELD_BRAKE_T: dst[0] = q.brake.t[0]; dst[1] = q.brake.t[1]; return 2;
        case RK_FIELD_CODES: dst[0] = q.limits; dst[1] = q.control_signs; dst[2] = q.direction; return 3;
        default: return 0;
    }
}

}  // extern "C"
