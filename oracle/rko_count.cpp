// rko_count.cpp — op-counting build of the parity oracle (liboracle_count.so).
//
// TEST / MEASUREMENT INFRASTRUCTURE ONLY (see rko_math.hpp). SURVEY.md §8(d)
// defines the algorithmic work of one trajectory as the FP64 operations the
// REFERENCE ALGORITHM executes for that input: every add / sub / mul / div /
// sqrt counts 1, every libm call (cbrt, sin, cos, acos, atan2, fmod) counts 1,
// compares, negations, abs, min/max and copies count 0, FMA is not assumed.
// This file compiles the oracle's restatement of rsruckig (rko_calc.hpp and
// what it includes, unchanged) with `double` replaced by a counting scalar, so
// the figure is a property of (algorithm, input distribution) and does not
// depend on how the CUDA pipeline schedules or recomputes anything.
//
// How: all system headers, the C ABI structs and rko_math.hpp are included
// first with the real `double`; then `#define double cnt` and the oracle
// headers are included. `cnt` converts implicitly FROM double and only
// explicitly TO double, so mixed expressions and `c ? x : 0.0` resolve to cnt.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#define RKO_PHASE(id) rkc::Scope rkc_scope_(id)
#include "../include/ruckig_b200.h"
#include "rko_math.hpp"  // glibc flavour: the libm calls stay opaque (1 op each)

namespace rkc {
enum { ADD = 0, MUL, DIV, SQRT, LIBM, NOPS };
static uint64_t g_ops[RKO_PH_COUNT][NOPS];  // single-threaded by construction (rko_count_calculate)
static int g_phase = RKO_PH_OTHER;
static constexpr void bump(int k) {
    if (!__builtin_is_constant_evaluated()) ++g_ops[g_phase][k];
}
// RKO_PHASE(id) at the top of a function of the oracle: operations until the end of that scope belong to phase `id`
// (the innermost marker wins: a profile check called from a Step-1 family counts as check, not as that family).
struct Scope {
    int saved;
    explicit Scope(int id) : saved(g_phase) { g_phase = id; }
    ~Scope() { g_phase = saved; }
};
}  // namespace rkc

struct cnt {
    double v;
    cnt() = default;
    constexpr cnt(double x) : v(x) {}
    constexpr cnt(float x) : v(x) {}
    constexpr cnt(long double x) : v((double)x) {}
    constexpr cnt(int x) : v(x) {}
    constexpr cnt(long x) : v((double)x) {}
    constexpr cnt(long long x) : v((double)x) {}
    constexpr cnt(unsigned x) : v(x) {}
    constexpr cnt(unsigned long x) : v((double)x) {}
    constexpr explicit operator double() const { return v; }
    constexpr explicit operator int() const { return (int)v; }
    constexpr explicit operator long() const { return (long)v; }
    constexpr explicit operator bool() const { return v != 0.0; }
    constexpr cnt operator-() const { return cnt(-v); }
    constexpr cnt operator+() const { return *this; }
    constexpr cnt& operator+=(cnt o) { rkc::bump(rkc::ADD); v += o.v; return *this; }
    constexpr cnt& operator-=(cnt o) { rkc::bump(rkc::ADD); v -= o.v; return *this; }
    constexpr cnt& operator*=(cnt o) { rkc::bump(rkc::MUL); v *= o.v; return *this; }
    constexpr cnt& operator/=(cnt o) { rkc::bump(rkc::DIV); v /= o.v; return *this; }
};
static_assert(sizeof(cnt) == sizeof(double), "cnt must alias double");

#define RKC_BIN(op, kind)                                                                                   \
    constexpr cnt operator op(cnt a, cnt b) { rkc::bump(rkc::kind); return cnt(a.v op b.v); }             \
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>                             \
    constexpr cnt operator op(cnt a, T b) { rkc::bump(rkc::kind); return cnt(a.v op (double)b); }         \
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>                             \
    constexpr cnt operator op(T a, cnt b) { rkc::bump(rkc::kind); return cnt((double)a op b.v); }
RKC_BIN(+, ADD)
RKC_BIN(-, ADD)
RKC_BIN(*, MUL)
RKC_BIN(/, DIV)
#undef RKC_BIN
#define RKC_CMP(op)                                                                                         \
    constexpr bool operator op(cnt a, cnt b) { return a.v op b.v; }                                        \
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>                             \
    constexpr bool operator op(cnt a, T b) { return a.v op (double)b; }                                    \
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>                             \
    constexpr bool operator op(T a, cnt b) { return (double)a op b.v; }
RKC_CMP(<)
RKC_CMP(>)
RKC_CMP(<=)
RKC_CMP(>=)
RKC_CMP(==)
RKC_CMP(!=)
#undef RKC_CMP

namespace std {
inline cnt fabs(cnt x) { return cnt(::fabs(x.v)); }
inline cnt abs(cnt x) { return cnt(::fabs(x.v)); }
inline cnt sqrt(cnt x) { rkc::bump(rkc::SQRT); return cnt(::sqrt(x.v)); }
inline cnt fmod(cnt a, cnt b) { rkc::bump(rkc::LIBM); return cnt(::fmod(a.v, b.v)); }
inline bool isinf(cnt x) { return std::isinf(x.v); }
inline bool isnan(cnt x) { return std::isnan(x.v); }
inline std::string to_string(cnt x) { return std::to_string(x.v); }
template <> struct numeric_limits<cnt> {
    static constexpr bool is_specialized = true;
    static constexpr cnt infinity() { return cnt(numeric_limits<double>::infinity()); }
    static constexpr cnt quiet_NaN() { return cnt(numeric_limits<double>::quiet_NaN()); }
    static constexpr cnt epsilon() { return cnt(numeric_limits<double>::epsilon()); }
    static constexpr cnt max() { return cnt(numeric_limits<double>::max()); }
    static constexpr cnt min() { return cnt(numeric_limits<double>::min()); }
    static constexpr cnt lowest() { return cnt(numeric_limits<double>::lowest()); }
};
}  // namespace std

namespace rko {
static inline cnt m_cbrt(cnt x) { rkc::bump(rkc::LIBM); return cnt(::cbrt(x.v)); }
static inline cnt m_sin(cnt x) { rkc::bump(rkc::LIBM); return cnt(::sin(x.v)); }
static inline cnt m_cos(cnt x) { rkc::bump(rkc::LIBM); return cnt(::cos(x.v)); }
static inline cnt m_acos(cnt x) { rkc::bump(rkc::LIBM); return cnt(::acos(x.v)); }
static inline cnt m_atan2(cnt y, cnt x) { rkc::bump(rkc::LIBM); return cnt(::atan2(y.v, x.v)); }
static inline cnt fmin_(cnt a, cnt b) { return (a != a) ? b : ((b != b) ? a : (b < a ? b : a)); }
static inline cnt fmax_(cnt a, cnt b) { return (a != a) ? b : ((b != b) ? a : (b > a ? b : a)); }
}  // namespace rko

#define double cnt
#include "rko_calc.hpp"
#undef double

extern "C" {

int32_t rko_count_phases() { return RKO_PH_COUNT; }

// Ruckig::calculate for every problem of the batch on the calling thread (fresh
// Ruckig + Trajectory each, as rko_batch_calculate does). ops[RKO_PH_COUNT][5]
// receives the totals {add+sub, mul, div, sqrt, libm} per phase (rko_math.hpp);
// status / duration (nullable) let the caller check that the counting build
// takes the same branches as the plain one.
void rko_count_calculate(const rk_batch_desc* d, const rk_batch_in* in, uint64_t* ops, int32_t* status, double* duration) {
    using namespace rko;
    const int64_t n = d->n;
    const int D = d->dof;
    const bool throw_mode = d->error_mode == RK_ERRORS_THROW;
    uint64_t acc[RKO_PH_COUNT][rkc::NOPS] = {{0}};
    InputParameter ip(D);
    for (int64_t i = 0; i < n; ++i) {
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
        std::memset(rkc::g_ops, 0, sizeof(rkc::g_ops));  // filling the input above is data movement, not algorithm
        rkc::g_phase = RKO_PH_OTHER;
        Ruckig otg(D, d->delta_time, throw_mode);
        Trajectory tr(D);
        Outcome o = otg.calculate(ip, tr);
        for (int p = 0; p < RKO_PH_COUNT; ++p)
            for (int k = 0; k < rkc::NOPS; ++k) acc[p][k] += rkc::g_ops[p][k];
        if (status) status[i] = o.code;
        if (duration) duration[i] = (double)tr.duration;
    }
    for (int p = 0; p < RKO_PH_COUNT; ++p)
        for (int k = 0; k < rkc::NOPS; ++k) ops[p * rkc::NOPS + k] = acc[p][k];
}

}  // extern "C"
