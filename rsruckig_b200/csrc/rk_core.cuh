// rk_core.cuh — device numeric core: candidate validation (the reference's
// Profile::check family), brake pre-trajectories, closed-form root solvers.
//
// A candidate profile on the device is just its seven phase durations t[7], a
// control value (jerk jf / acceleration pair / velocity) and two small codes;
// the p/v/a arrays of the reference's 568-byte Profile are a deterministic
// function of those and of the start state, so they are integrated in
// registers when a candidate is validated and once more when the winning
// profile is written to HBM (rk_store.cuh).
//
// Arithmetic follows the reference's association order and is compiled with
// --fmad=false: decisions such as `< f64::EPSILON` depend on the last bits.
// Included once per pass: rk_kernels.cu compiles these headers twice — into namespace rk (exact divisions with their fall-back
// branch) and, with RK_PASS2 defined and `rk` renamed, into namespace rk_lazy (rk_math.cuh, RK_LAZY_DDIV).
#ifndef RK_PASS2
#ifndef RK_H_CORE_1
#define RK_H_CORE_1
#define RK_H_CORE_GO
#endif
#else
#ifndef RK_H_CORE_2
#define RK_H_CORE_2
#define RK_H_CORE_GO
#endif
#endif
#ifdef RK_H_CORE_GO
#undef RK_H_CORE_GO
#include <cfloat>
#include <cstdint>

#include "rk_math.cuh"

namespace rk {

constexpr double EPS = DBL_EPSILON;
// profile.rs:14-22
constexpr double V_EPS = 1e-12, A_EPS = 1e-12, J_EPS = 1e-12;
constexpr double P_PRECISION = 1e-8, V_PRECISION = 1e-8, A_PRECISION = 1e-10;
constexpr double T_MAX = 1e12;
#define RK_INF (__longlong_as_double(0x7ff0000000000000LL))

// profile.rs:25-49
enum : int { L_ACC0_ACC1_VEL = 0, L_VEL = 1, L_ACC0 = 2, L_ACC1 = 3, L_ACC0_ACC1 = 4, L_ACC0_VEL = 5, L_ACC1_VEL = 6, L_NONE = 7 };
enum : int { UDDU = 0, UDUD = 1 };
enum : int { DIR_UP = 0, DIR_DOWN = 1 };

// Boundary state of one DoF after the brake pre-trajectory (Profile::p[0], v[0], a[0], pf, vf, af).
struct Bound {
    double p0, v0, a0, pf, vf, af;
};

// util.rs:27-33
__device__ __forceinline__ void integrate(double t, double p0, double v0, double a0, double j, double& p, double& v, double& a) {
    p = p0 + t * (v0 + t * (a0 / 2.0 + div6(t * j)));
    v = v0 + t * (a0 + t * j / 2.0);
    a = a0 + t * j;
}

// Sum of the seven durations in the reference's order (profile.rs:338-344).
__device__ __forceinline__ double t_sum6(const double (&t)[7]) {
    return (((((t[0] + t[1]) + t[2]) + t[3]) + t[4]) + t[5]) + t[6];
}

// Jerk pattern of a third-order position profile (profile.rs:373-393).
__device__ __forceinline__ void jerk_pattern_pos(const double (&t)[7], int cs, double jf, double (&j)[7]) {
    j[0] = (t[0] > 0.0) ? jf : 0.0;
    j[1] = 0.0;
    j[2] = (t[2] > 0.0) ? -jf : 0.0;
    j[3] = 0.0;
    if (cs == UDDU) {
        j[4] = (t[4] > 0.0) ? -jf : 0.0;
        j[6] = (t[6] > 0.0) ? jf : 0.0;
    } else {
        j[4] = (t[4] > 0.0) ? jf : 0.0;
        j[6] = (t[6] > 0.0) ? -jf : 0.0;
    }
    j[5] = 0.0;
}

__device__ __forceinline__ bool forces_a3_zero(int lim) {
    return lim == L_ACC0_ACC1_VEL || lim == L_ACC0_ACC1 || lim == L_ACC0_VEL || lim == L_ACC1_VEL || lim == L_VEL;
}

// Profile::check for the third-order position interface (profile.rs:323-482,
// set_limits == false as in every call of the reference, :498). Returns
// validity only; nothing is stored.
__device__ __forceinline__ bool check_pos3_body(double t0, double t1, double t2, double t3, double t4, double t5, double t6, int cs, int lim,
                                                double jf, double v_max, double v_min, double a_max, double a_min, double p0, double v0,
                                                double a0, double pf, double vf, double af) {
    // The reference tests t < 0 (profile.rs:336-340). A NaN duration passes that test there, poisons p[7] and fails the final
    // comparison, so it is rejected here at once as well: NaN operands would take the slow path of every division below.
    if (!(t0 >= 0.0) || !(t1 >= 0.0) || !(t2 >= 0.0) || !(t3 >= 0.0) || !(t4 >= 0.0) || !(t5 >= 0.0) || !(t6 >= 0.0)) return false;
    const double t[7] = {t0, t1, t2, t3, t4, t5, t6};

    if ((lim == L_ACC0_ACC1_VEL || lim == L_ACC0_VEL || lim == L_ACC1_VEL || lim == L_VEL) && t3 < EPS) return false;
    if ((lim == L_ACC0 || lim == L_ACC0_ACC1) && t1 < EPS) return false;
    if ((lim == L_ACC1 || lim == L_ACC0_ACC1) && t5 < EPS) return false;
    if (t_sum6(t) > T_MAX) return false;

    double j[7];
    jerk_pattern_pos(t, cs, jf, j);

    const bool up = v_max > 0.0;
    const double v_upp_lim = (up ? v_max : v_min) + V_EPS;
    const double v_low_lim = (up ? v_min : v_max) - V_EPS;
    const bool zero_a3 = forces_a3_zero(lim);

    double a[8], v[8], p = p0;
    a[0] = a0;
    v[0] = v0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        if (i & 1) {
            // j[1], j[3], j[5] are exactly +0 and every t[i] is finite and >= 0 here (NaN and the infinite sum were rejected
            // above), so t * j is +0 and the reference's `+ tj`, `+ tj / 2`, `+ tj / 6` add +0: they can only turn a -0 into +0,
            // which no comparison of this function sees (the integrated values themselves never leave it). Five operations
            // less per constant-acceleration phase.
            a[i + 1] = a[i];
            v[i + 1] = v[i] + t[i] * a[i];
            p = p + t[i] * (v[i] + t[i] * (a[i] / 2.0));
        } else {
            const double tj = t[i] * j[i];
            a[i + 1] = a[i] + tj;
            v[i + 1] = v[i] + t[i] * (a[i] + tj / 2.0);
            p = p + t[i] * (v[i] + t[i] * (a[i] / 2.0 + div6(tj)));
        }

        if (i == 2 && zero_a3) a[3] = 0.0;

        // (after a constant-acceleration phase a[i + 1] * a[i] is a square: the test can only fire for i = 2, 4, 6)
        if (i > 1 && !(i & 1) && a[i + 1] * a[i] < -EPS) {
            const double v_a_zero = v[i] - sdiv(a[i] * a[i], 2.0 * j[i]);
            if (v_a_zero > v_upp_lim || v_a_zero < v_low_lim) return false;
        }
    }

    const double a_upp_lim = (up ? a_max : a_min) + A_EPS;
    const double a_low_lim = (up ? a_min : a_max) - A_EPS;

    return fabs(p - pf) < P_PRECISION && fabs(v[7] - vf) < V_PRECISION && fabs(a[7] - af) < A_PRECISION
        && a[1] >= a_low_lim && a[1] <= a_upp_lim && a[3] >= a_low_lim && a[3] <= a_upp_lim && a[5] >= a_low_lim && a[5] <= a_upp_lim
        && v[3] <= v_upp_lim && v[3] >= v_low_lim && v[4] <= v_upp_lim && v[4] >= v_low_lim
        && v[5] <= v_upp_lim && v[5] >= v_low_lim && v[6] <= v_upp_lim && v[6] >= v_low_lim;
}

// Out-of-line copy for the families with many candidates (one body in the kernel instead of one per call site) ...
__device__ __noinline__ bool check_pos3(double t0, double t1, double t2, double t3, double t4, double t5, double t6, int cs, int lim,
                                        double jf, double v_max, double v_min, double a_max, double a_min, double p0, double v0,
                                        double a0, double pf, double vf, double af) {
    return check_pos3_body(t0, t1, t2, t3, t4, t5, t6, cs, lim, jf, v_max, v_min, a_max, a_min, p0, v0, a0, pf, vf, af);
}

struct Lim {  // direction-resolved limits handed to the family functions
    Den v_max, v_min, a_max, a_min, j_max;
};

__device__ __forceinline__ bool check_pos3(const double (&t)[7], int cs, int lim, double jf, const Lim& L, const Bound& b) {
    return check_pos3(t[0], t[1], t[2], t[3], t[4], t[5], t[6], cs, lim, jf, L.v_max, L.v_min, L.a_max, L.a_min, b.p0, b.v0, b.a0, b.pf, b.vf, b.af);
}

// ... and the inlined form for kernels with a single call site: no argument shuffling, and `cs` / `lim` fold to constants.
__device__ __forceinline__ bool check_pos3_inline(const double (&t)[7], int cs, int lim, double jf, const Lim& L, const Bound& b) {
    return check_pos3_body(t[0], t[1], t[2], t[3], t[4], t[5], t[6], cs, lim, jf, L.v_max, L.v_min, L.a_max, L.a_min, b.p0, b.v0, b.a0, b.pf, b.vf, b.af);
}

// profile.rs:502-516 check_with_timing_full
__device__ __forceinline__ bool check_pos3_full(const double (&t)[7], int cs, int lim, double jf, const Lim& L, const Bound& b, double j_max) {
    return (fabs(jf) < fabs(j_max) + J_EPS) && check_pos3(t, cs, lim, jf, L, b);
}

// Profile::check_for_velocity (profile.rs:121-207); note the UDUD pattern ends with +jf (Q7).
__device__ __noinline__ bool check_vel3(double t0, double t1, double t2, double t3, double t4, double t5, double t6, int cs, int lim,
                                        double jf, double a_max, double a_min, double v0, double a0, double vf, double af) {
    if (t0 < 0.0 || t1 < 0.0 || t2 < 0.0 || t3 < 0.0 || t4 < 0.0 || t5 < 0.0 || t6 < 0.0) return false;
    const double t[7] = {t0, t1, t2, t3, t4, t5, t6};
    if (lim == L_ACC0 && t1 < EPS) return false;
    if (t_sum6(t) > T_MAX) return false;

    double j[7];
    j[0] = (t0 > 0.0) ? jf : 0.0; j[1] = 0.0; j[2] = (t2 > 0.0) ? -jf : 0.0; j[3] = 0.0; j[5] = 0.0;
    if (cs == UDDU) { j[4] = (t4 > 0.0) ? -jf : 0.0; j[6] = (t6 > 0.0) ? jf : 0.0; }
    else { j[4] = (t4 > 0.0) ? jf : 0.0; j[6] = (t6 > 0.0) ? jf : 0.0; }

    double a[8], v = v0;
    a[0] = a0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        a[i + 1] = a[i] + t[i] * j[i];
        v = v + t[i] * (a[i] + t[i] * j[i] / 2.0);
    }
    const bool up = a_max > 0.0;
    const double a_upp_lim = (up ? a_max : a_min) + A_EPS;
    const double a_low_lim = (up ? a_min : a_max) - A_EPS;
    return fabs(v - vf) < V_PRECISION && fabs(a[7] - af) < A_PRECISION
        && a[1] >= a_low_lim && a[3] >= a_low_lim && a[5] >= a_low_lim && a[1] <= a_upp_lim && a[3] <= a_upp_lim && a[5] <= a_upp_lim;
}

__device__ __forceinline__ bool check_vel3(const double (&t)[7], int cs, int lim, double jf, double a_max, double a_min, const Bound& b) {
    return check_vel3(t[0], t[1], t[2], t[3], t[4], t[5], t[6], cs, lim, jf, a_max, a_min, b.v0, b.a0, b.vf, b.af);
}

// Profile::check_for_second_order (profile.rs:549-637); |v7 - vf| uses P_PRECISION (Q8).
__device__ __noinline__ bool check_pos2(double t0, double t1, double t2, double t3, double t4, double t5, double t6, int cs, double a_up,
                                        double a_down, double v_max, double v_min, double p0, double v0, double pf, double vf, double af) {
    if (t0 < 0.0 || t1 < 0.0 || t2 < 0.0 || t3 < 0.0 || t4 < 0.0 || t5 < 0.0 || t6 < 0.0) return false;
    const double t[7] = {t0, t1, t2, t3, t4, t5, t6};
    if (t_sum6(t) > T_MAX) return false;

    double a[7];
    a[0] = (t0 > 0.0) ? a_up : 0.0; a[1] = 0.0; a[2] = (t2 > 0.0) ? a_down : 0.0; a[3] = 0.0; a[5] = 0.0;
    if (cs == UDDU) { a[4] = (t4 > 0.0) ? a_down : 0.0; a[6] = (t6 > 0.0) ? a_up : 0.0; }
    else { a[4] = (t4 > 0.0) ? a_up : 0.0; a[6] = (t6 > 0.0) ? a_down : 0.0; }

    const bool up = v_max > 0.0;
    const double v_upp_lim = up ? v_max + V_EPS : v_min + V_EPS;
    const double v_low_lim = up ? v_min - V_EPS : v_max - V_EPS;

    double v[8], p = p0;
    v[0] = v0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        v[i + 1] = v[i] + t[i] * a[i];
        p = p + t[i] * (v[i] + t[i] * a[i] / 2.0);
    }
    (void)af;
    return fabs(p - pf) < P_PRECISION && fabs(v[7] - vf) < P_PRECISION
        && v[2] <= v_upp_lim && v[2] >= v_low_lim && v[3] <= v_upp_lim && v[3] >= v_low_lim && v[4] <= v_upp_lim && v[4] >= v_low_lim
        && v[5] <= v_upp_lim && v[5] >= v_low_lim && v[6] <= v_upp_lim && v[6] >= v_low_lim && v[7] <= v_upp_lim && v[7] >= v_low_lim;
}

__device__ __forceinline__ bool check_pos2(const double (&t)[7], int cs, double a_up, double a_down, double v_max, double v_min, const Bound& b) {
    return check_pos2(t[0], t[1], t[2], t[3], t[4], t[5], t[6], cs, a_up, a_down, v_max, v_min, b.p0, b.v0, b.pf, b.vf, b.af);
}

// Profile::check_for_second_order_velocity (profile.rs:256-294): only t[1] matters.
__device__ __forceinline__ bool check_vel2(const double (&t)[7], double a_up, const Bound& b) {
    if (t[1] < 0.0) return false;
    if (t[1] > T_MAX) return false;  // t_sum[6] == t[1]
    double a[7] = {0.0, (t[1] > 0.0) ? a_up : 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double v = b.v0;
#pragma unroll
    for (int i = 0; i < 7; ++i) v = v + t[i] * a[i];
    return fabs(v - b.vf) < V_PRECISION;
}

// Profile::check_for_first_order (profile.rs:685-726): only t[3] matters.
__device__ __forceinline__ bool check_pos1(const double (&t)[7], double v_up, const Bound& b) {
    if (t[3] < 0.0) return false;
    if (t[3] > T_MAX) return false;
    double v[7] = {0.0, 0.0, 0.0, (t[3] > 0.0) ? v_up : 0.0, 0.0, 0.0, 0.0};
    double p = b.p0;
#pragma unroll
    for (int i = 0; i < 7; ++i) p = p + t[i] * (v[i] + t[i] * 0.0 / 2.0);
    return fabs(p - b.pf) < P_PRECISION;
}

// ------------------------------------------------------------------ brake.rs
struct Brake {
    double t[2], j[2], a0;  // a0: BrakeProfile::a[0] of the second-order variant
};

__device__ __forceinline__ double v_at_t(double v0, double a0, double j, double t) { return v0 + t * (a0 + j * t / 2.0); }  // brake.rs:15
__device__ __forceinline__ double v_at_a_zero(double v0, double a0, double j) { return v0 + sdiv(a0 * a0, 2.0 * j); }       // brake.rs:20
constexpr double BRAKE_EPS = 2.2e-14;                                                                                       // brake.rs:12

// brake.rs:77-105
__device__ __forceinline__ void velocity_brake(Brake& b, double v0, double a0, double v_max, double v_min, double a_min, double j_max) {
    b.j[0] = -j_max;
    const double t_to_a_min = sdiv(a0 - a_min, j_max);
    const double t_to_v_max = sdiv(a0, j_max) + sqrt(a0 * a0 + 2.0 * j_max * (v0 - v_max)) / fabs(j_max);
    const double t_to_v_min = sdiv(a0, j_max) + sqrt(a0 * a0 / 2.0 + j_max * (v0 - v_min)) / fabs(j_max);
    const double t_min_to_v_max = rmin(t_to_v_max, t_to_v_min);
    if (t_to_a_min < t_min_to_v_max) {
        const double v_at_a_min = v_at_t(v0, a0, -j_max, t_to_a_min);
        const double t_to_v_max_with_constant = -(v_at_a_min - v_max) / a_min;
        const double t_to_v_min_with_constant = a_min / (2.0 * j_max) - (v_at_a_min - v_min) / a_min;
        b.t[0] = rmax(t_to_a_min - BRAKE_EPS, 0.0);
        b.t[1] = rmax(rmin(t_to_v_max_with_constant, t_to_v_min_with_constant), 0.0);
    } else {
        b.t[0] = rmax(t_min_to_v_max - BRAKE_EPS, 0.0);
    }
}

// brake.rs:46-75
__device__ __forceinline__ void acceleration_brake(Brake& b, double v0, double a0, double v_max, double v_min, double a_max, double a_min, double j_max) {
    b.j[0] = -j_max;
    const double t_to_a_max = sdiv(a0 - a_max, j_max);
    const double t_to_a_zero = sdiv(a0, j_max);
    const double v_at_a_max = v_at_t(v0, a0, -j_max, t_to_a_max);
    const double v_at_a_zero_ = v_at_t(v0, a0, -j_max, t_to_a_zero);
    if ((v_at_a_zero_ > v_max && j_max > 0.0) || (v_at_a_zero_ < v_max && j_max < 0.0)) {
        velocity_brake(b, v0, a0, v_max, v_min, a_min, j_max);
    } else if ((v_at_a_max < v_min && j_max > 0.0) || (v_at_a_max > v_min && j_max < 0.0)) {
        const double t_to_v_min = -(v_at_a_max - v_min) / a_max;
        const double t_to_v_max = -a_max / (2.0 * j_max) - (v_at_a_max - v_max) / a_max;
        b.t[0] = t_to_a_max + BRAKE_EPS;
        b.t[1] = rmin(t_to_v_min, rmax(t_to_v_max - BRAKE_EPS, 0.0));
    } else {
        b.t[0] = t_to_a_max + BRAKE_EPS;
    }
}

// brake.rs:107-139
__device__ __forceinline__ void position_brake(Brake& b, double v0, double a0, double v_max, double v_min, double a_max, double a_min, double j_max) {
    b.t[0] = 0.0; b.t[1] = 0.0; b.j[0] = 0.0; b.j[1] = 0.0;
    if (j_max == 0.0 || a_max == 0.0 || a_min == 0.0) return;
    if (a0 > a_max) {
        acceleration_brake(b, v0, a0, v_max, v_min, a_max, a_min, j_max);
    } else if (a0 < a_min) {
        acceleration_brake(b, v0, a0, v_min, v_max, a_min, a_max, -j_max);
    } else if ((v0 > v_max && v_at_a_zero(v0, a0, -j_max) > v_min) || (a0 > 0.0 && v_at_a_zero(v0, a0, j_max) > v_max)) {
        velocity_brake(b, v0, a0, v_max, v_min, a_min, j_max);
    } else if ((v0 < v_min && v_at_a_zero(v0, a0, j_max) < v_max) || (a0 < 0.0 && v_at_a_zero(v0, a0, -j_max) < v_min)) {
        velocity_brake(b, v0, a0, v_min, v_max, a_max, -j_max);
    }
}

// brake.rs:141-167
__device__ __forceinline__ void position_brake_second_order(Brake& b, double v0, double v_max, double v_min, double a_max, double a_min) {
    b.t[0] = 0.0; b.t[1] = 0.0; b.j[0] = 0.0; b.j[1] = 0.0; b.a0 = 0.0;
    if (a_max == 0.0 || a_min == 0.0) return;
    if (v0 > v_max) { b.a0 = a_min; b.t[0] = (v_max - v0) / a_min + EPS; }
    else if (v0 < v_min) { b.a0 = a_max; b.t[0] = (v_min - v0) / a_max + EPS; }
}

// brake.rs:169-186
__device__ __forceinline__ void velocity_brake_third_order(Brake& b, double a0, double a_max, double a_min, double j_max) {
    b.t[0] = 0.0; b.t[1] = 0.0; b.j[0] = 0.0; b.j[1] = 0.0;
    if (j_max == 0.0) return;
    if (a0 > a_max) { b.j[0] = -j_max; b.t[0] = sdiv(a0 - a_max, j_max) + EPS; }
    else if (a0 < a_min) { b.j[0] = j_max; b.t[0] = -sdiv(a0 - a_min, j_max) + EPS; }
}

// ------------------------------------------------------------------ roots.rs
// PositiveSet semantics (roots.rs:64-122): keeps values >= 0, drops exact duplicates, iterated ascending.
struct Roots {
    double r[4];
    int n;
    __device__ __forceinline__ Roots() : n(0) { r[0] = r[1] = r[2] = r[3] = RK_INF; }
    __device__ __forceinline__ void insert(double v) {
        if (!(v >= 0.0)) return;
        if ((n > 0 && r[0] == v) || (n > 1 && r[1] == v) || (n > 2 && r[2] == v) || (n > 3 && r[3] == v)) return;
        if (n == 0) r[0] = v; else if (n == 1) r[1] = v; else if (n == 2) r[2] = v; else r[3] = v;
        ++n;
    }
    // ascending order; unused slots hold +inf and stay behind. +inf itself can be a member (it is >= 0).
    __device__ __forceinline__ void sort() {
#define RK_CSWAP(i, k) { const double lo_ = r[i] <= r[k] ? r[i] : r[k]; const double hi_ = r[i] <= r[k] ? r[k] : r[i]; r[i] = lo_; r[k] = hi_; }
        RK_CSWAP(0, 1) RK_CSWAP(2, 3) RK_CSWAP(0, 2) RK_CSWAP(1, 3) RK_CSWAP(1, 2)
#undef RK_CSWAP
    }
    __device__ __forceinline__ double get(int i) const { return i == 0 ? r[0] : (i == 1 ? r[1] : (i == 2 ? r[2] : r[3])); }
};

constexpr double COS_120 = -0.50;
constexpr double SIN_120 = 0.866025403784438600;  // roots.rs:12

// roots.rs:132-231. Note the reference's shadowing in the d ~ 0 branch: the quadratic solved is a x^2 + b x + d.
template <bool SORTED = true>  // false: roots in insertion order (Set::get_data, Q9), for the position query of profile.rs:876
__device__ __noinline__ Roots solve_cub(double a, double b, double c, double d) {
    Roots roots;
    if (fabs(d) < EPS) {
        roots.insert(0.0);
        if (fabs(a) < EPS) {
            if (fabs(b) > EPS) roots.insert(-d / b);
        } else {
            const double discriminant = b * b - 4.0 * a * d;
            if (discriminant >= 0.0) {
                const double inv2b = 1.0 / (2.0 * a);
                const double y = sqrt(discriminant);
                roots.insert((-b + y) * inv2b);
                roots.insert((-b - y) * inv2b);
            }
        }
    } else if (fabs(a) < EPS) {
        if (fabs(b) < EPS) {
            if (fabs(c) > EPS) roots.insert(-d / c);
        } else {
            const double discriminant = c * c - 4.0 * b * d;
            if (discriminant >= 0.0) {
                const double inv2b = 1.0 / (2.0 * b);
                const double y = sqrt(discriminant);
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
            const double y = sqrt(yy);
            const double uuu = -halfq + y;
            const double vvv = -halfq - y;
            const double www = fabs(uuu) > fabs(vvv) ? uuu : vvv;
            const double w = m_cbrt(www);
            roots.insert(w - p / (3.0 * w) - bover3a);
        } else if (yy < -EPS) {
            const double x = -halfq;
            const double y = sqrt(-yy);
            double theta, r;
            if (fabs(x) > EPS) {
                theta = m_atan2(y, x);
                r = sqrt(x * x - yy);
            } else {
                theta = 3.14159265358979323846 / 2.0;
                r = y;
            }
            theta /= 3.0;
            r = 2.0 * m_cbrt(r);
            double sn, cs;
            m_sincos(theta, sn, cs);
            const double ux = cs * r;
            const double uyi = sn * r;
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
    if (SORTED) roots.sort();
    return roots;
}

// roots.rs:237-274
__device__ __forceinline__ int solve_resolvent(double (&x)[3], double a, double b, double c) {
    a = div3(a);
    const double a2 = a * a;
    double q = a2 - div3(b);
    const double r = (a * (2.0 * a2 - b) + c) / 2.0;
    const double r2 = r * r;
    const double q3 = q * q * q;
    if (r2 < q3) {
        const double q_sqrt = sqrt(q);
        const double t = rmax(rmin(sdiv(r, q * q_sqrt), 1.0), -1.0);
        q = -2.0 * q_sqrt;
        const double theta = div3(m_acos(t));  // acos(1) = 0 is common after the clamp
        double sn, cs;
        m_sincos(theta, sn, cs);
        const double ux = cs * q;
        const double uyi = sn * q;
        x[0] = ux - a;
        x[1] = ux * COS_120 - uyi * SIN_120 - a;
        x[2] = ux * COS_120 + uyi * SIN_120 - a;
        return 3;
    }
    double a_ = m_cbrt(-fabs(r) - sqrt(r2 - q3));
    if (r < 0.0) a_ = -a_;
    const double b_ = (a_ == 0.0) ? 0.0 : sdiv(q, a_);
    x[0] = (a_ + b_) - a;
    x[1] = -(a_ + b_) / 2.0 - a;
    x[2] = sqrt(3.0) * (a_ - b_) / 2.0;
    if (fabs(x[2]) < EPS) {
        x[2] = x[1];
        return 2;
    }
    return 1;
}

// roots.rs:278-370: real roots >= 0 of x^4 + a x^3 + b x^2 + c x + d, sorted ascending on return.
__device__ RK_QUART_FN Roots solve_quart_monic(double a, double b, double c, double d) {
    Roots roots;
    const double a_squared = a * a;
    const double four_b = 4.0 * b;

    if (fabs(d) < EPS) {
        if (fabs(c) < EPS) {
            roots.insert(0.0);
            const double d_ = a_squared - four_b;
            if (fabs(d_) < EPS) {
                roots.insert(-a / 2.0);
            } else if (d_ > 0.0) {
                const double sqrt_d = sqrt(d_);
                roots.insert((-a - sqrt_d) / 2.0);
                roots.insert((-a + sqrt_d) / 2.0);
            }
            roots.sort();
            return roots;
        }
        if (fabs(a) < EPS && fabs(b) < EPS) {
            roots.insert(0.0);
            roots.insert(-m_cbrt(c));
            roots.sort();
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
        if (fabs(x3[1]) > fabs(y)) y = x3[1];
        if (fabs(x3[2]) > fabs(y)) y = x3[2];
    }

    double q1, q2, p1, p2;
    double d_ = y * y - 4.0 * d;
    if (fabs(d_) < EPS) {
        q2 = y / 2.0;
        q1 = q2;
        d_ = a_squared - 4.0 * (b - y);
        if (fabs(d_) < EPS) {
            p2 = a / 2.0;
            p1 = p2;
        } else {
            const double sqrt_d = sqrt(d_);
            p1 = (a + sqrt_d) / 2.0;
            p2 = (a - sqrt_d) / 2.0;
        }
    } else {
        // d_ < 0 (or NaN): the reference carries NaN through q1, q2, p1, p2 and both discriminant tests below then fail, i.e.
        // no roots. Leave here instead of sending NaN operands down the slow path of the division.
        if (!(d_ > 0.0)) return roots;
        const double sqrt_d = sqrt(d_);
        q1 = (y + sqrt_d) / 2.0;
        q2 = (y - sqrt_d) / 2.0;
        const Den q12(q1 - q2);  // one reciprocal for the two quotients
        p1 = ddiv(a * q1 - c, q12);
        p2 = ddiv(c - a * q2, q12);
    }

    const double EPS16 = 16.0 * EPS;
    d_ = p1 * p1 - 4.0 * q1;
    if (fabs(d_) < EPS16) {
        roots.insert(-p1 / 2.0);
    } else if (d_ > 0.0) {
        const double sqrt_d = sqrt(d_);
        roots.insert((-p1 - sqrt_d) / 2.0);
        roots.insert((-p1 + sqrt_d) / 2.0);
    }

    d_ = p2 * p2 - 4.0 * q2;
    if (fabs(d_) < EPS16) {
        roots.insert(-p2 / 2.0);
    } else if (d_ > 0.0) {
        const double sqrt_d = sqrt(d_);
        roots.insert((-p2 - sqrt_d) / 2.0);
        roots.insert((-p2 + sqrt_d) / 2.0);
    }
    roots.sort();
    return roots;
}

// roots.rs:400-415: polynomial c[0] x^(N-1) + ... + c[N-1], summed from the constant term upward.
template <int N>
__device__ __forceinline__ double poly_eval(const double (&c)[N], double x) {
    double result = 0.0;
    if (fabs(x) < EPS) {
        result = c[N - 1];
    } else if (fabs(x - 1.0) < EPS) {
#pragma unroll
        for (int i = 0; i < N; ++i) result += c[i];
    } else {
        double xn = 1.0;
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            result += c[i] * xn;
            xn *= x;
        }
    }
    return result;
}

// Several polynomials at ONE point: poly_eval walks the powers 1, x, x * x, ... by repeated multiplication whatever the
// polynomial, so polynomials evaluated at the same x (f and f' of a Newton step; the quintic, its derivative and second
// derivative at an extremum) share them — the same products in the same order, hence the same bits as separate poly_eval
// calls, at N - 2 multiplications and one pair of range tests less per extra polynomial. (The compiler does not find this
// itself: the products sit behind each call's own x ~ 0 / x ~ 1 branches.)
template <int M>
struct PolyPoint {
    double xn[M];
    bool zero, one;
    __device__ __forceinline__ explicit PolyPoint(double x) {
        zero = fabs(x) < EPS;
        one = fabs(x - 1.0) < EPS;
        double p = 1.0;
#pragma unroll
        for (int k = 0; k < M; ++k) { xn[k] = p; p *= x; }
    }
    template <int N>
    __device__ __forceinline__ double eval(const double (&c)[N]) const {
        static_assert(N <= M, "PolyPoint holds too few powers");
        double result = 0.0;
        if (zero) {
            result = c[N - 1];
        } else if (one) {
#pragma unroll
            for (int i = 0; i < N; ++i) result += c[i];
        } else {
#pragma unroll
            for (int i = N - 1; i >= 0; --i) result += c[i] * xn[N - 1 - i];
        }
        return result;
    }
    template <int N1, int N2>
    __device__ __forceinline__ void eval2(const double (&c1)[N1], const double (&c2)[N2], double& r1, double& r2) const {
        static_assert(N1 <= M && N2 <= M, "PolyPoint holds too few powers");
        double a = 0.0, b = 0.0;
        if (zero) {
            a = c1[N1 - 1]; b = c2[N2 - 1];
        } else if (one) {
#pragma unroll
            for (int i = 0; i < N1; ++i) a += c1[i];
#pragma unroll
            for (int i = 0; i < N2; ++i) b += c2[i];
        } else {
#pragma unroll
            for (int i = N1 - 1; i >= 0; --i) a += c1[i] * xn[N1 - 1 - i];
#pragma unroll
            for (int i = N2 - 1; i >= 0; --i) b += c2[i] * xn[N2 - 1 - i];
        }
        r1 = a; r2 = b;
    }
};

// roots.rs:379-386
template <int N>
__device__ __forceinline__ void poly_deri(const double (&c)[N], double (&d)[N - 1]) {
#pragma unroll
    for (int i = 0; i < N - 1; ++i) d[i] = (double)(N - 1 - i) * c[i];
}

// roots.rs:389-397
template <int N>
__device__ __forceinline__ void poly_monic_deri(const double (&c)[N], double (&d)[N - 1]) {
    d[0] = 1.0;
#pragma unroll
    for (int i = 1; i < N - 1; ++i) {
        const double num = (double)(N - 1 - i) * c[i];
        // the callers' degrees are 5 and 6: a literal divisor, through its constant reciprocal (rk_math.cuh)
        d[i] = (N - 1 == 5) ? div_lit<5>(num) : ((N - 1 == 6) ? div_lit<6>(num) : sdiv(num, (double)(N - 1)));
    }
}

// roots.rs:424-481: safeguarded Newton / bisection, at most 128 iterations, tolerance 1e-14.
template <int N>
__device__ __forceinline__ double shrink_interval(const double (&p)[N], double l, double h) {
    double deriv[N - 1];
    poly_deri<N>(p, deriv);
    const double fl = poly_eval<N>(p, l);
    const double fh = poly_eval<N>(p, h);
    if (fl == 0.0) return l;
    if (fh == 0.0) return h;
    if (fl > 0.0) { const double tmp = l; l = h; h = tmp; }

    double rts = (l + h) / 2.0;
    double dxold = fabs(h - l);
    double dx = dxold;
    double f, df;
    PolyPoint<N>(rts).template eval2<N, N - 1>(p, deriv, f, df);
    for (int it = 0; it < 128; ++it) {
        if ((((rts - h) * df - f) * ((rts - l) * df - f) > 0.0) || fabs(2.0 * f) > fabs(dxold) * fabs(df)) {
            dxold = dx;
            dx = (h - l) / 2.0;
            rts = l + dx;
            if (fabs(l - rts) < EPS) break;
        } else {
            dxold = dx;
            dx = sdiv(f, df);
            const double temp = rts;
            rts -= dx;
            if (fabs(temp - rts) < EPS) break;
        }
        if (fabs(dx) < 1e-14) break;
        PolyPoint<N>(rts).template eval2<N, N - 1>(p, deriv, f, df);
        if (f < 0.0) l = rts; else h = rts;
    }
    return rts;
}

// shrink_interval as a resumable state machine: the same operations in the same order, one iteration per step() call, so
// that a lane can pick up the next bracket as soon as its own has converged (k2_polish).
template <int N>
struct ShrinkState {
    double p[N], deriv[N - 1];
    double l, h, rts, dxold, dx, f, df;
    int it;
    // returns true if the root is known without iterating (then rts holds it)
    __device__ __forceinline__ bool start(double lo, double hi) {
        poly_deri<N>(p, deriv);
        l = lo; h = hi;
        const double fl = poly_eval<N>(p, l);
        const double fh = poly_eval<N>(p, h);
        if (fl == 0.0) { rts = l; return true; }
        if (fh == 0.0) { rts = h; return true; }
        if (fl > 0.0) { const double tmp = l; l = h; h = tmp; }
        rts = (l + h) / 2.0;
        dxold = fabs(h - l);
        dx = dxold;
        PolyPoint<N>(rts).template eval2<N, N - 1>(p, deriv, f, df);
        it = 0;
        return false;
    }
    // one pass of the loop body of shrink_interval; returns true when rts is final
    __device__ __forceinline__ bool step() {
        if (it >= 128) return true;
        ++it;
        if ((((rts - h) * df - f) * ((rts - l) * df - f) > 0.0) || fabs(2.0 * f) > fabs(dxold) * fabs(df)) {
            dxold = dx;
            dx = (h - l) / 2.0;
            rts = l + dx;
            if (fabs(l - rts) < EPS) return true;
        } else {
            dxold = dx;
            dx = sdiv(f, df);
            const double temp = rts;
            rts -= dx;
            if (fabs(temp - rts) < EPS) return true;
        }
        if (fabs(dx) < 1e-14) return true;
        PolyPoint<N>(rts).template eval2<N, N - 1>(p, deriv, f, df);
        if (f < 0.0) l = rts; else h = rts;
        return false;
    }
};

}  // namespace rk
#endif  // RK_H_CORE_GO
