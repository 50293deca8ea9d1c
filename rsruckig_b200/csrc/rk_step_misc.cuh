// rk_step_misc.cuh — Step 1 / Step 2 for the cheap interface variants:
// third-order velocity interface (velocity_third_step{1,2}.rs), second-order
// position (position_second_step{1,2}.rs) and velocity (velocity_second_step
// {1,2}.rs) interfaces (max_jerk = inf), first-order position interface
// (position_first_step{1,2}.rs; max_acceleration = inf as well).
// All closed form, one sqrt at most; candidate order is the reference's.
// Included once per pass: rk_kernels.cu compiles these headers twice — into namespace rk (exact divisions with their fall-back
// branch) and, with RK_PASS2 defined and `rk` renamed, into namespace rk_lazy (rk_math.cuh, RK_LAZY_DDIV).
#ifndef RK_PASS2
#ifndef RK_H_STEP_MISC_1
#define RK_H_STEP_MISC_1
#define RK_H_STEP_MISC_GO
#endif
#else
#ifndef RK_H_STEP_MISC_2
#define RK_H_STEP_MISC_2
#define RK_H_STEP_MISC_GO
#endif
#endif
#ifdef RK_H_STEP_MISC_GO
#undef RK_H_STEP_MISC_GO
#include "rk_step1.cuh"

namespace rk {

__device__ __forceinline__ void zero_t(double (&t)[7]) {
#pragma unroll
    for (int i = 0; i < 7; ++i) t[i] = 0.0;
}

__device__ __forceinline__ void set_cand(Cand& c, const double (&t)[7], real ctrl, real ctrl2, int lim, int dir, int cs) {
    copy_t(c.t, t);
    c.dur = t_sum6(t);
    c.ctrl = ctrl; c.ctrl2 = ctrl2; c.lim = lim; c.dir = dir; c.cs = cs;
}

// ---------------------------------------------------------------- velocity interface, third order
struct Step1Vel3 {
    Bound b;
    real vd;
    Cand vp[6];
    int count;

    __device__ __forceinline__ bool emit(const double (&t)[7], int lim, real a_max, real a_min, double j_max) {
        if (!check_vel3(t, UDDU, lim, j_max, a_max, a_min, b)) return false;
        set_cand(vp[count], t, j_max, 0.0, lim, (a_max > 0.0) ? DIR_UP : DIR_DOWN, UDDU);
        if (count < 5) ++count;
        return true;
    }
    // velocity_third_step1.rs:46-62
    __device__ __forceinline__ void time_acc0(real a_max, real a_min, double j_max) {
        double t[7];
        zero_t(t);
        t[0] = (-b.a0 + a_max) / j_max;
        t[1] = (b.a0 * b.a0 + b.af * b.af) / (2.0 * a_max * j_max) - a_max / j_max + vd / a_max;
        t[2] = (-b.af + a_max) / j_max;
        emit(t, L_ACC0, a_max, a_min, j_max);
    }
    // velocity_third_step1.rs:64-117
    __device__ __forceinline__ void time_none(real a_max, real a_min, real j_max, bool first_only) {
        real h1 = (b.a0 * b.a0 + b.af * b.af) * 0.5 + j_max * vd;
        if (!(h1 >= 0.0)) return;
        h1 = sqrt(h1);
        double t[7];
        zero_t(t);
        t[0] = -(b.a0 + h1) / j_max;
        t[2] = -(b.af + h1) / j_max;
        if (emit(t, L_NONE, a_max, a_min, j_max) && first_only) return;
        t[0] = (-b.a0 + h1) / j_max;
        t[2] = (-b.af + h1) / j_max;
        emit(t, L_NONE, a_max, a_min, j_max);
    }
};

// velocity_third_step1.rs:163-242
__device__ __noinline__ bool step1_vel3(const Bound& b, real a_max, real a_min, real j_max, real bd, BlockRec& blk) {
    Step1Vel3 s;
    s.b = b;
    s.vd = b.vf - b.v0;
    s.count = 0;
    blk.has_a = 0; blk.has_b = 0;

    if (j_max == 0.0) {  // zero-limit special case (:165-177, :119-161)
        if (fabs(b.af - b.a0) > EPS) return false;
        double t[7];
        zero_t(t);
        bool ok = false;
        if (fabs(b.a0) > EPS) {
            t[3] = s.vd / b.a0;
            ok = check_vel3(t, UDDU, L_NONE, 0.0, a_max, a_min, b);
        } else if (fabs(s.vd) < EPS) {
            ok = check_vel3(t, UDDU, L_NONE, 0.0, a_max, a_min, b);
        }
        if (!ok) return false;
        set_cand(blk.p_min, t, 0.0, 0.0, L_NONE, (a_max > 0.0) ? DIR_UP : DIR_DOWN, UDDU);
        blk.t_min = blk.p_min.dur + bd + 0.0;
        if (fabs(b.a0) > EPS) { blk.has_a = 1; blk.a_left = blk.t_min; blk.a_right = RK_INF; blk.pa = blk.p_min; }
        return true;
    }

    if (fabs(b.af) < EPS) {
        const bool up = s.vd >= 0.0;
        const real am = up ? a_max : a_min, an = up ? a_min : a_max, jm = up ? j_max : -j_max;
        s.time_none(am, an, jm, true);
        if (s.count > 0) return calculate_block(blk, s.vp, s.count, bd);
        s.time_acc0(am, an, jm);
        if (s.count > 0) return calculate_block(blk, s.vp, s.count, bd);
        s.time_none(an, am, -jm, true);
        if (s.count > 0) return calculate_block(blk, s.vp, s.count, bd);
        s.time_acc0(an, am, -jm);
    } else {
        s.time_none(a_max, a_min, j_max, false);
        s.time_none(a_min, a_max, -j_max, false);
        s.time_acc0(a_max, a_min, j_max);
        s.time_acc0(a_min, a_max, -j_max);
    }
    return calculate_block(blk, s.vp, s.count, bd);
}

// velocity_third_step2.rs:42-191. On success c holds the profile; the reference then sets pf = p[7].
__device__ __forceinline__ bool step2_vel3_dir(const Bound& b, real tf, real vd, real ad, real a_max, real a_min, real j_max, Cand& c) {
    const int dir = (a_max > 0.0) ? DIR_UP : DIR_DOWN;
    double t[7];
    // time_acc0 — UD solution 1/2
    {
        const real h1 = sqrt((-ad * ad + 2.0 * j_max * ((b.a0 + b.af) * tf - 2.0 * vd)) / (j_max * j_max) + tf * tf);
        zero_t(t);
        t[0] = ad / twice(j_max) + (tf - h1) * 0.5;
        t[1] = h1;
        t[2] = tf - (t[0] + h1);
        if (check_vel3(t, UDDU, L_ACC0, j_max, a_max, a_min, b)) { set_cand(c, t, j_max, 0.0, L_ACC0, dir, UDDU); return true; }
    }
    // UU solution
    {
        const real h1 = -ad + j_max * tf;
        zero_t(t);
        t[0] = -ad * ad / (2.0 * j_max * h1) + (vd - b.a0 * tf) / h1;
        t[1] = -ad / j_max + tf;
        t[6] = tf - (t[0] + t[1]);
        if (check_vel3(t, UDDU, L_ACC0, j_max, a_max, a_min, b)) { set_cand(c, t, j_max, 0.0, L_ACC0, dir, UDDU); return true; }
    }
    // UU solution, 2 step
    {
        zero_t(t);
        t[1] = -ad / j_max + tf;
        t[6] = ad / j_max;
        if (check_vel3(t, UDDU, L_ACC0, j_max, a_max, a_min, b)) { set_cand(c, t, j_max, 0.0, L_ACC0, dir, UDDU); return true; }
    }
    // time_none
    if (fabs(b.a0) < EPS && fabs(b.af) < EPS && fabs(vd) < EPS) {
        zero_t(t);
        t[1] = tf;
        if (check_vel3(t, UDDU, L_NONE, j_max, a_max, a_min, b)) { set_cand(c, t, j_max, 0.0, L_NONE, dir, UDDU); return true; }
    }
    {
        const real h1 = 2.0 * (b.af * tf - vd);
        zero_t(t);
        t[0] = h1 / ad;
        t[1] = tf - t[0];
        const real jf = ad * ad / h1;  // not compared with j_max (Q10)
        if (check_vel3(t, UDDU, L_NONE, jf, a_max, a_min, b)) { set_cand(c, t, jf, 0.0, L_NONE, dir, UDDU); return true; }
    }
    return false;
}

__device__ __noinline__ bool step2_vel3(const Bound& b, real tf, real a_max, real a_min, real j_max, Cand& c) {
    const real vd = b.vf - b.v0, ad = b.af - b.a0;
    if (vd > 0.0) return step2_vel3_dir(b, tf, vd, ad, a_max, a_min, j_max, c) || step2_vel3_dir(b, tf, vd, ad, a_min, a_max, -j_max, c);
    return step2_vel3_dir(b, tf, vd, ad, a_min, a_max, -j_max, c) || step2_vel3_dir(b, tf, vd, ad, a_max, a_min, j_max, c);
}

// ---------------------------------------------------------------- position interface, second order
struct Step1Pos2 {
    Bound b;
    real pd;
    Cand vp[6];
    int count;

    // position_second_step1.rs:49-55: stores at most three (Q12)
    __device__ __forceinline__ bool emit(const double (&t)[7], int lim, real v_max, real v_min, real a_max, double a_min) {
        if (!check_pos2(t, UDDU, a_max, a_min, v_max, v_min, b)) return false;
        if (count < 3) {
            set_cand(vp[count], t, a_max, a_min, lim, (v_max > 0.0) ? DIR_UP : DIR_DOWN, UDDU);
            ++count;
        }
        return true;
    }
    // :57-88
    __device__ __forceinline__ void time_acc0(real v_max, real v_min, real a_max, double a_min) {
        double t[7];
        zero_t(t);
        t[0] = (-b.v0 + v_max) / a_max;
        t[1] = (a_min * b.v0 * b.v0 - a_max * b.vf * b.vf) / (2.0 * a_max * a_min * v_max) + v_max * (a_max - a_min) / (2.0 * a_max * a_min) + pd / v_max;
        t[2] = (b.vf - v_max) / a_min;
        emit(t, L_ACC0, v_max, v_min, a_max, a_min);
    }
    // :90-154
    __device__ __forceinline__ void time_none(real v_max, real v_min, real a_max, real a_min, bool first_only) {
        real h1 = (a_max * b.vf * b.vf - a_min * b.v0 * b.v0 - 2.0 * a_max * a_min * pd) / (a_max - a_min);
        if (!(h1 >= 0.0)) return;
        h1 = sqrt(h1);
        double t[7];
        zero_t(t);
        t[0] = -(b.v0 + h1) / a_max;
        t[2] = (b.vf + h1) / a_min;
        if (emit(t, L_NONE, v_max, v_min, a_max, a_min) && first_only) return;
        t[0] = (-b.v0 + h1) / a_max;
        t[2] = (b.vf - h1) / a_min;
        emit(t, L_NONE, v_max, v_min, a_max, a_min);
    }
};

// position_second_step1.rs:212-312
__device__ __noinline__ bool step1_pos2(const Bound& b, real v_max, real v_min, real a_max, real a_min, real bd, BlockRec& blk) {
    Step1Pos2 s;
    s.b = b;
    s.pd = b.pf - b.p0;
    s.count = 0;
    blk.has_a = 0; blk.has_b = 0;

    if (v_max == 0.0 && v_min == 0.0) {  // zero-limit special case (:214-226, :155-210)
        if (fabs(b.vf - b.v0) > EPS) return false;
        double t[7];
        zero_t(t);
        bool ok = check_pos2(t, UDDU, 0.0, 0.0, v_max, v_min, b);
        if (!ok) {
            if (fabs(b.v0) > EPS) {
                t[3] = s.pd / b.v0;
                ok = check_pos2(t, UDDU, 0.0, 0.0, v_max, v_min, b);
            } else if (fabs(s.pd) < EPS) {
                ok = check_pos2(t, UDDU, 0.0, 0.0, v_max, v_min, b);
            }
        }
        if (!ok) return false;
        set_cand(blk.p_min, t, 0.0, 0.0, L_NONE, (v_max > 0.0) ? DIR_UP : DIR_DOWN, UDDU);
        blk.t_min = blk.p_min.dur + bd + 0.0;
        if (fabs(b.v0) > EPS) { blk.has_a = 1; blk.a_left = blk.t_min; blk.a_right = RK_INF; blk.pa = blk.p_min; }
        return true;
    }

    if (fabs(b.vf) < EPS) {
        const bool up = s.pd >= 0.0;
        const real vm = up ? v_max : v_min, vn = up ? v_min : v_max, am = up ? a_max : a_min, an = up ? a_min : a_max;
        s.time_none(vm, vn, am, an, true);
        if (s.count > 0) return calculate_block(blk, s.vp, s.count, bd);
        s.time_acc0(vm, vn, am, an);
        if (s.count > 0) return calculate_block(blk, s.vp, s.count, bd);
        s.time_none(vn, vm, an, am, true);
        if (s.count > 0) return calculate_block(blk, s.vp, s.count, bd);
        s.time_acc0(vn, vm, an, am);
    } else {
        s.time_none(v_max, v_min, a_max, a_min, false);
        s.time_none(v_min, v_max, a_min, a_max, false);
        s.time_acc0(v_max, v_min, a_max, a_min);
        s.time_acc0(v_min, v_max, a_min, a_max);
    }
    return calculate_block(blk, s.vp, s.count, bd);
}

// position_second_step2.rs:48-216; on success the reference sets pf = p[7].
__device__ __forceinline__ bool step2_pos2_dir(const Bound& b, real tf, real pd, real vd, real v_max, real v_min, real a_max, real a_min, Cand& c) {
    const int dir = (v_max > 0.0) ? DIR_UP : DIR_DOWN;
    double t[7];
    {
        const real h1 = sqrt((2.0 * a_max * (pd - tf * b.vf) - 2.0 * a_min * (pd - tf * b.v0) + vd * vd) / (a_max * a_min) + tf * tf);
        zero_t(t);
        t[0] = (a_max * vd - a_max * a_min * (tf - h1)) / (a_max * (a_max - a_min));
        t[1] = h1;
        t[2] = tf - (t[0] + h1);
        if (check_pos2(t, UDDU, a_max, a_min, v_max, v_min, b)) { set_cand(c, t, a_max, a_min, L_ACC0, dir, UDDU); return true; }
    }
    {
        const real h1 = -vd + a_max * tf;
        zero_t(t);
        t[0] = -vd * vd / (2.0 * a_max * h1) + (pd - b.v0 * tf) / h1;
        t[1] = -vd / a_max + tf;
        t[6] = tf - (t[0] + t[1]);
        if (check_pos2(t, UDDU, a_max, a_min, v_max, v_min, b)) { set_cand(c, t, a_max, a_min, L_ACC0, dir, UDDU); return true; }
    }
    {
        zero_t(t);
        t[1] = -vd / a_max + tf;
        t[6] = vd / a_max;
        if (check_pos2(t, UDDU, a_max, a_min, v_max, v_min, b)) { set_cand(c, t, a_max, a_min, L_ACC0, dir, UDDU); return true; }
    }
    if (fabs(b.v0) < EPS && fabs(b.vf) < EPS && fabs(pd) < EPS) {
        zero_t(t);
        t[1] = tf;
        if (check_pos2(t, UDDU, a_max, a_min, v_max, v_min, b)) { set_cand(c, t, a_max, a_min, L_NONE, dir, UDDU); return true; }
    }
    {
        const real h1 = 2.0 * (b.vf * tf - pd);
        zero_t(t);
        t[0] = h1 / vd;
        t[1] = tf - t[0];
        const real af = vd * vd / h1;
        if ((a_min - 1e-12 < af) && (af < a_max + 1e-12) && check_pos2(t, UDDU, af, -af, v_max, v_min, b)) {
            set_cand(c, t, af, -af, L_NONE, dir, UDDU);
            return true;
        }
    }
    return false;
}

__device__ __noinline__ bool step2_pos2(const Bound& b, real tf, real v_max, real v_min, real a_max, real a_min, Cand& c) {
    const real pd = b.pf - b.p0, vd = b.vf - b.v0;
    if (pd > 0.0) return step2_pos2_dir(b, tf, pd, vd, v_max, v_min, a_max, a_min, c) || step2_pos2_dir(b, tf, pd, vd, v_min, v_max, a_min, a_max, c);
    return step2_pos2_dir(b, tf, pd, vd, v_min, v_max, a_min, a_max, c) || step2_pos2_dir(b, tf, pd, vd, v_max, v_min, a_max, a_min, c);
}

// ---------------------------------------------------------------- velocity interface, second order
// velocity_second_step1.rs:23-46
__device__ __forceinline__ bool step1_vel2(const Bound& b, real a_max, real a_min, real bd, BlockRec& blk) {
    const real vd = b.vf - b.v0;
    const real af = (vd > 0.0) ? a_max : a_min;
    double t[7];
    zero_t(t);
    t[1] = vd / af;
    blk.has_a = 0; blk.has_b = 0;
    if (!check_vel2(t, af, b)) return false;
    set_cand(blk.p_min, t, af, 0.0, L_ACC0, (af > 0.0) ? DIR_UP : DIR_DOWN, UDDU);
    blk.p_min.dur = t[1];  // t_sum[6] of this variant is t[1] (profile.rs:262-268)
    blk.t_min = blk.p_min.dur + bd + 0.0;
    return true;
}

// velocity_second_step2.rs:23-46 (+ profile.rs:308-320)
__device__ __forceinline__ bool step2_vel2(const Bound& b, real tf, real a_max, real a_min, Cand& c) {
    const real vd = b.vf - b.v0;
    const real af = vd / tf;
    double t[7];
    zero_t(t);
    t[1] = tf;
    if (!(a_min - A_EPS < af && af < a_max + A_EPS && check_vel2(t, af, b))) return false;
    set_cand(c, t, af, 0.0, L_ACC0, (af > 0.0) ? DIR_UP : DIR_DOWN, UDDU);
    c.dur = t[1];
    return true;
}

// ---------------------------------------------------------------- position interface, first order
// position_first_step1.rs:20-42
__device__ __forceinline__ bool step1_pos1(const Bound& b, real v_max, real v_min, real bd, BlockRec& blk) {
    const real pd = b.pf - b.p0;
    const real vf = (pd > 0.0) ? v_max : v_min;
    double t[7];
    zero_t(t);
    t[3] = pd / vf;
    blk.has_a = 0; blk.has_b = 0;
    if (!check_pos1(t, vf, b)) return false;
    set_cand(blk.p_min, t, vf, 0.0, L_VEL, (vf > 0.0) ? DIR_UP : DIR_DOWN, UDDU);
    blk.p_min.dur = t[3];  // t_sum[6] of this variant is t[3] (profile.rs:692-702)
    blk.t_min = blk.p_min.dur + bd + 0.0;
    return true;
}

// position_first_step2.rs:22-34
__device__ __forceinline__ bool step2_pos1(const Bound& b, real tf, Cand& c) {
    const real pd = b.pf - b.p0;
    const real vf = pd / tf;
    double t[7];
    zero_t(t);
    t[3] = tf;
    if (!check_pos1(t, vf, b)) return false;
    set_cand(c, t, vf, 0.0, L_NONE, (vf > 0.0) ? DIR_UP : DIR_DOWN, UDDU);
    c.dur = t[3];
    return true;
}

}  // namespace rk
#endif  // RK_H_STEP_MISC_GO
