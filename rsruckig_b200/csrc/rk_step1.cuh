// rk_step1.cuh — Step 1 on the device: extremal profiles of one DoF and the
// block of feasible durations they imply.
//
// Reference: position_third_step1.rs (third-order position interface),
// velocity_third_step1.rs, position_second_step1.rs, velocity_second_step1.rs,
// position_first_step1.rs, block.rs:40-154.
//
// Device design (not the reference's): every thread walks the SAME six-stage
// schedule (NONE/ACC0/ACC1 quartics for both directions, then ACC0_ACC1, then
// the VEL family) regardless of which branch of the reference its inputs take.
// The reference's "vf ~ 0 && af ~ 0" branch visits the same families in another
// order and stops at the first hit; each family is a pure function of the
// inputs, so here those threads run the common schedule and pick the winner by
// the reference's priority afterwards. Candidates are kept as t[7] + codes.
// Included once per pass: rk_kernels.cu compiles these headers twice — into namespace rk (exact divisions with their fall-back
// branch) and, with RK_PASS2 defined and `rk` renamed, into namespace rk_lazy (rk_math.cuh, RK_LAZY_DDIV).
#ifndef RK_PASS2
#ifndef RK_H_STEP1_1
#define RK_H_STEP1_1
#define RK_H_STEP1_GO
#endif
#else
#ifndef RK_H_STEP1_2
#define RK_H_STEP1_2
#define RK_H_STEP1_GO
#endif
#endif
#ifdef RK_H_STEP1_GO
#undef RK_H_STEP1_GO
#include "rk_core.cuh"

namespace rk {

// A validated extremal profile, compact form.
struct Cand {
    double t[7];
    double dur;    // t_sum[6]
    double ctrl;   // jf (third order) | a_up (second order) | v_up (first order)
    double ctrl2;  // a_down (second-order position interface)
    int lim, dir, cs;
};

struct BlockRec {
    double t_min, a_left, a_right, b_left, b_right;
    int has_a, has_b;
    Cand p_min, pa, pb;
};

__device__ __forceinline__ void copy_t(double (&dst)[7], const double (&src)[7]) {
#pragma unroll
    for (int i = 0; i < 7; ++i) dst[i] = src[i];
}

// block.rs:220-239 (brake duration bd is common to all profiles of the DoF; accel duration is always 0, Q21)
__device__ __forceinline__ void make_interval(const Cand& pl, const Cand& pr, real bd, double& left, double& right, Cand& prof) {
    const real ld = pl.dur + bd + 0.0;
    const real rd = pr.dur + bd + 0.0;
    if (ld < rd) { left = ld; right = rd; prof = pr; }
    else { left = rd; right = ld; prof = pl; }
}

// block.rs:40-154 on a list of n valid candidates (n <= 5; entries may be reordered by the n == 4 repair).
__device__ __noinline__ bool calculate_block(BlockRec& blk, Cand* vp, int n, double bd) {
    blk.has_a = 0; blk.has_b = 0;
    if (n == 1) {
        blk.p_min = vp[0];
        blk.t_min = vp[0].dur + bd + 0.0;
        return true;
    } else if (n == 2) {
        if (fabs(vp[0].dur - vp[1].dur) < 8.0 * EPS) {
            blk.p_min = vp[0];
            blk.t_min = vp[0].dur + bd + 0.0;
            return true;
        }
        const int idx_min = (vp[0].dur < vp[1].dur) ? 0 : 1;
        const int idx_else = (idx_min + 1) % 2;
        blk.p_min = vp[idx_min];
        blk.t_min = vp[idx_min].dur + bd + 0.0;
        make_interval(vp[idx_min], vp[idx_else], bd, blk.a_left, blk.a_right, blk.pa);
        blk.has_a = 1;
        return true;
    } else if (n == 4) {
        int rm;
        if (fabs(vp[0].dur - vp[1].dur) < 32.0 * EPS && vp[0].dir != vp[1].dir) {
            rm = 1;
        } else if ((fabs(vp[2].dur - vp[3].dur) < 256.0 * EPS && vp[2].dir != vp[3].dir)
                   || (fabs(vp[0].dur - vp[3].dur) < 256.0 * EPS && vp[0].dir != vp[3].dir)) {
            rm = 3;
        } else {
            return false;
        }
        for (int i = rm; i < 3; ++i) vp[i] = vp[i + 1];
        n = 3;
    } else if (n % 2 == 0) {
        return false;
    }

    int idx_min = 0;
    for (int i = 1; i < n; ++i)
        if (vp[i].dur < vp[idx_min].dur) idx_min = i;
    blk.p_min = vp[idx_min];
    blk.t_min = vp[idx_min].dur + bd + 0.0;

    if (n == 3) {
        make_interval(vp[(idx_min + 1) % 3], vp[(idx_min + 2) % 3], bd, blk.a_left, blk.a_right, blk.pa);
        blk.has_a = 1;
        return true;
    } else if (n == 5) {
        const int e1 = (idx_min + 1) % 5, e2 = (idx_min + 2) % 5, e3 = (idx_min + 3) % 5, e4 = (idx_min + 4) % 5;
        if (vp[e1].dir == vp[e2].dir) {
            make_interval(vp[e1], vp[e2], bd, blk.a_left, blk.a_right, blk.pa);
            make_interval(vp[e3], vp[e4], bd, blk.b_left, blk.b_right, blk.pb);
        } else {
            make_interval(vp[e1], vp[e4], bd, blk.a_left, blk.a_right, blk.pa);
            make_interval(vp[e2], vp[e3], bd, blk.b_left, blk.b_right, blk.pb);
        }
        blk.has_a = 1; blk.has_b = 1;
        return true;
    }
    return false;
}

// ===================================================== third-order position interface
enum : int { QUARTIC_ROOTS = 0, QUARTIC_CANDIDATE = 1 };

// Where validated candidates go: a local list (numerical rescues, zero-limit case) ...
struct LocalSink {
    static constexpr bool inline_check = false;
    static constexpr bool screen_only = false;
    int variant;
    __device__ __forceinline__ void push_root(double) {}
    Cand vp[6];
    int count;
    __device__ __forceinline__ void put(const double (&t)[7], int lim, const Lim& L) {
        Cand& c = vp[count];
        copy_t(c.t, t);
        c.dur = t_sum6(t);
        c.ctrl = L.j_max; c.ctrl2 = 0.0;
        c.lim = lim; c.dir = (L.v_max > 0.0) ? DIR_UP : DIR_DOWN; c.cs = UDDU;
        if (count < 5) ++count;  // the list saturates at five entries (Q11)
    }
};
// Sink of the two-phase quartic kernels: phase A collects in-range roots, phase B stores the validated candidate of ONE
// root into the sub-list slot that root owns (slot = rank of the root, so the ascending order survives).
struct QuarticSink {
    static constexpr bool inline_check = true;
    static constexpr bool screen_only = false;
    int variant;
    double roots[4];
    int n_roots;
    double* slot;  // the 64-byte candidate slot this root owns: t[7] + duration
    int count;
    __device__ __forceinline__ void push_root(double r) {
        if (n_roots == 0) roots[0] = r; else if (n_roots == 1) roots[1] = r; else if (n_roots == 2) roots[2] = r; else roots[3] = r;
        ++n_roots;
    }
    __device__ __forceinline__ void put(const double (&t)[7], int, const Lim&) {
#pragma unroll
        for (int k = 0; k < 7; ++k) slot[k] = t[k];
        slot[7] = t_sum6(t);
        ++count;
    }
};

// Sink of the closed-form families (ACC0_ACC1, VEL) in their candidate phase: `emit` only applies the cheap screens of the
// profile check, parks t[7] in the slot of this variant and notes the slot for the check kernel.
struct CandSink {
    static constexpr bool inline_check = true;
    static constexpr bool screen_only = true;
    __device__ __forceinline__ void push_root(double) {}
    double* base;     // slot 0 of this sub-list for this item (64-byte slots: t[7] + duration)
    int variant;      // slot inside the sub-list the next candidate goes to
    uint32_t parked;  // bit v set: slot v holds a candidate to be checked
    int count;
    __device__ __forceinline__ void put(const double (&t)[7], int, const Lim&) {
        double* q = base + variant * 8;
#pragma unroll
        for (int k = 0; k < 7; ++k) q[k] = t[k];
        q[7] = t_sum6(t);
        parked |= 1u << variant;
    }
};

// the tests of the profile check that need no integration (profile.rs:336-349)
__device__ __forceinline__ bool screen_pos3(const double (&t)[7], int lim) {
    if (!(t[0] >= 0.0) || !(t[1] >= 0.0) || !(t[2] >= 0.0) || !(t[3] >= 0.0) || !(t[4] >= 0.0) || !(t[5] >= 0.0) || !(t[6] >= 0.0)) return false;
    if ((lim == L_ACC0_ACC1_VEL || lim == L_ACC0_VEL || lim == L_ACC1_VEL || lim == L_VEL) && t[3] < EPS) return false;
    if ((lim == L_ACC0 || lim == L_ACC0_ACC1) && t[1] < EPS) return false;
    if ((lim == L_ACC1 || lim == L_ACC0_ACC1) && t[5] < EPS) return false;
    if (t_sum6(t) > T_MAX) return false;
    return true;
}

template <class Sink>
struct Step1Pos3 {
    // inputs (post-brake) and the products the reference precomputes (position_third_step1.rs:39-86)
    Bound b;
    real pd, v0_v0, vf_vf, a0_a0, a0_p3, a0_p4, af_af, af_p3, af_p4;
    Den jj;
    Sink sink;
    bool first_only;  // reference's return_after_found

    __device__ __forceinline__ void init(const Bound& bb, double j_max) {
        b = bb;
        pd = b.pf - b.p0;
        v0_v0 = b.v0 * b.v0; vf_vf = b.vf * b.vf;
        a0_a0 = b.a0 * b.a0; af_af = b.af * b.af;
        a0_p3 = b.a0 * a0_a0; a0_p4 = a0_a0 * a0_a0;
        af_p3 = b.af * af_af; af_p4 = af_af * af_af;
        jj = Den(j_max * j_max);
        sink.count = 0;
    }

    __device__ __forceinline__ void set_variant(int v) { sink.variant = v; }

    // Validate t; on success hand it to the sink. Returns validity.
    __device__ __forceinline__ bool emit(const double (&t)[7], int lim, const Lim& L) {
        if (Sink::screen_only) {  // candidate phase: never "found", every variant is parked for the check kernel
            if (screen_pos3(t, lim)) sink.put(t, lim, L);
            return false;
        }
        if (!(Sink::inline_check ? check_pos3_inline(t, UDDU, lim, L.j_max, L, b) : check_pos3(t, UDDU, lim, L.j_max, L, b))) return false;
        sink.put(t, lim, L);
        return true;
    }

    // position_third_step1.rs:97-240. All four variants are validated as ReachedLimits::Vel (Q6); at most one is kept.
    __device__ __forceinline__ void family_vel(const Lim& L) {
        const real v0 = b.v0, a0 = b.a0, vf = b.vf, af = b.af;
        const Den v_max = L.v_max, a_max = L.a_max, a_min = L.a_min, j_max = L.j_max;
        double t[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        // A variant whose cheap durations are already negative (or NaN: t_acc0 / t_acc1 are square roots) fails the profile
        // check whatever t[3] is, so the long t[3] expression is only evaluated for the others — with NaN operands it
        // would also take the slow path of its division.
        auto rest_ok = [&]() { return t[0] >= 0.0 && t[1] >= 0.0 && t[2] >= 0.0 && t[4] >= 0.0 && t[5] >= 0.0 && t[6] >= 0.0; };
        // ACC0_ACC1_VEL
        t[0] = (-a0 + a_max) / j_max;
        t[1] = (a0_a0 * 0.5 - a_max * a_max - j_max * (v0 - v_max)) / (a_max * j_max);
        t[2] = a_max / j_max;
        t[4] = -a_min / j_max;
        t[5] = -(af_af * 0.5 - a_min * a_min - j_max * (vf - v_max)) / (a_min * j_max);
        t[6] = t[4] + af / j_max;
        if (rest_ok()) {
            t[3] = (3.0 * (a0_p4 * a_min - af_p4 * a_max)
                    + 8.0 * a_max * a_min * (af_p3 - a0_p3 + 3.0 * j_max * (a0 * v0 - af * vf))
                    + 6.0 * a0_a0 * a_min * (a_max * a_max - 2.0 * j_max * v0)
                    - 6.0 * af_af * a_max * (a_min * a_min - 2.0 * j_max * vf)
                    - 12.0 * j_max * (a_max * a_min * (a_max * (v0 + v_max) - a_min * (vf + v_max) - 2.0 * j_max * pd)
                                      + (a_min - a_max) * j_max * v_max * v_max
                                      + j_max * (a_max * vf_vf - a_min * v0_v0)))
                   / (24.0 * a_max * a_min * jj * v_max);
            set_variant(0);
            if (emit(t, L_VEL, L)) return;
        }

        // ACC1_VEL (t[4..6] carry over)
        const real t_acc0 = sqrt(a0_a0 / twice(jj) + (v_max - v0) / j_max);
        t[0] = t_acc0 - a0 / j_max;
        t[1] = 0.0;
        t[2] = t_acc0;
        if (rest_ok()) {
            t[3] = -(3.0 * af_p4
                     - 8.0 * a_min * (af_p3 - a0_p3)
                     - 24.0 * a_min * j_max * (a0 * v0 - af * vf)
                     + 6.0 * af_af * (a_min * a_min - 2.0 * j_max * vf)
                     - 12.0 * j_max * (2.0 * a_min * j_max * pd
                                       + a_min * a_min * (vf + v_max)
                                       + j_max * (v_max * v_max - vf_vf)
                                       + a_min * t_acc0 * (a0_a0 - 2.0 * j_max * (v0 + v_max))))
                   / (24.0 * a_min * jj * v_max);
            set_variant(1);
            if (emit(t, L_VEL, L)) return;
        }

        // ACC0_VEL
        const real t_acc1 = sqrt(af_af / twice(jj) + (v_max - vf) / j_max);
        t[0] = (-a0 + a_max) / j_max;
        t[1] = (a0_a0 * 0.5 - a_max * a_max - j_max * (v0 - v_max)) / (a_max * j_max);
        t[2] = a_max / j_max;
        t[4] = t_acc1;
        t[5] = 0.0;
        t[6] = t_acc1 + af / j_max;
        if (rest_ok()) {
            t[3] = (3.0 * a0_p4
                    + 8.0 * a_max * (af_p3 - a0_p3)
                    + 24.0 * a_max * j_max * (a0 * v0 - af * vf)
                    + 6.0 * a0_a0 * (a_max * a_max - 2.0 * j_max * v0)
                    - 12.0 * j_max * (-2.0 * a_max * j_max * pd
                                      + a_max * a_max * (v0 + v_max)
                                      + j_max * (v_max * v_max - v0_v0)
                                      + a_max * t_acc1 * (-af_af + 2.0 * (vf + v_max) * j_max)))
                   / (24.0 * a_max * jj * v_max);
            set_variant(2);
            if (emit(t, L_VEL, L)) return;
        }

        // VEL (t[4..6] carry over)
        t[0] = t_acc0 - a0 / j_max;
        t[1] = 0.0;
        t[2] = t_acc0;
        if (rest_ok()) {
            t[3] = (af_p3 - a0_p3) / (3.0 * jj * v_max)
                   + (a0 * v0 - af * vf + (af_af * t_acc1 + a0_a0 * t_acc0) * 0.5) / (j_max * v_max)
                   - (v0 / v_max + 1.0) * t_acc0
                   - (vf / v_max + 1.0) * t_acc1
                   + pd / v_max;
            set_variant(3);
            emit(t, L_VEL, L);
        }
    }

    // position_third_step1.rs:241-327
    __device__ __forceinline__ void family_acc0_acc1(const Lim& L) {
        const real v0 = b.v0, a0 = b.a0, vf = b.vf, af = b.af;
        const Den a_max = L.a_max, a_min = L.a_min, j_max = L.j_max;
        real h1 = (3. * (af_p4 * a_max - a0_p4 * a_min)
                     + a_max * a_min * (8. * (a0_p3 - af_p3) + 3. * a_max * a_min * (a_max - a_min) + 6. * a_min * af_af - 6. * a_max * a0_a0)
                     + 12. * j_max * (a_max * a_min * ((a_max - 2. * a0) * v0 - (a_min - 2. * af) * vf) + a_min * a0_a0 * v0 - a_max * af_af * vf))
                    / (3. * (a_max - a_min) * jj)
                    + 4. * (a_max * vf_vf - a_min * v0_v0 - 2. * a_min * a_max * pd) / (a_max - a_min);
        if (!(h1 >= 0.)) return;
        h1 = sqrt(h1) * 0.5;
        const real h2 = a0_a0 / (2. * a_max * j_max) + (a_min - 2. * a_max) / twice(j_max) - v0 / a_max;
        const real h3 = -af_af / (2. * a_min * j_max) - (a_max - 2. * a_min) / twice(j_max) + vf / a_min;
        double t[7];
        // UDDU: Solution 2
        if (h2 > h1 / a_max && h3 > -h1 / a_min) {
            t[0] = (-a0 + a_max) / j_max;
            t[1] = h2 - h1 / a_max;
            t[2] = a_max / j_max;
            t[3] = 0.;
            t[4] = -a_min / j_max;
            t[5] = h3 + h1 / a_min;
            t[6] = t[4] + af / j_max;
            set_variant(0);
            if (emit(t, L_ACC0_ACC1, L) && first_only) return;
        }
        // UDDU: Solution 1
        if (h2 > -h1 / a_max && h3 > h1 / a_min) {
            t[0] = (-a0 + a_max) / j_max;
            t[1] = h2 + h1 / a_max;
            t[2] = a_max / j_max;
            t[3] = 0.;
            t[4] = -a_min / j_max;
            t[5] = h3 - h1 / a_min;
            t[6] = t[4] + af / j_max;
            set_variant(1);
            emit(t, L_ACC0_ACC1, L);
        }
    }

    // position_third_step1.rs:329-601 computes three monic quartics (NONE, ACC0, ACC1) per direction and polishes each
    // root with Newton steps. Here each quartic is its own function (and its own kernel, one thread per direction): the
    // three are independent of one another, and the candidate order NONE, ACC0, ACC1 is restored when the lists are read.
    template <int MODE>
    __device__ __forceinline__ void quartic_none(const Lim& L, double tt_in) {
        const real v0 = b.v0, a0 = b.a0, vf = b.vf, af = b.af;
        const Den a_max = L.a_max, a_min = L.a_min, j_max = L.j_max;
        const real jl = j_max * j_max;  // the reference mixes this local with the member computed from the UP jerk; equal values
        const real h2_none = (a0_a0 - af_af) / twice(j_max) + (vf - v0);
        const real h2_h2 = h2_none * h2_none;
        const real t_min_none = (a0 - af) / j_max;
        const real t_max_none = (a_max - a_min) / j_max;
        const real pn1 = -2.0 * (a0_a0 + af_af - 2.0 * j_max * (v0 + vf)) / (jl);
        const real pn2 = 4.0 * (a0_p3 - af_p3 + 3.0 * j_max * (af * vf - a0 * v0)) / (3.0 * jl * j_max) - 4.0 * pd / j_max;
        const real pn3 = -h2_h2 / (jl);
        double t[7];
        if (MODE == QUARTIC_ROOTS) {  // phase A: closed-form roots, keep the ones inside [t_min, t_max] in ascending order
            const Roots roots = solve_quart_monic(0.0, pn1, pn2, pn3);
#pragma unroll
            for (int ri = 0; ri < 4; ++ri) {
                const double r = roots.get(ri);
                if (ri < roots.n && !(r < t_min_none || r > t_max_none)) sink.push_root(r);
            }
            return;
        }
        {
            {  // phase B: one in-range root -> Newton polish -> t[7] -> profile check
                real tt = tt_in;
                if (tt > EPS) {  // single Newton step (regarding pd)
                    const real h1 = j_max * tt * tt;
                    const real orig = -h2_h2 / (4.0 * j_max * tt) + h2_none * (af / j_max + tt)
                                        + (4.0 * a0_p3 + 2.0 * af_p3 - 6.0 * a0_a0 * (af + 2.0 * j_max * tt) + 12.0 * (af - a0) * j_max * v0
                                           + 3.0 * jj * (-4.0 * pd + (h1 + 8.0 * v0) * tt))
                                          / (12.0 * jj);
                    const real deriv = h2_none + 2.0 * v0 - a0_a0 / j_max + h2_h2 / (4.0 * h1) + (3.0 * h1) * 0.25;
                    tt -= orig / deriv;
                }
                const real h0 = h2_none / (2.0 * j_max * tt);
                t[0] = h0 + tt * 0.5 - a0 / j_max;
                t[1] = 0.0;
                t[2] = tt;
                t[3] = 0.0; t[4] = 0.0; t[5] = 0.0;
                t[6] = -h0 + tt * 0.5 + af / j_max;
                emit(t, L_NONE, L);
            }
        }
    }

    template <int MODE>
    __device__ __forceinline__ void quartic_acc0(const Lim& L, double tt_in) {
        const real v0 = b.v0, a0 = b.a0, vf = b.vf, af = b.af;
        const Den a_max = L.a_max, a_min = L.a_min, j_max = L.j_max;
        const real jl = j_max * j_max;  // the reference mixes this local with the member computed from the UP jerk; equal values
        // ACC0
        const real h3_acc0 = (a0_a0 - af_af) / (2.0 * a_max * j_max) + (vf - v0) / a_max;
        const real t_min_acc0 = (a_max - af) / j_max;
        const real t_max_acc0 = (a_max - a_min) / j_max;
        const real h0_acc0 = 3.0 * (af_p4 - a0_p4)
                               + 8.0 * (a0_p3 - af_p3) * a_max
                               + 24.0 * a_max * j_max * (af * vf - a0 * v0)
                               - 6.0 * a0_a0 * (a_max * a_max - 2.0 * j_max * v0)
                               + 6.0 * af_af * (a_max * a_max - 2.0 * j_max * vf)
                               + 12.0 * j_max * (j_max * (vf_vf - v0_v0 - 2.0 * a_max * pd) - a_max * a_max * (vf - v0));
        const real h2_acc0 = -af_af + a_max * a_max + 2.0 * j_max * vf;
        const real pa0 = -2.0 * a_max / j_max;
        const real pa1 = h2_acc0 / (jl);
        const real pa2 = 0.0;
        const real pa3 = h0_acc0 / (12.0 * jj * jj);
        // sign pre-screen of the quartic shifted to t_min_acc0 (:404-415)
        const real m0 = pa0 + 4.0 * t_min_acc0;
        const real m1 = pa1 + (3.0 * pa0 + 6.0 * t_min_acc0) * t_min_acc0;
        const real m2 = pa2 + (2.0 * pa1 + (3.0 * pa0 + 4.0 * t_min_acc0) * t_min_acc0) * t_min_acc0;
        const real m3 = pa3 + (pa2 + (pa1 + (pa0 + t_min_acc0) * t_min_acc0) * t_min_acc0) * t_min_acc0;
        const bool acc0_has_solution = (m0 < 0.0) || (m1 < 0.0) || (m2 < 0.0) || (m3 <= 0.0);
        if (MODE == QUARTIC_ROOTS && !acc0_has_solution) return;
        double t[7];
        if (MODE == QUARTIC_ROOTS) {  // phase A: closed-form roots, keep the ones inside [t_min, t_max] in ascending order
            const Roots roots = solve_quart_monic(pa0, pa1, pa2, pa3);
#pragma unroll
            for (int ri = 0; ri < 4; ++ri) {
                const double r = roots.get(ri);
                if (ri < roots.n && !(r < t_min_acc0 || r > t_max_acc0)) sink.push_root(r);
            }
            return;
        }
        {
            {  // phase B: one in-range root -> Newton polish -> t[7] -> profile check
                real tt = tt_in;
                if (tt > EPS) {  // single Newton step (regarding pd)
                    const real h1 = j_max * tt;
                    const real orig = h0_acc0 / (12.0 * jj * tt) + tt * (h2_acc0 + h1 * (h1 - 2.0 * a_max));
                    const real deriv = 2.0 * (h2_acc0 + h1 * (2.0 * h1 - 3.0 * a_max));
                    tt -= orig / deriv;
                }
                t[0] = (-a0 + a_max) / j_max;
                t[1] = h3_acc0 - 2.0 * tt + (j_max / a_max) * tt * tt;
                t[2] = tt;
                t[3] = 0.0; t[4] = 0.0; t[5] = 0.0;
                t[6] = (af - a_max) / j_max + tt;
                emit(t, L_ACC0, L);
            }
        }
    }

    template <int MODE>
    __device__ __forceinline__ void quartic_acc1(const Lim& L, double tt_in) {
        const real v0 = b.v0, a0 = b.a0, vf = b.vf, af = b.af;
        const Den a_max = L.a_max, a_min = L.a_min, j_max = L.j_max;
        const real jl = j_max * j_max;  // the reference mixes this local with the member computed from the UP jerk; equal values
        // ACC1
        const real h3_acc1 = -(a0_a0 + af_af) / (2.0 * j_max * a_min) + a_min / j_max + (vf - v0) / a_min;
        const real t_min_acc1 = (a_min - a0) / j_max;
        const real t_max_acc1 = (a_max - a0) / j_max;
        const real h0_acc1 = (a0_p4 - af_p4) * 0.25
                               + 2.0 * (af_p3 - a0_p3) * a_min / THREE
                               + (a0_a0 - af_af) * a_min * a_min * 0.5
                               + j_max * (af_af * vf + a0_a0 * v0 + 2.0 * a_min * (j_max * pd - a0 * v0 - af * vf)
                                          + a_min * a_min * (v0 + vf) + j_max * (v0_v0 - vf_vf));
        const real h2_acc1 = a0_a0 - a0 * a_min + 2.0 * j_max * v0;
        const real pb0 = 2.0 * (2.0 * a0 - a_min) / j_max;
        const real pb1 = (5.0 * a0_a0 + a_min * (a_min - 6.0 * a0) + 2.0 * j_max * v0) / (jl);
        const real pb2 = 2.0 * (a0 - a_min) * h2_acc1 / (jl * j_max);
        const real pb3 = h0_acc1 / (jj * jj);
        const bool acc1_has_solution = (pb0 < 0.0) || (pb1 < 0.0) || (pb2 < 0.0) || (pb3 <= 0.0);
        if (MODE == QUARTIC_ROOTS && !acc1_has_solution) return;
        double t[7];
        if (MODE == QUARTIC_ROOTS) {  // phase A: closed-form roots, keep the ones inside [t_min, t_max] in ascending order
            const Roots roots = solve_quart_monic(pb0, pb1, pb2, pb3);
#pragma unroll
            for (int ri = 0; ri < 4; ++ri) {
                const double r = roots.get(ri);
                if (ri < roots.n && !(r < t_min_acc1 || r > t_max_acc1)) sink.push_root(r);
            }
            return;
        }
        {
            {  // phase B: one in-range root -> Newton polish -> t[7] -> profile check
                real tt = tt_in;
                if (tt > EPS) {  // up to three Newton steps (regarding pd)
                    const real h5 = a0_p3 + 2.0 * j_max * a0 * v0;
                    real h1 = j_max * tt;
                    real orig = -(h0_acc1 * 0.5
                                    + h1 * (h5 + a0 * (a_min - 2.0 * h1) * (a_min - h1) + a0_a0 * (5.0 * h1 * 0.5 - 2.0 * a_min)
                                            + a_min * a_min * h1 * 0.5 + j_max * (h1 * 0.5 - a_min) * (h1 * tt + 2.0 * v0)))
                                  / j_max;
                    real deriv = (a_min - a0 - h1) * (h2_acc1 + h1 * (4.0 * a0 - a_min + 2.0 * h1));
                    tt -= rmin(orig / deriv, tt);

                    h1 = j_max * tt;
                    orig = -(h0_acc1 * 0.5
                             + h1 * (h5 + a0 * (a_min - 2.0 * h1) * (a_min - h1) + a0_a0 * (5.0 * h1 * 0.5 - 2.0 * a_min)
                                     + a_min * a_min * h1 * 0.5 + j_max * (h1 * 0.5 - a_min) * (h1 * tt + 2.0 * v0)))
                           / j_max;
                    if (fabs(orig) > 1e-9) {
                        deriv = (a_min - a0 - h1) * (h2_acc1 + h1 * (4.0 * a0 - a_min + 2.0 * h1));
                        tt -= orig / deriv;

                        h1 = j_max * tt;
                        orig = -(h0_acc1 * 0.5
                                 + h1 * (h5 + a0 * (a_min - 2.0 * h1) * (a_min - h1) + a0_a0 * (5.0 * h1 * 0.5 - 2.0 * a_min)
                                         + a_min * a_min * h1 * 0.5 + j_max * (h1 * 0.5 - a_min) * (h1 * tt + 2.0 * v0)))
                               / j_max;
                        if (fabs(orig) > 1e-9) {
                            deriv = (a_min - a0 - h1) * (h2_acc1 + h1 * (4.0 * a0 - a_min + 2.0 * h1));
                            tt -= orig / deriv;
                        }
                    }
                }
                t[0] = tt;
                t[1] = 0.0;
                t[2] = (a0 - a_min) / j_max + tt;
                t[3] = 0.0; t[4] = 0.0;
                t[5] = h3_acc1 - (2.0 * a0 + j_max * tt) * tt / a_min;
                t[6] = (af - a_min) / j_max;
                emit(t, L_ACC1, L);
            }
        }
    }


    // ---- rare numerical rescues (position_third_step1.rs:603-895); each stops at its first valid candidate
    __device__ __forceinline__ void rescue_none(const Lim& L) {  // :843-895
        const Den j_max = L.j_max;
        double t[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const real h0 = sqrt((a0_a0 + af_af) * 0.5 + j_max * (b.vf - b.v0)) * fabs(j_max) / j_max;
        t[0] = (h0 - b.a0) / j_max;
        t[2] = (h0 - b.af) / j_max;
        if (emit(t, L_NONE, L)) return;
        t[0] = (b.af - b.a0) / j_max;
        t[2] = 0.0;
        emit(t, L_NONE, L);
    }

    __device__ __forceinline__ void rescue_acc0(const Lim& L) {  // :644-778
        const real v0 = b.v0, a0 = b.a0, vf = b.vf, af = b.af;
        const Den a_max = L.a_max, a_min = L.a_min, j_max = L.j_max;
        double t[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        // two step
        t[1] = (af_af - a0_a0 + 2.0 * j_max * (vf - v0)) / (2.0 * a0 * j_max);
        t[2] = (a0 - af) / j_max;
        if (emit(t, L_ACC0, L)) return;
        // three step, pf removed
        t[0] = (-a0 + a_max) / j_max;
        t[1] = (a0_a0 + af_af - 2.0 * a_max * a_max + 2.0 * j_max * (vf - v0)) / (2.0 * a_max * j_max);
        t[2] = (-af + a_max) / j_max;
        if (emit(t, L_ACC0, L)) return;
        // three step, a_max removed
        {
            const real h0 = 3.0 * (af_af - a0_a0 + 2.0 * j_max * (v0 + vf));
            const real h2 = a0_p3 + 2.0 * af_p3 + 6.0 * jj * pd + 6.0 * (af - a0) * j_max * vf - 3.0 * a0 * af_af;
            const real h1 = sqrt(2.0 * (2.0 * h2 * h2
                                          + h0 * (a0_p4 - 6.0 * a0_a0 * (af_af + 2.0 * j_max * vf)
                                                  + 8.0 * a0 * (af_p3 + 3.0 * jj * pd + 3.0 * af * j_max * vf)
                                                  - 3.0 * (af_p4 + 4.0 * af_af * j_max * vf + 4.0 * jj * (vf_vf - v0_v0)))))
                              * fabs(j_max) / j_max;
            t[0] = (4.0 * af_p3 + 2.0 * a0_p3 - 6.0 * a0 * af_af + 12.0 * jj * pd + 12.0 * (af - a0) * j_max * vf + h1) / (2.0 * j_max * h0);
            t[1] = -h1 / (j_max * h0);
            t[2] = (-4.0 * a0_p3 - 2.0 * af_p3 + 6.0 * a0_a0 * af + 12.0 * jj * pd - 12.0 * (af - a0) * j_max * v0 + h1) / (2.0 * j_max * h0);
            if (emit(t, L_ACC0, L)) return;
        }
        // three step with t = (a_max - a_min) / j_max
        {
            const real tt = (a_max - a_min) / j_max;
            t[0] = (-a0 + a_max) / j_max;
            t[1] = (a0_a0 - af_af) / (2.0 * a_max * j_max) + (vf - v0 + j_max * tt * tt) / a_max - 2.0 * tt;
            t[2] = tt;
            t[6] = (af - a_min) / j_max;
            emit(t, L_ACC0, L);
        }
    }

    __device__ __forceinline__ void rescue_vel(const Lim& L) {  // :780-841
        const real v0 = b.v0, a0 = b.a0, vf = b.vf, af = b.af;
        const Den v_max = L.v_max, j_max = L.j_max;
        const real h1 = sqrt(af_af / twice(jj) + (v_max - vf) / j_max);
        double t[7];
        t[0] = -a0 / j_max;
        t[1] = 0.0;
        t[2] = 0.0;
        t[3] = (af_p3 - a0_p3) / (3.0 * jj * v_max) + (a0 * v0 - af * vf + (af_af * h1) * 0.5) / (j_max * v_max) - (vf / v_max + 1.0) * h1 + pd / v_max;
        t[4] = h1;
        t[5] = 0.0;
        t[6] = h1 + af / j_max;
        if (emit(t, L_VEL, L)) return;
        t[0] = 0.0;
        t[1] = 0.0;
        t[2] = a0 / j_max;
        t[3] = (af_p3 - a0_p3) / (3.0 * jj * v_max) + (a0 * v0 - af * vf + (af_af * h1 + a0_p3 / j_max) * 0.5) / (j_max * v_max)
               - (v0 / v_max + 1.0) * a0 / j_max - (vf / v_max + 1.0) * h1 + pd / v_max;
        t[4] = h1;
        t[5] = 0.0;
        t[6] = h1 + af / j_max;
        emit(t, L_VEL, L);
    }

    __device__ __forceinline__ void rescue_acc1_vel(const Lim& L) {  // :603-642
        const real v0 = b.v0, a0 = b.a0, vf = b.vf, af = b.af;
        const Den v_max = L.v_max, a_min = L.a_min, j_max = L.j_max;
        double t[7];
        t[0] = 0.0;
        t[1] = 0.0;
        t[2] = a0 / j_max;
        t[3] = -(3.0 * af_p4 - 8.0 * a_min * (af_p3 - a0_p3) - 24.0 * a_min * j_max * (a0 * v0 - af * vf)
                 + 6.0 * af_af * (a_min * a_min - 2.0 * j_max * vf)
                 - 12.0 * j_max * (2. * a_min * j_max * pd + a_min * a_min * (vf + v_max) + j_max * (v_max * v_max - vf_vf)
                                   + a_min * a0 * (a0_a0 - 2.0 * j_max * (v0 + v_max)) / j_max))
               / (24.0 * a_min * jj * v_max);
        t[4] = -a_min / j_max;
        t[5] = -(af_af * 0.5 - a_min * a_min + j_max * (v_max - vf)) / (a_min * j_max);
        t[6] = t[4] + af / j_max;
        emit(t, L_ACC1_VEL, L);
    }
};

__device__ __forceinline__ Lim lim_dir(bool up, double v_max, double v_min, double a_max, double a_min, double j_max) {
    Lim L;
    L.v_max = Den(up ? v_max : v_min); L.v_min = Den(up ? v_min : v_max);
    L.a_max = Den(up ? a_max : a_min); L.a_min = Den(up ? a_min : a_max);
    L.j_max = Den(up ? j_max : -j_max);
    return L;
}

// The zero-limit special case of position_third_step1.rs:982-1004 (+ time_all_single_step :897-980): one
// constant-acceleration phase. bd = brake duration. Returns found; fills blk. (The general case runs as the kernel
// pipeline k1_family<0..2> + k1_block.)
__device__ __noinline__ bool step1_pos3_zero_limits(const Bound& b, real v_max, real v_min, real a_max, real a_min, real j_max, real bd, BlockRec& blk) {
    Step1Pos3<LocalSink> s;
    s.init(b, j_max);
    blk.has_a = 0; blk.has_b = 0;
    {
        if (fabs(b.af - b.a0) > EPS) return false;
        const Lim L = lim_dir(true, v_max, v_min, a_max, a_min, j_max);
        double t[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        bool ok = false;
        if (fabs(b.a0) > EPS) {
            const real q = sqrt(2.0 * b.a0 * s.pd + s.v0_v0);
            t[3] = (-b.v0 + q) / b.a0;
            ok = t[3] >= 0.0 && check_pos3(t, UDDU, L_NONE, 0.0, L, b);
            if (!ok) {
                t[3] = -(b.v0 + q) / b.a0;
                ok = t[3] >= 0.0 && check_pos3(t, UDDU, L_NONE, 0.0, L, b);
            }
        } else if (fabs(b.v0) > EPS) {
            t[3] = s.pd / b.v0;
            ok = check_pos3(t, UDDU, L_NONE, 0.0, L, b);
        } else if (fabs(s.pd) < EPS) {
            ok = check_pos3(t, UDDU, L_NONE, 0.0, L, b);
        }
        if (!ok) return false;
        Cand& c = blk.p_min;
        copy_t(c.t, t);
        c.dur = t_sum6(t); c.ctrl = 0.0; c.ctrl2 = 0.0; c.lim = L_NONE; c.dir = (v_max > 0.0) ? DIR_UP : DIR_DOWN; c.cs = UDDU;
        blk.t_min = c.dur + bd + 0.0;
        if (fabs(b.v0) > EPS || fabs(b.a0) > EPS) {
            blk.has_a = 1; blk.a_left = blk.t_min; blk.a_right = RK_INF;
            blk.pa = c;  // reference: Interval::default() profile; never selected because a.right is infinite
        }
        return true;
    }
}

}  // namespace rk
#endif  // RK_H_STEP1_GO
