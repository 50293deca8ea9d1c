// rk_step2_vel.cuh — Step 2 family VEL (position_third_step2.rs:607-974): the
// profile that cruises at v_max for part of tf. By far the most frequent
// winner of Step 2 on random inputs, and the only family that needs the roots
// of a 5th (UDDU) / 6th (UDUD) order polynomial: the extrema of the polynomial
// are found in closed form from its quartic (second) derivative, each sign
// change between consecutive extrema is then bracketed and polished by the
// safeguarded Newton iteration (shrink_interval).
// Included once per pass: rk_kernels.cu compiles these headers twice — into namespace rk (exact divisions with their fall-back
// branch) and, with RK_PASS2 defined and `rk` renamed, into namespace rk_lazy (rk_math.cuh, RK_LAZY_DDIV).
#ifndef RK_PASS2
#ifndef RK_H_STEP2_VEL_1
#define RK_H_STEP2_VEL_1
#define RK_H_STEP2_VEL_GO
#endif
#else
#ifndef RK_H_STEP2_VEL_2
#define RK_H_STEP2_VEL_2
#define RK_H_STEP2_VEL_GO
#endif
#endif
#ifdef RK_H_STEP2_VEL_GO
#undef RK_H_STEP2_VEL_GO
#include "rk_step2.cuh"

namespace rk {

constexpr double ROOT_TOLERANCE = 1e-14;  // roots.rs TOLERANCE

// check_root closure of the UDDU branch (:715-776)
__device__ __forceinline__ bool vel_uddu_root(const S2& s, const Lim& L, Hit& h, double tt) {
    RK_S2_LOCALS
    // Single Newton step (regarding pd)
    {
        const real h1 = sqrt((a0_a0 + af_af) / twice(j_max_j_max) + (2.0 * a0 * tt + j_max * tt * tt - vd) / j_max);
        if (h1 == h1) {  // a NaN h1 makes orig NaN and the reference's guard below skips the step: skip the arithmetic as well
            const real orig = -pd
                                - (2.0 * a0_p3 + 4.0 * af_p3 + 24.0 * a0 * j_max * tt * (af + j_max * (h1 + tt - tf)) + 6.0 * a0_a0 * (af + j_max * (2.0 * tt - tf))
                                   + 6.0 * (a0_a0 + af_af) * j_max * h1 + 12.0 * af * j_max * (j_max * tt * tt - vd)
                                   + 12.0 * j_max_j_max * (j_max * tt * tt * (h1 + tt - tf) - tf * v0 - h1 * vd))
                                      / (12.0 * j_max_j_max);
            const real deriv_newton = -(a0 + j_max * tt) * (3.0 * (h1 + tt) - 2.0 * tf + (a0 + 2.0 * af) / j_max);
            if (!(orig != orig) && !(deriv_newton != deriv_newton) && fabs(deriv_newton) > EPS) tt -= orig / deriv_newton;
        }
    }
    if (tt > tf || tt != tt) return false;

    const real h1 = sqrt((a0_a0 + af_af) / twice(j_max_j_max) + (tt * (2.0 * a0 + j_max * tt) - vd) / j_max);
    if (h1 != h1) return false;  // t[4] = NaN fails the profile check
    t[0] = tt;
    t[1] = 0.0;
    t[2] = tt + a0 / j_max;
    t[3] = tf - 2.0 * (tt + h1) - (a0 + af) / j_max;
    t[4] = h1;
    t[5] = 0.0;
    t[6] = h1 + af / j_max;
    RK_TRY_INLINE(UDDU, L_VEL, j_max)
    return false;
}

// check_root closure of the UDUD branch (:883-945)
__device__ __forceinline__ bool vel_udud_root(const S2& s, const Lim& L, Hit& h, double tt) {
    RK_S2_LOCALS
    // Double Newton step (regarding pd)
    {
        real h1 = sqrt((af_af - a0_a0) / twice(j_max_j_max) - ((2.0 * a0 + j_max * tt) * tt - vd) / j_max);
        if (h1 != h1) return false;  // NaN poisons tt and with it every duration: the check fails
        real orig = -pd + (af_p3 - a0_p3 + 3.0 * a0_a0 * j_max * (tf - 2.0 * tt)) / (6.0 * j_max_j_max) + (2.0 * a0 + j_max * tt) * tt * (tf - tt)
                      + (j_max * h1 - af) * h1 * h1 + tf * v0;
        real deriv_newton = (a0 + j_max * tt) * (2.0 * (af + j_max * tf) - 3.0 * j_max * (h1 + tt) - a0) / j_max;
        tt -= orig / deriv_newton;

        h1 = sqrt((af_af - a0_a0) / twice(j_max_j_max) - ((2.0 * a0 + j_max * tt) * tt - vd) / j_max);
        orig = -pd + (af_p3 - a0_p3 + 3.0 * a0_a0 * j_max * (tf - 2.0 * tt)) / (6.0 * j_max_j_max) + (2.0 * a0 + j_max * tt) * tt * (tf - tt)
               + (j_max * h1 - af) * h1 * h1 + tf * v0;
        if (fabs(orig) > 1e-9) {
            deriv_newton = (a0 + j_max * tt) * (2.0 * (af + j_max * tf) - 3.0 * j_max * (h1 + tt) - a0) / j_max;
            tt -= orig / deriv_newton;
        }
    }
    const real h1 = sqrt((af_af - a0_a0) / twice(j_max_j_max) - ((2.0 * a0 + j_max * tt) * tt - vd) / j_max);
    t[0] = tt;
    t[1] = 0.0;
    t[2] = tt + a0 / j_max;
    t[3] = tf - 2.0 * (tt + h1) + ad / j_max;
    t[4] = h1;
    t[5] = 0.0;
    t[6] = h1 - af / j_max;
    RK_TRY_INLINE(UDUD, L_VEL, j_max)
    return false;
}

// UDDU half of time_vel (:619-824)
__device__ RK_S2_FN bool s2_vel_uddu(const S2& s, const Lim& L, Hit& h) {
    RK_S2_LOCALS
    const real a0_p6 = s.a0_p6, af_p6 = s.af_p6;
    const real tz_min = rmax(0.0, -a0 / j_max);
    const real tz_max = rmin((tf - a0 / j_max) * 0.5, (a_max - a0) / j_max);

    if (fabs(v0) < EPS && fabs(a0) < EPS && fabs(vf) < EPS && fabs(af) < EPS) {
        const Roots roots = solve_cub(1.0, -tf * 0.5, 0.0, pd / twice(j_max));
#pragma unroll 1
        for (int ri = 0; ri < roots.n; ++ri) {
            real tt = roots.get(ri);
            if (tt > tf * 0.25) continue;
            // Single Newton step (regarding pd)
            if (tt > EPS) {
                const real orig = -pd + j_max * tt * tt * (tf - 2.0 * tt);
                const real deriv = 2.0 * j_max * tt * (tf - 3.0 * tt);
                tt -= orig / deriv;
            }
            t[0] = tt;
            t[1] = 0.0;
            t[2] = tt;
            t[3] = tf - 4.0 * tt;
            t[4] = tt;
            t[5] = 0.0;
            t[6] = tt;
            RK_TRY(UDDU, L_VEL, j_max)
        }
        return false;
    }

    const real p1 = af_af - 2.0 * j_max * (-2.0 * af * tf + j_max * tf_tf + 3.0 * vd);
    const real ph1 = af_p3 - 3.0 * j_max_j_max * g1 - 3.0 * af * j_max * vd;
    const real ph2 = af_p4 + 8.0 * af_p3 * j_max * tf
                       + 12.0 * j_max * (3.0 * j_max * vd_vd - af_af * vd + 2.0 * af * j_max * (g1 - tf * vd) - 2.0 * j_max_j_max * tf * g1);
    const real ph3 = a0 * (af - j_max * tf);
    const real ph4 = j_max * (-ad + j_max * tf);

    // 5th order polynomial, monic
    double pol[6];
    pol[0] = 1.0;
    pol[1] = (15.0 * a0_a0 + af_af + 4.0 * af * j_max * tf - 16.0 * ph3 - 2.0 * j_max * (j_max * tf_tf + 3.0 * vd)) / (4.0 * ph4);
    pol[2] = (29.0 * a0_p3 - 2.0 * af_p3 - 33.0 * a0 * ph3 + 6.0 * j_max_j_max * g1 + 6.0 * af * j_max * vd + 6.0 * a0 * p1) / (6.0 * j_max * ph4);
    pol[3] = (61.0 * a0_p4 - 76.0 * a0_a0 * ph3 - 16.0 * a0 * ph1 + 30.0 * a0_a0 * p1 + ph2) / (24.0 * j_max_j_max * ph4);
    pol[4] = (a0 * (7.0 * a0_p4 - 10.0 * a0_a0 * ph3 - 4.0 * a0 * ph1 + 6.0 * a0_a0 * p1 + ph2)) / (12.0 * j_max_j_max * j_max * ph4);
    pol[5] = (7.0 * a0_p6 + af_p6 - 12.0 * a0_p4 * ph3 + 48.0 * af_p3 * j_max_j_max * g1 - 8.0 * a0_p3 * ph1
              - 72.0 * j_max_j_max * j_max * (j_max * g1 * g1 + vd_vd * vd + 2.0 * af * g1 * vd)
              - 6.0 * af_p4 * j_max * vd + 36.0 * af_af * j_max_j_max * vd_vd + 9.0 * a0_p4 * p1 + 3.0 * a0_a0 * ph2)
             / (144.0 * j_max_j_max * j_max_j_max * ph4);

    double deriv[5], dderiv[4];
    poly_monic_deri<6>(pol, deriv);
    poly_deri<5>(deriv, dderiv);

    // extrema of the quintic: roots of its quartic derivative, closed form
    const Roots d_extremas = solve_quart_monic(deriv[1], deriv[2], deriv[3], deriv[4]);

    // Two phases instead of the reference's single loop (:778-823), with the same candidates in the same order:
    //  1. scan the extrema (plus a last pass for the interval that ends at tz_max) and only RECORD where a root candidate
    //     sits — directly at an extremum, or bracketed between the previous extremum and this one. Cheap, and all lanes of
    //     a warp walk it together.
    //  2. polish and validate the recorded candidates in order. Almost every item has its (usually single) candidate in
    //     slot 0, so the expensive part — safeguarded Newton bracketing and the profile check — runs with nearly full
    //     warps, where the interleaved loop ran it with the quarter of the lanes that had a candidate at that extremum.
    double c_lo[6], c_hi[6];  // bracket [lo, hi]; a direct candidate has lo == hi
    int n_cand = 0;
    real tz_current = tz_min;
    real val_current = poly_eval<6>(pol, tz_current);
#pragma unroll 1
    for (int ri = 0; ri <= d_extremas.n; ++ri) {
        const bool last = ri == d_extremas.n;
        real tz = last ? tz_max : real(d_extremas.get(ri));
        if (!last && tz >= tz_max) continue;
        bool direct = false, bracket = false;
        real val_new;
        if (!last) {
            const real orig = poly_eval<5>(deriv, tz);
            if (fabs(orig) > ROOT_TOLERANCE) tz -= orig / poly_eval<4>(dderiv, tz);
            val_new = poly_eval<6>(pol, tz);
            if (fabs(val_new) < 64.0 * fabs(poly_eval<4>(dderiv, tz)) * ROOT_TOLERANCE) direct = true;
            else if (val_current * val_new < 0.0) bracket = true;
        } else {
            val_new = poly_eval<6>(pol, tz_max);
            if (val_current * val_new < 0.0) bracket = true;
            else if (fabs(val_new) < 8.0 * EPS) direct = true;
        }
        if (direct || bracket) {
            c_lo[n_cand] = bracket ? (double)tz_current : (double)tz;
            c_hi[n_cand] = tz;
            ++n_cand;
        }
        tz_current = tz;
        val_current = val_new;
    }
#pragma unroll 1
    for (int c = 0; c < n_cand; ++c) {
        const double lo = c_lo[c], hi = c_hi[c];
        const double root = (lo != hi) ? shrink_interval<6>(pol, lo, hi) : hi;
        if (vel_uddu_root(s, L, h, root)) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------------------------
// Lockstep form of the UDDU half: the same arithmetic, cut into phases separated by block barriers so that all warps of a
// block walk the instruction stream together. The hot path of this family is ~50 KB of SASS, more than the SM's
// instruction cache holds, and without the barriers every warp streams it from L2 on its own (the "no instruction" stall
// was half of all warp time); in lockstep a line fetched for one warp is used by the others while it is still cached.
// Every thread of the block must call this (inactive ones with active = false).
#ifndef RK_PHASE_BARRIER
#define RK_PHASE_BARRIER() __syncthreads()
#endif
// Phases A-C of the UDDU half: the quintic, its extrema and the root candidates (brackets [lo, hi], lo == hi for a candidate
// that sits directly at an extremum). The rare all-zero boundary case is solved on the spot (returns true, h filled).
__device__ __forceinline__ bool vel_uddu_candidates(const S2& s, const Lim& L, Hit& h, bool active, double (&pol)[6], double (&c_lo)[6], double (&c_hi)[6],
                                                    int& n_cand) {
    RK_S2_LOCALS
    const real a0_p6 = s.a0_p6, af_p6 = s.af_p6;
    bool found = false;
    real tz_min = 0.0, tz_max = 0.0;
    double deriv[5], dderiv[4];
    n_cand = 0;

    if (active && fabs(v0) < EPS && fabs(a0) < EPS && fabs(vf) < EPS && fabs(af) < EPS) {  // rare, handled out of step
        active = false;
        const Roots roots = solve_cub(1.0, -tf * 0.5, 0.0, pd / twice(j_max));
#pragma unroll 1
        for (int ri = 0; ri < roots.n && !found; ++ri) {
            real tt = roots.get(ri);
            if (tt > tf * 0.25) continue;
            if (tt > EPS) {
                const real orig = -pd + j_max * tt * tt * (tf - 2.0 * tt);
                const real deriv_ = 2.0 * j_max * tt * (tf - 3.0 * tt);
                tt -= orig / deriv_;
            }
            t[0] = tt; t[1] = 0.0; t[2] = tt; t[3] = tf - 4.0 * tt; t[4] = tt; t[5] = 0.0; t[6] = tt;
            if (check_pos3(t, UDDU, L_VEL, j_max, L, s.b)) {
                h.jf = (double)(j_max); h.lim = L_VEL; h.cs = UDDU;
                h.dir = (L.v_max > 0.0) ? DIR_UP : DIR_DOWN;
                found = true;
            }
        }
    }

    // phase A: the monic quintic and its derivatives
    if (active) {
        tz_min = rmax(0.0, -a0 / j_max);
        tz_max = rmin((tf - a0 / j_max) * 0.5, (a_max - a0) / j_max);
        const real p1 = af_af - 2.0 * j_max * (-2.0 * af * tf + j_max * tf_tf + 3.0 * vd);
        const real ph1 = af_p3 - 3.0 * j_max_j_max * g1 - 3.0 * af * j_max * vd;
        const real ph2 = af_p4 + 8.0 * af_p3 * j_max * tf
                           + 12.0 * j_max * (3.0 * j_max * vd_vd - af_af * vd + 2.0 * af * j_max * (g1 - tf * vd) - 2.0 * j_max_j_max * tf * g1);
        const real ph3 = a0 * (af - j_max * tf);
        const real ph4 = j_max * (-ad + j_max * tf);
        pol[0] = 1.0;
        pol[1] = (15.0 * a0_a0 + af_af + 4.0 * af * j_max * tf - 16.0 * ph3 - 2.0 * j_max * (j_max * tf_tf + 3.0 * vd)) / (4.0 * ph4);
        pol[2] = (29.0 * a0_p3 - 2.0 * af_p3 - 33.0 * a0 * ph3 + 6.0 * j_max_j_max * g1 + 6.0 * af * j_max * vd + 6.0 * a0 * p1) / (6.0 * j_max * ph4);
        pol[3] = (61.0 * a0_p4 - 76.0 * a0_a0 * ph3 - 16.0 * a0 * ph1 + 30.0 * a0_a0 * p1 + ph2) / (24.0 * j_max_j_max * ph4);
        pol[4] = (a0 * (7.0 * a0_p4 - 10.0 * a0_a0 * ph3 - 4.0 * a0 * ph1 + 6.0 * a0_a0 * p1 + ph2)) / (12.0 * j_max_j_max * j_max * ph4);
        pol[5] = (7.0 * a0_p6 + af_p6 - 12.0 * a0_p4 * ph3 + 48.0 * af_p3 * j_max_j_max * g1 - 8.0 * a0_p3 * ph1
                  - 72.0 * j_max_j_max * j_max * (j_max * g1 * g1 + vd_vd * vd + 2.0 * af * g1 * vd)
                  - 6.0 * af_p4 * j_max * vd + 36.0 * af_af * j_max_j_max * vd_vd + 9.0 * a0_p4 * p1 + 3.0 * a0_a0 * ph2)
                 / (144.0 * j_max_j_max * j_max_j_max * ph4);
        poly_monic_deri<6>(pol, deriv);
        poly_deri<5>(deriv, dderiv);
    }
    RK_PHASE_BARRIER();

    // phase B: extrema of the quintic in closed form
    Roots d_extremas;
    if (active) d_extremas = solve_quart_monic(deriv[1], deriv[2], deriv[3], deriv[4]);
    RK_PHASE_BARRIER();

    // phase C: record the root candidates (see s2_vel_uddu)
    if (active) {
        real tz_current = tz_min;
        real val_current = poly_eval<6>(pol, tz_current);
#pragma unroll 1
        for (int ri = 0; ri <= d_extremas.n; ++ri) {
            const bool last = ri == d_extremas.n;
            real tz = last ? tz_max : real(d_extremas.get(ri));
            if (!last && tz >= tz_max) continue;
            bool direct = false, bracket = false;
            real val_new;
            if (!last) {
                {   // polynomials evaluated at the same point share its powers (PolyPoint, rk_core.cuh)
                    const PolyPoint<5> at(tz);
                    const real orig = at.eval<5>(deriv);
                    if (fabs(orig) > ROOT_TOLERANCE) tz -= orig / at.eval<4>(dderiv);
                }
                const PolyPoint<6> at(tz);
                double v_new, dd_new;
                at.eval2<6, 4>(pol, dderiv, v_new, dd_new);
                val_new = v_new;
                if (fabs(val_new) < 64.0 * fabs(dd_new) * ROOT_TOLERANCE) direct = true;
                else if (val_current * val_new < 0.0) bracket = true;
            } else {
                val_new = poly_eval<6>(pol, tz_max);
                if (val_current * val_new < 0.0) bracket = true;
                else if (fabs(val_new) < 8.0 * EPS) direct = true;
            }
            if (direct || bracket) {
                c_lo[n_cand] = bracket ? (double)tz_current : (double)tz;
                c_hi[n_cand] = tz;
                ++n_cand;
            }
            tz_current = tz;
            val_current = val_new;
        }
    }

    return found;
}

// UDUD half of time_vel (:825-973)
__device__ RK_S2_FN bool s2_vel_udud(const S2& s, const Lim& L, Hit& h) {
    RK_S2_LOCALS
    const real a0_p5 = s.a0_p5, a0_p6 = s.a0_p6, af_p6 = s.af_p6;
    const real tz_min = rmax(0.0, -a0 / j_max);
    const real tz_max = rmin((tf - a0 / j_max) * 0.5, (a_max - a0) / j_max);

    const real ph1 = af_af - 2.0 * j_max * (2.0 * af * tf + j_max * tf_tf - 3.0 * vd);
    const real ph2 = af_p3 - 3.0 * j_max_j_max * g1 + 3.0 * af * j_max * vd;
    const real ph3 = 2.0 * j_max * tf * g1 + 3.0 * vd_vd;
    const real ph4 = af_p4 - 8.0 * af_p3 * j_max * tf + 12.0 * j_max * (j_max * ph3 + af_af * vd + 2.0 * af * j_max * (g1 - tf * vd));
    const real ph5 = af + j_max * tf;

    // 6th order polynomial, monic
    double pol[7];
    pol[0] = 1.0;
    pol[1] = (5.0 * a0 - ph5) / j_max;
    pol[2] = (39.0 * a0_a0 - ph1 - 16.0 * a0 * ph5) / four_times(j_max_j_max);
    pol[3] = (55.0 * a0_p3 - 33.0 * a0_a0 * ph5 - 6.0 * a0 * ph1 + 2.0 * ph2) / (6.0 * j_max_j_max * j_max);
    pol[4] = (101.0 * a0_p4 + ph4 - 76.0 * a0_p3 * ph5 - 30.0 * a0_a0 * ph1 + 16.0 * a0 * ph2) / (24.0 * j_max_j_max * j_max_j_max);
    pol[5] = (a0 * (11.0 * a0_p4 + ph4 - 10.0 * a0_p3 * ph5 - 6.0 * a0_a0 * ph1 + 4.0 * a0 * ph2)) / (12.0 * j_max_j_max * j_max_j_max * j_max);
    pol[6] = (11.0 * a0_p6 - af_p6 - 12.0 * a0_p5 * ph5 - 48.0 * af_p3 * j_max_j_max * g1 - 9.0 * a0_p4 * ph1
              + 72.0 * j_max_j_max * j_max * (j_max * g1 * g1 - vd_vd * vd - 2.0 * af * g1 * vd)
              - 6.0 * af_p4 * j_max * vd - 36.0 * af_af * j_max_j_max * vd_vd + 8.0 * a0_p3 * ph2 + 3.0 * a0_a0 * ph4)
             / (144.0 * j_max_j_max * j_max_j_max * j_max_j_max);

    double deriv[6], dderiv[5], ddderiv[4];
    poly_monic_deri<7>(pol, deriv);
    poly_monic_deri<6>(deriv, dderiv);
    poly_deri<5>(dderiv, ddderiv);

    // sign changes of the first derivative between the extrema of the first derivative
    double iv_l[6], iv_r[6];
    int n_iv = 0;
    real dd_tz_current = tz_min;
    const Roots dd_extremas = solve_quart_monic(dderiv[1], dderiv[2], dderiv[3], dderiv[4]);
#pragma unroll 1
    for (int ri = 0; ri < dd_extremas.n; ++ri) {
        real tz = dd_extremas.get(ri);
        if (tz >= tz_max) continue;

        const real orig = poly_eval<5>(dderiv, tz);
        if (fabs(orig) > ROOT_TOLERANCE) tz -= orig / poly_eval<4>(ddderiv, tz);

        if (poly_eval<6>(deriv, dd_tz_current) * poly_eval<6>(deriv, tz) < 0.0) {
            bool dup = false;
            for (int k = 0; k < n_iv; ++k) dup |= (iv_l[k] == dd_tz_current && iv_r[k] == tz);
            if (!dup && n_iv < 6) { iv_l[n_iv] = dd_tz_current; iv_r[n_iv] = tz; ++n_iv; }
        }
        dd_tz_current = tz;
    }
    if (poly_eval<6>(deriv, dd_tz_current) * poly_eval<6>(deriv, tz_max) < 0.0) {
        bool dup = false;
        for (int k = 0; k < n_iv; ++k) dup |= (iv_l[k] == dd_tz_current && iv_r[k] == tz_max);
        if (!dup && n_iv < 6) { iv_l[n_iv] = dd_tz_current; iv_r[n_iv] = tz_max; ++n_iv; }
    }

    // record the candidates first, polish and validate them afterwards (see the UDDU half); pass k == n_iv is the interval
    // that ends at tz_max (:947-970)
    double c_lo[7], c_hi[7];
    int n_cand = 0;
    real tz_current = tz_min;
    real val_current = poly_eval<7>(pol, tz_current);
#pragma unroll 1
    for (int k = 0; k <= n_iv; ++k) {
        const bool last = k == n_iv;
        const real tz = last ? tz_max : real(shrink_interval<6>(deriv, iv_l[last ? 0 : k], iv_r[last ? 0 : k]));
        if (!last && tz >= tz_max) continue;
        const real p_val = poly_eval<7>(pol, tz);
        bool direct = false, bracket = false;
        if (!last && fabs(p_val) < 64.0 * fabs(poly_eval<5>(dderiv, tz)) * ROOT_TOLERANCE) direct = true;
        else if (val_current * p_val < 0.0) bracket = true;
        if (direct || bracket) {
            c_lo[n_cand] = bracket ? (double)tz_current : (double)tz;
            c_hi[n_cand] = tz;
            ++n_cand;
        }
        tz_current = tz;
        val_current = p_val;
    }
#pragma unroll 1
    for (int c = 0; c < n_cand; ++c) {
        const double lo = c_lo[c], hi = c_hi[c];
        const double root = (lo != hi) ? shrink_interval<7>(pol, lo, hi) : hi;
        if (vel_udud_root(s, L, h, root)) return true;
    }
    return false;
}

#ifndef RK_LS_BLOCK
#define RK_LS_BLOCK 128  // threads per block of the lockstep kernels (rk_pipeline.cuh)
#endif

// Lockstep form of the UDUD half: phases A-C as above, then the two bracketing levels inside the kernel.
__device__ __forceinline__ bool s2_vel_udud_lockstep(const S2& s, const Lim& L, Hit& h, bool active) {
    RK_S2_LOCALS
    const real a0_p5 = s.a0_p5, a0_p6 = s.a0_p6, af_p6 = s.af_p6;
    bool found = false;
    real tz_min = 0.0, tz_max = 0.0;
    double pol[7], deriv[6], dderiv[5], ddderiv[4];

    // phase A: the monic sextic and its derivatives
    if (active) {
        tz_min = rmax(0.0, -a0 / j_max);
        tz_max = rmin((tf - a0 / j_max) * 0.5, (a_max - a0) / j_max);
        const real ph1 = af_af - 2.0 * j_max * (2.0 * af * tf + j_max * tf_tf - 3.0 * vd);
        const real ph2 = af_p3 - 3.0 * j_max_j_max * g1 + 3.0 * af * j_max * vd;
        const real ph3 = 2.0 * j_max * tf * g1 + 3.0 * vd_vd;
        const real ph4 = af_p4 - 8.0 * af_p3 * j_max * tf + 12.0 * j_max * (j_max * ph3 + af_af * vd + 2.0 * af * j_max * (g1 - tf * vd));
        const real ph5 = af + j_max * tf;
        pol[0] = 1.0;
        pol[1] = (5.0 * a0 - ph5) / j_max;
        pol[2] = (39.0 * a0_a0 - ph1 - 16.0 * a0 * ph5) / four_times(j_max_j_max);
        pol[3] = (55.0 * a0_p3 - 33.0 * a0_a0 * ph5 - 6.0 * a0 * ph1 + 2.0 * ph2) / (6.0 * j_max_j_max * j_max);
        pol[4] = (101.0 * a0_p4 + ph4 - 76.0 * a0_p3 * ph5 - 30.0 * a0_a0 * ph1 + 16.0 * a0 * ph2) / (24.0 * j_max_j_max * j_max_j_max);
        pol[5] = (a0 * (11.0 * a0_p4 + ph4 - 10.0 * a0_p3 * ph5 - 6.0 * a0_a0 * ph1 + 4.0 * a0 * ph2)) / (12.0 * j_max_j_max * j_max_j_max * j_max);
        pol[6] = (11.0 * a0_p6 - af_p6 - 12.0 * a0_p5 * ph5 - 48.0 * af_p3 * j_max_j_max * g1 - 9.0 * a0_p4 * ph1
                  + 72.0 * j_max_j_max * j_max * (j_max * g1 * g1 - vd_vd * vd - 2.0 * af * g1 * vd)
                  - 6.0 * af_p4 * j_max * vd - 36.0 * af_af * j_max_j_max * vd_vd + 8.0 * a0_p3 * ph2 + 3.0 * a0_a0 * ph4)
                 / (144.0 * j_max_j_max * j_max_j_max * j_max_j_max);
        poly_monic_deri<7>(pol, deriv);
        poly_monic_deri<6>(deriv, dderiv);
        poly_deri<5>(dderiv, ddderiv);
    }
    RK_PHASE_BARRIER();

    // phase B: extrema of the first derivative in closed form
    Roots dd_extremas;
    if (active) dd_extremas = solve_quart_monic(dderiv[1], dderiv[2], dderiv[3], dderiv[4]);
    RK_PHASE_BARRIER();

    // phase C: sign changes of the first derivative between its extrema
    double iv_l[6], iv_r[6];
    int n_iv = 0;
    if (active) {
        real dd_tz_current = tz_min;
        real d_current = poly_eval<6>(deriv, dd_tz_current);  // the value at the left end is the previous pass's value at its right end
#pragma unroll 1
        for (int ri = 0; ri < dd_extremas.n; ++ri) {
            real tz = dd_extremas.get(ri);
            if (tz >= tz_max) continue;
            {   // polynomials evaluated at the same point share its powers (PolyPoint, rk_core.cuh)
                const PolyPoint<5> at(tz);
                const real orig = at.eval<5>(dderiv);
                if (fabs(orig) > ROOT_TOLERANCE) tz -= orig / at.eval<4>(ddderiv);
            }
            const real d_new = poly_eval<6>(deriv, tz);
            if (d_current * d_new < 0.0) {
                bool dup = false;
                for (int k = 0; k < n_iv; ++k) dup |= (iv_l[k] == dd_tz_current && iv_r[k] == tz);
                if (!dup && n_iv < 6) { iv_l[n_iv] = dd_tz_current; iv_r[n_iv] = tz; ++n_iv; }
            }
            dd_tz_current = tz;
            d_current = d_new;
        }
        if (d_current * poly_eval<6>(deriv, tz_max) < 0.0) {
            bool dup = false;
            for (int k = 0; k < n_iv; ++k) dup |= (iv_l[k] == dd_tz_current && iv_r[k] == tz_max);
            if (!dup && n_iv < 6) { iv_l[n_iv] = dd_tz_current; iv_r[n_iv] = tz_max; ++n_iv; }
        }
    }

    // phase D: extrema of the sextic by bracketing, candidates recorded; pass k == n_iv is the interval ending at tz_max
    double c_lo[7], c_hi[7];
    int n_cand = 0;
    real tz_current = tz_min;
    real val_current = 0.0;
    if (active) val_current = poly_eval<7>(pol, tz_current);
#pragma unroll 1
    for (int k = 0; k < 7; ++k) {
        const bool mine = active && k <= n_iv;
        if (!__syncthreads_or(mine)) break;
        if (mine) {
            const bool last = k == n_iv;
            const real tz = last ? tz_max : real(shrink_interval<6>(deriv, iv_l[last ? 0 : k], iv_r[last ? 0 : k]));
            if (last || !(tz >= tz_max)) {
                const PolyPoint<7> at(tz);
                const real p_val = at.eval<7>(pol);
                bool direct = false, bracket = false;
                if (!last && fabs(p_val) < 64.0 * fabs(at.eval<5>(dderiv)) * ROOT_TOLERANCE) direct = true;
                else if (val_current * p_val < 0.0) bracket = true;
                if (direct || bracket) {
                    c_lo[n_cand] = bracket ? (double)tz_current : (double)tz;
                    c_hi[n_cand] = tz;
                    ++n_cand;
                }
                tz_current = tz;
                val_current = p_val;
            }
        }
    }

    // phase E: polish and validate the candidates in order
#pragma unroll 1
    for (int c = 0; c < 7; ++c) {
        const bool mine = active && !found && c < n_cand;
        if (!__syncthreads_or(mine)) break;
        double root = 0.0;
        if (mine) {
            const double lo = c_lo[c], hi = c_hi[c];
            root = (lo != hi) ? shrink_interval<7>(pol, lo, hi) : hi;
        }
        RK_PHASE_BARRIER();
        if (mine) found = vel_udud_root(s, L, h, root);
    }
    return found;
}

// position_third_step2.rs:607-974
__device__ __forceinline__ bool s2_vel(const S2& s, const Lim& L, Hit& h) {
    return s2_vel_uddu(s, L, h) || s2_vel_udud(s, L, h);
}

}  // namespace rk
#endif  // RK_H_STEP2_VEL_GO
