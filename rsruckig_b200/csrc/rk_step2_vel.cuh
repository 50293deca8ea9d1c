// rk_step2_vel.cuh — Step 2 family VEL (position_third_step2.rs:607-974): the
// profile that cruises at v_max for part of tf. By far the most frequent
// winner of Step 2 on random inputs, and the only family that needs the roots
// of a 5th (UDDU) / 6th (UDUD) order polynomial: the extrema of the polynomial
// are found in closed form from its quartic (second) derivative, each sign
// change between consecutive extrema is then bracketed and polished by the
// safeguarded Newton iteration (shrink_interval).
#pragma once
#include "rk_step2.cuh"

namespace rk {

constexpr double ROOT_TOLERANCE = 1e-14;  // roots.rs TOLERANCE

// check_root closure of the UDDU branch (:715-776)
__device__ __forceinline__ bool vel_uddu_root(const S2& s, const Lim& L, Hit& h, double tt) {
    RK_S2_LOCALS
    // Single Newton step (regarding pd)
    {
        const double h1 = sqrt((a0_a0 + af_af) / (2.0 * j_max_j_max) + (2.0 * a0 * tt + j_max * tt * tt - vd) / j_max);
        const double orig = -pd
                            - (2.0 * a0_p3 + 4.0 * af_p3 + 24.0 * a0 * j_max * tt * (af + j_max * (h1 + tt - tf)) + 6.0 * a0_a0 * (af + j_max * (2.0 * tt - tf))
                               + 6.0 * (a0_a0 + af_af) * j_max * h1 + 12.0 * af * j_max * (j_max * tt * tt - vd)
                               + 12.0 * j_max_j_max * (j_max * tt * tt * (h1 + tt - tf) - tf * v0 - h1 * vd))
                                  / (12.0 * j_max_j_max);
        const double deriv_newton = -(a0 + j_max * tt) * (3.0 * (h1 + tt) - 2.0 * tf + (a0 + 2.0 * af) / j_max);
        if (!(orig != orig) && !(deriv_newton != deriv_newton) && fabs(deriv_newton) > EPS) tt -= orig / deriv_newton;
    }
    if (tt > tf || tt != tt) return false;

    const double h1 = sqrt((a0_a0 + af_af) / (2.0 * j_max_j_max) + (tt * (2.0 * a0 + j_max * tt) - vd) / j_max);
    t[0] = tt;
    t[1] = 0.0;
    t[2] = tt + a0 / j_max;
    t[3] = tf - 2.0 * (tt + h1) - (a0 + af) / j_max;
    t[4] = h1;
    t[5] = 0.0;
    t[6] = h1 + af / j_max;
    RK_TRY(UDDU, L_VEL, j_max)
    return false;
}

// check_root closure of the UDUD branch (:883-945)
__device__ __forceinline__ bool vel_udud_root(const S2& s, const Lim& L, Hit& h, double tt) {
    RK_S2_LOCALS
    // Double Newton step (regarding pd)
    {
        double h1 = sqrt((af_af - a0_a0) / (2.0 * j_max_j_max) - ((2.0 * a0 + j_max * tt) * tt - vd) / j_max);
        double orig = -pd + (af_p3 - a0_p3 + 3.0 * a0_a0 * j_max * (tf - 2.0 * tt)) / (6.0 * j_max_j_max) + (2.0 * a0 + j_max * tt) * tt * (tf - tt)
                      + (j_max * h1 - af) * h1 * h1 + tf * v0;
        double deriv_newton = (a0 + j_max * tt) * (2.0 * (af + j_max * tf) - 3.0 * j_max * (h1 + tt) - a0) / j_max;
        tt -= orig / deriv_newton;

        h1 = sqrt((af_af - a0_a0) / (2.0 * j_max_j_max) - ((2.0 * a0 + j_max * tt) * tt - vd) / j_max);
        orig = -pd + (af_p3 - a0_p3 + 3.0 * a0_a0 * j_max * (tf - 2.0 * tt)) / (6.0 * j_max_j_max) + (2.0 * a0 + j_max * tt) * tt * (tf - tt)
               + (j_max * h1 - af) * h1 * h1 + tf * v0;
        if (fabs(orig) > 1e-9) {
            deriv_newton = (a0 + j_max * tt) * (2.0 * (af + j_max * tf) - 3.0 * j_max * (h1 + tt) - a0) / j_max;
            tt -= orig / deriv_newton;
        }
    }
    const double h1 = sqrt((af_af - a0_a0) / (2.0 * j_max_j_max) - ((2.0 * a0 + j_max * tt) * tt - vd) / j_max);
    t[0] = tt;
    t[1] = 0.0;
    t[2] = tt + a0 / j_max;
    t[3] = tf - 2.0 * (tt + h1) + ad / j_max;
    t[4] = h1;
    t[5] = 0.0;
    t[6] = h1 - af / j_max;
    RK_TRY(UDUD, L_VEL, j_max)
    return false;
}

// UDDU half of time_vel (:619-824)
__device__ __noinline__ bool s2_vel_uddu(const S2& s, const Lim& L, Hit& h) {
    RK_S2_LOCALS
    const double a0_p6 = s.a0_p6, af_p6 = s.af_p6;
    const double tz_min = rmax(0.0, -a0 / j_max);
    const double tz_max = rmin((tf - a0 / j_max) / 2.0, (a_max - a0) / j_max);

    if (fabs(v0) < EPS && fabs(a0) < EPS && fabs(vf) < EPS && fabs(af) < EPS) {
        const Roots roots = solve_cub(1.0, -tf / 2.0, 0.0, pd / (2.0 * j_max));
#pragma unroll 1
        for (int ri = 0; ri < roots.n; ++ri) {
            double tt = roots.get(ri);
            if (tt > tf / 4.0) continue;
            // Single Newton step (regarding pd)
            if (tt > EPS) {
                const double orig = -pd + j_max * tt * tt * (tf - 2.0 * tt);
                const double deriv = 2.0 * j_max * tt * (tf - 3.0 * tt);
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

    const double p1 = af_af - 2.0 * j_max * (-2.0 * af * tf + j_max * tf_tf + 3.0 * vd);
    const double ph1 = af_p3 - 3.0 * j_max_j_max * g1 - 3.0 * af * j_max * vd;
    const double ph2 = af_p4 + 8.0 * af_p3 * j_max * tf
                       + 12.0 * j_max * (3.0 * j_max * vd_vd - af_af * vd + 2.0 * af * j_max * (g1 - tf * vd) - 2.0 * j_max_j_max * tf * g1);
    const double ph3 = a0 * (af - j_max * tf);
    const double ph4 = j_max * (-ad + j_max * tf);

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

    // One loop over the extrema plus a last pass for the interval that ends at tz_max (:778-823). Each pass decides
    // whether it has a root candidate (directly at the extremum, or bracketed between the previous extremum and this
    // one); the bracketing iteration and the candidate validation then have a single call site each, so the lanes of a
    // warp re-converge there instead of running four inlined copies one after the other.
    double tz_current = tz_min;
    double val_current = poly_eval<6>(pol, tz_current);
#pragma unroll 1
    for (int ri = 0; ri <= d_extremas.n; ++ri) {
        const bool last = ri == d_extremas.n;
        double tz = last ? tz_max : d_extremas.get(ri);
        if (!last && tz >= tz_max) continue;
        bool direct = false, bracket = false;
        double val_new;
        if (!last) {
            const double orig = poly_eval<5>(deriv, tz);
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
            const double root = bracket ? shrink_interval<6>(pol, tz_current, tz) : tz;
            if (vel_uddu_root(s, L, h, root)) return true;
        }
        tz_current = tz;
        val_current = val_new;
    }
    return false;
}

// UDUD half of time_vel (:825-973)
__device__ __noinline__ bool s2_vel_udud(const S2& s, const Lim& L, Hit& h) {
    RK_S2_LOCALS
    const double a0_p5 = s.a0_p5, a0_p6 = s.a0_p6, af_p6 = s.af_p6;
    const double tz_min = rmax(0.0, -a0 / j_max);
    const double tz_max = rmin((tf - a0 / j_max) / 2.0, (a_max - a0) / j_max);

    const double ph1 = af_af - 2.0 * j_max * (2.0 * af * tf + j_max * tf_tf - 3.0 * vd);
    const double ph2 = af_p3 - 3.0 * j_max_j_max * g1 + 3.0 * af * j_max * vd;
    const double ph3 = 2.0 * j_max * tf * g1 + 3.0 * vd_vd;
    const double ph4 = af_p4 - 8.0 * af_p3 * j_max * tf + 12.0 * j_max * (j_max * ph3 + af_af * vd + 2.0 * af * j_max * (g1 - tf * vd));
    const double ph5 = af + j_max * tf;

    // 6th order polynomial, monic
    double pol[7];
    pol[0] = 1.0;
    pol[1] = (5.0 * a0 - ph5) / j_max;
    pol[2] = (39.0 * a0_a0 - ph1 - 16.0 * a0 * ph5) / (4.0 * j_max_j_max);
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
    double dd_tz_current = tz_min;
    const Roots dd_extremas = solve_quart_monic(dderiv[1], dderiv[2], dderiv[3], dderiv[4]);
#pragma unroll 1
    for (int ri = 0; ri < dd_extremas.n; ++ri) {
        double tz = dd_extremas.get(ri);
        if (tz >= tz_max) continue;

        const double orig = poly_eval<5>(dderiv, tz);
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

    // same single-call-site arrangement as the UDDU half: pass k == n_iv is the interval that ends at tz_max (:947-970)
    double tz_current = tz_min;
    double val_current = poly_eval<7>(pol, tz_current);
#pragma unroll 1
    for (int k = 0; k <= n_iv; ++k) {
        const bool last = k == n_iv;
        const double tz = last ? tz_max : shrink_interval<6>(deriv, iv_l[last ? 0 : k], iv_r[last ? 0 : k]);
        if (!last && tz >= tz_max) continue;
        const double p_val = poly_eval<7>(pol, tz);
        bool direct = false, bracket = false;
        if (!last && fabs(p_val) < 64.0 * fabs(poly_eval<5>(dderiv, tz)) * ROOT_TOLERANCE) direct = true;
        else if (val_current * p_val < 0.0) bracket = true;
        if (direct || bracket) {
            const double root = bracket ? shrink_interval<7>(pol, tz_current, tz) : tz;
            if (vel_udud_root(s, L, h, root)) return true;
        }
        tz_current = tz;
        val_current = p_val;
    }
    return false;
}

// position_third_step2.rs:607-974
__device__ __forceinline__ bool s2_vel(const S2& s, const Lim& L, Hit& h) {
    return s2_vel_uddu(s, L, h) || s2_vel_udud(s, L, h);
}

}  // namespace rk
