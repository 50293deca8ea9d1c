// rk_kernels.cu — the four kernels of the batched calculator and the C ABI
// (include/ruckig_b200.h). sm_100a only; compiled with --fmad=false.
//
//   k_step1   one thread per (DoF, trajectory): validation of that DoF, brake
//             pre-trajectory, Step 1 (extremal profiles -> block of feasible
//             durations)                          calculator_target.rs:300-474
//   k_sync    one thread per trajectory: status from validation / Step-1
//             failures, synchronisation (common duration, limiting DoF),
//             phase synchronisation, per-DoF decision "reuse an extremal
//             profile or run Step 2"               calculator_target.rs:150-275, :475-747
//   k_step2   one thread per (DoF, trajectory): Step 2 where needed, expansion
//             of the winning compact profile into the full Profile arrays
//                                                   calculator_target.rs:749-830
//   k_at_time one thread per (DoF, trajectory), loop over K sample times
//                                                   trajectory.rs:52-189
//
// Items are laid out DoF-major / trajectory-fastest, so a warp always works on
// 32 consecutive trajectories of the same DoF: all global accesses are
// coalesced 256-byte rows, and every lane of a warp runs the same solver
// family with the same per-DoF options.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/ruckig_b200.h"
#include "rk_step1.cuh"
#include "rk_step_misc.cuh"
#include "rk_step2.cuh"
#include "rk_step2_vel.cuh"
#include "rk_step2_acc.cuh"
#include "rk_step2_none.cuh"
#include "rk_store.cuh"

namespace rk {

struct Params {
    int64_t n;        // trajectories in the batch (stride of the caller's arrays and of the trajectory storage)
    int64_t off;      // first trajectory of this chunk
    int64_t cn;       // trajectories in this chunk (stride of the scratch arrays)
    int32_t dof;
    int32_t discrete;
    int32_t throw_mode;
    int32_t check_current, check_target;
    double delta_time;
    int8_t ci[RK_MAX_DOF];
    int8_t sync[RK_MAX_DOF];
    // inputs, [dof][n]
    const double *p0, *v0, *a0, *pf, *vf, *af, *v_max, *a_max, *j_max, *v_min, *a_min, *min_dur;
    const uint8_t* enabled;
    // scratch, chunk-local
    double* ws;       // [WS_SLOTS][dof][cn]
    uint32_t* wmeta;  // [4][dof][cn]
    uint8_t* wsrc;    // [dof][cn]
    // trajectory storage
    double* prof;     // [PROFILE_SLOTS][dof][n]
    uint32_t* codes;  // [dof][n]
    int32_t* status;  // [n]
    double* duration; // [n]
    double* imd;      // [dof][n]
    int32_t* limiting;  // [n]
    int32_t* detail;    // [n]
};

struct DofIn {
    double p0, v0, a0, pf, vf, af, v_max, v_min, a_max, a_min, j_max;
    bool enabled;
};

__device__ __forceinline__ DofIn load_dof(const Params& P, int d, int64_t ig) {
    const int64_t g = (int64_t)d * P.n + ig;
    DofIn x;
    x.p0 = P.p0[g]; x.v0 = P.v0[g]; x.a0 = P.a0[g];
    x.pf = P.pf[g]; x.vf = P.vf[g]; x.af = P.af[g];
    x.v_max = P.v_max[g]; x.a_max = P.a_max[g]; x.j_max = P.j_max[g];
    x.v_min = P.v_min ? P.v_min[g] : -x.v_max;
    x.a_min = P.a_min ? P.a_min[g] : -x.a_max;
    x.enabled = P.enabled ? (P.enabled[g] != 0) : true;
    return x;
}

// InputParameter::validate for one DoF (input_parameter.rs:251-420); returns 0 or the id of the first failing check.
__device__ __forceinline__ int validate_dof(const DofIn& x, int ci, bool check_current, bool check_target) {
    const double j_max = x.j_max, a_max = x.a_max, a_min = x.a_min, a0 = x.a0, af = x.af, v0 = x.v0, vf = x.vf;
    if (j_max != j_max || j_max < 0.0) return RK_INVALID_MAX_JERK;
    if (a_max != a_max || a_max < 0.0) return RK_INVALID_MAX_ACCELERATION;
    if (a_min != a_min || a_min > 0.0) return RK_INVALID_MIN_ACCELERATION;
    if (a0 != a0) return RK_INVALID_CURRENT_ACCELERATION_NAN;
    if (af != af) return RK_INVALID_TARGET_ACCELERATION_NAN;
    if (check_current) {
        if (a0 > a_max) return RK_INVALID_CURRENT_ACCELERATION_ABOVE;
        if (a0 < a_min) return RK_INVALID_CURRENT_ACCELERATION_BELOW;
    }
    if (check_target) {
        if (af > a_max) return RK_INVALID_TARGET_ACCELERATION_ABOVE;
        if (af < a_min) return RK_INVALID_TARGET_ACCELERATION_BELOW;
    }
    if (v0 != v0) return RK_INVALID_CURRENT_VELOCITY_NAN;
    if (vf != vf) return RK_INVALID_TARGET_VELOCITY_NAN;
    if (ci == RK_POSITION) {
        if (x.p0 != x.p0) return RK_INVALID_CURRENT_POSITION_NAN;
        if (x.pf != x.pf) return RK_INVALID_TARGET_POSITION_NAN;
        const double v_max = x.v_max, v_min = x.v_min;
        if (v_max != v_max || v_max < 0.0) return RK_INVALID_MAX_VELOCITY;
        if (v_min != v_min || v_min > 0.0) return RK_INVALID_MIN_VELOCITY;
        if (check_current) {
            if (v0 > v_max) return RK_INVALID_CURRENT_VELOCITY_ABOVE;
            if (v0 < v_min) return RK_INVALID_CURRENT_VELOCITY_BELOW;
        }
        if (check_target) {
            if (vf > v_max) return RK_INVALID_TARGET_VELOCITY_ABOVE;
            if (vf < v_min) return RK_INVALID_TARGET_VELOCITY_BELOW;
        }
        if (check_current) {
            if (a0 > 0.0 && j_max > 0.0 && v_at_a_zero(v0, a0, j_max) > v_max) return RK_INVALID_CURRENT_WILL_EXCEED;
            if (a0 < 0.0 && j_max > 0.0 && v_at_a_zero(v0, a0, -j_max) < v_min) return RK_INVALID_CURRENT_WILL_UNDERCUT;
        }
        if (check_target) {
            if (af < 0.0 && j_max > 0.0 && v_at_a_zero(vf, af, j_max) > v_max) return RK_INVALID_TARGET_WILL_EXCEED;
            if (af > 0.0 && j_max > 0.0 && v_at_a_zero(vf, af, -j_max) < v_min) return RK_INVALID_TARGET_WILL_UNDERCUT;
        }
    }
    return 0;
}

__device__ __forceinline__ int kind_of(int ci, double a_max, double j_max) {
    const bool jinf = isinf(j_max), ainf = isinf(a_max);
    if (ci == RK_POSITION) return !jinf ? K_POS3 : (!ainf ? K_POS2 : K_POS1);
    if (ci == RK_VELOCITY) return !jinf ? K_VEL3 : K_VEL2;
    return K_NONE;  // ControlInterface::Acceleration is not implemented by the reference (Q20)
}

__device__ __forceinline__ void ws_put_cand(const Params& P, int64_t wi, int64_t wstride, int slot, const Cand& c) {
    double* base = P.ws + (int64_t)(W_CAND + slot * W_CAND_STRIDE) * wstride + wi;
#pragma unroll
    for (int k = 0; k < 7; ++k) base[(int64_t)k * wstride] = c.t[k];
    base[7 * wstride] = c.ctrl;
    base[8 * wstride] = c.ctrl2;
    P.wmeta[(int64_t)(1 + slot) * wstride + wi] = pack_codes(c.lim, c.cs, c.dir, 0);
}

__device__ __forceinline__ void ws_get_cand(const Params& P, int64_t wi, int64_t wstride, int slot, double (&t)[7], double& ctrl, double& ctrl2, uint32_t& codes) {
    const double* base = P.ws + (int64_t)(W_CAND + slot * W_CAND_STRIDE) * wstride + wi;
#pragma unroll
    for (int k = 0; k < 7; ++k) t[k] = base[(int64_t)k * wstride];
    ctrl = base[7 * wstride];
    ctrl2 = base[8 * wstride];
    codes = P.wmeta[(int64_t)(1 + slot) * wstride + wi];
}

// ------------------------------------------------------------------------------------------------ k_step1
__global__ void __launch_bounds__(128) k_step1(const Params P) {
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= wstride) return;
    const int d = (int)(wi / P.cn);
    const int64_t il = wi - (int64_t)d * P.cn;
    const int64_t ig = P.off + il;

    const DofIn x = load_dof(P, d, ig);
    const int ci = P.ci[d];
    const int vcode = validate_dof(x, ci, false, true);  // Ruckig::calculate validates with (false, true), ruckig.rs:215
    const int kind = kind_of(ci, x.a_max, x.j_max);

    ProfileSink out{P.prof, (int64_t)P.dof * P.n, (int64_t)d * P.n + ig};
    uint32_t meta = (uint32_t)kind | ((uint32_t)vcode << M_VCODE_SHIFT);
    if (x.enabled) meta |= M_ENABLED;
    if (x.a_max == 0.0 || x.a_min == 0.0 || x.j_max == 0.0) meta |= M_ZERO_LIMITS;

    // fresh Profile: brake and boundary
    double bdur = 0.0, bt0 = 0.0, bt1 = 0.0, bj0 = 0.0, bj1 = 0.0, ba0 = 0.0, ba1 = 0.0, bv0 = 0.0, bv1 = 0.0, bp0 = 0.0, bp1 = 0.0;
    double ps = x.p0, vs = x.v0, as = x.a0;
    double pf_store = 0.0, vf_store = 0.0, af_store = 0.0;
    double t_min = 0.0;
    BlockRec blk;
    blk.has_a = 0; blk.has_b = 0; blk.t_min = 0.0; blk.a_left = 0.0; blk.a_right = 0.0; blk.b_left = 0.0; blk.b_right = 0.0;
    bool found = false;

    if (x.enabled && (vcode == 0 || !P.throw_mode)) {
        // brake pre-trajectory (calculator_target.rs:330-387)
        Brake br;
        br.t[0] = 0.0; br.t[1] = 0.0; br.j[0] = 0.0; br.j[1] = 0.0; br.a0 = 0.0;
        if (kind == K_POS3) position_brake(br, x.v0, x.a0, x.v_max, x.v_min, x.a_max, x.a_min, x.j_max);
        else if (kind == K_POS2) position_brake_second_order(br, x.v0, x.v_max, x.v_min, x.a_max, x.a_min);
        else if (kind == K_VEL3) velocity_brake_third_order(br, x.a0, x.a_max, x.a_min, x.j_max);
        if (ci == RK_POSITION) { pf_store = x.pf; vf_store = x.vf; af_store = x.af; }
        else if (ci == RK_VELOCITY) { vf_store = x.vf; af_store = x.af; }
        bt0 = br.t[0]; bt1 = br.t[1]; bj0 = br.j[0]; bj1 = br.j[1];
        if (kind == K_POS3 || kind == K_VEL3) {  // brake.rs:195-220
            if (!(bt0 <= 0.0 && bt1 <= 0.0)) {
                bdur = bt0;
                bp0 = ps; bv0 = vs; ba0 = as;
                integrate(bt0, ps, vs, as, bj0, ps, vs, as);
                if (bt1 > 0.0) {
                    bdur += bt1;
                    bp1 = ps; bv1 = vs; ba1 = as;
                    integrate(bt1, ps, vs, as, bj1, ps, vs, as);
                }
            }
        } else if (kind == K_POS2) {  // brake.rs:222-235
            ba0 = br.a0;
            if (!(bt0 <= 0.0)) {
                bdur = bt0;
                bp0 = ps; bv0 = vs;
                integrate(bt0, ps, vs, ba0, 0.0, ps, vs, as);
            }
        }

        Bound b{ps, vs, as, x.pf, x.vf, x.af};
        switch (kind) {
            case K_POS3: found = step1_pos3(b, x.v_max, x.v_min, x.a_max, x.a_min, x.j_max, bdur, blk); break;
            case K_POS2: found = step1_pos2(b, x.v_max, x.v_min, x.a_max, x.a_min, bdur, blk); break;
            case K_POS1: found = step1_pos1(b, x.v_max, x.v_min, bdur, blk); break;
            case K_VEL3: found = step1_vel3(b, x.a_max, x.a_min, x.j_max, bdur, blk); break;
            case K_VEL2: found = step1_vel2(b, x.a_max, x.a_min, bdur, blk); break;
            default: found = false; break;
        }
        if (found) { meta |= M_FOUND; t_min = blk.t_min; }
    }
    if (found && blk.has_a) meta |= M_HAS_A;
    if (found && blk.has_b) meta |= M_HAS_B;

    // scratch for the synchronisation / Step-2 kernels
    double* w = P.ws + wi;
    w[(int64_t)W_P0 * wstride] = ps; w[(int64_t)W_V0 * wstride] = vs; w[(int64_t)W_A0 * wstride] = as;
    w[(int64_t)W_TMIN * wstride] = t_min;
    w[(int64_t)W_ALEFT * wstride] = blk.a_left; w[(int64_t)W_ARIGHT * wstride] = blk.a_right;
    w[(int64_t)W_BLEFT * wstride] = blk.b_left; w[(int64_t)W_BRIGHT * wstride] = blk.b_right;
    w[(int64_t)W_BRAKE_DUR * wstride] = bdur;
    if (found) {
        ws_put_cand(P, wi, wstride, 0, blk.p_min);
        if (blk.has_a) ws_put_cand(P, wi, wstride, 1, blk.pa);
        if (blk.has_b) ws_put_cand(P, wi, wstride, 2, blk.pb);
    }
    P.wmeta[wi] = meta;

    // the parts of the final Profile that do not depend on the later stages
    out.put(S_BRAKE_DUR, bdur);
    out.put(S_BRAKE_T, bt0); out.put(S_BRAKE_T + 1, bt1);
    out.put(S_BRAKE_J, bj0); out.put(S_BRAKE_J + 1, bj1);
    out.put(S_BRAKE_A, ba0); out.put(S_BRAKE_A + 1, ba1);
    out.put(S_BRAKE_V, bv0); out.put(S_BRAKE_V + 1, bv1);
    out.put(S_BRAKE_P, bp0); out.put(S_BRAKE_P + 1, bp1);
    out.put(S_PF, pf_store); out.put(S_VF, vf_store); out.put(S_AF, af_store);
    P.imd[(int64_t)d * P.n + ig] = found ? t_min : 0.0;
}

// ------------------------------------------------------------------------------------------------ k_sync
// block.rs:157-172
__device__ __forceinline__ bool is_blocked(double t, double t_min, bool has_a, double a_left, double a_right, bool has_b, double b_left, double b_right) {
    if (t < t_min) return true;
    if (has_a && t > a_left && t < a_right) return true;
    if (has_b && t > b_left && t < b_right) return true;
    return false;
}

__global__ void __launch_bounds__(128) k_sync(const Params P) {
    const int64_t il = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (il >= P.cn) return;
    const int64_t ig = P.off + il;
    const int D = P.dof;
    const int64_t wstride = (int64_t)D * P.cn;
    const double* W = P.ws;
#define WS(slot, d) W[(int64_t)(slot) * wstride + (int64_t)(d) * P.cn + il]
#define META(d) P.wmeta[(int64_t)(d) * P.cn + il]
#define SRC(d) P.wsrc[(int64_t)(d) * P.cn + il]

    for (int d = 0; d < D; ++d) SRC(d) = RK_SRC_NONE;
    int status = RK_RESULT_WORKING;
    int detail = 0;
    double duration = 0.0;
    int limiting = -1;

    // Err(ValidationError): first failing DoF (input_parameter.rs:251-420), then the delta_time rule (ruckig.rs:161-168)
    if (P.throw_mode) {
        for (int d = 0; d < D && status == 0; ++d) {
            const uint32_t vc = (META(d) >> M_VCODE_SHIFT) & 0xFFu;
            if (vc) { status = RK_RESULT_ERROR_INVALID_INPUT; detail = (int)(vc << 8) | d; }
        }
        if (status == 0 && P.delta_time <= 0.0 && P.discrete) { status = RK_RESULT_ERROR_INVALID_INPUT; detail = RK_INVALID_DELTA_TIME << 8; }
    }
    // Step-1 failure: lowest DoF index decides (calculator_target.rs:454-471)
    if (status == 0) {
        for (int d = 0; d < D && status == 0; ++d) {
            const uint32_t m = META(d);
            if ((m & M_ENABLED) && !(m & M_FOUND)) {
                status = (m & M_ZERO_LIMITS) ? RK_RESULT_ERROR_ZERO_LIMITS : RK_RESULT_ERROR_EXECUTION_TIME_CALCULATION;
                detail = (1 << 8) | d;
            }
        }
    }

    const double md_raw = P.min_dur ? P.min_dur[ig] : __longlong_as_double(0x7ff8000000000000LL);
    const bool has_md = md_raw == md_raw;
    const double md = has_md ? md_raw : 0.0;

    bool done = status != 0;
    if (!done && D == 1 && !has_md && !P.discrete) {  // :475-481
        duration = WS(W_TMIN, 0);
        if (META(0) & M_ENABLED) SRC(0) = RK_SRC_STEP1_MIN;
        done = true;
    }

    if (!done) {
        // ---- synchronize (:150-275)
        double cand[3 * RK_MAX_DOF + 1];
        int16_t idx[3 * RK_MAX_DOF + 1];
        const int NC = 3 * D + 1;
        bool any_interval = false;
        for (int d = 0; d < D; ++d) {
            if (P.sync[d] == RK_SYNC_NONE) {
                cand[d] = 0.0; cand[D + d] = RK_INF; cand[2 * D + d] = RK_INF;
                continue;
            }
            const uint32_t m = META(d);
            cand[d] = WS(W_TMIN, d);
            cand[D + d] = (m & M_HAS_A) ? WS(W_ARIGHT, d) : RK_INF;
            cand[2 * D + d] = (m & M_HAS_B) ? WS(W_BRIGHT, d) : RK_INF;
            any_interval |= (m & (M_HAS_A | M_HAS_B)) != 0;
        }
        cand[3 * D] = has_md ? md : RK_INF;
        any_interval |= has_md;

        if (P.discrete) {
            for (int i = 0; i < NC; ++i) {
                if (isinf(cand[i])) continue;
                const double remainder = fmod(cand[i], P.delta_time);
                if (remainder > EPS) cand[i] += P.delta_time - remainder;
            }
        }

        const int idx_end = any_interval ? NC : D;
        for (int i = 0; i < NC; ++i) idx[i] = (i < idx_end) ? (int16_t)i : (int16_t)0;  // fresh scratch: the tail is zero (Q4)
        for (int i = 1; i < idx_end; ++i) {  // stable insertion sort == slice::sort_by
            const int16_t key = idx[i];
            const double kv = cand[key];
            int k = i - 1;
            while (k >= 0 && cand[idx[k]] > kv) { idx[k + 1] = idx[k]; --k; }
            idx[k + 1] = key;
        }

        bool found_sync = false;
        int which = -1;
        for (int k = D - 1; k < NC && !found_sync; ++k) {
            const int i = idx[k];
            const double ts = cand[i];
            bool blocked = false;
            for (int d = 0; d < D && !blocked; ++d) {
                if (P.sync[d] == RK_SYNC_NONE) continue;
                const uint32_t m = META(d);
                blocked = is_blocked(ts, WS(W_TMIN, d), (m & M_HAS_A) != 0, WS(W_ALEFT, d), WS(W_ARIGHT, d), (m & M_HAS_B) != 0, WS(W_BLEFT, d), WS(W_BRIGHT, d));
            }
            if (blocked || ts < (has_md ? md : 0.0) || isinf(ts)) continue;
            duration = ts;
            found_sync = true;
            which = i;
        }

        if (!found_sync) {  // :492-518
            bool zero_limits = false;
            for (int d = 0; d < D; ++d) zero_limits |= (META(d) & M_ZERO_LIMITS) != 0;
            status = zero_limits ? RK_RESULT_ERROR_ZERO_LIMITS : RK_RESULT_ERROR_SYNCHRONIZATION_CALCULATION;
            detail = 2 << 8;
            done = true;
        } else {
            if (which != 3 * D) {
                limiting = which % D;
                SRC(limiting) = (uint8_t)(which / D);  // 0 p_min, 1 a.profile, 2 b.profile
            }
            // None synchronisation (:520-528)
            for (int d = 0; d < D; ++d) {
                if ((META(d) & M_ENABLED) && P.sync[d] == RK_SYNC_NONE) {
                    SRC(d) = RK_SRC_STEP1_MIN;
                    const double tm = WS(W_TMIN, d);
                    if (tm > duration) { duration = tm; limiting = d; }
                }
            }
            if (duration > 7.6e3) {  // :531-533
                status = RK_RESULT_ERROR_TRAJECTORY_DURATION;
                done = true;
            } else if (fabs(duration - 0.0) < EPS) {  // :535-541
                for (int d = 0; d < D; ++d) SRC(d) = (META(d) & M_ENABLED) ? RK_SRC_STEP1_MIN : RK_SRC_NONE;
                done = true;
            } else {
                bool all_none = true, any_phase = false, all_phase_or_none = true;
                for (int d = 0; d < D; ++d) {
                    const int s = P.sync[d];
                    all_none &= (s == RK_SYNC_NONE);
                    any_phase |= (s == RK_SYNC_PHASE);
                    all_phase_or_none &= (s == RK_SYNC_PHASE || s == RK_SYNC_NONE);
                }
                if (!P.discrete && all_none) done = true;  // :543-550

                // ---- phase synchronisation (:552-711, is_input_collinear :60-148)
                if (!done && limiting >= 0 && any_phase) {
                    double lt[7], lctrl, lctrl2;
                    uint32_t lcodes;
                    ws_get_cand(P, (int64_t)limiting * P.cn + il, wstride, SRC(limiting), lt, lctrl, lctrl2, lcodes);
                    const int l_lim = lcodes & 0xFF, l_cs = (lcodes >> 8) & 0xFF, l_dir = (lcodes >> 16) & 0xFF;

                    // scale vector selection
                    int scale_dof = -1, scale_vec = -1;  // 0 pd, 1 v0, 2 a0, 3 vf, 4 af
                    for (int d = 0; d < D && scale_dof < 0; ++d) {
                        if (P.sync[d] != RK_SYNC_PHASE) continue;
                        const DofIn x = load_dof(P, d, ig);
                        const double pd = x.pf - x.p0;
                        if (P.ci[d] == RK_POSITION && fabs(pd) > EPS) { scale_vec = 0; scale_dof = d; }
                        else if (fabs(x.v0) > EPS) { scale_vec = 1; scale_dof = d; }
                        else if (fabs(x.a0) > EPS) { scale_vec = 2; scale_dof = d; }
                        else if (fabs(x.vf) > EPS) { scale_vec = 3; scale_dof = d; }
                        else if (fabs(x.af) > EPS) { scale_vec = 4; scale_dof = d; }
                    }
                    bool collinear = scale_dof >= 0;
                    double pd_scale = 0.0, v0_scale = 0.0, vf_scale = 0.0, a0_scale = 0.0, af_scale = 0.0, scale_limiting = 0.0, control_limiting = 0.0;
                    auto pick = [&](const DofIn& x) {
                        return scale_vec == 0 ? (x.pf - x.p0) : (scale_vec == 1 ? x.v0 : (scale_vec == 2 ? x.a0 : (scale_vec == 3 ? x.vf : x.af)));
                    };
                    if (collinear) {
                        const DofIn xs = load_dof(P, scale_dof, ig);
                        const double scale = pick(xs);
                        pd_scale = (xs.pf - xs.p0) / scale;
                        v0_scale = xs.v0 / scale;
                        vf_scale = xs.vf / scale;
                        a0_scale = xs.a0 / scale;
                        af_scale = xs.af / scale;
                        const DofIn xl = load_dof(P, limiting, ig);
                        scale_limiting = pick(xl);
                        control_limiting = (l_dir == DIR_UP) ? xl.j_max : -xl.j_max;
                        if (isinf(xl.j_max)) control_limiting = (l_dir == DIR_UP) ? xl.a_max : xl.a_min;
                        for (int d = 0; d < D && collinear; ++d) {
                            if (P.sync[d] != RK_SYNC_PHASE) continue;
                            const DofIn x = load_dof(P, d, ig);
                            const double cs_ = pick(x);
                            if ((P.ci[d] == RK_POSITION && fabs((x.pf - x.p0) - pd_scale * cs_) > EPS) || fabs(x.v0 - v0_scale * cs_) > EPS
                                || fabs(x.a0 - a0_scale * cs_) > EPS || fabs(x.vf - vf_scale * cs_) > EPS || fabs(x.af - af_scale * cs_) > EPS)
                                collinear = false;
                        }
                    }
                    if (collinear && all_phase_or_none) {
                        // Pass 0 validates every phase DoF with the limiting DoF's timing; only if all pass (and every DoF is
                        // Phase or None, :701-708) pass 1 commits the scaled profiles. Otherwise nothing is kept: the
                        // reference overwrites its partial results in the time-synchronisation loop.
                        bool found_phase = true;
                        for (int pass = 0; pass < 2 && found_phase; ++pass) {
                            for (int d = 0; d < D; ++d) {
                                const uint32_t m = META(d);
                                if (!(m & M_ENABLED) || d == limiting || P.sync[d] != RK_SYNC_PHASE) continue;
                                const DofIn x = load_dof(P, d, ig);
                                const double npc = control_limiting * pick(x) / scale_limiting;
                                int kind = m & M_KIND_MASK;
                                // the reference routes a first-order DoF with a UDUD limiting profile to the second-order check (:629-641)
                                if (kind == K_POS1 && l_cs == UDUD) kind = K_POS2;
                                int dir = DIR_UP;
                                switch (kind) {
                                    case K_POS3: case K_POS2: dir = (x.v_max > 0.0) ? DIR_UP : DIR_DOWN; break;
                                    case K_VEL3: dir = (x.a_max > 0.0) ? DIR_UP : DIR_DOWN; break;
                                    default: dir = (npc > 0.0) ? DIR_UP : DIR_DOWN; break;
                                }
                                if (pass == 0) {
                                    const Bound b{WS(W_P0, d), WS(W_V0, d), WS(W_A0, d), x.pf, x.vf, x.af};
                                    bool ok = true;
                                    switch (kind) {
                                        case K_POS3: {
                                            const Lim L{x.v_max, x.v_min, x.a_max, x.a_min, x.j_max};
                                            ok = check_pos3_full(lt, l_cs, L_NONE, npc, L, b, x.j_max);
                                            break;
                                        }
                                        case K_POS2:
                                            ok = (x.a_min - A_EPS < npc) && (npc < x.a_max + A_EPS) && (x.a_min - A_EPS < -npc) && (-npc < x.a_max + A_EPS)
                                                 && check_pos2(lt, l_cs, npc, -npc, x.v_max, x.v_min, b);
                                            break;
                                        case K_POS1: ok = (x.v_min - V_EPS < npc) && (npc < x.v_max + V_EPS) && check_pos1(lt, npc, b); break;
                                        case K_VEL3: ok = (fabs(npc) < fabs(x.j_max) + J_EPS) && check_vel3(lt, l_cs, L_NONE, npc, x.a_max, x.a_min, b); break;
                                        case K_VEL2: ok = (x.a_min - A_EPS < npc) && (npc < x.a_max + A_EPS) && check_vel2(lt, npc, b); break;
                                        default: break;
                                    }
                                    found_phase &= ok;
                                } else {
                                    // slot 2 (the b-profile) of this DoF is free now: the trajectory is decided
                                    Cand c;
                                    copy_t(c.t, lt);
                                    c.ctrl = npc; c.ctrl2 = -npc; c.lim = l_lim; c.cs = l_cs; c.dir = dir; c.dur = 0.0;
                                    const int64_t wd = (int64_t)d * P.cn + il;
                                    ws_put_cand(P, wd, wstride, 2, c);
                                    P.wmeta[(int64_t)3 * wstride + wd] |= (uint32_t)(kind + 1) << 24;  // kind the profile is expanded with
                                    SRC(d) = RK_SRC_PHASE;
                                }
                            }
                        }
                        if (found_phase) done = true;
                    }
                }

                // ---- time synchronisation dispatch (:714-747)
                if (!done) {
                    for (int d = 0; d < D; ++d) {
                        const uint32_t m = META(d);
                        const bool skip = (d == limiting || P.sync[d] == RK_SYNC_NONE) && !P.discrete;
                        if (!(m & M_ENABLED) || skip) continue;
                        const double t_profile = duration - WS(W_BRAKE_DUR, d) - 0.0;
                        if (P.sync[d] == RK_SYNC_TIME_IF_NECESSARY) {
                            const DofIn x = load_dof(P, d, ig);
                            if (fabs(x.vf) < EPS && fabs(x.af) < EPS) { SRC(d) = RK_SRC_STEP1_MIN; continue; }
                        }
                        // extremal profile re-use; b is only looked at when a is absent (Q3), i.e. never (b implies a)
                        if (fabs(t_profile - WS(W_TMIN, d)) < 2.0 * EPS) { SRC(d) = RK_SRC_STEP1_MIN; continue; }
                        if (m & M_HAS_A) {
                            if (fabs(t_profile - WS(W_ARIGHT, d)) < 2.0 * EPS) { SRC(d) = RK_SRC_STEP1_A; continue; }
                        } else if (m & M_HAS_B) {
                            if (fabs(t_profile - WS(W_BRIGHT, d)) < 2.0 * EPS) { SRC(d) = RK_SRC_STEP1_B; continue; }
                        }
                        SRC(d) = RK_SRC_STEP2;
                    }
                }
            }
        }
    }
    P.status[ig] = status;
    P.duration[ig] = duration;
    P.limiting[ig] = limiting;
    P.detail[ig] = detail;
#undef WS
#undef META
#undef SRC
}

// ------------------------------------------------------------------------------------------------ k_step2
// position_third_step2.rs:2449-2482: 16 (family, direction) stages, first valid candidate wins.
__device__ __forceinline__ bool step2_pos3(const Bound& b, double tf, double v_max, double v_min, double a_max, double a_min, double j_max, Hit& h) {
    S2 s;
    s.init(b, tf, j_max);
    const bool up_first = s.pd > tf * b.v0;
    const Lim La = lim_dir(up_first, v_max, v_min, a_max, a_min, j_max);
    const Lim Lb = lim_dir(!up_first, v_max, v_min, a_max, a_min, j_max);
#pragma unroll 1
    for (int stage = 0; stage < 16; ++stage) {
        const Lim& L = ((stage >> 2) & 1) ? Lb : La;
        bool ok;
        switch (stage & 3) {
            case 0: ok = (stage < 8) ? s2_acc0_acc1_vel(s, L, h) : s2_acc0_acc1(s, L, h); break;
            case 1: ok = (stage < 8) ? s2_vel(s, L, h) : s2_acc0(s, L, h); break;
            case 2: ok = (stage < 8) ? s2_acc0_vel(s, L, h) : s2_acc1(s, L, h); break;
            default: ok = (stage < 8) ? s2_acc1_vel(s, L, h) : s2_none(s, L, h); break;
        }
        if (ok) return true;
    }
    return false;
}

__global__ void __launch_bounds__(128) k_step2(const Params P) {
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= wstride) return;
    const int d = (int)(wi / P.cn);
    const int64_t il = wi - (int64_t)d * P.cn;
    const int64_t ig = P.off + il;
    const int64_t gi = (int64_t)d * P.n + ig;

    const uint32_t meta = P.wmeta[wi];
    const int kind = meta & M_KIND_MASK;
    const int src = P.wsrc[wi];
    ProfileSink out{P.prof, (int64_t)P.dof * P.n, gi};
    const double* w = P.ws + wi;
    const double ps = w[(int64_t)W_P0 * wstride], vs = w[(int64_t)W_V0 * wstride], as = w[(int64_t)W_A0 * wstride];

    if (src == RK_SRC_NONE || kind == K_NONE) {
        // disabled DoF (calculator_target.rs:313-323) or a trajectory that ended with an error status
        const DofIn x = load_dof(P, d, ig);
        store_idle_profile(out, x.p0, x.v0, x.a0);
        P.codes[gi] = pack_codes(L_NONE, UDDU, DIR_UP, RK_SRC_NONE);
        return;
    }

    double t[7], ctrl, ctrl2;
    int lim, cs, dir, lim_check, store_kind = kind;
    const DofIn x = load_dof(P, d, ig);
    bool pf_from_p7 = false;

    if (src == RK_SRC_STEP2) {
        const double duration = P.duration[ig];
        const double tf = duration - w[(int64_t)W_BRAKE_DUR * wstride] - 0.0;
        const Bound b{ps, vs, as, x.pf, x.vf, x.af};
        bool ok = false;
        if (kind == K_POS3) {
            Hit h;
            ok = step2_pos3(b, tf, x.v_max, x.v_min, x.a_max, x.a_min, x.j_max, h);
            copy_t(t, h.t);
            ctrl = h.jf; ctrl2 = 0.0; lim = h.lim; cs = h.cs; dir = h.dir;
        } else {
            Cand c;
            if (kind == K_POS2) { ok = step2_pos2(b, tf, x.v_max, x.v_min, x.a_max, x.a_min, c); pf_from_p7 = true; }
            else if (kind == K_POS1) ok = step2_pos1(b, tf, c);
            else if (kind == K_VEL3) { ok = step2_vel3(b, tf, x.a_max, x.a_min, x.j_max, c); pf_from_p7 = true; }
            else { ok = step2_vel2(b, tf, x.a_max, x.a_min, c); pf_from_p7 = true; }
            copy_t(t, c.t);
            ctrl = c.ctrl; ctrl2 = c.ctrl2; lim = c.lim; cs = c.cs; dir = c.dir;
        }
        if (!ok) {  // :819-826 — every failing DoF stores the same value; the lowest index is reported in the detail word
            P.status[ig] = RK_RESULT_ERROR_EXECUTION_TIME_CALCULATION;
            const int mine = (3 << 8) | d;
            int seen = P.detail[ig];
            while (seen == 0 || mine < seen) {
                const int prev = atomicCAS(&P.detail[ig], seen, mine);
                if (prev == seen) break;
                seen = prev;
            }
            store_idle_profile(out, 0.0, 0.0, 0.0);
            P.codes[gi] = pack_codes(L_NONE, UDDU, DIR_UP, RK_SRC_STEP2);
            return;
        }
        lim_check = lim;
    } else {
        int slot = src;
        if (slot == RK_SRC_PHASE) slot = 2;
        uint32_t c;
        ws_get_cand(P, wi, wstride, slot, t, ctrl, ctrl2, c);
        lim = c & 0xFF; cs = (c >> 8) & 0xFF; dir = (c >> 16) & 0xFF;
        lim_check = lim;
        if (src == RK_SRC_PHASE) {  // validated with ReachedLimits::None, labelled with the limiting DoF's limits (Q25)
            lim_check = L_NONE;
            store_kind = (int)(c >> 24) - 1;
        }
    }

    const double p7 = store_profile(out, store_kind, t, ctrl, ctrl2, cs, lim_check, ps, vs, as, x.vf, x.af);
    if (pf_from_p7) out.put(S_PF, p7);
    P.codes[gi] = pack_codes(lim, cs, dir, src);
}

// ------------------------------------------------------------------------------------------------ k_validate
__global__ void __launch_bounds__(128) k_validate(const Params P, uint8_t* valid, int32_t* reason) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    int r = 0;
    for (int d = 0; d < P.dof && r == 0; ++d) {
        const DofIn x = load_dof(P, d, i);
        const int c = validate_dof(x, P.ci[d], P.check_current != 0, P.check_target != 0);
        if (c) r = (c << 8) | d;
    }
    if (r == 0 && P.delta_time <= 0.0 && P.discrete) r = RK_INVALID_DELTA_TIME << 8;
    if (valid) valid[i] = r == 0 ? 1 : 0;
    if (reason) reason[i] = r;
}

// ------------------------------------------------------------------------------------------------ k_at_time
struct SampleParams {
    int64_t n;
    int32_t dof, steps;
    double time_start, delta_time;
    const double* prof;
    const double* duration;
    double *pos, *vel, *acc, *jerk;
    int32_t* section;
};

// trajectory.rs:52-189 for K sample times; time accumulates by repeated addition (ruckig.rs:293).
__global__ void __launch_bounds__(256) k_at_time(const SampleParams S) {
    const int64_t total = (int64_t)S.dof * S.n;
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= total) return;
    const int d = (int)(gi / S.n);
    const int64_t i = gi - (int64_t)d * S.n;
    const double* q = S.prof + gi;
    const int64_t st = total;

    const double duration = S.duration[i];
    double ts[7], pp[8], pv[8], pa[8], pj[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) { ts[k] = q[(int64_t)(S_TSUM + k) * st]; pj[k] = q[(int64_t)(S_J + k) * st]; }
#pragma unroll
    for (int k = 0; k < 8; ++k) { pp[k] = q[(int64_t)(S_P + k) * st]; pv[k] = q[(int64_t)(S_V + k) * st]; pa[k] = q[(int64_t)(S_A + k) * st]; }
    const double bdur = q[(int64_t)S_BRAKE_DUR * st];
    const double bt0 = q[(int64_t)S_BRAKE_T * st];
    double bp[2], bv[2], ba[2], bj[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        bp[k] = q[(int64_t)(S_BRAKE_P + k) * st]; bv[k] = q[(int64_t)(S_BRAKE_V + k) * st];
        ba[k] = q[(int64_t)(S_BRAKE_A + k) * st]; bj[k] = q[(int64_t)(S_BRAKE_J + k) * st];
    }

    double time = S.time_start;
#pragma unroll 1
    for (int k = 0; k < S.steps; ++k) {
        time += S.delta_time;
        double t0, p0, v0, a0, j0;
        int section;
        if (time >= duration) {  // :60-83
            section = 1;
            t0 = time - (bdur + ts[6]);
            p0 = pp[7]; v0 = pv[7]; a0 = pa[7]; j0 = 0.0;
        } else {
            // cumulative_times[0] == duration > time, so the section index is 0 and t_diff == time (:85-99, Q13)
            section = 0;
            double td = time;
            bool set = false;
            if (bdur > 0.0) {
                if (td < bdur) {
                    const int index = (td < bt0) ? 0 : 1;
                    if (index > 0) td -= bt0;
                    t0 = td; p0 = bp[index]; v0 = bv[index]; a0 = ba[index]; j0 = bj[index];
                    set = true;
                } else {
                    td -= bdur;
                }
            }
            if (!set) {
                if (td >= ts[6]) {
                    t0 = td - ts[6]; p0 = pp[7]; v0 = pv[7]; a0 = pa[7]; j0 = 0.0;
                } else {
                    int index = 6;
#pragma unroll
                    for (int m = 6; m >= 0; --m) if (ts[m] > td) index = m;
                    if (index > 0) td -= ts[index - 1];
                    t0 = td; p0 = pp[index]; v0 = pv[index]; a0 = pa[index]; j0 = pj[index];
                }
            }
        }
        double p, v, a;
        integrate(t0, p0, v0, a0, j0, p, v, a);
        const int64_t o = (int64_t)k * total + gi;
        if (S.pos) __stcs(S.pos + o, p);
        if (S.vel) __stcs(S.vel + o, v);
        if (S.acc) __stcs(S.acc + o, a);
        if (S.jerk) __stcs(S.jerk + o, j0);
        if (S.section && d == 0) S.section[(int64_t)k * S.n + i] = section;
    }
}

// ------------------------------------------------------------------------------------------------ FP64 peak probe
// 8 independent DFMA chains per thread, all in registers; `iters` x 8 x 2 flop per thread.
__global__ void __launch_bounds__(256) k_fp64_peak(double* sink, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
            x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) sink[0] = s;  // keeps the chains alive
}

}  // namespace rk

// ================================================================================================ host side / C ABI
namespace {

thread_local std::string g_last_error;

rk_status fail(rk_status s, const std::string& m) {
    g_last_error = m;
    return s;
}

#define RK_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            return fail(e_ == cudaErrorMemoryAllocation ? RK_ERR_OUT_OF_MEMORY : RK_ERR_CUDA,          \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                           \
        }                                                                                               \
    } while (0)

constexpr int64_t CHUNK_ITEMS = 8ll << 20;  // (DoF, trajectory) items per chunk of scratch: 8 Mi items = 2.3 GB

}  // namespace

struct rk_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    // scratch
    int64_t ws_items = 0;
    double* ws = nullptr;
    uint32_t* wmeta = nullptr;
    uint8_t* wsrc = nullptr;
    // staging for host-memory calls
    void* stage_dev = nullptr;
    size_t stage_dev_bytes = 0;
    void* stage_pinned = nullptr;
    size_t stage_pinned_bytes = 0;
};

struct rk_trajectories {
    rk_handle* h = nullptr;
    int64_t n = 0;
    int32_t dof = 0;
    double* prof = nullptr;
    uint32_t* codes = nullptr;
    int32_t* status = nullptr;
    double* duration = nullptr;
    double* imd = nullptr;
    int32_t* limiting = nullptr;
    int32_t* detail = nullptr;
};

static rk_status ensure_scratch(rk_handle* h, int64_t items) {
    if (items <= h->ws_items) return RK_OK;
    if (h->ws) { cudaFree(h->ws); cudaFree(h->wmeta); cudaFree(h->wsrc); h->ws = nullptr; h->ws_items = 0; }
    RK_CUDA(cudaMalloc(&h->ws, sizeof(double) * rk::WS_SLOTS * items));
    RK_CUDA(cudaMalloc(&h->wmeta, sizeof(uint32_t) * 4 * items));
    RK_CUDA(cudaMalloc(&h->wsrc, items));
    h->ws_items = items;
    return RK_OK;
}

static rk_status ensure_stage(rk_handle* h, size_t dev_bytes, size_t pinned_bytes) {
    if (dev_bytes > h->stage_dev_bytes) {
        if (h->stage_dev) cudaFree(h->stage_dev);
        h->stage_dev = nullptr; h->stage_dev_bytes = 0;
        RK_CUDA(cudaMalloc(&h->stage_dev, dev_bytes));
        h->stage_dev_bytes = dev_bytes;
    }
    if (pinned_bytes > h->stage_pinned_bytes) {
        if (h->stage_pinned) cudaFreeHost(h->stage_pinned);
        h->stage_pinned = nullptr; h->stage_pinned_bytes = 0;
        RK_CUDA(cudaMallocHost(&h->stage_pinned, pinned_bytes));
        h->stage_pinned_bytes = pinned_bytes;
    }
    return RK_OK;
}

static rk_status fill_params(rk::Params& P, const rk_batch_desc* desc, const rk_batch_in* in) {
    if (!desc || !in) return fail(RK_ERR_INVALID_ARGUMENT, "desc / in is NULL");
    if (desc->n < 0 || desc->dof < 1 || desc->dof > RK_MAX_DOF) return fail(RK_ERR_INVALID_ARGUMENT, "n < 0 or dof outside 1..RK_MAX_DOF");
    if (!in->current_position || !in->current_velocity || !in->current_acceleration || !in->target_position || !in->target_velocity
        || !in->target_acceleration || !in->max_velocity || !in->max_acceleration || !in->max_jerk)
        return fail(RK_ERR_INVALID_ARGUMENT, "a mandatory rk_batch_in array is NULL");
    std::memset(&P, 0, sizeof(P));
    P.n = desc->n;
    P.dof = desc->dof;
    P.discrete = desc->duration_discretization == RK_DISCRETE;
    P.throw_mode = desc->error_mode == RK_ERRORS_THROW;
    P.delta_time = desc->delta_time;
    for (int d = 0; d < desc->dof; ++d) {
        P.ci[d] = (int8_t)(desc->per_dof_control_interface ? desc->per_dof_control_interface[d] : desc->control_interface);
        P.sync[d] = (int8_t)(desc->per_dof_synchronization ? desc->per_dof_synchronization[d] : desc->synchronization);
    }
    P.p0 = in->current_position; P.v0 = in->current_velocity; P.a0 = in->current_acceleration;
    P.pf = in->target_position; P.vf = in->target_velocity; P.af = in->target_acceleration;
    P.v_max = in->max_velocity; P.a_max = in->max_acceleration; P.j_max = in->max_jerk;
    P.v_min = in->min_velocity; P.a_min = in->min_acceleration;
    P.enabled = in->enabled; P.min_dur = in->minimum_duration;
    return RK_OK;
}

// Copy the caller's host arrays into one device staging block and re-point P at it.
static rk_status stage_inputs(rk_handle* h, rk::Params& P, size_t extra_dev_bytes, char** extra_dev) {
    const size_t per = sizeof(double) * (size_t)P.dof * (size_t)P.n;
    const double** slots[11] = {&P.p0, &P.v0, &P.a0, &P.pf, &P.vf, &P.af, &P.v_max, &P.a_max, &P.j_max, &P.v_min, &P.a_min};
    size_t bytes = 0;
    for (auto s : slots) if (*s) bytes += per;
    const size_t en_bytes = P.enabled ? (size_t)P.dof * (size_t)P.n : 0;
    const size_t md_bytes = P.min_dur ? sizeof(double) * (size_t)P.n : 0;
    const size_t in_bytes = bytes + md_bytes + ((en_bytes + 255) / 256) * 256;
    rk_status st = ensure_stage(h, in_bytes + extra_dev_bytes, 0);
    if (st != RK_OK) return st;
    char* dev = (char*)h->stage_dev;
    for (auto s : slots) {
        if (!*s) continue;
        RK_CUDA(cudaMemcpyAsync(dev, *s, per, cudaMemcpyHostToDevice, h->stream));
        *s = (const double*)dev;
        dev += per;
    }
    if (P.min_dur) {
        RK_CUDA(cudaMemcpyAsync(dev, P.min_dur, md_bytes, cudaMemcpyHostToDevice, h->stream));
        P.min_dur = (const double*)dev;
        dev += md_bytes;
    }
    if (P.enabled) {
        RK_CUDA(cudaMemcpyAsync(dev, P.enabled, en_bytes, cudaMemcpyHostToDevice, h->stream));
        P.enabled = (const uint8_t*)dev;
        dev += ((en_bytes + 255) / 256) * 256;
    }
    if (extra_dev) *extra_dev = dev;
    return RK_OK;
}

extern "C" {

rk_status rk_create(int32_t device, rk_handle** out) {
    if (!out) return fail(RK_ERR_INVALID_ARGUMENT, "out is NULL");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail(RK_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= count) return fail(RK_ERR_INVALID_ARGUMENT, "device index out of range");
    RK_CUDA(cudaSetDevice(device));
    rk_handle* h = new rk_handle();
    h->device = device;
    RK_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    *out = h;
    return RK_OK;
}

rk_status rk_destroy(rk_handle* h) {
    if (!h) return RK_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->ws) { cudaFree(h->ws); cudaFree(h->wmeta); cudaFree(h->wsrc); }
    if (h->stage_dev) cudaFree(h->stage_dev);
    if (h->stage_pinned) cudaFreeHost(h->stage_pinned);
    cudaStreamDestroy(h->stream);
    delete h;
    return RK_OK;
}

rk_status rk_trajectories_create(rk_handle* h, int64_t n, int32_t dof, rk_trajectories** out) {
    if (!h || !out || n < 0 || dof < 1 || dof > RK_MAX_DOF) return fail(RK_ERR_INVALID_ARGUMENT, "bad arguments to rk_trajectories_create");
    RK_CUDA(cudaSetDevice(h->device));
    rk_trajectories* t = new rk_trajectories();
    t->h = h; t->n = n; t->dof = dof;
    const size_t items = (size_t)(n > 0 ? n : 1) * dof, nn = (size_t)(n > 0 ? n : 1);
    RK_CUDA(cudaMalloc(&t->prof, sizeof(double) * rk::PROFILE_SLOTS * items));
    RK_CUDA(cudaMalloc(&t->codes, sizeof(uint32_t) * items));
    RK_CUDA(cudaMalloc(&t->imd, sizeof(double) * items));
    RK_CUDA(cudaMalloc(&t->status, sizeof(int32_t) * nn));
    RK_CUDA(cudaMalloc(&t->duration, sizeof(double) * nn));
    RK_CUDA(cudaMalloc(&t->limiting, sizeof(int32_t) * nn));
    RK_CUDA(cudaMalloc(&t->detail, sizeof(int32_t) * nn));
    RK_CUDA(cudaMemsetAsync(t->prof, 0, sizeof(double) * rk::PROFILE_SLOTS * items, h->stream));
    RK_CUDA(cudaMemsetAsync(t->duration, 0, sizeof(double) * nn, h->stream));
    RK_CUDA(cudaMemsetAsync(t->status, 0, sizeof(int32_t) * nn, h->stream));
    *out = t;
    return RK_OK;
}

rk_status rk_trajectories_destroy(rk_trajectories* t) {
    if (!t) return RK_OK;
    cudaSetDevice(t->h->device);
    cudaStreamSynchronize(t->h->stream);
    cudaFree(t->prof); cudaFree(t->codes); cudaFree(t->imd); cudaFree(t->status); cudaFree(t->duration); cudaFree(t->limiting); cudaFree(t->detail);
    delete t;
    return RK_OK;
}

rk_status rk_calculate_batch(rk_handle* h, const rk_batch_desc* desc, const rk_batch_in* in, rk_trajectories* traj, const rk_batch_out* out) {
    if (!h || !traj) return fail(RK_ERR_INVALID_ARGUMENT, "handle / trajectories is NULL");
    rk::Params P;
    rk_status st = fill_params(P, desc, in);
    if (st != RK_OK) return st;
    if (traj->n != desc->n || traj->dof != desc->dof) return fail(RK_ERR_INVALID_ARGUMENT, "trajectories were created for another (n, dof)");
    if (desc->n == 0) return RK_OK;
    RK_CUDA(cudaSetDevice(h->device));
    const bool host = desc->memory_space == RK_MEM_HOST;
    if (host) {
        st = stage_inputs(h, P, 0, nullptr);
        if (st != RK_OK) return st;
    }
    P.prof = traj->prof; P.codes = traj->codes; P.status = traj->status; P.duration = traj->duration; P.imd = traj->imd;
    P.limiting = traj->limiting; P.detail = traj->detail;

    int64_t cn_max = CHUNK_ITEMS / desc->dof;
    if (cn_max > desc->n) cn_max = desc->n;
    st = ensure_scratch(h, cn_max * desc->dof);
    if (st != RK_OK) return st;
    P.ws = h->ws; P.wmeta = h->wmeta; P.wsrc = h->wsrc;

    for (int64_t off = 0; off < desc->n; off += cn_max) {
        P.off = off;
        P.cn = (desc->n - off < cn_max) ? (desc->n - off) : cn_max;
        const int64_t items = P.cn * P.dof;
        rk::k_step1<<<(unsigned)((items + 127) / 128), 128, 0, h->stream>>>(P);
        rk::k_sync<<<(unsigned)((P.cn + 127) / 128), 128, 0, h->stream>>>(P);
        rk::k_step2<<<(unsigned)((items + 127) / 128), 128, 0, h->stream>>>(P);
        h->launches += 3;
    }
    RK_CUDA(cudaGetLastError());

    if (out) {
        const cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
        if (out->status) RK_CUDA(cudaMemcpyAsync(out->status, traj->status, sizeof(int32_t) * desc->n, kind, h->stream));
        if (out->duration) RK_CUDA(cudaMemcpyAsync(out->duration, traj->duration, sizeof(double) * desc->n, kind, h->stream));
        if (out->independent_min_durations)
            RK_CUDA(cudaMemcpyAsync(out->independent_min_durations, traj->imd, sizeof(double) * desc->n * desc->dof, kind, h->stream));
    }
    if (host) RK_CUDA(cudaStreamSynchronize(h->stream));
    return RK_OK;
}

rk_status rk_validate_batch(rk_handle* h, const rk_batch_desc* desc, const rk_batch_in* in, int32_t check_current, int32_t check_target,
                            uint8_t* valid, int32_t* reason) {
    if (!h) return fail(RK_ERR_INVALID_ARGUMENT, "handle is NULL");
    rk::Params P;
    rk_status st = fill_params(P, desc, in);
    if (st != RK_OK) return st;
    if (desc->n == 0) return RK_OK;
    RK_CUDA(cudaSetDevice(h->device));
    P.check_current = check_current; P.check_target = check_target;
    const bool host = desc->memory_space == RK_MEM_HOST;
    uint8_t* d_valid = valid;
    int32_t* d_reason = reason;
    if (host) {
        char* extra = nullptr;
        st = stage_inputs(h, P, (size_t)desc->n * 8, &extra);
        if (st != RK_OK) return st;
        d_reason = (int32_t*)extra;
        d_valid = (uint8_t*)(extra + (size_t)desc->n * 4);
    }
    rk::k_validate<<<(unsigned)((desc->n + 127) / 128), 128, 0, h->stream>>>(P, d_valid, d_reason);
    h->launches += 1;
    RK_CUDA(cudaGetLastError());
    if (host) {
        if (valid) RK_CUDA(cudaMemcpyAsync(valid, d_valid, (size_t)desc->n, cudaMemcpyDeviceToHost, h->stream));
        if (reason) RK_CUDA(cudaMemcpyAsync(reason, d_reason, (size_t)desc->n * 4, cudaMemcpyDeviceToHost, h->stream));
        RK_CUDA(cudaStreamSynchronize(h->stream));
    }
    return RK_OK;
}

rk_status rk_at_time_batch(rk_handle* h, const rk_trajectories* traj, const rk_sample_desc* desc, const rk_sample_out* out) {
    if (!h || !traj || !desc || !out) return fail(RK_ERR_INVALID_ARGUMENT, "NULL argument to rk_at_time_batch");
    if (desc->steps < 0) return fail(RK_ERR_INVALID_ARGUMENT, "steps < 0");
    if (desc->steps == 0 || traj->n == 0) return RK_OK;
    RK_CUDA(cudaSetDevice(h->device));
    rk::SampleParams S;
    S.n = traj->n; S.dof = traj->dof; S.steps = desc->steps;
    S.time_start = desc->time_start; S.delta_time = desc->delta_time;
    S.prof = traj->prof; S.duration = traj->duration;
    const bool host = desc->memory_space == RK_MEM_HOST;
    const size_t per = sizeof(double) * (size_t)desc->steps * traj->dof * traj->n;
    const size_t sec = sizeof(int32_t) * (size_t)desc->steps * traj->n;
    double* dst[4] = {out->position, out->velocity, out->acceleration, out->jerk};
    double* dev[4] = {out->position, out->velocity, out->acceleration, out->jerk};
    int32_t* dsec = out->section;
    if (host) {
        size_t bytes = 0;
        for (int k = 0; k < 4; ++k) if (dst[k]) bytes += per;
        if (out->section) bytes += sec;
        rk_status st = ensure_stage(h, bytes, 0);
        if (st != RK_OK) return st;
        char* p = (char*)h->stage_dev;
        for (int k = 0; k < 4; ++k) if (dst[k]) { dev[k] = (double*)p; p += per; }
        if (out->section) dsec = (int32_t*)p;
    }
    S.pos = dev[0]; S.vel = dev[1]; S.acc = dev[2]; S.jerk = dev[3]; S.section = dsec;
    const int64_t total = (int64_t)traj->dof * traj->n;
    rk::k_at_time<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(S);
    h->launches += 1;
    RK_CUDA(cudaGetLastError());
    if (host) {
        for (int k = 0; k < 4; ++k) if (dst[k]) RK_CUDA(cudaMemcpyAsync(dst[k], dev[k], per, cudaMemcpyDeviceToHost, h->stream));
        if (out->section) RK_CUDA(cudaMemcpyAsync(out->section, dsec, sec, cudaMemcpyDeviceToHost, h->stream));
        RK_CUDA(cudaStreamSynchronize(h->stream));
    }
    return RK_OK;
}

rk_status rk_trajectories_read(rk_handle* h, const rk_trajectories* t, int32_t field, void* dst, int32_t dst_memory_space) {
    if (!h || !t || !dst) return fail(RK_ERR_INVALID_ARGUMENT, "NULL argument to rk_trajectories_read");
    RK_CUDA(cudaSetDevice(h->device));
    const size_t items = (size_t)t->n * t->dof;
    const void* src = nullptr;
    size_t bytes = 0;
    auto slots = [&](int first, int count) { src = t->prof + (size_t)first * items; bytes = sizeof(double) * count * items; };
    switch (field) {
        case RK_FIELD_STATUS: src = t->status; bytes = sizeof(int32_t) * t->n; break;
        case RK_FIELD_DURATION: src = t->duration; bytes = sizeof(double) * t->n; break;
        case RK_FIELD_INDEPENDENT_MIN_DURATIONS: src = t->imd; bytes = sizeof(double) * items; break;
        case RK_FIELD_LIMITING_DOF: src = t->limiting; bytes = sizeof(int32_t) * t->n; break;
        case RK_FIELD_ERROR_DETAIL: src = t->detail; bytes = sizeof(int32_t) * t->n; break;
        case RK_FIELD_CODES: src = t->codes; bytes = sizeof(uint32_t) * items; break;
        case RK_FIELD_T: slots(rk::S_T, 7); break;
        case RK_FIELD_T_SUM: slots(rk::S_TSUM, 7); break;
        case RK_FIELD_J: slots(rk::S_J, 7); break;
        case RK_FIELD_A: slots(rk::S_A, 8); break;
        case RK_FIELD_V: slots(rk::S_V, 8); break;
        case RK_FIELD_P: slots(rk::S_P, 8); break;
        case RK_FIELD_BRAKE_DURATION: slots(rk::S_BRAKE_DUR, 1); break;
        case RK_FIELD_BRAKE_T: slots(rk::S_BRAKE_T, 2); break;
        case RK_FIELD_BRAKE_J: slots(rk::S_BRAKE_J, 2); break;
        case RK_FIELD_BRAKE_A: slots(rk::S_BRAKE_A, 2); break;
        case RK_FIELD_BRAKE_V: slots(rk::S_BRAKE_V, 2); break;
        case RK_FIELD_BRAKE_P: slots(rk::S_BRAKE_P, 2); break;
        case RK_FIELD_PF_VF_AF: slots(rk::S_PF, 3); break;
        default: return fail(RK_ERR_INVALID_ARGUMENT, "unknown field");
    }
    if (bytes == 0) return RK_OK;
    RK_CUDA(cudaMemcpyAsync(dst, src, bytes, dst_memory_space == RK_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, h->stream));
    if (dst_memory_space == RK_MEM_HOST) RK_CUDA(cudaStreamSynchronize(h->stream));
    return RK_OK;
}

rk_status rk_sync(rk_handle* h) {
    if (!h) return fail(RK_ERR_INVALID_ARGUMENT, "handle is NULL");
    RK_CUDA(cudaSetDevice(h->device));
    RK_CUDA(cudaStreamSynchronize(h->stream));
    return RK_OK;
}

rk_status rk_measure_fp64_peak(rk_handle* h, double milliseconds, double* tflops) {
    if (!h || !tflops) return fail(RK_ERR_INVALID_ARGUMENT, "NULL argument to rk_measure_fp64_peak");
    RK_CUDA(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    RK_CUDA(cudaGetDeviceProperties(&prop, h->device));
    rk_status st = ensure_stage(h, 256, 0);
    if (st != RK_OK) return st;
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    cudaEvent_t e0, e1;
    RK_CUDA(cudaEventCreate(&e0));
    RK_CUDA(cudaEventCreate(&e1));
    int iters = 2000;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        RK_CUDA(cudaEventRecord(e0, h->stream));
        rk::k_fp64_peak<<<blocks, threads, 0, h->stream>>>((double*)h->stage_dev, iters, 0.999999, 1e-9);
        RK_CUDA(cudaEventRecord(e1, h->stream));
        RK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        RK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep >= 2 && tf > best) best = tf;
        if (rep == 1 && ms > 0.f) iters = (int)(iters * (milliseconds > 0.0 ? milliseconds : 20.0) / ms) + 1;  // size the remaining runs
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    h->launches += 6;
    *tflops = best;
    return RK_OK;
}

void* rk_stream(rk_handle* h) { return h ? (void*)h->stream : nullptr; }
int64_t rk_kernel_launches(const rk_handle* h) { return h ? h->launches : 0; }

const char* rk_status_string(rk_status s) {
    switch (s) {
        case RK_OK: return "RK_OK";
        case RK_ERR_INVALID_ARGUMENT: return "RK_ERR_INVALID_ARGUMENT";
        case RK_ERR_CUDA: return "RK_ERR_CUDA";
        case RK_ERR_NO_DEVICE: return "RK_ERR_NO_DEVICE";
        case RK_ERR_OUT_OF_MEMORY: return "RK_ERR_OUT_OF_MEMORY";
        case RK_ERR_UNSUPPORTED: return "RK_ERR_UNSUPPORTED";
        default: return "RK_ERR_UNKNOWN";
    }
}
const char* rk_last_error(void) { return g_last_error.c_str(); }
const char* rk_version(void) { return "ruckig_b200 0.1.0 (sm_100a; parity target rsruckig 2.1.2)"; }

}  // extern "C"
