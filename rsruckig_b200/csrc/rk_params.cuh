// rk_params.cuh — kernel parameter block, per-DoF input loading, input validation and the scratch (workspace) accessors
// shared by all kernels of the pipeline.
// Included once per pass: rk_kernels.cu compiles these headers twice — into namespace rk (exact divisions with their fall-back
// branch) and, with RK_PASS2 defined and `rk` renamed, into namespace rk_lazy (rk_math.cuh, RK_LAZY_DDIV).
#ifndef RK_PASS2
#ifndef RK_H_PARAMS_1
#define RK_H_PARAMS_1
#define RK_H_PARAMS_GO
#endif
#else
#ifndef RK_H_PARAMS_2
#define RK_H_PARAMS_2
#define RK_H_PARAMS_GO
#endif
#endif
#ifdef RK_H_PARAMS_GO
#undef RK_H_PARAMS_GO
#include <cuda_runtime.h>

#include "../../include/ruckig_b200.h"
#include "rk_step1.cuh"
#include "rk_store.cuh"

namespace rk {

#define RK_N_COUNTERS 112

struct Params {
    int64_t n;        // trajectories in the batch (stride of the trajectory storage)
    int64_t off;      // first trajectory of this chunk
    int64_t in_n;     // stride of the input arrays: n for caller-resident device arrays, cn for a staged chunk
    int64_t in_shift; // input index of trajectory ig is ig + in_shift (0, or -off for a staged chunk)
    int64_t cn;       // trajectories in this chunk (stride of the scratch arrays)
    int32_t dof;
    int32_t discrete;
    int32_t throw_mode;
    int32_t check_current, check_target;
    int32_t compact;  // trajectory storage layout: 0 = full Profile (59 slots), 1 = compact record (COMPACT_SLOTS, rk_store.cuh)
    int32_t lim_per_dof;  // the five limit arrays are [dof] (RK_LIMITS_PER_DOF) instead of [dof][n]
    const void* self;     // this parameter block in global memory (written by the set-up kernel): what the out-of-line exact redo paths of the lazy kernels work from — a kernel parameter passed on by reference would be copied to local memory by every thread
    int32_t force_redo;   // test hook (RK_B200_FORCE_REDO=1): the lazy kernels redo a quarter of their work items through the exact path
    double delta_time;
    int8_t ci[RK_MAX_DOF];
    int8_t sync[RK_MAX_DOF];
    // inputs, [dof][n]
    const double *p0, *v0, *a0, *pf, *vf, *af, *v_max, *a_max, *j_max, *v_min, *a_min, *min_dur;
    const uint8_t* enabled;
    // scratch, chunk-local
    double* ws;       // [WS_SLOTS][dof][cn]
    double* vl;       // [dof * cn][36][8] Step-1 candidate slots (t[7] + duration), array of structures: 64-byte slots
    double2* rec;     // [dof * cn][6] packed per-item record written by k1_setup (load_rec): one 96-byte read instead of 14 sectors
    uint32_t* wmeta;  // [WM_WORDS][dof][cn]
    uint8_t* wsrc;    // [dof][cn]
    uint32_t* wqmask; // [10][dof][cn] valid-slot masks (one byte per slot) of the Step-1 candidate sub-lists
    uint2* rq_item;   // work queue of the two-phase Step-1 kernels: (item, direction | rank << 1) / (item, direction | sub << 1 | slot << 8)
    double* rq_root;  //   ... and the root itself (quartics); 8 resp. 12 entries per item at most
    uint32_t* list[2];   // ping-pong work lists of item indices for the Step-2 stages, [dof * cn] each
    uint32_t* pend;      // VEL stage: items whose root candidates are being polished, [dof * cn]
    uint32_t* counters;  // [0..18] list lengths per Step-2 stage, [20..22] quartic root queues, [24] rescue list, [32 + 2 * step], [33 + 2 * step] pending items / polish tasks of VEL stage `step`
    // trajectory storage
    double* prof;     // [PROFILE_SLOTS or COMPACT_SLOTS][dof][n]
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
    const int64_t g = (int64_t)d * P.in_n + (ig + P.in_shift);
    DofIn x;
    x.p0 = P.p0[g]; x.v0 = P.v0[g]; x.a0 = P.a0[g];
    x.pf = P.pf[g]; x.vf = P.vf[g]; x.af = P.af[g];
    const int64_t gl = P.lim_per_dof ? (int64_t)d : g;
    x.v_max = P.v_max[gl]; x.a_max = P.a_max[gl]; x.j_max = P.j_max[gl];
    x.v_min = P.v_min ? P.v_min[gl] : -x.v_max;
    x.a_min = P.a_min ? P.a_min[gl] : -x.a_max;
    x.enabled = P.enabled ? (P.enabled[g] != 0) : true;
    return x;
}

// The per-item record every kernel after k1_setup works from: post-brake boundary state, direction-neutral limits, brake
// duration. Packed (array of structures, 6 x 16 bytes) because most readers run over compacted lists, where a
// structure-of-arrays read costs one 32-byte sector per 8-byte value.
__device__ __forceinline__ void store_rec(const Params& P, int64_t wi, double ps, double vs, double as, const DofIn& x, double bdur) {
    double2* r = P.rec + wi * 6;
    r[0] = make_double2(ps, vs);
    r[1] = make_double2(as, x.pf);
    r[2] = make_double2(x.vf, x.af);
    r[3] = make_double2(x.v_max, x.v_min);
    r[4] = make_double2(x.a_max, x.a_min);
    r[5] = make_double2(x.j_max, bdur);
}

// x.p0 / v0 / a0 of the returned DofIn are the POST-brake values (b.p0 ...), bdur the brake duration.
__device__ __forceinline__ DofIn load_rec(const Params& P, int64_t wi, Bound& b, double& bdur) {
    const double2* r = P.rec + wi * 6;
    const double2 r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3], r4 = r[4], r5 = r[5];
    DofIn x;
    x.p0 = r0.x; x.v0 = r0.y; x.a0 = r1.x;
    x.pf = r1.y; x.vf = r2.x; x.af = r2.y;
    x.v_max = r3.x; x.v_min = r3.y; x.a_max = r4.x; x.a_min = r4.y; x.j_max = r5.x;
    x.enabled = true;
    b = Bound{r0.x, r0.y, r1.x, r1.y, r2.x, r2.y};
    bdur = r5.y;
    return x;
}

// Candidate slot `slot` of item wi: t[0..6], duration (t_sum6) in [7].
__device__ __forceinline__ double* vl_slot(const Params& P, int64_t wi, int slot) {
    return P.vl + (wi * VL_SLOTS + slot) * VL_SLOT_DOUBLES;
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
    P.wmeta[(int64_t)(WM_CAND + slot) * wstride + wi] = pack_codes(c.lim, c.cs, c.dir, 0);
}

// Read block profile `slot` (0 p_min, 1 a.profile, 2 b.profile). Bits 24..31 of its code word: 0 = the profile was copied
// into the W_CAND slots, r > 0 = it still sits in Step-1 sub-list slot r - 1 (third-order position interface; the control
// value is then j_max if bit 23 is set, else -j_max). Bits 12..15 carry the kind of a phase-synchronised profile (k_sync).
__device__ __forceinline__ void ws_get_cand(const Params& P, int64_t wi, int64_t wstride, int slot, double j_max, double (&t)[7], double& ctrl, double& ctrl2,
                                            uint32_t& codes) {
    codes = P.wmeta[(int64_t)(WM_CAND + slot) * wstride + wi];
    const int ref = (int)(codes >> 24);
    if (ref == 0) {
        const double* base = P.ws + (int64_t)(W_CAND + slot * W_CAND_STRIDE) * wstride + wi;
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = base[(int64_t)k * wstride];
        ctrl = base[7 * wstride];
        ctrl2 = base[8 * wstride];
    } else {
        const double* base = vl_slot(P, wi, ref - 1);
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = base[k];
        ctrl = ((codes >> 23) & 1u) ? j_max : -j_max;
        ctrl2 = 0.0;
    }
}


}  // namespace rk
#endif  // RK_H_PARAMS_GO
