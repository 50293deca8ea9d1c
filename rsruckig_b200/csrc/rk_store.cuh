// rk_store.cuh — HBM layout of the batch and the routines that expand a compact
// candidate (t[7] + control value + codes) into the full Profile arrays the
// reference keeps (profile.rs:76-118), bit-identical to what its check*
// functions leave behind.
//
// Layout: everything is structure-of-arrays with the trajectory index fastest.
//   profile field slot s of DoF d, trajectory i : prof[(s * dof + d) * n + i]
//   slots: T 0..6 | T_SUM 7..13 | J 14..20 | A 21..28 | V 29..36 | P 37..44 |
//          BRAKE_DURATION 45 | BRAKE_T 46,47 | BRAKE_J 48,49 | BRAKE_A 50,51 |
//          BRAKE_V 52,53 | BRAKE_P 54,55 | PF,VF,AF 56..58           (59 slots)
// A warp handles 32 consecutive trajectories of one DoF, so every load and
// store below is a fully coalesced 256-byte row.
// Included once per pass: rk_kernels.cu compiles these headers twice — into namespace rk (exact divisions with their fall-back
// branch) and, with RK_PASS2 defined and `rk` renamed, into namespace rk_lazy (rk_math.cuh, RK_LAZY_DDIV).
#ifndef RK_PASS2
#ifndef RK_H_STORE_1
#define RK_H_STORE_1
#define RK_H_STORE_GO
#endif
#else
#ifndef RK_H_STORE_2
#define RK_H_STORE_2
#define RK_H_STORE_GO
#endif
#endif
#ifdef RK_H_STORE_GO
#undef RK_H_STORE_GO
#include "rk_step1.cuh"

namespace rk {

enum : int { S_T = 0, S_TSUM = 7, S_J = 14, S_A = 21, S_V = 29, S_P = 37, S_BRAKE_DUR = 45, S_BRAKE_T = 46, S_BRAKE_J = 48, S_BRAKE_A = 50,
             S_BRAKE_V = 52, S_BRAKE_P = 54, S_PF = 56, S_VF = 57, S_AF = 58, PROFILE_SLOTS = 59 };

// Which reference solver family an item belongs to (calculator_target.rs:389-452).
enum : int { K_POS3 = 0, K_POS2 = 1, K_POS1 = 2, K_VEL3 = 3, K_VEL2 = 4, K_NONE = 5 };

// Scratch written by the Step-1 kernel (ws doubles, slot-major like the profiles).
enum : int { W_P0 = 0, W_V0 = 1, W_A0 = 2, W_TMIN = 3, W_ALEFT = 4, W_ARIGHT = 5, W_BLEFT = 6, W_BRIGHT = 7, W_CAND = 8, W_CAND_STRIDE = 9,
             W_BRAKE_DUR = 35,
             WS_SLOTS = 36,
             // Step-1 candidate sub-lists: 36 slots per item in their own array of structures (Params::vl), slot = t[7] + duration
             VL_SLOTS = 36, VL_SLOT_DOUBLES = 8, VL_ITEM_DOUBLES = 36 * 8 };
// Step 2, VEL family: a chain item re-uses its own (no longer needed) sub-list area for the polynomial, the root brackets and
// the polished roots that travel between the scan, polish and check kernels (double offsets into the item's vl area).
enum : int { V_POL = 0, V_LO = 8, V_HI = 16, V_ROOT = 24, V_TZ = 32, V1_LO = 40, V1_HI = 48, V1_ROOT = 56,
             V_SOL = 64 };  // the compact Step-2 solution of the item, t[7] + jerk: one 64-byte record (put_solution)
// meta words per item: [0] flags below, [1..3] codes of the three block profiles
enum : int { WM_META = 0, WM_CAND = 1, WM_WORDS = 4 };
enum : uint32_t { M_KIND_MASK = 0xFu, M_FOUND = 1u << 4, M_HAS_A = 1u << 5, M_HAS_B = 1u << 6, M_S1_PENDING = 1u << 7, M_ZERO_LIMITS = 1u << 8,
                  M_ENABLED = 1u << 9,
                  // what k1_block needs of a pending item's 96-byte record, as five bits (set by the set-up kernels): target at
                  // rest, nothing moves at all, pf >= p0 (post-brake), the signs of the velocity limits
                  M_REST_TARGET = 1u << 10, M_TRIVIAL = 1u << 11, M_PD_NONNEG = 1u << 12, M_VMAX_POS = 1u << 13, M_VMIN_POS = 1u << 14,
                  M_VCODE_SHIFT = 16 };

struct ProfileSink {
    double* prof;  // base of the trajectory storage
    int64_t stride;  // dof * n
    int64_t idx;     // d * n + i
    __device__ __forceinline__ void put(int slot, double v) const { prof[(int64_t)slot * stride + idx] = v; }
};

__device__ __forceinline__ uint32_t pack_codes(int lim, int cs, int dir, int src) {
    return (uint32_t)lim | ((uint32_t)cs << 8) | ((uint32_t)dir << 16) | ((uint32_t)src << 24);
}

// Expand and store one profile. `lim_check` is the limits value the reference passed to its check function (it decides
// whether a[3] is forced to zero); the label stored in the codes may differ after phase synchronisation (Q25).
// Returns p[7] (the velocity-interface / second-order Step 2 sets pf = p[7] afterwards).
__device__ __forceinline__ double store_profile(const ProfileSink& o, int kind, const double (&t)[7], double ctrl, double ctrl2, int cs, int lim_check,
                                                double p0, double v0, double a0, double vf, double af) {
    double ts[7], j[7], a[8], v[8], p[8];
    a[0] = a0; v[0] = v0; p[0] = p0;
    if (kind == K_POS3 || kind == K_VEL3) {
        ts[0] = t[0];
#pragma unroll
        for (int i = 0; i < 6; ++i) ts[i + 1] = ts[i] + t[i + 1];
        if (kind == K_POS3) {
            jerk_pattern_pos(t, cs, ctrl, j);
        } else {  // profile.rs:146-170
            j[0] = (t[0] > 0.0) ? ctrl : 0.0; j[1] = 0.0; j[2] = (t[2] > 0.0) ? -ctrl : 0.0; j[3] = 0.0; j[5] = 0.0;
            if (cs == UDDU) { j[4] = (t[4] > 0.0) ? -ctrl : 0.0; j[6] = (t[6] > 0.0) ? ctrl : 0.0; }
            else { j[4] = (t[4] > 0.0) ? ctrl : 0.0; j[6] = (t[6] > 0.0) ? ctrl : 0.0; }
        }
        const bool zero_a3 = (kind == K_POS3) && forces_a3_zero(lim_check);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const double tj = t[i] * j[i];  // as in check_pos3: +-0 / 6 == +-0 for the always-zero jerks
            a[i + 1] = a[i] + tj;
            v[i + 1] = v[i] + t[i] * (a[i] + tj / 2.0);
            p[i + 1] = p[i] + t[i] * (v[i] + t[i] * (a[i] / 2.0 + ((i & 1) ? tj : div6(tj))));
            if (i == 2 && zero_a3) a[3] = 0.0;
        }
    } else if (kind == K_POS2) {  // profile.rs:549-637
        ts[0] = t[0];
#pragma unroll
        for (int i = 0; i < 6; ++i) ts[i + 1] = ts[i] + t[i + 1];
#pragma unroll
        for (int i = 0; i < 7; ++i) j[i] = 0.0;
        a[0] = (t[0] > 0.0) ? ctrl : 0.0; a[1] = 0.0; a[2] = (t[2] > 0.0) ? ctrl2 : 0.0; a[3] = 0.0; a[5] = 0.0;
        if (cs == UDDU) { a[4] = (t[4] > 0.0) ? ctrl2 : 0.0; a[6] = (t[6] > 0.0) ? ctrl : 0.0; }
        else { a[4] = (t[4] > 0.0) ? ctrl : 0.0; a[6] = (t[6] > 0.0) ? ctrl2 : 0.0; }
        a[7] = af;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            v[i + 1] = v[i] + t[i] * a[i];
            p[i + 1] = p[i] + t[i] * (v[i] + t[i] * a[i] / 2.0);
        }
    } else if (kind == K_VEL2) {  // profile.rs:256-294
        ts[0] = 0.0;
#pragma unroll
        for (int i = 1; i < 7; ++i) ts[i] = t[1];
#pragma unroll
        for (int i = 0; i < 7; ++i) j[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = 0.0;
        a[1] = (t[1] > 0.0) ? ctrl : 0.0;
        a[7] = af;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            v[i + 1] = v[i] + t[i] * a[i];
            p[i + 1] = p[i] + t[i] * (v[i] + t[i] * a[i] / 2.0);
        }
    } else {  // K_POS1, profile.rs:685-726
        ts[0] = 0.0; ts[1] = 0.0; ts[2] = 0.0; ts[3] = t[3]; ts[4] = t[3]; ts[5] = t[3]; ts[6] = t[3];
#pragma unroll
        for (int i = 0; i < 7; ++i) { j[i] = 0.0; a[i] = 0.0; v[i] = 0.0; }
        a[7] = af;
        v[3] = (t[3] > 0.0) ? ctrl : 0.0;
        v[7] = vf;
#pragma unroll
        for (int i = 0; i < 7; ++i) p[i + 1] = p[i] + t[i] * (v[i] + t[i] * a[i] / 2.0);
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) { o.put(S_T + i, t[i]); o.put(S_TSUM + i, ts[i]); o.put(S_J + i, j[i]); }
#pragma unroll
    for (int i = 0; i < 8; ++i) { o.put(S_A + i, a[i]); o.put(S_V + i, v[i]); o.put(S_P + i, p[i]); }
    return p[7];
}

// A profile the calculation never touched (disabled DoF, calculator_target.rs:313-323, or an error status): the fields of a
// fresh Profile::default() except the end state, which at_time extrapolates from (trajectory.rs:122-132).
__device__ __forceinline__ void store_idle_profile(const ProfileSink& o, double p7, double v7, double a7) {
#pragma unroll
    for (int i = 0; i < 7; ++i) { o.put(S_T + i, 0.0); o.put(S_TSUM + i, 0.0); o.put(S_J + i, 0.0); }
#pragma unroll
    for (int i = 0; i < 7; ++i) { o.put(S_A + i, 0.0); o.put(S_V + i, 0.0); o.put(S_P + i, 0.0); }
    o.put(S_A + 7, a7); o.put(S_V + 7, v7); o.put(S_P + 7, p7);
}

// ------------------------------------------------------------------------------------------------ compact result record
// RK_RESULT_COMPACT: 20 doubles per (DoF, trajectory) instead of the 59 of the full Profile — what the
// calculation actually decided (t[7], the control values, the brake pre-trajectory) plus the boundary states it started from.
// Every other Profile field is a deterministic function of these: expand_compact() below runs the very operations the
// full-layout path runs (store_profile, the brake integration of k1_setup), so the expanded arrays are bit-identical to
// what RK_RESULT_FULL stores. Same structure-of-arrays rows: slot s of DoF d, trajectory i : prof[(s * dof + d) * n + i].
//   T 0..6 | CTRL 7 (jerk / acceleration / velocity of the control pattern) | CTRL2 8 | P0 V0 A0 9..11 (the input's current
//   state = state before the brake pre-trajectory) | BRAKE_T 12,13 | BRAKE_J 14,15 (second-order brake: slot 14 holds
//   BrakeProfile::a[0], its jerks are zero) | PF VF AF 16..18 | META 19
enum : int { C_T = 0, C_CTRL = 7, C_CTRL2 = 8, C_P0 = 9, C_V0 = 10, C_A0 = 11, C_BRAKE_T = 12, C_BRAKE_J = 14, C_PF = 16, C_VF = 17, C_AF = 18,
             C_META = 19, COMPACT_SLOTS = 20 };
// META (an exactly representable small integer): bits 0..2 solver kind of the DoF (K_*: decides the brake variant), 3..5 kind
// the profile is expanded with (differs after phase synchronisation), 6..7 CM_IDLE_*, 8 a[3] forced to zero, 9 pf = p[7],
// 10 control signs
enum : int { CM_PROFILE = 0, CM_IDLE_STATE = 1, CM_IDLE_ZERO = 2 };
__device__ __forceinline__ double compact_meta(int kind, int store_kind, int idle, bool zero_a3, bool pf_from_p7, int cs) {
    return (double)(kind | (store_kind << 3) | (idle << 6) | ((zero_a3 ? 1 : 0) << 8) | ((pf_from_p7 ? 1 : 0) << 9) | ((cs & 1) << 10));
}

// Compact record -> the 59 fields of the full layout, written through `o` (a stride-1 local array for the readers).
__device__ __forceinline__ void expand_compact(const double (&r)[COMPACT_SLOTS], const ProfileSink& o) {
    const int m = (int)r[C_META];
    const int kind = m & 7, store_kind = (m >> 3) & 7, idle = (m >> 6) & 3, cs = (m >> 10) & 1;
    const bool zero_a3 = (m >> 8) & 1, pf_from_p7 = (m >> 9) & 1;
    // brake pre-trajectory: the integration of k1_setup (brake.rs:195-235)
    const double bt0 = r[C_BRAKE_T], bt1 = r[C_BRAKE_T + 1];
    const double bj0 = (kind == K_POS2) ? 0.0 : r[C_BRAKE_J], bj1 = r[C_BRAKE_J + 1];
    double bdur = 0.0, ba0 = 0.0, ba1 = 0.0, bv0 = 0.0, bv1 = 0.0, bp0 = 0.0, bp1 = 0.0;
    double ps = r[C_P0], vs = r[C_V0], as = r[C_A0];
    if (kind == K_POS3 || kind == K_VEL3) {
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
    } else if (kind == K_POS2) {
        ba0 = r[C_BRAKE_J];
        if (!(bt0 <= 0.0)) {
            bdur = bt0;
            bp0 = ps; bv0 = vs;
            integrate(bt0, ps, vs, ba0, 0.0, ps, vs, as);
        }
    }
    o.put(S_BRAKE_DUR, bdur);
    o.put(S_BRAKE_T, bt0); o.put(S_BRAKE_T + 1, bt1);
    o.put(S_BRAKE_J, bj0); o.put(S_BRAKE_J + 1, bj1);
    o.put(S_BRAKE_A, ba0); o.put(S_BRAKE_A + 1, ba1);
    o.put(S_BRAKE_V, bv0); o.put(S_BRAKE_V + 1, bv1);
    o.put(S_BRAKE_P, bp0); o.put(S_BRAKE_P + 1, bp1);
    o.put(S_PF, r[C_PF]); o.put(S_VF, r[C_VF]); o.put(S_AF, r[C_AF]);
    if (idle == CM_IDLE_STATE) { store_idle_profile(o, r[C_P0], r[C_V0], r[C_A0]); return; }
    if (idle == CM_IDLE_ZERO) { store_idle_profile(o, 0.0, 0.0, 0.0); return; }
    double t[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) t[i] = r[C_T + i];
    // store_profile only asks forces_a3_zero(lim_check): L_VEL forces it, L_NONE does not
    const double p7 = store_profile(o, store_kind, t, r[C_CTRL], r[C_CTRL2], cs, zero_a3 ? (int)L_VEL : (int)L_NONE, ps, vs, as, r[C_VF], r[C_AF]);
    if (pf_from_p7) o.put(S_PF, p7);
}

}  // namespace rk
#endif  // RK_H_STORE_GO
