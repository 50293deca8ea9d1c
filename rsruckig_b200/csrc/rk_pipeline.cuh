// rk_pipeline.cuh — the calculation as a pipeline of small kernels.
//
// The monolithic "one thread does everything for its (DoF, trajectory)" form of this code ran at ~2 % FP64 utilisation:
// 300-390 KB of SASS per kernel thrashed the instruction cache (ncu: `no_instruction` was the top stall) and warps
// averaged 5-7 active lanes. The pipeline below keeps every kernel's code small and every warp on one solver family:
//
//   k1_setup            validation, brake pre-trajectory, the packed per-item record, zero-limit and non-third-order Step 1
//   k1_quartic_roots<T> the three quartic families of Step 1 (NONE, ACC0, ACC1), one thread per (direction, DoF, trajectory):
//     + _check<T>       coefficients + closed-form roots -> root queue -> Newton polish + profile check with full warps
//   k1_closed_cands     the closed-form families (ACC0_ACC1, four VEL variants): screened t[7] candidates -> queue ->
//     + k1_cand_check   profile check with full warps; valid candidates set a byte of their sub-list's mask word
//   k1_block            masks + stored durations -> block of feasible durations (block.rs) on slot references; k1_rescue for the
//                       rare items without any candidate
//   k_sync              one thread per trajectory: status, common duration, limiting DoF, phase sync, per-DoF decision,
//                       list of the items that need the Step-2 chain
//   k2_stage<F>         stage s of the chain (position_third_step2.rs:2449-2482) over the compacted list of items that are
//                       still unsolved; a solved item leaves t[7] + jerk + codes, the rest move to the next list.
//                       VEL/UDDU = k2_vel_scan_uddu -> k2_polish -> k2_vel_check_uddu, VEL/UDUD = k2_vel<8> (lockstep)
//   k2_fail             items that survive all 18 stages -> ErrorExecutionTimeCalculation
//   k_expand            every item: compact result -> the full profile arrays, coalesced rows
//
// Lists are compacted with warp-aggregated atomics; ordering inside a list does not matter (items are independent).
// Included once per pass: rk_kernels.cu compiles these headers twice — into namespace rk (exact divisions with their fall-back
// branch) and, with RK_PASS2 defined and `rk` renamed, into namespace rk_lazy (rk_math.cuh, RK_LAZY_DDIV).
#ifndef RK_PASS2
#ifndef RK_H_PIPELINE_1
#define RK_H_PIPELINE_1
#define RK_H_PIPELINE_GO
#endif
#else
#ifndef RK_H_PIPELINE_2
#define RK_H_PIPELINE_2
#define RK_H_PIPELINE_GO
#endif
#endif
#ifdef RK_H_PIPELINE_GO
#undef RK_H_PIPELINE_GO
#include "rk_params.cuh"
#include "rk_step_misc.cuh"
#include "rk_step2.cuh"
#include "rk_step2_vel.cuh"
#include "rk_step2_acc.cuh"
#include "rk_step2_none.cuh"

// Programmatic dependent launch: the pipeline is 36 short dependent kernels per chunk, so the launch latency between them
// shows in the latency of a small batch. With RK_PDL = 1 every launch carries the programmatic-stream-serialization
// attribute and every kernel starts with griddepcontrol.wait (= cudaGridDependencySynchronize: returns when the preceding
// kernel of the stream has completed and flushed), so the next kernel is set up while its predecessor drains. Every kernel
// waits unconditionally before its first memory access, hence completion of kernel k still implies completion of all
// kernels before it. 64 Ki x 7-DoF batch: 1.19 -> 1.12 ms, 16 Ki: 0.68 -> 0.60 ms; large batches unchanged (two chunks in
// flight already hide the gaps). RK_PDL = 2 (griddepcontrol.launch_dependents at kernel entry as well) was measured
// slower: 1.30 ms — the early successor blocks take the slots the other stream's kernel would have used.
#ifndef RK_PDL
#define RK_PDL 1
#endif
#if RK_PDL == 2
#define RK_PDL_ENTER() do { asm volatile("griddepcontrol.launch_dependents;"); asm volatile("griddepcontrol.wait;" ::: "memory"); } while (0)
#elif RK_PDL == 1
#define RK_PDL_ENTER() asm volatile("griddepcontrol.wait;" ::: "memory")
#else
#define RK_PDL_ENTER() ((void)0)
#endif
#ifndef RK_MIN_BLOCKS
#define RK_MIN_BLOCKS 4
#endif
// resident blocks per SM of the list-driven check kernels (A/B knobs; 128 threads per block)
#ifndef RK_QC_MINB
#define RK_QC_MINB(TYPE) ((TYPE) == 0 ? 6 : 8)
#endif
#ifndef RK_CK_MINB
#define RK_CK_MINB 8
#endif
#ifndef RK_VC_MINB
#define RK_VC_MINB 6
#endif
#ifndef RK_S0_MINB
#define RK_S0_MINB 6
#endif

namespace rk {

// ------------------------------------------------------------------------------------------------ candidate sub-lists in scratch
// Step-1 candidates of a third-order position item live in ten sub-lists, one per (family, direction slot):
//   sub 0..5 = quartic type (NONE, ACC0, ACC1) * 2 + slot, up to 4 candidates each; sub 6,7 = ACC0_ACC1, up to 2; sub 8,9 = VEL, one
//   slot per variant (4), of which the first valid one counts. Byte k of the sub-list's mask word: slot k holds a valid candidate.
// Direction slot 0 is UP for general items and sign(pd) for rest-target items (vf ~ 0 && af ~ 0). A slot holds t[7] and the duration.
__device__ __forceinline__ int sub_slot_base(int sub) { return sub < 6 ? sub * 4 : (sub < 8 ? 24 + (sub - 6) * 2 : 28 + (sub - 8) * 4); }
__device__ __forceinline__ int sub_limits(int sub) { return sub < 2 ? (int)L_NONE : (sub < 4 ? (int)L_ACC0 : (sub < 6 ? (int)L_ACC1 : (sub < 8 ? (int)L_ACC0_ACC1 : (int)L_VEL))); }

__device__ __forceinline__ void write_block(const Params& P, int64_t wi, int64_t wstride, const BlockRec& blk, bool found, uint32_t& meta) {
    double* w = P.ws + wi;
    w[(int64_t)W_TMIN * wstride] = found ? blk.t_min : 0.0;
    w[(int64_t)W_ALEFT * wstride] = blk.a_left; w[(int64_t)W_ARIGHT * wstride] = blk.a_right;
    w[(int64_t)W_BLEFT * wstride] = blk.b_left; w[(int64_t)W_BRIGHT * wstride] = blk.b_right;
    if (found) {
        meta |= M_FOUND;
        ws_put_cand(P, wi, wstride, 0, blk.p_min);
        if (blk.has_a) { meta |= M_HAS_A; ws_put_cand(P, wi, wstride, 1, blk.pa); }
        if (blk.has_b) { meta |= M_HAS_B; ws_put_cand(P, wi, wstride, 2, blk.pb); }
    }
}

// Append the items of this warp that are still unsolved to the next list (one atomic per warp).
__device__ __forceinline__ void append_unsolved(uint32_t* list, uint32_t* counter, bool unsolved, uint32_t item) {
    const unsigned mask = __ballot_sync(0xffffffffu, unsolved);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (unsolved) list[base + __popc(mask & ((1u << lane) - 1u))] = item;
}

// Lazy kernels (pass 2 of rk_kernels.cu, namespace rk_lazy): a division whose operands leave the fast path's domain raises the
// thread's flag instead of branching to the exact division (rk_math.cuh); the kernel clears the flag before a work item,
// looks at it afterwards and, if it is up, recomputes the item through the `*_exact` entry point of namespace rk — the same
// unit function compiled in pass 1, out of line. The flag is never up on ordinary data (tiny / huge intermediate values), so
// the test hook P.force_redo sends every fourth item through the exact path to keep that path under the parity tests.
// Launched from rk_lazy: k1_closed_cands (-10.6 % kernel time), k2_stage<0> (-4 %), k2_stage<3> (-12 %): +1.2 % on the whole step.
// The quartic kernels, the check kernels and the VEL check were built the same way and measured slower (+3 .. +10 %): they hold
// few `Den` divisions, and the out-of-line exact copy in the kernel costs their tight register budgets more than the branches did.
// With no redo behind it at all (a timing experiment) the lazy form of EVERY kernel is worth +5.1 %.
#undef RK_LAZY_BEGIN
#if RK_LAZY_DDIV
#define RK_LAZY_BEGIN() (rk_flag() = 0)
__device__ __forceinline__ bool redo_wanted(const Params& P, uint32_t key) {
    return (rk_flag() != 0) | ((P.force_redo != 0) & (((key * 2654435761u) >> 30) == 0u));
}
#define RK_EXACT_PARAMS(P) (*static_cast<const rk_exact::Params*>((P).self))
#else
#define RK_LAZY_BEGIN() ((void)0)
#endif

// The five bits of a pending third-order item that k1_block decides on (instead of re-reading the item's record there)
__device__ __forceinline__ uint32_t block_bits(double ps, double vs, double as, const DofIn& x) {
    const double pd = x.pf - ps;
    const bool rest_target = fabs(x.vf) < EPS && fabs(x.af) < EPS;
    const bool trivial = rest_target && fabs(vs) < EPS && fabs(as) < EPS && fabs(pd) < EPS;
    return (rest_target ? M_REST_TARGET : 0u) | (trivial ? M_TRIVIAL : 0u) | ((pd >= 0.0) ? M_PD_NONNEG : 0u)
           | ((x.v_max > 0.0) ? M_VMAX_POS : 0u) | ((x.v_min > 0.0) ? M_VMIN_POS : 0u);
}

// ------------------------------------------------------------------------------------------------ k1_setup
#ifndef RK_SU_MINB
#define RK_SU_MINB 4
#endif
__device__ __forceinline__ void setup_item(const Params& P, int64_t wi, int64_t wstride) {
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

        const Bound b{ps, vs, as, x.pf, x.vf, x.af};
        if (kind == K_POS3 && !(x.j_max == 0.0 || x.a_max == 0.0 || x.a_min == 0.0)) {
            meta |= M_S1_PENDING | block_bits(ps, vs, as, x);  // the three family kernels and k1_block take over
        } else {
            switch (kind) {
                case K_POS3: found = step1_pos3_zero_limits(b, x.v_max, x.v_min, x.a_max, x.a_min, x.j_max, bdur, blk); break;
                case K_POS2: found = step1_pos2(b, x.v_max, x.v_min, x.a_max, x.a_min, bdur, blk); break;
                case K_POS1: found = step1_pos1(b, x.v_max, x.v_min, bdur, blk); break;
                case K_VEL3: found = step1_vel3(b, x.a_max, x.a_min, x.j_max, bdur, blk); break;
                case K_VEL2: found = step1_vel2(b, x.a_max, x.a_min, bdur, blk); break;
                default: found = false; break;
            }
        }
    }

    double* w = P.ws + wi;
    w[(int64_t)W_P0 * wstride] = ps; w[(int64_t)W_V0 * wstride] = vs; w[(int64_t)W_A0 * wstride] = as;
    w[(int64_t)W_BRAKE_DUR * wstride] = bdur;
    store_rec(P, wi, ps, vs, as, x, bdur);
    if (!(meta & M_S1_PENDING)) {
        write_block(P, wi, wstride, blk, found, meta);
        P.imd[(int64_t)d * P.n + ig] = found ? blk.t_min : 0.0;
    }
    P.wmeta[(int64_t)WM_META * wstride + wi] = meta;

    // the parts of the final Profile that do not depend on the later stages
    if (P.compact) {  // compact record: the state before the brake pre-trajectory, the brake itself, the target
        out.put(C_P0, x.p0); out.put(C_V0, x.v0); out.put(C_A0, x.a0);
        out.put(C_BRAKE_T, bt0); out.put(C_BRAKE_T + 1, bt1);
        out.put(C_BRAKE_J, kind == K_POS2 ? ba0 : bj0); out.put(C_BRAKE_J + 1, bj1);
        out.put(C_PF, pf_store); out.put(C_VF, vf_store); out.put(C_AF, af_store);
        return;
    }
    out.put(S_BRAKE_DUR, bdur);
    out.put(S_BRAKE_T, bt0); out.put(S_BRAKE_T + 1, bt1);
    out.put(S_BRAKE_J, bj0); out.put(S_BRAKE_J + 1, bj1);
    out.put(S_BRAKE_A, ba0); out.put(S_BRAKE_A + 1, ba1);
    out.put(S_BRAKE_V, bv0); out.put(S_BRAKE_V + 1, bv1);
    out.put(S_BRAKE_P, bp0); out.put(S_BRAKE_P + 1, bp1);
    out.put(S_PF, pf_store); out.put(S_VF, vf_store); out.put(S_AF, af_store);
}

__device__ __forceinline__ void deposit_params(const Params& P) {
    if (P.self && blockIdx.x == 0 && threadIdx.x == 0) *static_cast<Params*>(const_cast<void*>(P.self)) = P;
}

__global__ void __launch_bounds__(128, RK_SU_MINB) k1_setup(const Params P) {
    RK_PDL_ENTER();
    deposit_params(P);
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= wstride) return;
    setup_item(P, wi, wstride);
}

// k1_setup for the case almost every item of a position-interface batch is: an enabled third-order DoF with non-zero limits
// that passes validation (or is not validated: ignoring handler). Only that path is compiled in — validation, the third-order
// brake pre-trajectory, the packed record — so the kernel runs at 8 blocks per SM instead of k1_setup's 4 (128 registers for
// the solver variants it carries). Every other item (disabled, invalid under the throwing handler, zero or infinite limits)
// is marked "not pending", so that the family kernels pass over it, and queued; k1_rescue runs the general setup_item on
// the queue before k_sync needs its block.
__global__ void __launch_bounds__(128, 8) k1_setup_lean(const Params P) {
    RK_PDL_ENTER();
    deposit_params(P);
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool other = false;
    if (wi < wstride) {
        const int d = (int)(wi / P.cn);
        const int64_t il = wi - (int64_t)d * P.cn;
        const int64_t ig = P.off + il;
        const DofIn x = load_dof(P, d, ig);
        const int ci = P.ci[d];
        const int vcode = validate_dof(x, ci, false, true);  // Ruckig::calculate validates with (false, true), ruckig.rs:215
        const int kind = kind_of(ci, x.a_max, x.j_max);
        other = !(kind == K_POS3 && x.enabled && (vcode == 0 || !P.throw_mode) && !(x.j_max == 0.0 || x.a_max == 0.0 || x.a_min == 0.0));
        if (other) {
            P.wmeta[(int64_t)WM_META * wstride + wi] = 0u;  // not pending: quartic / closed-form / block kernels skip the item
        } else {
            // brake pre-trajectory (calculator_target.rs:330-387, brake.rs:195-220)
            Brake br;
            br.t[0] = 0.0; br.t[1] = 0.0; br.j[0] = 0.0; br.j[1] = 0.0; br.a0 = 0.0;
            position_brake(br, x.v0, x.a0, x.v_max, x.v_min, x.a_max, x.a_min, x.j_max);
            const double bt0 = br.t[0], bt1 = br.t[1], bj0 = br.j[0], bj1 = br.j[1];
            double bdur = 0.0, ba0 = 0.0, ba1 = 0.0, bv0 = 0.0, bv1 = 0.0, bp0 = 0.0, bp1 = 0.0;
            double ps = x.p0, vs = x.v0, as = x.a0;
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
            double* w = P.ws + wi;
            w[(int64_t)W_P0 * wstride] = ps; w[(int64_t)W_V0 * wstride] = vs; w[(int64_t)W_A0 * wstride] = as;
            w[(int64_t)W_BRAKE_DUR * wstride] = bdur;
            store_rec(P, wi, ps, vs, as, x, bdur);
            P.wmeta[(int64_t)WM_META * wstride + wi] =
                (uint32_t)K_POS3 | ((uint32_t)vcode << M_VCODE_SHIFT) | M_ENABLED | M_S1_PENDING | block_bits(ps, vs, as, x);
            ProfileSink out{P.prof, (int64_t)P.dof * P.n, (int64_t)d * P.n + ig};
            if (P.compact) {
                out.put(C_P0, x.p0); out.put(C_V0, x.v0); out.put(C_A0, x.a0);
                out.put(C_BRAKE_T, bt0); out.put(C_BRAKE_T + 1, bt1);
                out.put(C_BRAKE_J, bj0); out.put(C_BRAKE_J + 1, bj1);
                out.put(C_PF, x.pf); out.put(C_VF, x.vf); out.put(C_AF, x.af);
            } else {
                out.put(S_BRAKE_DUR, bdur);
                out.put(S_BRAKE_T, bt0); out.put(S_BRAKE_T + 1, bt1);
                out.put(S_BRAKE_J, bj0); out.put(S_BRAKE_J + 1, bj1);
                out.put(S_BRAKE_A, ba0); out.put(S_BRAKE_A + 1, ba1);
                out.put(S_BRAKE_V, bv0); out.put(S_BRAKE_V + 1, bv1);
                out.put(S_BRAKE_P, bp0); out.put(S_BRAKE_P + 1, bp1);
                out.put(S_PF, x.pf); out.put(S_VF, x.vf); out.put(S_AF, x.af);
            }
        }
    }
    append_unsolved(P.list[1], P.counters + 25, other, (uint32_t)wi);
}

// warp-aggregated append of a variable number of entries per lane; returns the lane's first index
__device__ __forceinline__ uint32_t warp_reserve(uint32_t* counter, int mine) {
    const int lane = threadIdx.x & 31;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base = 0;
    if (total > 0 && lane == 31) base = atomicAdd(counter, (uint32_t)total);
    return __shfl_sync(0xffffffffu, base, 31) + (uint32_t)(incl - mine);
}

__device__ __forceinline__ Lim family_limits(const DofIn& x, const Bound& b, int ds, bool& rest_target, bool& trivial) {
    const double pd = b.pf - b.p0;
    rest_target = fabs(b.vf) < EPS && fabs(b.af) < EPS;
    trivial = rest_target && fabs(b.v0) < EPS && fabs(b.a0) < EPS && fabs(pd) < EPS;
    const bool up = rest_target ? ((ds == 0) == (pd >= 0.0)) : (ds == 0);
    return lim_dir(up, x.v_max, x.v_min, x.a_max, x.a_min, x.j_max);
}

// ------------------------------------------------------------------------------------------------ k1_closed_cands / k1_cand_check
// The closed-form families ACC0_ACC1 (position_third_step1.rs:241-327) and VEL (:97-240) in two phases, like the quartics
// below. Phase A, one thread per (direction slot, DoF, trajectory): the up to 2 + 4 candidate t[7] vectors, screened with the
// tests of the profile check that need no integration, parked in their sub-list slots and queued. Phase B, one thread per
// queued candidate: the full profile check, a valid candidate sets its byte of the sub-list's mask word. The order in which
// the reference would have met the candidates is restored by k1_block (for VEL only the first valid variant counts).
#ifndef RK_CC_MINB
#define RK_CC_MINB 8
#endif
#ifndef RK_CC_SPLIT
#define RK_CC_SPLIT 0  // 1: ACC0_ACC1 and VEL in two launches (PART 0 / 1) instead of one (PART 2)
#endif
// the work of one (direction slot, item): candidates parked in their slots, bit v of parked3 / parked4 = variant v is to be checked
template <int PART>
__device__ __forceinline__ void closed_cands_unit(const Params& P, int64_t wstride, int64_t wi, int ds, uint32_t& parked3, uint32_t& parked4) {
    parked3 = 0; parked4 = 0;
    const uint32_t meta = P.wmeta[(int64_t)WM_META * wstride + wi];
    Bound b;
    double bdur_;
    const DofIn x = load_rec(P, wi, b, bdur_);  // before the test of meta: two independent loads instead of a dependent pair
    if (meta & M_S1_PENDING) {
        bool rest_target, trivial;
        const Lim L = family_limits(x, b, ds, rest_target, trivial);
        if (PART != 1) P.wqmask[(int64_t)(6 + ds) * wstride + wi] = 0u;
        if (PART != 0) P.wqmask[(int64_t)(8 + ds) * wstride + wi] = 0u;
        if (!trivial) {  // :1015-1025 — nothing moves: only the quartics of the first direction are tried
            Step1Pos3<CandSink> s;
            s.init(b, x.j_max);
            s.first_only = false;
            if (PART != 1) {
                s.sink.parked = 0;
                s.sink.base = vl_slot(P, wi, sub_slot_base(6 + ds));
                s.family_acc0_acc1(L);
                parked3 = s.sink.parked;
            }
            if (PART != 0) {
                s.sink.parked = 0;
                s.sink.base = vl_slot(P, wi, sub_slot_base(8 + ds));
                s.family_vel(L);
                parked4 = s.sink.parked;
            }
        }
    }
}
template <int PART>
__device__ __noinline__ void closed_cands_exact(const Params& P, int64_t wstride, int64_t wi, int ds, uint32_t& parked3, uint32_t& parked4) {
    closed_cands_unit<PART>(P, wstride, wi, ds, parked3, parked4);
}

template <int PART>
__global__ void __launch_bounds__(128, RK_CC_MINB) k1_closed_cands(const Params P) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t parked3 = 0, parked4 = 0;
    int ds = 0;
    int64_t wi = 0;
    if (tid < 2 * wstride) {
        ds = (int)(tid & 1);  // the two directions of an item sit in adjacent lanes: its 96-byte record is fetched once
        wi = tid >> 1;
        RK_LAZY_BEGIN();
        closed_cands_unit<PART>(P, wstride, wi, ds, parked3, parked4);
#if RK_LAZY_DDIV
        if (redo_wanted(P, (uint32_t)tid)) rk_exact::closed_cands_exact<PART>(RK_EXACT_PARAMS(P), wstride, wi, ds, parked3, parked4);
#endif
    }
    const int mine = __popc(parked3) + __popc(parked4);
    uint32_t e = warp_reserve(P.counters + 23, mine);
#pragma unroll
    for (int v = 0; v < 2; ++v)
        if ((parked3 >> v) & 1u) P.rq_item[e++] = make_uint2((uint32_t)wi, (uint32_t)ds | (uint32_t)((6 + ds) << 1) | ((uint32_t)v << 8));
#pragma unroll
    for (int v = 0; v < 4; ++v)
        if ((parked4 >> v) & 1u) P.rq_item[e++] = make_uint2((uint32_t)wi, (uint32_t)ds | (uint32_t)((8 + ds) << 1) | ((uint32_t)v << 8));
}

// one queued candidate: the full profile check; byte v of the sub-list's mask word = valid
__device__ __forceinline__ void cand_check_unit(const Params& P, int64_t wstride, const uint2 it) {
    const int64_t wi = it.x;
    const int ds = it.y & 1, sub = (it.y >> 1) & 0x7F, v = it.y >> 8;
    Bound b;
    double bdur_;
    const DofIn x = load_rec(P, wi, b, bdur_);
    bool rest_target, trivial;
    const Lim L = family_limits(x, b, ds, rest_target, trivial);
    const double2* slot = reinterpret_cast<const double2*>(vl_slot(P, wi, sub_slot_base(sub) + v));
    const double2 s0 = slot[0], s1 = slot[1], s2 = slot[2], s3 = slot[3];
    const double t[7] = {s0.x, s0.y, s1.x, s1.y, s2.x, s2.y, s3.x};
    if (check_pos3_inline(t, UDDU, sub_limits(sub), L.j_max, L, b)) reinterpret_cast<uint8_t*>(P.wqmask + (int64_t)sub * wstride + wi)[v] = 1;
}

__global__ void __launch_bounds__(128, RK_CK_MINB) k1_cand_check(const Params P) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const uint32_t count = P.counters[23];
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        const uint2 it = P.rq_item[e];
        cand_check_unit(P, wstride, it);
    }
}

// ------------------------------------------------------------------------------------------------ k1_quartic_roots / _check
// The three quartic families in two phases. Phase A (one thread per (direction, DoF, trajectory)) builds the quartic,
// pre-screens it, solves it in closed form and queues every root inside [t_min, t_max]; phase B (one thread per queued
// root) does the Newton polish, builds t[7] and validates the profile. A (DoF, direction) has 0-4 in-range roots and most
// candidates die in the check, so phase B runs the expensive part with full warps instead of the few lanes that happened
// to have a root in the same loop iteration. Root k of a sub-list writes slot k, which keeps the ascending order.
// one (direction slot, item): the quartic of family TYPE, its in-range roots in ascending order
template <int TYPE>
__device__ __forceinline__ void quartic_roots_unit(const Params& P, int64_t wstride, int64_t wi, int ds, double (&roots)[4], int& n_roots) {
    n_roots = 0;
    const uint32_t meta = P.wmeta[(int64_t)WM_META * wstride + wi];
    Bound b;
    double bdur_;
    const DofIn x = load_rec(P, wi, b, bdur_);  // before the test of meta: two independent loads instead of a dependent pair
    if (meta & M_S1_PENDING) {
        bool rest_target, trivial;
        const Lim L = family_limits(x, b, ds, rest_target, trivial);
        P.wqmask[(int64_t)(TYPE * 2 + ds) * wstride + wi] = 0u;
        if (!(trivial && ds == 1)) {  // :1015-1025 — nothing moves: only the first direction is tried
            Step1Pos3<QuarticSink> s;
            s.init(b, x.j_max);
            s.sink.n_roots = 0;
            s.first_only = false;
            if (TYPE == 0) s.template quartic_none<QUARTIC_ROOTS>(L, 0.0);
            else if (TYPE == 1) s.template quartic_acc0<QUARTIC_ROOTS>(L, 0.0);
            else s.template quartic_acc1<QUARTIC_ROOTS>(L, 0.0);
            n_roots = s.sink.n_roots;
#pragma unroll
            for (int k = 0; k < 4; ++k) roots[k] = s.sink.roots[k];
        }
    }
}

#ifndef RK_QR_MINB
#define RK_QR_MINB 12  // 40 registers, 72 bytes of spills: the kernels are bound by issue slots under fp64 dependency stalls, and 48 warps per SM beat 32 (+0.8 % on the step)
#endif
template <int TYPE>
__global__ void __launch_bounds__(128, RK_QR_MINB) k1_quartic_roots(const Params P) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double roots[4];
    int n_roots = 0;
    int ds = 0;
    int64_t wi = 0;
    if (tid < 2 * wstride) {
        ds = (int)(tid & 1);  // the two directions of an item sit in adjacent lanes: its 96-byte record is fetched once
        wi = tid >> 1;
        quartic_roots_unit<TYPE>(P, wstride, wi, ds, roots, n_roots);
    }
    // warp-aggregated append of a variable number of entries per lane
    const int lane = threadIdx.x & 31;
    int incl = n_roots;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    uint32_t base = 0;
    if (lane == 31) base = atomicAdd(P.counters + 20 + TYPE, (uint32_t)total);
    base = __shfl_sync(0xffffffffu, base, 31) + (uint32_t)(incl - n_roots);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < n_roots) {
            P.rq_item[base + k] = make_uint2((uint32_t)wi, (uint32_t)ds | ((uint32_t)k << 1));
            P.rq_root[base + k] = roots[k];
        }
    }
}

// one queued root: Newton polish, t[7], profile check; a valid candidate goes into the slot the root owns
template <int TYPE>
__device__ __forceinline__ void quartic_check_unit(const Params& P, int64_t wstride, const uint2 it, double root) {
    const int64_t wi = it.x;
    const int ds = it.y & 1, rank = it.y >> 1;
    Bound b;
    double bdur_;
    const DofIn x = load_rec(P, wi, b, bdur_);
    bool rest_target, trivial;
    const Lim L = family_limits(x, b, ds, rest_target, trivial);
    const int sub = TYPE * 2 + ds;
    Step1Pos3<QuarticSink> s;
    s.init(b, x.j_max);
    s.first_only = false;
    s.sink.slot = vl_slot(P, wi, sub_slot_base(sub) + rank);
    if (TYPE == 0) s.template quartic_none<QUARTIC_CANDIDATE>(L, root);
    else if (TYPE == 1) s.template quartic_acc0<QUARTIC_CANDIDATE>(L, root);
    else s.template quartic_acc1<QUARTIC_CANDIDATE>(L, root);
    if (s.sink.count > 0) reinterpret_cast<uint8_t*>(P.wqmask + (int64_t)sub * wstride + wi)[rank] = 1;
}

template <int TYPE>
__global__ void __launch_bounds__(128, RK_QC_MINB(TYPE)) k1_quartic_check(const Params P) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const uint32_t count = P.counters[20 + TYPE];
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        const uint2 it = P.rq_item[e];
        const double root = P.rq_root[e];
        quartic_check_unit<TYPE>(P, wstride, it, root);
    }
}

// ------------------------------------------------------------------------------------------------ k1_block
// Candidate sub-lists -> block of feasible durations (block.rs:40-154). The candidates are never copied: the kernel works on
// (duration, sub-list slot) pairs held in registers, and the three profiles a block keeps (p_min, a.profile, b.profile) are
// recorded as references to their sub-list slots (bits 24..31 of the code word, slot + 1; 0 = the profile sits in the
// W_CAND slots, which the other Step-1 variants and the rescues still use). Items without any candidate go to k1_rescue.
struct RefList {
    double dur[5];
    int ref[5];  // sub-list slot | sub << 8 | direction << 16 | up << 24
    int n;
    __device__ __forceinline__ void push(double d, int r) {
#pragma unroll
        for (int q = 0; q < 5; ++q)
            if (q == n) { dur[q] = d; ref[q] = r; }
        ++n;
    }
    __device__ __forceinline__ double d(int i) const { return i == 0 ? dur[0] : (i == 1 ? dur[1] : (i == 2 ? dur[2] : (i == 3 ? dur[3] : dur[4]))); }
    __device__ __forceinline__ int r(int i) const { return i == 0 ? ref[0] : (i == 1 ? ref[1] : (i == 2 ? ref[2] : (i == 3 ? ref[3] : ref[4]))); }
    __device__ __forceinline__ int dir(int i) const { return (r(i) >> 16) & 0xFF; }
};

struct BlockRefs {
    double t_min, a_left, a_right, b_left, b_right;
    int has_a, has_b;
    int r_min, r_a, r_b;
};

// block.rs:220-239 on references
__device__ __forceinline__ void make_interval_ref(const RefList& v, int il, int ir, double bd, double& left, double& right, int& prof) {
    const double ld = v.d(il) + bd + 0.0;
    const double rd = v.d(ir) + bd + 0.0;
    if (ld < rd) { left = ld; right = rd; prof = v.r(ir); }
    else { left = rd; right = ld; prof = v.r(il); }
}

// block.rs:40-154 (same decisions as calculate_block in rk_step1.cuh)
__device__ __forceinline__ bool calculate_block_refs(BlockRefs& blk, RefList& v, double bd) {
    blk.has_a = 0; blk.has_b = 0;
    int n = v.n;
    if (n == 1) {
        blk.r_min = v.r(0);
        blk.t_min = v.d(0) + bd + 0.0;
        return true;
    } else if (n == 2) {
        if (fabs(v.d(0) - v.d(1)) < 8.0 * EPS) {
            blk.r_min = v.r(0);
            blk.t_min = v.d(0) + bd + 0.0;
            return true;
        }
        const int idx_min = (v.d(0) < v.d(1)) ? 0 : 1;
        const int idx_else = 1 - idx_min;
        blk.r_min = v.r(idx_min);
        blk.t_min = v.d(idx_min) + bd + 0.0;
        make_interval_ref(v, idx_min, idx_else, bd, blk.a_left, blk.a_right, blk.r_a);
        blk.has_a = 1;
        return true;
    } else if (n == 4) {
        if (fabs(v.d(0) - v.d(1)) < 32.0 * EPS && v.dir(0) != v.dir(1)) {
            v.dur[1] = v.dur[2]; v.ref[1] = v.ref[2];
            v.dur[2] = v.dur[3]; v.ref[2] = v.ref[3];
        } else if ((fabs(v.d(2) - v.d(3)) < 256.0 * EPS && v.dir(2) != v.dir(3)) || (fabs(v.d(0) - v.d(3)) < 256.0 * EPS && v.dir(0) != v.dir(3))) {
            // entry 3 is dropped
        } else {
            return false;
        }
        n = 3;
    } else if (n % 2 == 0) {
        return false;
    }

    int idx_min = 0;
#pragma unroll
    for (int i = 1; i < 5; ++i)
        if (i < n && v.d(i) < v.d(idx_min)) idx_min = i;
    blk.r_min = v.r(idx_min);
    blk.t_min = v.d(idx_min) + bd + 0.0;

    if (n == 3) {
        make_interval_ref(v, (idx_min + 1) % 3, (idx_min + 2) % 3, bd, blk.a_left, blk.a_right, blk.r_a);
        blk.has_a = 1;
        return true;
    } else if (n == 5) {
        const int e1 = (idx_min + 1) % 5, e2 = (idx_min + 2) % 5, e3 = (idx_min + 3) % 5, e4 = (idx_min + 4) % 5;
        if (v.dir(e1) == v.dir(e2)) {
            make_interval_ref(v, e1, e2, bd, blk.a_left, blk.a_right, blk.r_a);
            make_interval_ref(v, e3, e4, bd, blk.b_left, blk.b_right, blk.r_b);
        } else {
            make_interval_ref(v, e1, e4, bd, blk.a_left, blk.a_right, blk.r_a);
            make_interval_ref(v, e2, e3, bd, blk.b_left, blk.b_right, blk.r_b);
        }
        blk.has_a = 1; blk.has_b = 1;
        return true;
    }
    return false;
}

__device__ __forceinline__ uint32_t ref_codes(int ref) {
    const int slot = ref & 0xFF, sub = (ref >> 8) & 0xFF, dir = (ref >> 16) & 0xFF, up = (ref >> 24) & 1;
    return pack_codes(sub_limits(sub), UDDU, dir | (up << 7), slot + 1);
}

__global__ void __launch_bounds__(128, 4) k1_block(const Params P) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool rescue = false;
    if (wi < wstride) {
        uint32_t meta = P.wmeta[(int64_t)WM_META * wstride + wi];
        const double bdur = P.ws[(int64_t)W_BRAKE_DUR * wstride + wi];  // with the five bits of meta: all this kernel needs of the record
        if (meta & M_S1_PENDING) {
            const int d = (int)(wi / P.cn);
            const int64_t ig = P.off + (wi - (int64_t)d * P.cn);
            const bool rest_target = (meta & M_REST_TARGET) != 0, trivial = (meta & M_TRIVIAL) != 0;
            // direction slot 0 is UP for general items and sign(pd) for rest-target items; the Direction label of a profile
            // follows the sign of the direction-resolved v_max (profile.rs), the control value follows `up`
            const bool vmax_pos = (meta & M_VMAX_POS) != 0, vmin_pos = (meta & M_VMIN_POS) != 0;
            const bool up0 = rest_target ? ((meta & M_PD_NONNEG) != 0) : true;
            const int tag_slot0 = (((up0 ? vmax_pos : vmin_pos) ? DIR_UP : DIR_DOWN) << 16) | ((up0 ? 1 : 0) << 24);
            const int tag_slot1 = (((up0 ? vmin_pos : vmax_pos) ? DIR_UP : DIR_DOWN) << 16) | ((up0 ? 0 : 1) << 24);

            // byte k of a sub-list's mask word says whether slot k holds a validated candidate
            uint32_t qmask[10];
            int cnt[10];
#pragma unroll
            for (int sub = 0; sub < 10; ++sub) qmask[sub] = P.wqmask[(int64_t)sub * wstride + wi];
#pragma unroll
            for (int sub = 8; sub < 10; ++sub) {  // VEL keeps the first valid variant only (:97-240 returns after the first)
                const uint32_t m = qmask[sub];
                qmask[sub] = (m & 0xFFu) ? 0x1u : ((m & 0xFF00u) ? 0x100u : ((m & 0xFF0000u) ? 0x10000u : (m & 0xFF000000u)));
            }
#pragma unroll
            for (int sub = 0; sub < 10; ++sub) cnt[sub] = qmask[sub] ? 1 : 0;

            // Two passes: first WHICH candidates the reference's enumeration keeps (masks only, no memory), then their stored
            // durations with independent loads issued back to back (the kernel is bound by the latency of these scattered
            // 8-byte reads: ncu had half of its samples on the one dependent load per loop iteration).
            int sel[5];
            int n_sel = 0;
#pragma unroll
            for (int q = 0; q < 5; ++q) sel[q] = 0;
            auto take = [&](int sub, int k) {
                const int slot = sub_slot_base(sub) + k;
                const int r = slot | (sub << 8) | ((sub & 1) ? tag_slot1 : tag_slot0);
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    if (q == n_sel) sel[q] = r;
                ++n_sel;
            };
            if (rest_target) {
                // first hit in the reference's order (:1026-1089): VEL(d), quartics(d) = NONE, ACC0, ACC1, ACC0_ACC1(d), then the same for -d
                int pick = -1;
#pragma unroll
                for (int ds = 0; ds < 2; ++ds) {
                    if (trivial && ds == 1) break;
                    const int order[5] = {8 + ds, 0 + ds, 2 + ds, 4 + ds, 6 + ds};
#pragma unroll
                    for (int q = 0; q < 5; ++q)
                        if (pick < 0 && cnt[order[q]] > 0) pick = order[q];
                }
                if (pick >= 0) {
                    uint32_t m = 0;
#pragma unroll
                    for (int sub = 0; sub < 10; ++sub)
                        if (sub == pick) m = qmask[sub];
                    take(pick, (m & 0xFFu) ? 0 : ((m & 0xFF00u) ? 1 : ((m & 0xFF0000u) ? 2 : 3)));
                }
            } else {
                // enumeration order of :1091-1140: quartics(UP), quartics(DOWN), ACC0_ACC1(UP), (DOWN), VEL(UP), (DOWN); five are kept (Q11)
                const int order[10] = {0, 2, 4, 1, 3, 5, 6, 7, 8, 9};
#pragma unroll
                for (int q = 0; q < 10; ++q) {
                    const int sub = order[q];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (n_sel < 5 && ((qmask[sub] >> (8 * k)) & 1u)) take(sub, k);
                }
            }
            RefList v;
            v.n = n_sel;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                v.ref[q] = sel[q];
                v.dur[q] = (q < n_sel) ? vl_slot(P, wi, sel[q] & 0xFF)[7] : 0.0;  // [7] = t_sum6(t)
            }

            if (v.n == 0) {
                rescue = true;  // :1142-1255, handled by k1_rescue
            } else {
                BlockRefs blk;
                blk.a_left = 0.0; blk.a_right = 0.0; blk.b_left = 0.0; blk.b_right = 0.0; blk.t_min = 0.0;
                blk.r_min = 0; blk.r_a = 0; blk.r_b = 0;
                const bool found = calculate_block_refs(blk, v, bdur);
                meta &= ~M_S1_PENDING;
                double* wo = P.ws + wi;
                wo[(int64_t)W_TMIN * wstride] = found ? blk.t_min : 0.0;
                wo[(int64_t)W_ALEFT * wstride] = blk.a_left; wo[(int64_t)W_ARIGHT * wstride] = blk.a_right;
                wo[(int64_t)W_BLEFT * wstride] = blk.b_left; wo[(int64_t)W_BRIGHT * wstride] = blk.b_right;
                if (found) {
                    meta |= M_FOUND;
                    P.wmeta[(int64_t)(WM_CAND + 0) * wstride + wi] = ref_codes(blk.r_min);
                    if (blk.has_a) { meta |= M_HAS_A; P.wmeta[(int64_t)(WM_CAND + 1) * wstride + wi] = ref_codes(blk.r_a); }
                    if (blk.has_b) { meta |= M_HAS_B; P.wmeta[(int64_t)(WM_CAND + 2) * wstride + wi] = ref_codes(blk.r_b); }
                }
                P.wmeta[(int64_t)WM_META * wstride + wi] = meta;
                P.imd[(int64_t)d * P.n + ig] = found ? blk.t_min : 0.0;
            }
        }
    }
    append_unsolved(P.list[0], P.counters + 24, rescue, (uint32_t)wi);
}

__device__ __noinline__ void setup_item_rare(const Params& P, int64_t wi, int64_t wstride) { setup_item(P, wi, wstride); }

// Numerical rescues (:1142-1255) for the rare items none of the regular families produced a candidate for; the first rescue
// that yields one wins.
__global__ void __launch_bounds__(128) k1_rescue(const Params P) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    // the items k1_setup_lean passed over (the counter stays zero when the general k1_setup ran): complete set-up, including
    // the Step 1 of the cheap variants
    const uint32_t n_other = P.counters[25];
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_other; e += gridDim.x * blockDim.x) setup_item_rare(P, P.list[1][e], wstride);
    const uint32_t n_items = P.counters[24];
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_items; e += gridDim.x * blockDim.x) {
        const int64_t wi = P.list[0][e];
        uint32_t meta = P.wmeta[(int64_t)WM_META * wstride + wi];
        const int d = (int)(wi / P.cn);
        const int64_t ig = P.off + (wi - (int64_t)d * P.cn);
        const double* w = P.ws + wi;
        const double bdur = w[(int64_t)W_BRAKE_DUR * wstride];
        const DofIn x = load_dof(P, d, ig);
        const Bound b{w[(int64_t)W_P0 * wstride], w[(int64_t)W_V0 * wstride], w[(int64_t)W_A0 * wstride], x.pf, x.vf, x.af};
        Step1Pos3<LocalSink> s;
        s.init(b, x.j_max);
        s.first_only = false;
#pragma unroll 1
        for (int k = 0; k < 8 && s.sink.count == 0; ++k) {
            const Lim L = lim_dir((k & 1) == 0, x.v_max, x.v_min, x.a_max, x.a_min, x.j_max);
            if (k < 2) s.rescue_none(L);
            else if (k < 4) s.rescue_acc0(L);
            else if (k < 6) s.rescue_vel(L);
            else s.rescue_acc1_vel(L);
        }
        BlockRec blk;
        blk.a_left = 0.0; blk.a_right = 0.0; blk.b_left = 0.0; blk.b_right = 0.0; blk.t_min = 0.0;
        const bool found = calculate_block(blk, s.sink.vp, s.sink.count, bdur);
        meta &= ~M_S1_PENDING;
        write_block(P, wi, wstride, blk, found, meta);
        P.wmeta[(int64_t)WM_META * wstride + wi] = meta;
        P.imd[(int64_t)d * P.n + ig] = found ? blk.t_min : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------ k_sync
// block.rs:157-172
__device__ __forceinline__ bool is_blocked(double t, double t_min, bool has_a, double a_left, double a_right, bool has_b, double b_left, double b_right) {
    if (t < t_min) return true;
    if (has_a && t > a_left && t < a_right) return true;
    if (has_b && t > b_left && t < b_right) return true;
    return false;
}

__device__ __forceinline__ void sync_trajectory(const Params& P, int64_t il, uint32_t* pm, uint8_t* psrc) {
    const int64_t ig = P.off + il;
    const int D = P.dof;
    const int64_t wstride = (int64_t)D * P.cn;
    const double* W = P.ws;
#define WSG(slot, d) W[(int64_t)(slot) * wstride + (int64_t)(d) * P.cn + il]
    // Everything the decisions below look at, fetched up front with independent loads (the kernel is one thread per
    // trajectory and would otherwise walk a chain of dependent global loads); local arrays, L1-resident.
    double pw[6][RK_MAX_DOF];  // W_TMIN, W_ALEFT, W_ARIGHT, W_BLEFT, W_BRIGHT, W_BRAKE_DUR
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
        pm[d] = P.wmeta[(int64_t)d * P.cn + il];
        pw[0][d] = WSG(W_TMIN, d); pw[1][d] = WSG(W_ALEFT, d); pw[2][d] = WSG(W_ARIGHT, d);
        pw[3][d] = WSG(W_BLEFT, d); pw[4][d] = WSG(W_BRIGHT, d); pw[5][d] = WSG(W_BRAKE_DUR, d);
        psrc[d] = RK_SRC_NONE;
    }
#define WS(slot, d) ((slot) == W_TMIN ? pw[0][d] : (slot) == W_ALEFT ? pw[1][d] : (slot) == W_ARIGHT ? pw[2][d] : (slot) == W_BLEFT ? pw[3][d] : \
                     (slot) == W_BRIGHT ? pw[4][d] : (slot) == W_BRAKE_DUR ? pw[5][d] : WSG(slot, d))
#define META(d) pm[d]
#define SRC(d) psrc[d]
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

    const double md_raw = P.min_dur ? P.min_dur[ig + P.in_shift] : __longlong_as_double(0x7ff8000000000000LL);
    const bool has_md = md_raw == md_raw;
    const double md = has_md ? md_raw : 0.0;

    bool done = status != 0;
    if (!done && D == 1 && !has_md && !P.discrete) {  // :475-481
        duration = WS(W_TMIN, 0);
        SRC(0) = RK_SRC_STEP1_MIN;  // also for a disabled DoF: the reference copies blocks[0].p_min unconditionally (:478)
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
                for (int d = 0; d < D; ++d) SRC(d) = RK_SRC_STEP1_MIN;  // every DoF, disabled ones included (:537-539)
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
                    ws_get_cand(P, (int64_t)limiting * P.cn + il, wstride, SRC(limiting), load_dof(P, limiting, ig).j_max, lt, lctrl, lctrl2, lcodes);
                    const int l_lim = lcodes & 0xFF, l_cs = (lcodes >> 8) & 0xF, l_dir = (lcodes >> 16) & 0x7F;

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
                                    P.wmeta[(int64_t)3 * wstride + wd] |= (uint32_t)(kind + 1) << 12;  // kind the profile is expanded with
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
#undef WSG
#undef META
#undef SRC
}

// One thread per trajectory. After the decision every (DoF, trajectory) item is queued for the kernel that finishes it:
// third-order position items that need the Step-2 chain go to list[0] (stage 0 reads it); k_expand finishes all items.
#ifndef RK_SY_MINB
#define RK_SY_MINB 8
#endif
__global__ void __launch_bounds__(128, RK_SY_MINB) k_sync(const Params P) {
    RK_PDL_ENTER();
    const int64_t il = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = il < P.cn;
    uint32_t pm[RK_MAX_DOF];
    uint8_t psrc[RK_MAX_DOF];
    if (valid) sync_trajectory(P, il, pm, psrc);
    for (int d = 0; d < P.dof; ++d) {
        const uint32_t wi = (uint32_t)((int64_t)d * P.cn + il);
        const bool chain = valid && psrc[d] == RK_SRC_STEP2 && (pm[d] & M_KIND_MASK) == K_POS3;
        if (valid) P.wsrc[wi] = psrc[d];
        append_unsolved(P.list[0], P.counters + 0, chain, wi);
    }
}

// ------------------------------------------------------------------------------------------------ k_sync_lanes
// The synchronisation stage as warp-level reductions over the DoFs (the common configuration: at most 8 DoFs, continuous
// durations, no DoF with Phase synchronisation; everything else runs k_sync above). One lane per (trajectory, DoF), 8 lanes
// per trajectory, 4 trajectories per warp: every lane holds the block of ITS DoF in registers — t_min, the two blocked
// intervals, the brake duration — and the trajectory-level decisions of calculator_target.rs:150-275 / :454-550 / :714-747 are
// ballots and xor-shuffle min / max reductions inside the 8-lane group; nothing is sorted and nothing lives in local memory.
//
// Why no sort is needed (continuous durations). The reference sorts the 3 * DOF + 1 candidates (DOF minimum durations, the right
// ends of the blocked intervals, the minimum duration) stably by value and tests them in order FROM SORTED POSITION DOF - 1.
// Let T be the largest per-DoF minimum duration. A candidate below T is blocked by the DoF that owns T, so whether the scan
// starts before or after it changes nothing. Candidates equal to T all share one verdict; of those, the scan skips the first
// max(0, DOF - 1 - L) in index order (L = candidates below T) — that decides WHICH of several tied candidates is reported as the
// limiting one. Everything above T is visited in (value, index) order, and equal values again share a verdict. So: one max
// reduction for T, one count for L, one n-th-set-bit for the tie, then a min reduction per distinct value above T until one
// is not blocked. With Discrete durations the rounded candidates no longer dominate the raw minimum durations ulp for ulp, so
// that mode keeps the literal sort of k_sync.
__device__ __forceinline__ double group_max8(double v) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) { const double u = __shfl_xor_sync(0xffffffffu, v, o); v = (u > v) ? u : v; }
    return v;
}
__device__ __forceinline__ double group_min8(double v) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) { const double u = __shfl_xor_sync(0xffffffffu, v, o); v = (u < v) ? u : v; }
    return v;
}
__device__ __forceinline__ int group_sum8(int v) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256, 4) k_sync_lanes(const Params P) {
    RK_PDL_ENTER();
    __shared__ uint8_t src_of[8][32];  // the block's 32 trajectories x 8 DoFs, transposed for the DoF-major stores at the end
    const int lane = threadIdx.x & 31, d = lane & 7, gshift = lane & 24;  // gshift: first lane of this trajectory's group
    const int64_t il = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int D = P.dof;
    const bool tvalid = il < P.cn, mine = tvalid && d < D;
    const int64_t wstride = (int64_t)D * P.cn, wi = (int64_t)d * P.cn + il, ig = P.off + il;
    auto gbyte = [&](bool pred) { return (__ballot_sync(0xffffffffu, pred) >> gshift) & 0xFFu; };

    uint32_t m = 0;
    double t_min = 0.0, a_left = 0.0, a_right = 0.0, b_left = 0.0, b_right = 0.0, bdur = 0.0;
    if (mine) {
        const double* w = P.ws + wi;
        m = P.wmeta[wi];
        t_min = w[(int64_t)W_TMIN * wstride]; a_left = w[(int64_t)W_ALEFT * wstride]; a_right = w[(int64_t)W_ARIGHT * wstride];
        b_left = w[(int64_t)W_BLEFT * wstride]; b_right = w[(int64_t)W_BRIGHT * wstride]; bdur = w[(int64_t)W_BRAKE_DUR * wstride];
    }
    const int sync_d = (d < D) ? (int)P.sync[d] : (int)RK_SYNC_NONE;
    const bool none = sync_d == RK_SYNC_NONE;
    const bool enabled = mine && (m & M_ENABLED), has_a = (m & M_HAS_A) != 0, has_b = (m & M_HAS_B) != 0;
    const bool sync_on = mine && !none;  // takes part in candidates and blocking

    int status = RK_RESULT_WORKING, detail = 0, limiting = -1;
    double duration = 0.0;
    int src = RK_SRC_NONE;

    // NB: the four trajectories of a warp take different paths, so every ballot / shuffle below is executed unconditionally by
    // all 32 lanes and only the USE of its result is conditional.

    // Err(ValidationError): first failing DoF (input_parameter.rs:251-420)
    const uint32_t vc = (m >> M_VCODE_SHIFT) & 0xFFu;
    const uint32_t vbits = gbyte(mine && vc != 0);
    const int dfv = vbits ? __ffs(vbits) - 1 : 0;
    const uint32_t vc_first = __shfl_sync(0xffffffffu, vc, gshift + dfv);
    if (P.throw_mode && vbits) {
        status = RK_RESULT_ERROR_INVALID_INPUT;
        detail = (int)(vc_first << 8) | dfv;
    }
    // Step-1 failure: lowest DoF index decides (calculator_target.rs:454-471)
    const uint32_t fbits = gbyte(enabled && !(m & M_FOUND));
    const int dff = fbits ? __ffs(fbits) - 1 : 0;
    const int zl_first = __shfl_sync(0xffffffffu, (int)((m & M_ZERO_LIMITS) != 0), gshift + dff);
    if (status == 0 && fbits) {
        status = zl_first ? RK_RESULT_ERROR_ZERO_LIMITS : RK_RESULT_ERROR_EXECUTION_TIME_CALCULATION;
        detail = (1 << 8) | dff;
    }
    const double md_raw = (P.min_dur && tvalid) ? P.min_dur[ig + P.in_shift] : __longlong_as_double(0x7ff8000000000000LL);
    const bool has_md = md_raw == md_raw;
    const double md = has_md ? md_raw : 0.0;

    bool done = status != 0;
    const double t_min_first = __shfl_sync(0xffffffffu, t_min, gshift);
    if (!done && D == 1 && !has_md) {  // :475-481 (also for a disabled DoF: blocks[0].p_min is copied unconditionally)
        duration = t_min_first;
        if (d == 0) src = RK_SRC_STEP1_MIN;
        done = true;
    }

    // ---- synchronize (:150-275); every lane of the warp runs the reductions, `done` groups ignore the outcome
    const double c0 = sync_on ? t_min : (mine ? 0.0 : -RK_INF);  // lanes without a DoF never win a max and are never counted
    const double c1 = (sync_on && has_a) ? a_right : RK_INF;
    const double c2 = (sync_on && has_b) ? b_right : RK_INF;
    const double c3 = has_md ? md : RK_INF;  // candidate 3 * DOF, held by every lane, counted once (lane 0 of the group)
    const bool any_interval = gbyte(sync_on && (has_a || has_b)) != 0 || has_md;  // idx_end = any_interval ? NC : DOF (:203-215)
    auto blocked_by_me = [&](double ts) {  // block.rs:157-172
        return sync_on && (ts < t_min || (has_a && ts > a_left && ts < a_right) || (has_b && ts > b_left && ts < b_right));
    };
    auto value_ok = [&](double ts) {  // one verdict per candidate VALUE (:227-240)
        const bool blocked = gbyte(blocked_by_me(ts)) != 0;
        return !(blocked || ts < md || isinf(ts));
    };
    // bit s * 8 + d = candidate s of DoF d, bit 24 = the minimum duration: bit order == the reference's index order s * DOF + d
    // (the interval ends and the minimum duration are only in the sorted set when any_interval)
    auto equal_mask = [&](double v) {
        const uint32_t e0 = gbyte(mine && c0 == v), e1 = gbyte(mine && c1 == v), e2 = gbyte(mine && c2 == v);
        return any_interval ? (e0 | (e1 << 8) | (e2 << 16) | ((c3 == v) ? (1u << 24) : 0u)) : e0;
    };
    const double T = group_max8(c0);
    int below = mine && c0 < T;
    if (any_interval) below += (mine && c1 < T) + (mine && c2 < T) + (d == 0 && c3 < T);
    const int L = group_sum8(below);
    int which_bit = -1;
    double ts = T;
    {
        const uint32_t ties = equal_mask(T);
        const int skip = (D - 1 - L > 0) ? (D - 1 - L) : 0;
        const bool ok = value_ok(T);
        if (ok) which_bit = (int)__fns(ties, 0, skip + 1);  // the tie the scan meets first at or after sorted position DOF - 1
    }
    {
        // any_interval: the distinct values above T, ascending. The loop is warp-uniform: a group that is finished (or has no
        // interval at all) keeps voting with stale values and ignores the results.
        double cur = T;
        bool open = any_interval && which_bit < 0;
        while (__any_sync(0xffffffffu, open)) {
            double mn = RK_INF;
            if (mine) {
                if (c1 > cur && c1 < mn) mn = c1;
                if (c2 > cur && c2 < mn) mn = c2;
                if (c0 > cur && c0 < mn) mn = c0;
            }
            if (c3 > cur && c3 < mn) mn = c3;
            mn = group_min8(mn);
            const bool ok = value_ok(mn);
            const uint32_t eq = equal_mask(mn);
            if (open) {
                if (isinf(mn)) { open = false; }  // only infinite candidates are left: none is ever accepted
                else if (ok) { which_bit = __ffs(eq) - 1; ts = mn; open = false; }
                else cur = mn;
            }
        }
    }
    {
        // no interval: the tail of the (fresh) index array is zero, so after the DOF sorted minimum durations the scan meets
        // candidate 0 again and again (Q4)
        const double first = __shfl_sync(0xffffffffu, c0, gshift);
        const bool ok = value_ok(first);
        if (!any_interval && which_bit < 0 && ok) { which_bit = 0; ts = first; }
    }
    const bool zero_limits_any = gbyte(mine && (m & M_ZERO_LIMITS)) != 0;
    const bool all_none = gbyte(mine && !none) == 0;

    if (!done) {
        if (which_bit < 0) {  // :492-518
            status = zero_limits_any ? RK_RESULT_ERROR_ZERO_LIMITS : RK_RESULT_ERROR_SYNCHRONIZATION_CALCULATION;
            detail = 2 << 8;
            done = true;
        } else {
            duration = ts;
            if (which_bit < 24) {
                limiting = which_bit & 7;
                if (d == limiting) src = which_bit >> 3;  // 0 p_min, 1 a.profile, 2 b.profile
            }
        }
    }
    // None synchronisation (:520-528): in DoF order, a larger minimum duration takes over
    {
        const bool none_on = enabled && none;
        const double tm = none_on ? t_min : -RK_INF;
        const double tmax = group_max8(tm);
        const uint32_t at_max = gbyte(none_on && t_min == tmax);
        if (!done) {
            if (none_on) src = RK_SRC_STEP1_MIN;
            if (tmax > duration) { duration = tmax; limiting = __ffs(at_max) - 1; }
        }
    }
    if (!done) {
        if (duration > 7.6e3) {  // :531-533
            status = RK_RESULT_ERROR_TRAJECTORY_DURATION;
            done = true;
        } else if (fabs(duration - 0.0) < EPS) {  // :535-541: every DoF, disabled ones included
            src = RK_SRC_STEP1_MIN;
            done = true;
        } else if (all_none) {  // :543-550: all DoFs unsynchronised
            done = true;
        }
    }
    // ---- time synchronisation dispatch (:714-747), one DoF per lane
    if (!done && enabled && !(d == limiting || none)) {
        const double t_profile = duration - bdur - 0.0;
        bool rest = false;
        if (sync_d == RK_SYNC_TIME_IF_NECESSARY) {
            const int64_t g = (int64_t)d * P.in_n + (ig + P.in_shift);
            rest = fabs(P.vf[g]) < EPS && fabs(P.af[g]) < EPS;
        }
        // extremal profile re-use; b is only looked at when a is absent (Q3)
        if (rest || fabs(t_profile - t_min) < 2.0 * EPS) src = RK_SRC_STEP1_MIN;
        else if (has_a && fabs(t_profile - a_right) < 2.0 * EPS) src = RK_SRC_STEP1_A;
        else if (!has_a && has_b && fabs(t_profile - b_right) < 2.0 * EPS) src = RK_SRC_STEP1_B;
        else src = RK_SRC_STEP2;
    }
    if (tvalid && d == 0) {
        P.status[ig] = status;
        P.duration[ig] = duration;
        P.limiting[ig] = limiting;
        P.detail[ig] = detail;
    }
    // Per-DoF results leave the block DoF-major — warp w writes DoF w of the block's 32 consecutive trajectories — so that the
    // chain list keeps runs of consecutive trajectories per DoF like k_sync's (in (trajectory, DoF) order the Step-2 kernels
    // that walk the list lose the locality of their structure-of-arrays reads: measured -6 % on the whole step).
    src_of[d][threadIdx.x >> 3] = (uint8_t)((mine ? src : RK_SRC_NONE) | ((mine && (m & M_KIND_MASK) == K_POS3) ? 0x80 : 0));
    __syncthreads();
    {
        const int d2 = threadIdx.x >> 5, t2 = threadIdx.x & 31;
        const int64_t il2 = (int64_t)blockIdx.x * 32 + t2;
        const bool mine2 = il2 < P.cn && d2 < D;
        const uint8_t v = src_of[d2][t2];
        const uint32_t wi2 = (uint32_t)((int64_t)d2 * P.cn + il2);
        if (mine2) P.wsrc[wi2] = v & 0x7F;
        append_unsolved(P.list[0], P.counters + 0, mine2 && v == (0x80 | RK_SRC_STEP2), wi2);
    }
}

// ------------------------------------------------------------------------------------------------ Step-2 chain
// The 16 stages of position_third_step2.rs:2449-2482, with the VEL family split into its UDDU and UDUD halves (each half
// is a pure function of the inputs; UDUD only runs for items the UDDU half did not solve): 18 launches, families
// 0 ACC0_ACC1_VEL, 1 VEL/UDDU, 8 VEL/UDUD, 2 ACC0_VEL, 3 ACC1_VEL, 4 ACC0_ACC1, 5 ACC0, 6 ACC1, 7 NONE.
#define RK_S2_STEPS 18

struct ItemS2 {
    DofIn x;
    Bound b;
    double tf;
    int d;
    int64_t ig, gi, wi;
};

__device__ __forceinline__ ItemS2 load_item_s2(const Params& P, int64_t wi, int64_t wstride) {
    ItemS2 it;
    it.wi = wi;
    it.d = (int)(wi / P.cn);
    it.ig = P.off + (wi - (int64_t)it.d * P.cn);
    it.gi = (int64_t)it.d * P.n + it.ig;
    double bdur;
    it.x = load_rec(P, wi, it.b, bdur);  // NB: it.x.p0 / v0 / a0 are the post-brake values
    it.tf = P.duration[it.ig] - bdur - 0.0;  // :723 (accel duration is always 0)
    (void)wstride;
    return it;
}

// One stage of the chain position_third_step2.rs:2449-2482 for one item. stage = 4 * group + family slot; groups 0, 2 use the
// first direction (up_first = pd > tf * v0), groups 1, 3 the other one.
template <int F>
__device__ __forceinline__ bool run_family(const S2& s, const Lim& L, Hit& h) {
    if (F == 0) return s2_acc0_acc1_vel(s, L, h);
    if (F == 1) return s2_vel_uddu(s, L, h);
    if (F == 8) return s2_vel_udud(s, L, h);
    if (F == 2) return s2_acc0_vel(s, L, h);
    if (F == 3) return s2_acc1_vel(s, L, h);
    if (F == 4) return s2_acc0_acc1(s, L, h);
    if (F == 5) return s2_acc0(s, L, h);
    if (F == 6) return s2_acc1(s, L, h);
    return s2_none(s, L, h);
}

// A solved Step-2 item leaves only its compact solution behind — t[7] and the jerk as ONE 64-byte record in the item's own
// (no longer needed) candidate area, and the code word; k_expand turns it into the stored result later, with full warps and
// coalesced stores. (Round 1 kept the eight doubles in structure-of-arrays rows: eight partially filled sectors per item
// whenever the work list does not run over consecutive items.)
__device__ __forceinline__ void put_solution(const Params& P, const ItemS2& it, int64_t wstride, const Hit& h) {
    double2* rec = reinterpret_cast<double2*>(P.vl + it.wi * VL_ITEM_DOUBLES + V_SOL);
    rec[0] = make_double2(h.t[0], h.t[1]);
    rec[1] = make_double2(h.t[2], h.t[3]);
    rec[2] = make_double2(h.t[4], h.t[5]);
    rec[3] = make_double2(h.t[6], h.jf);
    P.codes[it.gi] = pack_codes(h.lim, h.cs, h.dir, RK_SRC_STEP2);
    (void)wstride;
}

template <int F>
__device__ __forceinline__ bool run_stage(const Params& P, const ItemS2& it, bool second_direction) {
    S2 s;
    s.init(it.b, it.tf, it.x.j_max);
    const bool up_first = s.pd > it.tf * it.b.v0;
    const Lim L = lim_dir(up_first != second_direction, it.x.v_max, it.x.v_min, it.x.a_max, it.x.a_min, it.x.j_max);
    Hit h;
    if (!run_family<F>(s, L, h)) return false;
    put_solution(P, it, (int64_t)P.dof * P.cn, h);
    return true;
}

__device__ __forceinline__ void report_step2_failure(const Params& P, int d, int64_t ig) {
    P.status[ig] = RK_RESULT_ERROR_EXECUTION_TIME_CALCULATION;  // :819-826; every failing DoF stores the same value
    const int mine = (3 << 8) | d;                              // the lowest failing DoF index is kept in the detail word
    int seen = P.detail[ig];
    while (seen == 0 || mine < seen) {
        const int prev = atomicCAS(&P.detail[ig], seen, mine);
        if (prev == seen) break;
        seen = prev;
    }
}

// one item of a chain stage: true = solved (its compact solution and code word are stored)
template <int F>
__device__ __forceinline__ bool stage_unit(const Params& P, uint32_t wi, int64_t wstride, bool second_direction) {
    const ItemS2 it = load_item_s2(P, wi, wstride);
    return run_stage<F>(P, it, second_direction);
}
template <int F>
__device__ __noinline__ bool stage_exact(const Params& P, uint32_t wi, int64_t wstride, bool second_direction) {
    return stage_unit<F>(P, wi, wstride, second_direction);
}

// ------------------------------------------------------------------------------------------------ k2_stage
// `step` numbers the launches of the chain (list 0 of step 0 is filled by k_sync); list `step & 1` is read, the other written.
template <int F>
__global__ void __launch_bounds__(128, F == 0 ? RK_S0_MINB : RK_MIN_BLOCKS) k2_stage(const Params P, int step, int second_dir) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const uint32_t* list_in = P.list[step & 1];
    uint32_t* list_out = P.list[(step + 1) & 1];
    const uint32_t count = P.counters[step];
    const bool second_direction = second_dir != 0;
    // whole warps iterate together so that the ballot in append_unsolved sees all 32 lanes
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (count + stride - 1) / stride;
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t r = 0; r < rounds; ++r, k += stride) {
        bool unsolved = false;
        uint32_t wi = 0;
        if (k < count) {
            wi = list_in[k];
            RK_LAZY_BEGIN();
            unsolved = !stage_unit<F>(P, wi, wstride, second_direction);
#if RK_LAZY_DDIV
            if (redo_wanted(P, k)) unsolved = !rk_exact::stage_exact<F>(RK_EXACT_PARAMS(P), wi, wstride, second_direction);
#endif
        }
        append_unsolved(list_out, P.counters + step + 1, unsolved, wi);
    }
}

// ------------------------------------------------------------------------------------------------ k2_vel
// The two VEL halves in lockstep form (rk_step2_vel.cuh): same list protocol as k2_stage, every thread of the block takes
// part in every round (inactive threads carry item 0 of the list and do no work between the barriers).
#ifndef RK_LS_BLOCK
#define RK_LS_BLOCK 128
#endif
#ifndef RK_LS_MINB
#define RK_LS_MINB 6
#endif
template <int F>
__global__ void __launch_bounds__(RK_LS_BLOCK, RK_LS_MINB) k2_vel(const Params P, int step, int second_dir) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const uint32_t* list_in = P.list[step & 1];
    uint32_t* list_out = P.list[(step + 1) & 1];
    const uint32_t count = P.counters[step];
    const bool second_direction = second_dir != 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (count + stride - 1) / stride;
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t r = 0; r < rounds; ++r, k += stride) {
        const bool active = k < count;
        const uint32_t wi = active ? list_in[k] : list_in[0];
        const ItemS2 it = load_item_s2(P, wi, wstride);
        S2 s;
        s.init(it.b, it.tf, it.x.j_max);
        const bool up_first = s.pd > it.tf * it.b.v0;
        const Lim L = lim_dir(up_first != second_direction, it.x.v_max, it.x.v_min, it.x.a_max, it.x.a_min, it.x.j_max);
        Hit h;
        static_assert(F == 8, "the UDDU half runs as k2_vel_scan_uddu -> k2_polish -> k2_vel_check_uddu");
        const bool found = s2_vel_udud_lockstep(s, L, h, active);
        RK_PHASE_BARRIER();
        if (found) put_solution(P, it, wstride, h);
        append_unsolved(list_out, P.counters + step + 1, active && !found, wi);
    }
}

// ------------------------------------------------------------------------------------------------ VEL as scan -> polish -> check
// Half of the VEL stage's instructions used to be the safeguarded Newton iteration (shrink_interval), run by the ~10 lanes of a
// warp that had a bracket, for as long as the slowest of them needed (mean 10 iterations, slowest of 32 about 28). The stage
// is therefore cut in three:
//   k2_vel_scan    phases A-C per item (lockstep): polynomial, extrema, root candidates. Items without a candidate go straight
//                  to the next list; the others park polynomial and brackets in their scratch area, join the pending list,
//                  and every true bracket becomes a task of the polish queue.
//   k2_polish<N>   one lane per task, and a lane that has converged takes the next task while its neighbours keep
//                  iterating (refill in batches), so the iteration runs with nearly full warps.
//   k2_vel_check   pending items: candidates in order, root -> Newton step -> profile check; first pass wins.
// Counters: [step + 1] the next list as before, [32 + 4 * step + {0, 1, 2}] = pending items, polish tasks, polish cursor.
__device__ __forceinline__ double* vslot(const Params& P, int64_t wi, int64_t wstride, int slot) {
    (void)wstride;
    return P.vl + wi * VL_ITEM_DOUBLES + slot;
}

__global__ void __launch_bounds__(RK_LS_BLOCK, RK_LS_MINB) k2_vel_scan_uddu(const Params P, int step, int second_dir) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const uint32_t* list_in = P.list[step & 1];
    uint32_t* list_out = P.list[(step + 1) & 1];
    const uint32_t count = P.counters[step];
    const bool second_direction = second_dir != 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (count + stride - 1) / stride;
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t r = 0; r < rounds; ++r, k += stride) {
        const bool active = k < count;
        const uint32_t wi = active ? list_in[k] : list_in[0];
        const ItemS2 it = load_item_s2(P, wi, wstride);
        S2 s;
        s.init(it.b, it.tf, it.x.j_max);
        const bool up_first = s.pd > it.tf * it.b.v0;
        const Lim L = lim_dir(up_first != second_direction, it.x.v_max, it.x.v_min, it.x.a_max, it.x.a_min, it.x.j_max);
        Hit h;
        double pol[6], c_lo[6], c_hi[6];
        int n_cand;
        const bool found = vel_uddu_candidates(s, L, h, active, pol, c_lo, c_hi, n_cand);
        RK_PHASE_BARRIER();
        if (found) put_solution(P, it, wstride, h);
        const bool pending = active && !found && n_cand > 0;
        int n_task = 0;
        if (pending) {
#pragma unroll
            for (int q = 0; q < 5; ++q) *vslot(P, wi, wstride, V_POL + q) = pol[q + 1];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                if (c < n_cand) {
                    *vslot(P, wi, wstride, V_LO + c) = c_lo[c];
                    *vslot(P, wi, wstride, V_HI + c) = c_hi[c];
                    if (c_lo[c] != c_hi[c]) ++n_task; else *vslot(P, wi, wstride, V_ROOT + c) = c_hi[c];
                }
            }
            P.wqmask[wi] = (uint32_t)n_cand;
        }
        // three appends (next list, pending list, polish tasks): the three reservations are issued back to back by one lane, so
        // the warp waits for one atomic round trip instead of three (ncu: 10.7 % of this kernel's samples sat on the atomics)
        const bool to_next = active && !found && n_cand == 0;
        const unsigned m_next = __ballot_sync(0xffffffffu, to_next), m_pend = __ballot_sync(0xffffffffu, pending);
        const int lane = threadIdx.x & 31;
        int incl = n_task;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const int n_tasks_warp = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t b_next = 0, b_pend = 0, b_task = 0;
        if (lane == 0) {
            if (m_next) b_next = atomicAdd(P.counters + step + 1, (uint32_t)__popc(m_next));
            if (m_pend) b_pend = atomicAdd(P.counters + 32 + 4 * step, (uint32_t)__popc(m_pend));
            if (n_tasks_warp > 0) b_task = atomicAdd(P.counters + 33 + 4 * step, (uint32_t)n_tasks_warp);
        }
        b_next = __shfl_sync(0xffffffffu, b_next, 0);
        b_pend = __shfl_sync(0xffffffffu, b_pend, 0);
        b_task = __shfl_sync(0xffffffffu, b_task, 0);
        const unsigned below = (1u << lane) - 1u;
        if (to_next) list_out[b_next + __popc(m_next & below)] = wi;
        if (pending) P.pend[b_pend + __popc(m_pend & below)] = wi;
        uint32_t e = b_task + (uint32_t)(incl - n_task);
        if (pending) {
#pragma unroll
            for (int c = 0; c < 6; ++c)
                if (c < n_cand && c_lo[c] != c_hi[c]) P.rq_item[e++] = make_uint2(wi, (uint32_t)c);
        }
    }
}

// Polish queue: task (item, c) = bracket [lo[c], hi[c]] of the item's monic polynomial of N coefficients (DERIV: of its monic
// derivative, the first pass of the UDUD half, slots V1_*). Result: root[c].
#ifndef RK_REFILL
#define RK_REFILL 16
#endif
#ifndef RK_PO_MINB
#define RK_PO_MINB 10  // 48 registers (+0.3 % on the step against 8 blocks / 64 registers)
#endif
template <int N, bool DERIV>
__global__ void __launch_bounds__(128, RK_PO_MINB) k2_polish(const Params P, int counter_index, int cursor_index) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const uint32_t count = P.counters[counter_index];
    uint32_t* cursor = P.counters + cursor_index;  // tasks are claimed dynamically: the lanes of the whole grid stay balanced
    const int lane = threadIdx.x & 31;
    ShrinkState<N> st;
    bool have = false, exhausted = false;
    double* out = nullptr;
    while (true) {
        const bool want = !have && !exhausted;
        const unsigned wanting = __ballot_sync(0xffffffffu, want);
        const unsigned busy = __ballot_sync(0xffffffffu, have);
        if (wanting == 0 && busy == 0) break;
        // refill in batches: a quarter of the warp idle, or nobody iterating
        if (wanting != 0 && (__popc(wanting) >= RK_REFILL || busy == 0)) {
            uint32_t base = 0;
            const int leader = __ffs(wanting) - 1;
            if (lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(wanting));
            base = __shfl_sync(0xffffffffu, base, leader);
            const uint32_t k = base + (uint32_t)__popc(wanting & ((1u << lane) - 1u));
            if (want && k >= count) exhausted = true;
            if (want && k < count) {
                const uint2 task = P.rq_item[k];
                const int64_t wi = task.x;
                const int c = (int)task.y;
                if (DERIV) {
                    double pol[N + 1];
                    pol[0] = 1.0;
#pragma unroll
                    for (int q = 0; q < N; ++q) pol[q + 1] = *vslot(P, wi, wstride, V_POL + q);
                    poly_monic_deri<N + 1>(pol, st.p);
                } else {
                    st.p[0] = 1.0;
#pragma unroll
                    for (int q = 0; q < N - 1; ++q) st.p[q + 1] = *vslot(P, wi, wstride, V_POL + q);
                }
                const double lo = *vslot(P, wi, wstride, (DERIV ? V1_LO : V_LO) + c);
                const double hi = *vslot(P, wi, wstride, (DERIV ? V1_HI : V_HI) + c);
                out = vslot(P, wi, wstride, (DERIV ? V1_ROOT : V_ROOT) + c);
                if (st.start(lo, hi)) *out = st.rts;
                else have = true;
            }
        }
        if (have && st.step()) {
            *out = st.rts;
            have = false;
        }
        if (have && st.step()) {  // a second iteration per trip: the refill test (two ballots, a count) is a tenth of a trip
            *out = st.rts;
            have = false;
        }
    }
}

// one pending item of the UDDU half: its polished candidates in order, the first valid one is stored; true = solved
__device__ __forceinline__ bool vel_check_unit(const Params& P, uint32_t wi, int64_t wstride, bool second_direction) {
    const ItemS2 it = load_item_s2(P, wi, wstride);
    S2 s;
    s.init(it.b, it.tf, it.x.j_max);
    const bool up_first = s.pd > it.tf * it.b.v0;
    const Lim L = lim_dir(up_first != second_direction, it.x.v_max, it.x.v_min, it.x.a_max, it.x.a_min, it.x.j_max);
    const int n_cand = (int)P.wqmask[wi];
    Hit h;
    bool found = false;
#pragma unroll 1
    for (int c = 0; c < n_cand && !found; ++c) found = vel_uddu_root(s, L, h, *vslot(P, wi, wstride, V_ROOT + c));
    if (found) put_solution(P, it, wstride, h);
    return found;
}

__global__ void __launch_bounds__(128, RK_VC_MINB) k2_vel_check_uddu(const Params P, int step, int second_dir) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    uint32_t* list_out = P.list[(step + 1) & 1];
    const uint32_t count = P.counters[32 + 4 * step];
    const bool second_direction = second_dir != 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (count + stride - 1) / stride;
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t r = 0; r < rounds; ++r, k += stride) {
        bool unsolved = false;
        uint32_t wi = 0;
        if (k < count) {
            wi = P.pend[k];
            unsolved = !vel_check_unit(P, wi, wstride, second_direction);
        }
        append_unsolved(list_out, P.counters + step + 1, unsolved, wi);
    }
}

// ------------------------------------------------------------------------------------------------ k2_tail
// The last eight launches of the chain (ACC0_ACC1, ACC0, ACC1, NONE in the first and then in the second direction,
// position_third_step2.rs:2466-2481) see less than 1 % of the items: as separate kernels they are nothing but launch and
// single-thread latency. k2_tail evaluates the eight (direction, family) variants of an item SIDE BY SIDE — warp w of a
// block runs family 4 + w for 16 items x 2 directions, so every warp stays in one family's code — and the first variant in
// the reference's order that produced a profile stores it (the families only read the item, so evaluating one that the
// reference would not have reached changes nothing). An item no variant solves gets k2_fail's bookkeeping. `step` is the
// launch index the first of the fused stages would have had (its input list and counter).
#ifndef RK_FUSE_TAIL
#define RK_FUSE_TAIL 1
#endif
template <int F>
__device__ __noinline__ bool tail_family(const ItemS2& it, bool second_direction, Hit& h) {
    S2 s;
    s.init(it.b, it.tf, it.x.j_max);
    const bool up_first = s.pd > it.tf * it.b.v0;
    const Lim L = lim_dir(up_first != second_direction, it.x.v_max, it.x.v_min, it.x.a_max, it.x.a_min, it.x.j_max);
    return run_family<F>(s, L, h);
}

__global__ void __launch_bounds__(128, RK_MIN_BLOCKS) k2_tail(const Params P, int step) {
    RK_PDL_ENTER();
    __shared__ uint32_t solved_by[16];  // per item of the block: bit (4 * direction + family slot) set = that variant has a profile
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const uint32_t* list_in = P.list[step & 1];
    const uint32_t count = P.counters[step];
    const int fam = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = lane & 15, dir = lane >> 4;
    const int variant = 4 * dir + fam;
    for (uint32_t base = blockIdx.x * 16u; base < count; base += gridDim.x * 16u) {  // uniform per block
        if (threadIdx.x < 16) solved_by[threadIdx.x] = 0u;
        __syncthreads();
        const uint32_t k = base + slot;
        const bool live = k < count;
        bool ok = false;
        Hit h;
        ItemS2 it;
        uint32_t wi = 0;
        if (live) {
            wi = list_in[k];
            it = load_item_s2(P, wi, wstride);
            if (fam == 0) ok = tail_family<4>(it, dir != 0, h);
            else if (fam == 1) ok = tail_family<5>(it, dir != 0, h);
            else if (fam == 2) ok = tail_family<6>(it, dir != 0, h);
            else ok = tail_family<7>(it, dir != 0, h);
            if (ok) atomicOr(&solved_by[slot], 1u << variant);
        }
        __syncthreads();
        if (live) {
            const uint32_t m = solved_by[slot];
            if (ok && (m & ((1u << variant) - 1u)) == 0u) put_solution(P, it, wstride, h);  // first in the reference's order
            if (m == 0u && variant == 0) {
                report_step2_failure(P, it.d, it.ig);
                P.wsrc[wi] = 0xFF;  // RK_SRC_FAILED: k_expand stores an idle profile
                P.codes[it.gi] = pack_codes(L_NONE, UDDU, DIR_UP, RK_SRC_STEP2);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ k2_fail
__global__ void __launch_bounds__(128) k2_fail(const Params P) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const uint32_t count = P.counters[RK_S2_STEPS];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        const uint32_t wi = P.list[RK_S2_STEPS & 1][k];
        const int d = (int)(wi / P.cn);
        const int64_t ig = P.off + (wi - (int64_t)d * P.cn);
        report_step2_failure(P, d, ig);
        P.wsrc[wi] = 0xFF;  // RK_SRC_FAILED: k_expand stores an idle profile
        P.codes[(int64_t)d * P.n + ig] = pack_codes(L_NONE, UDDU, DIR_UP, RK_SRC_STEP2);
    }
    (void)wstride;
}

// ------------------------------------------------------------------------------------------------ k_expand
// Last kernel, one thread per item: the compact result of every item — a Step-2 solution, a re-used extremal profile, a
// phase-synchronised profile, nothing at all (disabled DoF, error status, failed chain) — is integrated into the profile
// arrays the reference's check functions leave behind (profile.rs:76-118) and stored. Dense and coalesced: a warp writes 59
// full 256-byte rows. The cheap non-third-order Step 2 variants are solved right here.
#define RK_SRC_FAILED 0xFF
#ifndef RK_EX_MINB
#define RK_EX_MINB 4  // 128 registers, no spills: 162 -> 144 us per 300 k-trajectory chunk against the unconstrained 168
#endif
__global__ void __launch_bounds__(128, RK_EX_MINB) k_expand(const Params P) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    {
        const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (wi >= wstride) return;
        const uint32_t meta = P.wmeta[(int64_t)WM_META * wstride + wi];
        const int kind = meta & M_KIND_MASK;
        const int src = P.wsrc[wi];
        const ItemS2 it = load_item_s2(P, wi, wstride);
        ProfileSink out{P.prof, (int64_t)P.dof * P.n, it.gi};
        if (src == RK_SRC_NONE || kind == K_NONE) {
            // disabled DoF (calculator_target.rs:313-323) or a trajectory that ended with an error status
            if (P.compact) {
                out.put(C_META, compact_meta(kind, kind, CM_IDLE_STATE, false, false, UDDU));
            } else {
                const int64_t g = (int64_t)it.d * P.in_n + (it.ig + P.in_shift);
                store_idle_profile(out, P.p0[g], P.v0[g], P.a0[g]);
            }
            P.codes[it.gi] = pack_codes(L_NONE, UDDU, DIR_UP, RK_SRC_NONE);
            return;
        }
        if (!(meta & M_ENABLED)) {
            // A disabled DoF whose profile the reference overwrites with its block's p_min (zero duration :535-541, the
            // single-DoF shortcut :475-481, a disabled limiting DoF): the block of a disabled DoF on a fresh calculator holds
            // Profile::default().
            if (P.compact) out.put(C_META, compact_meta(kind, kind, CM_IDLE_ZERO, false, false, UDDU));
            else store_idle_profile(out, 0.0, 0.0, 0.0);
            P.codes[it.gi] = pack_codes(L_NONE, UDDU, DIR_UP, src);
            return;
        }
        if (src == RK_SRC_FAILED) {  // k2_fail has set the status and the code word
            if (P.compact) out.put(C_META, compact_meta(kind, kind, CM_IDLE_ZERO, false, false, UDDU));
            else store_idle_profile(out, 0.0, 0.0, 0.0);
            return;
        }
        double t[7], ctrl, ctrl2;
        int lim, cs, dir, lim_check, store_kind = kind;
        bool pf_from_p7 = false, ok = true;
        if (src == RK_SRC_STEP2 && kind == K_POS3) {  // solved by a stage of the chain (put_solution)
            const double2* rec = reinterpret_cast<const double2*>(P.vl + wi * VL_ITEM_DOUBLES + V_SOL);
            const double2 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3];
            t[0] = r0.x; t[1] = r0.y; t[2] = r1.x; t[3] = r1.y; t[4] = r2.x; t[5] = r2.y; t[6] = r3.x;
            ctrl = r3.y;
            ctrl2 = 0.0;
            const uint32_t c = P.codes[it.gi];
            lim = c & 0xFF; cs = (c >> 8) & 0xFF; dir = (c >> 16) & 0xFF;
            lim_check = lim;
        } else if (src == RK_SRC_STEP2) {
            Cand c;
            if (kind == K_POS2) { ok = step2_pos2(it.b, it.tf, it.x.v_max, it.x.v_min, it.x.a_max, it.x.a_min, c); pf_from_p7 = true; }
            else if (kind == K_POS1) ok = step2_pos1(it.b, it.tf, c);
            else if (kind == K_VEL3) { ok = step2_vel3(it.b, it.tf, it.x.a_max, it.x.a_min, it.x.j_max, c); pf_from_p7 = true; }
            else { ok = step2_vel2(it.b, it.tf, it.x.a_max, it.x.a_min, c); pf_from_p7 = true; }
            copy_t(t, c.t);
            ctrl = c.ctrl; ctrl2 = c.ctrl2; lim = c.lim; cs = c.cs; dir = c.dir;
            lim_check = lim;
        } else {
            const int slot = (src == RK_SRC_PHASE) ? 2 : src;
            uint32_t c;
            ws_get_cand(P, wi, wstride, slot, it.x.j_max, t, ctrl, ctrl2, c);
            lim = c & 0xFF; cs = (c >> 8) & 0xF; dir = (c >> 16) & 0x7F;
            lim_check = lim;
            if (src == RK_SRC_PHASE) {  // validated with ReachedLimits::None, labelled with the limiting DoF's limits (Q25)
                lim_check = L_NONE;
                store_kind = (int)((c >> 12) & 0xFu) - 1;
            }
        }
        if (ok) {
            if (P.compact) {
#pragma unroll
                for (int i = 0; i < 7; ++i) out.put(C_T + i, t[i]);
                out.put(C_CTRL, ctrl); out.put(C_CTRL2, ctrl2);
                out.put(C_META, compact_meta(kind, store_kind, CM_PROFILE, forces_a3_zero(lim_check), pf_from_p7, cs));
            } else {
                const double p7 = store_profile(out, store_kind, t, ctrl, ctrl2, cs, lim_check, it.b.p0, it.b.v0, it.b.a0, it.x.vf, it.x.af);
                if (pf_from_p7) out.put(S_PF, p7);
            }
            P.codes[it.gi] = pack_codes(lim, cs, dir, src);
        } else {
            report_step2_failure(P, it.d, it.ig);
            if (P.compact) out.put(C_META, compact_meta(kind, kind, CM_IDLE_ZERO, false, false, UDDU));
            else store_idle_profile(out, 0.0, 0.0, 0.0);
            P.codes[it.gi] = pack_codes(L_NONE, UDDU, DIR_UP, RK_SRC_STEP2);
        }
    }
}


// k_expand for the compact result layout (RK_RESULT_COMPACT): nothing is integrated, the decided t[7] and control values are
// copied into the record next to what k1_setup already stored (boundary states, brake, target). The common sources need
// neither the item's 96-byte record nor its inputs: a Step-2 solution is one 64-byte read (put_solution), a re-used Step-1
// profile is its candidate slot plus the jerk limit. Same decisions and code words as k_expand.
__global__ void __launch_bounds__(128, 8) k_expand_compact(const Params P) {
    RK_PDL_ENTER();
    const int64_t wstride = (int64_t)P.dof * P.cn;
    const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= wstride) return;
    const uint32_t meta = P.wmeta[(int64_t)WM_META * wstride + wi];
    const int src = P.wsrc[wi];
    const int kind = meta & M_KIND_MASK;
    const int d = (int)(wi / P.cn);
    const int64_t ig = P.off + (wi - (int64_t)d * P.cn);
    const int64_t gi = (int64_t)d * P.n + ig;
    // the common source (a Step-2 solution) is read before `src` is known: independent loads instead of a dependent chain
    const double2* sol = reinterpret_cast<const double2*>(P.vl + wi * VL_ITEM_DOUBLES + V_SOL);
    const double2 r0 = sol[0], r1 = sol[1], r2 = sol[2], r3 = sol[3];
    const uint32_t code_in = P.codes[gi];
    ProfileSink out{P.prof, (int64_t)P.dof * P.n, gi};
    if (src == RK_SRC_NONE || kind == K_NONE) {  // disabled DoF (calculator_target.rs:313-323) or an error status
        out.put(C_META, compact_meta(kind, kind, CM_IDLE_STATE, false, false, UDDU));
        P.codes[gi] = pack_codes(L_NONE, UDDU, DIR_UP, RK_SRC_NONE);
        return;
    }
    if (!(meta & M_ENABLED)) {  // a disabled DoF the reference overwrites with its block's (default) p_min
        out.put(C_META, compact_meta(kind, kind, CM_IDLE_ZERO, false, false, UDDU));
        P.codes[gi] = pack_codes(L_NONE, UDDU, DIR_UP, src);
        return;
    }
    if (src == RK_SRC_FAILED) {  // the chain has set the status and the code word
        out.put(C_META, compact_meta(kind, kind, CM_IDLE_ZERO, false, false, UDDU));
        return;
    }
    double t[7], ctrl, ctrl2 = 0.0;
    int lim, cs, dir, lim_check, store_kind = kind;
    bool pf_from_p7 = false, ok = true;
    if (src == RK_SRC_STEP2 && kind == K_POS3) {  // solved by a stage of the chain (put_solution)
        t[0] = r0.x; t[1] = r0.y; t[2] = r1.x; t[3] = r1.y; t[4] = r2.x; t[5] = r2.y; t[6] = r3.x;
        ctrl = r3.y;
        const uint32_t c = code_in;
        lim = c & 0xFF; cs = (c >> 8) & 0xFF; dir = (c >> 16) & 0xFF;
        lim_check = lim;
    } else if (src == RK_SRC_STEP2) {  // the cheap non-third-order Step 2 variants are solved right here
        const ItemS2 it = load_item_s2(P, wi, wstride);
        Cand c;
        if (kind == K_POS2) { ok = step2_pos2(it.b, it.tf, it.x.v_max, it.x.v_min, it.x.a_max, it.x.a_min, c); pf_from_p7 = true; }
        else if (kind == K_POS1) ok = step2_pos1(it.b, it.tf, c);
        else if (kind == K_VEL3) { ok = step2_vel3(it.b, it.tf, it.x.a_max, it.x.a_min, it.x.j_max, c); pf_from_p7 = true; }
        else { ok = step2_vel2(it.b, it.tf, it.x.a_max, it.x.a_min, c); pf_from_p7 = true; }
        copy_t(t, c.t);
        ctrl = c.ctrl; ctrl2 = c.ctrl2; lim = c.lim; cs = c.cs; dir = c.dir;
        lim_check = lim;
    } else {
        const int slot = (src == RK_SRC_PHASE) ? 2 : src;
        uint32_t c;
        ws_get_cand(P, wi, wstride, slot, P.rec[wi * 6 + 5].x, t, ctrl, ctrl2, c);  // [5].x = j_max
        lim = c & 0xFF; cs = (c >> 8) & 0xF; dir = (c >> 16) & 0x7F;
        lim_check = lim;
        if (src == RK_SRC_PHASE) {  // validated with ReachedLimits::None, labelled with the limiting DoF's limits (Q25)
            lim_check = L_NONE;
            store_kind = (int)((c >> 12) & 0xFu) - 1;
        }
    }
    if (ok) {
#pragma unroll
        for (int i = 0; i < 7; ++i) out.put(C_T + i, t[i]);
        out.put(C_CTRL, ctrl); out.put(C_CTRL2, ctrl2);
        out.put(C_META, compact_meta(kind, store_kind, CM_PROFILE, forces_a3_zero(lim_check), pf_from_p7, cs));
        P.codes[gi] = pack_codes(lim, cs, dir, src);
    } else {
        report_step2_failure(P, d, ig);
        out.put(C_META, compact_meta(kind, kind, CM_IDLE_ZERO, false, false, UDDU));
        P.codes[gi] = pack_codes(L_NONE, UDDU, DIR_UP, RK_SRC_STEP2);
    }
}

}  // namespace rk
#endif  // RK_H_PIPELINE_GO
