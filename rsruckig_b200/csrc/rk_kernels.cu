// rk_kernels.cu — the C ABI (include/ruckig_b200.h) and the kernels that are not part of the calculation pipeline.
// sm_100a only; compiled with --fmad=false.
//
//   rk_calculate_batch           launches the pipeline of rk_pipeline.cuh (see the map there) per chunk of ~4 Mi (DoF,
//                                trajectory) items; chunks rotate over two scratch sets / streams; host inputs are staged
//   k_validate                   InputParameter::validate with arbitrary flags          input_parameter.rs:251-420
//   k_at_time                    Trajectory::at_time for K sample times                  trajectory.rs:52-189
//   k_position_extrema,
//   k_first_time_at_position     the trajectory query helpers                            profile.rs:754-895
//   k_upd_dirty/gather/scatter/step   the closed-loop cycle (rk_update_batch)            ruckig.rs:266-321
//   k_fp64_peak, k_selftest_division  measurement / self-test helpers
//
// Dense arrays are laid out DoF-major / trajectory-fastest, so a warp works on 32 consecutive trajectories of one DoF and
// its global accesses are coalesced 256-byte rows; data read through compacted lists is packed per item (rk_params.cuh).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ruckig_b200.h"
#include "rk_pipeline.cuh"  // pass 1: namespace rk — every division exact on the spot (fast path + fall-back branch)

// Pass 2: the same headers into namespace rk_lazy, where a division by a `Den` raises a per-thread flag instead of branching
// to its fall-back (rk_math.cuh). The hot kernels are launched from rk_lazy and redo a flagged work item through the
// `*_exact` entry points of pass 1; everything else runs from namespace rk.
namespace rk_exact = rk;
#define RK_PASS2
#undef RK_LAZY_DDIV
#define RK_LAZY_DDIV 1
#define rk rk_lazy
#include "rk_pipeline.cuh"
#undef rk
#undef RK_LAZY_DDIV
#define RK_LAZY_DDIV 0


namespace rk {

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
    int32_t dof, steps, k_chunk;
    int32_t rewind;  // sample times do not grow (delta_time < 0 or NaN): every sample runs the general search
    int32_t n_starts;                  // > 0: start[y] = the accumulated time before the first sample of slice y
    double start[16];
    double time_start, delta_time;
    const double* prof;
    const double* duration;
    double *pos, *vel, *acc, *jerk;
    int32_t* section;
};

// trajectory.rs:52-189 for K sample times; time accumulates by repeated addition (ruckig.rs:293).
//
// Sample times only grow, so a thread spends long runs of samples inside one segment of its profile (one of the seven
// phases, the rest after its own end, the rest after the trajectory's end). The state to integrate from and the offsets to
// subtract live in registers per segment; a sample first re-evaluates the reference's own conditions for STAYING in the
// segment (three compares) and only falls back to the general search of trajectory.rs:60-140 when one of them fails. The
// first version ran the general search for every sample: 186 instructions per sample, 84 % of the issue slots busy at
// 3.2 TB/s; this form is bound by the output stream.
#ifndef RK_AT_BLOCK
#define RK_AT_BLOCK 128
#endif
#ifndef RK_AT_STORE
#define RK_AT_STORE __stcs
#endif
#ifndef RK_AT_MINB
#define RK_AT_MINB 0
#endif
// Reader side of the two trajectory layouts: the full Profile is read in place (row stride dof * n); a compact record
// (RK_RESULT_COMPACT, rk_store.cuh) is expanded once per thread into a 59-double local array (stride 1) by the same
// operations that fill the full layout, so every reader below sees bit-identical values either way.
__device__ __forceinline__ void load_compact_profile(const double* prof, int64_t gi, int64_t total, double* local) {
    double r[COMPACT_SLOTS];
#pragma unroll
    for (int s = 0; s < COMPACT_SLOTS; ++s) r[s] = prof[(int64_t)s * total + gi];
    expand_compact(r, ProfileSink{local, 1, 0});
}

#define RK_AT_MAX_STARTS 16
// PVA_ONLY: position, velocity and acceleration are all requested, jerk and section are not (the RL rollout). COMPACT:
// trajectory layout (the one-off expansion of a compact record would otherwise set the register count: 117 instead of 68,
// half the occupancy — hence the second launch bound).
//
// Per sample the fast path is: time += dt; td = time - A; t0 = td - B; three compares that restate the reference's own
// conditions for staying in the current segment; integrate; three streaming stores. (A, B) and the three limits are
// per-segment registers: trajectory end (A = brake + t_sum[6], B = 0, no limit), own end (A = brake duration, B = t_sum[6],
// limit `time < duration`), phase `index` (A = brake duration, B = t_sum[index - 1], limits duration, t_sum[6], t_sum[index]).
// x - 0.0 == x for every x, so an absent offset is a zero. ncu on the previous form of this loop: 128 warp instructions
// per sample, 75 % of the issue slots — it was issue-bound at 5.1 TB/s; this form needs about a third of that.
template <bool PVA_ONLY, bool COMPACT>
__global__ void __launch_bounds__(RK_AT_BLOCK, COMPACT ? 7 : RK_AT_MINB) k_at_time(const SampleParams S) {
    const int64_t total = (int64_t)S.dof * S.n;
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= total) return;
    const int d = (int)(gi / S.n);
    const int64_t i = gi - (int64_t)d * S.n;
    double local[COMPACT ? PROFILE_SLOTS : 1];
    if (COMPACT) load_compact_profile(S.prof, gi, total, local);
    const double* q = COMPACT ? local : S.prof + gi;
    const int64_t st = COMPACT ? 1 : total;

    const double duration = S.duration[i];
    const double ts6 = q[(int64_t)(S_TSUM + 6) * st];
    const double p7 = q[(int64_t)(S_P + 7) * st], v7 = q[(int64_t)(S_V + 7) * st], a7 = q[(int64_t)(S_A + 7) * st];
    const double bdur = q[(int64_t)S_BRAKE_DUR * st];
    const double bt0 = q[(int64_t)S_BRAKE_T * st];
    const double bsub = (bdur > 0.0) ? bdur : 0.0;  // the reference subtracts bdur only if it is positive
    const double end_off = bdur + ts6;
    // all seven t_sum in registers: the phase search used to walk t_sum[index] with one dependent global load per step,
    // zero-length phases included (ncu: 6.8 warp-cycles of long_scoreboard per issued instruction). Prefetching the start state
    // of the next phase as well was measured and changes nothing: the kernel is bound by the DRAM write stream by now.
    double ts[7];
#pragma unroll
    for (int r = 0; r < 6; ++r) ts[r] = q[(int64_t)(S_TSUM + r) * st];
    ts[6] = ts6;
    const double INF = RK_INF;

    // current segment: offsets, limits (h_lim = -inf: no segment, the next sample runs the general search), start state
    double A = 0.0, B = 0.0, d_lim = INF, t_lim = INF, h_lim = -INF;
    double sp = 0.0, sv = 0.0, sa = 0.0, sj = 0.0;
    int section = 0;

    // blockIdx.y picks a slice of the K sample times; the time of sample k is k + 1 repeated additions of delta_time
    // (ruckig.rs:293). The host has done the additions up to each slice's first sample (same IEEE additions, same result);
    // only with more than RK_AT_MAX_STARTS slices (tiny batches) a slice replays them itself.
    const int k_begin = blockIdx.y * S.k_chunk;
    const int k_end = (k_begin + S.k_chunk < S.steps) ? (k_begin + S.k_chunk) : S.steps;
    double time = S.time_start;
    if (S.n_starts > 0) {
        time = S.start[blockIdx.y];
    } else {
#pragma unroll 1
        for (int k = 0; k < k_begin; ++k) time += S.delta_time;
    }
    const int64_t stride = total * (int64_t)sizeof(double);
    const int64_t o0 = (int64_t)(k_begin - 1) * total + gi;  // pointers are advanced before each store
    char* out_p = S.pos ? (char*)(S.pos + o0) : nullptr;
    char* out_v = S.vel ? (char*)(S.vel + o0) : nullptr;
    char* out_a = S.acc ? (char*)(S.acc + o0) : nullptr;
    char* out_j = S.jerk ? (char*)(S.jerk + o0) : nullptr;
    int32_t* out_s = (S.section && d == 0) ? S.section + (int64_t)(k_begin - 1) * S.n + i : nullptr;
    const double dt = S.delta_time;
#pragma unroll 1
    for (int k = k_begin; k < k_end; ++k) {
        time += dt;
        if (!PVA_ONLY && S.rewind) h_lim = -INF;  // Trajectory::at_time is order-independent: every sample runs the general search
        double td = time - A;
        double t0 = td - B;
        const bool stay = !(time >= d_lim) & !(td >= t_lim) & (h_lim > td);
        if (!stay) {
            if (time >= duration) {  // :60-83; stays true while time grows
                section = 1;
                A = end_off; B = 0.0; d_lim = INF; t_lim = INF; h_lim = INF;
                t0 = time - end_off;
                sp = p7; sv = v7; sa = a7; sj = 0.0;
            } else {
                // cumulative_times[0] == duration > time, so the section index is 0 and t_diff == time (:85-99, Q13)
                section = 0;
                td = time;
                bool set = false;
                if (bdur > 0.0) {
                    if (td < bdur) {  // brake pre-trajectory (:103-121): no segment, it lasts a few samples at most
                        const int bi = (td < bt0) ? 0 : 1;
                        if (bi > 0) td -= bt0;
                        t0 = td;
                        sp = q[(int64_t)(S_BRAKE_P + bi) * st]; sv = q[(int64_t)(S_BRAKE_V + bi) * st];
                        sa = q[(int64_t)(S_BRAKE_A + bi) * st]; sj = q[(int64_t)(S_BRAKE_J + bi) * st];
                        h_lim = -INF;
                        set = true;
                    } else {
                        td -= bdur;
                    }
                }
                if (!set) {
                    A = bsub; d_lim = duration;
                    if (td >= ts6) {  // past this DoF's own end (:122-132); td >= ts6 keeps holding while time grows
                        t0 = td - ts6;
                        sp = p7; sv = v7; sa = a7; sj = 0.0;
                        B = ts6; t_lim = INF; h_lim = INF;
                    } else {
                        // "first i with t_sum[i] > t", default 6 (trajectory.rs:134-140)
                        int index = 6;
#pragma unroll
                        for (int r = 5; r >= 0; --r) if (ts[r] > td) index = r;
                        double t_hi = ts[6];  // t_sum[index]
                        double lsub = 0.0;    // t_sum[index - 1] if index > 0 (else nothing is subtracted: 0.0)
#pragma unroll
                        for (int r = 0; r < 6; ++r) if (index == r) t_hi = ts[r];
#pragma unroll
                        for (int r = 1; r < 7; ++r) if (index == r) lsub = ts[r - 1];
                        t0 = td - lsub;
                        sp = q[(int64_t)(S_P + index) * st]; sv = q[(int64_t)(S_V + index) * st];
                        sa = q[(int64_t)(S_A + index) * st]; sj = q[(int64_t)(S_J + index) * st];
                        B = lsub; t_lim = ts6;
                        // index == 6 is the default of the search, not a test of t_sum[6] > td: stay only on the reference's terms
                        h_lim = (t_hi > td) ? t_hi : -INF;
                    }
                }
            }
        }
        // util.rs:27-33
        const double tj = t0 * sj;
        const double p = sp + t0 * (sv + t0 * (sa / 2.0 + div6(tj)));
        const double v = sv + t0 * (sa + tj / 2.0);
        const double a = sa + tj;
        if (PVA_ONLY) {
            out_p += stride; RK_AT_STORE((double*)out_p, p);
            out_v += stride; RK_AT_STORE((double*)out_v, v);
            out_a += stride; RK_AT_STORE((double*)out_a, a);
        } else {
            if (out_p) { out_p += stride; RK_AT_STORE((double*)out_p, p); }
            if (out_v) { out_v += stride; RK_AT_STORE((double*)out_v, v); }
            if (out_a) { out_a += stride; RK_AT_STORE((double*)out_a, a); }
            if (out_j) { out_j += stride; RK_AT_STORE((double*)out_j, sj); }
            if (out_s) { out_s += S.n; *out_s = section; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ closed-loop update (rk_update_batch)
struct UpdateParams {
    int64_t n;
    int32_t dof;
    int32_t lim_per_dof;            // the caller's five limit arrays (in[6..10]) are [dof]; the stored copies are always [dof][n]
    // the caller's inputs ([dof][n]; v_min / a_min / enabled / min_dur may be null) and the copies of the last calculation
    const double* in[11];
    double* last[11];
    const uint8_t* enabled; uint8_t* last_enabled;
    const double* min_dur; double* last_min_dur;
    int32_t force_all;              // batch-level settings changed (or different optional arrays): everything is dirty
    uint8_t* initialized;           // [n]
    uint8_t* dirty;                 // [n]
    uint32_t* idx;                  // [n] compacted indices of the dirty environments
    uint32_t* count;                // [1]
    // gathered inputs of the dirty environments, [dof][m]
    double* gin[11];
    uint8_t* g_enabled; double* g_min_dur;
};

// Which environments must be re-solved: never calculated, or input != current_input (input_parameter.rs:190-211; f64 ==, so a
// NaN never compares equal).
__global__ void __launch_bounds__(256) k_upd_dirty(const UpdateParams U) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool dirty = false;
    if (i < U.n) {
        dirty = U.force_all != 0 || U.initialized[i] == 0;
        for (int d = 0; d < U.dof && !dirty; ++d) {
            const int64_t g = (int64_t)d * U.n + i;
#pragma unroll
            for (int f = 0; f < 11; ++f)
                if (U.in[f] && !(U.in[f][(f >= 6 && U.lim_per_dof) ? (int64_t)d : g] == U.last[f][g])) dirty = true;
            if (U.enabled && U.enabled[g] != U.last_enabled[g]) dirty = true;
        }
        if (!dirty && U.min_dur) {
            const double a = U.min_dur[i], b = U.last_min_dur[i];
            // Option<f64>: None == None, Some(x) == Some(y) iff x == y; NaN encodes None at this boundary
            if (!((a != a && b != b) || a == b)) dirty = true;
        }
        U.dirty[i] = dirty ? 1 : 0;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, dirty);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(U.count, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (dirty) U.idx[base + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)i;
}

// The order in which the warps appended is arbitrary; environments are independent, so it only has to be the same order for
// gather and scatter.
__global__ void __launch_bounds__(256) k_upd_gather(const UpdateParams U, int64_t m) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * U.dof) return;
    const int d = (int)(t / m);
    const int64_t j = t - (int64_t)d * m;
    const int64_t g = (int64_t)d * U.n + U.idx[j];
#pragma unroll
    for (int f = 0; f < 11; ++f)
        if (U.in[f]) U.gin[f][t] = U.in[f][(f >= 6 && U.lim_per_dof) ? (int64_t)d : g];
    if (U.enabled) U.g_enabled[t] = U.enabled[g];
    if (d == 0 && U.min_dur) U.g_min_dur[j] = U.min_dur[U.idx[j]];
}

struct ScatterParams {
    int64_t n, m;
    int32_t dof, throw_mode, slots;
    const uint32_t* idx;
    const double* s_prof; const uint32_t* s_codes; const int32_t* s_status; const double* s_duration; const double* s_imd;
    const int32_t* s_limiting; const int32_t* s_detail;
    double* prof; uint32_t* codes; int32_t* status; double* duration; double* imd; int32_t* limiting; int32_t* detail;
};

// Results of the re-solved subset -> the environments' own trajectories; the input they were solved for becomes current_input.
__global__ void __launch_bounds__(256) k_upd_scatter(const ScatterParams C, const UpdateParams U) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= C.m * C.dof) return;
    const int d = (int)(t / C.m);
    const int64_t j = t - (int64_t)d * C.m;
    const int64_t i = C.idx[j];
    const int64_t g = (int64_t)d * C.n + i;
    const int64_t s_stride = (int64_t)C.dof * C.m, stride = (int64_t)C.dof * C.n;
#pragma unroll 1
    for (int r = 0; r < C.slots; ++r) C.prof[(int64_t)r * stride + g] = C.s_prof[(int64_t)r * s_stride + t];
    C.codes[g] = C.s_codes[t];
    C.imd[g] = C.s_imd[t];
    const int st = C.s_status[j];
    // Err(..) under the throwing handler leaves current_input untouched (the `?` of ruckig.rs:285): the environment is solved
    // again next cycle
    const bool thrown = C.throw_mode && (st == RK_RESULT_ERROR_INVALID_INPUT || st == RK_RESULT_ERROR_ZERO_LIMITS
                                         || st == RK_RESULT_ERROR_EXECUTION_TIME_CALCULATION || st == RK_RESULT_ERROR_SYNCHRONIZATION_CALCULATION);
    if (!thrown) {
#pragma unroll
        for (int f = 0; f < 11; ++f)
            if (U.in[f]) U.last[f][g] = U.gin[f][t];
        if (U.enabled) U.last_enabled[g] = U.g_enabled[t];
    }
    if (d == 0) {
        C.status[i] = st;
        C.duration[i] = C.s_duration[j];
        C.limiting[i] = C.s_limiting[j];
        C.detail[i] = C.s_detail[j];
        if (!thrown) {
            if (U.min_dur) U.last_min_dur[i] = U.g_min_dur[j];
            U.initialized[i] = 1;
        }
    }
}

struct StepParams {
    int64_t n;
    int32_t dof, throw_mode, reset_time, compact;
    double delta_time;
    const double* prof; const double* duration; const int32_t* status;
    const uint8_t* dirty;
    const double* time_in; double* time_out;
    double* last_p; double* last_v; double* last_a;      // current_input.current_* (pass_to_input, ruckig.rs:314)
    double* in_p; double* in_v; double* in_a;            // the caller's arrays, or null
    double *new_p, *new_v, *new_a, *new_j;
    int32_t* result; uint8_t* new_calc; double* time_user;
};

// Trajectory::at_time (trajectory.rs:52-189) for one sample time.
__device__ __forceinline__ void sample_once(const double* q, int64_t st, double duration, double time, double& p, double& v, double& a, double& j) {
    const double ts6 = q[(int64_t)(S_TSUM + 6) * st];
    const double bdur = q[(int64_t)S_BRAKE_DUR * st];
    double t0, p0, v0, a0, j0;
    if (time >= duration) {  // :60-83
        t0 = time - (bdur + ts6);
        p0 = q[(int64_t)(S_P + 7) * st]; v0 = q[(int64_t)(S_V + 7) * st]; a0 = q[(int64_t)(S_A + 7) * st]; j0 = 0.0;
    } else {
        double td = time;
        bool set = false;
        if (bdur > 0.0) {
            if (td < bdur) {  // :103-121
                const int bi = (td < q[(int64_t)S_BRAKE_T * st]) ? 0 : 1;
                if (bi > 0) td -= q[(int64_t)S_BRAKE_T * st];
                t0 = td;
                p0 = q[(int64_t)(S_BRAKE_P + bi) * st]; v0 = q[(int64_t)(S_BRAKE_V + bi) * st];
                a0 = q[(int64_t)(S_BRAKE_A + bi) * st]; j0 = q[(int64_t)(S_BRAKE_J + bi) * st];
                set = true;
            } else {
                td -= bdur;
            }
        }
        if (!set) {
            if (td >= ts6) {  // :122-132
                t0 = td - ts6;
                p0 = q[(int64_t)(S_P + 7) * st]; v0 = q[(int64_t)(S_V + 7) * st]; a0 = q[(int64_t)(S_A + 7) * st]; j0 = 0.0;
            } else {  // :134-140
                int index = 0;
                while (index < 6 && !(q[(int64_t)(S_TSUM + index) * st] > td)) ++index;
                if (index > 0) td -= q[(int64_t)(S_TSUM + index - 1) * st];
                t0 = td;
                p0 = q[(int64_t)(S_P + index) * st]; v0 = q[(int64_t)(S_V + index) * st];
                a0 = q[(int64_t)(S_A + index) * st]; j0 = q[(int64_t)(S_J + index) * st];
            }
        }
    }
    integrate(t0, p0, v0, a0, j0, p, v, a);
    j = j0;
}

// One control cycle of every environment: time += delta_time, at_time, pass_to_input, Working / Finished (ruckig.rs:292-321).
template <bool COMPACT>  // trajectory layout (rollouts keep the full layout; the compact instantiation carries the expansion)
__global__ void __launch_bounds__(256) k_upd_step(const StepParams T) {
    const int64_t total = (int64_t)T.dof * T.n;
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= total) return;
    const int d = (int)(gi / T.n);
    const int64_t i = gi - (int64_t)d * T.n;
    const int st = T.status[i];
    const bool fresh = T.dirty[i] != 0;
    const bool thrown = fresh && T.throw_mode && (st == RK_RESULT_ERROR_INVALID_INPUT || st == RK_RESULT_ERROR_ZERO_LIMITS
                                                  || st == RK_RESULT_ERROR_EXECUTION_TIME_CALCULATION
                                                  || st == RK_RESULT_ERROR_SYNCHRONIZATION_CALCULATION);
    double time = T.time_in[i];
    if (thrown) {  // Err(..): update returned before touching the output
        if (d == 0) {
            T.time_out[i] = time;
            if (T.result) T.result[i] = st;
            if (T.new_calc) T.new_calc[i] = 0;
            if (T.time_user) T.time_user[i] = time;
        }
        return;
    }
    if (fresh && T.reset_time) time = 0.0;
    time += T.delta_time;
    const double duration = T.duration[i];
    double p, v, a, j;
    if (COMPACT) {
        double local[COMPACT ? PROFILE_SLOTS : 1];
        load_compact_profile(T.prof, gi, total, local);
        sample_once(local, 1, duration, time, p, v, a, j);
    } else {
        sample_once(T.prof + gi, total, duration, time, p, v, a, j);
    }
    if (T.new_p) T.new_p[gi] = p;
    if (T.new_v) T.new_v[gi] = v;
    if (T.new_a) T.new_a[gi] = a;
    if (T.new_j) T.new_j[gi] = j;
    T.last_p[gi] = p; T.last_v[gi] = v; T.last_a[gi] = a;
    if (T.in_p) { T.in_p[gi] = p; T.in_v[gi] = v; T.in_a[gi] = a; }
    if (d == 0) {
        T.time_out[i] = time;
        if (T.result) T.result[i] = (time > duration) ? RK_RESULT_FINISHED : RK_RESULT_WORKING;
        if (T.new_calc) T.new_calc[i] = fresh ? 1 : 0;
        if (T.time_user) T.time_user[i] = time;
    }
}

// ------------------------------------------------------------------------------------------------ trajectory query helpers
// profile.rs:754-776
__device__ __forceinline__ void check_position_extremum(double t_ext, double t_sum, double t, double p, double v, double a, double j,
                                                        double& e_min, double& e_max, double& e_tmin, double& e_tmax) {
    if (0.0 < t_ext && t_ext < t) {
        double p_ext, v_ext, a_ext;
        integrate(t_ext, p, v, a, j, p_ext, v_ext, a_ext);
        if (a_ext > 0.0 && p_ext < e_min) { e_min = p_ext; e_tmin = t_sum + t_ext; }
        else if (a_ext < 0.0 && p_ext > e_max) { e_max = p_ext; e_tmax = t_sum + t_ext; }
    }
}

// profile.rs:778-806
__device__ __forceinline__ void check_step_for_position_extremum(double t_sum, double t, double p, double v, double a, double j,
                                                                 double& e_min, double& e_max, double& e_tmin, double& e_tmax) {
    if (p < e_min) { e_min = p; e_tmin = t_sum; }
    if (p > e_max) { e_max = p; e_tmax = t_sum; }
    if (j != 0.0) {
        const double dsc = a * a - 2.0 * j * v;
        if (fabs(dsc) < EPS) {
            check_position_extremum(sdiv(-a, j), t_sum, t, p, v, a, j, e_min, e_max, e_tmin, e_tmax);
        } else if (dsc > 0.0) {
            const double d_sqrt = sqrt(dsc);
            check_position_extremum(sdiv(-a - d_sqrt, j), t_sum, t, p, v, a, j, e_min, e_max, e_tmin, e_tmax);
            check_position_extremum(sdiv(-a + d_sqrt, j), t_sum, t, p, v, a, j, e_min, e_max, e_tmin, e_tmax);
        }
    }
}

struct QueryParams {
    int64_t n;
    int32_t dof, compact;
    const double* prof;
    double *e_min, *e_max, *e_tmin, *e_tmax;   // position extrema, [dof][n], nullable
    const double* position; double* time;      // first time at position, [dof][n]
};

// Trajectory::get_position_extrema (trajectory.rs:211-231, profile.rs:808-863), one thread per (DoF, trajectory)
__global__ void __launch_bounds__(256) k_position_extrema(const QueryParams Q) {
    const int64_t total = (int64_t)Q.dof * Q.n;
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= total) return;
    double local[PROFILE_SLOTS];
    if (Q.compact) load_compact_profile(Q.prof, gi, total, local);
    const double* q = Q.compact ? local : Q.prof + gi;
    const int64_t st = Q.compact ? 1 : total;
    double e_min = RK_INF, e_max = -RK_INF, e_tmin = 0.0, e_tmax = 0.0;
    const double bdur = q[(int64_t)S_BRAKE_DUR * st];
    const double bt0 = q[(int64_t)S_BRAKE_T * st], bt1 = q[(int64_t)(S_BRAKE_T + 1) * st];
    if (bdur > 0.0 && bt0 > 0.0) {
        check_step_for_position_extremum(0.0, bt0, q[(int64_t)S_BRAKE_P * st], q[(int64_t)S_BRAKE_V * st], q[(int64_t)S_BRAKE_A * st],
                                         q[(int64_t)S_BRAKE_J * st], e_min, e_max, e_tmin, e_tmax);
        if (bt1 > 0.0)
            check_step_for_position_extremum(bt0, bt1, q[(int64_t)(S_BRAKE_P + 1) * st], q[(int64_t)(S_BRAKE_V + 1) * st],
                                             q[(int64_t)(S_BRAKE_A + 1) * st], q[(int64_t)(S_BRAKE_J + 1) * st], e_min, e_max, e_tmin, e_tmax);
    }
    double t_current_sum = 0.0;
#pragma unroll 1
    for (int i = 0; i < 7; ++i) {
        if (i > 0) t_current_sum = q[(int64_t)(S_TSUM + i - 1) * st];
        check_step_for_position_extremum(t_current_sum + bdur, q[(int64_t)(S_T + i) * st], q[(int64_t)(S_P + i) * st], q[(int64_t)(S_V + i) * st],
                                         q[(int64_t)(S_A + i) * st], q[(int64_t)(S_J + i) * st], e_min, e_max, e_tmin, e_tmax);
    }
    const double pf = q[(int64_t)S_PF * st], ts6 = q[(int64_t)(S_TSUM + 6) * st];
    if (pf < e_min) { e_min = pf; e_tmin = ts6 + bdur; }
    if (pf > e_max) { e_max = pf; e_tmax = ts6 + bdur; }
    if (Q.e_min) Q.e_min[gi] = e_min;
    if (Q.e_max) Q.e_max[gi] = e_max;
    if (Q.e_tmin) Q.e_tmin[gi] = e_tmin;
    if (Q.e_tmax) Q.e_tmax[gi] = e_tmax;
}

// Trajectory::get_first_time_at_position (trajectory.rs:233-246, profile.rs:865-895) for one query position per (DoF,
// trajectory); NaN = None. The cubic's roots are visited in insertion order (Q9).
__global__ void __launch_bounds__(256) k_first_time_at_position(const QueryParams Q) {
    const int64_t total = (int64_t)Q.dof * Q.n;
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= total) return;
    double local[PROFILE_SLOTS];
    if (Q.compact) load_compact_profile(Q.prof, gi, total, local);
    const double* q = Q.compact ? local : Q.prof + gi;
    const int64_t st = Q.compact ? 1 : total;
    const double pt = Q.position[gi];
    double result = __longlong_as_double(0x7ff8000000000000LL);
    bool found = false;
#pragma unroll 1
    for (int i = 0; i < 7 && !found; ++i) {
        const double pi = q[(int64_t)(S_P + i) * st];
        const double before = (i > 0) ? q[(int64_t)(S_TSUM + i - 1) * st] : 0.0;
        if (fabs(pi - pt) < EPS) {
            result = 0.0 + before;
            found = true;
            break;
        }
        const double ti = q[(int64_t)(S_T + i) * st];
        if (ti == 0.0) continue;
        const Roots r = solve_cub<false>(sdiv(q[(int64_t)(S_J + i) * st], 6.0), q[(int64_t)(S_A + i) * st] / 2.0, q[(int64_t)(S_V + i) * st], pi - pt);
        for (int k = 0; k < r.n && !found; ++k) {
            const double t_ = r.get(k);
            if (0.0 < t_ && t_ <= ti) {
                result = 0.0 + t_ + before;
                found = true;
            }
        }
    }
    if (!found && fabs(q[(int64_t)S_PF * st] - pt) < 1e-9) result = 0.0 + q[(int64_t)(S_TSUM + 6) * st];
    Q.time[gi] = result;
}

// rk_trajectories_read on a compact trajectory set: slots [first, first + count) of the full layout, expanded on the fly.
__global__ void __launch_bounds__(256) k_read_expanded(const double* prof, int64_t total, int first, int count, double* dst) {
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= total) return;
    double local[PROFILE_SLOTS];
    load_compact_profile(prof, gi, total, local);
#pragma unroll 1
    for (int s = 0; s < count; ++s) dst[(int64_t)s * total + gi] = local[first + s];
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

// ------------------------------------------------------------------------------------------------ division self-test
__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// Operand generator: half of the draws are "ordinary" magnitudes (2^-40 .. 2^40), the rest are raw bit patterns, values
// at the edges of the fast-path window, and the special values the solvers produce on purpose (0, inf, NaN).
__device__ __forceinline__ double test_operand(uint64_t r) {
    const uint64_t kind = r & 15u;
    const uint64_t mant = (r >> 12) & 0x000fffffffffffffull;
    const uint64_t sign = (r >> 11) & 1ull;
    uint64_t e;
    if (kind < 8) e = 1023 - 40 + ((r >> 4) % 81);
    else if (kind < 10) e = (r >> 4) % 2047;                       // any finite exponent incl. subnormals
    else if (kind == 10) e = 523 - 3 + ((r >> 4) % 7);             // lower edge of the window
    else if (kind == 11) e = 1523 - 3 + ((r >> 4) % 7);            // upper edge
    else if (kind == 12) return sign ? -0.0 : 0.0;
    else if (kind == 13) return __longlong_as_double((long long)((sign << 63) | 0x7ff0000000000000ull));
    else if (kind == 14) return __longlong_as_double(0x7ff8000000000000LL);
    else { e = 1023; return __longlong_as_double((long long)((sign << 63) | (e << 52) | ((r >> 4) & 0xff)));  }  // near +-1 with few bits
    return __longlong_as_double((long long)((sign << 63) | (e << 52) | mant));
}

__global__ void __launch_bounds__(256) k_selftest_division(int64_t n, uint64_t seed, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = test_operand(mix64(seed + 2 * (uint64_t)i));
        const double b = test_operand(mix64(seed + 2 * (uint64_t)i + 1));
        const double ref = a / b;
        const Den d(b);
        const double q1 = ddiv(a, d);
        const double q2 = sdiv(a, b);
        const Den dn(-b);
        const double q3 = ddiv(a, dn);
        const double ref3 = a / (-b);
        const bool nan_ref = ref != ref;
        const bool ok1 = nan_ref ? (q1 != q1) : (__double_as_longlong(q1) == __double_as_longlong(ref));
        const bool ok2 = nan_ref ? (q2 != q2) : (__double_as_longlong(q2) == __double_as_longlong(ref));
        const bool ok3 = (ref3 != ref3) ? (q3 != q3) : (__double_as_longlong(q3) == __double_as_longlong(ref3));
        const double q6 = div6(a), ref6 = a / 6.0;  // the constant-reciprocal form used for t * j / 6 (util.rs:27-33)
        const bool ok6 = (ref6 != ref6) ? (q6 != q6) : (__double_as_longlong(q6) == __double_as_longlong(ref6));
        const double q3c = div3(a), ref3c = a / 3.0;  // the three divisions by 3 of the resolvent (roots.rs:237-252)
        const bool ok3c = (ref3c != ref3c) ? (q3c != q3c) : (__double_as_longlong(q3c) == __double_as_longlong(ref3c));
        const double q5 = div_lit<5>(a), ref5 = a / 5.0;  // poly_monic_deri of the quintic (roots.rs:389-397)
        const bool ok5 = (ref5 != ref5) ? (q5 != q5) : (__double_as_longlong(q5) == __double_as_longlong(ref5));
        bad += (ok1 ? 0 : 1) + (ok2 ? 0 : 1) + (ok3 ? 0 : 1) + (ok6 ? 0 : 1) + (ok3c ? 0 : 1) + (ok5 ? 0 : 1);
    }
    if (bad) atomicAdd(mismatches, bad);
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

#ifndef RK_CHUNK_MI
#define RK_CHUNK_MI 4
#endif
#ifndef RK_LANES
#define RK_LANES 2
#endif
#ifndef RK_BIG_LANES
#define RK_BIG_LANES 2  // scratch sets / streams used by batches of more than one full chunk (a third set: no gain, +12 GB)
#endif
#ifndef RK_SMALL_SPLIT
#define RK_SMALL_SPLIT 2  // 64 Ki x 7-DoF: 1.27 -> 1.19 ms; 4 parts on 2 or 4 streams measured slower
#endif
#ifndef RK_SMALL_SPLIT_MIN_ITEMS
#define RK_SMALL_SPLIT_MIN_ITEMS (192 * 1024)
#endif
#ifndef RK_HOST_RAMP
#define RK_HOST_RAMP 4  // host-buffer calls: the first chunk is computed in this many pieces (pipeline fill)
#endif
#ifndef RK_STAGE_BUFS
#define RK_STAGE_BUFS 4  // host-buffer calls: input chunks are copied this many chunks ahead of the kernels
#endif
constexpr int64_t CHUNK_ITEMS = (int64_t)RK_CHUNK_MI << 20;  // (DoF, trajectory) items per chunk of scratch: 4 Mi items = 8.4 GB of scratch

}  // namespace

struct rk_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    // scratch: two sets, so that consecutive chunks of a large batch run on two streams (the many short, latency-bound
    // kernels of one chunk overlap with the long ones of the other)
    struct Scratch {
        int64_t ws_items = 0;
        double* ws = nullptr;
        double2* rec = nullptr;
        double* vl = nullptr;
        uint32_t* wmeta = nullptr;
        uint8_t* wsrc = nullptr;
        uint32_t* wqmask = nullptr;
        uint2* rq_item = nullptr;
        double* rq_root = nullptr;
        uint32_t* list[2] = {nullptr, nullptr};
        uint32_t* pend = nullptr;
        uint32_t* counters = nullptr;
    } sc[RK_LANES];
    cudaStream_t lane_stream[RK_LANES] = {};  // [0] is `stream`
    cudaEvent_t ev_fork = nullptr, ev_join[RK_LANES] = {};
    int sm_count = 148;
    // staging for host-memory calls
    void* stage_dev = nullptr;
    size_t stage_dev_bytes = 0;
    void* stage_pinned = nullptr;
    size_t stage_pinned_bytes = 0;
    // host-buffer calculate: input chunks are copied on a second stream while the previous chunk is computed
    cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_d2h = nullptr;
    cudaEvent_t ev_copied[RK_STAGE_BUFS] = {}, ev_computed[RK_STAGE_BUFS] = {};
    void* chunk_in[RK_STAGE_BUFS] = {};
    size_t chunk_in_bytes = 0;
    double* lim_dev = nullptr;  // [5][RK_MAX_DOF]: the limit arrays of a host-buffer call with RK_LIMITS_PER_DOF
};

struct rk_trajectories {
    rk_handle* h = nullptr;
    int64_t n = 0;
    int32_t dof = 0;
    int32_t layout = RK_RESULT_FULL;  // RK_RESULT_FULL: 59 profile slots per (DoF, trajectory); RK_RESULT_COMPACT: 20 (rk_store.cuh)
    int slots() const { return layout == RK_RESULT_COMPACT ? (int)rk::COMPACT_SLOTS : (int)rk::PROFILE_SLOTS; }
    double* prof = nullptr;
    uint32_t* codes = nullptr;
    int32_t* status = nullptr;
    double* duration = nullptr;
    double* imd = nullptr;
    int32_t* limiting = nullptr;
    int32_t* detail = nullptr;
};

static bool poison_scratch() {
    static const bool on = [] { const char* e = std::getenv("RK_B200_POISON"); return e && e[0] == '1'; }();
    return on;
}

static void free_scratch(rk_handle::Scratch& c) {
    if (!c.ws) return;
    cudaFree(c.ws); cudaFree(c.rec); cudaFree(c.vl); cudaFree(c.wmeta); cudaFree(c.wsrc); cudaFree(c.wqmask); cudaFree(c.rq_item); cudaFree(c.rq_root);
    cudaFree(c.list[0]); cudaFree(c.list[1]); cudaFree(c.pend); cudaFree(c.counters);
    c = rk_handle::Scratch();
}

static rk_status ensure_scratch(rk_handle* h, int set, int64_t items) {
    rk_handle::Scratch& c = h->sc[set];
    if (items <= c.ws_items) return RK_OK;
    free_scratch(c);
    RK_CUDA(cudaMalloc(&c.ws, sizeof(double) * rk::WS_SLOTS * items));
    RK_CUDA(cudaMalloc(&c.vl, sizeof(double) * rk::VL_ITEM_DOUBLES * items));
    RK_CUDA(cudaMalloc(&c.rec, sizeof(double2) * 6 * items));
    RK_CUDA(cudaMalloc(&c.wmeta, sizeof(uint32_t) * rk::WM_WORDS * items));
    RK_CUDA(cudaMalloc(&c.wsrc, items));
    RK_CUDA(cudaMalloc(&c.wqmask, sizeof(uint32_t) * 10 * items));
    RK_CUDA(cudaMalloc(&c.rq_item, sizeof(uint2) * 12 * items));
    RK_CUDA(cudaMalloc(&c.rq_root, sizeof(double) * 8 * items));
    RK_CUDA(cudaMalloc(&c.list[0], sizeof(uint32_t) * items));
    RK_CUDA(cudaMalloc(&c.list[1], sizeof(uint32_t) * items));
    RK_CUDA(cudaMalloc(&c.pend, sizeof(uint32_t) * items));
    RK_CUDA(cudaMalloc(&c.counters, 512 + sizeof(rk::Params)));  // [0, 448): the work-list counters; [512, ...): the chunk's parameter block (Params::self)
    static_assert(sizeof(uint32_t) * RK_N_COUNTERS <= 512, "counters overlap the parameter block");
    c.ws_items = items;
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

// One launch of a pipeline kernel. With RK_PDL the launch carries the programmatic-stream-serialization attribute and the
// kernel itself waits for its predecessor (RK_PDL_ENTER).
template <typename... KA, typename... A>
static inline void rk_launch(void (*kern)(KA...), unsigned grid, unsigned block, cudaStream_t s, A... args) {
#if RK_PDL
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, KA(args)...);
#else
    kern<<<grid, block, 0, s>>>(KA(args)...);
#endif
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
    P.in_n = desc->n; P.in_shift = 0;
    if (desc->limits_layout != RK_LIMITS_PER_TRAJECTORY && desc->limits_layout != RK_LIMITS_PER_DOF)
        return fail(RK_ERR_INVALID_ARGUMENT, "limits_layout is neither RK_LIMITS_PER_TRAJECTORY nor RK_LIMITS_PER_DOF");
    P.lim_per_dof = desc->limits_layout == RK_LIMITS_PER_DOF;
    {   // test hook: the lazy kernels send every fourth work item through their exact redo path (tests/test_parity.py)
        static const int forced = [] { const char* e = std::getenv("RK_B200_FORCE_REDO"); return (e && e[0] == '1') ? 1 : 0; }();
        P.force_redo = forced;
    }
    return RK_OK;
}

// Copy the caller's host arrays into one device staging block and re-point P at it.
static rk_status stage_inputs(rk_handle* h, rk::Params& P, size_t extra_dev_bytes, char** extra_dev) {
    const size_t per = sizeof(double) * (size_t)P.dof * (size_t)P.n;
    const double** slots[11] = {&P.p0, &P.v0, &P.a0, &P.pf, &P.vf, &P.af, &P.v_max, &P.a_max, &P.j_max, &P.v_min, &P.a_min};
    const size_t per_lim = P.lim_per_dof ? sizeof(double) * (size_t)P.dof : per;  // slots 6..10 are the limit arrays
    size_t bytes = 0;
    for (int k = 0; k < 11; ++k) if (*slots[k]) bytes += ((k >= 6 ? per_lim : per) + 255) / 256 * 256;
    const size_t en_bytes = P.enabled ? (size_t)P.dof * (size_t)P.n : 0;
    const size_t md_bytes = P.min_dur ? sizeof(double) * (size_t)P.n : 0;
    const size_t in_bytes = bytes + md_bytes + ((en_bytes + 255) / 256) * 256;
    rk_status st = ensure_stage(h, in_bytes + extra_dev_bytes, 0);
    if (st != RK_OK) return st;
    char* dev = (char*)h->stage_dev;
    for (int k = 0; k < 11; ++k) {
        const double** s = slots[k];
        if (!*s) continue;
        const size_t sz = k >= 6 ? per_lim : per;
        RK_CUDA(cudaMemcpyAsync(dev, *s, sz, cudaMemcpyHostToDevice, h->stream));
        *s = (const double*)dev;
        dev += (sz + 255) / 256 * 256;
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
    RK_CUDA(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
    RK_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    RK_CUDA(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    RK_CUDA(cudaEventCreateWithFlags(&h->ev_d2h, cudaEventDisableTiming));
    h->lane_stream[0] = h->stream;
    for (int k = 1; k < RK_LANES; ++k) RK_CUDA(cudaStreamCreateWithFlags(&h->lane_stream[k], cudaStreamNonBlocking));
    RK_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    for (int k = 0; k < RK_LANES; ++k) RK_CUDA(cudaEventCreateWithFlags(&h->ev_join[k], cudaEventDisableTiming));
    for (int k = 0; k < RK_STAGE_BUFS; ++k) {
        RK_CUDA(cudaEventCreateWithFlags(&h->ev_copied[k], cudaEventDisableTiming));
        RK_CUDA(cudaEventCreateWithFlags(&h->ev_computed[k], cudaEventDisableTiming));
    }
    *out = h;
    return RK_OK;
}

rk_status rk_destroy(rk_handle* h) {
    if (!h) return RK_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int k = 0; k < RK_LANES; ++k) {
        free_scratch(h->sc[k]);
        if (k > 0 && h->lane_stream[k]) cudaStreamDestroy(h->lane_stream[k]);
        if (h->ev_join[k]) cudaEventDestroy(h->ev_join[k]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->stage_dev) cudaFree(h->stage_dev);
    if (h->stage_pinned) cudaFreeHost(h->stage_pinned);
    for (int k = 0; k < RK_STAGE_BUFS; ++k) {
        if (h->chunk_in[k]) cudaFree(h->chunk_in[k]);
        if (h->ev_copied[k]) cudaEventDestroy(h->ev_copied[k]);
        if (h->ev_computed[k]) cudaEventDestroy(h->ev_computed[k]);
    }
    if (h->lim_dev) cudaFree(h->lim_dev);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    if (h->ev_d2h) cudaEventDestroy(h->ev_d2h);
    cudaStreamDestroy(h->stream);
    delete h;
    return RK_OK;
}

rk_status rk_trajectories_create_ex(rk_handle* h, int64_t n, int32_t dof, int32_t result_layout, rk_trajectories** out) {
    if (!h || !out || n < 0 || dof < 1 || dof > RK_MAX_DOF) return fail(RK_ERR_INVALID_ARGUMENT, "bad arguments to rk_trajectories_create");
    if (result_layout != RK_RESULT_FULL && result_layout != RK_RESULT_COMPACT) return fail(RK_ERR_INVALID_ARGUMENT, "unknown result layout");
    RK_CUDA(cudaSetDevice(h->device));
    rk_trajectories* t = new rk_trajectories();
    t->h = h; t->n = n; t->dof = dof; t->layout = result_layout;
    const size_t items = (size_t)(n > 0 ? n : 1) * dof, nn = (size_t)(n > 0 ? n : 1);
    const size_t prof_bytes = sizeof(double) * (size_t)t->slots() * items;
    cudaError_t e = cudaMalloc(&t->prof, prof_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&t->codes, sizeof(uint32_t) * items);
    if (e == cudaSuccess) e = cudaMalloc(&t->imd, sizeof(double) * items);
    if (e == cudaSuccess) e = cudaMalloc(&t->status, sizeof(int32_t) * nn);
    if (e == cudaSuccess) e = cudaMalloc(&t->duration, sizeof(double) * nn);
    if (e == cudaSuccess) e = cudaMalloc(&t->limiting, sizeof(int32_t) * nn);
    if (e == cudaSuccess) e = cudaMalloc(&t->detail, sizeof(int32_t) * nn);
    if (e == cudaSuccess) e = cudaMemsetAsync(t->prof, 0, prof_bytes, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(t->duration, 0, sizeof(double) * nn, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(t->status, 0, sizeof(int32_t) * nn, h->stream);
    if (e != cudaSuccess) {  // nothing of a partially built object is leaked (the likely cause is a failed allocation)
        rk_trajectories_destroy(t);
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? RK_ERR_OUT_OF_MEMORY : RK_ERR_CUDA, std::string("rk_trajectories_create: ") + cudaGetErrorString(e));
    }
    *out = t;
    return RK_OK;
}

rk_status rk_trajectories_create(rk_handle* h, int64_t n, int32_t dof, rk_trajectories** out) {
    return rk_trajectories_create_ex(h, n, dof, RK_RESULT_FULL, out);
}

rk_status rk_trajectories_destroy(rk_trajectories* t) {
    if (!t) return RK_OK;
    cudaSetDevice(t->h->device);
    cudaStreamSynchronize(t->h->stream);
    cudaFree(t->prof); cudaFree(t->codes); cudaFree(t->imd); cudaFree(t->status); cudaFree(t->duration); cudaFree(t->limiting); cudaFree(t->detail);
    delete t;
    return RK_OK;
}

}  // extern "C"

// Ruckig::calculate for trajectories [first, first + desc->n) of caller arrays whose DoF rows are `stride` elements apart
// (rk_calculate_batch: first = 0, stride = n; rk_calculate_batch_multi: one index range per device of one host batch).
static rk_status calculate_impl(rk_handle* h, const rk_batch_desc* desc, const rk_batch_in* in_full, rk_trajectories* traj, const rk_batch_out* out_full,
                                int64_t stride, int64_t first) {
    if (!h || !traj) return fail(RK_ERR_INVALID_ARGUMENT, "handle / trajectories is NULL");
    if (!desc || !in_full) return fail(RK_ERR_INVALID_ARGUMENT, "desc / in is NULL");
    rk_batch_in in_shifted = *in_full;
    {
        const double** d11[11] = {&in_shifted.current_position, &in_shifted.current_velocity, &in_shifted.current_acceleration, &in_shifted.target_position,
                                  &in_shifted.target_velocity, &in_shifted.target_acceleration, &in_shifted.max_velocity, &in_shifted.max_acceleration,
                                  &in_shifted.max_jerk, &in_shifted.min_velocity, &in_shifted.min_acceleration};
        const int n_shift = desc->limits_layout == RK_LIMITS_PER_DOF ? 6 : 11;  // [dof] limit arrays belong to every part
        for (int k = 0; k < n_shift; ++k) if (*d11[k]) *d11[k] += first;
        if (in_shifted.enabled) in_shifted.enabled += first;
        if (in_shifted.minimum_duration) in_shifted.minimum_duration += first;
    }
    const rk_batch_in* in = &in_shifted;
    rk_batch_out out_shifted{};
    const rk_batch_out* out = nullptr;
    if (out_full) {
        out_shifted = *out_full;
        if (out_shifted.status) out_shifted.status += first;
        if (out_shifted.duration) out_shifted.duration += first;
        if (out_shifted.independent_min_durations) out_shifted.independent_min_durations += first;
        out = &out_shifted;
    }
    rk::Params P;
    rk_status st = fill_params(P, desc, in);
    if (st != RK_OK) return st;
    P.in_n = stride;
    if (traj->n != desc->n || traj->dof != desc->dof) return fail(RK_ERR_INVALID_ARGUMENT, "trajectories were created for another (n, dof)");
    if (desc->n == 0) return RK_OK;
    RK_CUDA(cudaSetDevice(h->device));
    const bool host = desc->memory_space == RK_MEM_HOST;
    P.prof = traj->prof; P.codes = traj->codes; P.status = traj->status; P.duration = traj->duration; P.imd = traj->imd;
    P.limiting = traj->limiting; P.detail = traj->detail;
    P.compact = traj->layout == RK_RESULT_COMPACT;

    // chunks of about CHUNK_ITEMS (DoF, trajectory) items, equally sized (8 Mi x 7 DoF = exactly 14 x 4 Mi items must not
    // become 14 chunks and a 15th of four trajectories)
    int64_t n_chunks = ((int64_t)desc->n * desc->dof + CHUNK_ITEMS - 1) / CHUNK_ITEMS;
    // a batch that fits one chunk but is large enough to fill the GPU several times over is still split, so that the
    // short list-driven kernels of one part run beside the long kernels of the other (latency of a 64 Ki batch)
    const bool small_split = n_chunks < RK_SMALL_SPLIT && (int64_t)desc->n * desc->dof >= (int64_t)RK_SMALL_SPLIT_MIN_ITEMS;
    if (small_split) n_chunks = RK_SMALL_SPLIT;
    const int64_t cn_max = (desc->n + n_chunks - 1) / n_chunks;
    // more than one chunk: consecutive chunks rotate over RK_LANES scratch sets and streams
    const int max_lanes = small_split ? RK_LANES : RK_BIG_LANES;
    const int lanes = (int)(n_chunks < max_lanes ? n_chunks : max_lanes);
    for (int k = 0; k < lanes; ++k) {
        st = ensure_scratch(h, k, cn_max * desc->dof);
        if (st != RK_OK) return st;
    }
    if (lanes > 1) {  // the extra streams join the caller-visible one at both ends
        RK_CUDA(cudaEventRecord(h->ev_fork, h->stream));
        for (int k = 1; k < lanes; ++k) RK_CUDA(cudaStreamWaitEvent(h->lane_stream[k], h->ev_fork, 0));
    }

    // the synchronisation stage as warp-level reductions (k_sync_lanes) where its preconditions hold, else the generic kernel
    bool sync_lanes = desc->dof <= 8 && !P.discrete;
    for (int d = 0; d < desc->dof; ++d) sync_lanes = sync_lanes && P.sync[d] != RK_SYNC_PHASE;
    {
        static const int forced = [] { const char* e = std::getenv("RK_B200_SYNC_KERNEL"); return e ? (e[0] == 'g' ? 1 : 0) : 0; }();
        if (forced) sync_lanes = false;  // RK_B200_SYNC_KERNEL=generic: A/B and parity of the two kernels against each other
    }
    // position interface on every DoF: the lean set-up kernel (third-order items inline, the rest queued for k1_rescue)
    bool setup_lean = true;
    for (int d = 0; d < desc->dof; ++d) setup_lean = setup_lean && P.ci[d] == RK_POSITION;
    // host buffers: per-chunk staging, double buffered
    const rk::Params host_ptrs = P;
    const double* rk::Params::*dslots[11] = {&rk::Params::p0, &rk::Params::v0, &rk::Params::a0, &rk::Params::pf, &rk::Params::vf, &rk::Params::af,
                                             &rk::Params::v_max, &rk::Params::a_max, &rk::Params::j_max, &rk::Params::v_min, &rk::Params::a_min};
    if (host) {
        const size_t per = sizeof(double) * (size_t)desc->dof * (size_t)cn_max;
        size_t bytes = 0;
        for (int k = 0; k < (P.lim_per_dof ? 6 : 11); ++k) if (host_ptrs.*dslots[k]) bytes += per;
        if (host_ptrs.min_dur) bytes += sizeof(double) * (size_t)cn_max;
        if (host_ptrs.enabled) bytes += (((size_t)desc->dof * (size_t)cn_max + 255) / 256) * 256;
        if (bytes > h->chunk_in_bytes) {
            for (int k = 0; k < RK_STAGE_BUFS; ++k) {
                if (h->chunk_in[k]) cudaFree(h->chunk_in[k]);
                h->chunk_in[k] = nullptr;
            }
            h->chunk_in_bytes = 0;
            for (int k = 0; k < RK_STAGE_BUFS; ++k) RK_CUDA(cudaMalloc(&h->chunk_in[k], bytes));
            h->chunk_in_bytes = bytes;
        }
        if (P.lim_per_dof) {  // one value per DoF: copied once, ahead of the first chunk on the same stream
            if (!h->lim_dev) RK_CUDA(cudaMalloc(&h->lim_dev, sizeof(double) * 5 * RK_MAX_DOF));
            for (int k = 6; k < 11; ++k) {
                if (!(host_ptrs.*dslots[k])) continue;
                double* dst = h->lim_dev + (size_t)(k - 6) * RK_MAX_DOF;
                RK_CUDA(cudaMemcpyAsync(dst, host_ptrs.*dslots[k], sizeof(double) * (size_t)desc->dof, cudaMemcpyHostToDevice, h->copy_stream));
                P.*dslots[k] = dst;
            }
        }
    }

    // Host buffers: the kernels of the first chunk cannot start before its inputs have arrived, so the first chunk is cut
    // into RK_HOST_RAMP pieces — the pipeline fills after a fraction of a chunk's copy time instead of a whole one. Results
    // (status, duration, independent minimum durations) go back chunk by chunk on their own stream while later chunks
    // are still computed.
    const bool ramp = host && n_chunks >= 2 && cn_max >= (int64_t)RK_HOST_RAMP * 16384;
    const bool d2h_per_chunk = host && out && n_chunks >= 2;
    if (d2h_per_chunk) {  // earlier work on the caller-visible stream may still write the result arrays' device copies
        RK_CUDA(cudaEventRecord(h->ev_fork, h->stream));
        RK_CUDA(cudaStreamWaitEvent(h->d2h_stream, h->ev_fork, 0));
    }
    int chunk_index = 0;
    for (int64_t off = 0, step_n = 0; off < desc->n; off += step_n, ++chunk_index) {
        step_n = cn_max;
        if (ramp && off < cn_max) {
            const int64_t piece = (cn_max + RK_HOST_RAMP - 1) / RK_HOST_RAMP;
            step_n = (cn_max - off < piece) ? (cn_max - off) : piece;
        }
        P.off = off;
        P.cn = (desc->n - off < step_n) ? (desc->n - off) : step_n;
        const int lane_set = chunk_index % lanes;
        cudaStream_t cs = h->lane_stream[lane_set];
        const rk_handle::Scratch& c = h->sc[lane_set];
        P.ws = c.ws; P.rec = c.rec; P.vl = c.vl; P.wmeta = c.wmeta; P.wsrc = c.wsrc; P.wqmask = c.wqmask; P.rq_item = c.rq_item; P.rq_root = c.rq_root;
        P.list[0] = c.list[0]; P.list[1] = c.list[1]; P.pend = c.pend; P.counters = c.counters;
        if (host) {
            const int buf = chunk_index % RK_STAGE_BUFS;
            // the buffer may still be read by the chunk computed two iterations ago
            if (chunk_index >= RK_STAGE_BUFS) RK_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_computed[buf], 0));
            char* dev = (char*)h->chunk_in[buf];
            const size_t per = sizeof(double) * (size_t)desc->dof * (size_t)P.cn;
            for (int k = 0; k < (P.lim_per_dof ? 6 : 11); ++k) {
                const auto m = dslots[k];
                if (!(host_ptrs.*m)) continue;
                RK_CUDA(cudaMemcpy2DAsync(dev, sizeof(double) * P.cn, host_ptrs.*m + off, sizeof(double) * stride, sizeof(double) * P.cn,
                                          (size_t)desc->dof, cudaMemcpyHostToDevice, h->copy_stream));
                P.*m = (const double*)dev;
                dev += per;
            }
            if (host_ptrs.min_dur) {
                RK_CUDA(cudaMemcpyAsync(dev, host_ptrs.min_dur + off, sizeof(double) * P.cn, cudaMemcpyHostToDevice, h->copy_stream));
                P.min_dur = (const double*)dev;
                dev += sizeof(double) * (size_t)P.cn;
            }
            if (host_ptrs.enabled) {
                RK_CUDA(cudaMemcpy2DAsync(dev, (size_t)P.cn, host_ptrs.enabled + off, (size_t)stride, (size_t)P.cn, (size_t)desc->dof,
                                          cudaMemcpyHostToDevice, h->copy_stream));
                P.enabled = (const uint8_t*)dev;
            }
            RK_CUDA(cudaEventRecord(h->ev_copied[buf], h->copy_stream));
            RK_CUDA(cudaStreamWaitEvent(cs, h->ev_copied[buf], 0));
            P.in_n = P.cn;
            P.in_shift = -off;
        }
        const int64_t items = P.cn * P.dof;
        const unsigned grid_items = (unsigned)((items + 127) / 128);
        // list-driven stages: a few resident waves per SM, grid-stride over the (device-side) list length
#ifndef RK_LIST_WAVES
#define RK_LIST_WAVES 16
#endif
        unsigned grid_list = (unsigned)h->sm_count * (unsigned)RK_LIST_WAVES;
        if (grid_list > grid_items) grid_list = grid_items;
        const unsigned grid_vel = (grid_list * 128u + RK_LS_BLOCK - 1u) / RK_LS_BLOCK;
        if (poison_scratch()) {  // debugging aid (RK_B200_POISON=1): no kernel may depend on what an earlier call left behind
            const size_t it = (size_t)c.ws_items;
            RK_CUDA(cudaMemsetAsync(c.ws, 0xFF, sizeof(double) * rk::WS_SLOTS * it, cs));
            RK_CUDA(cudaMemsetAsync(c.vl, 0xFF, sizeof(double) * rk::VL_ITEM_DOUBLES * it, cs));
            RK_CUDA(cudaMemsetAsync(c.rec, 0xFF, sizeof(double2) * 6 * it, cs));
            RK_CUDA(cudaMemsetAsync(c.wmeta, 0xFF, sizeof(uint32_t) * rk::WM_WORDS * it, cs));
            RK_CUDA(cudaMemsetAsync(c.wsrc, 0xFF, it, cs));
            RK_CUDA(cudaMemsetAsync(c.wqmask, 0xFF, sizeof(uint32_t) * 10 * it, cs));
            RK_CUDA(cudaMemsetAsync(c.rq_item, 0xFF, sizeof(uint2) * 12 * it, cs));
            RK_CUDA(cudaMemsetAsync(c.rq_root, 0xFF, sizeof(double) * 8 * it, cs));
            RK_CUDA(cudaMemsetAsync(c.list[0], 0xFF, sizeof(uint32_t) * it, cs));
            RK_CUDA(cudaMemsetAsync(c.list[1], 0xFF, sizeof(uint32_t) * it, cs));
            RK_CUDA(cudaMemsetAsync(c.pend, 0xFF, sizeof(uint32_t) * it, cs));
        }
        RK_CUDA(cudaMemsetAsync(c.counters, 0, sizeof(uint32_t) * RK_N_COUNTERS, cs));
        static_assert(sizeof(rk_lazy::Params) == sizeof(rk::Params), "the two passes share one parameter block");
        P.self = reinterpret_cast<const char*>(c.counters) + 512;  // the set-up kernel deposits the parameter block there (exact redo paths)
        const rk_lazy::Params& LP = reinterpret_cast<const rk_lazy::Params&>(P);  // kernels of pass 2 (see the top of this file)
        if (setup_lean) rk_launch(rk::k1_setup_lean, grid_items, 128, cs, P);
        else rk_launch(rk::k1_setup, grid_items, 128, cs, P);
        const unsigned grid_fam = (unsigned)((2 * items + 127) / 128);
        rk_launch(rk::k1_quartic_roots<0>, grid_fam, 128, cs, P);
        rk_launch(rk::k1_quartic_check<0>, grid_list, 128, cs, P);
        rk_launch(rk::k1_quartic_roots<1>, grid_fam, 128, cs, P);
        rk_launch(rk::k1_quartic_check<1>, grid_list, 128, cs, P);
        rk_launch(rk::k1_quartic_roots<2>, grid_fam, 128, cs, P);
        rk_launch(rk::k1_quartic_check<2>, grid_list, 128, cs, P);
#if RK_CC_SPLIT
        rk_launch(rk::k1_closed_cands<0>, grid_fam, 128, cs, P);
        rk_launch(rk::k1_closed_cands<1>, grid_fam, 128, cs, P);
#else
        rk_launch(rk_lazy::k1_closed_cands<2>, grid_fam, 128, cs, LP);
#endif
        rk_launch(rk::k1_cand_check, grid_list, 128, cs, P);
        rk_launch(rk::k1_block, grid_items, 128, cs, P);
        rk_launch(rk::k1_rescue, (unsigned)h->sm_count * 4u, 128, cs, P);
        if (sync_lanes) rk_launch(rk::k_sync_lanes, (unsigned)((P.cn + 31) / 32), 256, cs, P);
        else rk_launch(rk::k_sync, (unsigned)((P.cn + 127) / 128), 128, cs, P);
        static const int fam[RK_S2_STEPS] = {0, 1, 8, 2, 3, 0, 1, 8, 2, 3, 4, 5, 6, 7, 4, 5, 6, 7};
        static const int dir2[RK_S2_STEPS] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1};
        // with RK_FUSE_TAIL the last eight stages (families 4-7 in both directions) and k2_fail are one launch (k2_tail)
        const int n_steps = RK_FUSE_TAIL ? 10 : RK_S2_STEPS;
        for (int step = 0; step < n_steps; ++step) {
            switch (fam[step]) {
                case 0: rk_launch(rk_lazy::k2_stage<0>, grid_list, 128, cs, LP, step, dir2[step]); break;
                case 1:
                    rk_launch(rk::k2_vel_scan_uddu, grid_vel, RK_LS_BLOCK, cs, P, step, dir2[step]);
                    rk_launch(rk::k2_polish<6, false>, (unsigned)h->sm_count * 8u, 128, cs, P, 33 + 4 * step, 34 + 4 * step);
                    rk_launch(rk::k2_vel_check_uddu, grid_list, 128, cs, P, step, dir2[step]);
                    h->launches += 2;
                    break;
                case 2: rk_launch(rk::k2_stage<2>, grid_list, 128, cs, P, step, dir2[step]); break;
                case 3: rk_launch(rk_lazy::k2_stage<3>, grid_list, 128, cs, LP, step, dir2[step]); break;
                case 4: rk_launch(rk::k2_stage<4>, grid_list, 128, cs, P, step, dir2[step]); break;
                case 5: rk_launch(rk::k2_stage<5>, grid_list, 128, cs, P, step, dir2[step]); break;
                case 6: rk_launch(rk::k2_stage<6>, grid_list, 128, cs, P, step, dir2[step]); break;
                case 7: rk_launch(rk::k2_stage<7>, grid_list, 128, cs, P, step, dir2[step]); break;
                default: rk_launch(rk::k2_vel<8>, grid_vel, RK_LS_BLOCK, cs, P, step, dir2[step]); break;
            }
        }
#if RK_FUSE_TAIL
        rk_launch(rk::k2_tail, grid_list, 128, cs, P, n_steps);
#else
        rk_launch(rk::k2_fail, (unsigned)h->sm_count, 128, cs, P);
#endif
        if (P.compact) rk_launch(rk::k_expand_compact, grid_items, 128, cs, P);
        else rk_launch(rk::k_expand, grid_items, 128, cs, P);
        h->launches += 14 + RK_CC_SPLIT + n_steps;  // fused tail: k2_tail stands in for k2_fail
        if (host) RK_CUDA(cudaEventRecord(h->ev_computed[chunk_index % RK_STAGE_BUFS], cs));
        if (d2h_per_chunk) {
            RK_CUDA(cudaStreamWaitEvent(h->d2h_stream, h->ev_computed[chunk_index % RK_STAGE_BUFS], 0));
            if (out->status)
                RK_CUDA(cudaMemcpyAsync(out->status + off, traj->status + off, sizeof(int32_t) * P.cn, cudaMemcpyDeviceToHost, h->d2h_stream));
            if (out->duration)
                RK_CUDA(cudaMemcpyAsync(out->duration + off, traj->duration + off, sizeof(double) * P.cn, cudaMemcpyDeviceToHost, h->d2h_stream));
            if (out->independent_min_durations)
                RK_CUDA(cudaMemcpy2DAsync(out->independent_min_durations + off, sizeof(double) * stride, traj->imd + off, sizeof(double) * desc->n,
                                          sizeof(double) * P.cn, (size_t)desc->dof, cudaMemcpyDeviceToHost, h->d2h_stream));
        }
    }
    if (d2h_per_chunk) {
        RK_CUDA(cudaEventRecord(h->ev_d2h, h->d2h_stream));
        RK_CUDA(cudaStreamWaitEvent(h->stream, h->ev_d2h, 0));
    }
    for (int k = 1; k < lanes; ++k) {
        RK_CUDA(cudaEventRecord(h->ev_join[k], h->lane_stream[k]));
        RK_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join[k], 0));
    }
    RK_CUDA(cudaGetLastError());

    if (out && !d2h_per_chunk) {
        const cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
        if (out->status) RK_CUDA(cudaMemcpyAsync(out->status, traj->status, sizeof(int32_t) * desc->n, kind, h->stream));
        if (out->duration) RK_CUDA(cudaMemcpyAsync(out->duration, traj->duration, sizeof(double) * desc->n, kind, h->stream));
        if (out->independent_min_durations)
            RK_CUDA(cudaMemcpy2DAsync(out->independent_min_durations, sizeof(double) * stride, traj->imd, sizeof(double) * desc->n,
                                      sizeof(double) * desc->n, (size_t)desc->dof, kind, h->stream));
    }
    if (host) RK_CUDA(cudaStreamSynchronize(h->stream));
    return RK_OK;
}

extern "C" {

rk_status rk_calculate_batch(rk_handle* h, const rk_batch_desc* desc, const rk_batch_in* in, rk_trajectories* traj, const rk_batch_out* out) {
    if (!desc) return fail(RK_ERR_INVALID_ARGUMENT, "desc / in is NULL");
    return calculate_impl(h, desc, in, traj, out, desc->n, 0);
}

void rk_shard_range(int64_t n, int32_t parts, int32_t part, int64_t* first, int64_t* count) {
    const int64_t base = n / parts, rem = n % parts;
    const int64_t lo = (int64_t)part * base + (part < rem ? part : rem);
    if (first) *first = lo;
    if (count) *count = base + (part < rem ? 1 : 0);
}

rk_status rk_calculate_batch_multi(rk_handle* const* handles, rk_trajectories* const* trajs, int32_t parts, const rk_batch_desc* desc,
                                   const rk_batch_in* in, const rk_batch_out* out) {
    if (!handles || !trajs || parts < 1 || !desc || !in) return fail(RK_ERR_INVALID_ARGUMENT, "NULL / empty argument to rk_calculate_batch_multi");
    if (desc->memory_space != RK_MEM_HOST) return fail(RK_ERR_INVALID_ARGUMENT, "rk_calculate_batch_multi takes host arrays (device memory belongs to one device)");
    for (int g = 0; g < parts; ++g) {
        if (!handles[g] || !trajs[g]) return fail(RK_ERR_INVALID_ARGUMENT, "a handle / trajectories entry is NULL");
        int64_t lo, cnt;
        rk_shard_range(desc->n, parts, g, &lo, &cnt);
        if (trajs[g]->n != cnt || trajs[g]->dof != desc->dof) return fail(RK_ERR_INVALID_ARGUMENT, "trajectories[g] must hold rk_shard_range(n, parts, g) trajectories");
        for (int k = 0; k < g; ++k)
            if (handles[k] == handles[g]) return fail(RK_ERR_INVALID_ARGUMENT, "every part needs its own handle (a handle is one stream + scratch)");
    }
    // one host thread per part: each part stages, computes and returns its own index range; parts on different devices (or on
    // different handles of one device) overlap completely, there is no exchange between them
    std::vector<rk_status> result((size_t)parts, RK_OK);
    std::vector<std::string> message((size_t)parts);
    std::vector<std::thread> workers;
    for (int g = 0; g < parts; ++g) {
        workers.emplace_back([&, g] {
            int64_t lo, cnt;
            rk_shard_range(desc->n, parts, g, &lo, &cnt);
            rk_batch_desc sub = *desc;
            sub.n = cnt;
            result[(size_t)g] = cnt > 0 ? calculate_impl(handles[g], &sub, in, trajs[g], out, desc->n, lo) : RK_OK;
            if (result[(size_t)g] != RK_OK) message[(size_t)g] = g_last_error;
        });
    }
    for (auto& w : workers) w.join();
    for (int g = 0; g < parts; ++g)
        if (result[(size_t)g] != RK_OK) return fail(result[(size_t)g], "part " + std::to_string(g) + ": " + message[(size_t)g]);
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
    S.rewind = !(desc->delta_time >= 0.0);
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
    // K is sliced over blockIdx.y only when (DoF x n) alone cannot fill the GPU: every slice re-reads the 45-double profile
    // and replays the time additions, which costs more than it hides once ~300 k threads exist (measured: 64 Ki x 7-DoF x
    // 1000 cycles runs at 3.1 TB/s unsliced, 2.4 TB/s in slices of 50).
    int64_t slices = (300000 + total - 1) / total;
    if (slices < 1) slices = 1;
    const bool pva_only = dev[0] && dev[1] && dev[2] && !dev[3] && !dsec && !S.rewind;
    const bool compact = traj->layout == RK_RESULT_COMPACT;
    void (*kern)(const rk::SampleParams) = compact ? (pva_only ? rk::k_at_time<true, true> : rk::k_at_time<false, true>)
                                                   : (pva_only ? rk::k_at_time<true, false> : rk::k_at_time<false, false>);
    {   // ... and a few slices more when that evens out the last wave of blocks (64 Ki x 7 DoF alone is four waves and a
        // sliver): pick the slice count whose last wave is fullest, charging 2 % per extra slice for the re-read (measured on
        // 64 Ki x 7 x 1000: 5.45 / 5.61 / 5.64 / 5.61 / 5.55 / 5.47 TB/s with 1 / 2 / 3 / 4 / 6 / 8 slices)
        static int resident[4] = {0, 0, 0, 0};  // blocks per SM of the four instantiations
        int& blocks_per_sm = resident[(compact ? 2 : 0) + (pva_only ? 1 : 0)];
        if (blocks_per_sm == 0) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, RK_AT_BLOCK, 0) != cudaSuccess || blocks_per_sm < 1) blocks_per_sm = 6;
        }
        const double wave = (double)h->sm_count * (double)blocks_per_sm * (double)RK_AT_BLOCK;
        double best = -1.0;
        int64_t best_s = slices;
        for (int64_t sl = slices; sl <= slices + 7; ++sl) {
            const double waves = (double)total * (double)sl / wave;
            const double eff = waves / (double)(int64_t)(waves + 0.999999) - 0.02 * (double)(sl - slices);
            if (eff > best + 1e-9) { best = eff; best_s = sl; }
        }
        slices = best_s;
    }
    {   // tuning aid: RK_B200_AT_SLICES overrides the heuristic
        static const int forced = [] { const char* e = std::getenv("RK_B200_AT_SLICES"); return e ? std::atoi(e) : 0; }();
        if (forced > 0) slices = forced;
    }
    if (slices > (desc->steps + 31) / 32) slices = (desc->steps + 31) / 32;
    if (slices > 65535) slices = 65535;  // gridDim.y
    if (slices < 1) slices = 1;
    S.k_chunk = (int32_t)((desc->steps + slices - 1) / slices);
    const dim3 grid((unsigned)((total + RK_AT_BLOCK - 1) / RK_AT_BLOCK), (unsigned)((desc->steps + S.k_chunk - 1) / S.k_chunk), 1);
    S.n_starts = 0;
    if (grid.y <= RK_AT_MAX_STARTS) {  // the repeated additions of ruckig.rs:293 up to each slice's first sample, done once here
        S.n_starts = (int32_t)grid.y;
        volatile double t = desc->time_start;  // volatile: plain IEEE double additions, one at a time
        for (int k = 0, y = 0; y < (int)grid.y; ++k) {
            if (k == y * S.k_chunk) S.start[y++] = t;
            t = t + desc->delta_time;
        }
    }
    kern<<<grid, RK_AT_BLOCK, 0, h->stream>>>(S);
    h->launches += 1;
    RK_CUDA(cudaGetLastError());
    if (host) {
        for (int k = 0; k < 4; ++k) if (dst[k]) RK_CUDA(cudaMemcpyAsync(dst[k], dev[k], per, cudaMemcpyDeviceToHost, h->stream));
        if (out->section) RK_CUDA(cudaMemcpyAsync(out->section, dsec, sec, cudaMemcpyDeviceToHost, h->stream));
        RK_CUDA(cudaStreamSynchronize(h->stream));
    }
    return RK_OK;
}

// ------------------------------------------------------------------------------------------------ rk_update_batch
struct rk_rollout {
    rk_handle* h = nullptr;
    int64_t n = 0;
    int32_t dof = 0;
    rk_trajectories* traj = nullptr;  // the environments' trajectories
    rk_trajectories* tmp = nullptr;   // results of the re-solved subset, compact [..][dof][m] inside buffers of capacity n
    double* last[11] = {};            // current_input, [dof][n] each
    double* gin[11] = {};             // gathered inputs of the subset
    uint8_t *last_enabled = nullptr, *g_enabled = nullptr;
    double *last_min_dur = nullptr, *g_min_dur = nullptr;
    double* time[2] = {nullptr, nullptr};
    int time_cur = 0;
    uint8_t *initialized = nullptr, *dirty = nullptr;
    uint32_t *idx = nullptr, *count = nullptr;
    uint32_t* count_host = nullptr;   // pinned
    // batch-level part of the last calculated input
    bool have_desc = false;
    rk_batch_desc last_desc{};
    int8_t last_ci[RK_MAX_DOF] = {}, last_sync[RK_MAX_DOF] = {};
    bool last_present[13] = {};
};

static rk_status rollout_alloc(rk_handle* h, rk_rollout* r) {
    const int64_t n = r->n;
    const size_t items = (size_t)n * r->dof;
    rk_status st = rk_trajectories_create(h, n, r->dof, &r->traj);
    if (st == RK_OK) st = rk_trajectories_create(h, n, r->dof, &r->tmp);
    if (st != RK_OK) return st;
    for (int f = 0; f < 11; ++f) {
        RK_CUDA(cudaMalloc(&r->last[f], sizeof(double) * items));
        RK_CUDA(cudaMalloc(&r->gin[f], sizeof(double) * items));
        RK_CUDA(cudaMemsetAsync(r->last[f], 0, sizeof(double) * items, h->stream));
    }
    RK_CUDA(cudaMalloc(&r->last_enabled, items)); RK_CUDA(cudaMalloc(&r->g_enabled, items));
    RK_CUDA(cudaMalloc(&r->last_min_dur, sizeof(double) * n)); RK_CUDA(cudaMalloc(&r->g_min_dur, sizeof(double) * n));
    RK_CUDA(cudaMalloc(&r->time[0], sizeof(double) * n)); RK_CUDA(cudaMalloc(&r->time[1], sizeof(double) * n));
    RK_CUDA(cudaMalloc(&r->initialized, n)); RK_CUDA(cudaMalloc(&r->dirty, n));
    RK_CUDA(cudaMalloc(&r->idx, sizeof(uint32_t) * n)); RK_CUDA(cudaMalloc(&r->count, sizeof(uint32_t)));
    RK_CUDA(cudaMallocHost(&r->count_host, sizeof(uint32_t)));
    return rk_rollout_reset(h, r);
}

rk_status rk_rollout_create(rk_handle* h, int64_t n, int32_t dof, rk_rollout** out) {
    if (!h || !out || n < 1 || dof < 1 || dof > RK_MAX_DOF) return fail(RK_ERR_INVALID_ARGUMENT, "bad arguments to rk_rollout_create");
    if (n > (int64_t)0xFFFFFFFFll) return fail(RK_ERR_INVALID_ARGUMENT, "rk_rollout_create: n exceeds 2^32 - 1 (environment indices are 32-bit)");
    RK_CUDA(cudaSetDevice(h->device));
    rk_rollout* r = new rk_rollout();
    r->h = h; r->n = n; r->dof = dof;
    const rk_status st = rollout_alloc(h, r);
    if (st != RK_OK) {  // nothing of a partially built rollout is leaked
        const std::string why = g_last_error;
        rk_rollout_destroy(r);
        cudaGetLastError();
        return fail(st, why);
    }
    *out = r;
    return RK_OK;
}

rk_status rk_rollout_reset(rk_handle* h, rk_rollout* r) {
    if (!h || !r) return fail(RK_ERR_INVALID_ARGUMENT, "NULL argument to rk_rollout_reset");
    RK_CUDA(cudaSetDevice(h->device));
    RK_CUDA(cudaMemsetAsync(r->initialized, 0, r->n, h->stream));
    RK_CUDA(cudaMemsetAsync(r->time[0], 0, sizeof(double) * r->n, h->stream));
    RK_CUDA(cudaMemsetAsync(r->time[1], 0, sizeof(double) * r->n, h->stream));
    r->time_cur = 0;
    r->have_desc = false;
    return RK_OK;
}

rk_status rk_rollout_destroy(rk_rollout* r) {
    if (!r) return RK_OK;
    cudaSetDevice(r->h->device);
    cudaStreamSynchronize(r->h->stream);
    rk_trajectories_destroy(r->traj); rk_trajectories_destroy(r->tmp);
    for (int f = 0; f < 11; ++f) { cudaFree(r->last[f]); cudaFree(r->gin[f]); }
    cudaFree(r->last_enabled); cudaFree(r->g_enabled); cudaFree(r->last_min_dur); cudaFree(r->g_min_dur);
    cudaFree(r->time[0]); cudaFree(r->time[1]); cudaFree(r->initialized); cudaFree(r->dirty); cudaFree(r->idx); cudaFree(r->count);
    cudaFreeHost(r->count_host);
    delete r;
    return RK_OK;
}

rk_trajectories* rk_rollout_trajectories(rk_rollout* r) { return r ? r->traj : nullptr; }

rk_status rk_update_batch(rk_handle* h, const rk_batch_desc* desc, const rk_batch_in* in, rk_rollout* r, const rk_update_out* out,
                          int64_t* recalculated) {
    if (!h || !r || !out) return fail(RK_ERR_INVALID_ARGUMENT, "NULL argument to rk_update_batch");
    rk::Params P;
    rk_status st = fill_params(P, desc, in);
    if (st != RK_OK) return st;
    if (desc->n != r->n || desc->dof != r->dof) return fail(RK_ERR_INVALID_ARGUMENT, "the rollout was created for another (n, dof)");
    if (desc->memory_space != RK_MEM_DEVICE) return fail(RK_ERR_INVALID_ARGUMENT, "rk_update_batch takes device pointers");
    RK_CUDA(cudaSetDevice(h->device));
    const int64_t n = r->n;
    const int dof = r->dof;

    // batch-level fields of InputParameter's PartialEq, and which optional arrays are present
    const bool present[13] = {true, true, true, true, true, true, true, true, true, in->min_velocity != nullptr, in->min_acceleration != nullptr,
                              in->enabled != nullptr, in->minimum_duration != nullptr};
    bool force_all = !r->have_desc || r->last_desc.duration_discretization != desc->duration_discretization;
    for (int d = 0; d < dof && !force_all; ++d) force_all = r->last_ci[d] != P.ci[d] || r->last_sync[d] != P.sync[d];
    for (int f = 0; f < 13 && !force_all; ++f) force_all = r->last_present[f] != present[f];

    rk::UpdateParams U{};
    U.n = n; U.dof = dof;
    const double* ins[11] = {in->current_position, in->current_velocity, in->current_acceleration, in->target_position, in->target_velocity,
                             in->target_acceleration, in->max_velocity, in->max_acceleration, in->max_jerk, in->min_velocity, in->min_acceleration};
    for (int f = 0; f < 11; ++f) { U.in[f] = ins[f]; U.last[f] = r->last[f]; U.gin[f] = r->gin[f]; }
    U.enabled = in->enabled; U.last_enabled = r->last_enabled; U.g_enabled = r->g_enabled;
    U.min_dur = in->minimum_duration; U.last_min_dur = r->last_min_dur; U.g_min_dur = r->g_min_dur;
    U.force_all = force_all ? 1 : 0;
    U.lim_per_dof = P.lim_per_dof;
    U.initialized = r->initialized; U.dirty = r->dirty; U.idx = r->idx; U.count = r->count;

    RK_CUDA(cudaMemsetAsync(r->count, 0, sizeof(uint32_t), h->stream));
    rk::k_upd_dirty<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(U);
    RK_CUDA(cudaMemcpyAsync(r->count_host, r->count, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    RK_CUDA(cudaStreamSynchronize(h->stream));
    const int64_t m = (int64_t)*r->count_host;
    h->launches += 1;
    if (recalculated) *recalculated = m;

    if (m > 0) {
        rk::k_upd_gather<<<(unsigned)((m * dof + 255) / 256), 256, 0, h->stream>>>(U, m);
        h->launches += 1;
        rk_batch_desc sub = *desc;
        sub.n = m;
        sub.limits_layout = RK_LIMITS_PER_TRAJECTORY;  // the gathered inputs are per trajectory
        rk_batch_in gi{};
        gi.current_position = r->gin[0]; gi.current_velocity = r->gin[1]; gi.current_acceleration = r->gin[2];
        gi.target_position = r->gin[3]; gi.target_velocity = r->gin[4]; gi.target_acceleration = r->gin[5];
        gi.max_velocity = r->gin[6]; gi.max_acceleration = r->gin[7]; gi.max_jerk = r->gin[8];
        gi.min_velocity = in->min_velocity ? r->gin[9] : nullptr;
        gi.min_acceleration = in->min_acceleration ? r->gin[10] : nullptr;
        gi.enabled = in->enabled ? r->g_enabled : nullptr;
        gi.minimum_duration = in->minimum_duration ? r->g_min_dur : nullptr;
        rk_trajectories view = *r->tmp;  // the same buffers, laid out for m trajectories
        view.n = m;
        st = rk_calculate_batch(h, &sub, &gi, &view, nullptr);
        if (st != RK_OK) return st;
        rk::ScatterParams C{};
        C.n = n; C.m = m; C.dof = dof; C.throw_mode = desc->error_mode == RK_ERRORS_THROW; C.idx = r->idx;
        C.slots = r->traj->slots();
        C.s_prof = view.prof; C.s_codes = view.codes; C.s_status = view.status; C.s_duration = view.duration; C.s_imd = view.imd;
        C.s_limiting = view.limiting; C.s_detail = view.detail;
        C.prof = r->traj->prof; C.codes = r->traj->codes; C.status = r->traj->status; C.duration = r->traj->duration; C.imd = r->traj->imd;
        C.limiting = r->traj->limiting; C.detail = r->traj->detail;
        rk::k_upd_scatter<<<(unsigned)((m * dof + 255) / 256), 256, 0, h->stream>>>(C, U);
        h->launches += 1;
    }
    r->have_desc = true;
    r->last_desc = *desc;
    for (int d = 0; d < dof; ++d) { r->last_ci[d] = P.ci[d]; r->last_sync[d] = P.sync[d]; }
    for (int f = 0; f < 13; ++f) r->last_present[f] = present[f];

    rk::StepParams T{};
    T.n = n; T.dof = dof; T.throw_mode = desc->error_mode == RK_ERRORS_THROW; T.reset_time = out->reset_time_on_new_calculation;
    T.delta_time = desc->delta_time;
    T.compact = r->traj->layout == RK_RESULT_COMPACT;
    T.prof = r->traj->prof; T.duration = r->traj->duration; T.status = r->traj->status; T.dirty = r->dirty;
    T.time_in = r->time[r->time_cur]; T.time_out = r->time[r->time_cur ^ 1];
    T.last_p = r->last[0]; T.last_v = r->last[1]; T.last_a = r->last[2];
    if (out->pass_to_input) {
        T.in_p = const_cast<double*>(in->current_position); T.in_v = const_cast<double*>(in->current_velocity);
        T.in_a = const_cast<double*>(in->current_acceleration);
    }
    T.new_p = out->new_position; T.new_v = out->new_velocity; T.new_a = out->new_acceleration; T.new_j = out->new_jerk;
    T.result = out->result; T.new_calc = out->new_calculation; T.time_user = out->time;
    if (T.compact) rk::k_upd_step<true><<<(unsigned)((n * dof + 255) / 256), 256, 0, h->stream>>>(T);
    else rk::k_upd_step<false><<<(unsigned)((n * dof + 255) / 256), 256, 0, h->stream>>>(T);
    h->launches += 1;
    r->time_cur ^= 1;
    RK_CUDA(cudaGetLastError());
    return RK_OK;
}

// ------------------------------------------------------------------------------------------------ trajectory query helpers (host)
rk_status rk_position_extrema_batch(rk_handle* h, const rk_trajectories* traj, double* p_min, double* p_max, double* t_min, double* t_max,
                                    int32_t memory_space) {
    if (!h || !traj) return fail(RK_ERR_INVALID_ARGUMENT, "NULL argument to rk_position_extrema_batch");
    if (traj->n == 0) return RK_OK;
    RK_CUDA(cudaSetDevice(h->device));
    const bool host = memory_space == RK_MEM_HOST;
    const size_t per = sizeof(double) * (size_t)traj->n * (size_t)traj->dof;
    double* dst[4] = {p_min, p_max, t_min, t_max};
    double* dev[4] = {p_min, p_max, t_min, t_max};
    if (host) {
        rk_status st = ensure_stage(h, 4 * per, 0);
        if (st != RK_OK) return st;
        for (int k = 0; k < 4; ++k) dev[k] = dst[k] ? (double*)((char*)h->stage_dev + k * per) : nullptr;
    }
    rk::QueryParams Q{};
    Q.n = traj->n; Q.dof = traj->dof; Q.prof = traj->prof; Q.compact = traj->layout == RK_RESULT_COMPACT;
    Q.e_min = dev[0]; Q.e_max = dev[1]; Q.e_tmin = dev[2]; Q.e_tmax = dev[3];
    rk::k_position_extrema<<<(unsigned)(((int64_t)traj->n * traj->dof + 255) / 256), 256, 0, h->stream>>>(Q);
    h->launches += 1;
    RK_CUDA(cudaGetLastError());
    if (host) {
        for (int k = 0; k < 4; ++k) if (dst[k]) RK_CUDA(cudaMemcpyAsync(dst[k], dev[k], per, cudaMemcpyDeviceToHost, h->stream));
        RK_CUDA(cudaStreamSynchronize(h->stream));
    }
    return RK_OK;
}

rk_status rk_first_time_at_position_batch(rk_handle* h, const rk_trajectories* traj, const double* position, double* time, int32_t memory_space) {
    if (!h || !traj || !position || !time) return fail(RK_ERR_INVALID_ARGUMENT, "NULL argument to rk_first_time_at_position_batch");
    if (traj->n == 0) return RK_OK;
    RK_CUDA(cudaSetDevice(h->device));
    const bool host = memory_space == RK_MEM_HOST;
    const size_t per = sizeof(double) * (size_t)traj->n * (size_t)traj->dof;
    const double* d_pos = position;
    double* d_time = time;
    if (host) {
        rk_status st = ensure_stage(h, 2 * per, 0);
        if (st != RK_OK) return st;
        RK_CUDA(cudaMemcpyAsync(h->stage_dev, position, per, cudaMemcpyHostToDevice, h->stream));
        d_pos = (const double*)h->stage_dev;
        d_time = (double*)((char*)h->stage_dev + per);
    }
    rk::QueryParams Q{};
    Q.n = traj->n; Q.dof = traj->dof; Q.prof = traj->prof; Q.position = d_pos; Q.time = d_time;
    Q.compact = traj->layout == RK_RESULT_COMPACT;
    rk::k_first_time_at_position<<<(unsigned)(((int64_t)traj->n * traj->dof + 255) / 256), 256, 0, h->stream>>>(Q);
    h->launches += 1;
    RK_CUDA(cudaGetLastError());
    if (host) {
        RK_CUDA(cudaMemcpyAsync(time, d_time, per, cudaMemcpyDeviceToHost, h->stream));
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
    int ex_first = -1, ex_count = 0;  // compact layout: the field is expanded by a kernel instead of copied
    auto slots = [&](int first, int count) {
        src = t->prof + (size_t)first * items; bytes = sizeof(double) * count * items;
        if (t->layout == RK_RESULT_COMPACT) { ex_first = first; ex_count = count; }
    };
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
    if (bytes == 0 || t->n == 0) return RK_OK;
    if (ex_first >= 0) {
        double* dev = (double*)dst;
        if (dst_memory_space == RK_MEM_HOST) {
            rk_status st = ensure_stage(h, bytes, 0);
            if (st != RK_OK) return st;
            dev = (double*)h->stage_dev;
        }
        rk::k_read_expanded<<<(unsigned)((items + 255) / 256), 256, 0, h->stream>>>(t->prof, (int64_t)items, ex_first, ex_count, dev);
        h->launches += 1;
        RK_CUDA(cudaGetLastError());
        if (dst_memory_space != RK_MEM_HOST) return RK_OK;
        src = dev;
    }
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

rk_status rk_selftest_division(rk_handle* h, int64_t n, uint64_t seed, int64_t* mismatches) {
    if (!h || !mismatches || n < 0) return fail(RK_ERR_INVALID_ARGUMENT, "bad argument to rk_selftest_division");
    RK_CUDA(cudaSetDevice(h->device));
    rk_status st = ensure_stage(h, 256, 0);
    if (st != RK_OK) return st;
    RK_CUDA(cudaMemsetAsync(h->stage_dev, 0, 8, h->stream));
    rk::k_selftest_division<<<h->sm_count * 8, 256, 0, h->stream>>>(n, seed, (unsigned long long*)h->stage_dev);
    h->launches += 1;
    RK_CUDA(cudaGetLastError());
    unsigned long long bad = 0;
    RK_CUDA(cudaMemcpyAsync(&bad, h->stage_dev, 8, cudaMemcpyDeviceToHost, h->stream));
    RK_CUDA(cudaStreamSynchronize(h->stream));
    *mismatches = (int64_t)bad;
    return RK_OK;
}

rk_status rk_debug_counters(rk_handle* h, uint32_t* out) {
    if (!h || !out) return fail(RK_ERR_INVALID_ARGUMENT, "NULL argument to rk_debug_counters");
    if (!h->sc[0].counters) return fail(RK_ERR_INVALID_ARGUMENT, "no calculation has run on this handle yet");
    RK_CUDA(cudaSetDevice(h->device));
    RK_CUDA(cudaStreamSynchronize(h->stream));
    RK_CUDA(cudaMemcpy(out, h->sc[0].counters, sizeof(uint32_t) * RK_N_COUNTERS, cudaMemcpyDeviceToHost));
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

