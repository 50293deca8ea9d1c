/* ruckig_b200.h — C ABI of libruckig_b200.so (B200 / sm_100a batched Ruckig).
 *
 * The reference (petrikosk/rsruckig v2.1.2) has no FFI layer: callers use the
 * Rust API directly. Each entry point below names the reference interface it
 * replaces for a whole batch of independent problems (file:line under
 * /root/reference/lib/src/rsruckig/). A Rust `-sys` crate binds exactly these
 * symbols.
 *
 * Conventions
 *  - every function returns rk_status (0 = RK_OK, negative = argument / CUDA
 *    error); nothing throws across the ABI; rk_last_error() gives the text.
 *  - all arrays are structure-of-arrays, trajectory index fastest:
 *    element (dof d, trajectory i) of a per-DoF field is ptr[d * n + i];
 *    element (slot s, dof d, trajectory i) of a profile field is
 *    ptr[(s * dof + d) * n + i].
 *  - all pointers are caller-owned. `memory_space` says whether the data
 *    pointers are host or device pointers; the small per-DoF option arrays in
 *    rk_batch_desc are always host pointers.
 *  - calls on one handle are ordered on the handle's CUDA stream and are not
 *    thread-safe; one handle per device (multi-GPU = one handle per GPU, no
 *    exchange between them).
 *  - there is no CPU fallback: without a usable CUDA device rk_create fails.
 */
#ifndef RUCKIG_B200_H
#define RUCKIG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t rk_status;
enum {
    RK_OK = 0,
    RK_ERR_INVALID_ARGUMENT = -1,
    RK_ERR_CUDA = -2,
    RK_ERR_NO_DEVICE = -3,
    RK_ERR_OUT_OF_MEMORY = -4,
    RK_ERR_UNSUPPORTED = -5
};

/* input_parameter.rs:35-45 (ControlInterface), :72-85 (Synchronization),
 * :110-117 (DurationDiscretization) — discriminant order of the reference. */
enum { RK_POSITION = 0, RK_VELOCITY = 1, RK_ACCELERATION = 2 };
enum { RK_SYNC_TIME = 0, RK_SYNC_TIME_IF_NECESSARY = 1, RK_SYNC_PHASE = 2, RK_SYNC_NONE = 3 };
enum { RK_CONTINUOUS = 0, RK_DISCRETE = 1 };

/* result.rs:34-61 (RuckigResult) — per-trajectory result codes. Under
 * RK_ERRORS_THROW a trajectory for which the reference would return
 * Err(ValidationError) gets RK_RESULT_ERROR_INVALID_INPUT and is not
 * calculated; Err(CalculatorError) maps to the code the IgnoreErrorHandler
 * path returns (-104 / -110 / -111). */
enum {
    RK_RESULT_WORKING = 0,
    RK_RESULT_FINISHED = 1,
    RK_RESULT_ERROR = -1,
    RK_RESULT_ERROR_INVALID_INPUT = -100,
    RK_RESULT_ERROR_TRAJECTORY_DURATION = -101,
    RK_RESULT_ERROR_POSITIONAL_LIMITS = -102,
    RK_RESULT_ERROR_ZERO_LIMITS = -104,
    RK_RESULT_ERROR_EXECUTION_TIME_CALCULATION = -110,
    RK_RESULT_ERROR_SYNCHRONIZATION_CALCULATION = -111
};

/* error.rs:98-122: ThrowErrorHandler / IgnoreErrorHandler. */
enum { RK_ERRORS_THROW = 0, RK_ERRORS_IGNORE = 1 };
enum { RK_MEM_HOST = 0, RK_MEM_DEVICE = 1 };
/* Layout of the five limit arrays of rk_batch_in (max_velocity, max_acceleration, max_jerk, min_velocity, min_acceleration):
 * one value per (DoF, trajectory) like every other field of InputParameter, or one value per DoF shared by the whole batch — a
 * fleet of identical robots — in which case the arrays are [dof] and a host-buffer call copies a third fewer bytes
 * (336 instead of 504 per 7-DoF trajectory). Results are the same bits either way. */
enum { RK_LIMITS_PER_TRAJECTORY = 0, RK_LIMITS_PER_DOF = 1 };

#define RK_MAX_DOF 64

/* profile.rs:25-49 */
enum { RK_LIMITS_ACC0_ACC1_VEL = 0, RK_LIMITS_VEL, RK_LIMITS_ACC0, RK_LIMITS_ACC1, RK_LIMITS_ACC0_ACC1,
       RK_LIMITS_ACC0_VEL, RK_LIMITS_ACC1_VEL, RK_LIMITS_NONE };
enum { RK_DIRECTION_UP = 0, RK_DIRECTION_DOWN = 1 };
enum { RK_SIGNS_UDDU = 0, RK_SIGNS_UDUD = 1 };

typedef struct rk_handle rk_handle;             /* replaces Ruckig<DOF,E> (ruckig.rs:76-83): stream + scratch */
typedef struct rk_trajectories rk_trajectories; /* replaces Vec<Trajectory<DOF>> (trajectory.rs:14-21), device resident */

/* Ruckig::new(dofs, delta_time) + the option fields of InputParameter that are
 * shared by the whole batch (ruckig.rs:110-119, input_parameter.rs:147-188). */
typedef struct rk_batch_desc {
    int64_t n;                                /* trajectories in the batch */
    int32_t dof;                              /* 1 .. RK_MAX_DOF */
    int32_t control_interface;                /* RK_POSITION | RK_VELOCITY | RK_ACCELERATION */
    int32_t synchronization;                  /* RK_SYNC_* */
    int32_t duration_discretization;          /* RK_CONTINUOUS | RK_DISCRETE */
    int32_t error_mode;                       /* RK_ERRORS_THROW | RK_ERRORS_IGNORE */
    int32_t memory_space;                     /* RK_MEM_HOST | RK_MEM_DEVICE for rk_batch_in / rk_batch_out */
    int32_t limits_layout;                    /* RK_LIMITS_PER_TRAJECTORY ([dof][n], the default) | RK_LIMITS_PER_DOF ([dof]) */
    double delta_time;                        /* control cycle (Ruckig::delta_time) */
    const int32_t* per_dof_control_interface; /* [dof] host, or NULL */
    const int32_t* per_dof_synchronization;   /* [dof] host, or NULL */
} rk_batch_desc;

/* The per-trajectory fields of InputParameter<DOF> (input_parameter.rs:147-188). */
typedef struct rk_batch_in {
    const double* current_position;     /* [dof][n] */
    const double* current_velocity;
    const double* current_acceleration;
    const double* target_position;
    const double* target_velocity;
    const double* target_acceleration;
    const double* max_velocity;         /* the five limit arrays: [dof][n], or [dof] with RK_LIMITS_PER_DOF */
    const double* max_acceleration;     /* +inf = second-order (no jerk limit needs max_jerk = +inf too) */
    const double* max_jerk;
    const double* min_velocity;         /* or NULL (= -max_velocity) */
    const double* min_acceleration;     /* or NULL (= -max_acceleration) */
    const uint8_t* enabled;             /* [dof][n] or NULL (= all enabled) */
    const double* minimum_duration;     /* [n] or NULL; NaN entry = None */
} rk_batch_in;

/* Results copied out right after the calculation; every pointer may be NULL. */
typedef struct rk_batch_out {
    int32_t* status;                    /* [n] RuckigResult value */
    double* duration;                   /* [n] Trajectory::duration */
    double* independent_min_durations;  /* [dof][n] */
} rk_batch_out;

/* Fields of the device-resident trajectories readable with rk_trajectories_read. */
enum {
    RK_FIELD_STATUS = 0,                /* int32 [n] */
    RK_FIELD_DURATION,                  /* f64 [n] */
    RK_FIELD_INDEPENDENT_MIN_DURATIONS, /* f64 [dof][n] */
    RK_FIELD_LIMITING_DOF,              /* int32 [n], -1 = none */
    RK_FIELD_CODES,                     /* uint32 [dof][n]: limits | control_signs<<8 | direction<<16 | source<<24 */
    RK_FIELD_T,                         /* f64 [7][dof][n]  Profile::t      (profile.rs:76-118) */
    RK_FIELD_T_SUM,                     /* f64 [7][dof][n]  Profile::t_sum */
    RK_FIELD_J,                         /* f64 [7][dof][n]  Profile::j */
    RK_FIELD_A,                         /* f64 [8][dof][n]  Profile::a */
    RK_FIELD_V,                         /* f64 [8][dof][n]  Profile::v */
    RK_FIELD_P,                         /* f64 [8][dof][n]  Profile::p */
    RK_FIELD_BRAKE_DURATION,            /* f64 [1][dof][n]  BrakeProfile::duration (brake.rs:25-32) */
    RK_FIELD_BRAKE_T,                   /* f64 [2][dof][n] */
    RK_FIELD_BRAKE_J,                   /* f64 [2][dof][n] */
    RK_FIELD_BRAKE_A,                   /* f64 [2][dof][n] */
    RK_FIELD_BRAKE_V,                   /* f64 [2][dof][n] */
    RK_FIELD_BRAKE_P,                   /* f64 [2][dof][n] */
    RK_FIELD_PF_VF_AF,                  /* f64 [3][dof][n]  Profile::pf, vf, af */
    RK_FIELD_ERROR_DETAIL,              /* int32 [n]: for error statuses (stage << 8) | dof; stage 1 = Step 1,
                                           2 = synchronisation, 3 = Step 2 (the texts of calculator_target.rs:459-470,
                                           :509-517, :820-825 are rebuilt from it on the host); 0 otherwise */
    RK_FIELD_COUNT
};
/* `source` byte of RK_FIELD_CODES: which stage produced the final profile. */
enum { RK_SRC_STEP1_MIN = 0, RK_SRC_STEP1_A = 1, RK_SRC_STEP1_B = 2, RK_SRC_PHASE = 3, RK_SRC_STEP2 = 4, RK_SRC_NONE = 5 };

/* Trajectory::at_time for K sample times per trajectory (trajectory.rs:158-189).
 * Sample k is taken at time_k = the value after k+1 repeated additions of
 * delta_time to time_start, which is what K calls of Ruckig::update produce
 * (ruckig.rs:293). Outputs are [K][dof][n]; each may be NULL. */
typedef struct rk_sample_desc {
    double time_start;
    double delta_time;
    int32_t steps;       /* K */
    int32_t memory_space;
} rk_sample_desc;

typedef struct rk_sample_out {
    double* position;
    double* velocity;
    double* acceleration;
    double* jerk;
    int32_t* section;    /* [K][n] new_section as at_time reports it, or NULL */
} rk_sample_out;

/* ---- lifetime --------------------------------------------------------- */
rk_status rk_create(int32_t device, rk_handle** out);                 /* Ruckig::new, ruckig.rs:110 */
rk_status rk_destroy(rk_handle* h);
rk_status rk_trajectories_create(rk_handle* h, int64_t n, int32_t dof, rk_trajectories** out); /* Trajectory::new, trajectory.rs:37-50; RK_RESULT_FULL */
/* Layout of the device-resident result (SURVEY.md §8b "compact and/or full, selectable"):
 *   RK_RESULT_FULL     the reference's Profile verbatim (profile.rs:76-118): 59 doubles per (DoF, trajectory), 3.3 KB per 7-DoF trajectory;
 *   RK_RESULT_COMPACT  what the calculation decided — t[7], the control values, the brake pre-trajectory, the boundary states:
 *                      20 doubles per (DoF, trajectory), 1.1 KB per 7-DoF trajectory. Every consumer (rk_at_time_batch, the
 *                      query helpers, rk_trajectories_read of any field) integrates the remaining Profile fields on the fly
 *                      with the operations the full layout was filled with: results are bit-identical in both layouts. */
enum { RK_RESULT_FULL = 0, RK_RESULT_COMPACT = 1 };
rk_status rk_trajectories_create_ex(rk_handle* h, int64_t n, int32_t dof, int32_t result_layout, rk_trajectories** out);
rk_status rk_trajectories_destroy(rk_trajectories* t);

/* ---- the hot path ------------------------------------------------------ */
/* Ruckig::calculate for every trajectory of the batch (ruckig.rs:210-218 ->
 * calculator_target.rs:278-833). Asynchronous on the handle's stream when
 * memory_space is RK_MEM_DEVICE (batches of 192 Ki (DoF, trajectory) items or
 * more are internally spread over a second stream that joins the handle's
 * stream again before the call returns);
 * with RK_MEM_HOST the inputs are staged chunk by chunk through device buffers
 * (pin the host arrays for full PCIe speed) and the call returns after the
 * results reached `out`. */
rk_status rk_calculate_batch(rk_handle* h, const rk_batch_desc* desc, const rk_batch_in* in,
                             rk_trajectories* traj, const rk_batch_out* out);

/* The same for ONE host batch spread over several handles — normally one per GPU of the box (north_star: "sharded trivially
 * across the 8 GPUs of one box with no collectives beyond a final host gather"), so a Rust / C++ host uses every device
 * without one process per GPU. Part g computes the contiguous index range rk_shard_range(n, parts, g) on handles[g] into
 * trajs[g] (created for exactly that many trajectories); the parts run concurrently on one host thread each and write their
 * slices of `out` (the gather). Host arrays only. Every part needs its own handle; two handles may share a device. */
void rk_shard_range(int64_t n, int32_t parts, int32_t part, int64_t* first, int64_t* count);
rk_status rk_calculate_batch_multi(rk_handle* const* handles, rk_trajectories* const* trajs, int32_t parts,
                                   const rk_batch_desc* desc, const rk_batch_in* in, const rk_batch_out* out);

/* Ruckig::validate_input(input, check_current, check_target) (ruckig.rs:150-171,
 * input_parameter.rs:251-420): valid[i] = 1 if trajectory i passes. reason[i] (may be NULL) = 0 if valid, else
 * (check_id << 8) | dof of the FIRST failing check in the reference's order; check ids are RK_INVALID_* below. */
rk_status rk_validate_batch(rk_handle* h, const rk_batch_desc* desc, const rk_batch_in* in,
                            int32_t check_current_state_within_limits, int32_t check_target_state_within_limits,
                            uint8_t* valid, int32_t* reason);
enum {
    RK_INVALID_MAX_JERK = 1, RK_INVALID_MAX_ACCELERATION, RK_INVALID_MIN_ACCELERATION, RK_INVALID_CURRENT_ACCELERATION_NAN,
    RK_INVALID_TARGET_ACCELERATION_NAN, RK_INVALID_CURRENT_ACCELERATION_ABOVE, RK_INVALID_CURRENT_ACCELERATION_BELOW,
    RK_INVALID_TARGET_ACCELERATION_ABOVE, RK_INVALID_TARGET_ACCELERATION_BELOW, RK_INVALID_CURRENT_VELOCITY_NAN,
    RK_INVALID_TARGET_VELOCITY_NAN, RK_INVALID_CURRENT_POSITION_NAN, RK_INVALID_TARGET_POSITION_NAN, RK_INVALID_MAX_VELOCITY,
    RK_INVALID_MIN_VELOCITY, RK_INVALID_CURRENT_VELOCITY_ABOVE, RK_INVALID_CURRENT_VELOCITY_BELOW, RK_INVALID_TARGET_VELOCITY_ABOVE,
    RK_INVALID_TARGET_VELOCITY_BELOW, RK_INVALID_CURRENT_WILL_EXCEED, RK_INVALID_CURRENT_WILL_UNDERCUT, RK_INVALID_TARGET_WILL_EXCEED,
    RK_INVALID_TARGET_WILL_UNDERCUT, RK_INVALID_DELTA_TIME
};

/* Trajectory::at_time for every trajectory and K sample times (trajectory.rs:158-189). */
rk_status rk_at_time_batch(rk_handle* h, const rk_trajectories* traj, const rk_sample_desc* desc, const rk_sample_out* out);

/* Trajectory::get_position_extrema (trajectory.rs:211-231, profile.rs:754-863): per (DoF, trajectory) the smallest and largest
 * position of the whole motion and the times at which they are reached. [dof][n] arrays, each nullable. */
rk_status rk_position_extrema_batch(rk_handle* h, const rk_trajectories* traj, double* p_min, double* p_max, double* t_min, double* t_max,
                                    int32_t memory_space);

/* Trajectory::get_first_time_at_position (trajectory.rs:233-246, profile.rs:865-895) for one query position per (DoF,
 * trajectory): position[dof][n] -> time[dof][n], NaN where the reference returns None. */
rk_status rk_first_time_at_position_batch(rk_handle* h, const rk_trajectories* traj, const double* position, double* time,
                                          int32_t memory_space);

/* ---- closed-loop control cycles: the batched Ruckig::update (ruckig.rs:266-321) ------------------------------------------
 * An rk_rollout is n independent (Ruckig, OutputParameter) pairs: per environment the input of its last calculation
 * (Ruckig::current_input), its trajectory and OutputParameter::time. One rk_update_batch call is one control cycle for all
 * of them: environments that were never calculated or whose input differs from their current_input (the PartialEq of
 * input_parameter.rs:190-211, batch-level settings included) are re-solved — only those —, then every environment advances
 * its time by desc->delta_time and samples its trajectory (Trajectory::at_time), and the new state is written into
 * current_input (output.pass_to_input(&mut self.current_input), ruckig.rs:314) and, if asked, into the caller's
 * current_* arrays. As in the reference, time is NOT reset by a new calculation unless
 * reset_time_on_new_calculation is set, and Ok(Error*) results of the calculation are dropped (Q1); under RK_ERRORS_THROW a
 * validation / calculator error is returned as the environment's result and its cycle is not advanced (the `?` of :285).
 * Device pointers only (desc->memory_space must be RK_MEM_DEVICE); stream-ordered on rk_stream(h) except for one small
 * device-to-host read per cycle (the number of environments to re-solve). */
typedef struct rk_rollout rk_rollout;

typedef struct rk_update_out {
    double* new_position;       /* [dof][n], nullable */
    double* new_velocity;       /* [dof][n], nullable */
    double* new_acceleration;   /* [dof][n], nullable */
    double* new_jerk;           /* [dof][n], nullable */
    int32_t* result;            /* [n], nullable: 0 Working, 1 Finished, < 0 error of the calculation (throw semantics) */
    uint8_t* new_calculation;   /* [n], nullable: OutputParameter::new_calculation */
    double* time;               /* [n], nullable: OutputParameter::time after the cycle */
    int32_t pass_to_input;      /* != 0: also write the new state into in->current_* (OutputParameter::pass_to_input) */
    int32_t reset_time_on_new_calculation; /* 0 = the reference's behaviour (Q1) */
} rk_update_out;

rk_status rk_rollout_create(rk_handle* h, int64_t n, int32_t dof, rk_rollout** out);
rk_status rk_rollout_destroy(rk_rollout* r);
rk_status rk_rollout_reset(rk_handle* h, rk_rollout* r);          /* Ruckig::reset (ruckig.rs:125-127) + a fresh OutputParameter */
rk_trajectories* rk_rollout_trajectories(rk_rollout* r);          /* the trajectories of the n environments (owned by r) */
rk_status rk_update_batch(rk_handle* h, const rk_batch_desc* desc, const rk_batch_in* in, rk_rollout* r, const rk_update_out* out,
                          int64_t* recalculated /* nullable: number of environments re-solved in this cycle */);

/* Copy one field of the device-resident trajectories to `dst`. */
rk_status rk_trajectories_read(rk_handle* h, const rk_trajectories* traj, int32_t field, void* dst, int32_t dst_memory_space);

/* ---- plumbing ---------------------------------------------------------- */
rk_status rk_sync(rk_handle* h);                      /* wait for the handle's stream */
void* rk_stream(rk_handle* h);                        /* the cudaStream_t, for event timing by the caller */
int64_t rk_kernel_launches(const rk_handle* h);       /* kernels launched through this handle so far */
/* Diagnostics: the 112 work-list counters the last chunk computed on scratch set 0 left behind (how many (DoF, trajectory)
 * items entered each stage of the Step-2 chain, queue lengths of the Step-1 kernels; layout in rk_params.cuh). */
rk_status rk_debug_counters(rk_handle* h, uint32_t* out /* [112] host */);
/* Roofline denominator for this path (it is FP64-pipe bound): runs a register-resident chain of
 * dependent-free DFMAs on every SM for about `milliseconds` and reports the sustained rate, an FMA counted as 2 flop. */
rk_status rk_measure_fp64_peak(rk_handle* h, double milliseconds, double* tflops);
/* Self-test of the library's own fp64 division paths (shared-reciprocal division, zero-safe division; rk_math.cuh)
 * against the compiler's IEEE division on n pseudo-random operand pairs that include zeros, infinities, NaNs,
 * subnormals and extreme exponents; *mismatches = number of pairs whose result bits differ (NaN == NaN). */
rk_status rk_selftest_division(rk_handle* h, int64_t n, uint64_t seed, int64_t* mismatches);
const char* rk_status_string(rk_status s);
const char* rk_last_error(void);
const char* rk_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RUCKIG_B200_H */
