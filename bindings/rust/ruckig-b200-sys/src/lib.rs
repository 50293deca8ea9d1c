//! Raw bindings to `libruckig_b200.so`, one to one with `include/ruckig_b200.h`.
//! SOURCE ONLY: never compiled where this repository is built (no Rust toolchain there), see ../README.md.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub type rk_status = i32;

pub const RK_OK: rk_status = 0;
pub const RK_POSITION: i32 = 0;
pub const RK_VELOCITY: i32 = 1;
pub const RK_ACCELERATION: i32 = 2;
pub const RK_SYNC_TIME: i32 = 0;
pub const RK_SYNC_TIME_IF_NECESSARY: i32 = 1;
pub const RK_SYNC_PHASE: i32 = 2;
pub const RK_SYNC_NONE: i32 = 3;
pub const RK_CONTINUOUS: i32 = 0;
pub const RK_DISCRETE: i32 = 1;
pub const RK_ERRORS_THROW: i32 = 0;
pub const RK_ERRORS_IGNORE: i32 = 1;
pub const RK_MEM_HOST: i32 = 0;
pub const RK_MEM_DEVICE: i32 = 1;
pub const RK_LIMITS_PER_TRAJECTORY: i32 = 0;
pub const RK_LIMITS_PER_DOF: i32 = 1;  // max_* / min_* arrays are [dof], one value per DoF for the whole batch
pub const RK_RESULT_FULL: i32 = 0;     // profile.rs:76-118 verbatim, 59 doubles per (DoF, trajectory)
pub const RK_RESULT_COMPACT: i32 = 1;  // 20 doubles per (DoF, trajectory), expanded on the fly by every reader
pub const RK_MAX_DOF: usize = 64;
// fields of rk_trajectories_read (enum in the header, same order)
pub const RK_FIELD_STATUS: i32 = 0;
pub const RK_FIELD_DURATION: i32 = 1;
pub const RK_FIELD_INDEPENDENT_MIN_DURATIONS: i32 = 2;
pub const RK_FIELD_LIMITING_DOF: i32 = 3;
pub const RK_FIELD_CODES: i32 = 4;
pub const RK_FIELD_T: i32 = 5;
pub const RK_FIELD_T_SUM: i32 = 6;
pub const RK_FIELD_J: i32 = 7;
pub const RK_FIELD_A: i32 = 8;
pub const RK_FIELD_V: i32 = 9;
pub const RK_FIELD_P: i32 = 10;
pub const RK_FIELD_BRAKE_DURATION: i32 = 11;
pub const RK_FIELD_BRAKE_T: i32 = 12;
pub const RK_FIELD_BRAKE_J: i32 = 13;
pub const RK_FIELD_BRAKE_A: i32 = 14;
pub const RK_FIELD_BRAKE_V: i32 = 15;
pub const RK_FIELD_BRAKE_P: i32 = 16;
pub const RK_FIELD_PF_VF_AF: i32 = 17;
pub const RK_FIELD_ERROR_DETAIL: i32 = 18;
#[repr(C)] pub struct rk_handle { _p: [u8; 0] }
#[repr(C)] pub struct rk_trajectories { _p: [u8; 0] }
#[repr(C)] pub struct rk_rollout { _p: [u8; 0] }
#[repr(C)]
pub struct rk_update_out {           // OutputParameter fields, [dof][n] / [n] device arrays, all nullable
    pub new_position: *mut f64, pub new_velocity: *mut f64, pub new_acceleration: *mut f64, pub new_jerk: *mut f64,
    pub result: *mut i32, pub new_calculation: *mut u8, pub time: *mut f64,
    pub pass_to_input: i32,                    // OutputParameter::pass_to_input into the caller's current_* arrays
    pub reset_time_on_new_calculation: i32,    // 0 = rsruckig's behaviour (time keeps running), 1 = upstream Ruckig's
}

#[repr(C)]
pub struct rk_batch_desc {
    pub n: i64,
    pub dof: i32,
    pub control_interface: i32,        // ControlInterface as i32   (input_parameter.rs:35-45)
    pub synchronization: i32,          // Synchronization as i32    (input_parameter.rs:72-85)
    pub duration_discretization: i32,  // DurationDiscretization    (input_parameter.rs:110-117)
    pub error_mode: i32,               // 0 = ThrowErrorHandler, 1 = IgnoreErrorHandler (error.rs:98-122)
    pub memory_space: i32,             // 0 = host pointers, 1 = device pointers
    pub limits_layout: i32,            // 0 = limit arrays [dof][n], 1 = [dof] shared by the whole batch (RK_LIMITS_PER_DOF)
    pub delta_time: f64,               // Ruckig::delta_time
    pub per_dof_control_interface: *const i32, // [dof] or null
    pub per_dof_synchronization: *const i32,   // [dof] or null
}

#[repr(C)]
pub struct rk_batch_in {               // SoA, element (dof d, trajectory i) at ptr[d * n + i]
    pub current_position: *const f64, pub current_velocity: *const f64, pub current_acceleration: *const f64,
    pub target_position: *const f64,  pub target_velocity: *const f64,  pub target_acceleration: *const f64,
    pub max_velocity: *const f64,     pub max_acceleration: *const f64, pub max_jerk: *const f64,
    pub min_velocity: *const f64,     pub min_acceleration: *const f64, // null = -max
    pub enabled: *const u8,                                               // null = all enabled
    pub minimum_duration: *const f64,                                     // [n], NaN = None; null = none
}

#[repr(C)]
pub struct rk_batch_out { pub status: *mut i32, pub duration: *mut f64, pub independent_min_durations: *mut f64 }
#[repr(C)]
pub struct rk_sample_desc { pub time_start: f64, pub delta_time: f64, pub steps: i32, pub memory_space: i32 }
#[repr(C)]
pub struct rk_sample_out { pub position: *mut f64, pub velocity: *mut f64, pub acceleration: *mut f64, pub jerk: *mut f64, pub section: *mut i32 }

#[link(name = "ruckig_b200")]
extern "C" {
    pub fn rk_create(device: i32, out: *mut *mut rk_handle) -> rk_status;                      // Ruckig::new        ruckig.rs:110
    pub fn rk_destroy(h: *mut rk_handle) -> rk_status;
    pub fn rk_trajectories_create(h: *mut rk_handle, n: i64, dof: i32, out: *mut *mut rk_trajectories) -> rk_status; // Trajectory::new trajectory.rs:37
    pub fn rk_trajectories_create_ex(h: *mut rk_handle, n: i64, dof: i32, result_layout: i32, out: *mut *mut rk_trajectories) -> rk_status; // RK_RESULT_FULL | RK_RESULT_COMPACT
    pub fn rk_trajectories_destroy(t: *mut rk_trajectories) -> rk_status;
    pub fn rk_debug_counters(h: *mut rk_handle, out: *mut u32) -> rk_status; // diagnostics: 112 work-list counters of the last chunk
    pub fn rk_calculate_batch(h: *mut rk_handle, desc: *const rk_batch_desc, input: *const rk_batch_in,
                              traj: *mut rk_trajectories, out: *const rk_batch_out) -> rk_status;                 // Ruckig::calculate  ruckig.rs:210
    pub fn rk_shard_range(n: i64, parts: i32, part: i32, first: *mut i64, count: *mut i64);
    pub fn rk_calculate_batch_multi(handles: *const *mut rk_handle, trajs: *const *mut rk_trajectories, parts: i32, desc: *const rk_batch_desc,
                                    input: *const rk_batch_in, out: *const rk_batch_out) -> rk_status;                // one host batch over all GPUs
    pub fn rk_validate_batch(h: *mut rk_handle, desc: *const rk_batch_desc, input: *const rk_batch_in,
                             check_current: i32, check_target: i32, valid: *mut u8, reason: *mut i32) -> rk_status; // validate_input ruckig.rs:150
    pub fn rk_at_time_batch(h: *mut rk_handle, traj: *const rk_trajectories, desc: *const rk_sample_desc,
                            out: *const rk_sample_out) -> rk_status;                                              // Trajectory::at_time trajectory.rs:158
    pub fn rk_trajectories_read(h: *mut rk_handle, traj: *const rk_trajectories, field: i32, dst: *mut c_void, dst_memory_space: i32) -> rk_status;
    pub fn rk_position_extrema_batch(h: *mut rk_handle, traj: *const rk_trajectories, p_min: *mut f64, p_max: *mut f64,
                                     t_min: *mut f64, t_max: *mut f64, memory_space: i32) -> rk_status;              // Trajectory::get_position_extrema trajectory.rs:211
    pub fn rk_first_time_at_position_batch(h: *mut rk_handle, traj: *const rk_trajectories, position: *const f64, time: *mut f64,
                                           memory_space: i32) -> rk_status;                                          // get_first_time_at_position      trajectory.rs:233
    pub fn rk_rollout_create(h: *mut rk_handle, n: i64, dof: i32, out: *mut *mut rk_rollout) -> rk_status;         // n x (Ruckig, OutputParameter)
    pub fn rk_rollout_destroy(r: *mut rk_rollout) -> rk_status;
    pub fn rk_rollout_reset(h: *mut rk_handle, r: *mut rk_rollout) -> rk_status;                                    // Ruckig::reset      ruckig.rs:125
    pub fn rk_rollout_trajectories(r: *mut rk_rollout) -> *mut rk_trajectories;
    pub fn rk_update_batch(h: *mut rk_handle, desc: *const rk_batch_desc, input: *const rk_batch_in, r: *mut rk_rollout,
                           out: *const rk_update_out, recalculated: *mut i64) -> rk_status;                          // Ruckig::update     ruckig.rs:266
    pub fn rk_sync(h: *mut rk_handle) -> rk_status;
    pub fn rk_stream(h: *mut rk_handle) -> *mut c_void;
    pub fn rk_kernel_launches(h: *const rk_handle) -> i64;
    pub fn rk_measure_fp64_peak(h: *mut rk_handle, milliseconds: f64, tflops: *mut f64) -> rk_status;
    pub fn rk_selftest_division(h: *mut rk_handle, n: i64, seed: u64, mismatches: *mut i64) -> rk_status;
    pub fn rk_status_string(s: rk_status) -> *const c_char;
    pub fn rk_last_error() -> *const c_char;
    pub fn rk_version() -> *const c_char;
}
