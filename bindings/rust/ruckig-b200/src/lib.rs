//! rsruckig's public API on top of `libruckig_b200.so` (B200, sm_100a).
//!
//! SOURCE ONLY: never compiled where this repository is built (no Rust toolchain there), see ../README.md. The logic is a
//! transliteration of the two bindings that ARE built and tested — `include/ruckig_b200.hpp` (C++) and
//! `rsruckig_b200/api.py` (Python).
//!
//! * [`GpuRuckig<DOF, E>`] — `Ruckig::<DOF, E>::{new, reset, validate_input, calculate, update}` (ruckig.rs:110-321) for one
//!   problem, computed on the GPU as a batch of one.
//! * [`GpuTrajectory<DOF>`] — `Trajectory::{at_time, get_duration, get_independent_min_durations}` (trajectory.rs:158-209);
//!   the profiles stay on the device.
//! * [`BatchRuckig<DOF, E>`] — `calculate` / `at_time` for n independent problems in SoA buffers ([`BatchInput`]).
//!
//! Inputs, result codes, errors and the error-handler trait are rsruckig's own types. There is no CPU path: without a CUDA
//! device `new` returns `RuckigError::CalculatorError("RK_ERR_NO_DEVICE: ...")`.
use core::marker::PhantomData;
use std::ffi::CStr;
use std::ptr;

use rsruckig::prelude::*;
use ruckig_b200_sys as sys;

fn lib_error(st: sys::rk_status) -> RuckigError {
    // SAFETY: both functions return pointers to static / thread-local NUL-terminated strings owned by the library.
    let (name, detail) = unsafe {
        (CStr::from_ptr(sys::rk_status_string(st)).to_string_lossy().into_owned(), CStr::from_ptr(sys::rk_last_error()).to_string_lossy().into_owned())
    };
    RuckigError::CalculatorError(format!("{}: {}", name, detail))
}

fn check(st: sys::rk_status) -> Result<(), RuckigError> {
    if st == sys::RK_OK { Ok(()) } else { Err(lib_error(st)) }
}

/// result.rs:34-61
pub fn result_from_code(code: i32) -> RuckigResult {
    match code {
        0 => RuckigResult::Working,
        1 => RuckigResult::Finished,
        -100 => RuckigResult::ErrorInvalidInput,
        -101 => RuckigResult::ErrorTrajectoryDuration,
        -102 => RuckigResult::ErrorPositionalLimits,
        -104 => RuckigResult::ErrorZeroLimits,
        -110 => RuckigResult::ErrorExecutionTimeCalculation,
        -111 => RuckigResult::ErrorSynchronizationCalculation,
        _ => RuckigResult::Error,
    }
}

/// Texts of `InputParameter::validate` (input_parameter.rs:251-420), keyed by the `RK_INVALID_*` id of the first failing check.
pub fn validation_message(reason: i32) -> String {
    const TEXT: [&str; 25] = [
        "invalid input",
        "Maximum jerk limit of DoF {} should be larger than or equal to zero.",
        "maximum acceleration limit of DoF {} should be larger than or equal to zero.",
        "minimum acceleration limit of DoF {} should be smaller than or equal to zero.",
        "current acceleration of DoF {} should be a valid number.",
        "target acceleration of DoF {} should be a valid number.",
        "current acceleration of DoF {} exceeds its maximum acceleration limit.",
        "current acceleration of DoF {} undercuts its minimum acceleration limit.",
        "target acceleration of DoF {} exceeds its maximum acceleration limit.",
        "target acceleration of DoF {} undercuts its minimum acceleration limit.",
        "current velocity of DoF {} should be a valid number.",
        "target velocity of DoF {} should be a valid number.",
        "current position of DoF {} should be a valid number.",
        "target position of DoF {} should be a valid number.",
        "maximum velocity limit of DoF {} should be larger than or equal to zero.",
        "minimum velocity limit of DoF {} should be smaller than or equal to zero.",
        "current velocity of DoF {} exceeds its maximum velocity limit.",
        "current velocity of DoF {} undercuts its minimum velocity limit.",
        "target velocity of DoF {} exceeds its maximum velocity limit.",
        "target velocity of DoF {} undercuts its minimum velocity limit.",
        "DoF {} will inevitably reach a velocity from the current kinematic state that will exceed its maximum velocity limit.",
        "DoF {} will inevitably reach a velocity from the current kinematic state that will undercut its minimum velocity limit.",
        "DoF {} will inevitably have reached a velocity from the target kinematic state that will exceed its maximum velocity limit.",
        "DoF {} will inevitably have reached a velocity from the target kinematic state that will undercut its minimum velocity limit.",
        "delta time (control rate) parameter should be larger than zero.",
    ];
    let (id, dof) = ((reason >> 8) as usize, reason & 0xFF);
    TEXT[if (1..TEXT.len()).contains(&id) { id } else { 0 }].replace("{}", &dof.to_string())
}

/// Texts of calculator_target.rs:459-470, :509-517, :820-825 rebuilt from the status and the error-detail word.
pub fn calculator_message(code: i32, detail: i32) -> String {
    let (stage, dof, zero) = (detail >> 8, detail & 0xFF, code == -104);
    match stage {
        1 if zero => format!("zero limits conflict in step 1, dof: {}", dof),
        1 => format!("error in step 1, dof: {}", dof),
        2 if zero => "zero limits conflict with other degrees of freedom in time synchronization".to_string(),
        2 => "error in time synchronization".to_string(),
        _ => format!("error in step 2, dof: {}", dof),
    }
}

fn is_calculator_error(code: i32) -> bool {
    code == -104 || code == -110 || code == -111
}

// ------------------------------------------------------------------------------------------------------------- BatchInput
/// SoA input buffers `[dof][n]` (element (d, i) at `d * n + i`): the per-trajectory fields of `InputParameter<DOF>`
/// (input_parameter.rs:147-188) for n problems, plus the settings that are shared by the batch.
#[derive(Clone, Debug)]
pub struct BatchInput<const DOF: usize> {
    pub n: usize,
    pub degrees_of_freedom: usize,
    pub control_interface: ControlInterface,
    pub synchronization: Synchronization,
    pub duration_discretization: DurationDiscretization,
    pub per_dof_control_interface: Option<Vec<ControlInterface>>,
    pub per_dof_synchronization: Option<Vec<Synchronization>>,
    pub current_position: Vec<f64>,
    pub current_velocity: Vec<f64>,
    pub current_acceleration: Vec<f64>,
    pub target_position: Vec<f64>,
    pub target_velocity: Vec<f64>,
    pub target_acceleration: Vec<f64>,
    pub max_velocity: Vec<f64>,
    pub max_acceleration: Vec<f64>,
    pub max_jerk: Vec<f64>,
    pub min_velocity: Option<Vec<f64>>,
    pub min_acceleration: Option<Vec<f64>>,
    pub enabled: Vec<u8>,
    /// `[n]`, NaN = None
    pub minimum_duration: Option<Vec<f64>>,
}

fn control_interface_code(c: ControlInterface) -> i32 {
    match c { ControlInterface::Position => sys::RK_POSITION, ControlInterface::Velocity => sys::RK_VELOCITY, ControlInterface::Acceleration => sys::RK_ACCELERATION }
}
fn synchronization_code(s: Synchronization) -> i32 {
    match s {
        Synchronization::Time => sys::RK_SYNC_TIME,
        Synchronization::TimeIfNecessary => sys::RK_SYNC_TIME_IF_NECESSARY,
        Synchronization::Phase => sys::RK_SYNC_PHASE,
        Synchronization::None => sys::RK_SYNC_NONE,
    }
}
fn discretization_code(d: &DurationDiscretization) -> i32 {
    match d { DurationDiscretization::Continuous => sys::RK_CONTINUOUS, DurationDiscretization::Discrete => sys::RK_DISCRETE }
}

impl<const DOF: usize> BatchInput<DOF> {
    /// Transposes `inputs` into SoA buffers. The batch-level settings (control interface, synchronisation, discretisation,
    /// per-DoF settings, presence of minimum limits) are taken from `inputs[0]` and must be the same for every element.
    pub fn from_inputs(inputs: &[InputParameter<DOF>]) -> Result<Self, RuckigError> {
        let first = inputs.first().ok_or_else(|| RuckigError::ValidationError("empty batch".to_string()))?;
        let (n, d) = (inputs.len(), first.degrees_of_freedom);
        let soa = |get: fn(&InputParameter<DOF>, usize) -> f64| -> Vec<f64> {
            let mut v = vec![0.0; d * n];
            for (i, ip) in inputs.iter().enumerate() { for k in 0..d { v[k * n + i] = get(ip, k); } }
            v
        };
        for ip in inputs {
            if ip.degrees_of_freedom != d
                || ip.control_interface != first.control_interface
                || ip.synchronization != first.synchronization
                || ip.duration_discretization != first.duration_discretization
                || ip.per_dof_control_interface != first.per_dof_control_interface
                || ip.per_dof_synchronization != first.per_dof_synchronization
                || ip.min_velocity.is_some() != first.min_velocity.is_some()
                || ip.min_acceleration.is_some() != first.min_acceleration.is_some()
            {
                return Err(RuckigError::ValidationError("the batch-level settings differ between the inputs of one batch".to_string()));
            }
        }
        let mut enabled = vec![1u8; d * n];
        for (i, ip) in inputs.iter().enumerate() { for k in 0..d { enabled[k * n + i] = ip.enabled[k] as u8; } }
        let minimum_duration = if inputs.iter().any(|ip| ip.minimum_duration.is_some()) {
            Some(inputs.iter().map(|ip| ip.minimum_duration.unwrap_or(f64::NAN)).collect())
        } else {
            None
        };
        Ok(Self {
            n,
            degrees_of_freedom: d,
            control_interface: first.control_interface,
            synchronization: first.synchronization,
            duration_discretization: first.duration_discretization.clone(),
            per_dof_control_interface: first.per_dof_control_interface.as_ref().map(|v| v.iter().copied().collect()),
            per_dof_synchronization: first.per_dof_synchronization.as_ref().map(|v| v.iter().copied().collect()),
            current_position: soa(|ip, k| ip.current_position[k]),
            current_velocity: soa(|ip, k| ip.current_velocity[k]),
            current_acceleration: soa(|ip, k| ip.current_acceleration[k]),
            target_position: soa(|ip, k| ip.target_position[k]),
            target_velocity: soa(|ip, k| ip.target_velocity[k]),
            target_acceleration: soa(|ip, k| ip.target_acceleration[k]),
            max_velocity: soa(|ip, k| ip.max_velocity[k]),
            max_acceleration: soa(|ip, k| ip.max_acceleration[k]),
            max_jerk: soa(|ip, k| ip.max_jerk[k]),
            min_velocity: first.min_velocity.as_ref().map(|_| soa(|ip, k| ip.min_velocity.as_ref().unwrap()[k])),
            min_acceleration: first.min_acceleration.as_ref().map(|_| soa(|ip, k| ip.min_acceleration.as_ref().unwrap()[k])),
            enabled,
            minimum_duration,
        })
    }

    /// The C structs that describe these buffers. `ci` / `sy` receive the per-DoF codes and must outlive the returned structs.
    fn as_ffi(&self, delta_time: f64, error_mode: i32, ci: &mut Vec<i32>, sy: &mut Vec<i32>) -> (sys::rk_batch_desc, sys::rk_batch_in) {
        let opt = |v: &Option<Vec<f64>>| v.as_ref().map_or(ptr::null(), |v| v.as_ptr());
        if let Some(v) = &self.per_dof_control_interface { *ci = v.iter().map(|c| control_interface_code(*c)).collect(); }
        if let Some(v) = &self.per_dof_synchronization { *sy = v.iter().map(|s| synchronization_code(*s)).collect(); }
        let desc = sys::rk_batch_desc {
            n: self.n as i64,
            dof: self.degrees_of_freedom as i32,
            control_interface: control_interface_code(self.control_interface),
            synchronization: synchronization_code(self.synchronization),
            duration_discretization: discretization_code(&self.duration_discretization),
            error_mode,
            memory_space: sys::RK_MEM_HOST,
            limits_layout: sys::RK_LIMITS_PER_TRAJECTORY,
            delta_time,
            per_dof_control_interface: if self.per_dof_control_interface.is_some() { ci.as_ptr() } else { ptr::null() },
            per_dof_synchronization: if self.per_dof_synchronization.is_some() { sy.as_ptr() } else { ptr::null() },
        };
        let input = sys::rk_batch_in {
            current_position: self.current_position.as_ptr(),
            current_velocity: self.current_velocity.as_ptr(),
            current_acceleration: self.current_acceleration.as_ptr(),
            target_position: self.target_position.as_ptr(),
            target_velocity: self.target_velocity.as_ptr(),
            target_acceleration: self.target_acceleration.as_ptr(),
            max_velocity: self.max_velocity.as_ptr(),
            max_acceleration: self.max_acceleration.as_ptr(),
            max_jerk: self.max_jerk.as_ptr(),
            min_velocity: opt(&self.min_velocity),
            min_acceleration: opt(&self.min_acceleration),
            enabled: self.enabled.as_ptr(),
            minimum_duration: opt(&self.minimum_duration),
        };
        (desc, input)
    }
}

// ---------------------------------------------------------------------------------------------------------------- engine
/// One `rk_handle` (stream + scratch) per device and thread, shared by every object of this crate created on that thread and
/// never destroyed (the C++ and Python bindings do the same): trajectories may outlive the calculator that filled them, and
/// the library needs their handle to free them. Calls on a handle are stream ordered and not thread safe; the objects below
/// hold the raw pointer, so they are neither `Send` nor `Sync` and stay on the thread whose handle they use.
fn engine(device: i32) -> Result<*mut sys::rk_handle, RuckigError> {
    thread_local! {
        static ENGINES: std::cell::RefCell<std::collections::HashMap<i32, *mut sys::rk_handle>> = Default::default();
    }
    ENGINES.with(|engines| {
        if let Some(&h) = engines.borrow().get(&device) {
            return Ok(h);
        }
        let mut h = ptr::null_mut();
        // SAFETY: `h` is a valid out-pointer; on failure the library leaves it null.
        check(unsafe { sys::rk_create(device, &mut h) })?;
        engines.borrow_mut().insert(device, h);
        Ok(h)
    })
}

/// Device-resident trajectories of one batch (`Vec<Trajectory<DOF>>`, trajectory.rs:14-21).
struct Trajectories { t: *mut sys::rk_trajectories, h: *mut sys::rk_handle, n: usize, dof: usize }

/// What stays on the device per (DoF, trajectory): the reference's `Profile` verbatim (59 doubles) or the 20 doubles the
/// calculation decided; every reader returns identical values either way (ruckig_b200.h, RK_RESULT_*).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum ResultLayout { Full = sys::RK_RESULT_FULL as isize, Compact = sys::RK_RESULT_COMPACT as isize }

impl Trajectories {
    fn new(h: *mut sys::rk_handle, n: usize, dof: usize) -> Result<Self, RuckigError> { Self::with_layout(h, n, dof, ResultLayout::Full) }
    fn with_layout(h: *mut sys::rk_handle, n: usize, dof: usize, layout: ResultLayout) -> Result<Self, RuckigError> {
        let mut t = ptr::null_mut();
        // SAFETY: valid handle and out-pointer.
        check(unsafe { sys::rk_trajectories_create_ex(h, n as i64, dof as i32, layout as i32, &mut t) })?;
        Ok(Self { t, h, n, dof })
    }
    fn read_i32(&self, field: i32, len: usize) -> Result<Vec<i32>, RuckigError> {
        let mut v = vec![0i32; len];
        // SAFETY: `v` has the element count the header documents for `field`.
        check(unsafe { sys::rk_trajectories_read(self.h, self.t, field, v.as_mut_ptr().cast(), sys::RK_MEM_HOST) })?;
        Ok(v)
    }
}
impl Drop for Trajectories {
    fn drop(&mut self) {
        // SAFETY: created by rk_trajectories_create and destroyed exactly once; its handle lives as long as the thread.
        unsafe { sys::rk_trajectories_destroy(self.t) };
    }
}

// ------------------------------------------------------------------------------------------------------------ MultiBatchRuckig
/// One host batch over several CUDA devices (`rk_calculate_batch_multi`): contiguous index ranges (`rk_shard_range`), one
/// handle per entry of `devices`, the parts run concurrently inside the library and gather status / duration into the
/// caller's vectors. north_star: "sharded trivially across the 8 GPUs of one box with no collectives beyond a final host gather".
pub struct MultiBatchRuckig<const DOF: usize> {
    pub degrees_of_freedom: usize,
    pub delta_time: f64,
    layout: ResultLayout,
    handles: Vec<*mut sys::rk_handle>,
    parts: Vec<Trajectories>,
    n: usize,
}

impl<const DOF: usize> MultiBatchRuckig<DOF> {
    pub fn new(dofs: Option<usize>, delta_time: f64, devices: &[i32], layout: ResultLayout) -> Result<Self, RuckigError> {
        let mut handles = Vec::new();
        for &d in devices {
            let mut h = ptr::null_mut();
            // SAFETY: `h` is a valid out-pointer; every part owns its own handle (a handle is one stream + scratch).
            check(unsafe { sys::rk_create(d, &mut h) })?;
            handles.push(h);
        }
        Ok(Self { degrees_of_freedom: dofs.unwrap_or(DOF), delta_time, layout, handles, parts: Vec::new(), n: usize::MAX })
    }

    /// `Ruckig::calculate` for every problem; `status[i]` is the `RuckigResult` value (-100 = `Err(ValidationError)` under the
    /// throwing semantics the library is asked for here).
    pub fn calculate(&mut self, input: &BatchInput<DOF>, status: &mut Vec<i32>, durations: &mut Vec<f64>) -> Result<(), RuckigError> {
        let (n, d, g) = (input.n, input.degrees_of_freedom, self.handles.len());
        if n != self.n {
            self.parts.clear();
            for (k, &h) in self.handles.iter().enumerate() {
                let (mut first, mut count) = (0i64, 0i64);
                // SAFETY: plain arithmetic on the two out-pointers.
                unsafe { sys::rk_shard_range(n as i64, g as i32, k as i32, &mut first, &mut count) };
                self.parts.push(Trajectories::with_layout(h, count as usize, d, self.layout)?);
            }
            self.n = n;
        }
        status.resize(n, 0);
        durations.resize(n, 0.0);
        let (mut ci, mut sy) = (Vec::new(), Vec::new());
        let (desc, inp) = input.as_ffi(self.delta_time, sys::RK_ERRORS_THROW, &mut ci, &mut sy);
        let out = sys::rk_batch_out { status: status.as_mut_ptr(), duration: durations.as_mut_ptr(), independent_min_durations: ptr::null_mut() };
        let trajs: Vec<*mut sys::rk_trajectories> = self.parts.iter().map(|p| p.t).collect();
        // SAFETY: `handles` and `trajs` have `g` live entries each; desc / inp / out point to live host buffers of the documented sizes.
        check(unsafe { sys::rk_calculate_batch_multi(self.handles.as_ptr(), trajs.as_ptr(), g as i32, &desc, &inp, &out) })
    }
}

impl<const DOF: usize> Drop for MultiBatchRuckig<DOF> {
    fn drop(&mut self) {
        self.parts.clear();
        // SAFETY: created by rk_create above, destroyed exactly once, after the trajectories that refer to them.
        for &h in &self.handles { unsafe { sys::rk_destroy(h) }; }
    }
}

// ------------------------------------------------------------------------------------------------------------ BatchRuckig
/// The batched entry point: `Ruckig::calculate` and `Trajectory::at_time` for n independent problems.
pub struct BatchRuckig<const DOF: usize, E: RuckigErrorHandler> {
    pub degrees_of_freedom: usize,
    pub delta_time: f64,
    traj: Option<Trajectories>,
    handle: *mut sys::rk_handle,
    ignores_validation_errors: bool,
    _e: PhantomData<E>,
}

impl<const DOF: usize, E: RuckigErrorHandler> BatchRuckig<DOF, E> {
    /// `Ruckig::new(dofs, delta_time)` (ruckig.rs:110-119) on CUDA device `device`.
    pub fn new(dofs: Option<usize>, delta_time: f64, device: i32) -> Result<Self, RuckigError> {
        Ok(Self { degrees_of_freedom: dofs.unwrap_or(DOF), delta_time, traj: None, handle: engine(device)?, ignores_validation_errors: false, _e: PhantomData })
    }

    /// `Ruckig::calculate` for every problem of the batch. `results[i]` is what the reference returns for problem i:
    /// a validation failure goes through `E::handle_validation_error` (`Err` under `ThrowErrorHandler`; a handler that
    /// ignores it makes the calculation proceed, as in ruckig.rs:215-216 — the batch is then computed in the library's
    /// ignore mode from that call on), `ErrorZeroLimits` / `ErrorExecutionTimeCalculation` /
    /// `ErrorSynchronizationCalculation` go through `E::handle_calculator_error` and come back as `Ok(code)` if it lets
    /// them pass, `ErrorTrajectoryDuration` is always `Ok(code)` (calculator_target.rs:531-533).
    /// The outer `Err` is a failure of the library itself (no device, out of memory).
    pub fn calculate(&mut self, input: &BatchInput<DOF>, durations: &mut Vec<f64>) -> Result<Vec<Result<RuckigResult, RuckigError>>, RuckigError> {
        let (n, d) = (input.n, input.degrees_of_freedom);
        if self.traj.as_ref().map_or(true, |t| t.n != n || t.dof != d) {
            self.traj = None;
            self.traj = Some(Trajectories::new(self.handle, n, d)?);
        }
        let mut status = vec![0i32; n];
        durations.resize(n, 0.0);
        loop {
            let (mut ci, mut sy) = (Vec::new(), Vec::new());
            let mode = if self.ignores_validation_errors { sys::RK_ERRORS_IGNORE } else { sys::RK_ERRORS_THROW };
            let (desc, inp) = input.as_ffi(self.delta_time, mode, &mut ci, &mut sy);
            let out = sys::rk_batch_out { status: status.as_mut_ptr(), duration: durations.as_mut_ptr(), independent_min_durations: ptr::null_mut() };
            let traj = self.traj.as_ref().unwrap();
            // SAFETY: every pointer of desc / inp / out refers to a live buffer of the documented size; host memory space.
            check(unsafe { sys::rk_calculate_batch(self.handle, &desc, &inp, traj.t, &out) })?;
            if self.ignores_validation_errors || !status.iter().any(|&s| s == -100) { break; }
            // first validation failure seen by this object: ask the handler what it does with one
            let detail = traj.read_i32(sys::RK_FIELD_ERROR_DETAIL, n)?;
            let i = status.iter().position(|&s| s == -100).unwrap();
            if E::handle_validation_error(&validation_message(detail[i])).is_err() { break; }
            self.ignores_validation_errors = true;  // the handler lets invalid inputs through: compute them as the reference does
        }
        let traj = self.traj.as_ref().unwrap();
        let detail = if status.iter().any(|&s| s == -100 || is_calculator_error(s)) { traj.read_i32(sys::RK_FIELD_ERROR_DETAIL, n)? } else { Vec::new() };
        Ok(status.iter().enumerate().map(|(i, &s)| {
            if s == -100 { E::handle_validation_error(&validation_message(detail[i]))?; }
            if is_calculator_error(s) { E::handle_calculator_error(&calculator_message(s, detail[i]))?; }
            Ok(result_from_code(s))
        }).collect())
    }

    /// `Trajectory::at_time` for `steps` control cycles of every trajectory (sample k at `time_start` + (k + 1) additions of
    /// `cycle_time`, what `steps` calls of `Ruckig::update` produce): `position` etc. are `[steps][dof][n]`.
    pub fn at_time(&self, time_start: f64, cycle_time: f64, steps: usize, position: &mut [f64], velocity: Option<&mut [f64]>,
                   acceleration: Option<&mut [f64]>, jerk: Option<&mut [f64]>) -> Result<(), RuckigError> {
        let traj = self.traj.as_ref().ok_or_else(|| RuckigError::CalculatorError("calculate() has not been called".to_string()))?;
        let want = steps * traj.dof * traj.n;
        let out_ptr = |s: Option<&mut [f64]>| -> Result<*mut f64, RuckigError> {
            match s {
                None => Ok(ptr::null_mut()),
                Some(s) if s.len() == want => Ok(s.as_mut_ptr()),
                Some(_) => Err(RuckigError::CalculatorError("output slice must hold steps * dof * n values".to_string())),
            }
        };
        let so = sys::rk_sample_out { position: out_ptr(Some(position))?, velocity: out_ptr(velocity)?, acceleration: out_ptr(acceleration)?, jerk: out_ptr(jerk)?, section: ptr::null_mut() };
        let sd = sys::rk_sample_desc { time_start, delta_time: cycle_time, steps: steps as i32, memory_space: sys::RK_MEM_HOST };
        // SAFETY: non-null outputs hold steps * dof * n doubles (checked above).
        check(unsafe { sys::rk_at_time_batch(self.handle, traj.t, &sd, &so) })
    }
}

// ------------------------------------------------------------------------------------------------- GpuTrajectory / GpuRuckig
/// `Trajectory<DOF>` (trajectory.rs:14-21) whose profiles live on the GPU.
pub struct GpuTrajectory<const DOF: usize> {
    pub degrees_of_freedom: usize,
    duration: f64,
    independent_min_durations: Vec<f64>,
    dev: Option<Trajectories>,
}

impl<const DOF: usize> GpuTrajectory<DOF> {
    pub fn new(dofs: Option<usize>) -> Self {
        let d = dofs.unwrap_or(DOF);
        Self { degrees_of_freedom: d, duration: 0.0, independent_min_durations: vec![0.0; d], dev: None }
    }
    pub fn get_duration(&self) -> f64 { self.duration }                                             // trajectory.rs:197-199
    pub fn get_independent_min_durations(&self) -> &[f64] { &self.independent_min_durations }       // :207-209

    /// trajectory.rs:158-189; every output is optional, `new_section` as in the reference.
    pub fn at_time(&self, time: f64, new_position: &mut Option<&mut [f64]>, new_velocity: &mut Option<&mut [f64]>,
                   new_acceleration: &mut Option<&mut [f64]>, new_jerk: &mut Option<&mut [f64]>, new_section: &mut Option<usize>) -> Result<(), RuckigError> {
        let traj = self.dev.as_ref().ok_or_else(|| RuckigError::CalculatorError("calculate() has not filled this trajectory".to_string()))?;
        let d = self.degrees_of_freedom;
        let out_ptr = |s: &mut Option<&mut [f64]>| -> *mut f64 { match s { Some(s) if s.len() >= d => s.as_mut_ptr(), _ => ptr::null_mut() } };
        let mut section = 0i32;
        let so = sys::rk_sample_out { position: out_ptr(new_position), velocity: out_ptr(new_velocity), acceleration: out_ptr(new_acceleration), jerk: out_ptr(new_jerk), section: &mut section };
        // one sample, taken after one addition of delta_time = 0.0 to time_start = time
        let sd = sys::rk_sample_desc { time_start: time, delta_time: 0.0, steps: 1, memory_space: sys::RK_MEM_HOST };
        // SAFETY: non-null outputs hold at least `dof` doubles.
        check(unsafe { sys::rk_at_time_batch(traj.h, traj.t, &sd, &so) })?;
        *new_section = Some(section as usize);
        Ok(())
    }
}

/// `OutputParameter<DOF>` (output_parameter.rs:36-61) around a [`GpuTrajectory`].
pub struct GpuOutput<const DOF: usize> {
    pub degrees_of_freedom: usize,
    pub trajectory: GpuTrajectory<DOF>,
    pub new_position: Vec<f64>,
    pub new_velocity: Vec<f64>,
    pub new_acceleration: Vec<f64>,
    pub new_jerk: Vec<f64>,
    pub time: f64,
    pub new_section: usize,
    pub did_section_change: bool,
    pub new_calculation: bool,
    pub was_calculation_interrupted: bool,
    pub calculation_duration: f64,
}

impl<const DOF: usize> GpuOutput<DOF> {
    pub fn new(dofs: Option<usize>) -> Self {
        let d = dofs.unwrap_or(DOF);
        Self { degrees_of_freedom: d, trajectory: GpuTrajectory::new(dofs), new_position: vec![0.0; d], new_velocity: vec![0.0; d], new_acceleration: vec![0.0; d],
               new_jerk: vec![0.0; d], time: 0.0, new_section: 0, did_section_change: false, new_calculation: false, was_calculation_interrupted: false, calculation_duration: 0.0 }
    }
    /// output_parameter.rs:110-120
    pub fn pass_to_input(&self, input: &mut InputParameter<DOF>) {
        for k in 0..self.degrees_of_freedom {
            input.current_position[k] = self.new_position[k];
            input.current_velocity[k] = self.new_velocity[k];
            input.current_acceleration[k] = self.new_acceleration[k];
        }
    }
}

/// `Ruckig<DOF, E>` (ruckig.rs:76-321) for a single problem; every calculation runs on the GPU as a batch of one.
pub struct GpuRuckig<const DOF: usize, E: RuckigErrorHandler> {
    pub degrees_of_freedom: usize,
    pub delta_time: f64,
    current_input: InputParameter<DOF>,
    current_input_initialized: bool,
    handle: *mut sys::rk_handle,
    _e: PhantomData<E>,
}

impl<const DOF: usize, E: RuckigErrorHandler> GpuRuckig<DOF, E> {
    pub fn new(dofs: Option<usize>, delta_time: f64) -> Result<Self, RuckigError> {
        Ok(Self { degrees_of_freedom: dofs.unwrap_or(DOF), delta_time, current_input: InputParameter::new(dofs), current_input_initialized: false, handle: engine(0)?, _e: PhantomData })
    }

    pub fn reset(&mut self) { self.current_input_initialized = false; }  // ruckig.rs:125-127

    /// ruckig.rs:150-171
    pub fn validate_input(&self, input: &InputParameter<DOF>, check_current_state_within_limits: bool, check_target_state_within_limits: bool) -> Result<bool, RuckigError> {
        let batch = BatchInput::from_inputs(core::slice::from_ref(input))?;
        let (mut ci, mut sy) = (Vec::new(), Vec::new());
        let (desc, inp) = batch.as_ffi(self.delta_time, sys::RK_ERRORS_THROW, &mut ci, &mut sy);
        let (mut valid, mut reason) = (0u8, 0i32);
        // SAFETY: n = 1; `valid` and `reason` hold one element each.
        check(unsafe { sys::rk_validate_batch(self.handle, &desc, &inp, check_current_state_within_limits as i32, check_target_state_within_limits as i32, &mut valid, &mut reason) })?;
        if valid == 0 { E::handle_validation_error(&validation_message(reason))?; }
        Ok(valid != 0)
    }

    /// ruckig.rs:210-218
    pub fn calculate(&mut self, input: &InputParameter<DOF>, traj: &mut GpuTrajectory<DOF>) -> Result<RuckigResult, RuckigError> {
        let d = self.degrees_of_freedom;
        let batch = BatchInput::from_inputs(core::slice::from_ref(input))?;
        if traj.dev.as_ref().map_or(true, |t| t.h != self.handle || t.dof != d) {
            traj.dev = None;
            traj.dev = Some(Trajectories::new(self.handle, 1, d)?);
        }
        let dev = traj.dev.as_ref().unwrap();
        let mut mode = sys::RK_ERRORS_THROW;
        let (mut status, mut duration, mut imd) = (0i32, 0.0f64, vec![0.0f64; d]);
        loop {
            let (mut ci, mut sy) = (Vec::new(), Vec::new());
            let (desc, inp) = batch.as_ffi(self.delta_time, mode, &mut ci, &mut sy);
            let out = sys::rk_batch_out { status: &mut status, duration: &mut duration, independent_min_durations: imd.as_mut_ptr() };
            // SAFETY: n = 1; every pointer refers to a live buffer of the documented size.
            check(unsafe { sys::rk_calculate_batch(self.handle, &desc, &inp, dev.t, &out) })?;
            if status != -100 || mode == sys::RK_ERRORS_IGNORE { break; }
            // ruckig.rs:215-216: the handler decides; if it lets the invalid input through, the calculation proceeds
            let detail = dev.read_i32(sys::RK_FIELD_ERROR_DETAIL, 1)?;
            E::handle_validation_error(&validation_message(detail[0]))?;
            mode = sys::RK_ERRORS_IGNORE;
        }
        let detail = if is_calculator_error(status) { dev.read_i32(sys::RK_FIELD_ERROR_DETAIL, 1)?[0] } else { 0 };
        traj.duration = duration;
        traj.independent_min_durations = imd;
        if is_calculator_error(status) {
            E::handle_calculator_error(&calculator_message(status, detail))?;
        }
        Ok(result_from_code(status))
    }

    /// ruckig.rs:266-321 (output.time is not reset by a new calculation, as in the reference)
    pub fn update(&mut self, input: &InputParameter<DOF>, output: &mut GpuOutput<DOF>) -> Result<RuckigResult, RuckigError> {
        output.new_calculation = false;
        if !self.current_input_initialized || *input != self.current_input {
            self.calculate(input, &mut output.trajectory)?;  // Ok(Error*) codes are dropped, ruckig.rs:285
            output.new_calculation = true;
            self.current_input = input.clone();
            self.current_input_initialized = true;
        }
        output.time += self.delta_time;
        let time = output.time;
        let GpuOutput { trajectory, new_position, new_velocity, new_acceleration, new_jerk, .. } = &mut *output;
        trajectory.at_time(time, &mut Some(&mut new_position[..]), &mut Some(&mut new_velocity[..]), &mut Some(&mut new_acceleration[..]),
                           &mut Some(&mut new_jerk[..]), &mut None)?;
        output.did_section_change = false;  // ruckig.rs:300-302: new_section is passed by value
        output.pass_to_input(&mut self.current_input);
        Ok(if output.time > output.trajectory.get_duration() { RuckigResult::Finished } else { RuckigResult::Working })
    }
}
