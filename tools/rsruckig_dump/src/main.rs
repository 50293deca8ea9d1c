//! rsruckig_dump <inputs.rkb2> <results.rkb2>
//!
//! Reads a batch input file (format: rsruckig_b200/batchfile.py), runs `Ruckig::calculate` of the real rsruckig crate on every
//! trajectory with a fresh calculator and a fresh `Trajectory` (the batched library's semantics), and writes status,
//! duration, profile codes and `Profile::t` in the results format, for `python tools/compare_results.py`.
//! Untested here: the repository's build environment has no Rust toolchain.
use rsruckig::prelude::*;
use std::convert::TryInto;
use std::io::Write;

fn f64s(b: &[u8], off: &mut usize, count: usize) -> Vec<f64> {
    let v = b[*off..*off + 8 * count].chunks_exact(8).map(|c| f64::from_le_bytes(c.try_into().unwrap())).collect();
    *off += 8 * count;
    v
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let b = std::fs::read(&args[1]).expect("inputs file");
    assert_eq!(&b[0..8], b"RKB2IN\0\0");
    assert_eq!(u32::from_le_bytes(b[8..12].try_into().unwrap()), 1);
    let dof = i32::from_le_bytes(b[12..16].try_into().unwrap()) as usize;
    let n = i64::from_le_bytes(b[16..24].try_into().unwrap()) as usize;
    let ci = i32::from_le_bytes(b[24..28].try_into().unwrap());
    let sync = i32::from_le_bytes(b[28..32].try_into().unwrap());
    let disc = i32::from_le_bytes(b[32..36].try_into().unwrap());
    let mask = u32::from_le_bytes(b[36..40].try_into().unwrap());
    let delta_time = f64::from_le_bytes(b[40..48].try_into().unwrap());
    let mut off = 64;
    let arrays: Vec<Vec<f64>> = (0..9).map(|_| f64s(&b, &mut off, dof * n)).collect();
    let min_v = if mask & 1 != 0 { Some(f64s(&b, &mut off, dof * n)) } else { None };
    let min_a = if mask & 2 != 0 { Some(f64s(&b, &mut off, dof * n)) } else { None };
    let enabled = if mask & 4 != 0 { let e = b[off..off + dof * n].to_vec(); off += dof * n; Some(e) } else { None };
    let min_dur = if mask & 8 != 0 { Some(f64s(&b, &mut off, n)) } else { None };

    let (mut status, mut duration) = (vec![0i32; n], vec![0f64; n]);
    let (mut codes, mut t) = (vec![0u32; dof * n], vec![0f64; 7 * dof * n]);
    for i in 0..n {
        let mut otg = Ruckig::<0, ThrowErrorHandler>::new(Some(dof), delta_time);
        let mut input = InputParameter::new(Some(dof));
        let mut traj = Trajectory::new(Some(dof));
        for d in 0..dof {
            let g = d * n + i;
            input.current_position[d] = arrays[0][g]; input.current_velocity[d] = arrays[1][g]; input.current_acceleration[d] = arrays[2][g];
            input.target_position[d] = arrays[3][g]; input.target_velocity[d] = arrays[4][g]; input.target_acceleration[d] = arrays[5][g];
            input.max_velocity[d] = arrays[6][g]; input.max_acceleration[d] = arrays[7][g]; input.max_jerk[d] = arrays[8][g];
            if let Some(e) = &enabled { input.enabled[d] = e[g] != 0; }
        }
        if let Some(v) = &min_v { input.min_velocity = Some(DataArrayOrVec::Heap((0..dof).map(|d| v[d * n + i]).collect::<Vec<f64>>())); }
        if let Some(v) = &min_a { input.min_acceleration = Some(DataArrayOrVec::Heap((0..dof).map(|d| v[d * n + i]).collect::<Vec<f64>>())); }
        if let Some(v) = &min_dur { if !v[i].is_nan() { input.minimum_duration = Some(v[i]); } }
        input.control_interface = if ci == 1 { ControlInterface::Velocity } else { ControlInterface::Position };
        input.synchronization = match sync { 1 => Synchronization::TimeIfNecessary, 2 => Synchronization::Phase, 3 => Synchronization::None, _ => Synchronization::Time };
        input.duration_discretization = if disc == 1 { DurationDiscretization::Discrete } else { DurationDiscretization::Continuous };
        status[i] = match otg.calculate(&input, &mut traj) {
            Ok(r) => r as i32,
            Err(RuckigError::ValidationError(_)) => -100,
            Err(RuckigError::CalculatorError(m)) => if m.contains("zero limits") { -104 } else if m.contains("synchronization") { -111 } else { -110 },
        };
        duration[i] = traj.get_duration();
        for d in 0..dof {
            let p = &traj.get_profiles()[0][d];
            codes[d * n + i] = (p.limits as u32) | ((p.control_signs as u32) << 8) | ((p.direction as u32) << 16);
            for k in 0..7 { t[(k * dof + d) * n + i] = p.t[k]; }
        }
    }
    let mut out = std::fs::File::create(&args[2]).expect("results file");
    let mut header = Vec::with_capacity(64);
    header.extend_from_slice(b"RKB2RS\0\0");
    header.extend_from_slice(&1u32.to_le_bytes());
    header.extend_from_slice(&(dof as i32).to_le_bytes());
    header.extend_from_slice(&(n as i64).to_le_bytes());
    header.resize(64, 0);
    out.write_all(&header).unwrap();
    for v in &status { out.write_all(&v.to_le_bytes()).unwrap(); }
    for v in &duration { out.write_all(&v.to_le_bytes()).unwrap(); }
    for v in &codes { out.write_all(&v.to_le_bytes()).unwrap(); }
    for v in &t { out.write_all(&v.to_le_bytes()).unwrap(); }
}
