//! ffi.rs — binding of libodesat_b200 for the reference crate (UNVERIFIED SOURCE).
//!
//! No Rust toolchain exists in the build image (no cargo/rustc), so this file has never been
//! compiled; it is the shim a maintainer of AHartNtkn/odesat would add as `src/ffi.rs` (plus
//! `pub mod ffi;` in `src/lib.rs` and the `build.rs` shown in INTEGRATION.md).  It mirrors the
//! declarations of include/odesat_b200.h one to one and offers GPU twins of the three call sites
//! of the hot path: `simulate` (main.rs:176), `batch`'s loop (main.rs:278-308) and
//! `simulate_inter` (main.rs:360).

use crate::cnf::CNFFormula;
use crate::system::State;
use std::os::raw::{c_char, c_double, c_int};

#[repr(C)]
pub struct OdesatFormula { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct OdesatParams {
    pub tolerance: c_double,     // Option<f64>: NaN = None
    pub step_size: c_double,     // Option<f64>: NaN = None (adaptive)
    pub steps: i64,              // Option<usize>: < 0 = None
    pub learning_rate: c_double, // Option<f64>: NaN = None (density rule)
    pub precision: i32,          // 0 = f64, 1 = f32
    pub engine: i32,             // 0 = auto
    pub schedule: i32,           // 0 = exact summation order, 1 = balanced
    pub chunk: i32,              // steps between early-exit polls; <= 0 = 32
    pub n_gpus: i32,             // replica batches: devices of this process to shard over; <= 0 = 1
    pub sub_batches: i32,        // shards per device (upload / compute overlap); <= 0 = auto
}

#[link(name = "odesat_b200")]
extern "C" {
    fn odesat_last_error() -> *const c_char;
    fn odesat_device_count() -> c_int;
    fn odesat_formula_create(varnum: i64, n_clauses: i64, clause_off: *const i64, lits: *const i32,
                             out: *mut *mut OdesatFormula) -> c_int;
    fn odesat_formula_destroy(f: *mut OdesatFormula);
    fn odesat_simulate(f: *const OdesatFormula, v: *mut c_double, xs: *mut c_double, xl: *mut c_double,
                       params: *const OdesatParams, assignment: *mut u8, steps_taken: *mut i64,
                       allsat: *mut c_int, final_dt: *mut c_double) -> c_int;
    fn odesat_simulate_batch(f: *const OdesatFormula, r: i64, v: *mut c_double, xs: *mut c_double,
                             xl: *mut c_double, seed: u64, replica_offset: i64,
                             params: *const OdesatParams, mode: i32, write_back: i32,
                             solved_step: *mut i64, verified: *mut u8, winner: *mut i64,
                             assignment: *mut u8, steps_run: *mut i64) -> c_int;
    fn odesat_simulate_inter(f: *const OdesatFormula, r: i64, v: *mut c_double, xs: *mut c_double,
                             xl: *mut c_double, params: *const OdesatParams, assignment: *mut u8,
                             winner: *mut i64, steps_taken: *mut i64) -> c_int;
}

fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(odesat_last_error()) };
        panic!("odesat_b200: status {}: {}", rc, msg.to_string_lossy());
    }
}

fn params(tolerance: Option<f64>, step_size: Option<f64>, steps: Option<usize>, learning_rate: Option<f64>) -> OdesatParams {
    OdesatParams {
        tolerance: tolerance.unwrap_or(f64::NAN),
        step_size: step_size.unwrap_or(f64::NAN),
        steps: steps.map(|s| s as i64).unwrap_or(-1),
        learning_rate: learning_rate.unwrap_or(f64::NAN),
        precision: 0, engine: 0, schedule: 0, chunk: 0,
        // `batch` / `inter` replicas are independent: one process drives every visible GPU (single-state calls ignore it)
        n_gpus: unsafe { odesat_device_count() }.max(1), sub_batches: 0,
    }
}

/// `&CNFFormula` flattened to the CSR the ABI takes; owns the device copy.
pub struct GpuFormula { h: *mut OdesatFormula, varnum: usize, n_clauses: usize }

impl GpuFormula {
    pub fn new(formula: &CNFFormula) -> Self {
        let mut off: Vec<i64> = vec![0];
        let mut lits: Vec<i32> = Vec::new();
        for clause in formula.clauses.iter() {
            for l in clause.literals.iter() {
                let x = (l.variable + 1) as i32;
                lits.push(if l.is_negated { -x } else { x });
            }
            off.push(lits.len() as i64);
        }
        let mut h = std::ptr::null_mut();
        check(unsafe { odesat_formula_create(formula.varnum as i64, formula.clauses.len() as i64, off.as_ptr(), lits.as_ptr(), &mut h) });
        GpuFormula { h, varnum: formula.varnum, n_clauses: formula.clauses.len() }
    }
}
impl Drop for GpuFormula { fn drop(&mut self) { unsafe { odesat_formula_destroy(self.h) } } }

/// Drop-in for `system::simulate` (system.rs:156-163).
pub fn simulate(state: &mut State, formula: &GpuFormula, tolerance: Option<f64>, step_size: Option<f64>,
                steps: Option<usize>, learning_rate: Option<f64>) -> Vec<bool> {
    let p = params(tolerance, step_size, steps, learning_rate);
    let mut a = vec![0u8; formula.varnum];
    check(unsafe {
        odesat_simulate(formula.h, state.v.as_mut_ptr(), state.xs.as_mut_ptr(), state.xl.as_mut_ptr(), &p,
                        a.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut())
    });
    a.into_iter().map(|x| x != 0).collect()
}

/// Drop-in for `system::simulate_inter` (system.rs:241-248): fixed step, or — `step_size = None` — adaptive
/// with ONE dt shared by the replicas, stepped one after the other as the reference does (system.rs:312-349).
pub fn simulate_inter(states: &mut Vec<State>, formula: &GpuFormula, tolerance: Option<f64>, step_size: Option<f64>,
                      steps: Option<usize>, learning_rate: Option<f64>) -> Vec<bool> {
    let (r, n, m) = (states.len(), formula.varnum, formula.n_clauses);
    let mut v: Vec<f64> = Vec::with_capacity(r * n);
    let mut xs: Vec<f64> = Vec::with_capacity(r * m);
    let mut xl: Vec<f64> = Vec::with_capacity(r * m);
    for s in states.iter() { v.extend(s.v.iter()); xs.extend(s.xs.iter()); xl.extend(s.xl.iter()); }
    let p = params(tolerance, step_size, steps, learning_rate);
    let mut a = vec![0u8; n];
    let (mut winner, mut taken) = (0i64, 0i64);
    check(unsafe { odesat_simulate_inter(formula.h, r as i64, v.as_mut_ptr(), xs.as_mut_ptr(), xl.as_mut_ptr(), &p, a.as_mut_ptr(), &mut winner, &mut taken) });
    for (k, s) in states.iter_mut().enumerate() {
        s.v.assign(&ndarray::ArrayView1::from(&v[k * n..(k + 1) * n]));
        s.xs.assign(&ndarray::ArrayView1::from(&xs[k * m..(k + 1) * m]));
        s.xl.assign(&ndarray::ArrayView1::from(&xl[k * m..(k + 1) * m]));
    }
    a.into_iter().map(|x| x != 0).collect()
}

/// The whole loop of `batch` (main.rs:278-308) as one call: states generated on the device from
/// `seed`, every replica verified exactly, returns (winner index or -1, its assignment).
pub fn batch(formula: &GpuFormula, batch_size: usize, seed: u64, tolerance: Option<f64>, step_size: Option<f64>,
             steps: usize, learning_rate: Option<f64>) -> (i64, Vec<bool>) {
    let p = params(tolerance, step_size, Some(steps), learning_rate);
    let mut a = vec![0u8; formula.varnum];
    let mut solved = vec![0i64; batch_size];
    let mut verified = vec![0u8; batch_size];
    let (mut winner, mut run) = (0i64, 0i64);
    check(unsafe {
        odesat_simulate_batch(formula.h, batch_size as i64, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(),
                              seed, 0, &p, 0, 0, solved.as_mut_ptr(), verified.as_mut_ptr(), &mut winner, a.as_mut_ptr(), &mut run)
    });
    (winner, a.into_iter().map(|x| x != 0).collect())
}
