//! `krets-solver-b200`: the krets-solver analysis API on top of the krets_b200 CUDA library.
//!
//! Same seam as `krets_solver::solver::Solver` (solver.rs:17-52): `Solver::new(circuit, config).solve(analysis)`, same
//! `AnalysisResult` and `Error` types, so `krets-cli` switches backend by changing one `use` (main.rs:82-93).  The Rust
//! side keeps parsing, the MNA index map (parser.rs:105-125) and the sweep grids; everything numeric runs in
//! libkrets_b200.so (include/krets_b200.h).  SOURCE ONLY — see Cargo.toml.
mod ffi;

use faer::c64;
use ffi::*;
use krets_parser::analyses::{AcAnalysis, Analysis, DcAnalysis, TransientAnalysis};
use krets_parser::circuit::Circuit;
use krets_parser::elements::Element;
use krets_solver::AnalysisResult;
use krets_solver::config::SolverConfig;
use krets_solver::error::Error;
use std::collections::HashMap;
use std::ffi::CStr;
use std::ptr::{null, null_mut};

pub type Result<T> = std::result::Result<T, Error>;

pub struct Solver {
    circuit: Circuit,
    config: SolverConfig,
}

/// Owns a `krb_circuit` handle.
struct CircuitHandle(*mut krb_circuit);
impl Drop for CircuitHandle {
    fn drop(&mut self) {
        unsafe { krb_circuit_destroy(self.0) }
    }
}
/// Owns a `krb_plan` handle (one-time symbolic analysis: pattern, static pivots, executor tables).
struct PlanHandle(*mut krb_plan);
impl Drop for PlanHandle {
    fn drop(&mut self) {
        unsafe { krb_plan_destroy(self.0) }
    }
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(krb_last_error_string()).to_string_lossy().into_owned() }
}

/// Status code -> `krets_solver::Error` (error.rs:4-36), 1:1 as documented in the header.
fn check(rc: i32, max_iter: usize) -> Result<()> {
    match rc {
        KRB_OK => Ok(()),
        KRB_ERR_ELEMENT_NOT_FOUND => Err(Error::ElementNotFound(last_error())),
        KRB_ERR_INVALID_ELEMENT_FORMAT => Err(Error::InvalidElementFormat(last_error())),
        KRB_ERR_MAX_ITERATIONS => Err(Error::MaximumIterationsExceeded(max_iter)),
        KRB_ERR_DECOMPOSITION_FAILED => Err(Error::DecompositionFailed),
        _ => Err(Error::Unexpected(last_error())),
    }
}

/// The reference returns at the first failing point; lanes report their condition in `lane_status[]`.
fn first_failed_lane(status: &[i32], max_iter: usize) -> Result<()> {
    for &s in status {
        if s & KRB_LANE_ZERO_PIVOT != 0 {
            return Err(Error::DecompositionFailed); // op.rs:50-51, dc.rs:83-84, ac.rs:70, transient.rs:76-77
        }
        if s & KRB_LANE_NOT_CONVERGED != 0 {
            return Err(Error::MaximumIterationsExceeded(max_iter)); // op.rs:79-81
        }
        // inf/nan rows and lanes whose static pivots lost accuracy (and that the caller did not re-solve on a
        // KRB_EXEC_STAGED plan with refine_steps > 0, as the C++ host mirror does) are not returned as successes
        if s & (KRB_LANE_NONFINITE | KRB_LANE_INACCURATE) != 0 {
            return Err(Error::DecompositionFailed);
        }
    }
    Ok(())
}

impl Solver {
    pub const fn new(circuit: Circuit, config: SolverConfig) -> Self {
        Self { circuit, config }
    }

    fn idx(&self, key: String) -> i32 {
        self.circuit.index_map.get(&key).map(|&i| i as i32).unwrap_or(-1)
    }

    /// Flattens `Circuit.elements` (netlist order) into the POD table of krets_b200.h.
    fn element_table(&self) -> Result<Vec<krb_element>> {
        let mut table = Vec::with_capacity(self.circuit.elements.len());
        for e in &self.circuit.elements {
            let nodes = e.nodes();
            let mut pod = krb_element {
                kind: 0,
                ip: self.idx(format!("V({})", nodes[0])),
                im: self.idx(format!("V({})", nodes[1])),
                ic: self.idx(format!("I({})", e.identifier())),
                waveform: KRB_WAVE_NONE,
                flags: 0,
                value: 0.0,
                ac_amplitude: 0.0,
                wave: [0.0; 7],
                is: 1e-12,
                n: 1.0,
            };
            match e {
                Element::Resistor(r) => {
                    pod.kind = KRB_R;
                    pod.value = r.value;
                }
                Element::Capacitor(c) => {
                    pod.kind = KRB_C;
                    pod.value = c.value;
                    if c.g2 {
                        pod.flags |= KRB_ELEM_FLAG_G2;
                    }
                }
                Element::Inductor(l) => {
                    pod.kind = KRB_L;
                    pod.value = l.value;
                }
                Element::CurrentSource(i) => {
                    pod.kind = KRB_I;
                    pod.value = i.value;
                }
                Element::VoltageSource(v) => {
                    pod.kind = KRB_V;
                    pod.value = v.dc_value;
                    pod.ac_amplitude = v.ac_amplitude;
                    // transient_value_at prefers PULSE over SIN (voltage_source.rs:244-252)
                    if let Some(p) = &v.pulse {
                        pod.waveform = KRB_WAVE_PULSE;
                        pod.wave = [p.initial_value, p.pulsed_value, p.delay_time, p.rise_time, p.fall_time, p.pulse_width, p.period];
                    } else if let Some(s) = &v.sinusoidal {
                        pod.waveform = KRB_WAVE_SIN;
                        pod.wave = [s.offset, s.amplitude, s.frequency, s.delay, s.damping, s.phase, 0.0];
                    }
                }
                Element::Diode(d) => {
                    pod.kind = KRB_D;
                    pod.is = d.model.saturation_current;
                    pod.n = d.model.emission_coefficient;
                }
                Element::NMOSFET(m) => {
                    // drain / gate in ip / im; the reference's index lookups never match (stampable.rs:842-844),
                    // so the element stamps nothing in DC/transient and is rejected in AC, as upstream's todo!()
                    pod.kind = KRB_NMOS;
                    // source node and level-1 parameters travel with the element either way; only the feature
                    // `nmos-level1` (a krets build whose NMOS stamp is fixed) sets KRB_ELEM_FLAG_LEVEL1, which makes
                    // the backend stamp the companion model of nmosfet.rs:51-102 (krets_b200.h)
                    pod.ic = self.idx(format!("V({})", m.source));
                    pod.wave[0] = m.threshold_voltage();
                    pod.wave[1] = m.beta();
                    pod.wave[2] = m.lambda();
                    if cfg!(feature = "nmos-level1") {
                        pod.flags |= KRB_ELEM_FLAG_LEVEL1;
                    }
                }
                Element::BJT(_) => return Err(Error::Unexpected("BJT stamps are todo!() upstream (stampable.rs:327)".into())),
                Element::SubcktInstance(_) => return Err(Error::Unexpected("subcircuits are flattened by the parser (parser.rs:93-98)".into())),
            }
            table.push(pod);
        }
        Ok(table)
    }

    fn names_by_index(&self) -> Vec<&String> {
        let mut names: Vec<(&String, &usize)> = self.circuit.index_map.iter().collect();
        names.sort_by_key(|(_, i)| **i);
        names.into_iter().map(|(k, _)| k).collect()
    }

    fn create_circuit(&self) -> Result<CircuitHandle> {
        let table = self.element_table()?;
        // convergence_check picks the current tolerance for keys that start with 'I' (solver.rs:96-100)
        let is_current: Vec<u8> = self.names_by_index().iter().map(|k| k.starts_with('I') as u8).collect();
        let mut h: *mut krb_circuit = null_mut();
        check(
            unsafe { krb_circuit_create(table.as_ptr(), table.len() as i32, is_current.len() as i32, is_current.as_ptr(), &mut h) },
            self.config.maximum_iterations,
        )?;
        Ok(CircuitHandle(h))
    }

    fn analyze(&self, kc: &CircuitHandle, mode: i32, edit: impl FnOnce(&mut krb_plan_opts)) -> Result<PlanHandle> {
        let mut o: krb_plan_opts = unsafe { std::mem::zeroed() };
        unsafe { krb_default_plan_opts(&mut o) };
        edit(&mut o);
        let mut p: *mut krb_plan = null_mut();
        check(unsafe { krb_analyze(kc.0, mode, &o, &mut p) }, self.config.maximum_iterations)?;
        Ok(PlanHandle(p))
    }

    fn cfg(&self) -> krb_config {
        krb_config {
            relative_tolerance: self.config.relative_tolerance,
            current_absolute_tolerance: self.config.current_absolute_tolerance,
            voltage_absolute_tolerance: self.config.voltage_absolute_tolerance,
            maximum_iterations: self.config.maximum_iterations as i32,
        }
    }

    /// One result row per point: unknown name -> value, plus the pseudo-unknown the reference inserts
    /// ("step" dc.rs:97, "time" transient.rs:95).
    fn rows_f64(&self, x: &[f64], n: usize, axis_key: Option<&str>, axis: impl Fn(usize) -> f64) -> Vec<HashMap<String, f64>> {
        let names = self.names_by_index();
        x.chunks(n.max(1))
            .enumerate()
            .map(|(r, row)| {
                let mut m: HashMap<String, f64> = names.iter().zip(row).map(|(k, v)| ((*k).clone(), *v)).collect();
                if let Some(key) = axis_key {
                    m.insert(key.to_string(), axis(r));
                }
                m
            })
            .collect()
    }

    /// Same entry point as `krets_solver::solver::Solver::solve` (solver.rs:31-52).
    pub fn solve(&mut self, analysis: Analysis) -> Result<AnalysisResult> {
        let n = self.circuit.index_map.len();
        let max_iter = self.config.maximum_iterations;
        let kc = self.create_circuit()?;
        let cfg = self.cfg();
        match analysis {
            Analysis::Op => {
                let plan = self.analyze(&kc, KRB_MODE_DC, |_| {})?;
                let mut x = vec![0.0f64; n];
                let mut st = [0i32; 1];
                check(unsafe { krb_solve_op(plan.0, &cfg, 1, null(), x.as_mut_ptr(), null_mut(), st.as_mut_ptr(), null()) }, max_iter)?;
                first_failed_lane(&st, max_iter)?;
                Ok(AnalysisResult::Op(self.rows_f64(&x, n, None, |_| 0.0).remove(0))) // op.rs:60-63
            }
            Analysis::Dc(DcAnalysis { element, start, stop, step_size }) => {
                let sweep = self
                    .circuit
                    .elements
                    .iter()
                    .position(|x| x.identifier() == element)
                    .ok_or_else(|| Error::ElementNotFound(element.clone()))?; // dc.rs:21-28
                let num_steps = ((stop - start) / step_size).abs() as usize + 1; // dc.rs:44-45
                let values: Vec<f64> = (0..num_steps).map(|i| start + (i as f64 * step_size)).collect(); // dc.rs:48
                let plan = self.analyze(&kc, KRB_MODE_DC, |o| o.sweep_element = sweep as i32)?; // V / I check: dc.rs:31-39
                let mut x = vec![0.0f64; num_steps * n];
                if self.circuit.elements.iter().any(|e| e.is_nonlinear()) {
                    // dc.rs:58,112: Newton-Raphson at every point starts from the previous point's solution, so the points are
                    // sequential: one lane whose result rows are the sweep points (x[0][row][unknown] is already the result layout)
                    let mut st1 = [0i32; 1];
                    check(
                        unsafe {
                            krb_solve_dc_sweep_sequential(
                                plan.0,
                                &cfg,
                                start,
                                step_size,
                                num_steps as i64,
                                1,
                                null(),
                                x.as_mut_ptr(),
                                null_mut(),
                                st1.as_mut_ptr(),
                                null(),
                            )
                        },
                        max_iter,
                    )?;
                    first_failed_lane(&st1, max_iter)?;
                    return Ok(AnalysisResult::Dc(self.rows_f64(&x, n, Some("step"), |r| r as f64))); // dc.rs:92-97
                }
                // linear circuit: one solve per point, nothing carried between points — the points are parallel lanes
                let mut st = vec![0i32; num_steps];
                check(
                    unsafe { krb_solve_dc_sweep(plan.0, &cfg, values.as_ptr(), num_steps as i64, null(), x.as_mut_ptr(), st.as_mut_ptr(), null()) },
                    max_iter,
                )?;
                first_failed_lane(&st, max_iter)?;
                Ok(AnalysisResult::Dc(self.rows_f64(&x, n, Some("step"), |r| r as f64))) // dc.rs:92-97
            }
            Analysis::Ac(ac) => {
                // ac.rs:25: the DC operating point comes first (diode small-signal conductances, stampable.rs:614-641)
                let plan_dc = self.analyze(&kc, KRB_MODE_DC, |_| {})?;
                let mut op = vec![0.0f64; n];
                let mut st1 = [0i32; 1];
                check(unsafe { krb_solve_op(plan_dc.0, &cfg, 1, null(), op.as_mut_ptr(), null_mut(), st1.as_mut_ptr(), null()) }, max_iter)?;
                first_failed_lane(&st1, max_iter)?;
                let freqs: Vec<f64> = AcAnalysis::generate_frequencies(ac).into_iter().filter(|f| *f > 0.0).collect(); // ac.rs:40-45
                let fmin = freqs.first().copied().unwrap_or(1.0);
                // static pivots are chosen at the lowest frequency of the sweep (smallest |jwL| branch diagonals)
                let plan = self.analyze(&kc, KRB_MODE_AC, |o| o.nominal_frequency = fmin)?;
                let lanes = freqs.len();
                let mut x = vec![0.0f64; 2 * lanes * n];
                let mut st = vec![0i32; lanes];
                check(
                    unsafe {
                        krb_solve_ac(plan.0, &cfg, freqs.as_ptr(), lanes as i64, null(), op.as_ptr(), 1, x.as_mut_ptr(), st.as_mut_ptr(), null())
                    },
                    max_iter,
                )?;
                first_failed_lane(&st, max_iter)?;
                let names = self.names_by_index();
                let rows = x
                    .chunks(2 * n.max(1))
                    .zip(&freqs)
                    .map(|(row, f)| {
                        let mut m: HashMap<String, c64> =
                            names.iter().enumerate().map(|(k, nm)| ((*nm).clone(), c64::new(row[2 * k], row[2 * k + 1]))).collect();
                        m.insert("frequency".to_string(), c64::new(*f, 0.0)); // ac.rs:97-103
                        m
                    })
                    .collect();
                Ok(AnalysisResult::Ac(rows))
            }
            Analysis::Transient(TransientAnalysis { time_step, stop_time }) => {
                let n_steps = (stop_time / time_step).round() as usize; // transient.rs:29
                let plan_dc = self.analyze(&kc, KRB_MODE_DC, |_| {})?;
                let plan_tr = self.analyze(&kc, KRB_MODE_TRAN, |o| {
                    o.time_step = time_step;
                    o.integrator = KRB_INTEGRATOR_BE_KRETS; // backward Euler incl. the reference's source lag
                })?;
                let mut x = vec![0.0f64; (n_steps + 1) * n];
                let mut st = [0i32; 1];
                check(
                    unsafe { krb_solve_transient(plan_dc.0, plan_tr.0, &cfg, n_steps as i64, 1, null(), x.as_mut_ptr(), st.as_mut_ptr(), null()) },
                    max_iter,
                )?;
                first_failed_lane(&st, max_iter)?;
                Ok(AnalysisResult::Transient(self.rows_f64(&x, n, Some("time"), |r| r as f64 * time_step))) // transient.rs:44,95
            }
        }
    }
}
