//! `QEqSolver` with cheq's signature, backed by libcheq_b200.so.
//!
//! Mirrors /root/reference/src/solver/implementation.rs:60-266: the NoAtoms check (:180-182), `fetch_element_data`
//! (:348-362), the point-charge element lookup (:378-387) and the `ExternalPotential::is_empty` short-circuit
//! (:245-247) stay in Rust; everything numerical is one FFI call.  NOT compiled in the build container.
use cheq::{AtomView, CalculationResult, CheqError, ExternalPotential, Parameters, SolverOptions};
use cheq::{BasisType, DampingStrategy};
use cheq_b200_sys as sys;
use std::sync::Mutex;

pub struct QEqSolver<'p> {
    parameters: &'p Parameters,
    options: SolverOptions,
    ctx: Mutex<*mut sys::cheq_ctx>, // one CUDA stream + workspace; calls on one solver are serialised
}
unsafe impl<'p> Send for QEqSolver<'p> {}
unsafe impl<'p> Sync for QEqSolver<'p> {}

impl<'p> QEqSolver<'p> {
    pub fn new(parameters: &'p Parameters) -> Self {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { sys::cheq_ctx_create(0, &mut ctx) };
        assert_eq!(rc, sys::CHEQ_OK, "cheq-b200 needs a CUDA device (no CPU fallback)");
        Self { parameters, options: SolverOptions::default(), ctx: Mutex::new(ctx) }
    }

    pub fn with_options(mut self, options: SolverOptions) -> Self {
        self.options = options;
        self
    }

    pub fn solve<A: AtomView>(&self, atoms: &[A], total_charge: f64) -> Result<CalculationResult, CheqError> {
        self.solve_impl(atoms, total_charge, None)
    }

    pub fn solve_in_field<A: AtomView>(
        &self, atoms: &[A], total_charge: f64, external: &ExternalPotential,
    ) -> Result<CalculationResult, CheqError> {
        if external.is_empty() {
            return self.solve_impl(atoms, total_charge, None);
        }
        self.solve_impl(atoms, total_charge, Some(external))
    }

    fn solve_impl<A: AtomView>(
        &self, atoms: &[A], total_charge: f64, external: Option<&ExternalPotential>,
    ) -> Result<CalculationResult, CheqError> {
        let n = atoms.len();
        if n == 0 {
            return Err(CheqError::NoAtoms);
        }
        let (mut xyz, mut chi, mut hard, mut rad, mut nq) =
            (Vec::with_capacity(3 * n), Vec::with_capacity(n), Vec::with_capacity(n), Vec::with_capacity(n), Vec::with_capacity(n));
        for a in atoms {
            let z = a.atomic_number();
            let d = self.parameters.elements.get(&z).ok_or(CheqError::ParameterNotFound(z))?;
            xyz.extend_from_slice(&a.position());
            chi.push(d.electronegativity);
            hard.push(d.hardness);
            rad.push(d.radius);
            nq.push(d.principal_quantum_number);
        }
        let (mut pxyz, mut pq, mut prad, mut pn) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        let ext = match external {
            None => None,
            Some(e) => {
                for pc in e.point_charges() {
                    let d = self.parameters.elements.get(&pc.atomic_number)
                        .ok_or(CheqError::ParameterNotFound(pc.atomic_number))?;
                    pxyz.extend_from_slice(&pc.position);
                    pq.push(pc.charge);
                    prad.push(d.radius);
                    pn.push(d.principal_quantum_number);
                }
                Some(sys::cheq_external {
                    num_point_charges: pq.len() as u64, xyz: pxyz.as_ptr(), charge: pq.as_ptr(),
                    radius: prad.as_ptr(), n: pn.as_ptr(), field: e.uniform_field(),
                })
            }
        };
        let (kind, value) = match self.options.damping {
            DampingStrategy::None => (0u8, 1.0),
            DampingStrategy::Fixed(d) => (1, d),
            DampingStrategy::Auto { initial } => (2, initial),
        };
        let opt = sys::cheq_options {
            tolerance: self.options.tolerance, lambda_scale: self.options.lambda_scale, damping_value: value,
            max_iterations: self.options.max_iterations, hydrogen_scf: self.options.hydrogen_scf as u8,
            basis: matches!(self.options.basis_type, BasisType::Sto) as u8, damping_kind: kind, reserved: 0,
        };
        let mut charges = vec![0.0f64; n];
        let (mut mu, mut its, mut delta) = (0.0f64, 0u32, 0.0f64);
        let ctx = self.ctx.lock().unwrap();
        let rc = unsafe {
            sys::cheq_solve(*ctx, n as u64, xyz.as_ptr(), chi.as_ptr(), hard.as_ptr(), rad.as_ptr(), nq.as_ptr(),
                total_charge, &opt, ext.as_ref().map_or(std::ptr::null(), |e| e as *const _),
                charges.as_mut_ptr(), &mut mu, &mut its, &mut delta)
        };
        match rc {
            sys::CHEQ_OK => Ok(CalculationResult { charges, equilibrated_potential: mu, iterations: its }),
            sys::CHEQ_ERR_NO_ATOMS => Err(CheqError::NoAtoms),
            sys::CHEQ_ERR_NOT_CONVERGED => Err(CheqError::NotConverged { max_iterations: self.options.max_iterations, delta }),
            _ => {
                let msg = unsafe { std::ffi::CStr::from_ptr(sys::cheq_last_error(*ctx)) }.to_string_lossy().into_owned();
                Err(CheqError::LinalgError(msg))
            }
        }
    }
}

impl<'p> Drop for QEqSolver<'p> {
    fn drop(&mut self) {
        unsafe { sys::cheq_ctx_destroy(*self.ctx.lock().unwrap()) }
    }
}
