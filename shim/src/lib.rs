//! zkb200-sys: the Rust side of the drop-in boundary (include/zkb200.h).
//!
//! `halo2curves::bn256::{Fr, Fq}` are `#[repr(transparent)]`-like wrappers over `[u64; 4]` holding Montgomery
//! limbs, `G1Affine { x, y }` is 64 bytes and `G1 { x, y, z }` 96 bytes, so slices of them are passed to C as plain
//! `*const u64` without any conversion.  Every wrapper keeps the exact signature of the halo2-axiom function it
//! replaces and panics on a non-zero status, matching the `.unwrap()` / `.expect()` convention of the callers
//! (/root/reference/aggregator/src/wrapper.rs:107-108,137).  There is no CPU fallback behind these calls.
//!
//! NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no cargo/rustc).  INTEGRATION.md shows the `[patch]` of
//! halo2-axiom that routes `best_multiexp` / `best_fft` / `EvaluationDomain` here.

use halo2curves::bn256::{Fr, G1Affine, G1};
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[allow(non_camel_case_types)]
type size_t = usize;

extern "C" {
    pub fn zkb_init(devices: *const c_int, ndev: c_int) -> c_int;
    pub fn zkb_shutdown();
    pub fn zkb_last_error() -> *const c_char;
    pub fn zkb_msm_g1(scalars: *const u64, bases: *const u64, n: size_t, out_jac: *mut u64) -> c_int;
    pub fn zkb_srs_register(bases: *const u64, n: size_t, handle: *mut u64) -> c_int;
    pub fn zkb_srs_release(handle: u64) -> c_int;
    pub fn zkb_msm_g1_srs(handle: u64, scalars: *const u64, n: size_t, out_jac: *mut u64) -> c_int;
    pub fn zkb_msm_g1_srs_batch(handle: u64, scalars: *const *const u64, ncols: size_t, n: size_t, out_jac: *mut u64) -> c_int;
    pub fn zkb_ntt_fr(a: *mut u64, omega: *const u64, log_n: u32) -> c_int;
    pub fn zkb_ntt_fr_batch(cols: *const *mut u64, ncols: size_t, omega: *const u64, log_n: u32) -> c_int;
    pub fn zkb_lagrange_to_coeff(a: *mut u64, k: u32) -> c_int;
    pub fn zkb_coeff_to_extended(input: *const u64, out: *mut u64, k: u32, extended_k: u32) -> c_int;
    pub fn zkb_extended_to_coeff(a: *mut u64, k: u32, extended_k: u32) -> c_int;
    pub fn zkb_lagrange_to_coeff_batch(cols: *const *mut u64, ncols: size_t, k: u32) -> c_int;
    pub fn zkb_coeff_to_extended_batch(input: *const *const u64, out: *const *mut u64, ncols: size_t, k: u32, extended_k: u32) -> c_int;
    pub fn zkb_g1_batch_normalize(points_jac: *const u64, n: size_t, out_affine: *mut u64) -> c_int;
    pub fn zkb_kzg_setup(k: u32, s: *const u64, g_out: *mut u64, g_lagrange_out: *mut u64) -> c_int;
    pub fn zkb_kzg_setup_resident(k: u32, s: *const u64, handle_g: *mut u64, handle_g_lagrange: *mut u64) -> c_int;
    pub fn zkb_srs_download(handle: u64, bases_out: *mut u64, n: size_t) -> c_int;
    pub fn zkb_host_register(ptr: *mut c_void, bytes: size_t) -> c_int;
    pub fn zkb_host_unregister(ptr: *mut c_void) -> c_int;
    pub fn zkb_graph_evaluate(graph: *const zkb_graph, inputs: *const zkb_graph_inputs, values: u64) -> c_int;
    pub fn zkb_msm_g1_srs_dev(handle: u64, offset: size_t, d_scalars: *const c_void, n: size_t, out_jac: *mut u64, stream: *mut c_void) -> c_int;
    pub fn zkb_fr_zeta(out: *mut u64) -> c_int;
    pub fn zkb_bound_devices(devices: *mut c_int, capacity: c_int) -> c_int;
    pub fn zkb_poly_upload(values: *const u64, n: size_t, handle: *mut u64) -> c_int;
    pub fn zkb_poly_load_file(path: *const c_char, offset: u64, n: size_t, handle: *mut u64) -> c_int;
    pub fn zkb_poly_write(handle: u64, offset: size_t, values: *const u64, n: size_t) -> c_int;
    pub fn zkb_poly_free(handle: u64) -> c_int;
    pub fn zkb_lookup_permute_expression_pair(input: u64, table: u64, usable_rows: size_t, permuted_input: *mut u64, permuted_table: *mut u64) -> c_int;
}

/// Call once before the first proof (e.g. from the patched `ParamsKZG::setup` / `read`).  `devices`: CUDA ordinals to drive from
/// this ONE process — the reference's prover is one process (`create_proof`, /root/reference/aggregator/src/wrapper.rs:129-137,
/// chained by `gen_recursion_snark`, wrapper.rs:869-902), so the 8 GPUs of a box are reached by listing them here (or by setting
/// ZKB_DEVICES=all and passing an empty slice): commits are then sharded by SRS point range, batches by column and a large
/// transform over NVLink peer memory behind the same wrappers below.
/// Also pins the one constant the two sides must agree on and that first-principles tests cannot see: the cube root of unity used
/// as the coset generator.  halo2curves releases have shipped either primitive root as `Fr::ZETA`; a mismatch would evaluate
/// `coeff_to_extended` on a different coset than the host-side pieces of the prover (valid-looking but unverifiable proofs).
pub fn init(devices: &[i32]) {
    use halo2curves::ff::WithSmallOrderMulGroup;
    let rc = unsafe { zkb_init(if devices.is_empty() { std::ptr::null() } else { devices.as_ptr() }, devices.len() as c_int) };
    check(rc, "init");
    let mut zeta = [0u64; 4];
    check(unsafe { zkb_fr_zeta(zeta.as_mut_ptr()) }, "fr_zeta");
    let crate_zeta: Fr = <Fr as WithSmallOrderMulGroup<3>>::ZETA;
    let crate_limbs: [u64; 4] = unsafe { std::mem::transmute(crate_zeta) };
    assert_eq!(zeta, crate_limbs, "zkb200 and halo2curves disagree on Fr::ZETA (coset generator of the extended domain)");
}

/// `plonk::evaluation::ValueSource` / `Calculation` / `GraphEvaluator` as the C ABI reads them (include/zkb200.h).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct zkb_value_source {
    pub kind: u32,     // ZKB_SRC_*: 0 Constant 1 Intermediate 2 Fixed 3 Advice 4 Instance 5 Challenge 6 Beta 7 Gamma 8 Theta 9 Y 10 PreviousValue
    pub index: u32,
    pub rotation: u32, // index into `rotations`
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct zkb_calculation {
    pub op: u32,       // ZKB_CALC_*: 0 Add 1 Sub 2 Mul 3 Square 4 Double 5 Negate 6 Store 7 MulAdd (one Horner step)
    pub target: u32,
    pub a: zkb_value_source,
    pub b: zkb_value_source,
    pub c: zkb_value_source,
}
#[repr(C)]
pub struct zkb_graph {
    pub calculations: *const zkb_calculation,
    pub num_calculations: size_t,
    pub num_intermediates: u32,
    pub constants: *const u64,
    pub num_constants: size_t,
    pub rotations: *const i32,
    pub num_rotations: size_t,
}
#[repr(C)]
pub struct zkb_graph_inputs {
    pub fixed: *const u64,
    pub num_fixed: size_t,
    pub advice: *const u64,
    pub num_advice: size_t,
    pub instance: *const u64,
    pub num_instance: size_t,
    pub challenges: *const u64,
    pub num_challenges: size_t,
    pub beta: *const u64,
    pub gamma: *const u64,
    pub theta: *const u64,
    pub y: *const u64,
    pub rot_scale: i32,
}

#[inline]
fn check(rc: c_int, what: &str) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(zkb_last_error()) }.to_string_lossy().into_owned();
        panic!("zkb200 {what} failed ({rc}): {msg}");
    }
}

const _: () = assert!(std::mem::size_of::<Fr>() == 32);
const _: () = assert!(std::mem::size_of::<G1Affine>() == 64);
const _: () = assert!(std::mem::size_of::<G1>() == 96);

/// Drop-in body for `halo2_proofs::arithmetic::best_multiexp::<G1Affine>`.
pub fn best_multiexp(coeffs: &[Fr], bases: &[G1Affine]) -> G1 {
    assert_eq!(coeffs.len(), bases.len());
    let mut out = std::mem::MaybeUninit::<G1>::uninit();
    let rc = unsafe { zkb_msm_g1(coeffs.as_ptr() as *const u64, bases.as_ptr() as *const u64, coeffs.len(), out.as_mut_ptr() as *mut u64) };
    check(rc, "best_multiexp");
    unsafe { out.assume_init() }
}

/// Drop-in body for `halo2_proofs::arithmetic::best_fft::<Fr, Fr>`.
pub fn best_fft(a: &mut [Fr], omega: Fr, log_n: u32) {
    assert_eq!(a.len(), 1 << log_n);
    let rc = unsafe { zkb_ntt_fr(a.as_mut_ptr() as *mut u64, &omega as *const Fr as *const u64, log_n) };
    check(rc, "best_fft");
}

/// An SRS (`ParamsKZG.g` or `ParamsKZG.g_lagrange`) resident in HBM for the life of the params.
pub struct ResidentSrs {
    handle: u64,
    len: usize,
}

impl ResidentSrs {
    pub fn new(bases: &[G1Affine]) -> Self {
        let mut handle = 0u64;
        check(unsafe { zkb_srs_register(bases.as_ptr() as *const u64, bases.len(), &mut handle) }, "srs_register");
        Self { handle, len: bases.len() }
    }
    /// `ParamsKZG::commit` / `commit_lagrange`: `best_multiexp(&poly, &bases[..poly.len()])`.
    pub fn commit(&self, poly: &[Fr]) -> G1 {
        assert!(poly.len() <= self.len);
        let mut out = std::mem::MaybeUninit::<G1>::uninit();
        check(unsafe { zkb_msm_g1_srs(self.handle, poly.as_ptr() as *const u64, poly.len(), out.as_mut_ptr() as *mut u64) }, "commit");
        unsafe { out.assume_init() }
    }
    /// The prover's per-column commit loop in one call.
    pub fn commit_batch(&self, polys: &[&[Fr]]) -> Vec<G1> {
        if polys.is_empty() {
            return vec![];
        }
        let n = polys[0].len();
        assert!(polys.iter().all(|p| p.len() == n) && n <= self.len);
        let ptrs: Vec<*const u64> = polys.iter().map(|p| p.as_ptr() as *const u64).collect();
        let mut out = Vec::<G1>::with_capacity(polys.len());
        check(unsafe { zkb_msm_g1_srs_batch(self.handle, ptrs.as_ptr(), polys.len(), n, out.as_mut_ptr() as *mut u64) }, "commit_batch");
        unsafe { out.set_len(polys.len()) };
        out
    }
}

impl Drop for ResidentSrs {
    fn drop(&mut self) {
        unsafe { zkb_srs_release(self.handle) };
    }
}

/// `EvaluationDomain::lagrange_to_coeff` body (in place on the polynomial's values).
pub fn lagrange_to_coeff(values: &mut [Fr], k: u32) {
    assert_eq!(values.len(), 1 << k);
    check(unsafe { zkb_lagrange_to_coeff(values.as_mut_ptr() as *mut u64, k) }, "lagrange_to_coeff");
}

/// `EvaluationDomain::coeff_to_extended` body: returns the 2^extended_k coset evaluations.
pub fn coeff_to_extended(values: &[Fr], k: u32, extended_k: u32) -> Vec<Fr> {
    assert_eq!(values.len(), 1 << k);
    let mut out = Vec::<Fr>::with_capacity(1 << extended_k);
    check(unsafe { zkb_coeff_to_extended(values.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64, k, extended_k) }, "coeff_to_extended");
    unsafe { out.set_len(1 << extended_k) };
    out
}

/// `EvaluationDomain::extended_to_coeff` body: in place, then truncated to n * quotient_poly_degree.
pub fn extended_to_coeff(mut values: Vec<Fr>, k: u32, extended_k: u32, quotient_poly_degree: u64) -> Vec<Fr> {
    assert_eq!(values.len(), 1 << extended_k);
    check(unsafe { zkb_extended_to_coeff(values.as_mut_ptr() as *mut u64, k, extended_k) }, "extended_to_coeff");
    values.truncate(((1u64 << k) * quotient_poly_degree) as usize);
    values
}

/// The evaluate_h loop of `create_proof`: `coeff_to_extended` of many polynomials in one call, so that the upload of
/// polynomial i+1, the coset NTT of polynomial i and the download of polynomial i-1 overlap (three-stream scheduler).
pub fn coeff_to_extended_batch(polys: &[&[Fr]], k: u32, extended_k: u32) -> Vec<Vec<Fr>> {
    let n_ext = 1usize << extended_k;
    assert!(polys.iter().all(|p| p.len() == 1 << k));
    let mut outs: Vec<Vec<Fr>> = polys.iter().map(|_| Vec::with_capacity(n_ext)).collect();
    let ins: Vec<*const u64> = polys.iter().map(|p| p.as_ptr() as *const u64).collect();
    let ptrs: Vec<*mut u64> = outs.iter_mut().map(|o| o.as_mut_ptr() as *mut u64).collect();
    check(unsafe { zkb_coeff_to_extended_batch(ins.as_ptr(), ptrs.as_ptr(), polys.len(), k, extended_k) }, "coeff_to_extended_batch");
    for o in outs.iter_mut() {
        unsafe { o.set_len(n_ext) };
    }
    outs
}

/// Drop-in body for `halo2curves::group::Curve::batch_normalize(&[G1], &mut [G1Affine])`.
pub fn batch_normalize(p: &[G1], q: &mut [G1Affine]) {
    assert_eq!(p.len(), q.len());
    check(unsafe { zkb_g1_batch_normalize(p.as_ptr() as *const u64, p.len(), q.as_mut_ptr() as *mut u64) }, "batch_normalize");
}

/// The G1 side of `ParamsKZG::<Bn256>::setup(k, rng)`: the caller samples `s` exactly as upstream does
/// (`let s = <E::Scalar>::random(rng);`), the arrays are generated in HBM and registered as resident SRS handles.
pub fn kzg_setup_resident(k: u32, s: Fr) -> (ResidentSrs, ResidentSrs) {
    let (mut hg, mut hgl) = (0u64, 0u64);
    check(unsafe { zkb_kzg_setup_resident(k, &s as *const Fr as *const u64, &mut hg, &mut hgl) }, "ParamsKZG::setup");
    (ResidentSrs { handle: hg, len: 1 << k }, ResidentSrs { handle: hgl, len: 1 << k })
}

impl ResidentSrs {
    /// Copy the resident bases back (`ParamsKZG::write`, `get_g()`).
    pub fn download(&self) -> Vec<G1Affine> {
        let mut out = Vec::<G1Affine>::with_capacity(self.len);
        check(unsafe { zkb_srs_download(self.handle, out.as_mut_ptr() as *mut u64, self.len) }, "srs_download");
        unsafe { out.set_len(self.len) };
        out
    }
}

/// Page-locks a long-lived `Vec<Fr>` (advice column, extended polynomial) for its lifetime so the transfer scheduler
/// DMAs it directly instead of staging it through pinned buffers.
pub struct PinnedVec {
    pub values: Vec<Fr>,
}

impl PinnedVec {
    pub fn new(mut values: Vec<Fr>) -> Self {
        check(unsafe { zkb_host_register(values.as_mut_ptr() as *mut c_void, values.len() * 32) }, "host_register");
        Self { values }
    }
}

impl Drop for PinnedVec {
    fn drop(&mut self) {
        unsafe { zkb_host_unregister(self.values.as_mut_ptr() as *mut c_void) };
    }
}

/// The row loop of `evaluate_h` over a compiled `GraphEvaluator`, on extended cosets resident in HBM (polynomial handles):
/// `values[idx] = graph.evaluate(.., &values[idx], idx, rot_scale, isize)` for every row.
#[allow(clippy::too_many_arguments)]
pub fn graph_evaluate(graph: &zkb_graph, fixed: &[u64], advice: &[u64], instance: &[u64], challenges: &[Fr], beta: &Fr, gamma: &Fr,
                      theta: &Fr, y: &Fr, rot_scale: i32, values: u64) {
    let inputs = zkb_graph_inputs {
        fixed: fixed.as_ptr(), num_fixed: fixed.len(), advice: advice.as_ptr(), num_advice: advice.len(),
        instance: instance.as_ptr(), num_instance: instance.len(),
        challenges: challenges.as_ptr() as *const u64, num_challenges: challenges.len(),
        beta: beta as *const Fr as *const u64, gamma: gamma as *const Fr as *const u64,
        theta: theta as *const Fr as *const u64, y: y as *const Fr as *const u64, rot_scale,
    };
    check(unsafe { zkb_graph_evaluate(graph, &inputs, values) }, "graph_evaluate");
}
