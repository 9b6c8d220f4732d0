//! aleo-b200-sys: FFI declarations for libaleo_b200.so (include/aleo_b200.h) and the two call-site
//! replacements inside a patched snarkvm-algorithms 0.14.5.  NOT compiled in this environment (no
//! Rust toolchain); the layouts it assumes are asserted on the C++ side in include/aleo_b200.hpp.
use std::os::raw::{c_char, c_int, c_void};

extern "C" {
    pub fn aleo_b200_strerror(code: c_int) -> *const c_char;
    pub fn aleo_b200_ntt_fr(inout_host: *mut c_void, log_n: u32, direction: c_int, kind: c_int) -> c_int;
    pub fn aleo_b200_msm_g1(
        out_projective_host: *mut c_void,
        bases_host: *const c_void,
        n: usize,
        scalars_host: *const c_void,
        affine_stride: usize,
    ) -> c_int;
}

// the full argument lists of upstream's optional FFI (snarkvm_ntt with its `order`, snarkvm_polymul)
extern "C" {
    pub fn aleo_b200_ntt_fr_ordered(inout_host: *mut c_void, log_n: u32, direction: c_int, kind: c_int, order: c_int) -> c_int;
    pub fn aleo_b200_polymul(
        out_host: *mut c_void,
        pcount: usize,
        polynomials_host: *const *const c_void,
        plens: *const usize,
        ecount: usize,
        evaluations_host: *const *const c_void,
        elens: *const usize,
        log_n: u32,
    ) -> c_int;
}
pub const NTT_ORDER_II: c_int = 0; // upstream NN
pub const NTT_ORDER_IO: c_int = 1; // upstream NR
pub const NTT_ORDER_OI: c_int = 2; // upstream RN

/// Body of `PolyMultiplier::multiply` (src/fft/polynomial/multiplier.rs): coefficients of
/// `ifft(prod fft(p_i) * prod e_j)` over the domain of size `2^log_n`; every operand crosses PCIe once.
/// # Safety
/// `T` must be `Fp256<FrParameters>`; every evaluation vector must hold `2^log_n` elements.
pub unsafe fn polymul<T: Clone + Default>(polynomials: &[&[T]], evaluations: &[&[T]], log_n: u32) -> Vec<T> {
    assert_eq!(std::mem::size_of::<T>(), 32);
    let mut out = vec![T::default(); 1usize << log_n];
    let pp: Vec<*const c_void> = polynomials.iter().map(|p| p.as_ptr() as *const c_void).collect();
    let pl: Vec<usize> = polynomials.iter().map(|p| p.len()).collect();
    let ep: Vec<*const c_void> = evaluations.iter().map(|e| e.as_ptr() as *const c_void).collect();
    let el: Vec<usize> = evaluations.iter().map(|e| e.len()).collect();
    let rc = aleo_b200_polymul(out.as_mut_ptr() as *mut c_void, pp.len(), pp.as_ptr(), pl.as_ptr(), ep.len(), ep.as_ptr(), el.as_ptr(), log_n);
    if rc != 0 {
        die("PolyMultiplier::multiply", rc);
    }
    out
}

pub const NTT_FORWARD: c_int = 0;
pub const NTT_INVERSE: c_int = 1;
pub const NTT_STANDARD: c_int = 0;
pub const NTT_COSET: c_int = 1;

#[cold]
fn die(what: &str, rc: c_int) -> ! {
    let msg = unsafe { std::ffi::CStr::from_ptr(aleo_b200_strerror(rc)) }.to_string_lossy().into_owned();
    // north_star: no CPU fallback -- a failed device call is fatal
    panic!("{what}: aleo_b200 error {rc}: {msg}");
}

/// Body of `VariableBase::msm` for `G == bls12_377::G1Affine` (src/msm/variable_base/mod.rs):
/// ```ignore
/// pub fn msm<G: AffineCurve>(bases: &[G], scalars: &[<G::ScalarField as PrimeField>::BigInteger]) -> G::Projective {
///     if TypeId::of::<G>() == TypeId::of::<G1Affine>() {
///         return unsafe { aleo_b200_sys::msm_g1(bases, scalars) };
///     }
///     unimplemented!("only BLS12-377 G1 is on the proving path")
/// }
/// ```
/// # Safety
/// `A` must be the 104-byte `G1Affine`, `S` the 32-byte `BigInteger256`, `P` the 144-byte `G1Projective`.
pub unsafe fn msm_g1<A, S, P>(bases: &[A], scalars: &[S]) -> P {
    assert_eq!(std::mem::size_of::<A>(), 104);
    assert_eq!(std::mem::size_of::<S>(), 32);
    assert_eq!(std::mem::size_of::<P>(), 144);
    let n = bases.len().min(scalars.len());
    let mut out = std::mem::MaybeUninit::<P>::uninit();
    let rc = aleo_b200_msm_g1(
        out.as_mut_ptr() as *mut c_void,
        bases.as_ptr() as *const c_void,
        n,
        scalars.as_ptr() as *const c_void,
        std::mem::size_of::<A>(),
    );
    if rc != 0 {
        die("VariableBase::msm", rc);
    }
    out.assume_init()
}

/// Body of `EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place` when `size_of::<T>() == 32`
/// (src/fft/domain.rs): the caller has already done `x_s.resize(self.size(), T::zero())`.
/// # Safety
/// `T` must be `Fp256<FrParameters>` (32-byte Montgomery residue).
pub unsafe fn ntt_fr<T>(x_s: &mut [T], log_size_of_group: u32, direction: c_int, kind: c_int) {
    assert_eq!(std::mem::size_of::<T>(), 32);
    assert_eq!(x_s.len(), 1usize << log_size_of_group);
    let rc = aleo_b200_ntt_fr(x_s.as_mut_ptr() as *mut c_void, log_size_of_group, direction, kind);
    if rc != 0 {
        die("EvaluationDomain::fft", rc);
    }
}

// ---- resident SRS / KZG10 (src/polycommit/kzg10/mod.rs) -------------------------------------------------------------
extern "C" {
    pub fn aleo_b200_srs_create(handle_out: *mut *mut c_void, bases_host: *const c_void, n: usize, affine_stride: usize) -> c_int;
    pub fn aleo_b200_srs_destroy(handle: *mut c_void) -> c_int;
    pub fn aleo_b200_srs_msm(handle: *const c_void, out_projective_host: *mut c_void, scalars_host: *const c_void, n_used: usize) -> c_int;
    pub fn aleo_b200_kzg_commit(handle: *const c_void, out_compressed48_host: *mut c_void, coeffs_montgomery_host: *const c_void, n_coeffs: usize) -> c_int;
}

/// `Powers::powers_of_beta_g` (or `lagrange_basis_at_beta_g`) kept on the device: uploaded and expanded once, then every
/// `KZG10::commit` against it only moves the polynomial's coefficients.
pub struct SrsHandle(*mut c_void);
unsafe impl Send for SrsHandle {}
unsafe impl Sync for SrsHandle {} // the library is re-entrant; a handle is immutable after creation

impl SrsHandle {
    /// # Safety
    /// `A` must be the 104-byte `G1Affine`.
    pub unsafe fn new<A>(powers: &[A]) -> Self {
        assert_eq!(std::mem::size_of::<A>(), 104);
        let mut h: *mut c_void = std::ptr::null_mut();
        let rc = aleo_b200_srs_create(&mut h, powers.as_ptr() as *const c_void, powers.len(), 104);
        if rc != 0 {
            die("aleo_b200_srs_create", rc);
        }
        SrsHandle(h)
    }

    /// Non-hiding `KZG10::commit`: the polynomial's coefficients as it holds them (`Fp256`, Montgomery) -> the 48-byte
    /// compressed commitment (`G1Affine::deserialize_compressed` reads it back).
    /// # Safety
    /// `T` must be `Fp256<FrParameters>`.
    pub unsafe fn commit<T>(&self, coeffs: &[T]) -> [u8; 48] {
        assert_eq!(std::mem::size_of::<T>(), 32);
        let mut out = [0u8; 48];
        let rc = aleo_b200_kzg_commit(self.0, out.as_mut_ptr() as *mut c_void, coeffs.as_ptr() as *const c_void, coeffs.len());
        if rc != 0 {
            die("KZG10::commit", rc);
        }
        out
    }
}

impl Drop for SrsHandle {
    fn drop(&mut self) {
        unsafe { aleo_b200_srs_destroy(self.0) };
    }
}
