//! Safe wrappers over `libbpgpu`, shaped like the call sites of renegade-fi/mpc-bulletproof that
//! they replace.  Scalars cross the boundary as 32-byte little-endian canonical encodings and
//! points as 32-byte compressed encodings; the two traits below are what the host crate's
//! `Scalar` / point types must provide (for curve25519-dalek: `Scalar::to_bytes`,
//! `RistrettoPoint::compress().to_bytes()`, `CompressedRistretto::decompress`).
//!
//! Not compiled in the environment this was written in (no Rust toolchain); it mirrors
//! `mpc_bulletproof_b200/csrc/host/protocol.cpp`, which is the compiled and tested host side.

use bpgpu_sys as sys;
use std::ffi::CStr;
use std::os::raw::c_int;
use std::ptr;

/// `R1CSError` / `ProofError` as the reference names them (src/errors.rs:13-55, 150-177).
#[derive(Debug, Clone, PartialEq, Eq)]
pub enum Error {
    /// `R1CSError::InvalidGeneratorsLength`
    InvalidGeneratorsLength,
    /// `ProofError::FormatError` / `R1CSError::FormatError`
    FormatError,
    /// `ProofError::VerificationError` / `R1CSError::VerificationError`
    VerificationError,
    /// lengths differ / not a power of two: the reference panics (`assert!`), src/inner_product_proof.rs:59-70
    InvalidInput(&'static str),
    /// CUDA failure or no device; there is no CPU path
    Device(String),
}

fn check(code: c_int) -> Result<(), Error> {
    match code {
        sys::BPG_OK => Ok(()),
        sys::BPG_ERR_CAPACITY => Err(Error::InvalidGeneratorsLength),
        sys::BPG_ERR_DECODE => Err(Error::FormatError),
        sys::BPG_ERR_VERIFY => Err(Error::VerificationError),
        sys::BPG_ERR_LEN => Err(Error::InvalidInput("vector lengths differ")),
        sys::BPG_ERR_POW2 => Err(Error::InvalidInput("length is not a power of two")),
        sys::BPG_ERR_ARG => Err(Error::InvalidInput("bad argument")),
        other => {
            let msg = unsafe { CStr::from_ptr(sys::bpg_strerror(other)) }.to_string_lossy().into_owned();
            Err(Error::Device(msg))
        }
    }
}

/// 32-byte little-endian canonical scalar (mod l).
pub trait ScalarBytes {
    fn to_le_bytes(&self) -> [u8; 32];
}
/// 32-byte compressed group element.
pub trait PointBytes: Sized {
    fn to_compressed(&self) -> [u8; 32];
    fn from_compressed(bytes: &[u8; 32]) -> Option<Self>;
}

/// Raw encodings are points too (generators that were derived on the device never leave byte form).
impl PointBytes for [u8; 32] {
    fn to_compressed(&self) -> [u8; 32] { *self }
    fn from_compressed(bytes: &[u8; 32]) -> Option<Self> { Some(*bytes) }
}

/// One GPU, one proving thread (`&mut Transcript` exclusivity, src/r1cs/prover.rs:27-28).
pub struct Context {
    raw: *mut sys::bpg_ctx,
}
// a context may move between threads but is used by one at a time
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::bpg_init(device, &mut raw) })?;
        Ok(Context { raw })
    }

    /// Drop-in for `StarkPoint::msm_iter(scalars, points)`
    /// (src/inner_product_proof.rs:90,103,159,166,353; src/r1cs/prover.rs:465-555; src/r1cs/verifier.rs:516).
    pub fn msm_iter<S, P, I, J>(&self, scalars: I, points: J) -> Result<P, Error>
    where
        S: ScalarBytes,
        P: PointBytes,
        I: IntoIterator<Item = S>,
        J: IntoIterator<Item = P>,
    {
        let s: Vec<u8> = scalars.into_iter().flat_map(|x| x.to_le_bytes()).collect();
        let p: Vec<u8> = points.into_iter().flat_map(|x| x.to_compressed()).collect();
        if s.len() != p.len() {
            return Err(Error::InvalidInput("vector lengths differ"));
        }
        let mut out = [0u8; 32];
        check(unsafe { sys::bpg_msm(self.raw, s.as_ptr(), p.as_ptr(), s.len() / 32, out.as_mut_ptr()) })?;
        P::from_compressed(&out).ok_or(Error::FormatError)
    }

    /// MPC open (src/r1cs_mpc/mpc_prover.rs:630-657): the parties' partial sums (n_sets x 128 bytes
    /// each, laid out [party][set]) have been exchanged over the fabric; add and encode.
    pub fn open_partials(&self, parts: &[u8], n_parties: usize, n_sets: usize) -> Result<Vec<[u8; 32]>, Error> {
        if parts.len() != n_parties * n_sets * 128 {
            return Err(Error::InvalidInput("partials buffer length"));
        }
        let mut out = vec![[0u8; 32]; n_sets];
        check(unsafe {
            sys::bpg_sum_encode(self.raw, parts.as_ptr(), n_parties as c_int, n_sets as c_int, out.as_mut_ptr() as *mut u8)
        })?;
        Ok(out)
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::bpg_free(self.raw) }
    }
}

/// Generators resident in HBM, owned next to `BulletproofGens` / `PedersenGens`
/// (src/generators.rs:32-71, 158-235): one windowed table [G (cap) | H (cap) | B | B_blinding]
/// and a fixed-base comb for (B, B_blinding).  Built once per capacity.
pub struct GpuGens<'c> {
    ctx: &'c Context,
    table: *mut sys::bpg_table,
    comb: *mut sys::bpg_comb,
    pub capacity: usize,
}

impl<'c> GpuGens<'c> {
    pub fn new<P: PointBytes>(ctx: &'c Context, g: &[P], h: &[P], b: &P, b_blinding: &P) -> Result<Self, Error> {
        if g.len() != h.len() {
            return Err(Error::InvalidInput("vector lengths differ"));
        }
        let cap = g.len();
        let mut all = Vec::with_capacity((2 * cap + 2) * 32);
        for p in g.iter().chain(h.iter()) {
            all.extend_from_slice(&p.to_compressed());
        }
        all.extend_from_slice(&b.to_compressed());
        all.extend_from_slice(&b_blinding.to_compressed());
        let mut table = ptr::null_mut();
        check(unsafe { sys::bpg_table_upload(ctx.raw, all.as_ptr(), 2 * cap + 2, &mut table) })?;
        let mut gens = GpuGens { ctx, table, comb: ptr::null_mut(), capacity: cap };
        check(unsafe { sys::bpg_table_set_windows(ctx.raw, gens.table, 0) })?;
        let bases = &all[2 * cap * 32..];
        check(unsafe { sys::bpg_comb_create(ctx.raw, bases.as_ptr(), 2, &mut gens.comb) })?;
        Ok(gens)
    }
    /// `BulletproofGens::new(capacity, ..).share(party)` + `PedersenGens::default()` with the chains
    /// derived on the device (src/generators.rs:61-71, 80-125, 182-235): labels "G"/"H" || u32le(party).
    /// Returns the resident generators and the compressed (G, H) for the host-side structs.
    pub fn derive(ctx: &'c Context, capacity: usize, party: u32, b: &[u8; 32], b_blinding: &[u8; 32]) -> Result<(Self, Vec<[u8; 32]>, Vec<[u8; 32]>), Error> {
        let mut g = vec![[0u8; 32]; capacity];
        let mut h = vec![[0u8; 32]; capacity];
        for (tag, out) in [(b'G', &mut g), (b'H', &mut h)] {
            let mut label = vec![tag];
            label.extend_from_slice(&party.to_le_bytes());
            check(unsafe { sys::bpg_gens_chain(ctx.raw, label.as_ptr(), label.len(), 0, capacity, out.as_mut_ptr() as *mut u8) })?;
        }
        let gens = Self::new(ctx, &g, &h, b, b_blinding)?;
        Ok((gens, g, h))
    }
    pub fn g_base(&self) -> usize { 0 }
    pub fn h_base(&self) -> usize { self.capacity }
    pub fn b_id(&self) -> usize { 2 * self.capacity }
    pub fn b_blinding_id(&self) -> usize { 2 * self.capacity + 1 }

    /// `PedersenGens::commit(value, blinding)` batched (src/generators.rs:41-43): V_j, T_1..T_6.
    pub fn commit_batch<S: ScalarBytes>(&self, values: &[S], blindings: &[S]) -> Result<Vec<[u8; 32]>, Error> {
        if values.len() != blindings.len() {
            return Err(Error::InvalidInput("vector lengths differ"));
        }
        let n = values.len();
        let sc: Vec<u8> = values.iter().chain(blindings.iter()).flat_map(|x| x.to_le_bytes()).collect();
        let mut out = vec![[0u8; 32]; n];
        check(unsafe { sys::bpg_comb_mul(self.ctx.raw, self.comb, sc.as_ptr(), n, out.as_mut_ptr() as *mut u8) })?;
        Ok(out)
    }

    /// One party's share commitments over the first `n` G generators and the blinding base:
    /// `msm_authenticated_iter` on the share and MAC vectors (src/r1cs_mpc/mpc_prover.rs:621-657);
    /// returns n_sets x 128 bytes to be exchanged and passed to `Context::open_partials`.
    pub fn share_partial(&self, scalars_set_major: &[u8], n: usize, n_sets: usize) -> Result<Vec<u8>, Error> {
        if scalars_set_major.len() != n * n_sets * 32 {
            return Err(Error::InvalidInput("scalar buffer length"));
        }
        let mut out = vec![0u8; n_sets * 128];
        check(unsafe {
            sys::bpg_msm_table_partial(self.ctx.raw, self.table, 0, n, scalars_set_major.as_ptr(), n_sets as c_int, out.as_mut_ptr())
        })?;
        Ok(out)
    }
}

impl Drop for GpuGens<'_> {
    fn drop(&mut self) {
        unsafe {
            if !self.comb.is_null() { sys::bpg_comb_free(self.comb) }
            sys::bpg_table_free(self.table)
        }
    }
}

/// The round loop of `InnerProductProof::create` (src/inner_product_proof.rs:49-193) with the
/// vectors resident in HBM.  The transcript stays with the caller:
///
/// ```ignore
/// transcript.innerproduct_domain_sep(n as u64);                          // :72
/// let mut s = IppSession::begin_shared(&gens, n, &w, &g_factors, &h_factors, &a, &b)?;
/// while s.rounds_left() > 0 {
///     let (l, r) = s.round_lr()?;                                          // :87-114 / :156-172
///     transcript.append_point(b"L", &l); transcript.append_point(b"R", &r);   // :119-120
///     let u = transcript.challenge_scalar(b"u");                           // :122
///     s.round_fold(&u, &u.invert())?;                                      // fold_witness :202-248
///     l_vec.push(l); r_vec.push(r);
/// }
/// let (a, b) = s.finish()?;                                                // :187-192
/// ```
pub struct IppSession<'g> {
    raw: *mut sys::bpg_ipp,
    _gens: std::marker::PhantomData<&'g ()>,
}

impl<'g> IppSession<'g> {
    /// Generators and the base of Q live in the shared generator table: Q = q_mul * B
    /// (the R1CS prover's Q = w * B, src/r1cs/prover.rs:687).
    pub fn begin_shared<S: ScalarBytes>(
        gens: &'g GpuGens<'_>, n: usize, q_mul: &S, g_factors: &[S], h_factors: &[S], a: &[S], b: &[S],
    ) -> Result<Self, Error> {
        if g_factors.len() != n || h_factors.len() != n || a.len() != n || b.len() != n {
            return Err(Error::InvalidInput("vector lengths differ"));
        }
        let ser = |v: &[S]| -> Vec<u8> { v.iter().flat_map(|x| x.to_le_bytes()).collect() };
        let (gf, hf, av, bv) = (ser(g_factors), ser(h_factors), ser(a), ser(b));
        let mut raw = ptr::null_mut();
        check(unsafe {
            sys::bpg_ipp_begin_shared(
                gens.ctx.raw, gens.table, gens.g_base(), gens.h_base(), gens.b_id(), q_mul.to_le_bytes().as_ptr(), n,
                gf.as_ptr(), hf.as_ptr(), av.as_ptr(), bv.as_ptr(), &mut raw,
            )
        })?;
        Ok(IppSession { raw, _gens: std::marker::PhantomData })
    }
    pub(crate) fn from_raw(raw: *mut sys::bpg_ipp) -> Self {
        IppSession { raw, _gens: std::marker::PhantomData }
    }
    pub fn rounds_left(&self) -> usize {
        unsafe { sys::bpg_ipp_rounds_left(self.raw) }
    }
    pub fn round_lr(&mut self) -> Result<([u8; 32], [u8; 32]), Error> {
        let (mut l, mut r) = ([0u8; 32], [0u8; 32]);
        check(unsafe { sys::bpg_ipp_round_LR(self.raw, l.as_mut_ptr(), r.as_mut_ptr()) })?;
        Ok((l, r))
    }
    pub fn round_fold<S: ScalarBytes>(&mut self, u: &S, u_inv: &S) -> Result<(), Error> {
        check(unsafe { sys::bpg_ipp_round_fold(self.raw, u.to_le_bytes().as_ptr(), u_inv.to_le_bytes().as_ptr()) })
    }
    pub fn finish(self) -> Result<([u8; 32], [u8; 32]), Error> {
        let (mut a, mut b) = ([0u8; 32], [0u8; 32]);
        check(unsafe { sys::bpg_ipp_finish(self.raw, a.as_mut_ptr(), b.as_mut_ptr()) })?;
        Ok((a, b))
    }
}

impl Drop for IppSession<'_> {
    fn drop(&mut self) {
        unsafe { sys::bpg_ipp_free(self.raw) }
    }
}

/// The O(n) scalar vectors of `Prover::prove` / `Verifier::verify` kept in HBM
/// (src/r1cs/prover.rs:342-379, 465-494, 589-708; src/r1cs/verifier.rs:323-362, 468-547).
/// Vectors are passed as Montgomery limbs (`[u32; 8]` per scalar, x * 2^256 mod l), which is what
/// a 4x64-limb Montgomery `Scalar` already holds in memory.
pub struct R1csDevice<'g> {
    raw: *mut sys::bpg_r1cs_dev,
    gens: &'g GpuGens<'g>,
}

impl<'g> R1csDevice<'g> {
    pub fn new(gens: &'g GpuGens<'g>, capacity: usize) -> Result<Self, Error> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::bpg_r1cs_dev_new(gens.ctx.raw, capacity.max(1), &mut raw) })?;
        Ok(R1csDevice { raw, gens })
    }
    /// second-phase multipliers appeared (src/r1cs/prover.rs:501-530)
    pub fn reserve(&mut self, capacity: usize) -> Result<(), Error> {
        check(unsafe { sys::bpg_r1cs_dev_reserve(&mut self.raw, capacity) })
    }
    /// (A_I, A_O, S) over gens[first .. first + a_l.len()); s_L, s_R are generated on the device
    /// from `vec_key` (one draw of the prover's RNG).  blind3 = i_blinding | o_blinding | s_blinding.
    pub fn commit_phase(
        &mut self, first: usize, a_l: &[sys::Mont], a_r: &[sys::Mont], a_o: &[sys::Mont], vec_key: u64, blind3: &[u8; 96],
    ) -> Result<[[u8; 32]; 3], Error> {
        if a_r.len() != a_l.len() || a_o.len() != a_l.len() {
            return Err(Error::InvalidInput("vector lengths differ"));
        }
        let mut out = [[0u8; 32]; 3];
        let g = self.gens;
        check(unsafe {
            sys::bpg_r1cs_dev_commit(
                self.raw, g.table, g.g_base(), g.h_base(), g.b_blinding_id(), first, a_l.len(), a_l.as_ptr() as *const _,
                a_r.as_ptr() as *const _, a_o.as_ptr() as *const _, vec_key, blind3.as_ptr(), out.as_mut_ptr() as *mut u8,
            )
        })?;
        Ok(out)
    }
    /// `flattened_constraints(z)`: terms in CSR order (code = kind << 28 | index with kinds
    /// 1 left, 2 right, 3 output, 4 committed, 5 one; row; coefficient).  Returns (wV, wc).
    pub fn flatten(
        &mut self, n: usize, m: usize, t_code: &[u32], t_row: &[u32], t_coeff: &[sys::Mont], z_pow: &sys::PowTable,
    ) -> Result<(Vec<sys::Mont>, sys::Mont), Error> {
        if t_row.len() != t_code.len() || t_coeff.len() != t_code.len() {
            return Err(Error::InvalidInput("vector lengths differ"));
        }
        let mut wv = vec![[0u32; 8]; m + 1];
        check(unsafe {
            sys::bpg_r1cs_dev_flatten(
                self.raw, n, m, t_code.len(), t_code.as_ptr(), t_row.as_ptr(), t_coeff.as_ptr() as *const _,
                z_pow.as_ptr() as *const _, wv.as_mut_ptr() as *mut _,
            )
        })?;
        let wc = wv.pop().unwrap();
        Ok((wv, wc))
    }
    /// The same with the terms in two lists: `unit` terms have coefficient +1, or -1 with bit 31 of the code, and
    /// carry no coefficient (8 bytes on the bus instead of 40).  Finds the lists resident when `prefetch_terms`
    /// named the same slices and they have not changed since.
    pub fn flatten_terms(
        &mut self, n: usize, m: usize, general: (&[u32], &[u32], &[sys::Mont]), unit: (&[u32], &[u32]), z_pow: &sys::PowTable,
    ) -> Result<(Vec<sys::Mont>, sys::Mont), Error> {
        let t = terms_of(general, unit)?;
        let mut wv = vec![[0u32; 8]; m + 1];
        check(unsafe {
            sys::bpg_r1cs_dev_flatten_terms(self.raw, n, m, &t, z_pow.as_ptr() as *const _, wv.as_mut_ptr() as *mut _)
        })?;
        let wc = wv.pop().unwrap();
        Ok((wv, wc))
    }
    /// t_1..t_6 as canonical scalars (src/util.rs:152-170)
    pub fn poly_t(&mut self, n: usize, y_pow: &sys::PowTable, y_inv_pow: &sys::PowTable) -> Result<[[u8; 32]; 6], Error> {
        let mut t = [[0u8; 32]; 6];
        check(unsafe {
            sys::bpg_r1cs_dev_poly_t(self.raw, n, y_pow.as_ptr() as *const _, y_inv_pow.as_ptr() as *const _, t.as_mut_ptr() as *mut u8)
        })?;
        Ok(t)
    }
    /// l(x), r(x), padding and factors on the device, then the IPP over the shared table with
    /// Q = w * B (src/r1cs/prover.rs:650-708).
    pub fn ipp_begin(
        &mut self, w: &[u8; 32], n: usize, n1: usize, padded_n: usize, x: &sys::Mont, u: &sys::Mont, y_pow: &sys::PowTable,
        y_inv_pow: &sys::PowTable,
    ) -> Result<IppSession<'g>, Error> {
        let mut raw = ptr::null_mut();
        let g = self.gens;
        check(unsafe {
            sys::bpg_r1cs_dev_ipp_begin(
                self.raw, g.table, g.g_base(), g.h_base(), g.b_id(), w.as_ptr(), n, n1, padded_n, x.as_ptr() as *const _,
                u.as_ptr() as *const _, y_pow.as_ptr() as *const _, y_inv_pow.as_ptr() as *const _, &mut raw,
            )
        })?;
        Ok(IppSession::from_raw(raw))
    }
    /// The verifier's mega-MSM (src/r1cs/verifier.rs:516-547); `Ok(())` iff the sum is the identity (:549).
    pub fn verify_msm(
        &mut self, adhoc_points: &[[u8; 32]], adhoc_scalars: &[[u8; 32]], bb_scalar: &[u8; 32], params: &sys::bpg_verify_params,
    ) -> Result<(), Error> {
        if adhoc_points.len() != adhoc_scalars.len() {
            return Err(Error::InvalidInput("vector lengths differ"));
        }
        let mut out = [0u8; 32];
        let g = self.gens;
        check(unsafe {
            sys::bpg_r1cs_dev_verify_msm(
                self.raw, g.table, g.g_base(), g.h_base(), g.b_id(), adhoc_points.as_ptr() as *const u8,
                adhoc_scalars.as_ptr() as *const u8, adhoc_points.len(), bb_scalar.as_ptr(), params, out.as_mut_ptr(),
            )
        })?;
        if out == [0u8; 32] { Ok(()) } else { Err(Error::VerificationError) }
    }
}

impl Drop for R1csDevice<'_> {
    fn drop(&mut self) {
        unsafe { sys::bpg_r1cs_dev_free(self.raw) }
    }
}

fn terms_of(general: (&[u32], &[u32], &[sys::Mont]), unit: (&[u32], &[u32])) -> Result<sys::bpg_terms, Error> {
    let (t_code, t_row, t_coeff) = general;
    let (u_code, u_row) = unit;
    if t_row.len() != t_code.len() || t_coeff.len() != t_code.len() || u_row.len() != u_code.len() {
        return Err(Error::InvalidInput("vector lengths differ"));
    }
    Ok(sys::bpg_terms {
        n_terms: t_code.len(),
        t_code: t_code.as_ptr(),
        t_row: t_row.as_ptr(),
        t_coeff: t_coeff.as_ptr() as *const _,
        n_unit: u_code.len(),
        u_code: u_code.as_ptr(),
        u_row: u_row.as_ptr(),
    })
}

/// Work that depends on no challenge, started before the transcript is replayed.
/// `ctx` is the context the later calls run on.
///
/// # Safety
/// The slices of `prefetch_terms` must stay alive and unchanged until the matching `flatten_terms`
/// (or `terms_wait`); page-locked memory (`bpg_host_alloc`) lets the copy run beside other work.
pub unsafe fn prefetch_terms(
    ctx: *mut sys::bpg_ctx, general: (&[u32], &[u32], &[sys::Mont]), unit: (&[u32], &[u32]), after_commit_uploads: bool,
) -> Result<(), Error> {
    let t = terms_of(general, unit)?;
    check(sys::bpg_r1cs_terms_prefetch(ctx, &t, after_commit_uploads as std::os::raw::c_int))
}
/// The prefetched copy has left the host arrays (call before second-phase constraints are appended).
pub unsafe fn terms_wait(ctx: *mut sys::bpg_ctx) -> Result<(), Error> {
    check(sys::bpg_r1cs_terms_wait(ctx))
}
/// The points of the verifier's final check (src/r1cs/verifier.rs:516-547), handed over as soon as the proof is
/// parsed: their doubling chains run beside the transcript replay, and the MSM that later names the same
/// encodings in the same order adds comb entries instead of running its own double-and-add.
pub unsafe fn prefetch_points(ctx: *mut sys::bpg_ctx, points: &[[u8; 32]]) -> Result<(), Error> {
    check(sys::bpg_adhoc_prefetch(ctx, points.as_ptr() as *const u8, points.len()))
}
