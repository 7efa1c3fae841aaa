//! Raw bindings to `include/bpgpu.h`.  One declaration per entry point; comments name the
//! reference interface each one stands behind (paths relative to renegade-fi/mpc-bulletproof).
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

macro_rules! opaque {
    ($($name:ident),*) => { $( #[repr(C)] pub struct $name { _private: [u8; 0] } )* };
}
opaque!(bpg_ctx, bpg_table, bpg_ipp, bpg_comb, bpg_r1cs_dev, bpg_vbatch, bpg_stark_table, bpg_stark_ipp, bpg_peer, bpg_msm_job);

pub const BPG_OK: c_int = 0;
pub const BPG_ERR_ARG: c_int = -1;
pub const BPG_ERR_LEN: c_int = -2;
/// `assert!(n.is_power_of_two())`, src/inner_product_proof.rs:69
pub const BPG_ERR_POW2: c_int = -3;
/// `R1CSError::InvalidGeneratorsLength`, src/r1cs/prover.rs:450
pub const BPG_ERR_CAPACITY: c_int = -4;
/// `ProofError::FormatError` / `R1CSError::FormatError`
pub const BPG_ERR_DECODE: c_int = -5;
/// `ProofError::VerificationError` / `R1CSError::VerificationError`
pub const BPG_ERR_VERIFY: c_int = -6;
pub const BPG_ERR_CUDA: c_int = -7;
pub const BPG_ERR_NOMEM: c_int = -8;

/// Montgomery limbs (x * 2^256 mod l), eight little-endian u32 words.
pub type Mont = [u32; 8];
/// base^(2^k), k < 32, Montgomery limbs.
pub type PowTable = [Mont; 32];

/// Parameters of the verifier's scalar preparation, src/r1cs/verifier.rs:468-501, 527-529.
#[repr(C)]
pub struct bpg_verify_params {
    pub y_inv_pow: PowTable,
    /// u_j^2 in creation order, src/inner_product_proof.rs:288-292
    pub u_sq: [Mont; 32],
    pub allinv: Mont,
    pub x: Mont,
    pub a: Mont,
    pub b: Mont,
    pub u: Mont,
    /// scalar of B = c0 + c1 * delta
    pub c0: Mont,
    pub c1: Mont,
    pub lg_n: u32,
    pub n: u32,
    pub n1: u32,
    pub padded_n: u32,
}

/// Parameters of `InnerProductProof::verify`'s scalars, src/inner_product_proof.rs:283-307.
#[repr(C)]
pub struct bpg_ipp_verify_params {
    pub u_sq: [Mont; 32],
    pub allinv: Mont,
    pub a: Mont,
    pub b: Mont,
    pub lg_n: u32,
    pub padded_n: u32,
}

/// The flat constraint terms, general ones (with a coefficient) and unit ones (+1, or -1 with bit 31 of the code).
#[repr(C)]
pub struct bpg_terms {
    pub n_terms: usize,
    pub t_code: *const u32,
    pub t_row: *const u32,
    pub t_coeff: *const c_void,
    pub n_unit: usize,
    pub u_code: *const u32,
    pub u_row: *const u32,
}

extern "C" {
    // ---- context -----------------------------------------------------------------------
    pub fn bpg_init(device: c_int, out: *mut *mut bpg_ctx) -> c_int;
    pub fn bpg_free(ctx: *mut bpg_ctx);
    pub fn bpg_sync(ctx: *mut bpg_ctx) -> c_int;
    pub fn bpg_strerror(code: c_int) -> *const c_char;
    pub fn bpg_last_cuda_error(ctx: *const bpg_ctx) -> c_int;

    // ---- generator tables: BulletproofGens / PedersenGens as data, src/generators.rs:158-235
    pub fn bpg_table_upload(ctx: *mut bpg_ctx, points: *const u8, n: usize, out: *mut *mut bpg_table) -> c_int;
    pub fn bpg_table_set_windows(ctx: *mut bpg_ctx, t: *mut bpg_table, c: c_int) -> c_int;
    pub fn bpg_table_len(t: *const bpg_table) -> usize;
    pub fn bpg_table_free(t: *mut bpg_table);
    // generator derivation on the device (GeneratorsChain, src/generators.rs:80-125, ristretto255 form)
    pub fn bpg_points_from_uniform(ctx: *mut bpg_ctx, uniform64: *const u8, n: usize, out: *mut u8) -> c_int;
    pub fn bpg_gens_chain(ctx: *mut bpg_ctx, label: *const u8, label_len: usize, skip: usize, n: usize, out: *mut u8) -> c_int;

    // ---- StarkPoint::msm_iter / ::msm (all call sites of SURVEY.md 2.2) ---------------
    pub fn bpg_msm(ctx: *mut bpg_ctx, scalars: *const u8, points: *const u8, n: usize, out: *mut u8) -> c_int;
    pub fn bpg_msm_table(
        ctx: *mut bpg_ctx, t: *const bpg_table, offset: usize, n: usize, scalars: *const u8, n_sets: c_int, out: *mut u8,
    ) -> c_int;
    pub fn bpg_msm_table_indexed(
        ctx: *mut bpg_ctx, t: *const bpg_table, point_ids: *const u32, set_ids: *const u8, scalars: *const u8,
        n_terms: usize, n_sets: c_int, out: *mut u8,
    ) -> c_int;
    pub fn bpg_msm_mixed(
        ctx: *mut bpg_ctx, adhoc_points: *const u8, n_adhoc: usize, tabs: *const *const bpg_table, offs: *const usize,
        lens: *const usize, nsegs: c_int, scalars: *const u8, out: *mut u8,
    ) -> c_int;
    /// the same ad-hoc points ahead of their scalars: doubling chains beside the transcript replay
    pub fn bpg_adhoc_prefetch(ctx: *mut bpg_ctx, points: *const u8, n: usize) -> c_int;
    /// one rank's / one MPC party's partial sums: n_sets x 128 bytes (X|Y|Z|T)
    pub fn bpg_msm_table_partial(
        ctx: *mut bpg_ctx, t: *const bpg_table, offset: usize, n: usize, scalars: *const u8, n_sets: c_int,
        out_ext: *mut u8,
    ) -> c_int;
    /// out[s] = encode(sum_p parts[p][s])
    pub fn bpg_sum_encode(ctx: *mut bpg_ctx, parts_ext: *const u8, n_parts: c_int, n_sets: c_int, out: *mut u8) -> c_int;

    // ---- sharded sums, one process per GPU: exchange fused with the combine (peer-mapped buffers) ----
    pub fn bpg_dev_msm_table(
        ctx: *mut bpg_ctx, t: *const bpg_table, offset: usize, n: usize, d_scalars: *const c_void, n_sets: c_int,
        d_out_ext: *mut c_void,
    ) -> c_int;
    pub fn bpg_peer_create(
        ctx: *mut bpg_ctx, world: c_int, rank: c_int, max_sets: c_int, out: *mut *mut bpg_peer, handle_out: *mut u8,
    ) -> c_int;
    pub fn bpg_peer_connect(p: *mut bpg_peer, handles: *const u8) -> c_int;
    pub fn bpg_dev_exchange_sum_encode(
        ctx: *mut bpg_ctx, p: *mut bpg_peer, d_part: *const c_void, n_sets: c_int, d_out_bytes: *mut c_void,
        d_out_ext: *mut c_void,
    ) -> c_int;
    pub fn bpg_peer_status(p: *mut bpg_peer, status_out: *mut c_int) -> c_int;
    pub fn bpg_peer_free(p: *mut bpg_peer);

    // ---- InnerProductProof::create, src/inner_product_proof.rs:49-193 -------------------
    pub fn bpg_ipp_begin(
        ctx: *mut bpg_ctx, g: *const bpg_table, g_off: usize, h: *const bpg_table, h_off: usize, n: usize, q: *const u8,
        g_factors: *const u8, h_factors: *const u8, a: *const u8, b: *const u8, out: *mut *mut bpg_ipp,
    ) -> c_int;
    pub fn bpg_ipp_begin_shared(
        ctx: *mut bpg_ctx, shared: *const bpg_table, g_base: usize, h_base: usize, q_id: usize, q_mul: *const u8,
        n: usize, g_factors: *const u8, h_factors: *const u8, a: *const u8, b: *const u8, out: *mut *mut bpg_ipp,
    ) -> c_int;
    pub fn bpg_ipp_rounds_left(st: *const bpg_ipp) -> usize;
    pub fn bpg_ipp_round_LR(st: *mut bpg_ipp, l: *mut u8, r: *mut u8) -> c_int;
    pub fn bpg_ipp_round_fold(st: *mut bpg_ipp, u: *const u8, u_inv: *const u8) -> c_int;
    pub fn bpg_ipp_finish(st: *mut bpg_ipp, a: *mut u8, b: *mut u8) -> c_int;
    pub fn bpg_ipp_free(st: *mut bpg_ipp);
    /// InnerProductProof::verify, src/inner_product_proof.rs:317-372
    pub fn bpg_ipp_verify_msm(
        ctx: *mut bpg_ctx, g: *const bpg_table, g_off: usize, h: *const bpg_table, h_off: usize, adhoc_points: *const u8,
        adhoc_scalars: *const u8, n_adhoc: usize, g_factors: *const u8, h_factors: *const u8,
        params: *const bpg_ipp_verify_params, out: *mut u8,
    ) -> c_int;

    // ---- PedersenGens::commit, src/generators.rs:41-43 -----------------------------------
    pub fn bpg_comb_create(ctx: *mut bpg_ctx, bases: *const u8, nbases: c_int, out: *mut *mut bpg_comb) -> c_int;
    pub fn bpg_comb_mul(ctx: *mut bpg_ctx, comb: *const bpg_comb, scalars: *const u8, n: usize, out: *mut u8) -> c_int;
    pub fn bpg_comb_free(c: *mut bpg_comb);

    // ---- R1CS scalar preparation in HBM --------------------------------------------------
    pub fn bpg_r1cs_dev_new(ctx: *mut bpg_ctx, capacity: usize, out: *mut *mut bpg_r1cs_dev) -> c_int;
    pub fn bpg_r1cs_dev_reserve(st: *mut *mut bpg_r1cs_dev, capacity: usize) -> c_int;
    pub fn bpg_r1cs_dev_free(st: *mut bpg_r1cs_dev);
    /// (A_I, A_O, S) of one phase, src/r1cs/prover.rs:465-494, 532-565
    pub fn bpg_r1cs_dev_commit(
        st: *mut bpg_r1cs_dev, gens: *const bpg_table, g_base: usize, h_base: usize, bb_id: usize, first: usize,
        cnt: usize, a_l: *const c_void, a_r: *const c_void, a_o: *const c_void, vec_key: u64, blind3: *const u8,
        out: *mut u8,
    ) -> c_int;
    /// flattened_constraints, src/r1cs/prover.rs:342-379, src/r1cs/verifier.rs:323-362
    pub fn bpg_r1cs_dev_flatten(
        st: *mut bpg_r1cs_dev, n: usize, m: usize, n_terms: usize, t_code: *const u32, t_row: *const u32,
        t_coeff: *const c_void, z_pow: *const c_void, wv_out: *mut c_void,
    ) -> c_int;
    /// batch verification: the generator scalars of several proofs side by side, combined into one MSM
    pub fn bpg_vbatch_new(ctx: *mut bpg_ctx, n: usize, capacity: usize, out: *mut *mut bpg_vbatch) -> c_int;
    pub fn bpg_vbatch_free(b: *mut bpg_vbatch);
    pub fn bpg_vbatch_put(
        b: *mut bpg_vbatch, k: usize, st: *mut bpg_r1cs_dev, bb_scalar: *const u8, params: *const bpg_verify_params,
    ) -> c_int;
    pub fn bpg_vbatch_check(
        b: *mut bpg_vbatch, gens: *const bpg_table, g_base: usize, h_base: usize, b_id: usize, idx: *const u32, cnt: usize,
        rho: *const u8, adhoc_points: *const u8, adhoc_scalars: *const u8, n_adhoc: usize, out: *mut u8,
    ) -> c_int;
    /// flattened_constraints with the +1 / -1 terms in a list of their own (no coefficient stored)
    pub fn bpg_r1cs_dev_flatten_terms(
        st: *mut bpg_r1cs_dev, n: usize, m: usize, terms: *const bpg_terms, z_pow: *const c_void, wv_out: *mut c_void,
    ) -> c_int;
    /// early upload of the constraint terms (they depend on no challenge); consumed by the next flatten
    pub fn bpg_r1cs_terms_prefetch(ctx: *mut bpg_ctx, terms: *const bpg_terms, after_commit_uploads: c_int) -> c_int;
    /// the prefetched copy has left the host arrays: they may change (second-phase constraints) or be freed
    pub fn bpg_r1cs_terms_wait(ctx: *mut bpg_ctx) -> c_int;
    /// t_1..t_6, src/util.rs:152-170
    pub fn bpg_r1cs_dev_poly_t(
        st: *mut bpg_r1cs_dev, n: usize, y_pow: *const c_void, y_inv_pow: *const c_void, t_out: *mut u8,
    ) -> c_int;
    /// l(x), r(x), padding, factors, then the IPP state, src/r1cs/prover.rs:650-708
    pub fn bpg_r1cs_dev_ipp_begin(
        st: *mut bpg_r1cs_dev, gens: *const bpg_table, g_base: usize, h_base: usize, q_id: usize, q_mul: *const u8,
        n: usize, n1: usize, padded_n: usize, x: *const c_void, u: *const c_void, y_pow: *const c_void,
        y_inv_pow: *const c_void, out: *mut *mut bpg_ipp,
    ) -> c_int;
    /// the verifier's mega-MSM, src/r1cs/verifier.rs:468-547
    pub fn bpg_r1cs_dev_verify_msm(
        st: *mut bpg_r1cs_dev, gens: *const bpg_table, g_base: usize, h_base: usize, b_id: usize,
        adhoc_points: *const u8, adhoc_scalars: *const u8, n_adhoc: usize, bb_scalar: *const u8,
        params: *const bpg_verify_params, out: *mut u8,
    ) -> c_int;

    // ---- Stark-curve policy: mpc_stark::algebra::stark_curve::StarkPoint::msm_iter / ::msm ----
    // points are affine x || y, 32 bytes little-endian each (src/util.rs:274-289); identity = 64 zero bytes
    pub fn bpg_stark_table_upload(ctx: *mut bpg_ctx, points_xy: *const u8, n: usize, out: *mut *mut bpg_stark_table) -> c_int;
    pub fn bpg_stark_table_set_windows(ctx: *mut bpg_ctx, t: *mut bpg_stark_table, c: c_int) -> c_int;
    pub fn bpg_stark_table_len(t: *const bpg_stark_table) -> usize;
    pub fn bpg_stark_table_free(t: *mut bpg_stark_table);
    pub fn bpg_stark_msm_table(
        ctx: *mut bpg_ctx, t: *const bpg_stark_table, offset: usize, n: usize, scalars: *const u8, n_sets: c_int,
        out_xy: *mut u8,
    ) -> c_int;
    pub fn bpg_stark_msm(ctx: *mut bpg_ctx, scalars: *const u8, points_xy: *const u8, n: usize, out_xy: *mut u8) -> c_int;

    // InnerProductProof::create over the Stark curve, split at the transcript (src/inner_product_proof.rs:49-193)
    pub fn bpg_stark_ipp_begin(
        ctx: *mut bpg_ctx, g: *const bpg_stark_table, g_off: usize, h: *const bpg_stark_table, h_off: usize, n: usize,
        q_xy: *const u8, g_factors: *const u8, h_factors: *const u8, a: *const u8, b: *const u8,
        out: *mut *mut bpg_stark_ipp,
    ) -> c_int;
    pub fn bpg_stark_ipp_rounds_left(st: *const bpg_stark_ipp) -> usize;
    pub fn bpg_stark_ipp_round_LR(st: *mut bpg_stark_ipp, l_xy: *mut u8, r_xy: *mut u8) -> c_int;
    pub fn bpg_stark_ipp_round_fold(st: *mut bpg_stark_ipp, u: *const u8, u_inv: *const u8) -> c_int;
    pub fn bpg_stark_ipp_finish(st: *mut bpg_stark_ipp, a: *mut u8, b: *mut u8) -> c_int;
    pub fn bpg_stark_ipp_free(st: *mut bpg_stark_ipp);

    // ---- round 2 additions ----------------------------------------------------------------
    /// combs of a resident table: inner-product rounds without the bucket method (csrc/comb_kernels.cuh)
    pub fn bpg_table_build_comb(ctx: *mut bpg_ctx, t: *mut bpg_table) -> c_int;
    pub fn bpg_table_has_comb(t: *const bpg_table) -> c_int;
    pub fn bpg_table_entry_bytes(t: *const bpg_table) -> usize;
    /// a stream of host-buffer MSMs, two in flight: uploads behind kernels
    pub fn bpg_msm_table_submit(
        ctx: *mut bpg_ctx, t: *const bpg_table, offset: usize, n: usize, scalars: *const u8, n_sets: c_int,
        job: *mut *mut bpg_msm_job,
    ) -> c_int;
    pub fn bpg_msm_job_wait(job: *mut bpg_msm_job, out: *mut u8) -> c_int;
    /// open of additively shared points (the parties' compressed shares): out[s] = sum_p points[p][s]
    pub fn bpg_points_sum(ctx: *mut bpg_ctx, points: *const u8, n_parts: c_int, n_sets: c_int, out: *mut u8) -> c_int;
    /// SharedInnerProductProof::create, one party's local work (src/r1cs_mpc/mpc_inner_product.rs:52-228):
    /// `lanes` (a, b) pairs (value shares, MAC shares) over public generators; cross terms from the fabric
    pub fn bpg_ipp_begin_shares(
        ctx: *mut bpg_ctx, shared: *const bpg_table, g_base: usize, h_base: usize, q_id: usize, q_mul: *const u8,
        n: usize, lanes: c_int, g_factors: *const u8, h_factors: *const u8, a: *const u8, b: *const u8,
        out: *mut *mut bpg_ipp,
    ) -> c_int;
    pub fn bpg_ipp_lanes(st: *const bpg_ipp) -> c_int;
    pub fn bpg_ipp_len(st: *const bpg_ipp) -> usize;
    pub fn bpg_ipp_read_ab(st: *mut bpg_ipp, a_out: *mut u8, b_out: *mut u8) -> c_int;
    pub fn bpg_ipp_round_LR_shares(
        st: *mut bpg_ipp, c_l: *const u8, c_r: *const u8, l_out: *mut u8, r_out: *mut u8,
    ) -> c_int;
    pub fn bpg_ipp_finish_shares(st: *mut bpg_ipp, a: *mut u8, b: *mut u8) -> c_int;
    /// (A_I, A_O, S) with s_L, s_R expanded from a 256-bit key (ChaCha20): the production form
    pub fn bpg_r1cs_dev_commit_keyed(
        st: *mut bpg_r1cs_dev, gens: *const bpg_table, g_base: usize, h_base: usize, bb_id: usize, first: usize,
        cnt: usize, a_l: *const c_void, a_r: *const c_void, a_o: *const c_void, vec_key: *const u8, blind3: *const u8,
        out: *mut u8,
    ) -> c_int;
    /// the fork's Stark-curve hashing conventions (src/util.rs:252-267, src/generators.rs:80-125)
    pub fn bpg_keccak256(data: *const u8, len: usize, out: *mut u8);
    pub fn bpg_stark_hash_to_scalar(low: *const u8, out: *mut u8);
    pub fn bpg_stark_gens_chain(ctx: *mut bpg_ctx, state0: *const u8, skip: usize, n: usize, out_xy: *mut u8) -> c_int;

    // ---- page-locked staging ---------------------------------------------------------------
    pub fn bpg_host_alloc(bytes: usize) -> *mut c_void;
    pub fn bpg_host_free(p: *mut c_void);
}
