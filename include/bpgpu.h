/* bpgpu.h — C ABI of libbpgpu, the B200 (sm_100a) engine for the multiscalar
 * multiplications and inner-product-argument folding of Bulletproofs R1CS.
 *
 * Drop-in boundary for renegade-fi/mpc-bulletproof: the reference has no FFI of
 * its own; the seam is its calls into the group dependency.  Each entry point
 * below names the reference interface it stands behind.  INTEGRATION.md shows
 * the Rust `extern "C"` block that binds them.
 *
 * Conventions
 *  - group: ristretto255; points on the wire are 32-byte
 *    compressed encodings (RFC 9496), scalars are 32-byte little-endian,
 *    canonical (< l).
 *  - every function returns BPG_OK (0) or a negative error code; nothing aborts.
 *  - host-buffer calls are synchronous: outputs are valid on return.
 *  - `bpg_dev_*` calls take DEVICE pointers and enqueue on the context's stream
 *    (set it with bpg_set_stream); they return after enqueueing.
 *  - a bpg_ctx is single-owner (one proving thread); contexts are independent.
 *  - there is no CPU fallback: without a CUDA device every call fails with
 *    BPG_ERR_CUDA.
 */
#ifndef BPGPU_H
#define BPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPG_OK 0
#define BPG_ERR_ARG -1      /* null pointer / bad size                                        */
#define BPG_ERR_LEN -2      /* vector lengths differ (reference: assert_eq!, inner_product_proof.rs:59-66) */
#define BPG_ERR_POW2 -3     /* length not a power of two (inner_product_proof.rs:69)         */
#define BPG_ERR_CAPACITY -4 /* R1CSError::InvalidGeneratorsLength (r1cs/prover.rs:450-452)   */
#define BPG_ERR_DECODE -5   /* ProofError::FormatError / R1CSError::FormatError              */
#define BPG_ERR_VERIFY -6   /* ProofError::VerificationError / R1CSError::VerificationError  */
#define BPG_ERR_CUDA -7     /* CUDA runtime failure (bpg_last_cuda_error gives the code)     */
#define BPG_ERR_NOMEM -8

typedef struct bpg_ctx bpg_ctx;
typedef struct bpg_table bpg_table; /* points resident in HBM in affine-Niels form */

/* ---- context --------------------------------------------------------------------- */
int bpg_init(int device, bpg_ctx** out);
void bpg_free(bpg_ctx* ctx);
/* use_own != 0: the context's own non-blocking stream (the default after bpg_init);
 * otherwise cuda_stream is a cudaStream_t used verbatim (NULL = legacy default stream). */
int bpg_set_stream(bpg_ctx* ctx, void* cuda_stream, int use_own);
int bpg_sync(bpg_ctx* ctx);
const char* bpg_strerror(int code);
int bpg_last_cuda_error(const bpg_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t bpg_launch_count(const bpg_ctx* ctx);
/* force the Pippenger window width (0 = choose from the size); for tuning/tests */
int bpg_set_window(bpg_ctx* ctx, int c);
/* tuning/test hook: bucket groups per set on windowed tables (0 = chosen from the launch size) */
int bpg_set_groups(bpg_ctx* ctx, int gsub);

/* ---- per-phase device timing --------------------------------------------------------
 * When enabled, CUDA events are recorded on the launch stream around each kernel
 * phase; bpg_profile_read waits for the last event and returns, per phase, the
 * summed milliseconds and the number of intervals since the last reset. */
#define BPG_PROF_HIST 0
#define BPG_PROF_SCAN 1
#define BPG_PROF_SCATTER 2
#define BPG_PROF_ACCUM 3
#define BPG_PROF_ACCUM_BIG 4
#define BPG_PROF_REDUCE 5
#define BPG_PROF_COMBINE 6
#define BPG_PROF_HORNER 7
#define BPG_PROF_ENCODE 8
#define BPG_PROF_OTHER 9
#define BPG_PROF_NPHASE 10
int bpg_profile_enable(bpg_ctx* ctx, int on);
int bpg_profile_reset(bpg_ctx* ctx);
int bpg_profile_read(bpg_ctx* ctx, double* ms, uint64_t* count, int n);
const char* bpg_profile_phase_name(int phase);

/* ---- point tables ------------------------------------------------------------------
 * Upload n compressed points once; they are decoded and kept as affine-Niels
 * (y+x, y-x, 2dxy), 96 bytes each.  Stands behind `BulletproofGens::new` /
 * `.share(j).G(n)/.H(n)` and `PedersenGens` as *data* (reference
 * src/generators.rs:32-71,158-235).  BPG_ERR_DECODE if any encoding is invalid. */
int bpg_table_upload(bpg_ctx* ctx, const uint8_t* points_compressed, size_t n, bpg_table** out);
int bpg_table_upload_dev(bpg_ctx* ctx, const void* d_points_compressed, size_t n, bpg_table** out);
/* Turn a table into a *windowed* table: besides P_i it then holds 2^(c w) P_i for every
 * window w of a c-bit signed-digit decomposition (W = ceil(255/c) times the memory), so
 * that MSMs over it need no doublings at all.  c = 0 picks c from the table length.  One-time
 * cost of ~240 doublings per point; meant for generator tables that live in HBM for many
 * proofs ("uploaded once in precomputed affine-Niels form").  MSMs over a windowed table
 * always use its c. */
int bpg_table_set_windows(bpg_ctx* ctx, bpg_table* t, int c);
int bpg_table_window(const bpg_table* t); /* 0 = plain */
/* Give every point of a resident table a COMB: the 64 x 8 affine-Niels multiples (d+1) 16^j P_i, 48 KB per
 * point, so that k P_i is 64 additions with no doublings, no buckets and no sort.  One-time cost (like
 * bpg_table_set_windows).  The inner-product rounds use the combs of a generator table for short vectors and
 * to materialise the folded generators of long ones (csrc/comb_kernels.cuh); bpg_gens_new builds them when
 * they fit BPG_COMB_MAX_GB (default 32) gigabytes. */
int bpg_table_build_comb(bpg_ctx* ctx, bpg_table* t);
int bpg_table_has_comb(const bpg_table* t);
size_t bpg_table_len(const bpg_table* t);
size_t bpg_table_entry_bytes(const bpg_table* t); /* bytes of HBM per resident multiple of a point */
void bpg_table_free(bpg_table* t);

/* ---- multiscalar multiplication ----------------------------------------------------
 * out = sum_i scalars[i] * points[i].
 * Stands behind `StarkPoint::msm_iter(scalars, points)` / `StarkPoint::msm(&s, &p)`
 * (reference src/inner_product_proof.rs:90-114,159-172,226-227,353;
 * src/r1cs/verifier.rs:516-547). */
int bpg_msm(bpg_ctx* ctx, const uint8_t* scalars_le, const uint8_t* points_compressed, size_t n,
            uint8_t out[32]);

/* n_sets sums over the same n table points table[offset .. offset+n):
 *   out[s] = sum_i scalars[s*n + i] * table[offset + i]
 * One launch serves A_I/A_O/S (reference src/r1cs/prover.rs:465-494) or a party's
 * share and MAC vectors (src/r1cs_mpc/mpc_prover.rs:621-657). */
int bpg_msm_table(bpg_ctx* ctx, const bpg_table* table, size_t offset, size_t n,
                  const uint8_t* scalars_le, int n_sets, uint8_t* out /* n_sets*32 */);

/* Pipelined form of bpg_msm_table for a STREAM of MSMs over a resident table (batch verification, a prover's
 * successive commitments, the throughput benchmark): submit returns as soon as the upload and the kernels are
 * enqueued, wait blocks for the 32-byte results.  At most two jobs are in flight per context; the scalars of
 * the second travel host -> device (its own staging buffer, the copy engine) while the first computes.
 * Page-locked scalars (bpg_host_alloc) make the upload asynchronous; pageable ones work but submit then
 * blocks for the staging copy.  BPG_ERR_ARG if two jobs are already in flight. */
typedef struct bpg_msm_job bpg_msm_job;
int bpg_msm_table_submit(bpg_ctx* ctx, const bpg_table* table, size_t offset, size_t n, const uint8_t* scalars_le,
                         int n_sets, bpg_msm_job** job);
int bpg_msm_job_wait(bpg_msm_job* job, uint8_t* out /* n_sets*32 */);

/* Indexed form: term t = scalars[t] * table[point_ids[t]], added into out[set_ids[t]]
 * (set_ids NULL = all set 0).  One launch forms A_I, A_O and S from (B_blinding, G, H)
 * (reference src/r1cs/prover.rs:465-494, 532-565). */
int bpg_msm_table_indexed(bpg_ctx* ctx, const bpg_table* table, const uint32_t* point_ids, const uint8_t* set_ids,
                          const uint8_t* scalars_le, size_t n_terms, int n_sets, uint8_t* out /* n_sets*32 */);

/* One sum over n_adhoc compressed points followed by nsegs ranges of resident tables:
 *   out = sum_{i<n_adhoc} s[i]*adhoc[i] + sum_seg sum_{k<lens[seg]} s[..]*tabs[seg][offs[seg]+k]
 * with the scalars in that order.  This is the verifier's single "mega" MSM over
 * [A_I1..S2, V_*, T_*, B, B_blinding | G | H | L_*, R_*] (reference
 * src/r1cs/verifier.rs:516-547) and `InnerProductProof::verify`'s MSM (:353-368). */
int bpg_msm_mixed(bpg_ctx* ctx, const uint8_t* adhoc_points, size_t n_adhoc, const bpg_table* const* tabs,
                  const size_t* offs, const size_t* lens, int nsegs, const uint8_t* scalars_le, uint8_t out[32]);
/* The ad-hoc points of a coming bpg_msm_mixed / bpg_r1cs_dev_verify_msm / bpg_ipp_verify_msm, handed over before
 * their scalars exist.  A verifier holds every point of its final check as soon as it has the proof (reference
 * src/r1cs/verifier.rs:516-547), and the 252 doublings of a variable-base multiplication do not depend on the
 * scalar: decoding, doubling chains and digit multiples run on an auxiliary stream while the caller replays its
 * transcript; the later call that names the same encodings in the same order adds comb entries instead.  At most 256
 * points (more: no effect).  An invalid encoding is reported by that later call (BPG_ERR_DECODE), as without this. */
int bpg_adhoc_prefetch(bpg_ctx* ctx, const uint8_t* points, size_t n);

/* Device-resident form: d_scalars (n_sets*n*32 bytes, 16-byte aligned) already in
 * HBM; writes n_sets extended points (4x8 uint32 limbs X,Y,Z,T = 128 bytes each)
 * to d_out_ext.  This is the per-rank partial sum of a sharded MSM. */
int bpg_dev_msm_table(bpg_ctx* ctx, const bpg_table* table, size_t offset, size_t n,
                      const void* d_scalars, int n_sets, void* d_out_ext);

/* d_parts: n_parts*n_sets extended points laid out [part][set] (e.g. the
 * all-gather of every rank's partial sums).  Adds the parts of each set and
 * encodes: d_out_bytes (n_sets*32, device) and/or d_out_ext (n_sets*128, device);
 * either may be NULL. */
int bpg_dev_sum_encode(bpg_ctx* ctx, const void* d_parts, int n_parts, int n_sets, void* d_out_bytes,
                       void* d_out_ext);
/* Host-buffer forms of the two halves of a sharded sum: the partial sums of one
 * rank / one MPC party as extended points (n_sets x 128 bytes: X|Y|Z|T, 32 bytes LE each), and
 * the combine step  out[s] = encode(sum_p parts[p][s])  once the partials have been exchanged
 * (parts laid out [part][set]).  For the MPC prover (reference src/r1cs_mpc/mpc_prover.rs:621-657)
 * a party's additive share of the scalars gives its additive share of the commitment; "open" is
 * the exchange plus bpg_sum_encode. */
int bpg_msm_table_partial(bpg_ctx* ctx, const bpg_table* table, size_t offset, size_t n, const uint8_t* scalars_le,
                          int n_sets, uint8_t* out_ext);
int bpg_sum_encode(bpg_ctx* ctx, const uint8_t* parts_ext, int n_parts, int n_sets, uint8_t* out);

/* Open of additively shared points that crossed the party-to-party link as compressed encodings (the MPC
 * prover: `.open()` / `open_authenticated` of a commitment, reference src/r1cs_mpc/mpc_prover.rs:630-657,
 * mpc_inner_product.rs:131,191): out[s] = encode(sum_p decode(points[p][s])), points laid out [part][set],
 * n_parts <= 32.  BPG_ERR_DECODE if any share is not a valid encoding. */
int bpg_points_sum(bpg_ctx* ctx, const uint8_t* points, int n_parts, int n_sets, uint8_t* out /* n_sets*32 */);

/* Exchange fused with the combine, for one process per GPU on one node: each rank
 * owns an exchange buffer that its peers map through CUDA IPC; bpg_dev_exchange_sum_encode is ONE
 * kernel that stores this rank's partial sums into every rank's buffer over NVLink, raises its flag,
 * waits for all ranks' flags, adds the partials and encodes -- in place of an all-gather followed by a
 * combine launch.  Setup: every rank calls bpg_peer_create, the 64-byte handles are exchanged by the
 * caller (any channel) and passed in rank order to bpg_peer_connect.  world <= 8.  Every rank must call
 * the exchange the same number of times.  A peer that does not show up within ~1 s sets the status. */
typedef struct bpg_peer bpg_peer;
int bpg_peer_create(bpg_ctx* ctx, int world, int rank, int max_sets, bpg_peer** out, uint8_t handle_out[64]);
int bpg_peer_connect(bpg_peer* p, const uint8_t* handles /* world*64 */);
int bpg_dev_exchange_sum_encode(bpg_ctx* ctx, bpg_peer* p, const void* d_part, int n_sets, void* d_out_bytes,
                                void* d_out_ext);
int bpg_peer_status(bpg_peer* p, int* status_out);
void bpg_peer_free(bpg_peer* p);

/* ---- inner-product argument --------------------------------------------------------
 * Device-resident state for `InnerProductProof::create` (reference
 * src/inner_product_proof.rs:49-193).  The transcript stays with the caller, hence the
 * split per round:
 *     bpg_ipp_begin(...)                       // vectors go to HBM once
 *     while (bpg_ipp_rounds_left(st)) {
 *        bpg_ipp_round_LR(st, L, R);           // :87-114 / :156-172  cross terms + the two MSMs
 *        transcript.append(L), append(R); u = challenge            (caller, :119-123)
 *        bpg_ipp_round_fold(st, u, u_inv);     // fold_witness, :202-248
 *     }
 *     bpg_ipp_finish(st, a, b);                // :187-192
 * G, H: tables holding the generator vectors G_vec = G[g_off .. g_off+n), H_vec likewise;
 * G_factors/H_factors: n scalars each or NULL for all-ones (:51-52); a, b: n scalars each.
 * Errors: BPG_ERR_POW2 if n is not a power of two (assert at :69), BPG_ERR_CAPACITY if a
 * table is too short, BPG_ERR_DECODE for an invalid Q. */
typedef struct bpg_ipp bpg_ipp;
int bpg_ipp_begin(bpg_ctx* ctx, const bpg_table* G, size_t g_off, const bpg_table* H, size_t h_off, size_t n,
                  const uint8_t Q[32], const uint8_t* G_factors, const uint8_t* H_factors, const uint8_t* a,
                  const uint8_t* b, bpg_ipp** out);
/* same with the four scalar vectors already in HBM (device pointers; factors may be NULL) */
int bpg_ipp_begin_dev(bpg_ctx* ctx, const bpg_table* G, size_t g_off, const bpg_table* H, size_t h_off, size_t n,
                      const uint8_t Q[32], const void* d_G_factors, const void* d_H_factors, const void* d_a,
                      const void* d_b, bpg_ipp** out);
/* Generators and the base of Q all live in one *windowed* table: G_vec = shared[g_base..),
 * H_vec = shared[h_base..), Q = q_mul * shared[q_id] (q_mul NULL = 1).  This is how the R1CS
 * prover calls it (Q = w*B, reference src/r1cs/prover.rs:686-708): no per-proof table work. */
int bpg_ipp_begin_shared(bpg_ctx* ctx, const bpg_table* shared, size_t g_base, size_t h_base, size_t q_id,
                         const uint8_t q_mul[32], size_t n, const uint8_t* G_factors, const uint8_t* H_factors,
                         const uint8_t* a, const uint8_t* b, bpg_ipp** out);
size_t bpg_ipp_rounds_left(const bpg_ipp* st);
int bpg_ipp_round_LR(bpg_ipp* st, uint8_t L[32], uint8_t R[32]);
int bpg_ipp_round_fold(bpg_ipp* st, const uint8_t u[32], const uint8_t u_inv[32]);
int bpg_ipp_finish(bpg_ipp* st, uint8_t a[32], uint8_t b[32]);
void bpg_ipp_free(bpg_ipp* st);

/* The same rounds on SECRET SHARES: `SharedInnerProductProof::create` (reference
 * src/r1cs_mpc/mpc_inner_product.rs:52-228).  A party holds additive shares of a and b -- `lanes` (a, b)
 * pairs: lane 0 the value shares, further lanes e.g. the MAC shares of an authenticated fabric (all lanes
 * fold with the same public challenge) -- while G, H, the factors and Q are public.  Everything linear is
 * local and runs here: the party's shares of L and R (ONE MSM with 2 x lanes outputs over one pass of the
 * generator table, :104-126, 172-186) and the folds (:136-137, 196-197).  The cross terms c_L = <a_lo, b_hi>,
 * c_R = <a_hi, b_lo> multiply shared values: they come out of the fabric's multiplication protocol (Beaver
 * triples over the network, :104-105, 172-173), which is the caller's:
 *     bpg_ipp_begin_shares(...)                          // share (and MAC) vectors go to HBM once
 *     while (bpg_ipp_rounds_left(st)) {
 *        bpg_ipp_read_ab(st, a_cur, b_cur);              // what the multiplication protocol needs
 *        (c_L, c_R) = fabric: shares of <a_lo, b_hi>, <a_hi, b_lo>                       (caller, network)
 *        bpg_ipp_round_LR_shares(st, c_L, c_R, L_sh, R_sh);   // this party's shares of L, R (compressed)
 *        L, R = open (exchange + bpg_points_sum); transcript; u                          (caller, network)
 *        bpg_ipp_round_fold(st, u, u_inv);
 *     }
 *     bpg_ipp_finish_shares(st, a_sh, b_sh);             // shares of the final a, b (:217-228), opened by the caller
 * Generators: G_vec = shared[g_base..), H_vec = shared[h_base..) of ONE windowed table, Q = q_mul * shared[q_id]
 * (the MPC prover's Q = w*B, src/r1cs_mpc/mpc_prover.rs:926-927).  a, b: lanes x n x 32 bytes, lane-major. */
#define BPG_IPP_MAX_LANES 4
int bpg_ipp_begin_shares(bpg_ctx* ctx, const bpg_table* shared, size_t g_base, size_t h_base, size_t q_id,
                         const uint8_t q_mul[32], size_t n, int lanes, const uint8_t* G_factors,
                         const uint8_t* H_factors, const uint8_t* a, const uint8_t* b, bpg_ipp** out);
int bpg_ipp_lanes(const bpg_ipp* st);
size_t bpg_ipp_len(const bpg_ipp* st); /* current vector length m */
int bpg_ipp_read_ab(bpg_ipp* st, uint8_t* a_out /* lanes*m*32 */, uint8_t* b_out);
int bpg_ipp_round_LR_shares(bpg_ipp* st, const uint8_t* c_L /* lanes*32 */, const uint8_t* c_R, uint8_t* L_out /* lanes*32 */,
                            uint8_t* R_out);
int bpg_ipp_finish_shares(bpg_ipp* st, uint8_t* a /* lanes*32 */, uint8_t* b);

/* ---- fixed-base multiplication (comb) ---------------------------------------------
 * A comb holds, for each of nbases points P_t, the 64x8 affine-Niels multiples
 * (d+1)*16^j*P_t, so that  out[i] = sum_t scalars[t*n + i] * P_t  costs 64 mixed
 * additions per base and no doublings.  With bases {B, B_blinding} this is
 * `PedersenGens::commit(v, v_blinding)` batched over i (reference
 * src/generators.rs:41-43; src/r1cs/prover.rs:319-329,627-631; Q = w*B at :687).
 * With one base it is `Scalar * StarkPoint` for a fixed point. */
typedef struct bpg_comb bpg_comb;
int bpg_comb_create(bpg_ctx* ctx, const uint8_t* bases_compressed /* nbases*32 */, int nbases, bpg_comb** out);
void bpg_comb_free(bpg_comb* comb);
int bpg_comb_mul(bpg_ctx* ctx, const bpg_comb* comb, const uint8_t* scalars_le /* nbases*n*32 */, size_t n,
                 uint8_t* out /* n*32 */);
/* device-resident: either output may be NULL (bytes: n*32, ext: n*128) */
int bpg_dev_comb_mul(bpg_ctx* ctx, const bpg_comb* comb, const void* d_scalars, size_t n, void* d_out_bytes,
                     void* d_out_ext);

/* ---- R1CS scalar preparation on the device ------------------------------------------
 * The O(n) scalar vectors around the prover's and verifier's MSMs are computed in HBM
 * (reference src/r1cs/prover.rs:589-619,650-697; src/r1cs/verifier.rs:468-501;
 * src/inner_product_proof.rs:283-307).  Vectors marked "Montgomery" are n x 8 uint32
 * little-endian limbs of x*2^256 mod l (the host mirror's in-memory form); pow tables hold
 * base^(2^k), k < 32, in the same form. */
typedef struct {
  uint32_t y_inv_pow[32][8]; /* (y^-1)^(2^k), Montgomery                                        */
  uint32_t u_sq[32][8];      /* u_j^2 in creation order (inner_product_proof.rs:288-292)         */
  uint32_t allinv[8], x[8], a[8], b[8], u[8];
  uint32_t c0[8], c1[8];     /* scalar of B = c0 + c1*delta (verifier.rs:527-529)               */
  uint32_t lg_n, n, n1, N;   /* N = 2^lg_n padded size, n multipliers, n1 first-phase           */
} bpg_verify_params;
/* InnerProductProof::verify (src/inner_product_proof.rs:317-372): s_i from (allinv, u_j^2) by
 * its closed form, g_i = a s_i G_factors[i], h_i = b s_{N-1-i} H_factors[i] on the device, then
 * one MSM over [adhoc (Q, L_*, R_*; host scalars) | G[g_off..+N) | H[h_off..+N)].  Factors are
 * canonical bytes or NULL (all ones). */
typedef struct {
  uint32_t u_sq[32][8]; /* Montgomery, creation order */
  uint32_t allinv[8], a[8], b[8];
  uint32_t lg_n, N;
} bpg_ipp_verify_params;
int bpg_ipp_verify_msm(bpg_ctx* ctx, const bpg_table* G, size_t g_off, const bpg_table* H, size_t h_off,
                       const uint8_t* adhoc_points, const uint8_t* adhoc_scalars, size_t n_adhoc,
                       const uint8_t* G_factors, const uint8_t* H_factors, const bpg_ipp_verify_params* params,
                       uint8_t out[32]);

/* Prover side: witness rows, blinding vectors and flattened weights stay resident. */
typedef struct bpg_r1cs_dev bpg_r1cs_dev;
int bpg_r1cs_dev_new(bpg_ctx* ctx, size_t capacity, bpg_r1cs_dev** out);
void bpg_r1cs_dev_free(bpg_r1cs_dev* st);
/* Batch verification (BASELINE config 4; the reference verifies proof by proof, verifier.rs:393): the 2 + 2N generator
 * scalars of several proofs' final checks kept side by side on the device (bpg_vbatch_put: what
 * bpg_r1cs_dev_verify_msm computes before its MSM), then sum_j rho_j * check_j as ONE multiscalar multiplication over
 * [adhoc points | B | B_blinding | G | H] (bpg_vbatch_check; the caller weights the adhoc scalars by rho itself).
 * out = 32 zero bytes iff the combination is the identity.  A single slot with rho = 1 is that proof's own check. */
typedef struct bpg_vbatch bpg_vbatch;
int bpg_vbatch_new(bpg_ctx* ctx, size_t N, size_t capacity, bpg_vbatch** out);
void bpg_vbatch_free(bpg_vbatch* b);
int bpg_vbatch_put(bpg_vbatch* b, size_t k, bpg_r1cs_dev* st, const uint8_t bb_scalar[32], const bpg_verify_params* params);
int bpg_vbatch_check(bpg_vbatch* b, const bpg_table* gens, size_t g_base, size_t h_base, size_t b_id, const uint32_t* idx,
                     size_t cnt, const uint8_t* rho, const uint8_t* adhoc_points, const uint8_t* adhoc_scalars,
                     size_t n_adhoc, uint8_t out[32]);

/* grow to `capacity` rows keeping a_L, a_R, a_O, s_L, s_R (second-phase multipliers, prover.rs:501-530) */
int bpg_r1cs_dev_reserve(bpg_r1cs_dev** st, size_t capacity);
/* (A_I, A_O, S) of one phase over gens[first .. first+cnt) (prover.rs:465-494, 532-565):
 * aL/aR/aO Montgomery rows; the blinding vectors s_L, s_R are generated on the device as the
 * consecutive 64-byte blocks s_L[0], s_R[0], s_L[1], ... of the SplitMix64 stream keyed by
 * vec_key, each reduced mod l; blind3 = i_blinding, o_blinding, s_blinding (canonical). */
int bpg_r1cs_dev_commit(bpg_r1cs_dev* st, const bpg_table* gens, size_t g_base, size_t h_base, size_t bb_id,
                        size_t first, size_t cnt, const void* aL, const void* aR, const void* aO, uint64_t vec_key,
                        const uint8_t blind3[96], uint8_t out[96]);
/* same with s_L, s_R expanded from a 256-bit key: ChaCha20 (RFC 8439) blocks 2j, 2j+1 under vec_key with
 * nonce "bpg sLsR v01", each 64-byte block reduced mod l -- the production form (bpg_prover_prove) */
int bpg_r1cs_dev_commit_keyed(bpg_r1cs_dev* st, const bpg_table* gens, size_t g_base, size_t h_base, size_t bb_id,
                              size_t first, size_t cnt, const void* aL, const void* aR, const void* aO,
                              const uint8_t vec_key[32], const uint8_t blind3[96], uint8_t out[96]);
/* Page-locked host memory for buffers the library reads repeatedly (witness rows, constraint
 * terms): uploads from it run at PCIe rate.  Falls back to malloc when pinning fails. */
void* bpg_host_alloc(size_t bytes);
void bpg_host_free(void* p);
/* flattened_constraints (prover.rs:342-379, verifier.rs:323-362) as a sparse product on the
 * device: term t of constraint row t_row[t] names variable t_code[t] = kind << 28 | index
 * (kinds 1 left, 2 right, 3 output, 4 committed, 5 constant one) with coefficient t_coeff[t]
 * (Montgomery); z_pow is the pow table of z.  wL, wR, wO stay resident in the state; wv_out
 * receives wV[0..m) and w_c, (m + 1) x 32 bytes of Montgomery limbs. */
int bpg_r1cs_dev_flatten(bpg_r1cs_dev* st, size_t n, size_t m, size_t n_terms, const uint32_t* t_code,
                         const uint32_t* t_row, const void* t_coeff, const void* z_pow, void* wv_out);
/* The same with the terms in two lists.  Most coefficients of an R1CS are +1 or -1 (copies, differences, booleans):
 * such a term carries no coefficient -- bit 31 of its code says -1 -- and costs 8 bytes on the bus instead of 40
 * and no multiplication in the kernel.  `t_*` are the general terms as above, `u_*` the unit ones. */
typedef struct {
  size_t n_terms;
  const uint32_t* t_code;
  const uint32_t* t_row;
  const void* t_coeff;
  size_t n_unit;
  const uint32_t* u_code;
  const uint32_t* u_row;
} bpg_terms;
int bpg_r1cs_dev_flatten_terms(bpg_r1cs_dev* st, size_t n, size_t m, const bpg_terms* terms, const void* z_pow,
                               void* wv_out);
/* The terms depend on no challenge: handing them over when the circuit is built lets their upload run on an
 * auxiliary stream beside the commitments (prover.rs:465-494) or the transcript replay (verifier.rs:400-440).
 * The next bpg_r1cs_dev_flatten_terms on this context that names the same arrays (same pointers and lengths,
 * unchanged contents) uses the resident copy; any other uploads its own.  Arrays from bpg_host_alloc copy at bus
 * rate.  after_commit_uploads != 0 queues the copy behind the witness rows of the next bpg_r1cs_dev_commit* (they
 * are on the prover's critical path and share the bus). */
int bpg_r1cs_terms_prefetch(bpg_ctx* ctx, const bpg_terms* terms, int after_commit_uploads);
/* Blocks until a prefetched copy has left the host arrays (and drops a request still waiting for its commitment): call
 * it before the arrays change or are freed while a prefetch may be in flight -- e.g. before second-phase constraints
 * are appended (prover.rs:383-402).  The next flatten then names longer arrays and uploads them itself. */
int bpg_r1cs_terms_wait(bpg_ctx* ctx);
/* t_1..t_6 (util.rs:152-170) from the resident vectors.  t_out: six canonical scalars. */
int bpg_r1cs_dev_poly_t(bpg_r1cs_dev* st, size_t n, const void* y_pow, const void* y_inv_pow, uint8_t t_out[192]);
/* The verifier's mega-MSM (verifier.rs:516-547) with g_scalars, h_scalars and delta computed on
 * the device from the resident weights.  Terms: [adhoc points | B | B_blinding | G[0..N) | H[0..N)];
 * the caller supplies the adhoc scalars and the B_blinding scalar (canonical bytes).
 * out = 32 zero bytes iff the sum is the identity (verifier.rs:549; tested projectively, X = 0 or Y = 0,
 * so no encoding and no inversion); anything else means reject.  When the generators are not one
 * windowed table the general path returns the compressed sum, to be compared with 32 zero bytes as well. */
int bpg_r1cs_dev_verify_msm(bpg_r1cs_dev* st, const bpg_table* gens, size_t g_base, size_t h_base, size_t b_id,
                            const uint8_t* adhoc_points, const uint8_t* adhoc_scalars, size_t n_adhoc,
                            const uint8_t bb_scalar[32], const bpg_verify_params* params, uint8_t out[32]);
/* l(x), r(x) with padding and the G/H factors (prover.rs:650-697) computed in HBM, then the IPP
 * state over the shared generator table with Q = q_mul * gens[q_id]. */
int bpg_r1cs_dev_ipp_begin(bpg_r1cs_dev* st, const bpg_table* gens, size_t g_base, size_t h_base, size_t q_id,
                           const uint8_t q_mul[32], size_t n, size_t n1, size_t N, const void* x_mont,
                           const void* u_mont, const void* y_pow, const void* y_inv_pow, bpg_ipp** out);

/* ==== Stark-curve policy ============================================
 * The mounted fork computes over the Stark curve (y^2 = x^3 + x + beta over F_p,
 * p = 2^251 + 17*2^192 + 1) through mpc-stark's StarkPoint (reference Cargo.toml:13,21,
 * src/generators.rs:11-16).  Same Pippenger pipeline, second field/curve policy.  Points cross
 * the boundary as affine x || y, 32 bytes little-endian each (the fork's transcript encoding,
 * src/util.rs:274-289); the identity is 64 zero bytes; scalars are 32 bytes little-endian, reduced
 * mod the group order.  BPG_ERR_DECODE for a coordinate >= p or a point off the curve.
 * Round 1 covers `StarkPoint::msm_iter` / `::msm`; the protocol layers above keep using the
 * ristretto255 instantiation. */
typedef struct bpg_stark_table bpg_stark_table;
int bpg_stark_table_upload(bpg_ctx* ctx, const uint8_t* points_xy /* n*64 */, size_t n, bpg_stark_table** out);
/* precompute 2^(c w) P_i for every window (c = 0: chosen from the table length); one-time cost, no doublings afterwards */
int bpg_stark_table_set_windows(bpg_ctx* ctx, bpg_stark_table* t, int c);
int bpg_stark_table_window(const bpg_stark_table* t);
size_t bpg_stark_table_len(const bpg_stark_table* t);
void bpg_stark_table_free(bpg_stark_table* t);
/* out[s] = sum_i scalars[s*n + i] * table[offset + i], n_sets x 64 bytes */
int bpg_stark_msm_table(bpg_ctx* ctx, const bpg_stark_table* table, size_t offset, size_t n, const uint8_t* scalars_le,
                        int n_sets, uint8_t* out_xy);
int bpg_stark_msm(bpg_ctx* ctx, const uint8_t* scalars_le, const uint8_t* points_xy, size_t n, uint8_t out_xy[64]);

/* Inner-product-argument rounds over the Stark curve: `InnerProductProof::create`
 * (src/inner_product_proof.rs:49-193) with the vectors resident in HBM, split at the transcript as
 * bpg_ipp_* is (the fork's hash-chain transcript stays with the caller): round_LR returns the two
 * points (x || y), the caller derives u and passes (u, u^-1) to round_fold; finish returns the last
 * a, b.  G, H: ranges of resident Stark tables; factors may be NULL (all ones); scalars 32 bytes LE,
 * canonical mod the Stark group order. */
typedef struct bpg_stark_ipp bpg_stark_ipp;
int bpg_stark_ipp_begin(bpg_ctx* ctx, const bpg_stark_table* G, size_t g_off, const bpg_stark_table* H, size_t h_off,
                        size_t n, const uint8_t Q_xy[64], const uint8_t* G_factors, const uint8_t* H_factors,
                        const uint8_t* a, const uint8_t* b, bpg_stark_ipp** out);
size_t bpg_stark_ipp_rounds_left(const bpg_stark_ipp* st);
int bpg_stark_ipp_round_LR(bpg_stark_ipp* st, uint8_t L_xy[64], uint8_t R_xy[64]);
int bpg_stark_ipp_round_fold(bpg_stark_ipp* st, const uint8_t u[32], const uint8_t u_inv[32]);
int bpg_stark_ipp_finish(bpg_stark_ipp* st, uint8_t a[32], uint8_t b[32]);
void bpg_stark_ipp_free(bpg_stark_ipp* st);

/* The fork's Stark-curve conventions that ARE in /root/reference (src/util.rs:252-289, src/generators.rs:80-125):
 * legacy Keccak-256 (`merlin::keccak256`), hash_to_scalar = (low || keccak256(low)) as a 512-bit little-endian
 * integer mod the group order (out: 32 bytes little-endian), and the generator chain
 *     state <- keccak256(state);  G_i = hash_to_scalar(state) * generator
 * from its initial state state0 = keccak256(pad_label("GeneratorsChain" || label)), which the caller forms
 * (`pad_label` belongs to the un-vendored merlin fork).  skip = GeneratorsChain::fast_forward.  The hash chain
 * runs on the host, the wide reductions and fixed-base multiplications on the device; out_xy: n x 64 bytes. */
void bpg_keccak256(const uint8_t* data, size_t len, uint8_t out[32]);
void bpg_stark_hash_to_scalar(const uint8_t low[32], uint8_t out[32]);
int bpg_stark_gens_chain(bpg_ctx* ctx, const uint8_t state0[32], size_t skip, size_t n, uint8_t* out_xy /* n*64 */);
/* building block of the chain: out[i] = ((wide[i] as a 512-bit little-endian integer) mod order) * generator */
int bpg_stark_wide_mul_generator(bpg_ctx* ctx, const uint8_t* wide64 /* n*64 */, size_t n, uint8_t* out_xy /* n*64 */);

/* ==== host mirror of the reference's protocol layer ==================================
 * C bindings of the C++ host code in mpc_bulletproof_b200/csrc/host/: the transcript,
 * the generators as resident tables, InnerProductProof and the R1CS Prover/Verifier with
 * the reference's method names, argument meaning and error behaviour.  Group work is done
 * by the entry points above; constraint-system bookkeeping and the transcript stay on the
 * CPU as in the reference. */

/* ---- transcript: merlin::Transcript + TranscriptProtocol (reference src/transcript.rs:25-121) */
typedef struct bpg_transcript bpg_transcript;
bpg_transcript* bpg_transcript_new(const uint8_t* label, size_t len);
bpg_transcript* bpg_transcript_clone(const bpg_transcript* t);
void bpg_transcript_free(bpg_transcript* t);
void bpg_transcript_append_message(bpg_transcript* t, const char* label, const uint8_t* msg, size_t len);
void bpg_transcript_append_u64(bpg_transcript* t, const char* label, uint64_t v);
void bpg_transcript_challenge_bytes(bpg_transcript* t, const char* label, uint8_t* out, size_t len);
void bpg_transcript_challenge_scalar(bpg_transcript* t, const char* label, uint8_t out[32]);

/* merlin::TranscriptRngBuilder / TranscriptRng (the prover's blinding source, reference
 * src/r1cs/prover.rs:435-445): build_rng clones the transcript state; rekey binds witness bytes;
 * finalize mixes in 32 bytes of external randomness; fill_bytes squeezes. */
typedef struct bpg_transcript_rng bpg_transcript_rng;
bpg_transcript_rng* bpg_transcript_build_rng(const bpg_transcript* t);
void bpg_transcript_rng_rekey_with_witness_bytes(bpg_transcript_rng* r, const char* label, const uint8_t* witness, size_t len);
void bpg_transcript_rng_finalize(bpg_transcript_rng* r, const uint8_t random_bytes[32]);
void bpg_transcript_rng_fill_bytes(bpg_transcript_rng* r, uint8_t* out, size_t len);
void bpg_transcript_rng_free(bpg_transcript_rng* r);

/* ---- generators: PedersenGens{B, B_blinding} + BulletproofGens party 0 (reference
 * src/generators.rs:32-71,158-235) uploaded once as ONE windowed table
 * [G (capacity) | H (capacity) | B | B_blinding] plus a comb for (B, B_blinding). */
typedef struct bpg_gens bpg_gens;
int bpg_gens_new(bpg_ctx* ctx, const uint8_t* G, const uint8_t* H, size_t capacity, const uint8_t B[32],
                 const uint8_t B_blinding[32], bpg_gens** out);
/* Generator derivation on the device, ristretto255 instantiation.
 * bpg_points_from_uniform: out[i] = element derivation (RFC 9496 4.3.4, dalek's from_uniform_bytes) of
 *   the i-th 64-byte block: the per-point work of GeneratorsChain::next (reference src/generators.rs:107-125).
 * bpg_gens_chain: points [skip, skip+n) of the chain SHAKE256("GeneratorsChain" || label) -- the XOF is
 *   squeezed on the host, `skip` is GeneratorsChain::fast_forward (src/generators.rs:80-100).
 * bpg_gens_derive: BulletproofGens::new(gens_capacity, ..).share(party) (src/generators.rs:182-235, labels
 *   "G"/"H" || u32le(party)) + PedersenGens::default() (src/generators.rs:61-71: B = basepoint,
 *   B_blinding = hash-to-group(SHA3-512(B))), resident as bpg_gens_new leaves them; the compressed
 *   generators are also returned where the pointers are not NULL. */
int bpg_points_from_uniform(bpg_ctx* ctx, const uint8_t* uniform64 /* n*64 */, size_t n, uint8_t* out_compressed /* n*32 */);
int bpg_gens_chain(bpg_ctx* ctx, const uint8_t* label, size_t label_len, size_t skip, size_t n, uint8_t* out /* n*32 */);
int bpg_gens_derive(bpg_ctx* ctx, size_t gens_capacity, uint32_t party, uint8_t* G_out, uint8_t* H_out,
                    uint8_t B_out[32], uint8_t Bb_out[32], bpg_gens** out);
void bpg_gens_free(bpg_gens* g);
size_t bpg_gens_capacity(const bpg_gens* g);
const bpg_table* bpg_gens_table(const bpg_gens* g);
/* PedersenGens::commit, batched: out[i] = values[i]*B + blindings[i]*B_blinding */
int bpg_pedersen_commit(bpg_ctx* ctx, const bpg_gens* g, const uint8_t* values, const uint8_t* blindings, size_t n,
                        uint8_t* out);

/* ---- InnerProductProof::create / ::verify (reference src/inner_product_proof.rs:49-193, 317-372).
 * Proof bytes: (L_j R_j)_j || a || b, 32 bytes each (:388-397).  create appends to the
 * transcript exactly as the reference does (dom-sep, n, then L, R per round). */
int bpg_ipp_create(bpg_ctx* ctx, bpg_transcript* t, const uint8_t Q[32], const uint8_t* G_factors,
                   const uint8_t* H_factors, const bpg_table* G, size_t g_off, const bpg_table* H, size_t h_off,
                   size_t n, const uint8_t* a, const uint8_t* b, uint8_t* proof_out, size_t proof_cap,
                   size_t* proof_len);
/* BPG_OK, BPG_ERR_VERIFY (ProofError::VerificationError) or BPG_ERR_DECODE (FormatError) */
int bpg_ipp_verify(bpg_ctx* ctx, bpg_transcript* t, size_t n, const uint8_t* G_factors, const uint8_t* H_factors,
                   const uint8_t P[32], const uint8_t Q[32], const bpg_table* G, size_t g_off, const bpg_table* H,
                   size_t h_off, const uint8_t* proof, size_t proof_len);

/* ---- r1cs::Prover / r1cs::Verifier (reference src/r1cs/prover.rs, src/r1cs/verifier.rs).
 * A bpg_cs is one of the two; both implement the ConstraintSystem methods
 * (src/r1cs/constraint_system.rs:55-208).  Variables are opaque 64-bit handles, linear
 * combinations are arrays of (variable, coefficient) terms. */
typedef struct bpg_cs bpg_cs;
typedef uint64_t bpg_var;
typedef struct {
  bpg_var var;
  uint8_t coeff[32];
} bpg_term;
typedef int (*bpg_randomized_cb)(bpg_cs* cs, void* user);
bpg_var bpg_var_one(void); /* Variable::One() */
int bpg_prover_new(bpg_ctx* ctx, const bpg_gens* gens, bpg_transcript* t, bpg_cs** out);   /* Prover::new   */
int bpg_verifier_new(bpg_ctx* ctx, const bpg_gens* gens, bpg_transcript* t, bpg_cs** out); /* Verifier::new */
void bpg_cs_free(bpg_cs* cs);
int bpg_prover_commit(bpg_cs* cs, const uint8_t v[32], const uint8_t v_blinding[32], uint8_t V_out[32], bpg_var* var);
/* n calls of bpg_prover_commit in one (same transcript, same variables; one batched fixed-base launch) */
int bpg_prover_commit_batch(bpg_cs* cs, const uint8_t* v /* n*32 */, const uint8_t* v_blinding /* n*32 */, size_t n,
                            uint8_t* V_out /* n*32 */, bpg_var* vars /* n */);
int bpg_verifier_commit(bpg_cs* cs, const uint8_t V[32], bpg_var* var);
int bpg_cs_commit_public(bpg_cs* cs, const uint8_t value[32], bpg_var* var);
int bpg_cs_multiply(bpg_cs* cs, const bpg_term* left, size_t nl, const bpg_term* right, size_t nr, bpg_var out[3]);
int bpg_cs_allocate(bpg_cs* cs, const uint8_t* assignment /* NULL on the verifier */, bpg_var* out);
int bpg_cs_allocate_multiplier(bpg_cs* cs, const uint8_t* l, const uint8_t* r, bpg_var out[3]);
int bpg_cs_constrain(bpg_cs* cs, const bpg_term* lc, size_t n);
int bpg_cs_specify_randomized_constraints(bpg_cs* cs, bpg_randomized_cb cb, void* user);
int bpg_cs_challenge_scalar(bpg_cs* cs, const char* label, uint8_t out[32]); /* only inside a callback */
int bpg_cs_eval(bpg_cs* cs, const bpg_term* lc, size_t n, uint8_t out[32]);
/* The reference's benchmark circuit (benches/r1cs.rs:24-32): n chained squarings starting from `var`;
 * equivalent to n calls of bpg_cs_multiply(var, var).  out (may be NULL) = the last output variable. */
int bpg_gadget_square_chain(bpg_cs* cs, bpg_var var, size_t n, bpg_var* out);
/* The shuffle gadget of the reference's bench and tests (benches/shuffle.rs:30-69, tests/r1cs.rs:22-63): y[0..k) is a
 * permutation of x[0..k), by the randomized product check (one deferred callback, 2 (k - 1) multipliers; k = 1: the
 * constraint y_0 - x_0).  The same constraint system as the callback form built through bpg_cs_*. */
int bpg_gadget_shuffle(bpg_cs* cs, const bpg_var* x, const bpg_var* y, size_t k);
/* BASELINE.json config 4: n_mult multipliers with uniform a_L, a_R and n_cons random
 * linear constraints over (a_L, a_R, a_O, v) with constants c0 fixed from the witness; everything from
 * xoshiro256**(seed).  c0: n_cons x 32 bytes, written by the prover, read by the verifier. */
int bpg_gadget_random_circuit(bpg_cs* cs, uint64_t seed, size_t n_mult, size_t n_cons, uint8_t* c0);
size_t bpg_cs_num_multipliers(const bpg_cs* cs);
size_t bpg_cs_num_constraints(const bpg_cs* cs);
/* Prover::prove (prover.rs:412-727).  BPG_ERR_CAPACITY = InvalidGeneratorsLength.
 * Blinding scalars as in the reference (prover.rs:435-445): a merlin TranscriptRng forked from the
 * transcript after "m", rekeyed with every v_blinding and finalized with 32 bytes of external
 * randomness, drawn in the reference's order (:457-462, 519-530, 621-625).  The two blinding vectors
 * of a phase, s_L and s_R, are expanded on the device from ONE 32-byte draw of that RNG (ChaCha20
 * blocks 2j / 2j+1 reduced mod l) instead of 2n sequential draws.
 *   bpg_prover_prove                 the 32 bytes come from the operating system (getrandom): THE entry
 *                                    point for production use;
 *   bpg_prover_prove_with_rng_bytes  the caller supplies them (its own CSPRNG; fixed bytes make the
 *                                    proof reproducible -- parity tests);
 *   bpg_prover_prove_deterministic   TEST / BENCH ONLY: every blinding from xoshiro256**(rng_seed), a
 *                                    64-bit non-cryptographic seed.  Reusing a seed leaks the witness. */
int bpg_prover_prove(bpg_cs* cs, uint8_t* proof_out, size_t proof_cap, size_t* proof_len);
int bpg_prover_prove_with_rng_bytes(bpg_cs* cs, const uint8_t rng_bytes[32], uint8_t* proof_out, size_t proof_cap,
                                    size_t* proof_len);
int bpg_prover_prove_deterministic(bpg_cs* cs, uint64_t rng_seed, uint8_t* proof_out, size_t proof_cap,
                                   size_t* proof_len);
/* Verifier::verify: BPG_OK, BPG_ERR_VERIFY, BPG_ERR_DECODE (FormatError), BPG_ERR_CAPACITY.
 * bpg_verifier_verify derives the scalar r that batches the t(x) check with the inner-product check
 * as the mounted fork does, r = challenge_scalar("r") (verifier.rs:506): a public function of the
 * transcript.  bpg_verifier_verify_with_rng_bytes draws r from a TranscriptRng finalized with 32 bytes
 * the prover cannot know (rng_bytes; NULL = from the operating system), as upstream's
 * `build_rng().finalize(&mut thread_rng())` does: same accept set, and r is unpredictable to whoever
 * produced the proof.  Use it whenever proofs come from an untrusted party. */
int bpg_verifier_verify(bpg_cs* cs, const uint8_t* proof, size_t proof_len);
int bpg_verifier_verify_with_rng_bytes(bpg_cs* cs, const uint8_t* proof, size_t proof_len, const uint8_t* rng_bytes);
/* Many proofs, proof by proof as the reference does (verifier.rs:393); every verifier is consumed.
 * ok[i] = 1 iff proof i verifies (a malformed proof is a reject).  Non-zero return only for
 * failures of the machinery.  Across GPUs whole proofs are sharded by the caller (one context
 * per GPU); there is no data-path collective. */
int bpg_batch_verify(bpg_cs* const* verifiers, const uint8_t* const* proofs, const size_t* proof_lens, size_t n,
                     uint8_t* ok);

#ifdef __cplusplus
}
#endif
#endif /* BPGPU_H */
