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
size_t bpg_table_len(const bpg_table* t);
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

#ifdef __cplusplus
}
#endif
#endif /* BPGPU_H */
