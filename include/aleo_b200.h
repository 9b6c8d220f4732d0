/* aleo_b200.h -- C ABI of libaleo_b200.so: BLS12-377 G1 MSM and Fr NTT on NVIDIA B200 (sm_100a).
 *
 * The reference (demox-labs/aleo = Aleo SDK 0.5.1) has no FFI of its own for this path: it reaches
 * the arithmetic through the crates.io pin snarkvm-algorithms "=0.14.5" (reference Cargo.toml:28-53).
 * The entry points below are what a `[patch.crates-io] snarkvm-algorithms` shim binds (see
 * INTEGRATION.md and rust_shim/); each one names the snarkVM function it stands in for and the
 * reference call site that reaches it.  Data contracts:
 *   Fr element      : 32 B, Montgomery form (R = 2^256), 4 little-endian u64 limbs, fully reduced.
 *   MSM scalar      : 32 B, canonical (non-Montgomery) BigInteger256, little-endian, < r.
 *   G1Affine        : x (48 B) | y (48 B) Montgomery (R = 2^384) little-endian; with
 *                     affine_stride == 104 a bool `infinity` follows at byte 96 (Rust layout,
 *                     size_of::<G1Affine>()); with affine_stride == 96 (packed) the identity is
 *                     encoded as x = y = 0.
 *   G1Projective    : Jacobian x | y | z, 3 x 48 B Montgomery.  Results are returned NORMALISED:
 *                     z = 1 (Montgomery one), or (0, 1, 0) for the identity, so that to_affine()
 *                     on the Rust side is a copy.
 * All functions return 0 on success and a negative ALEO_B200_E* code otherwise; nothing throws,
 * aborts or falls back to the CPU.  Every function is re-entrant: calls may come concurrently from
 * many host threads (snarkVM commits polynomials on a rayon ExecutionPool); host-pointer calls
 * use an internal per-thread stream, *_dev calls run on the caller's stream.
 */
#ifndef ALEO_B200_H
#define ALEO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  ALEO_B200_OK = 0,
  ALEO_B200_EINVAL = -1,     /* null pointer, bad enum value, bad stride                         */
  ALEO_B200_ETOOLARGE = -2,  /* log_n above the supported maximum (32; the field allows 47)      */
  ALEO_B200_ENODEVICE = -3,  /* no CUDA device, or the device is not sm_100 (B200): no fallback  */
  ALEO_B200_ECUDA = -4,      /* a CUDA runtime call or kernel failed; see aleo_b200_last_cuda_error */
  ALEO_B200_ENOMEM = -5      /* device or pinned-host allocation failed                          */
};

/* enums mirror upstream's optional cuda FFI so that a shim is a few lines */
enum { ALEO_B200_NTT_FORWARD = 0, ALEO_B200_NTT_INVERSE = 1 };
enum { ALEO_B200_NTT_STANDARD = 0, ALEO_B200_NTT_COSET = 1 };

const char* aleo_b200_version(void);
const char* aleo_b200_strerror(int code);
/* text of the last CUDA error seen by the calling thread ("" if none) */
const char* aleo_b200_last_cuda_error(void);

/* Selects `device` for the calling thread and prepares it (uploads field constants).  Optional:
 * every other entry point prepares the calling thread's current device on first use. */
int aleo_b200_init(int device);
int aleo_b200_device_count(void);
/* Frees cached plans / workspaces of the current device. */
int aleo_b200_shutdown(void);

/* ---- Fr NTT --------------------------------------------------------------------------------
 * Stands in for snarkvm_algorithms::fft::EvaluationDomain::{fft_in_place, ifft_in_place,
 * coset_fft_in_place, coset_ifft_in_place} (snarkvm-algorithms 0.14.5 src/fft/domain.rs), reached
 * from the reference at rust/src/program/execute.rs:74,177,219 and rust/src/program/transfer.rs:99
 * (prove_execution / vm.execute) and rust/src/program/deploy.rs:142 (vm.deploy: index polynomials).
 * inout holds n = 2^log_n elements (the caller has already zero-padded, as Vec::resize does
 * upstream); natural order in, natural order out.  log_n = 0 is the identity. */
int aleo_b200_ntt_fr(void* inout_host, uint32_t log_n, int direction, int kind);
/* Same on device memory, asynchronous on `stream` (a cudaStream_t; NULL = default stream).
 * `batch` transforms of the same size, contiguous, each in place. */
int aleo_b200_ntt_fr_dev(void* inout_dev, uint32_t log_n, size_t batch, int direction, int kind, void* stream);
/* Out-of-order variants (snarkVM's FFTOrder, src/fft/domain.rs: fft_helper_in_place_with_pc / ifft_helper_in_place_with_pc /
 * out_order_fft_in_place_with_pc, which the Varuna prover calls with its cached FFTPrecomputation; SURVEY.md 8a
 * row 11 -- the precomputation itself is not needed here: twiddle tables are cached per size inside the library).
 *   II: natural order in, natural order out (= aleo_b200_ntt_fr_dev);  IO: natural in, bit-reversed out
 *   (out[bitrev(i)] = X[i]);  OI: bit-reversed in, natural out.  Coset scaling always refers to the natural index. */
enum { ALEO_B200_NTT_ORDER_II = 0, ALEO_B200_NTT_ORDER_IO = 1, ALEO_B200_NTT_ORDER_OI = 2 };
int aleo_b200_ntt_fr_ordered_dev(void* inout_dev, uint32_t log_n, size_t batch, int direction, int kind, int order, void* stream);
/* The same on a HOST buffer: the full argument list of upstream's optional FFI `snarkvm_ntt(inout, lg_domain_size, order,
 * direction, ntt_type)` (upstream's NN / NR / RN are II / IO / OI here).  One H2D, the transform, one D2H. */
int aleo_b200_ntt_fr_ordered(void* inout_host, uint32_t log_n, int direction, int kind, int order);
/* Product of polynomials over the domain of size n = 2^log_n, shaped like upstream's `snarkvm_polymul(out, pcount,
 * polynomials, plens, ecount, evaluations, elens, lg_domain_size)` (the CUDA arm of snarkVM's
 * PolyMultiplier::multiply, src/fft/polynomial/multiplier.rs [U]):
 *     out = ifft( prod_i fft(pad_n(polynomials[i])) * prod_j evaluations[j] )          (n coefficients, natural order)
 * polynomials[i]: plens[i] <= n coefficients; evaluations[j]: elens[j] == n values on the domain; all Montgomery Fr;
 * pcount + ecount >= 1, each <= 64.  The host call moves every operand over PCIe once and the result back once -- a
 * prover that called fft / mul / ifft through the host NTT entry would pay 2 (pcount + 1) transfers of n elements instead
 * (aleo_b200_ntt_fr at 2^24: 23 ms per call against 3.3 ms on the device).  `_dev`: HOST arrays of device pointers,
 * asynchronous on `stream`; out_dev may not alias an input. */
int aleo_b200_polymul(void* out_host, size_t pcount, const void* const* polynomials_host, const size_t* plens, size_t ecount,
                      const void* const* evaluations_host, const size_t* elens, uint32_t log_n);
int aleo_b200_polymul_dev(void* out_dev, size_t pcount, const void* const* polynomials_dev_ptrs_host, const size_t* plens_host,
                          size_t ecount, const void* const* evaluations_dev_ptrs_host, const size_t* elens_host, uint32_t log_n,
                          void* stream);
/* One transform with CUDA events around every pass: pass_ms4[i] = device time of pass i (unused
 * entries 0).  Synchronises `stream`.  Measurement aid for bench.py's roofline, same kernels. */
int aleo_b200_ntt_fr_dev_profile(void* inout_dev, uint32_t log_n, int direction, int kind, void* stream, float* pass_ms4);
/* Multi-GPU four-step NTT, twiddle step between the two local stages: for a rows x cols row-major
 * matrix of Fr on the device, data[r][c] *= w^((r + row0) * (c + col0)) with w the root of the GLOBAL
 * domain of size 2^log_n_global (its inverse for ALEO_B200_NTT_INVERSE).  12 <= log_n_global <= 32. */
int aleo_b200_ntt_twiddle_dev(void* data_dev, uint32_t log_n_global, int direction, uint32_t rows, uint32_t cols,
                              uint32_t row0, uint32_t col0, void* stream);
/* ---- one Fr NTT spread over the GPUs of a node (one process per GPU), exchange fused into the transform ---------
 * BASELINE config 4 / SURVEY.md 8e: the four-step decomposition with ONE exchange.  The pass split is the single-GPU
 * one (radices 2^K_0 .. 2^K_last); the transposing exchange is not a separate collective: the last-but-one pass
 * stores every result element straight into the receive buffer of the rank that needs it (peer memory over
 * NVLink, CUDA IPC between the processes), so the transfer overlaps the butterflies tile by tile.
 *   input  (rank j): column block of the (N / R_last) x R_last row-major matrix of x:
 *                    local[a * R_last / g + b] = x[a * R_last + j * R_last / g + b]
 *   output (rank t): column block of the (N / R_0) x R_0 row-major matrix of X (natural order):
 *                    local[a * R_0 / g + b] = X[a * R_0 + t * R_0 / g + b]
 * Usage, every rank, same call sequence:  ctx = create;  gather every rank's 128-byte `handles` (e.g. an
 * all_gather);  open(all handles);  then per transform: transform(in, out).  No collective runs per transform: the CTA
 * that finishes the exchange pass last raises this rank's flag in every rank's receive buffer (st.release.sys after a
 * system fence, over NVLink) and the last pass spins on its own `world` flags (ld.acquire.sys) before it reads what the
 * peers stored.  Two receive buffers alternate, so the next transform's first stage may be enqueued at once.
 * stage1 / stage2 are the two halves of transform() (a caller may put other work between them); profile() is
 * transform() with events: stage_ms4 = {local passes, exchange pass, wait for the peers' flags, last pass}, synchronises.
 * world = 1, 2, 4, 8; all four transform kinds (the coset powers are taken at the global index). */
int aleo_b200_ntt_dist_layout(uint32_t log_n, int world, uint32_t* log_r_first_out, uint32_t* log_r_last_out, int* passes_out);
int aleo_b200_ntt_dist_create(void** ctx_out, uint32_t log_n, int rank, int world);
int aleo_b200_ntt_dist_handles(void* ctx, void* handles128_out);
int aleo_b200_ntt_dist_open(void* ctx, const void* all_handles /* world x 128 bytes, rank-major */);
int aleo_b200_ntt_dist_stage1(void* ctx, const void* local_in_dev, int direction, int kind, void* stream);
int aleo_b200_ntt_dist_stage2(void* ctx, void* local_out_dev, int direction, int kind, void* stream);
int aleo_b200_ntt_dist_transform(void* ctx, const void* local_in_dev, void* local_out_dev, int direction, int kind, void* stream);
int aleo_b200_ntt_dist_profile(void* ctx, const void* local_in_dev, void* local_out_dev, int direction, int kind, void* stream,
                               float* stage_ms4);
int aleo_b200_ntt_dist_destroy(void* ctx);
/* kernel launches one transform of this size issues (for launch accounting) */
int aleo_b200_ntt_launches(uint32_t log_n);

/* ---- G1 MSM --------------------------------------------------------------------------------
 * Stands in for snarkvm_algorithms::msm::VariableBase::msm::<G1Affine>(bases, scalars)
 * (snarkvm-algorithms 0.14.5 src/msm/variable_base/mod.rs) as called by KZG10::commit /
 * commit_lagrange / open (src/polycommit/kzg10/mod.rs), reached from the same reference call sites.
 * Like upstream the two slices are zipped: n is the shorter length.  n = 0 gives the identity.
 * The bases are points of the prime-order subgroup G1, which is what a snarkVM G1Affine is (deserialisation checks it,
 * the group law preserves it): up to 2^20 points the scalars are split with the curve's endomorphism (GLV), which is
 * multiplication by -u^2 on G1 only.  ALEO_B200_MSM_GLV=0 turns the split off for arbitrary curve points. */
int aleo_b200_msm_g1(void* out_projective_host, const void* bases_host, size_t n, const void* scalars_host,
                     size_t affine_stride);
/* The same MSM spread over the first `n_devices` GPUs of the node from ONE process (how a single prover process uses
 * the box; BASELINE config 4 / SURVEY.md 8e): contiguous point ranges, one host thread and stream per device, each
 * range copied and accumulated as in aleo_b200_msm_g1, and a single final combine of the 144-byte partial sums on
 * device 0.  n_devices <= aleo_b200_device_count(); 1 is aleo_b200_msm_g1. */
int aleo_b200_msm_g1_multi(void* out_projective_host, const void* bases_host, size_t n, const void* scalars_host,
                           size_t affine_stride, int n_devices);
/* Device-resident operands (bases typically stay resident: the SRS is reused by every commitment).
 * out_projective_dev: 144 B of device memory.  Asynchronous on `stream`. */
int aleo_b200_msm_g1_dev(void* out_projective_dev, const void* bases_dev, size_t n, const void* scalars_dev,
                         size_t affine_stride, void* stream);
/* One MSM with CUDA events around its phases: phase_ms3 = {recode + counting sort + task plan,
 * bucket accumulation kernel, combine + bucket reduction + final}.  Synchronises `stream`. */
int aleo_b200_msm_g1_dev_profile(void* out_projective_dev, const void* bases_dev, size_t n, const void* scalars_dev,
                                 size_t affine_stride, void* stream, float* phase_ms3);
/* Sum of `count` Jacobian points (144 B each, device) -> one normalised Jacobian point (device).
 * This is the single final combine of the point-range-sharded multi-GPU MSM. */
int aleo_b200_g1_sum_dev(void* out_projective_dev, const void* points_dev, size_t count, void* stream);
/* ---- resident SRS / KZG10::commit ---------------------------------------------------------------
 * snarkVM commits every polynomial against the SAME bases (`powers_of_beta_g`, or its Lagrange form):
 * KZG10::commit(powers, polynomial, ..) = VariableBase::msm(&powers.powers_of_beta_g[..d+1], coeffs.to_bigint())
 * (snarkvm-algorithms 0.14.5 src/polycommit/kzg10/mod.rs).  An SRS
 * handle keeps the bases on the device, expanded once to 2^(c w) * P_i for every window w, so that a
 * commitment needs one shared bucket set, ~30 % fewer additions (larger windows; scalars above (r-1)/2 are
 * recoded as r - s on the negated point, so 2^24 points need 11 windows of 23 bits) and no doubling tail.
 * Device memory: n * ceil(253 / c) * 96 bytes (c by cost model, 8..23: 2^16 -> 15, 2^24 -> 23).  A handle belongs to the
 * device that was current when it was created; MSMs over any PREFIX of the bases are supported. */
int aleo_b200_srs_create(void** handle_out, const void* bases_host, size_t n, size_t affine_stride);
int aleo_b200_srs_create_dev(void** handle_out, const void* bases_dev, size_t n, size_t affine_stride, void* stream);
int aleo_b200_srs_destroy(void* handle);
int aleo_b200_srs_info(const void* handle, size_t* n_out, int* window_bits_out, int* windows_out, size_t* bytes_out);
/* sum_i scalars[i] * P_i over the first n_used bases; scalars canonical 32-byte LE */
int aleo_b200_srs_msm(const void* handle, void* out_projective_host, const void* scalars_host, size_t n_used);
int aleo_b200_srs_msm_dev(const void* handle, void* out_projective_dev, const void* scalars_dev, size_t n_used, void* stream);
int aleo_b200_srs_msm_dev_profile(const void* handle, void* out_projective_dev, const void* scalars_dev, size_t n_used,
                                  void* stream, float* phase_ms3);
int aleo_b200_srs_msm_launches(const void* handle, size_t n_used);
/* KZG10::commit without hiding: coefficients are Fr in MONTGOMERY form (the polynomial as snarkVM holds
 * it); converts them to canonical integers on the device (to_bigint), runs the MSM and returns the
 * commitment as the 48-byte compressed G1 wire format (x LE | bit 383: larger y | bit 382: infinity). */
int aleo_b200_kzg_commit(const void* handle, void* out_compressed48_host, const void* coeffs_montgomery_host, size_t n_coeffs);
int aleo_b200_kzg_commit_dev(const void* handle, void* out_compressed48_dev, const void* coeffs_montgomery_dev, size_t n_coeffs,
                             void* stream);
/* KZG10::commit with a hiding bound: commitment = msm(powers_of_beta_g, coeffs) + msm(powers_of_beta_times_gamma_g,
 * random_coeffs), the second polynomial being the blinding polynomial the caller drew (src/polycommit/kzg10/mod.rs
 * `commit`: `random_ints` against `powers_of_beta_times_gamma_g`).  Two resident handles, both
 * vectors Montgomery Fr on the device, result 48-byte compressed. */
int aleo_b200_kzg_commit_hiding_dev(const void* handle_beta, const void* handle_beta_gamma, void* out_compressed48_dev,
                                    const void* coeffs_montgomery_dev, size_t n_coeffs, const void* random_coeffs_montgomery_dev,
                                    size_t n_random, void* stream);
/* `count` (<= 64) commitments against the same resident SRS in ONE launch sequence: snarkVM commits the ~13 polynomials
 * of a proof against the same powers (SonicKZG10::commit over an ExecutionPool, src/polycommit/sonic_pc/mod.rs;
 * SURVEY.md 8f rank 1).  The coefficient vectors are sorted together, accumulated by one kernel (every polynomial
 * owns a bucket set), and each result is normalised and compressed by its own CTA, so the latency-bound tails of
 * proof-sized MSMs (2^15 .. 2^18) run side by side instead of one after the other.
 * coeffs_dev_ptrs_host / n_coeffs_host: HOST arrays of `count` device pointers / lengths (Montgomery Fr);
 * out_compressed48_dev: count x 48 bytes. */
int aleo_b200_kzg_commit_batch_dev(const void* handle, void* out_compressed48_dev, const void* const* coeffs_dev_ptrs_host,
                                   const size_t* n_coeffs_host, size_t count, void* stream);
/* window size c (bits) the MSM uses for n device-resident points, and the kernel launches one MSM issues */
int aleo_b200_msm_window_bits(size_t n);
int aleo_b200_msm_launches(size_t n);
/* batch-affine pair-tree levels (csrc/msm_ba.cuh) the MSM runs in front of its XYZZ accumulation for n device-resident
 * points (0 below 2^21 points): after L levels 1 - 2^-L of the additions were done at 5 products + 1 square instead of
 * 8 + 2 -- bench.py needs it to count the multiplies the accumulation phase issues */
int aleo_b200_msm_ba_levels(size_t n);
/* The host-pointer entry points (aleo_b200_msm_g1, aleo_b200_srs_msm, aleo_b200_kzg_commit) copy and accumulate
 * the MSM in point ranges so that the host->device copy of range k+1 overlaps the accumulation of range k:
 * how many ranges n points are cut into for PAGEABLE caller memory (what a Rust Vec is; pinned buffers use one range
 * fewer from 2^22 points), and the window size that goes with it. */
int aleo_b200_msm_host_plan(size_t n, int* ranges_out, int* window_bits_out);

/* ---- elementwise field arithmetic on device vectors ------------------------------------------------
 * out[i] = a[i] op b[i] over n elements in Montgomery form (Fr: 32 B, Fq: 48 B; in-place allowed).  Stands in
 * for the pointwise arithmetic snarkVM's prover does on evaluation vectors between its FFTs
 * (snarkvm-algorithms 0.14.5 src/fft/evaluations.rs Mul / Sub / Add, snarkvm-fields batch_inversion;
 * SURVEY.md 8f rank 2), reached from the reference at rust/src/program/execute.rs:74,219, so that prover
 * rounds can stay on the device between transforms.  b is ignored for SQR / INV / NEG; INV maps 0 to 0
 * (batch_inversion leaves zeros untouched).  Also what the parity tests use to pin the field core. */
enum { ALEO_B200_FIELD_FR = 0, ALEO_B200_FIELD_FQ = 1 };
enum { ALEO_B200_OP_ADD = 0, ALEO_B200_OP_SUB = 1, ALEO_B200_OP_MUL = 2, ALEO_B200_OP_SQR = 3, ALEO_B200_OP_INV = 4,
       ALEO_B200_OP_NEG = 5 };
int aleo_b200_field_op_dev(int field, int op, void* out_dev, const void* a_dev, const void* b_dev, size_t n, void* stream);

/* Polynomial helpers of the KZG / Varuna call sites; Fr vectors in Montgomery form on the device, the scalar
 * arguments (g, k, z) are 32-byte Montgomery Fr on the HOST.
 *   distribute_powers : a[i] *= k * g^i   (EvaluationDomain::distribute_powers[_and_mul_by_const], src/fft/domain.rs;
 *                       k_host NULL = 1) -- the coset shift of a polynomial by an arbitrary generator
 *   poly_eval         : out = sum_i c[i] z^i   (DensePolynomial::evaluate, src/fft/polynomial/dense.rs), 32 bytes
 *   divide_by_linear  : q(x) = (p(x) - p(z)) / (x - z), the witness polynomial of KZG10::open
 *                       (KZG10::compute_witness_polynomial, src/polycommit/kzg10/mod.rs); quotient_dev holds n
 *                       elements, q[n-1] = 0; must not alias coeffs_dev
 *   kzg_open          : commitment (48-byte compressed G1) to that witness polynomial against the resident SRS =
 *                       the non-hiding part of KZG10::open */
/*   axpy              : y[i] += a * x[i]: the linear combination sum_i xi^i p_i of SonicKZG10::open_combinations /
 *                       batch_open (src/polycommit/sonic_pc/mod.rs); a: 32-byte Montgomery Fr on the host */
/*   lagrange_coeffs   : out[i] = L_i(tau) over the domain of size 2^log_n (EvaluationDomain::
 *                       evaluate_all_lagrange_coefficients, src/fft/domain.rs: (tau^n - 1) / n * w^i / (tau - w^i), the
 *                       indicator vector when tau lies in the domain); synchronises the stream once */
/*   divide_by_vanishing_on_coset : evals[i] /= Z_n(g w_m^i), Z_n(x) = x^n - 1, for the evaluations of a polynomial on the
 *                       coset g H_m (m = 2^log_m >= n = 2^log_n; g = 22 for coset_fft): EvaluationDomain::
 *                       divide_by_vanishing_poly_on_coset_in_place (src/fft/domain.rs) when m = n, and the quotient by the
 *                       constraint-domain vanishing polynomial over a larger multiplication domain in the prover rounds;
 *                       the divisor has period m / n, so the kernel inverts m / n values and multiplies */
int aleo_b200_fr_divide_by_vanishing_on_coset_dev(void* evals_inout_dev, uint32_t log_m, uint32_t log_n, const void* g_host,
                                                  void* stream);
int aleo_b200_fr_lagrange_coeffs_dev(void* out_dev, uint32_t log_n, const void* tau_host, void* stream);
int aleo_b200_fr_axpy_dev(void* y_inout_dev, const void* x_dev, const void* a_host, size_t n, void* stream);
int aleo_b200_fr_distribute_powers_dev(void* inout_dev, size_t n, const void* g_host, const void* k_host, void* stream);
int aleo_b200_fr_poly_eval_dev(void* out_dev, const void* coeffs_dev, size_t n, const void* z_host, void* stream);
int aleo_b200_fr_divide_by_linear_dev(void* quotient_dev, const void* coeffs_dev, size_t n, const void* z_host, void* stream);
int aleo_b200_kzg_open_dev(const void* handle, void* out_compressed48_dev, const void* coeffs_montgomery_dev, size_t n_coeffs,
                           const void* z_host, void* stream);
/* m openings in one call -- the shape of SonicKZG10::open_combinations / kzg10 batch_open (src/polycommit/sonic_pc/mod.rs;
 * SURVEY.md 8a row 14): opening k is of the linear combination p_k = sum_i lc[k][i] * poly_i at the point z_k, its
 * witness polynomial (p_k(x) - p_k(z_k)) / (x - z_k) is built on the device, and the m witness commitments go through
 * ONE launch sequence (the batch path of aleo_b200_kzg_commit_batch_dev).  polys_dev_ptrs_host / n_coeffs_host: HOST
 * arrays of n_polys device pointers / lengths (Montgomery Fr); lc_coeffs_host: m x n_polys Montgomery Fr on the host,
 * row-major, a zero coefficient = the polynomial is not part of that combination; points_host: m x 32 bytes;
 * out_compressed48_dev: m x 48 bytes.  m <= 64, n_polys <= 1024.  (Hiding openings add the caller's random_v.) */
int aleo_b200_kzg_open_combinations_dev(const void* handle, void* out_compressed48_dev, const void* const* polys_dev_ptrs_host,
                                        const size_t* n_coeffs_host, size_t n_polys, const void* lc_coeffs_host,
                                        const void* points_host, size_t m, void* stream);

/* ---- compressed G1 wire format ---------------------------------------------------------
 * snarkVM's CanonicalSerialize / CanonicalDeserialize of G1Affine -- every commitment inside a proof and the points of
 * key / SRS files: 48 bytes, x canonical little-endian, bit 383 = y is the larger root, bit 382 = infinity.  Pinned by
 * the reference's own proof string (wasm/src/programs/transaction.rs:100; tests/golden/proof_fixture.json).
 * decompress: n x 48 bytes -> n affine points (stride 104 / 96, Montgomery); returns the number of invalid encodings
 * (stored as the identity), < 0 on error; synchronises the stream.  Invalid = x >= p, x^3 + 1 not a square, or -- as
 * deserialize_compressed (Validate::Yes -> is_in_correct_subgroup_assuming_on_curve) rejects it upstream -- a curve
 * point outside the prime-order subgroup (BLS12-377's G1 cofactor is large); the subgroup test is phi(P) == -[u^2] P.
 * decompress_unchecked: deserialize_compressed_unchecked (Validate::No): no subgroup test, for trusted SRS / key files.
 * compress: the inverse, asynchronous. */
int aleo_b200_g1_decompress_dev(void* out_affine_dev, size_t affine_stride, const void* in48_dev, size_t n, void* stream);
int aleo_b200_g1_decompress_unchecked_dev(void* out_affine_dev, size_t affine_stride, const void* in48_dev, size_t n, void* stream);
int aleo_b200_g1_compress_dev(void* out48_dev, const void* affine_dev, size_t affine_stride, size_t n, void* stream);

/* ---- synthetic workload generation and on-device checks (bench / tests) --------------------
 * bases[i] = (s0 + (first_index + i) * d) * G, G the G1 generator; s0, d canonical 32-byte scalars.
 * Known discrete logs make the MSM result checkable at any size. */
int aleo_b200_gen_bases_dev(void* bases_dev, size_t n, size_t affine_stride, const void* s0_host, const void* d_host,
                            uint64_t first_index, void* stream);
/* scalars[i] = splitmix-style hash of (seed, first_index + i) masked to 252 bits (< r), canonical */
int aleo_b200_gen_scalars_dev(void* scalars_dev, size_t n, uint64_t seed, uint64_t first_index, int montgomery,
                              void* stream);
/* out (32 B canonical, device) = sum_i scalars[i] * (s0 + (first_index + i) * d) mod r */
int aleo_b200_dlog_dot_dev(void* out_scalar_dev, const void* scalars_dev, size_t n, const void* s0_host,
                           const void* d_host, uint64_t first_index, void* stream);
/* 1 if every point is on y^2 = x^3 + 1 (or the identity), 0 if not, < 0 on error */
int aleo_b200_check_on_curve_dev(const void* bases_dev, size_t n, size_t affine_stride, void* stream);
/* dependent-free IMAD / IMAD.WIDE issue-rate microbenchmark: returns elapsed ms for
 * `iters` x 64 chains per thread on a full-chip grid, and the op count through *ops_out.
 * kind 0 = mad.lo (IMAD), 1 = 32x32+64 (IMAD.WIDE), 2 = mad.lo.cc/madc.hi.cc carry chains (IMAD.WIDE.U32.X),
 * 3 = dependent Fq products, 4 = DFMA chains (FP64 pipe), 5 = DFMA and IMAD.WIDE interleaved one to one
 * (ops = pairs): do the two pipes overlap?  6 / 7 = the dependent Fq product / square chains of kind 3 through the
 * FP64-pipe multiplier (csrc/mont_fp64.cuh; ops counted as 276-MAC products, so the rates compare directly), 8 = dedicated
 * IMAD.WIDE squares. */
int aleo_b200_bench_imad(int kind, int iters, double* ms_out, double* ops_out);
/* EXPERIMENT: out[i] = a[i] * b[i] (square != 0: a[i]^2) over Fq, 48-byte Montgomery images, through
 * the FP64-pipe (DFMA, 48-bit limbs) multiplier -- bit-identical to aleo_b200_field_op_dev(FQ, MUL / SQR); parity tests
 * and the A/B measurement use it. */
int aleo_b200_fq_mul_fp64_dev(void* out_dev, const void* a_dev, const void* b_dev, size_t n, int square, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ALEO_B200_H */
