/*
 * uzkge_cuda.h -- C ABI of the B200 (sm_100a) MSM / NTT backend for zypher-game/uzkge.
 *
 * The reference has no FFI today; this header is the seam a thin Rust `extern "C"` crate
 * binds (INTEGRATION.md shows the crate and the two patched call sites).  Each entry point names the
 * reference code it replaces.
 *
 * Data layout (identical to arkworks' in-memory types, so Rust copies limbs, never converts):
 *   field element  : 4 x uint64 little-endian limbs, MONTGOMERY form, R = 2^256   (ark_ff::Fp, BigInt<4>)
 *   affine G1      : x[4], y[4]  (64 B).  The identity is x = y = 0 and is skipped
 *                    (the padded monomial SRS holds identities, uzkge/src/gen_params/mod.rs:160-161)
 *   Jacobian G1    : X[4], Y[4], Z[4]  (96 B), Z == 0 <=> identity               (G1Projective)
 *
 * Ownership: the caller owns every host buffer; the library copies in / out and owns all device memory.
 * Errors: 0 = OK, otherwise one of UZKGE_ERR_*; uzkge_cuda_last_error() gives a thread-local message.
 * There is NO CPU fallback: without a usable CUDA device every call fails with UZKGE_ERR_NO_DEVICE.
 * Threading: every entry point may be called from any thread; calls are serialised PER DEVICE (one mutex and one internal stream per
 * GPU), so threads working on different GPUs run concurrently.
 * Devices: a call goes to the device of the SRS handle it names (handles carry their device), otherwise to the device the calling
 * thread selected with uzkge_cuda_init / uzkge_cuda_set_device, otherwise to the first device initialised.  One process can drive all
 * GPUs of the box: uzkge_cuda_init_devices + uzkge_cuda_srs_upload_multi (below); one process per GPU (uzkge_b200/dist.py) works too.
 * Streams: the *_device entry points enqueue on the caller's stream.  Workspaces shared by all streams of a device (the MSM buffers
 * of an SRS handle, the scan workspace of the polynomial calls, the small result buffer) are ordered internally: when consecutive
 * calls use different streams the later stream waits (cudaStreamWaitEvent) for the earlier call's work; nothing blocks the host.
 */
#ifndef UZKGE_CUDA_H
#define UZKGE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define UZKGE_API __attribute__((visibility("default")))
#else
#define UZKGE_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define UZKGE_OK 0
#define UZKGE_ERR_NO_DEVICE 1   /* no CUDA device / driver: the backend refuses to run (no CPU path)   */
#define UZKGE_ERR_SIZE 2        /* unsupported domain size, length mismatch, range outside the SRS     */
#define UZKGE_ERR_CUDA 3        /* CUDA runtime failure (message in uzkge_cuda_last_error)             */
#define UZKGE_ERR_OOM 4         /* device allocation failed                                            */
#define UZKGE_ERR_HANDLE 5      /* unknown SRS handle                                                  */
#define UZKGE_ERR_ARG 6         /* null pointer / bad flag                                             */
#define UZKGE_ERR_INTERNAL 7

/* Initialise CUDA device `device` (-1: the calling thread's device, else the process default, else the CUDA runtime's current device,
 * e.g. the one torch selected) and make it the calling thread's device.  Idempotent and thread-safe; the first device initialised
 * becomes the process default.  Every entry point initialises its device on first use, so calling this is optional. */
UZKGE_API int32_t uzkge_cuda_init(int32_t device);
/* Route the calling thread's subsequent calls (those that do not name an SRS handle) to `device`; uzkge_cuda_get_device: the device
 * they currently go to (-1 before anything was initialised). */
UZKGE_API int32_t uzkge_cuda_set_device(int32_t device);
UZKGE_API int32_t uzkge_cuda_get_device(void);
UZKGE_API int32_t uzkge_cuda_device_count(void);
/* One process, several GPUs (SURVEY 8b `uzkge_cuda_init(device_count)`, 8e): initialise devices 0 .. device_count - 1 (0 = all visible)
 * as the device GROUP and enable peer access between them.  The reference is a single process (shuffle/src/sdk.rs:196-227), so this
 * is how its Rust host reaches the other GPUs of the box. */
UZKGE_API int32_t uzkge_cuda_init_devices(int32_t device_count);
UZKGE_API int32_t uzkge_cuda_group_size(void);
UZKGE_API const char* uzkge_cuda_last_error(void);
/* "uzkge-b200 <version> sm_100a" */
UZKGE_API const char* uzkge_cuda_version(void);

/* ---- SRS ------------------------------------------------------------------------------------------
 * Upload the G1 part of an SRS (monomial or Lagrange) once.  Replaces the per-commit
 * `G1Projective::normalize_batch(&self.public_parameter_group_1[0..degree + 1])`
 * (uzkge/src/poly_commit/kzg_poly_commitment.rs:287-288): bases live on the GPU in affine form together with
 * the fixed-base window tables 2^(c*f) * P_i that the MSM uses (built on the device, once).
 * window_bits = 0 lets the library choose c from n. */
UZKGE_API int32_t uzkge_cuda_srs_upload(const uint64_t* affine_xy, size_t n, uint32_t window_bits, uint64_t* handle);
UZKGE_API int32_t uzkge_cuda_srs_free(uint64_t handle);
/* The same SRS over the device group; the handle is accepted by uzkge_cuda_msm_g1, uzkge_cuda_msm_g1_batch, uzkge_cuda_srs_info and
 * uzkge_cuda_srs_free (the device-pointer variants need a single-device handle).
 *   UZKGE_MULTI_SPLIT      : device i keeps bases [i n / G, (i + 1) n / G) (tables included) -- "MSM points are split per GPU and the
 *                            partial G1 sums are combined with one projective add each": every MSM runs on all devices at once, one
 *                            worker thread per device, the G partial sums (96 B) are added on the host.
 *   UZKGE_MULTI_REPLICATED : every device keeps the whole SRS -- "a PlonK round's independent polynomial commitments are distributed
 *                            across GPUs" (plonk/prover.rs:132-192, helpers.rs:1323-1408): uzkge_cuda_msm_g1_batch deals MSM j to
 *                            device j mod G; a single uzkge_cuda_msm_g1 runs on the first device. */
#define UZKGE_MULTI_SPLIT 0
#define UZKGE_MULTI_REPLICATED 1
UZKGE_API int32_t uzkge_cuda_srs_upload_multi(const uint64_t* affine_xy, size_t n, uint32_t window_bits, int32_t mode, uint64_t* handle);

/* Setup path: out[i] = tau^i * G (affine, Montgomery), i < n -- the G1 half of KZGCommitmentScheme::new
 * (kzg_poly_commitment.rs:183-204: n sequential scalar multiplications on the CPU).  tau: Montgomery Fr. */
UZKGE_API int32_t uzkge_cuda_srs_generate(const uint64_t tau[4], size_t n, uint64_t* out_affine_xy);

/* Setup path, Lagrange basis: out[i] = L_i(tau) * G for the size-n domain (n = 2^k), i.e. the SRS `lagrange-srs-*.bin` holds for the
 * bundled circuits (uzkge/src/gen_params/mod.rs:42-65) and that prover_with_lagrange commits evaluation vectors against
 * (plonk/prover.rs:119-146).  The scalars L_i(tau) = iNTT(tau^0 .. tau^(n-1))_i are formed on the device.  A deployment without the
 * trapdoor derives the same points from the monomial SRS with a G1 inverse transform (not built). */
UZKGE_API int32_t uzkge_cuda_srs_generate_lagrange(const uint64_t tau[4], size_t n, uint64_t* out_affine_xy);

/* The same SRS WITHOUT the trapdoor: out[i] = (1 / n) sum_j w^(-i j) monomial[j], an inverse transform over G1 points ("iNTT in the
 * exponent") of the first n = 2^k points of a monomial SRS (identity points allowed).  (n / 2) log2 n scalar
 * multiplications: milliseconds at the bundled circuit sizes, seconds at 2^22; run once per SRS. */
UZKGE_API int32_t uzkge_cuda_srs_lagrange_from_monomial(const uint64_t* monomial_affine_xy, size_t n, uint64_t* out_affine_xy);

/* ---- MSM ------------------------------------------------------------------------------------------
 * out = sum_{i < n} scalars[i] * srs[base_offset + i].  Replaces `G1Projective::msm(&points_raw, &coefs)`
 * (kzg_poly_commitment.rs:290) for KZGCommitmentSchemeBN254::commit (:278-293).
 * n == 0 gives the identity.  base_offset + n must not exceed the SRS length (UZKGE_ERR_SIZE; the Rust
 * caller raises DegreeError before calling, kzg_poly_commitment.rs:283-285). */
UZKGE_API int32_t uzkge_cuda_msm_g1(uint64_t handle, size_t base_offset, const uint64_t* scalars, size_t n, uint64_t out_jac[12]);
/* k independent MSMs over the same SRS prefix (the round-1 wire commitments, the split quotient, ...:
 * uzkge/src/plonk/prover.rs:132-192, helpers.rs:1323-1408).  out_jac holds k * 12 words. */
UZKGE_API int32_t uzkge_cuda_msm_g1_batch(uint64_t handle, const uint64_t* const* scalars, const size_t* n, size_t k, uint64_t* out_jac);

/* *d_out_jac (+)= sum_{j < k} scalars[j] * srs[idx[j]], k <= 32, scalars on the host: PolyComScheme::apply_blind_factors
 * (kzg_poly_commitment.rs:299-313) folded onto a device-resident commitment (accumulate = 1). */
UZKGE_API int32_t uzkge_cuda_msm_g1_small_device(uint64_t handle, const size_t* idx, const uint64_t* scalars_host, size_t k, int32_t accumulate,
                                                 void* d_out_jac, void* stream);

/* ---- NTT ------------------------------------------------------------------------------------------
 * In-place transform of `inout` (capacity domain_size elements; the first len_in are the input, the rest
 * is treated as zero).  domain_size = 2^k (Radix2EvaluationDomain) or 3 * 2^k (MixedRadixEvaluationDomain),
 * natural order in and out, generator = arkworks' get_root_of_unity(domain_size).
 *   inverse = 0, coset_shift = NULL : FpPolynomial::fft_with_domain        (field_polynomial.rs:583-586)
 *   inverse = 0, coset_shift = k    : FpPolynomial::coset_fft_with_domain  (:589-591)  c_j * k^j, then fft
 *   inverse = 1, coset_shift = NULL : domain.ifft of ifft_with_domain      (:594-597)  incl. 1/N; the
 *                                     trailing-zero trim stays in Rust (`from_coefs`)
 *   inverse = 1, coset_shift = k^-1 : coset_ifft_with_domain               (:601-607)  ifft, then c_j * (k^-1)^j */
UZKGE_API int32_t uzkge_cuda_ntt_fr(uint64_t* inout, size_t len_in, size_t domain_size, int32_t inverse, const uint64_t* coset_shift);
/* k transforms over the same domain in one call (the 5 wire iFFTs of a prover round, the 7 coset FFTs of t_poly, the ~33 of the
 * indexer: plonk/prover.rs:155-170, helpers.rs:256-266, indexer.rs:296-470), software-pipelined over three device buffers and three
 * streams: the H2D copy of vector j + 1 and the D2H copy of vector j - 1 run under transform j.  inouts[j] holds domain_size elements,
 * the first len_in[j] are the input. */
UZKGE_API int32_t uzkge_cuda_ntt_fr_batch(uint64_t* const* inouts, const size_t* len_in, size_t k, size_t domain_size, int32_t inverse,
                                          const uint64_t* coset_shift);
/* ONE transform over the whole device group (uzkge_cuda_init_devices; 2, 4 or 8 GPUs), same contract as uzkge_cuda_ntt_fr: the
 * four-step decomposition with both exchanges done by the kernels' own loads and stores over peer memory (NVLink) and the natural
 * output order folded into the last pass.  Every GPU moves only its 1/G of the vector over its own host link, which is what bounds a
 * host-pointer transform.  2^k domains of at least G^2 points; anything else is forwarded to uzkge_cuda_ntt_fr on the calling
 * thread's device. */
UZKGE_API int32_t uzkge_cuda_ntt_fr_multi(uint64_t* inout, size_t len_in, size_t domain_size, int32_t inverse, const uint64_t* coset_shift);
/* generator of the size-n domain (Montgomery), for the Rust side's consistency assert against
 * `domain.group_gen`; UZKGE_ERR_SIZE if n is not 3^a 2^b with a <= 2, b <= 28. */
UZKGE_API int32_t uzkge_cuda_fr_root_of_unity(size_t n, uint64_t out[4]);

/* ---- device-resident variants ----------------------------------------------------------------------
 * Same semantics with DEVICE pointers and a caller-provided cudaStream_t (NULL = default stream); nothing
 * is copied and, in the steady state, nothing synchronises: this is what a device-resident prover pipeline and the
 * HBM-resident benchmark call.  d_out may equal d_in for the NTT; d_scratch holds domain_size elements.
 * Exceptions (first-use set-up, hence not CUDA-graph capturable until warmed up): the FIRST transform of a domain size, and the first
 * use of a (size, coset shift) pair, build their twiddle / coset tables (cudaMalloc + one stream synchronisation); at most 16 coset
 * tables (and 8 GiB of them) are kept per device, rotating more (size, shift) pairs rebuilds them (cudaFree + cudaMalloc);
 * uzkge_cuda_grand_product_fr_device and uzkge_cuda_fr_trimmed_len_device read one value back and say so below. */
UZKGE_API int32_t uzkge_cuda_msm_g1_device(uint64_t handle, size_t base_offset, const void* d_scalars, size_t n, void* d_out_jac, void* stream);
/* k MSMs over srs[base_offset ..] in one pass (one sort, one accumulate launch, concurrent reductions); d_scalars is a HOST
 * array of k device pointers, d_out_jac holds k * 12 words on the device. */
UZKGE_API int32_t uzkge_cuda_msm_g1_batch_device(uint64_t handle, size_t base_offset, const void* const* d_scalars, const size_t* n,
                                                 size_t k, void* d_out_jac, void* stream);
UZKGE_API int32_t uzkge_cuda_ntt_fr_device(const void* d_in, void* d_out, void* d_scratch, size_t len_in, size_t domain_size,
                                 int32_t inverse, const uint64_t* coset_shift_host, void* stream);

/* k <= 16 independent transforms over ONE domain (and coset shift) in one launch per pass -- the 5 (+3) wire interpolations and the
 * 5 (+3) coset evaluations of a prover round (plonk/prover.rs:155-170, helpers.rs:256-266): a 2^14-point vector alone fills a tenth
 * of the SMs and its passes are launch-bound.  d_ins / d_outs / len_in: HOST arrays of k device pointers / lengths; d_scratch holds
 * k * domain_size elements.  d_outs[j] may equal d_ins[j]; distinct vectors must not overlap. */
UZKGE_API int32_t uzkge_cuda_ntt_fr_batch_device(const void* const* d_ins, void* const* d_outs, void* d_scratch, const size_t* len_in, size_t k,
                                                 size_t domain_size, int32_t inverse, const uint64_t* coset_shift_host, void* stream);

/* Cross-GPU step of a distributed transform of size n_total = G * L over G = 2^log_ranks ranks (four-step NTT, SURVEY 8e:
 * "NTTs of size 2^22 and above use a four-step transpose over NVLink with NCCL all-to-all").  d_in holds G rows of
 * `cols` elements: row n1 = the columns [col_offset, col_offset + cols) of rank n1's contiguous slice (what the first
 * all-to-all delivers).  d_out row k1 = w^(n2 * k1) * sum_n1 in[n1][.] * w_G^(n1 * k1): it travels to rank k1 (second
 * all-to-all), which then runs the size-L transform with uzkge_cuda_ntt_fr_device.  uzkge_b200/dist.py drives it. */
UZKGE_API int32_t uzkge_cuda_ntt_cross_fr_device(const void* d_in, void* d_out, uint32_t log_ranks, size_t cols, size_t col_offset,
                                                 size_t n_total, int32_t inverse, void* stream);

/* The same step with one base pointer per row (2^log_ranks HOST arrays of device pointers).  With rows mapped from PEER GPUs
 * (uzkge_cuda_ipc_*) the kernel itself performs both exchanges of the four-step transform over NVLink: d_in_rows[n1] = rank n1's
 * slice + col_offset (P2P loads instead of the first all-to-all), d_out_rows[k1] = rank k1's receive buffer + col_offset (P2P
 * stores instead of the second).  The caller orders the ranks around it (inputs ready before, stores complete after). */
UZKGE_API int32_t uzkge_cuda_ntt_cross_rows_fr_device(const void* const* d_in_rows, void* const* d_out_rows, uint32_t log_ranks, size_t cols,
                                                      size_t col_offset, size_t n_total, int32_t inverse, void* stream);
/* Device buffers that other processes of the box can map (one process per GPU): cudaMalloc + cudaIpcGetMemHandle /
 * cudaIpcOpenMemHandle.  `handle` is 64 opaque bytes to ship through any host channel (torch.distributed all_gather_object). */
UZKGE_API int32_t uzkge_cuda_dev_alloc(size_t bytes, void** d_ptr);
UZKGE_API int32_t uzkge_cuda_dev_free(void* d_ptr);
/* The local size-n transform of a distributed four-step transform with the NATURAL output order folded into its final store: rank
 * `rank` of 2^log_ranks holds, after the cross step, row k1 = rank; output k2 of its size-n transform is X[k1 + G k2] of the whole
 * transform and is written straight into its owner's contiguous slice, d_out_rows[k2 / (n / G)] + k1 + G (k2 mod n / G), over peer
 * memory -- no third exchange.  d_in: n elements (left intact: the passes work in d_scratch, n elements); d_out_rows: HOST array of G
 * device pointers (peer mappings). */
UZKGE_API int32_t uzkge_cuda_ntt_fr_scatter_device(const void* d_in, void* const* d_out_rows, void* d_scratch, size_t n, int32_t inverse,
                                                   uint32_t log_ranks, uint32_t rank, void* stream);
/* Synchronous host <-> device copies for callers without their own CUDA binding (the Rust crate, the C++ host layer): ordered with
 * the *_device entry points called with stream = NULL.  What `Vec<Fr>` <-> device buffer conversions of a device-resident
 * prover_with_lagrange (plonk/prover.rs:88-394) need at its ends. */
UZKGE_API int32_t uzkge_cuda_dev_copy_in(void* d_dst, const void* h_src, size_t bytes);
UZKGE_API int32_t uzkge_cuda_dev_copy_out(void* h_dst, const void* d_src, size_t bytes);
UZKGE_API int32_t uzkge_cuda_ipc_export(const void* d_ptr, uint8_t handle[64]);
UZKGE_API int32_t uzkge_cuda_ipc_open(const uint8_t handle[64], void** d_ptr);
UZKGE_API int32_t uzkge_cuda_ipc_close(void* d_ptr);

/* ---- polynomial glue over Fr (the prover's O(n) serial loops as scans) ----------------------------------
 * FpPolynomial::eval (field_polynomial.rs:198-209): out = sum_j coefs[j] * x^j.  n >= 1. */
UZKGE_API int32_t uzkge_cuda_poly_eval_fr(const uint64_t* coefs, size_t n, const uint64_t x[4], uint64_t out[4]);
/* FpPolynomial::div_rem by the divisor (X - z) (field_polynomial.rs:519-550 as used by KZG `prove`,
 * kzg_poly_commitment.rs:322-335): quotient holds n - 1 coefficients (untrimmed), rem = p(z). */
UZKGE_API int32_t uzkge_cuda_poly_div_linear_fr(const uint64_t* coefs, size_t n, const uint64_t z[4], uint64_t* quotient, uint64_t rem[4]);
/* Device-resident Horner scan: *d_value = p(z); d_quotient (n - 1 elements, may be NULL) = p / (X - z).  No copies, no sync. */
UZKGE_API int32_t uzkge_cuda_poly_horner_fr_device(const void* d_coefs, size_t n, const uint64_t z_host[4], void* d_quotient, void* d_value,
                                                   void* stream);
/* k <= UZKGE_EVAL_BATCH_MAX evaluations in two launches: d_values[j] = polys[j](points[point_index[j]]), at most two distinct
 * points (the prover opens 11 polynomials at zeta and 4 at zeta * omega, plonk/prover.rs:217-244).  No copies, no sync. */
#define UZKGE_EVAL_BATCH_MAX 32
UZKGE_API int32_t uzkge_cuda_poly_eval_batch_fr_device(const void* const* d_polys, const size_t* lens, const uint32_t* point_index, size_t k,
                                                       const uint64_t* points_host, size_t npoints, void* d_values, void* stream);
/* The grand product of z_poly (plonk/helpers.rs:204-217: batch_inversion of the denominators, then the running product):
 * out[0] = 1, out[i + 1] = out[i] * num[i] / den[i], i < n  (n + 1 outputs).  UZKGE_ERR_ARG if a denominator is zero. */
UZKGE_API int32_t uzkge_cuda_grand_product_fr(const uint64_t* num, const uint64_t* den, size_t n, uint64_t* out);

/* ---- the quotient polynomial's pointwise map -------------------------------------------------------------
 * t_poly's loop over the m = factor * n points of the coset k[1] * <w_m> (plonk/helpers.rs:284-669, the terms 1-11 that exist
 * without the `shuffle` feature): every pointer is a DEVICE array of m Fr elements in natural order -- the coset evaluations of
 * the wire, public-input and z polynomials (helpers.rs:256-266) and of the preprocessed polynomials in PlonkProverParams
 * (plonk/indexer.rs:77-139).  The omega-shifted reads use index (point + factor) % m (helpers.rs:308).  z_h_inv holds the
 * `factor` inverses of the vanishing polynomial on the coset (helpers.rs:244-253); scalars are Montgomery Fr on the host. */
typedef struct {
    const void* w[5];             /* w_polys_coset_evals */
    const void* q[9];             /* q_coset_evals (N_SELECTORS = 9, constraint_system/turbo/mod.rs:23) */
    const void* pi;               /* pi_coset_evals */
    const void* z;                /* z_coset_evals */
    const void* s[5];             /* s_coset_evals */
    const void* coset_quotient;   /* k[1] * w_m^point (plonk/indexer.rs:278-282) */
    const void* l1;               /* l1_coset_evals */
    const void* qb;               /* qb_coset_eval */
    const void* q_prk[4];         /* q_prk_coset_evals */
    uint64_t k[5][4];
    uint64_t alpha[4], beta[4], gamma[4];
    uint64_t anemoi_generator[4], anemoi_generator_inv[4];
    uint64_t z_h_inv[16][4];
    size_t m, factor;
} uzkge_quotient_args;
UZKGE_API int32_t uzkge_cuda_plonk_quotient_fr_device(const uzkge_quotient_args* args, void* d_out, void* stream);
/* The same map for the `shuffle` feature set (what zshuffle is built with): terms 12-18 of t_poly (plonk/helpers.rs:416-640) on top of
 * terms 1-11 -- the remark gates' elliptic-curve additions selected by the witness selectors, and the selector constraints.  Device
 * arrays of m elements: the coset evaluations of the 3 witness-selector polynomials (prover.rs:148-165), of q_ecc and of the 12 + 12
 * shuffle public-key / generator selector polynomials (indexer.rs:447-501; order x_00..x_11, y_00..y_11, dxy_00..dxy_11). */
typedef struct {
    const void* w_sel[3];
    const void* q_ecc;
    const void* pk[12];
    const void* gen[12];
    uint64_t edwards_a[4];
} uzkge_quotient_shuffle_args;
UZKGE_API int32_t uzkge_cuda_plonk_quotient_shuffle_fr_device(const uzkge_quotient_args* args, const uzkge_quotient_shuffle_args* shuffle,
                                                              void* d_out, void* stream);

/* The same map (either feature set: shuffle may be NULL) on the points p = start + step * i, i < count, of the size-m arrays; d_out is
 * written at the same positions.  (start, step, count) = (j, factor, n) is the coset g_j <w_n>, g_j = k[1] w_m^j, of the quotient
 * domain: the omega-shifted point stays inside the coset, so the cosets of one quotient are independent jobs -- the unit by which a
 * device group splits the round (uzkge_cuda_plonk_prove over a multi-device parameter handle). */
UZKGE_API int32_t uzkge_cuda_plonk_quotient_range_fr_device(const uzkge_quotient_args* args, const uzkge_quotient_shuffle_args* shuffle,
                                                            uint64_t start, uint64_t step, uint64_t count, void* d_out, void* stream);

/* The inverse of that split: d_u holds, for every coset j < factor, u_j = the size-n coset iFFT (shift g_j^-1) of the quotient's values
 * on coset j (compact, u_j at [j n, (j + 1) n)); d_out receives the factor * n coefficients of t -- what coset_ifft_with_domain over the
 * whole 6n domain (helpers.rs:673-677) returns.  Per coefficient index r a factor-point inverse DFT over the cosets (11 products for
 * factor = 6).  k1: the quotient coset's shift k[1] (host, Montgomery).  d_out must not alias d_u. */
UZKGE_API int32_t uzkge_cuda_plonk_coset_combine_fr_device(const void* d_u, size_t n, size_t factor, const uint64_t k1_host[4], void* d_out,
                                                           void* stream);

/* ---- elementwise glue of a device-resident prover; DEVICE pointers, caller's stream, no copies, no sync ------
 * out[i] = sum_{j < k} coefs[j] * polys[j][i] for i < out_len, where polys[j][i] = 0 for i >= lens[j]; coefs: k Montgomery Fr on
 * the HOST.  Replaces the mul / add_assign chains of r_poly_or_comm (plonk/helpers.rs:716-745, 986-993) and of batch_prove
 * (poly_commit/pcs.rs:124-131).  d_out may alias one of the inputs.  1 <= k <= UZKGE_LINCOMB_MAX. */
#define UZKGE_LINCOMB_MAX 24
UZKGE_API int32_t uzkge_cuda_fr_lincomb_device(const void* const* d_polys, const size_t* lens, const uint64_t* coefs_host, size_t k,
                                               void* d_out, size_t out_len, void* stream);
/* poly[idx[j]] += vals[j], j < k <= UZKGE_SPARSE_MAX, applied in order (indices may repeat): FpPolynomial::add_coef_assign as used by
 * hide_polynomial (plonk/helpers.rs:139-154) and split_t_and_commit (helpers.rs:1351-1361). */
#define UZKGE_SPARSE_MAX 16
UZKGE_API int32_t uzkge_cuda_fr_add_sparse_device(void* d_poly, const size_t* idx, const uint64_t* vals_host, size_t k, void* stream);
/* polys[j][idx[j]] += vals[j], j < k <= UZKGE_SPARSE_MULTI_MAX, over SEVERAL polynomials in one launch; the k entries must address distinct
 * elements (they are applied in parallel): the blinds of all the polynomials of a prover round (helpers.rs:139-154). */
#define UZKGE_SPARSE_MULTI_MAX 64
UZKGE_API int32_t uzkge_cuda_fr_add_sparse_multi_device(void* const* d_polys, const size_t* idx, const uint64_t* vals_host, size_t k, void* stream);
/* out[i] = scale * base^i, i < n (scale NULL = 1): `domain.elements()` and the coset k[1] * w_m^i (plonk/indexer.rs:276-282). */
UZKGE_API int32_t uzkge_cuda_fr_powers_device(const uint64_t base_host[4], const uint64_t* scale_host, size_t n, void* d_out, void* stream);
/* out[i] = src[idx[i]], idx: n uint32 on the device: ConstraintSystem::extend_witness (plonk/constraint_system/mod.rs:103-111). */
UZKGE_API int32_t uzkge_cuda_fr_gather_device(const void* d_src, const void* d_idx_u32, size_t n, void* d_out, void* stream);
/* dst[dst_start + dst_step * i] = src[src_start + src_step * i], i < count: one coset of an interleaved domain <-> a compact vector. */
UZKGE_API int32_t uzkge_cuda_fr_strided_copy_device(const void* d_src, size_t src_start, size_t src_step, void* d_dst, size_t dst_start,
                                                    size_t dst_step, size_t count, void* stream);
/* dst[dst_idx[j]] = src[src_idx[j]], j < k; both index arrays are uint32 on the device: pi_poly's evaluation vector (the public inputs on
 * their constraint rows, plonk/helpers.rs:111-131) filled straight from the device-resident witness. */
UZKGE_API int32_t uzkge_cuda_fr_gather_scatter_device(const void* d_src, const void* d_src_idx_u32, void* d_dst, const void* d_dst_idx_u32, size_t k,
                                                      void* stream);
/* out[i] = a[i] * b[i], i < n (witness generation of synthetic circuits; FpPolynomial's pointwise products). */
UZKGE_API int32_t uzkge_cuda_fr_mul_device(const void* d_a, const void* d_b, size_t n, void* d_out, void* stream);
/* *len_out = number of coefficients after FpPolynomial::from_coefs' trailing-zero trim (field_polynomial.rs:86-90): 1 + the largest
 * index holding a non-zero element, 0 for the zero vector.  Synchronises the stream (the prover's split_t_and_commit branches on it,
 * plonk/helpers.rs:1334-1347). */
UZKGE_API int32_t uzkge_cuda_fr_trimmed_len_device(const void* d_poly, size_t n, size_t* len_out, void* stream);
/* Device-resident grand product (see uzkge_cuda_grand_product_fr): d_out holds n + 1 elements, d_tmp 2 n + 2.  Synchronises the
 * stream once (the single field inversion runs on the host). */
UZKGE_API int32_t uzkge_cuda_grand_product_fr_device(const void* d_num, const void* d_den, size_t n, void* d_out, void* d_tmp, void* stream);
/* z_poly's evaluations (plonk/helpers.rs:160-220) from device-resident inputs: d_w[j] / d_sigma[j] = the n values of wire j and of
 * its encoded permutation k_j' * w^i' (indexer.rs:195-208, :305-313), d_group = w^i; k: 5 Montgomery Fr on the host.
 * d_z[0] = 1, d_z[i + 1] = d_z[i] * prod_j (w_j + beta k_j w^i + gamma) / prod_j (w_j + beta sigma_j + gamma);  d_tmp: 4 n elements. */
UZKGE_API int32_t uzkge_cuda_plonk_z_evals_fr_device(const void* const d_w[5], const void* const d_sigma[5], const void* d_group,
                                                     const uint64_t* k_host, const uint64_t beta_host[4], const uint64_t gamma_host[4], size_t n,
                                                     void* d_z, void* d_tmp, void* stream);

/* ---- the whole prover behind the boundary ---------------------------------------------------------------------
 * prover_with_lagrange (plonk/prover.rs:88-394) as ONE call: the caller (the patched Rust body) keeps what is serial and cheap -- building
 * the constraint system, transcript_init_plonk (plonk/transcript.rs:8-31), the RNG draws -- and hands over everything that touches a
 * polynomial.  All five rounds run device-resident; the library hashes the Fiat-Shamir transcript itself (Keccak-256, same bytes as
 * utils/transcript.rs) between the rounds.
 *
 * 1. uzkge_cuda_plonk_params_upload: once per circuit, from PlonkProverParams (plonk/indexer.rs:77-139).  Only the COEFFICIENT forms and the
 *    permutation travel (host pointers, Montgomery Fr; `len` coefficients each, a NULL pointer or len 0 is the zero polynomial); the
 *    evaluations on the quotient coset (q_coset_evals, s_coset_evals, ... 6 n values per polynomial in the reference's struct) are
 *    recomputed on the device, as are `group`, `coset_quotient`, l1 and Z_H^-1 on the coset.
 * 2. uzkge_cuda_srs_upload_lagrange_commit: the bases of the `commit` closure of prover_with_lagrange (prover.rs:131-146) as one SRS:
 *    [L_0(tau) G .. L_(n-1)(tau) G | SRS[0..3) | SRS[n..n+3)], so that lagrange_pcs.commit(evals) + apply_blind_factors(blinds, n)
 *    (kzg_poly_commitment.rs:299-313) is a single MSM over [evals | b | -b].
 * 3. uzkge_cuda_plonk_prove. */
typedef struct {
    uint64_t n;                          /* cs.size(), a power of two */
    uint64_t m;                          /* cs.quot_eval_dom_size(): 6 n (n > 8) or 16 n */
    uint64_t num_vars;                   /* witness length */
    const uint32_t* wiring;              /* 5 n variable indices, wire-major: ConstraintSystem::extend_witness (constraint_system/mod.rs:103-111) */
    const uint64_t* permutation;         /* 5 n positions (usize as u64): PlonkProverParams::permutation */
    uint64_t k[5][4];                    /* verifier_params.k, Montgomery */
    const uint64_t* q_polys[9];          /* N_SELECTORS = 9 */
    size_t q_len[9];
    const uint64_t* s_polys[5];
    size_t s_len[5];
    const uint64_t* qb_poly;
    size_t qb_len;
    const uint64_t* q_prk_polys[4];
    size_t q_prk_len[4];
    uint64_t anemoi_generator[4], anemoi_generator_inv[4];
    const uint64_t* public_vars_constraint_indices;   /* n_public rows */
    const uint64_t* public_vars_witness_indices;      /* n_public variable indices */
    size_t n_public;
    int32_t shuffle;                     /* 1: the `shuffle` feature set (zshuffle, zmatchmaking): the fields below are read */
    int32_t reserved;
    const uint64_t* q_ecc_poly;
    size_t q_ecc_len;
    const uint64_t* q_shuffle_generator_polys[12];
    size_t gen_len[12];
    const uint64_t* q_shuffle_public_key_polys[12];
    size_t pk_len[12];
    uint64_t edwards_a[4];
} uzkge_plonk_params_desc;
UZKGE_API int32_t uzkge_cuda_plonk_params_upload(const uzkge_plonk_params_desc* desc, uint64_t* params_handle);
/* The same parameters on every device of the group (uzkge_cuda_init_devices), uploaded side by side.  uzkge_cuda_plonk_prove over such
 * a handle -- with `srs` / `lagrange_srs` multi-device handles in UZKGE_MULTI_SPLIT mode -- produces ONE proof on all the GPUs: every
 * device runs the prover on replicated polynomials (same witness, blinds and transcript, so the same challenges), but commits only its
 * slice of the SRS points (the 96-byte partial sums are added on the host: "MSM points are split per GPU") and evaluates the quotient
 * only on its cosets of the 6n domain, which the devices exchange over peer memory before the inverse transform.  The proof is
 * byte-identical to the single-device proof.  set_public_key / free accept the handle. */
UZKGE_API int32_t uzkge_cuda_plonk_params_upload_multi(const uzkge_plonk_params_desc* desc, uint64_t* params_handle);
/* refresh_prover_params_public_key (shuffle/src/gen_params/params.rs:57-129): replace the 12 public-key selector polynomials. */
UZKGE_API int32_t uzkge_cuda_plonk_params_set_public_key(uint64_t params_handle, const uint64_t* const polys[12], const size_t len[12]);
UZKGE_API int32_t uzkge_cuda_plonk_params_free(uint64_t params_handle);
/* lagrange_xy: the n points of the size-n Lagrange SRS; monomial_xy: at least n + 3 points of the monomial SRS (only [0, 3) and
 * [n, n + 3) are read -- exactly what the bundled srs-padding.bin keeps, gen_params/mod.rs:147-171). */
UZKGE_API int32_t uzkge_cuda_srs_upload_lagrange_commit(const uint64_t* lagrange_xy, size_t n, const uint64_t* monomial_xy, size_t monomial_len,
                                                        uint32_t window_bits, uint64_t* handle);
/* The same SRS split over the device group (UZKGE_MULTI_SPLIT), for uzkge_cuda_plonk_prove over a multi-device parameter handle. */
UZKGE_API int32_t uzkge_cuda_srs_upload_lagrange_commit_multi(const uint64_t* lagrange_xy, size_t n, const uint64_t* monomial_xy,
                                                              size_t monomial_len, uint32_t window_bits, uint64_t* handle);

typedef struct {
    uint64_t params;                     /* uzkge_cuda_plonk_params_upload */
    uint64_t srs;                        /* monomial SRS (>= n + 3 points); may be 0 when lagrange_all is set */
    uint64_t lagrange_srs;               /* uzkge_cuda_srs_upload_lagrange_commit, or 0: every commitment over the monomial SRS */
    int32_t lagrange_all;                /* 1: also the witness selectors, quotient pieces and opening proofs over the Lagrange bases
                                            (helpers.rs:1363-1391, pcs.rs:139-163) -- required when the monomial SRS has holes */
    int32_t witness_on_device;           /* 0: `witness` is a host pointer; 1: a device pointer */
    const uint64_t* witness;             /* num_vars Montgomery Fr */
    const uint64_t* w_sel_evals[3];      /* shuffle feature set: compute_witness_selectors (turbo/mod.rs:171-191), n Fr each (host);
                                            all NULL when the circuit has no remark gate */
    const uint64_t* blinds;              /* the prover's Fr::rand draws in the reference's order (Montgomery): wires 3,3,3,2,2;
                                            [3 x 2 witness selectors]; z 3; 5 quotient-split blinds -- 21 values, 27 with shuffle */
    size_t n_blinds;
    const uint8_t* transcript;           /* transcript state after transcript_init_plonk */
    size_t transcript_len;
} uzkge_plonk_prove_args;
/* PlonkProof (plonk/indexer.rs:33-75).  Commitments: affine x, y in Montgomery form (x = y = 0: the identity); evaluations: Montgomery Fr. */
typedef struct {
    uint64_t cm_w[5][8], cm_w_sel[3][8], cm_t[5][8], cm_z[8];
    uint64_t prk_3_poly_eval_zeta[4], prk_4_poly_eval_zeta[4];
    uint64_t w_polys_eval_zeta[5][4], w_polys_eval_zeta_omega[3][4], z_eval_zeta_omega[4], s_polys_eval_zeta[4][4];
    uint64_t q_ecc_poly_eval_zeta[4], w_sel_polys_eval_zeta[3][4];
    uint64_t opening_witness_zeta[8], opening_witness_zeta_omega[8];
    uint8_t transcript_state[32];        /* the transcript after the proof: the state is always one 32-byte slot here */
    uint32_t launches;                   /* kernels launched for this proof */
    uint32_t msm, ifft_n, fft_n, coset_fft_m, coset_ifft_m, evals;    /* what ran: the proof's operation inventory */
    double rounds_ms[6];                 /* wall clock per stage: wires, z, quotient, commit t, evaluations + r, openings */
} uzkge_plonk_proof;
/* Errors: UZKGE_ERR_SIZE = a polynomial does not fit the SRS (the reference's DegreeError: an unsatisfied witness makes the quotient
 * too long) or too few blinds; UZKGE_ERR_ARG = the opening remainder is not zero (PCSProveEvalError) / bad arguments. */
UZKGE_API int32_t uzkge_cuda_plonk_prove(const uzkge_plonk_prove_args* args, uzkge_plonk_proof* proof);

/* ---- small group helpers (combine per-GPU partial MSMs; blinds) --------------------------------------
 * out = a + b on Jacobian points (host pointers, tiny device kernel).  Used for the G - 1 projective adds
 * that merge per-GPU partial sums. */
UZKGE_API int32_t uzkge_cuda_g1_add(const uint64_t a_jac[12], const uint64_t b_jac[12], uint64_t out_jac[12]);
/* d_out[j] = sum_{r < count} d_parts[r * stride + j], j < k (Jacobian, 12 words each; DEVICE pointers, caller's stream, no sync; d_out
 * may alias the first k parts): the combine step after the k partial sums of every GPU of a point-split MSM batch were gathered
 * rank-major (NCCL all-gather of k * 96 bytes per GPU). */
UZKGE_API int32_t uzkge_cuda_g1_sum_device(const void* d_parts_jac, size_t count, size_t stride, size_t k, void* d_out_jac, void* stream);
/* Jacobian -> affine (x = y = 0 for the identity): the normalisation `into_affine` performs before a
 * commitment is serialised (kzg_poly_commitment.rs:37-53). */
UZKGE_API int32_t uzkge_cuda_g1_to_affine(const uint64_t in_jac[12], uint64_t out_affine[8]);

/* ---- pinned host memory ---------------------------------------------------------------------------------
 * Optional.  Every host-pointer entry point accepts pageable memory (a Rust Vec); buffers obtained here, or
 * registered in place, are page-locked so the PCIe copies run at full rate.  The Rust side can register the
 * long-lived vectors (SRS, prover-parameter evaluations) once. */
UZKGE_API int32_t uzkge_cuda_host_alloc(size_t bytes, void** out);
UZKGE_API int32_t uzkge_cuda_host_free(void* p);
UZKGE_API int32_t uzkge_cuda_host_register(void* p, size_t bytes);
UZKGE_API int32_t uzkge_cuda_host_unregister(void* p);

/* ---- introspection for benchmarks / profiling --------------------------------------------------------*/
typedef struct {
    uint32_t window_bits;   /* c */
    uint32_t windows;       /* ceil(255 / c) = number of fixed-base tables */
    uint64_t n;             /* SRS length */
    uint64_t device_bytes;  /* device bytes held for this SRS (tables + MSM workspace) */
    uint32_t batch_slots;   /* independent MSMs one pass can carry (uzkge_cuda_msm_g1_batch) */
    uint32_t reserved;
    double precompute_ms;   /* device time spent building the tables */
} uzkge_srs_info;
UZKGE_API int32_t uzkge_cuda_srs_info(uint64_t handle, uzkge_srs_info* info);
/* number of kernel launches issued by this library since init (bench.py's gpu_launches) */
UZKGE_API uint64_t uzkge_cuda_launch_count(void);
/* Per-phase device timing with CUDA events recorded on the launching stream (off by default).
 *   kind 0 = MSM: phase_ms[0..5] = recode, sort, offsets, accumulate, oversized buckets, reduce
 *   kind 1 = NTT: phase_ms[0..3] = radix-3 pre-pass, pass 0, pass 1, pass 2
 * profile_read synchronises the device, returns the sums over the `runs` engine runs since the last read and
 * resets them. */
UZKGE_API int32_t uzkge_cuda_profile_enable(int32_t on);
UZKGE_API int32_t uzkge_cuda_profile_read(int32_t kind, double phase_ms[8], uint64_t* runs);
/* tuning knobs for experiments: "msm_lanes" (0 = auto, else 1..32 lanes per bucket), "ntt_log_tile",
 * "ntt_max_log_r", "ntt_two_pass_max" (affect plans created afterwards); "ntt_radix4" (transforms of at least 2^value points walk two
 * butterfly stages per shared-memory round trip; 0 = off, default 19); "quotient_min_blocks"; "l2_fetch_granularity" (32 / 64 / 128:
 * cudaLimitMaxL2FetchGranularity of the calling thread's device); "virtual_devices" (tests: the next uzkge_cuda_init_devices forms a
 * group of that many members over the visible GPUs, 0 = one member per GPU); "group_deal_min_log_n" (a device group deals a round's
 * interpolations and the linearisation polynomial to its members for circuits of at least 2^value gates; default 19). */
UZKGE_API int32_t uzkge_cuda_configure(const char* key, uint64_t value);
/* K1 field kernels, exposed for parity tests and for the integer-pipe roof measurement:
 *   out[i] = a[i] * b[i] (Montgomery), field = 0: Fr, 1: Fq; host pointers. */
UZKGE_API int32_t uzkge_cuda_field_mul(int32_t field, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* `chains` dependent multiplication chains of length `iters` per thread over the whole GPU; reports achieved
 * field multiplications per second (device time). */
UZKGE_API int32_t uzkge_cuda_bench_field_mul(int32_t field, uint32_t iters, double* muls_per_s);

#ifdef __cplusplus
}
#endif
#endif /* UZKGE_CUDA_H */
