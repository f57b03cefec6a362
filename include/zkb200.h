/*
 * zkb200.h — C ABI of libzkb200.so: B200-native BN254 G1 MSM and Fr NTT behind halo2's call signatures.
 *
 * This is the drop-in boundary for the zksnap provers.  The reference
 * (aerius-labs/zksnap-circuits-halo2) has no FFI layer of its own; it reaches the hot path only through
 * halo2-axiom's keygen_vk / keygen_pk / create_proof
 *   /root/reference/aggregator/src/wrapper.rs:106-109   (gen_pk -> keygen_vk, keygen_pk)
 *   /root/reference/aggregator/src/wrapper.rs:129-137   (create_proof::<_, ProverGWC<_>, ...>)
 *   /root/reference/voter/benches/voter_circuit.rs:60-62,80
 *   /root/reference/aggregator/benches/state_transition_circuit.rs:64-66,84
 *   /root/reference/aggregator/benches/wrapper_circuit.rs:107,140
 * so each entry point below names the halo2-axiom / halo2curves function whose body it replaces (sources
 * are un-vendored third-party crates; see INTEGRATION.md for the Rust shim that binds these symbols).
 *
 * Encoding (exactly what the Rust side holds in memory):
 *   Fr / Fq     4 little-endian u64 limbs, Montgomery form (a * 2^256 mod m), canonical (< m)
 *   G1Affine    8 u64: x, y; the identity is (0, 0)
 *   G1          12 u64: x, y, z Jacobian (affine = (x/z^2, y/z^3)); identity has z = 0 and is returned as
 *               (0, R, 0); every other result is returned normalised to z = R (Montgomery one), so the bytes
 *               are deterministic and equal to `to_affine()` of the CPU result.
 * Pointers need only 8-byte alignment.  The caller owns every buffer; the library keeps no reference past
 * return except the device copy of SRS bases registered with zkb_srs_register.
 *
 * All functions return 0 on success or a negative ZKB_ERR_* code; zkb_last_error() gives the thread-local
 * message.  There is no CPU fallback: without a usable CUDA device every compute call fails with
 * ZKB_ERR_NO_DEVICE.  Entry points are thread-safe (serialised per device) and synchronous: host buffers are
 * valid when the call returns.  The *_dev variants take device pointers and a CUDA stream and are
 * asynchronous with respect to the host unless stated otherwise.
 */
#ifndef ZKB200_H
#define ZKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define ZKB_OK 0
#define ZKB_ERR_ARG (-1)        /* bad argument (NULL, size mismatch, log_n out of range) */
#define ZKB_ERR_CUDA (-2)       /* CUDA runtime error */
#define ZKB_ERR_OOM (-3)        /* device or pinned-host allocation failed */
#define ZKB_ERR_NO_DEVICE (-4)  /* no CUDA device / library not initialised */
#define ZKB_ERR_HANDLE (-5)     /* unknown SRS handle */

/* ---- lifecycle ------------------------------------------------------------------------------------------- */

/* Bind this process to 1..8 GPUs of one box.  devices == NULL or ndev == 0 selects the devices named by the environment
 * variable ZKB_DEVICES ("all" or a list such as "0,1,2,3") or, when it is unset, the current CUDA device — so an
 * unmodified caller drives the whole box by setting one variable.  Idempotent for the same list.
 * With several devices the ONE process the reference's prover is (create_proof, /root/reference/aggregator/src/wrapper.rs:
 * 129-137, chained by gen_recursion_snark, wrapper.rs:869-902) reaches all of them through the same entry points:
 *   - zkb_srs_register & co. replicate the SRS to every device (peer copies over NVLink);
 *   - zkb_msm_g1 / zkb_msm_g1_srs / _range shard one commit by SRS point range: each device uploads only its share of the
 *     scalars over its own PCIe link, the <= 8 partial sums (96 B each) are folded on the host;
 *   - the *_batch forms (commits and NTT-family ops) split the batch by column, no communication;
 *   - one large transform (zkb_ntt_fr, zkb_lagrange_to_coeff, zkb_coeff_to_lagrange, zkb_extended_to_coeff of >= 2^22
 *     elements) is sharded over the devices with the exchange fused into the NTT passes over NVLink peer memory.
 * Results are bit-identical to the single-device ones.  Polynomial handles, quotient evaluation and the *_dev variants act
 * on one device: the first of the list (home device), or for *_dev the device that owns the pointer.
 * The one-process-per-GPU deployment (torchrun, one rank per device, zkb_dist_* over CUDA IPC) remains available. */
int zkb_init(const int* devices, int ndev);
/* Binds the CALLING THREAD to one of the devices given to zkb_init: from then on every call this thread makes acts on that device
 * alone (its SRS replica, its streams and workspaces; polynomial handles it creates live there and must be used from threads bound
 * to the same device) and never fans out.  Threads that never call this act on the home device and get the sharded behaviour of
 * zkb_init.  This is how one process keeps resident work on every GPU: one host thread per device. */
int zkb_thread_bind_device(int device);
/* Thresholds of the multi-device paths, as log2 sizes (0 = default): smallest per-device share of points for which a commit
 * is sharded (default 16), smallest transform whose batch is split by column (16), smallest single transform sharded over
 * the devices (22; needs >= 2 NTT passes, i.e. >= 2^11).  Environment: ZKB_MULTI_MIN_SHARE_LOG, ZKB_MULTI_NTT_MIN_LOG,
 * ZKB_MULTI_DIST_MIN_LOG. */
int zkb_multi_device_set(int msm_min_share_log, int batch_ntt_min_log, int dist_ntt_min_log);
/* devices bound by zkb_init, home device first; returns their number (0 before zkb_init) */
int zkb_bound_devices(int* devices, int capacity);
void zkb_shutdown(void);
const char* zkb_last_error(void);
/* Number of visible CUDA devices (0 if none) — never fails. */
int zkb_device_count(void);
const char* zkb_version(void);

/* ---- MSM: halo2_proofs::arithmetic::best_multiexp / ParamsKZG::{commit, commit_lagrange} ------------------ */

/* best_multiexp(coeffs, bases) -> G1.  scalars: n x 4, bases: n x 8, out: 12. */
int zkb_msm_g1(const uint64_t* scalars, const uint64_t* bases, size_t n, uint64_t out_jac[12]);

/* Keep an SRS (ParamsKZG.g or ParamsKZG.g_lagrange) resident in HBM; returns an opaque handle. */
int zkb_srs_register(const uint64_t* bases, size_t n, uint64_t* handle);
int zkb_srs_release(uint64_t handle);
/* ParamsKZG::commit / commit_lagrange: best_multiexp(scalars, srs[..n]).  n <= registered length. */
int zkb_msm_g1_srs(uint64_t handle, const uint64_t* scalars, size_t n, uint64_t out_jac[12]);
/* Same over the sub-range srs[offset .. offset+n) — the point-range shard of a multi-GPU MSM. */
int zkb_msm_g1_srs_range(uint64_t handle, size_t offset, const uint64_t* scalars, size_t n, uint64_t out_jac[12]);
/* ncols commitments against the same SRS (the prover's per-column commit loop).  out: ncols x 12. */
int zkb_msm_g1_srs_batch(uint64_t handle, const uint64_t* const* scalars, size_t ncols, size_t n, uint64_t* out_jac);

/* SRS window table: T[w][i] = 2^(c*w) * P_i for every digit window w, kept in HBM next to the bases (W x the SRS size,
 * e.g. 12 x 256 MiB at k = 22).  With it all windows of all scalars share one bucket set, so a commit needs one
 * bucket reduction and no Horner fold (~18 % faster: 2^22 11.5 vs 14.0 ms).  Building it costs about as much as it saves over
 * ~100 commits (2^22: 0.33 s), so the policy is, per zkb_srs_set_precompute(mode) or ZKB_SRS_PRECOMPUTE=0|1|auto:
 *   0 off;  1 eager: build on the first commit against the handle;  2 automatic (default): build once the handle has served
 *   96 commitments — a process that proves once never pays, a long-running prover pays once.
 * zkb_srs_precompute forces the build now and reports the window bits / bytes used (0 / 0 when off or when the table would not
 * fit).  Results are identical either way. */
int zkb_srs_set_precompute(int mode);
int zkb_srs_precompute(uint64_t handle, uint32_t* window_bits, uint64_t* table_bytes);
/* state of a handle without forcing anything: window bits / bytes of its table (0 / 0 if not built) and commitments served */
int zkb_srs_table_info(uint64_t handle, uint32_t* window_bits, uint64_t* table_bytes, uint64_t* commits);

/* Host-side combination of partial results (multi-GPU point-range shards): out = sum of `count` Jacobian points. */
int zkb_g1_sum(const uint64_t* points_jac, size_t count, uint64_t out_jac[12]);

/* out[i] = [s_i] G as G1Affine (n x 8) — the fixed-base multiples ParamsKZG::setup computes; also used to synthesise
 * bases with known discrete logs.  16-bit windows over a table of multiples of G resident in HBM (64 MiB, built on first
 * use), <= 16 mixed additions per scalar, batched normalisation.  The _naive form is the plain double-and-add kernel the
 * table itself is built with (an independent path, kept for cross-checks). */
int zkb_g1_fixed_base_mul(const uint64_t* scalars, size_t n, uint64_t* out_affine);
int zkb_g1_fixed_base_mul_naive(const uint64_t* scalars, size_t n, uint64_t* out_affine);

/* halo2curves Curve::batch_normalize: n Jacobian G1 (n x 12) -> n G1Affine (n x 8), identity -> (0, 0); one field
 * inversion per 16 points (Montgomery's trick).  The step before commitments are written to the transcript. */
int zkb_g1_batch_normalize(const uint64_t* points_jac, size_t n, uint64_t* out_affine);

/* ---- ParamsKZG::<Bn256>::setup(k, rng) — the G1 side (reference call sites: /root/reference/voter/benches/
 * voter_circuit.rs:60, aggregator/benches/state_transition_circuit.rs:64).  s is the toxic-waste scalar the Rust side
 * samples (Montgomery Fr);  g[i] = [s^i] G  and  g_lagrange[i] = [l_i(s)] G,  l_i(s) = omega^i (s^n - 1) / (n (s - omega^i)),
 * n = 2^k.  Either output may be NULL.  Fails with ZKB_ERR_ARG if s is an n-th root of unity (upstream panics there).
 * The G2 elements of the params ([1]G2, [s]G2) are two scalar multiplications and stay on the host.
 * zkb_kzg_setup_resident leaves the arrays in HBM and returns SRS handles (as from zkb_srs_register) — nothing crosses
 * PCIe; zkb_srs_download copies a registered SRS back when the caller wants to serialise it. */
int zkb_kzg_setup(uint32_t k, const uint64_t s[4], uint64_t* g_out, uint64_t* g_lagrange_out);
int zkb_kzg_setup_resident(uint32_t k, const uint64_t s[4], uint64_t* handle_g, uint64_t* handle_g_lagrange);
int zkb_srs_download(uint64_t handle, uint64_t* bases_out, size_t n);
/* best_fft with G = G1 (halo2's FftGroup impl for curve points): out[i] = sum_j [omega^(i j)] P_j, affine in / affine out,
 * log_n <= 26.  zkb_srs_g_to_lagrange is halo2_proofs::arithmetic::g_to_lagrange on a resident SRS: g_lagrange =
 * (1/n) * inverse G1 FFT of g[..2^k] — for params read from a file that carries only the monomial basis. */
/* SRS file -> HBM without a host copy of the arrays: n raw G1Affine (64 B each, Montgomery limbs — the element encoding of
 * halo2's SerdeFormat::RawBytes / RawBytesUnchecked) are read from `path` at byte `offset` through two pinned staging buffers
 * and registered as an SRS handle.  check_points != 0 verifies on the device that every point is the identity or has canonical
 * coordinates on y^2 = x^3 + 3 (what RawBytes checks and RawBytesUnchecked skips).  The container layout (ParamsKZG::write:
 * u32 k, g[2^k], g_lagrange[2^k], then the G2 elements) stays with the caller, who passes the offsets. */
int zkb_srs_load_file(const char* path, uint64_t offset, size_t n, int check_points, uint64_t* handle);
int zkb_g1_ntt(const uint64_t* points_affine, uint64_t* out_affine, const uint64_t omega[4], uint32_t log_n);
int zkb_srs_g_to_lagrange(uint64_t handle_g, uint32_t k, uint64_t* handle_g_lagrange);

/* ---- NTT: halo2_proofs::arithmetic::best_fft and EvaluationDomain::{lagrange_to_coeff, coeff_to_extended,
 *      extended_to_coeff} ----------------------------------------------------------------------------------- */

/* best_fft(a, omega, log_n): in place, natural order in and out, a has 2^log_n elements (1 <= log_n <= 28). */
int zkb_ntt_fr(uint64_t* a, const uint64_t omega[4], uint32_t log_n);
int zkb_ntt_fr_batch(uint64_t* const* cols, size_t ncols, const uint64_t omega[4], uint32_t log_n);

/* EvaluationDomain::lagrange_to_coeff: best_fft(a, omega_inv(k), k) then * 2^-k.  In place, 2^k elements. */
int zkb_lagrange_to_coeff(uint64_t* a, uint32_t k);
int zkb_lagrange_to_coeff_batch(uint64_t* const* cols, size_t ncols, uint32_t k);
/* EvaluationDomain::coeff_to_lagrange: best_fft(a, omega(k), k). */
int zkb_coeff_to_lagrange(uint64_t* a, uint32_t k);

/* EvaluationDomain::coeff_to_extended: in = 2^k coefficients, out = 2^extended_k evaluations on the coset
 * zeta * <extended_omega>  (a_i *= zeta^(i mod 3), zero-extend, best_fft(extended_omega)). */
int zkb_coeff_to_extended(const uint64_t* in, uint64_t* out, uint32_t k, uint32_t extended_k);
int zkb_coeff_to_extended_batch(const uint64_t* const* in, uint64_t* const* out, size_t ncols, uint32_t k,
                                uint32_t extended_k);
/* EvaluationDomain::extended_to_coeff: in place over 2^extended_k elements (inverse FFT, * 2^-extended_k,
 * inverse coset shift).  The caller truncates to n * quotient_poly_degree like the Rust original. */
int zkb_extended_to_coeff(uint64_t* a, uint32_t k, uint32_t extended_k);

/* omega(k) = Fr::ROOT_OF_UNITY^(2^(28-k)) in Montgomery limbs (EvaluationDomain::new) */
int zkb_fr_omega(uint32_t k, uint64_t out[4]);
/* The cube root of unity this library uses as the coset generator of the extended domain (EvaluationDomain::g_coset =
 * Fr::ZETA), Montgomery limbs.  halo2curves releases have shipped either primitive cube root as bn256 Fr::ZETA; the shim
 * asserts at start-up that this equals the linked crate's constant (coeff_to_extended / extended_to_coeff would otherwise
 * evaluate on a different coset than the host-side pieces of the prover). */
int zkb_fr_zeta(uint64_t out[4]);

/* ---- device-resident variants (data already in HBM; `stream` is a cudaStream_t, NULL = default stream) ----- */

/* d_scalars: n x 32 B on the device; result written to host out_jac (synchronises `stream`). */
int zkb_msm_g1_srs_dev(uint64_t handle, size_t offset, const void* d_scalars, size_t n, uint64_t out_jac[12], void* stream);
/* ncols NTTs of 2^log_n elements, column c at d_data + c * 2^log_n * 32; result overwrites d_data
 * (d_scratch must hold the same number of bytes).  Asynchronous on `stream`. */
int zkb_ntt_fr_dev(void* d_data, void* d_scratch, size_t ncols, const uint64_t omega[4], uint32_t log_n, void* stream);
/* d_in: ncols x 2^k, d_out: ncols x 2^extended_k, d_scratch: as d_out.  Asynchronous on `stream`. */
int zkb_coeff_to_extended_dev(const void* d_in, void* d_out, void* d_scratch, size_t ncols, uint32_t k,
                              uint32_t extended_k, void* stream);
int zkb_extended_to_coeff_dev(void* d_data, void* d_scratch, size_t ncols, uint32_t k, uint32_t extended_k, void* stream);
int zkb_lagrange_to_coeff_dev(void* d_data, void* d_scratch, size_t ncols, uint32_t k, void* stream);

/* ---- polynomials resident in HBM ------------------------------------------------------
 * The prover commits, transforms and evaluates each polynomial several times.  Under a handle it is uploaded once and the
 * chain  commit_lagrange -> lagrange_to_coeff -> commit -> coeff_to_extended -> eval_polynomial -> kate_division  runs
 * without PCIe traffic; only 96-byte commitments and 32-byte evaluations come back.  Fixed / permutation polynomials of the
 * proving key can stay resident across proofs the same way.  Handles are process-wide; every op is ordered on the library
 * stream, downloads synchronise.
 *   zkb_poly_commit              ParamsKZG::commit / commit_lagrange of the resident values against a registered SRS
 *   zkb_poly_lagrange_to_coeff   EvaluationDomain::lagrange_to_coeff, in place   (2^k elements)
 *   zkb_poly_coeff_to_lagrange   EvaluationDomain::coeff_to_lagrange, in place
 *   zkb_poly_coeff_to_extended   EvaluationDomain::coeff_to_extended -> new handle with 2^extended_k elements
 *   zkb_poly_extended_to_coeff   EvaluationDomain::extended_to_coeff, in place over 2^extended_k elements
 *   zkb_poly_eval                arithmetic::eval_polynomial(poly, x)
 *   zkb_poly_kate_division       arithmetic::kate_division(poly, b): new handle with len - 1 coefficients of poly / (X - b)
 *   zkb_poly_batch_invert        ff::BatchInvert over the values, in place (zeros stay zero)
 *   zkb_poly_scale_add           poly[i] = poly[i] * k + other[i] (other = 0: scale only) — the multiopen provers' Horner fold
 *                                of their query polynomials with powers of a challenge
 *   zkb_poly_add_const           coefficient 0 += c — f(X) - f(x) ahead of kate_division
 *   zkb_poly_mul                 poly[i] *= other[i]
 *   zkb_poly_prefix_product      in place z[0] = 1, z[i] = prod_{j<i} v[j] — with batch_invert and mul, the grand products
 *                                z of the permutation and lookup arguments (numerators * inverted denominators, scanned) */
int zkb_poly_upload(const uint64_t* values, size_t n, uint64_t* handle);
int zkb_poly_alloc(size_t n, uint64_t* handle); /* zero-filled */
/* Proving-key residency: n raw Fr (32-byte Montgomery limbs — the element encoding of
 * SerdeFormat::RawBytesUnchecked, the format the reference keeps its proving keys in, /root/reference/aggregator/src/wrapper.rs:
 * 970-988, 1006-1034, 1072-1106) are read from `path` at byte `offset` through two pinned staging buffers straight into a
 * polynomial handle.  The fixed / selector / sigma polynomials and cosets, l_0, l_last and l_active of a ProvingKey are loaded
 * once this way and reused by every proof of the IVC loop (wrapper.rs:884-900); the container layout (which polynomial sits
 * where) stays with the caller, who passes offsets — see zksnap-circuits-halo2_b200/plonk.py: ProvingKey.read. */
int zkb_poly_load_file(const char* path, uint64_t offset, size_t n, uint64_t* handle);
/* Host->device bytes the library has copied since load (or since the last call with reset != 0), all devices: lets a caller
 * assert that a second proof against a resident proving key uploads witness data only. */
int zkb_transfer_stats(uint64_t* h2d_bytes, int reset);
int zkb_poly_len(uint64_t handle, size_t* n);
/* poly[offset .. offset + n) = values (host -> device): e.g. the random blinding rows at the end of a product column */
int zkb_poly_write(uint64_t handle, size_t offset, const uint64_t* values, size_t n);
int zkb_poly_download(uint64_t handle, uint64_t* out, size_t n);
int zkb_poly_free(uint64_t handle);
/* new handle with a copy of poly[offset .. offset + n) — the pieces of h(X) after extended_to_coeff (n coefficients each), which the
 * prover commits and opens one by one */
int zkb_poly_slice(uint64_t poly, size_t offset, size_t n, uint64_t* out_handle);
int zkb_poly_commit(uint64_t srs_handle, uint64_t poly, uint64_t out_jac[12]);
int zkb_poly_lagrange_to_coeff(uint64_t poly, uint32_t k);
int zkb_poly_coeff_to_lagrange(uint64_t poly, uint32_t k);
int zkb_poly_coeff_to_extended(uint64_t poly, uint32_t k, uint32_t extended_k, uint64_t* out_handle);
int zkb_poly_extended_to_coeff(uint64_t poly, uint32_t k, uint32_t extended_k);
int zkb_poly_eval(uint64_t poly, const uint64_t x[4], uint64_t out[4]);
int zkb_poly_kate_division(uint64_t poly, const uint64_t b[4], uint64_t* out_handle);
int zkb_poly_batch_invert(uint64_t poly);
int zkb_poly_mul(uint64_t poly, uint64_t other);
/* poly[i] *= table[i mod period] (period a power of two <= 16, table: period x 4 Montgomery limbs): EvaluationDomain::
 * divide_by_vanishing_poly — on the extended coset 1 / (X^n - 1) takes only 2^(extended_k - k) distinct values, so no
 * extended-size array of them is ever built or uploaded */
int zkb_poly_mul_periodic(uint64_t poly, const uint64_t* table, uint32_t period);
int zkb_poly_scale_add(uint64_t poly, const uint64_t k[4], uint64_t other);
int zkb_poly_add_const(uint64_t poly, const uint64_t c[4]);
int zkb_poly_prefix_product(uint64_t poly);
/* plonk::lookup::prover::permute_expression_pair (halo2-axiom plonk/lookup/prover.rs; reached from create_proof,
 * /root/reference/aggregator/src/wrapper.rs:129-137) on resident columns.  `input` / `table`: the theta-compressed input and table
 * expressions (Lagrange basis, same length); rows [0, usable_rows) take part.  New handles receive
 *   permuted_input  A' = the input values sorted ascending by canonical value (Fr's Ord),
 *   permuted_table  S' with S'[i] = A'[i] wherever A'[i] != A'[i-1], and the table elements left over (ascending) on the repeated
 *                   rows, assigned from the last repeated row backwards (upstream's `repeated_input_rows.pop()`),
 * rows >= usable_rows are zero (the caller writes its random blinding rows with zkb_poly_write).  Fails with ZKB_ERR_ARG when an
 * input value does not occur in the table (upstream: Error::ConstraintSystemFailure). */
int zkb_lookup_permute_expression_pair(uint64_t input, uint64_t table, size_t usable_rows, uint64_t* permuted_input,
                                       uint64_t* permuted_table);
/* host-buffer forms of the three helpers (upload + op + download) */
int zkb_fr_eval_polynomial(const uint64_t* coeffs, size_t n, const uint64_t x[4], uint64_t out[4]);
int zkb_fr_kate_division(const uint64_t* coeffs, size_t n, const uint64_t b[4], uint64_t* out /* n - 1 */);
int zkb_fr_batch_invert(uint64_t* values, size_t n);

/* ---- quotient evaluation on resident cosets ---------------------------------------------------
 * halo2-axiom `plonk/evaluation.rs` (un-vendored; reached from create_proof, /root/reference/aggregator/src/wrapper.rs:129-137)
 * compiles the custom gates and the lookup input / table expressions into a `GraphEvaluator`: constants, rotations and a
 * list of calculations over value sources, and `evaluate_h` runs it for every row of the extended domain,
 *     values[idx] = graph.evaluate(&mut data, fixed, advice, instance, challenges, &beta, &gamma, &theta, &y, &values[idx],
 *                                  idx, rot_scale, isize)
 * with rotated rows read at (idx + rotation * rot_scale).rem_euclid(isize).  zkb_graph_evaluate is that loop over polynomials
 * resident in HBM, so the extended cosets never cross PCIe.  The structures carry the GraphEvaluator's own fields:
 *   zkb_value_source   ValueSource::{Constant(i), Intermediate(i), Fixed(col, rot), Advice(col, rot), Instance(col, rot),
 *                      Challenge(i), Beta, Gamma, Theta, Y, PreviousValue}; `rotation` indexes `rotations`
 *   zkb_calculation    Calculation::{Add, Sub, Mul, Square, Double, Negate, Store} writing intermediate `target`;
 *                      Horner(start, parts, factor) is passed as one ZKB_CALC_STORE of `start` followed by one
 *                      ZKB_CALC_MUL_ADD (target = a * b + c: a = Intermediate(target), b = factor, c = part) per part.
 * The result of a row is the target of the last calculation (zero for an empty graph), written to `values`; PreviousValue
 * reads values[idx] before it is overwritten.  Every polynomial must hold exactly the same power-of-two number of elements
 * (the extended domain, isize).  An intermediate must be written before it is read.  Other terms of h(X) (permutation and
 * lookup products, with l_0, l_last, l_active, the sigma cosets and the coset of X held as fixed columns) are graphs over
 * the same sources. */
#define ZKB_SRC_CONSTANT 0
#define ZKB_SRC_INTERMEDIATE 1
#define ZKB_SRC_FIXED 2
#define ZKB_SRC_ADVICE 3
#define ZKB_SRC_INSTANCE 4
#define ZKB_SRC_CHALLENGE 5
#define ZKB_SRC_BETA 6
#define ZKB_SRC_GAMMA 7
#define ZKB_SRC_THETA 8
#define ZKB_SRC_Y 9
#define ZKB_SRC_PREVIOUS 10
typedef struct zkb_value_source {
    uint32_t kind;     /* ZKB_SRC_* */
    uint32_t index;    /* constant / intermediate / column / challenge index */
    uint32_t rotation; /* index into `rotations` (columns only) */
} zkb_value_source;
#define ZKB_CALC_ADD 0
#define ZKB_CALC_SUB 1
#define ZKB_CALC_MUL 2
#define ZKB_CALC_SQUARE 3
#define ZKB_CALC_DOUBLE 4
#define ZKB_CALC_NEGATE 5
#define ZKB_CALC_STORE 6
#define ZKB_CALC_MUL_ADD 7
typedef struct zkb_calculation {
    uint32_t op;     /* ZKB_CALC_* */
    uint32_t target; /* intermediate written */
    zkb_value_source a, b, c;
} zkb_calculation;
typedef struct zkb_graph {
    const zkb_calculation* calculations;
    size_t num_calculations;
    uint32_t num_intermediates;
    const uint64_t* constants; /* num_constants x 4, Montgomery Fr */
    size_t num_constants;
    const int32_t* rotations;
    size_t num_rotations;
} zkb_graph;
typedef struct zkb_graph_inputs {
    const uint64_t* fixed;    /* polynomial handles, one per column */
    size_t num_fixed;
    const uint64_t* advice;
    size_t num_advice;
    const uint64_t* instance;
    size_t num_instance;
    const uint64_t* challenges; /* num_challenges x 4 */
    size_t num_challenges;
    const uint64_t* beta; /* 4 u64 each; may be NULL when the graph does not use the source */
    const uint64_t* gamma;
    const uint64_t* theta;
    const uint64_t* y;
    int32_t rot_scale; /* 1 << (extended_k - k) */
} zkb_graph_inputs;
/* values: polynomial handle of isize elements, read as PreviousValue and overwritten with the row results. */
int zkb_graph_evaluate(const zkb_graph* graph, const zkb_graph_inputs* inputs, uint64_t values);
/* Device pointers and the caller's stream (asynchronous): the fixed / advice / instance arrays of `inputs` hold device addresses
 * (16-byte aligned) instead of handles, d_values holds `rows` elements.  window == 0: the whole domain, rows a power of two,
 * rotated rows wrap.  window != 0: a ROW WINDOW — one rank's share of the extended domain when the quotient evaluation is
 * sharded by rows over the GPUs of a box: every column buffer holds halo_lo + rows + halo_hi elements (the last
 * rows of the previous shard, the shard, the first rows of the next), row i reads element halo_lo + i + rotation * rot_scale,
 * nothing wraps, and a rotation that reaches outside the halo is ZKB_ERR_ARG.  The halo exchange between neighbouring ranks is
 * the path's one collective step (zksnap-circuits-halo2_b200/distributed.py: ShardedQuotient). */
int zkb_graph_evaluate_dev(const zkb_graph* graph, const zkb_graph_inputs* inputs, void* d_values, size_t rows, int window,
                           size_t halo_lo, size_t halo_hi, void* stream);
/* What the last zkb_graph_evaluate was lowered to: device instructions, shared-memory slots after liveness analysis,
 * distinct polynomials read, algorithmic bytes per row (32 x (polynomials read + previous value + result)). */
int zkb_graph_last_info(uint32_t* instructions, uint32_t* slots, uint32_t* polys_read, uint32_t* bytes_per_row);

/* ---- one NTT sharded over the GPUs of a box ("single NTT larger than one GPU's share") ------------
 * One process per GPU.  Rank r holds the contiguous slice [r N/G, (r+1) N/G) of the natural-order input and receives
 * the same slice of the natural-order output of best_fft(a, omega, log_n).  The exchange is not a separate collective:
 * pass 0 of the NTT gathers its tiles from every rank's HBM and the last pass scatters its results to the owning rank,
 * both over NVLink peer mappings (CUDA IPC), with device-side barriers in between.
 *   zkb_dist_create   allocates this rank's symmetric slices (input, work, output: 3 x 2^max_log_n/world x 32 B) and
 *                     returns an opaque blob of ZKB_DIST_HANDLE_BYTES to be all-gathered by the caller (any transport:
 *                     torch.distributed, MPI, a file);
 *   zkb_dist_connect  maps every peer's slices from the gathered blobs (world x ZKB_DIST_HANDLE_BYTES, rank order).
 * All ranks must call the zkb_dist_ntt_* functions collectively, in the same order.  world is 1, 2, 4 or 8; the transform
 * needs at least two passes (log_n >= 11).  A rank that never arrives makes the others fail with ZKB_ERR_CUDA after a
 * bounded wait instead of hanging the GPU. */
#define ZKB_DIST_HANDLE_BYTES 256
int zkb_dist_create(int rank, int world, uint32_t max_log_n, uint8_t* handle_out);
int zkb_dist_connect(const uint8_t* all_handles);
int zkb_dist_destroy(void);
/* One process, several devices (zkb_init with a device list): allocates and connects the symmetric slices of every bound device
 * (largest power of two <= their number) through peer access — no handles to exchange.  The host-buffer entry points do this
 * themselves for transforms of >= 2^22 elements; call it to use zkb_dist_ntt_fr_dev / zkb_dist_buffers / zkb_dist_status from one
 * host thread per device (zkb_thread_bind_device), rank = position of the device in zkb_init's list. */
int zkb_dist_create_inprocess(uint32_t max_log_n);
/* host slices (N/world x 4 u64 each); synchronous */
int zkb_dist_ntt_fr(const uint64_t* in_slice, uint64_t* out_slice, const uint64_t omega[4], uint32_t log_n);
/* device slices; NULL d_in_slice / d_out_slice = use the symmetric slices directly (zkb_dist_buffers).  Asynchronous on
 * `stream`; call zkb_dist_status(stream) to synchronise and learn whether every barrier completed. */
int zkb_dist_ntt_fr_dev(const void* d_in_slice, void* d_out_slice, const uint64_t omega[4], uint32_t log_n, void* stream);
int zkb_dist_buffers(void** d_in_slice, void** d_out_slice, size_t* slice_bytes);
int zkb_dist_status(void* stream);
/* Bound of the device-side barrier waits (default 120 s, or ZKB_DIST_TIMEOUT_MS; 0 restores the default).  When a peer never
 * arrives the waiting rank's remaining NTT passes return at once (no half-exchanged data is read or written), zkb_dist_status /
 * zkb_dist_ntt_fr report ZKB_ERR_CUDA once and clear the flag.  The ranks' barrier epochs are then out of step: every rank must
 * zkb_dist_destroy and create / connect again before the next collective (the in-process form does that by itself). */
int zkb_dist_set_timeout_ms(uint64_t ms);

/* ---- host memory and the transfer scheduler --------------------------------------------------------------------
 * The host-buffer batch entry points (zkb_*_batch, and the single-column forms through them) run a three-stream
 * pipeline: column group i+1 is copied host->device while group i computes and group i-1 is copied device->host.
 * Caller memory that is page-locked (cudaMallocHost, or registered with zkb_host_register — e.g. the prover's
 * long-lived advice / extended-polynomial Vecs) is DMA'd directly; pageable memory is staged through internal
 * pinned buffers by two pools of host threads (stage-in on the calling thread, stage-out on a drainer thread so that the two host
 * copies overlap; ZKB_STAGE_THREADS threads per pool, default clamp(cores/4, 2, 8)). */
int zkb_host_register(void* ptr, size_t bytes);
int zkb_host_unregister(void* ptr);
/* depth: column groups in flight, 1 = serial, 0 = default (3); group_bytes: output bytes per group, 0 = automatic
 * (a quarter of the batch, clamped to [8 MiB, 256 MiB]). */
int zkb_pipeline_set(int depth, size_t group_bytes);
/* Host-scalar commits (zkb_msm_g1_srs / _range) of >= 2^21 points against an SRS with a window table are cut into
 * point-range slices: slice i+1 is uploaded while slice i runs its digit / sort / accumulate kernels; later slices fill
 * a scratch bucket array that is folded into the main one, and the last one reduces.  slices: 0 = automatic (2 for
 * page-locked scalars, 4 for pageable ones, more above 2^24 points), 1 = upload everything first. */
int zkb_msm_set_slices(int slices);

/* ---- tuning and measurement -------------------------------------------------------------------------------- */

/* MSM window bits / level-0 chunk length override (0 = automatic). */
int zkb_msm_set_params(uint32_t window_bits, uint32_t chunk);
int zkb_msm_get_params(size_t n, uint32_t* window_bits, uint32_t* num_windows, uint32_t* chunk);
/* Bucket additions the last MSM on the home device performed = its non-zero signed digits (zero digits are skipped, which is
 * most of a witness column): the work figure to use for roofline fractions of non-uniform scalar distributions. */
int zkb_msm_last_entries(uint64_t* entries);

/* Per-kernel CUDA-event timers.  Names: "msm_digits", "msm_sort", "msm_accumulate", "msm_reduce",
 * "ntt_pass", "graph_evaluate", "dist_ntt_pass0", "dist_ntt_middle", "dist_ntt_final", "dist_barrier".  zkb_prof_get returns the summed milliseconds and launch count since the last reset. */
int zkb_prof_enable(int on);
int zkb_prof_reset(void);
int zkb_prof_get(const char* name, double* total_ms, uint64_t* launches);
/* Measured integer-pipe peak of this GPU: independent unrolled mad.wide.u32 (32x32+64) chains, MACs per second —
 * the denominator of the MSM roofline. */
int zkb_measure_imad_peak(double* wide_macs_per_s);
/* Device self-test of the field arithmetic (halo2curves bn256::{Fr, Fq} mul / add / sub / square / neg / double / to_repr):
 * out[i] = a[i] (op) b[i] on the GPU's PTX carry chains.  field: 0 Fr, 1 Fq; op: 0 mul, 1 add, 2 sub, 3 sqr(a), 4 neg(a),
 * 5 double(a), 6 from-Montgomery(a); 7 / 8 / 9: the lazy mul / add / sub used inside the bucket accumulation and the NTT
 * butterflies (inputs only < 2M, result reduced for comparison), 10: the raw lazy product (must be < 2M).  Montgomery limbs. */
int zkb_field_vec_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* Device self-test of the MSM's bucket sort (csrc/bucket_sort.cuh): n (key, value) pairs in host memory are sorted on the GPU by the
 * low key_bits (<= 24) bits of the key and copied back; the order of the values inside one key is unspecified.  tile: entries per
 * CTA tile (0 = automatic; otherwise a multiple of 1024 up to 8192) — lets the tests exercise every tile / segment geometry. */
int zkb_bucket_sort_pairs(uint32_t* keys, uint32_t* vals, size_t n, uint32_t key_bits, uint32_t tile);
/* Kernels launched by this library since load (the bench's gpu_launches claim). */
uint64_t zkb_launch_count(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */
