/* lsp_b200.h -- C ABI of the B200-native proving backend for the Plonky3
 * prover loop that distributed-lab/linea-stark-prover drives on its Linea
 * permutation AIR.
 *
 * The reference has no FFI of its own: its backend-selection point is the set
 * of type aliases in reference `bin/src/config.rs:9-25` and the single
 * `prove(..)` call at `bin/src/main.rs:80-86`.  Each entry point below names
 * the Plonky3 trait method (as instantiated by those aliases) that a Rust
 * `extern "C"` shim would forward to it; INTEGRATION.md shows the shim.
 *
 * Conventions
 *  - Field elements are BLS12-377 Fr in the host type's own memory format:
 *    4 x uint64_t little-endian limbs of the Montgomery representative
 *    (a * 2^256 mod r), fully reduced (`Bls12_377Fr`, `trace/src/permutation.rs:102`).
 *  - Host matrices are row-major (`RowMajorMatrix<Val>`, `trace/src/lib.rs:94-106`).
 *  - Every function returns 0 on success or a negative LSP_ERR_* code; nothing
 *    throws or aborts across the boundary.  `lsp_last_error` gives the text.
 *    (lsp_verify_* additionally return a positive LSP_VERIFY_* reason for a rejected proof.)
 *  - One `lsp_ctx` per host thread; calls on a ctx are serialised on its stream.
 *  - There is no CPU fallback: without a CUDA device `lsp_ctx_create` fails.
 */
#ifndef LSP_B200_H
#define LSP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSP_OK 0
#define LSP_ERR_PARAM (-1)   /* bad argument / unsupported shape            */
#define LSP_ERR_CUDA (-2)    /* CUDA runtime error, see lsp_last_error      */
#define LSP_ERR_STATE (-3)   /* call order (e.g. Poseidon2 constants unset) */
#define LSP_ERR_NOMEM (-4)
#define LSP_ERR_COMM (-5)    /* NCCL / multi-GPU plumbing                   */

#define LSP_ABI_VERSION 1

typedef struct lsp_ctx lsp_ctx;
typedef struct lsp_mat lsp_mat;               /* device-resident matrix of Fr            */
typedef struct lsp_tree lsp_tree;             /* MerkleTreeMmcs::ProverData              */
typedef struct lsp_challenger lsp_challenger; /* device-resident HashChallenger<Val,Hash,1> */

/* ---- context ----------------------------------------------------------- */
int lsp_abi_version(void);
int lsp_ctx_create(int device, lsp_ctx** out);
void lsp_ctx_destroy(lsp_ctx* ctx);
const char* lsp_last_error(const lsp_ctx* ctx);
int lsp_ctx_sync(lsp_ctx* ctx);
/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches). */
uint64_t lsp_kernel_launches(const lsp_ctx* ctx);

/* Per-kernel CUDA-event timing (off by default).  enable = 1 clears the records and times every
 * launch (two events each: ~4 % on a 400-launch prove); enable = 2 times only the Poseidon2
 * leaf-hash launches (the dominant kernel; free).  The report is JSON
 * [{"phase","kernel","launches","ms"},...]. */
int lsp_kernel_timing(lsp_ctx* ctx, int enable);
int lsp_kernel_timing_report(lsp_ctx* ctx, char* buf, size_t cap);
/* Measured rate (32x32->64 multiply-accumulates per second) of each instruction form the multiply can take on
 * this device, with data-dependent operands: [0] IMAD.WIDE.U32 multiply-only + IADD3/IADD3.X accumulate,
 * [1] IMAD.WIDE.U32 with fused 64-bit accumulate, [2] IMAD.WIDE.U32.X (carry in and out), [3] the IMAD + IMAD.HI
 * pair (SASS: profiles/sass_k_int_peak.txt).  lsp_int_peak returns the fastest of them: the denominator of the
 * integer roofline the benchmark reports. */
#define LSP_INT_PEAK_FORMS 4
int lsp_int_peaks(lsp_ctx* ctx, double mac32_per_s[LSP_INT_PEAK_FORMS]);
int lsp_int_peak(lsp_ctx* ctx, double* mac32_per_s);

/* `Perm::new_from_rng(8, 22, &mut rng)` (bin/src/main.rs:49): the host draws the
 * constants and hands them over.  `constants` holds, in draw order,
 * rounds_f/2 x 3 initial external, rounds_f/2 x 3 terminal external, then
 * rounds_p internal constants; `internal_diag_m1` the 3 diagonal entries
 * (matrix = diag + all-ones).  sbox_d in {3,5,7,11,17}. */
int lsp_set_poseidon2(lsp_ctx* ctx, int width, int sbox_d, int rounds_f, int rounds_p,
                      const uint64_t* constants, const uint64_t* internal_diag_m1);

/* Field parameters fixed by the fork-only `p3-bls12-377-fr` crate, which is not available to check:
 * `Bls12_377Fr::GENERATOR` (the coset shift of every LDE `TwoAdicFriPcs` commits: `Val::GENERATOR / domain.shift`)
 * and `two_adic_generator(47)` (every smaller root is a power of it), both as Montgomery limbs.  Defaults:
 * arkworks' FrConfig values, GENERATOR = 22 and TWO_ADIC_ROOT_OF_UNITY = 22^((r-1)/2^47).  The call validates
 * them (non-zero generator outside every two-adic subgroup, primitive 2^47-th root) and drops the cached tables. */
int lsp_set_field_consts(lsp_ctx* ctx, const uint64_t generator[4], const uint64_t two_adic_root_2_47[4]);
/* Transcript order of `TwoAdicFriPcs::open`.  alpha_before_openings != 0: the batching challenge
 * is sampled before the opened values are computed; observe_opened_values != 0: the opened values (trace at zeta,
 * trace at zeta*g, every quotient chunk at zeta) are observed by the challenger.  The pinned fork is (1, 0) -- the
 * default; upstream Plonky3 after early 2025 is (0, 1).  Applies to prove, sharded prove and verify alike. */
int lsp_set_transcript_flags(lsp_ctx* ctx, int alpha_before_openings, int observe_opened_values);

/* ---- parity probes (SURVEY.md section 4 tier 3) --------------------------- */
/* Elementwise Fr ops on host arrays: op 0 add, 1 sub, 2 mul, 3 inverse (b unused),
 * 4 halve (b unused).  Replaces ark-ff `Fp256` arithmetic behind `Val` (bin/src/config.rs:9). */
int lsp_fr_op(lsp_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* `Permutation::permute` of `Perm` on n independent [Fr;3] states. */
int lsp_poseidon2_permute(lsp_ctx* ctx, const uint64_t* states_in, uint64_t* states_out, size_t n);
/* `CryptographicHasher::hash_iter` of `Hash` (bin/src/config.rs:12) on each of
 * `rows` rows of a host row-major matrix. */
int lsp_hash_rows(lsp_ctx* ctx, const uint64_t* rowmajor, size_t rows, size_t width, uint64_t* digests_out);

/* ---- matrices ---------------------------------------------------------- */
int lsp_mat_upload(lsp_ctx* ctx, const uint64_t* rowmajor, size_t rows, size_t width, lsp_mat** out);
int lsp_mat_download(lsp_ctx* ctx, const lsp_mat* m, uint64_t* rowmajor_out);
int lsp_mat_download_rows(lsp_ctx* ctx, const lsp_mat* m, size_t row0, size_t nrows, uint64_t* rowmajor_out);
size_t lsp_mat_rows(const lsp_mat* m);
size_t lsp_mat_width(const lsp_mat* m);
void lsp_mat_free(lsp_ctx* ctx, lsp_mat* m);

/* ---- TwoAdicSubgroupDft<Val> for `Dft` (bin/src/config.rs:22) -------------- */
/* `coset_lde_batch(mat, added_bits, shift)`: out has rows << added_bits rows and
 * holds, at row bitrev(j), the values p_col(shift * omega^j) -- i.e. the storage of
 * `.bit_reverse_rows().to_row_major_matrix()` that TwoAdicFriPcs::commit consumes.
 * `in` is not consumed.  If `coeffs_out` is non-NULL it receives the N x W
 * coefficient matrix of the interpolants (kept by the PCS for openings). */
int lsp_coset_lde_batch(lsp_ctx* ctx, const lsp_mat* in, int added_bits, const uint64_t shift[4],
                        lsp_mat** out_bitrev, lsp_mat** coeffs_out);

/* ---- Mmcs<Val> for `ValMmcs` / `ChallengeMmcs` (bin/src/config.rs:19-20) --- */
/* `commit(Vec<M>)`: all matrices must share one height (the only case on the
 * reference's path).  The tree borrows the matrices; they must outlive it. */
int lsp_merkle_commit(lsp_ctx* ctx, const lsp_mat* const* mats, int n_mats, uint64_t root_out[4], lsp_tree** out);
/* `open_batch(index, &prover_data)`: rows_out receives the opened row of every
 * matrix back to back (sum of widths elements), siblings_out log2(height) digests. */
int lsp_merkle_open_batch(lsp_ctx* ctx, const lsp_tree* t, size_t index, uint64_t* rows_out, uint64_t* siblings_out);
/* One digest layer (0 = leaf digests), for parity tests. */
int lsp_merkle_layer(lsp_ctx* ctx, const lsp_tree* t, int layer, uint64_t* digests_out);
size_t lsp_merkle_height(const lsp_tree* t);
void lsp_tree_free(lsp_ctx* ctx, lsp_tree* t);

/* ---- AIR config: `AirPermutationConfig` (air/src/air_permutation.rs:2-7) ---- */
typedef struct {
    uint32_t n_cols;         /* a_columns_ids.len() == b_columns_ids.len() */
    const uint32_t* a_ids;   /* a_columns_ids */
    const uint32_t* b_ids;   /* b_columns_ids */
    uint32_t b_inverse_id;
    uint32_t check_id;
} lsp_perm_air_cfg;

/* ---- AIR config: `AirLookupConfig` (air/src/air_lookup.rs:2-11), the LogUp argument ---- */
typedef struct {
    uint32_t n_a_cols;              /* a_columns_ids.len()                                   */
    const uint32_t* a_ids;          /* a_columns_ids                                         */
    uint32_t n_tables;              /* b_columns_ids.len()                                   */
    uint32_t n_b_cols;              /* b_columns_ids[t].len(), the same for every table      */
    const uint32_t* b_ids;          /* b_columns_ids, table-major: n_tables x n_b_cols       */
    uint32_t a_filter_id;
    const uint32_t* b_filter_ids;   /* b_filter_id      [n_tables] */
    uint32_t a_inverses_id;
    const uint32_t* b_inverses_ids; /* b_inverses_id    [n_tables] */
    const uint32_t* occurrences_ids;/* occurrences_id   [n_tables] */
    uint32_t check_id;
} lsp_lookup_air_cfg;

/* `quotient_values` of p3-uni-stark with `LineaAIR::eval` -> `eval_permutation`
 * (air/src/lib.rs:47-54,116-167) folded by powers of `alpha`, times 1/Z_H.
 * `lde_bitrev` is the committed trace LDE; the first N<<log_q rows are read.
 * Output: q = 1<<log_q chunk matrices (N x 1, natural order), chunk c = rows
 * c, c+q, ... of the quotient vector (`split_evals`). */
int lsp_quotient_permutation(lsp_ctx* ctx, const lsp_mat* lde_bitrev, int log_n, int log_q,
                             const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4],
                             const uint64_t alpha[4], lsp_mat** chunks_out);

/* The same for a full `LineaAIR` (air/src/lib.rs:47-54): `eval_lookup` (:57-114) for every lookup
 * config, then `eval_permutation` for every permutation config -- the order in which
 * `RawTrace::push_traces` emits them (trace/src/lib.rs:80-89).  log_q = 2 when a lookup is
 * present (its first-row constraint has degree 4), else 1. */
int lsp_quotient_air(lsp_ctx* ctx, const lsp_mat* lde_bitrev, int log_n, int log_q,
                     const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* perms, int n_perms,
                     const uint64_t publics[2][4], const uint64_t alpha[4], lsp_mat** chunks_out);
/* log2 of the number of quotient chunks `prove` uses for this AIR: `get_log_quotient_degree` of p3-uni-stark, i.e.
 * `LineaAIR::eval` (air/src/lib.rs:47-54) evaluated on symbolic degrees, log2_ceil(max(d_max, 2) - 1).  The `_cfg`
 * form walks the actual configs (what prove and verify use); the count form assumes non-empty configs. */
int lsp_air_log_quotient_degree_cfg(const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* perms, int n_perms);
int lsp_air_log_quotient_degree(int n_lookups, int n_perms);

/* ---- FRI pieces of `TwoAdicFriPcs` (bin/src/config.rs:24-25) -------------- */
/* `fold_matrix(beta, m)`: in = vector of 2h elements viewed as h rows of 2; out h elements. */
int lsp_fri_fold(lsp_ctx* ctx, const lsp_mat* in, const uint64_t beta[4], lsp_mat** out);

/* `Pcs::open` of `TwoAdicFriPcs`, piece by piece, for a trait-level drop-in that keeps its own challenger
 * (inside lsp_prove_* the same kernels run without leaving the device).
 * lsp_eval_at: "compute opened values with Lagrange interpolation" (bench.log:34) -- the value at `z` of every column's
 * interpolant, from the coefficient matrix lsp_coset_lde_batch returned (coefficients over the subgroup H_N: for a matrix
 * committed on the domain s*H_N pass z/s).  values_out: width elements.
 * lsp_reduce_openings: "reduce rows" (bench.log:35) -- the FRI input vector
 *     sum over entries e of  alpha^(offset_e) * (sum_c alpha^c m_e[c](x) - sum_c alpha^c y_e[c]) / (x - z_e),   offset_{e+1} = offset_e + width_e
 * over the L rows of the committed LDEs (x = GENERATOR * w_L^bitrev(row)); one entry per (matrix, point) in the order
 * `open` walks them (trace at zeta, trace at zeta*g, then every quotient chunk at zeta). */
int lsp_eval_at(lsp_ctx* ctx, const lsp_mat* coeffs, const uint64_t z[4], uint64_t* values_out);
int lsp_reduce_openings(lsp_ctx* ctx, const lsp_mat* const* ldes, const uint64_t* points /* n_entries x 4 */,
                        const uint64_t* const* opened /* per entry: width values */, int n_entries, const uint64_t alpha[4],
                        lsp_mat** fri_input_out);

/* ---- witness generation ---------------------------------------------------- */
/* `RawPermutationTrace::get_trace` + `RawTrace::get_trace` (trace/src/permutation.rs:24-93,
 * trace/src/lib.rs:94-106): from the a/b input columns (host row-major rows x 2*n_cols, a
 * columns first) and publics = [alpha, delta], builds the rows x (2*n_cols+2) trace
 * a.., b.., 1/(b_comb+delta), running product.  Fails (LSP_ERR_PARAM) when the running
 * product does not end at 1, like the reference's assert (permutation.rs:76-79). */
int lsp_permutation_trace(lsp_ctx* ctx, const uint64_t* ab_rowmajor, size_t rows, uint32_t n_cols,
                          const uint64_t publics[2][4], lsp_mat** trace_out);

/* ---- input wire format: `RawPermutationTrace` as CBOR (trace/src/permutation.rs:9-22) ---------- */
/* `read_file` (:17-22): pure parsing, no GPU.  `_shape` reports the height (tallest column) and the
 * number of a (= b) columns; `_decode` writes every element's 32 big-endian bytes row-major,
 * rows x 2*n_cols (a columns first), zero-padding short columns (`resize`, :134-142). */
int lsp_cbor_permutation_shape(const uint8_t* cbor, size_t len, size_t* rows, uint32_t* n_cols, char* name, size_t name_cap);
int lsp_cbor_permutation_decode(const uint8_t* cbor, size_t len, uint8_t* be_rowmajor, size_t rows, uint32_t n_cols);
/* `_shape` + `_decode` in one call and one structure pass over the file: the output buffer is allocated by the
 * library (malloc) and handed to the caller, who releases it with lsp_host_free.  The structure pass is a parallel
 * scan on all host threads for files in serde's regular layout (~4 GB/s on 16 cores), serial otherwise. */
int lsp_cbor_permutation_read(const uint8_t* cbor, size_t len, size_t* rows, uint32_t* n_cols, char* name, size_t name_cap,
                              uint8_t** be_rowmajor_out);
/* `RawTrace::push_traces` (trace/src/lib.rs:62-79) resizes every sub-trace to the tallest one before its witness is built
 * (`resize`, permutation.rs:134-142: zero rows).  `_decode` therefore accepts any `rows` >= the file's own height, and
 * `_read_rows` allocates max(min_rows, height) rows; the rows past the file's height are zero (filters included). */
int lsp_cbor_permutation_read_rows(const uint8_t* cbor, size_t len, size_t min_rows, size_t* rows, uint32_t* n_cols, char* name,
                                   size_t name_cap, uint8_t** be_rowmajor_out);
void lsp_host_free(void* p);
/* Process-wide switch: after lsp_host_pinned(1) the `_read` functions hand out PAGE-LOCKED buffers (portable across the
 * process's devices), recycled through a small pool on lsp_host_free, so that `lsp_*_trace_be` uploads at PCIe rate and
 * several ranks may read one buffer at once.  lsp_host_pinned(0) returns to malloc and releases the pooled blocks. */
int lsp_host_pinned(int enable);
/* `RawLookupTrace::read_file` (trace/src/lookup.rs:20-44), same conventions.  Row layout of the decoded
 * buffer: a columns, b columns table by table, a_filter, one b_filter per table -- i.e. the first
 * n_a + T*n_b + 1 + T columns of the trace `get_trace` emits (:63-71).  Filter entries the file omits
 * default to one up to the length of the column they guard (:25-41); `resize` padding is zero (:230-246). */
int lsp_cbor_lookup_shape(const uint8_t* cbor, size_t len, size_t* rows, uint32_t* n_a_cols, uint32_t* n_tables,
                          uint32_t* n_b_cols, char* name, size_t name_cap);
int lsp_cbor_lookup_decode(const uint8_t* cbor, size_t len, uint8_t* be_rowmajor, size_t rows, uint32_t n_a_cols,
                           uint32_t n_tables, uint32_t n_b_cols);
int lsp_cbor_lookup_read(const uint8_t* cbor, size_t len, size_t* rows, uint32_t* n_a_cols, uint32_t* n_tables,
                         uint32_t* n_b_cols, char* name, size_t name_cap, uint8_t** be_rowmajor_out);
int lsp_cbor_lookup_read_rows(const uint8_t* cbor, size_t len, size_t min_rows, size_t* rows, uint32_t* n_a_cols, uint32_t* n_tables,
                              uint32_t* n_b_cols, char* name, size_t name_cap, uint8_t** be_rowmajor_out);
/* `get_columns` (`from_be_bytes_mod_order`, :95-118) + `get_trace` (:24-93) on the device: like
 * lsp_permutation_trace, from raw 32-byte big-endian values (any value < 2^256, reduced mod r). */
int lsp_permutation_trace_be(lsp_ctx* ctx, const uint8_t* be_rowmajor, size_t rows, uint32_t n_cols,
                             const uint64_t publics[2][4], lsp_mat** trace_out);

/* `RawLookupTrace::get_trace` (trace/src/lookup.rs:46-176) on the device: from the decoded input
 * (row layout as lsp_cbor_lookup_decode: a.., b.., a_filter, b_filter[T]) and publics = [alpha, delta],
 * builds the rows x (n_a + T*(n_b+3) + 3) lookup trace -- inverses, multiplicities (the count of every
 * enabled A row goes to the first enabled B row holding the same combination) and the running
 * log-derivative sum.  Fails (LSP_ERR_PARAM) when the sum does not end at 0 (the assert of :165-168). */
int lsp_lookup_trace(lsp_ctx* ctx, const uint64_t* in_rowmajor, size_t rows, uint32_t n_a_cols, uint32_t n_tables,
                     uint32_t n_b_cols, const uint64_t publics[2][4], lsp_mat** trace_out);
int lsp_lookup_trace_be(lsp_ctx* ctx, const uint8_t* be_rowmajor, size_t rows, uint32_t n_a_cols, uint32_t n_tables,
                        uint32_t n_b_cols, const uint64_t publics[2][4], lsp_mat** trace_out);
/* `RawTrace::push_traces` (trace/src/lib.rs:62-92): sub-traces of one height side by side, in the order given. */
int lsp_mat_hconcat(lsp_ctx* ctx, const lsp_mat* const* mats, int n_mats, lsp_mat** out);

/* ---- `prove` (bin/src/main.rs:80-86) ------------------------------------- */
typedef struct {
    uint32_t log_blowup;          /* FriConfig.log_blowup        (main.rs:59) */
    uint32_t log_final_poly_len;  /* FriConfig.log_final_poly_len (main.rs:60) */
    uint32_t num_queries;         /* FriConfig.num_queries        (main.rs:61) */
    uint32_t proof_of_work_bits;  /* FriConfig.proof_of_work_bits (main.rs:62) */
} lsp_fri_config;

/* Number of uint64_t words `lsp_prove_permutation` writes for this shape. */
size_t lsp_proof_words(uint32_t log_n, uint32_t width, uint32_t log_q, const lsp_fri_config* fri);

/* Full uni-STARK prove of the permutation AIR, device-resident end to end
 * (commit trace, quotient, commit quotient, open, FRI), transcript included
 * (`HashChallenger<Val,Hash,1>` with empty initial state, main.rs:78).
 * `trace` is a host row-major N x W matrix.  The proof is written as a flat
 * array of Fr (Montgomery limbs) in the order documented in DESIGN.md
 * ("proof layout"); `timings_ms_out` (optional, 8 floats) receives per-stage
 * CUDA-event times named after the reference's tracing spans. */
int lsp_prove_permutation(lsp_ctx* ctx, const lsp_fri_config* fri, const uint64_t* trace, size_t rows,
                          size_t width, const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4],
                          uint64_t* proof_out, size_t proof_words, float* timings_ms_out);
/* Same, for a trace already uploaded with lsp_mat_upload (bench "value" leg). */
int lsp_prove_permutation_dev(lsp_ctx* ctx, const lsp_fri_config* fri, const lsp_mat* trace,
                              const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4],
                              uint64_t* proof_out, size_t proof_words, float* timings_ms_out);

/* `prove` for a `LineaAIR` holding lookup and permutation configs (what the reference's `main`
 * proves at HEAD, bin/src/main.rs:37-43,76-86).  Proof layout as lsp_prove_permutation with
 * q = 1 << lsp_air_log_quotient_degree(..) quotient chunks. */
int lsp_prove_air(lsp_ctx* ctx, const lsp_fri_config* fri, const uint64_t* trace, size_t rows, size_t width,
                  const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* perms, int n_perms,
                  const uint64_t publics[2][4], uint64_t* proof_out, size_t proof_words, float* timings_ms_out);
int lsp_prove_air_dev(lsp_ctx* ctx, const lsp_fri_config* fri, const lsp_mat* trace,
                      const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* perms, int n_perms,
                      const uint64_t publics[2][4], uint64_t* proof_out, size_t proof_words, float* timings_ms_out);

/* ---- `verify` (bin/src/main.rs:88-96) and `Mmcs::verify_batch` ---------------------- */
/* Rejection reasons: POSITIVE return values of lsp_verify_* (0 = proof accepted, negative =
 * LSP_ERR_*: the call itself failed).  They name the checks of p3_uni_stark::verify /
 * TwoAdicFriPcs::verify in the order that verifier meets them; the first failing one is returned. */
#define LSP_VERIFY_INVALID_PROOF_SHAPE 1      /* VerificationError::InvalidProofShape                                  */
#define LSP_VERIFY_TRACE_OPENING 2            /* InvalidOpeningArgument(InputError(..)): trace round, Merkle root mismatch */
#define LSP_VERIFY_QUOTIENT_OPENING 3         /* InvalidOpeningArgument(InputError(..)): quotient-chunk round           */
#define LSP_VERIFY_COMMIT_PHASE_OPENING 4     /* InvalidOpeningArgument(CommitPhaseMmcsError(..))                       */
#define LSP_VERIFY_FINAL_POLY_MISMATCH 5      /* InvalidOpeningArgument(FinalPolyMismatch)                              */
#define LSP_VERIFY_INVALID_POW_WITNESS 6      /* InvalidOpeningArgument(InvalidPowWitness)                              */
#define LSP_VERIFY_OOD_EVALUATION_MISMATCH 7  /* VerificationError::OodEvaluationMismatch                               */
#define LSP_VERIFY_ROOT_MISMATCH 8            /* lsp_merkle_verify_batch only: MerkleTreeError::RootMismatch            */

/* `verify(&config, &air, &mut challenger, &proof, &publics)` on the device: replays the transcript,
 * checks the proof of work, every query's Merkle openings and fold chain against the final polynomial,
 * and the out-of-domain identity  sum_i zp_i(zeta) chunk_i(zeta) = constraints(zeta) / Z_H(zeta).
 * `proof` is the flat array lsp_prove_* wrote (host memory); `log_n` is the proof's `degree_bits`,
 * `width` the AIR width.  `device_ms_out` (optional) receives the CUDA-event time of the whole check,
 * upload included.  Every element must be canonical (< r), as the reference's deserialiser requires: a proof carrying
 * x + r in place of x is INVALID_PROOF_SHAPE.  A proof with zero commit-phase rounds (log_final_poly_len == log_n) is rejected
 * with FINAL_POLY_MISMATCH, as the pinned verifier does (the reduced opening only enters inside a round). */
int lsp_verify_air(lsp_ctx* ctx, const lsp_fri_config* fri, uint32_t log_n, size_t width,
                   const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* perms, int n_perms,
                   const uint64_t publics[2][4], const uint64_t* proof, size_t proof_words, float* device_ms_out);
int lsp_verify_permutation(lsp_ctx* ctx, const lsp_fri_config* fri, uint32_t log_n, size_t width,
                           const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4],
                           const uint64_t* proof, size_t proof_words, float* device_ms_out);
/* `verify_batch(&commit, dims, index, &opened_values, &proof)` for matrices of one height 2^log_height:
 * `row` = the opened rows back to back (what lsp_merkle_open_batch returned), `siblings` = log_height digests.
 * Returns 0 when the recomputed root equals `root`, LSP_VERIFY_ROOT_MISMATCH otherwise. */
int lsp_merkle_verify_batch(lsp_ctx* ctx, const uint64_t root[4], uint32_t log_height, size_t index,
                            const uint64_t* row, size_t row_len, const uint64_t* siblings);

/* ---- proof serialisation (host only) ------------------------------------------------------------------- */
/* The flat array of lsp_prove_* <-> a byte stream in the field order of `p3_uni_stark::Proof` as bincode (fixed-width
 * little-endian integers, a u64 length before every Vec) writes the derived `Serialize`, every field element as its
 * canonical integer in 32 little-endian bytes (ark-serialize's CanonicalSerialize), behind a 32-byte header
 * ("LSPP", version, FriConfig, width, log_q).  The reference never serialises a proof and the fork's serde impl is not
 * available: this is this library's documented format (host/serialize.cu), not a claim about the fork's bytes.
 * `_deserialize` with proof_out == NULL only reports the shape; non-canonical elements, wrong lengths and trailing bytes
 * are LSP_ERR_PARAM.  The per-query indices are not part of `Proof`: the rebuilt flat array carries a marker in those
 * slots and lsp_verify_air substitutes the indices it samples. */
size_t lsp_proof_serialized_bytes(uint32_t log_n, uint32_t width, uint32_t log_q, const lsp_fri_config* fri);
int lsp_proof_serialize(const uint64_t* proof, size_t proof_words, uint32_t log_n, uint32_t width, uint32_t log_q,
                        const lsp_fri_config* fri, uint8_t* out, size_t out_cap, size_t* out_len);
int lsp_proof_deserialize(const uint8_t* bytes, size_t len, uint32_t* log_n_out, uint32_t* width_out, uint32_t* log_q_out,
                          lsp_fri_config* fri_out, uint64_t* proof_out, size_t proof_words_cap, size_t* proof_words_out);

/* ---- multi-GPU: one proof sharded by row ranges of the LDE ---------- */
typedef struct lsp_comm lsp_comm;
/* NCCL bootstrap: rank 0 calls lsp_nccl_unique_id, ships the 128 bytes to the other ranks
 * (torch.distributed / MPI / a file), every rank calls lsp_comm_init_nccl on its own ctx. */
int lsp_nccl_unique_id(uint8_t out[128]);
int lsp_comm_init_nccl(lsp_ctx* ctx, int rank, int world, const uint8_t unique_id[128], lsp_comm** out);
/* All `world` ranks hosted by this process on ctx's device, collectives replaced by device
 * copies: the same sharded code path on a single GPU (tests, 1-GPU boxes). */
int lsp_comm_init_local(lsp_ctx* ctx, int world, lsp_comm** out);
void lsp_comm_destroy(lsp_comm* comm);
/* `prove` sharded over comm's ranks (world any power of two; with more ranks than the 2^log_blowup cosets a rank
 * owns a fraction of one, at least 8 rows of it).  Every rank passes the same trace and receives the same proof,
 * bit-identical to lsp_prove_permutation -- which is this same code with one rank. */
int lsp_prove_permutation_sharded(lsp_comm* comm, const lsp_fri_config* fri, const uint64_t* trace, size_t rows,
                                  size_t width, const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4],
                                  uint64_t* proof_out, size_t proof_words, float* timings_ms_out);
/* Same, for a trace every rank has already uploaded with lsp_mat_upload / lsp_permutation_trace. */
int lsp_prove_permutation_sharded_dev(lsp_comm* comm, const lsp_fri_config* fri, const lsp_mat* trace,
                                      const lsp_perm_air_cfg* cfgs, int n_cfgs, const uint64_t publics[2][4],
                                      uint64_t* proof_out, size_t proof_words, float* timings_ms_out);

/* The same for a `LineaAIR` with lookup configs (4 quotient chunks; needs log_blowup >= 2). */
int lsp_prove_air_sharded(lsp_comm* comm, const lsp_fri_config* fri, const uint64_t* trace, size_t rows, size_t width,
                          const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* perms, int n_perms,
                          const uint64_t publics[2][4], uint64_t* proof_out, size_t proof_words, float* timings_ms_out);
int lsp_prove_air_sharded_dev(lsp_comm* comm, const lsp_fri_config* fri, const lsp_mat* trace,
                              const lsp_lookup_air_cfg* lookups, int n_lookups, const lsp_perm_air_cfg* perms, int n_perms,
                              const uint64_t publics[2][4], uint64_t* proof_out, size_t proof_words, float* timings_ms_out);

#ifdef __cplusplus
}
#endif
#endif /* LSP_B200_H */
