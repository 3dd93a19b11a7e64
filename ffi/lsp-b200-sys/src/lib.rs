//! One declaration per entry point of `include/lsp_b200.h` that the Rust host uses.  Field elements cross as
//! `*const u64` / `*mut u64`: 4 little-endian limbs of the Montgomery representative, the memory format of
//! `Bls12_377Fr` itself (reference `trace/src/permutation.rs:102`).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

macro_rules! opaque { ($($n:ident),*) => { $(#[repr(C)] pub struct $n { _p: [u8; 0] })* } }
opaque!(lsp_ctx, lsp_mat, lsp_tree, lsp_comm);

pub const LSP_OK: c_int = 0;
pub const LSP_INT_PEAK_FORMS: usize = 4;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct lsp_fri_config { pub log_blowup: u32, pub log_final_poly_len: u32, pub num_queries: u32, pub proof_of_work_bits: u32 }
#[repr(C)]
pub struct lsp_perm_air_cfg { pub n_cols: u32, pub a_ids: *const u32, pub b_ids: *const u32, pub b_inverse_id: u32, pub check_id: u32 }
#[repr(C)]
pub struct lsp_lookup_air_cfg {
    pub n_a_cols: u32, pub a_ids: *const u32, pub n_tables: u32, pub n_b_cols: u32, pub b_ids: *const u32,
    pub a_filter_id: u32, pub b_filter_ids: *const u32, pub a_inverses_id: u32, pub b_inverses_ids: *const u32,
    pub occurrences_ids: *const u32, pub check_id: u32,
}

extern "C" {
    pub fn lsp_abi_version() -> c_int;
    pub fn lsp_ctx_create(device: c_int, out: *mut *mut lsp_ctx) -> c_int;
    pub fn lsp_ctx_destroy(ctx: *mut lsp_ctx);
    pub fn lsp_last_error(ctx: *const lsp_ctx) -> *const c_char;
    pub fn lsp_ctx_sync(ctx: *mut lsp_ctx) -> c_int;
    pub fn lsp_set_poseidon2(ctx: *mut lsp_ctx, width: c_int, sbox_d: c_int, rounds_f: c_int, rounds_p: c_int,
                             constants: *const u64, internal_diag_m1: *const u64) -> c_int;
    pub fn lsp_set_field_consts(ctx: *mut lsp_ctx, generator: *const u64, two_adic_root_2_47: *const u64) -> c_int;
    pub fn lsp_set_transcript_flags(ctx: *mut lsp_ctx, alpha_before_openings: c_int, observe_opened_values: c_int) -> c_int;
    // parity probes
    pub fn lsp_fr_op(ctx: *mut lsp_ctx, op: c_int, a: *const u64, b: *const u64, out: *mut u64, n: usize) -> c_int;
    pub fn lsp_poseidon2_permute(ctx: *mut lsp_ctx, states_in: *const u64, states_out: *mut u64, n: usize) -> c_int;
    pub fn lsp_hash_rows(ctx: *mut lsp_ctx, rowmajor: *const u64, rows: usize, width: usize, digests_out: *mut u64) -> c_int;
    // matrices
    pub fn lsp_mat_upload(ctx: *mut lsp_ctx, rowmajor: *const u64, rows: usize, width: usize, out: *mut *mut lsp_mat) -> c_int;
    pub fn lsp_mat_download(ctx: *mut lsp_ctx, m: *const lsp_mat, rowmajor_out: *mut u64) -> c_int;
    pub fn lsp_mat_download_rows(ctx: *mut lsp_ctx, m: *const lsp_mat, row0: usize, nrows: usize, rowmajor_out: *mut u64) -> c_int;
    pub fn lsp_mat_rows(m: *const lsp_mat) -> usize;
    pub fn lsp_mat_width(m: *const lsp_mat) -> usize;
    pub fn lsp_mat_free(ctx: *mut lsp_ctx, m: *mut lsp_mat);
    // TwoAdicSubgroupDft / Mmcs / pieces of Pcs::open
    pub fn lsp_coset_lde_batch(ctx: *mut lsp_ctx, input: *const lsp_mat, added_bits: c_int, shift: *const u64,
                               out_bitrev: *mut *mut lsp_mat, coeffs_out: *mut *mut lsp_mat) -> c_int;
    pub fn lsp_merkle_commit(ctx: *mut lsp_ctx, mats: *const *const lsp_mat, n_mats: c_int, root_out: *mut u64, out: *mut *mut lsp_tree) -> c_int;
    pub fn lsp_merkle_open_batch(ctx: *mut lsp_ctx, t: *const lsp_tree, index: usize, rows_out: *mut u64, siblings_out: *mut u64) -> c_int;
    pub fn lsp_merkle_verify_batch(ctx: *mut lsp_ctx, root: *const u64, log_height: u32, index: usize, row: *const u64,
                                   row_len: usize, siblings: *const u64) -> c_int;
    pub fn lsp_merkle_height(t: *const lsp_tree) -> usize;
    pub fn lsp_tree_free(ctx: *mut lsp_ctx, t: *mut lsp_tree);
    pub fn lsp_eval_at(ctx: *mut lsp_ctx, coeffs: *const lsp_mat, z: *const u64, values_out: *mut u64) -> c_int;
    pub fn lsp_reduce_openings(ctx: *mut lsp_ctx, ldes: *const *const lsp_mat, points: *const u64, opened: *const *const u64,
                               n_entries: c_int, alpha: *const u64, fri_input_out: *mut *mut lsp_mat) -> c_int;
    pub fn lsp_fri_fold(ctx: *mut lsp_ctx, input: *const lsp_mat, beta: *const u64, out: *mut *mut lsp_mat) -> c_int;
    // prove / verify
    pub fn lsp_air_log_quotient_degree_cfg(lookups: *const lsp_lookup_air_cfg, n_lookups: c_int, perms: *const lsp_perm_air_cfg, n_perms: c_int) -> c_int;
    pub fn lsp_proof_words(log_n: u32, width: u32, log_q: u32, fri: *const lsp_fri_config) -> usize;
    pub fn lsp_prove_air(ctx: *mut lsp_ctx, fri: *const lsp_fri_config, trace: *const u64, rows: usize, width: usize,
                         lookups: *const lsp_lookup_air_cfg, n_lookups: c_int, perms: *const lsp_perm_air_cfg, n_perms: c_int,
                         publics: *const u64, proof_out: *mut u64, proof_words: usize, timings_ms_out: *mut f32) -> c_int;
    pub fn lsp_verify_air(ctx: *mut lsp_ctx, fri: *const lsp_fri_config, log_n: u32, width: usize,
                          lookups: *const lsp_lookup_air_cfg, n_lookups: c_int, perms: *const lsp_perm_air_cfg, n_perms: c_int,
                          publics: *const u64, proof: *const u64, proof_words: usize, device_ms_out: *mut f32) -> c_int;
    // multi-GPU
    pub fn lsp_nccl_unique_id(out: *mut u8) -> c_int;
    pub fn lsp_comm_init_nccl(ctx: *mut lsp_ctx, rank: c_int, world: c_int, unique_id: *const u8, out: *mut *mut lsp_comm) -> c_int;
    pub fn lsp_comm_destroy(comm: *mut lsp_comm);
    pub fn lsp_prove_air_sharded(comm: *mut lsp_comm, fri: *const lsp_fri_config, trace: *const u64, rows: usize, width: usize,
                                 lookups: *const lsp_lookup_air_cfg, n_lookups: c_int, perms: *const lsp_perm_air_cfg, n_perms: c_int,
                                 publics: *const u64, proof_out: *mut u64, proof_words: usize, timings_ms_out: *mut f32) -> c_int;
    // input files and witness (trace/ crate)
    pub fn lsp_cbor_permutation_read_rows(cbor: *const u8, len: usize, min_rows: usize, rows: *mut usize, n_cols: *mut u32,
                                          name: *mut c_char, name_cap: usize, be_rowmajor_out: *mut *mut u8) -> c_int;
    pub fn lsp_cbor_lookup_read_rows(cbor: *const u8, len: usize, min_rows: usize, rows: *mut usize, n_a_cols: *mut u32, n_tables: *mut u32,
                                     n_b_cols: *mut u32, name: *mut c_char, name_cap: usize, be_rowmajor_out: *mut *mut u8) -> c_int;
    pub fn lsp_host_free(p: *mut c_void);
    pub fn lsp_permutation_trace_be(ctx: *mut lsp_ctx, be_rowmajor: *const u8, rows: usize, n_cols: u32, publics: *const u64, trace_out: *mut *mut lsp_mat) -> c_int;
    pub fn lsp_lookup_trace_be(ctx: *mut lsp_ctx, be_rowmajor: *const u8, rows: usize, n_a_cols: u32, n_tables: u32, n_b_cols: u32,
                               publics: *const u64, trace_out: *mut *mut lsp_mat) -> c_int;
    pub fn lsp_mat_hconcat(ctx: *mut lsp_ctx, mats: *const *const lsp_mat, n_mats: c_int, out: *mut *mut lsp_mat) -> c_int;
}
