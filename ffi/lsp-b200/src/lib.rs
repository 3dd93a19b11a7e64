//! Plonky3 trait implementations over `liblsp_b200.so` -- the drop-in for the aliases of the reference's
//! `bin/src/config.rs:9-25` and the `prove(..)` call of `bin/src/main.rs:80-86`.  NOT COMPILED in the build image (no
//! rustc, fork not vendored): written against upstream Plonky3 of the fork's era; `FORK:` marks every
//! line that depends on something only the fork's source can confirm.
use core::{mem::{align_of, size_of}, ptr};
use std::ffi::CStr;

use air::{air_lookup::AirLookupConfig, air_permutation::AirPermutationConfig, AirConfig};
use lsp_b200_sys as sys;
use p3_bls12_377_fr::Bls12_377Fr as Val;
use p3_commit::{Mmcs, Pcs, TwoAdicMultiplicativeCoset};
use p3_dft::TwoAdicSubgroupDft;
use p3_field::{Field, FieldAlgebra, TwoAdicField};
use p3_fri::{BatchOpening, CommitPhaseProofStep, FriConfig, FriProof, QueryProof};
use p3_matrix::{bitrev::{BitReversableMatrix, BitReversedMatrixView}, dense::RowMajorMatrix, Dimensions, Matrix};
use p3_symmetric::{CryptographicHasher, Hash, Permutation};
use p3_uni_stark::{Commitments, OpenedValues, Proof, StarkGenericConfig};

// `Bls12_377Fr` wraps `ark_ff::Fp256<MontBackend<FrConfig,4>>` = `BigInt<4>([u64;4])`, Montgomery form
// (trace/src/permutation.rs:3,102): the very [u64;4] the library reads and writes.  FORK: the wrapper is not known to be
// repr(transparent), so the layout is asserted instead of assumed.
const _: () = assert!(size_of::<Val>() == 32 && align_of::<Val>() == 8);
fn limbs(v: &[Val]) -> *const u64 { v.as_ptr().cast() }
fn limbs_mut(v: &mut [Val]) -> *mut u64 { v.as_mut_ptr().cast() }
fn zeros(n: usize) -> Vec<Val> { vec![Val::ZERO; n] }   // FORK: `FieldAlgebra::ZERO` in this era

/// One CUDA device + stream + Poseidon2 parameters.  One per host thread (calls are serialised on its stream).
pub struct GpuCtx { raw: *mut sys::lsp_ctx }
unsafe impl Send for GpuCtx {}
impl Drop for GpuCtx { fn drop(&mut self) { unsafe { sys::lsp_ctx_destroy(self.raw) } } }
impl GpuCtx {
    /// `constants`: what `Perm::new_from_rng(8, 22, rng)` drew (main.rs:49), in draw order: 4x3 initial external,
    /// 4x3 terminal external, 22 internal.  FORK: the S-box degree and the internal diagonal (1,1,2) live in the fork.
    pub fn new(device: i32, constants: &[Val], sbox_degree: i32) -> Result<Self, String> {
        let mut raw = ptr::null_mut();
        if unsafe { sys::lsp_ctx_create(device, &mut raw) } != 0 { return Err("no CUDA device: the backend has no CPU fallback".into()); }
        let ctx = GpuCtx { raw };
        let diag_m1 = [Val::ONE, Val::ONE, Val::TWO];
        ctx.check(unsafe { sys::lsp_set_poseidon2(raw, 3, sbox_degree, 8, 22, limbs(constants), limbs(&diag_m1)) })?;
        // never assumed: the coset shift and the two-adic root come from the host's own field type
        ctx.check(unsafe { sys::lsp_set_field_consts(raw, limbs(&[Val::GENERATOR]), limbs(&[Val::two_adic_generator(47)])) })?;
        Ok(ctx)
    }
    fn check(&self, rc: i32) -> Result<(), String> {
        if rc == 0 { Ok(()) } else { Err(unsafe { CStr::from_ptr(sys::lsp_last_error(self.raw)) }.to_string_lossy().into_owned()) }
    }
    /// The prover side of Plonky3 panics on misuse (`assert!`/`unwrap`, trace/src/permutation.rs:18-20): so does the shim.
    fn must(&self, rc: i32) { self.check(rc).unwrap_or_else(|e| panic!("lsp_b200: {e}")) }

    /// run the host's own `Perm`, `Hash`, `Dft` beside the device once per process and abort on
    /// the first mismatch -- this is how the fork-only details (S-box degree, matrices, generators) get caught.
    pub fn parity_probe<P: Permutation<[Val; 3]>, H: CryptographicHasher<Val, [Val; 1]>, D: TwoAdicSubgroupDft<Val>>(&self, perm: &P, hash: &H, dft: &D) {
        let x: Vec<Val> = (1..=24u64).map(|i| Val::from_canonical_u64(i.wrapping_mul(0x9e3779b97f4a7c15) % 0xffff_fffb)).collect();
        let mut got = zeros(3);
        self.must(unsafe { sys::lsp_poseidon2_permute(self.raw, limbs(&x[..3]), limbs_mut(&mut got), 1) });
        assert_eq!(perm.permute([x[0], x[1], x[2]]).to_vec(), got, "Poseidon2 permutation differs: S-box degree / linear layers?");
        let mut dig = zeros(3);
        self.must(unsafe { sys::lsp_hash_rows(self.raw, limbs(&x[..21]), 3, 7, limbs_mut(&mut dig)) });
        for r in 0..3 { assert_eq!(hash.hash_iter(x[7 * r..7 * r + 7].iter().copied())[0], dig[r], "sponge differs"); }
        let m = RowMajorMatrix::new(x[..24].to_vec(), 3);
        let want = dft.coset_lde_batch(m.clone(), 2, Val::GENERATOR).bit_reverse_rows().to_row_major_matrix();
        assert_eq!(GpuDft::new(self).coset_lde_batch(m, 2, Val::GENERATOR).bit_reverse_rows().to_row_major_matrix(), want, "coset LDE differs");
    }
}

/// A device-resident matrix; freed with its context still alive.
pub struct DevMat<'c> { ctx: &'c GpuCtx, raw: *mut sys::lsp_mat }
impl Drop for DevMat<'_> { fn drop(&mut self) { unsafe { sys::lsp_mat_free(self.ctx.raw, self.raw) } } }
impl<'c> DevMat<'c> {
    pub fn upload(ctx: &'c GpuCtx, m: &RowMajorMatrix<Val>) -> Self {
        let mut raw = ptr::null_mut();
        ctx.must(unsafe { sys::lsp_mat_upload(ctx.raw, limbs(&m.values), m.height(), m.width(), &mut raw) });
        DevMat { ctx, raw }
    }
    pub fn download(&self) -> RowMajorMatrix<Val> {
        let (h, w) = unsafe { (sys::lsp_mat_rows(self.raw), sys::lsp_mat_width(self.raw)) };
        let mut v = zeros(h * w);
        self.ctx.must(unsafe { sys::lsp_mat_download(self.ctx.raw, self.raw, limbs_mut(&mut v)) });
        RowMajorMatrix::new(v, w)
    }
}

/// `Dft = Radix2DitParallel<Val>` (config.rs:22) -> `GpuDft`.  The PCS only calls `coset_lde_batch`.
#[derive(Clone)]
pub struct GpuDft<'c> { ctx: &'c GpuCtx }
impl<'c> GpuDft<'c> { pub fn new(ctx: &'c GpuCtx) -> Self { GpuDft { ctx } } }
impl Default for GpuDft<'_> { fn default() -> Self { unimplemented!("a GpuDft is bound to a GpuCtx: construct it with GpuDft::new") } }
impl TwoAdicSubgroupDft<Val> for GpuDft<'_> {
    type Evaluations = BitReversedMatrixView<RowMajorMatrix<Val>>;
    fn dft_batch(&self, mat: RowMajorMatrix<Val>) -> Self::Evaluations { self.coset_lde_batch(mat, 0, Val::ONE) }
    fn coset_lde_batch(&self, mat: RowMajorMatrix<Val>, added_bits: usize, shift: Val) -> Self::Evaluations {
        let d = DevMat::upload(self.ctx, &mat);
        let mut out = ptr::null_mut();
        self.ctx.must(unsafe { sys::lsp_coset_lde_batch(self.ctx.raw, d.raw, added_bits as i32, limbs(&[shift]), &mut out, ptr::null_mut()) });
        // the library returns the storage of `.bit_reverse_rows().to_row_major_matrix()`: view it back as natural order
        BitReversedMatrixView::new(DevMat { ctx: self.ctx, raw: out }.download())
    }
}

/// `ValMmcs = ChallengeMmcs = MerkleTreeMmcs<Val,Val,Hash,Compress,1>` (config.rs:19-20) -> `GpuMmcs`.  `get_matrices` must
/// hand out host references, so the prover data keeps a host mirror: use `gpu_prove` for full device residency.
#[derive(Clone)]
pub struct GpuMmcs<'c> { ctx: &'c GpuCtx }
pub struct GpuProverData<'c, M> { host: Vec<M>, dev: Vec<DevMat<'c>>, tree: *mut sys::lsp_tree, ctx: &'c GpuCtx }
impl<M> Drop for GpuProverData<'_, M> { fn drop(&mut self) { unsafe { sys::lsp_tree_free(self.ctx.raw, self.tree) } } }
impl<'c> Mmcs<Val> for GpuMmcs<'c> {
    type ProverData<M> = GpuProverData<'c, M>;
    type Commitment = Hash<Val, Val, 1>;
    type Proof = Vec<[Val; 1]>;
    type Error = ();
    fn commit<M: Matrix<Val>>(&self, inputs: Vec<M>) -> (Self::Commitment, Self::ProverData<M>) {
        let dev: Vec<DevMat> = inputs.iter().map(|m| DevMat::upload(self.ctx, &m.to_row_major_matrix())).collect();
        let raws: Vec<*const sys::lsp_mat> = dev.iter().map(|d| d.raw as *const _).collect();
        let (mut root, mut tree) = (zeros(1), ptr::null_mut());
        self.ctx.must(unsafe { sys::lsp_merkle_commit(self.ctx.raw, raws.as_ptr(), raws.len() as i32, limbs_mut(&mut root), &mut tree) });
        (Hash::from([root[0]]), GpuProverData { host: inputs, dev, tree, ctx: self.ctx })
    }
    fn open_batch<M: Matrix<Val>>(&self, index: usize, data: &Self::ProverData<M>) -> (Vec<Vec<Val>>, Self::Proof) {
        let widths: Vec<usize> = data.host.iter().map(|m| m.width()).collect();
        let log_h = unsafe { sys::lsp_merkle_height(data.tree) }.trailing_zeros() as usize;
        let (mut rows, mut sib) = (zeros(widths.iter().sum()), zeros(log_h));
        self.ctx.must(unsafe { sys::lsp_merkle_open_batch(self.ctx.raw, data.tree, index, limbs_mut(&mut rows), limbs_mut(&mut sib)) });
        let mut o = 0;
        let opened = widths.iter().map(|&w| { let r = rows[o..o + w].to_vec(); o += w; r }).collect();
        (opened, sib.into_iter().map(|d| [d]).collect())
    }
    fn get_matrices<'a, M: Matrix<Val>>(&self, data: &'a Self::ProverData<M>) -> Vec<&'a M> { data.host.iter().collect() }
    fn verify_batch(&self, commit: &Self::Commitment, dims: &[Dimensions], index: usize, opened: &[Vec<Val>], proof: &Self::Proof) -> Result<(), ()> {
        let row: Vec<Val> = opened.iter().flatten().copied().collect();
        let sib: Vec<Val> = proof.iter().map(|d| d[0]).collect();
        let root: [Val; 1] = (*commit).into();
        let log_h = dims.iter().map(|d| d.height).max().unwrap().trailing_zeros();
        match unsafe { sys::lsp_merkle_verify_batch(self.ctx.raw, limbs(&root), log_h, index, limbs(&row), row.len(), limbs(&sib)) } { 0 => Ok(()), _ => Err(()) }
    }
}

/// Device-resident `Pcs::commit` for `TwoAdicFriPcs` (config.rs:24-25): LDE and tree never leave the GPU.
pub struct GpuPcs<'c> { pub ctx: &'c GpuCtx, pub log_blowup: usize }
pub struct GpuPcsData<'c> { pub ldes: Vec<DevMat<'c>>, pub coeffs: Vec<DevMat<'c>>, tree: *mut sys::lsp_tree, ctx: &'c GpuCtx }
impl Drop for GpuPcsData<'_> { fn drop(&mut self) { unsafe { sys::lsp_tree_free(self.ctx.raw, self.tree) } } }
impl<'c> GpuPcs<'c> {
    pub fn natural_domain_for_degree(&self, degree: usize) -> TwoAdicMultiplicativeCoset<Val> {
        TwoAdicMultiplicativeCoset { log_n: degree.trailing_zeros() as usize, shift: Val::ONE }
    }
    /// `commit(vec![(domain, evals)])`: shift = GENERATOR / domain.shift, as TwoAdicFriPcs does.
    pub fn commit(&self, evaluations: Vec<(TwoAdicMultiplicativeCoset<Val>, RowMajorMatrix<Val>)>) -> (Hash<Val, Val, 1>, GpuPcsData<'c>) {
        let (mut ldes, mut coeffs) = (Vec::new(), Vec::new());
        for (domain, evals) in &evaluations {
            let shift = Val::GENERATOR * domain.shift.inverse();
            let (d, mut lde, mut co) = (DevMat::upload(self.ctx, evals), ptr::null_mut(), ptr::null_mut());
            self.ctx.must(unsafe { sys::lsp_coset_lde_batch(self.ctx.raw, d.raw, self.log_blowup as i32, limbs(&[shift]), &mut lde, &mut co) });
            ldes.push(DevMat { ctx: self.ctx, raw: lde });
            coeffs.push(DevMat { ctx: self.ctx, raw: co });
        }
        let raws: Vec<*const sys::lsp_mat> = ldes.iter().map(|d| d.raw as *const _).collect();
        let (mut root, mut tree) = (zeros(1), ptr::null_mut());
        self.ctx.must(unsafe { sys::lsp_merkle_commit(self.ctx.raw, raws.as_ptr(), raws.len() as i32, limbs_mut(&mut root), &mut tree) });
        (Hash::from([root[0]]), GpuPcsData { ldes, coeffs, tree, ctx: self.ctx })
    }
    /// `get_evaluations_on_domain(data, idx, quotient_domain)`: the first N*q rows of the committed LDE, bit-reversed order.
    pub fn get_evaluations_on_domain(&self, data: &GpuPcsData<'c>, idx: usize, domain: TwoAdicMultiplicativeCoset<Val>) -> BitReversedMatrixView<RowMajorMatrix<Val>> {
        let (rows, w) = (1usize << domain.log_n, unsafe { sys::lsp_mat_width(data.ldes[idx].raw) });
        let mut v = zeros(rows * w);
        self.ctx.must(unsafe { sys::lsp_mat_download_rows(self.ctx.raw, data.ldes[idx].raw, 0, rows, limbs_mut(&mut v)) });
        BitReversedMatrixView::new(RowMajorMatrix::new(v, w))
    }
    /// Opened values of matrix `idx` at `z` (barycentric `interpolate_coset` in the reference; same field element).
    pub fn eval_at(&self, data: &GpuPcsData<'c>, idx: usize, z: Val) -> Vec<Val> {
        let mut out = zeros(unsafe { sys::lsp_mat_width(data.coeffs[idx].raw) });
        self.ctx.must(unsafe { sys::lsp_eval_at(self.ctx.raw, data.coeffs[idx].raw, limbs(&[z]), limbs_mut(&mut out)) });
        out
    }
}

struct CfgStore { ids: Vec<Vec<u32>>, perms: Vec<sys::lsp_perm_air_cfg>, lookups: Vec<sys::lsp_lookup_air_cfg> }
/// `Vec<AirConfig>` (what `RawTrace::push_traces` returns, trace/src/lib.rs:62-92) -> the C structs, lookups first.
fn c_configs(cfgs: &[AirConfig]) -> CfgStore {
    let mut s = CfgStore { ids: Vec::new(), perms: Vec::new(), lookups: Vec::new() };
    let keep = |v: &[usize], s: &mut CfgStore| -> *const u32 { s.ids.push(v.iter().map(|&x| x as u32).collect()); s.ids.last().unwrap().as_ptr() };
    for c in cfgs {
        match c {
            AirConfig::Permutation(AirPermutationConfig { a_columns_ids, b_columns_ids, b_inverse_id, check_id }) => {
                let (a, b) = (keep(a_columns_ids, &mut s), keep(b_columns_ids, &mut s));
                s.perms.push(sys::lsp_perm_air_cfg { n_cols: a_columns_ids.len() as u32, a_ids: a, b_ids: b, b_inverse_id: *b_inverse_id as u32, check_id: *check_id as u32 });
            }
            AirConfig::Lookup(l @ AirLookupConfig { .. }) => {
                let flat_b: Vec<usize> = l.b_columns_ids.iter().flatten().copied().collect();
                let (a, b, bf, bi, oc) = (keep(&l.a_columns_ids, &mut s), keep(&flat_b, &mut s), keep(&l.b_filter_id, &mut s), keep(&l.b_inverses_id, &mut s), keep(&l.occurrences_id, &mut s));
                s.lookups.push(sys::lsp_lookup_air_cfg { n_a_cols: l.a_columns_ids.len() as u32, a_ids: a, n_tables: l.b_columns_ids.len() as u32,
                    n_b_cols: l.b_columns_ids[0].len() as u32, b_ids: b, a_filter_id: l.a_filter_id as u32, b_filter_ids: bf,
                    a_inverses_id: l.a_inverses_id as u32, b_inverses_ids: bi, occurrences_ids: oc, check_id: l.check_id as u32 });
            }
        }
    }
    s
}
fn c_fri<M>(f: &FriConfig<M>) -> sys::lsp_fri_config {
    sys::lsp_fri_config { log_blowup: f.log_blowup as u32, log_final_poly_len: f.log_final_poly_len as u32, num_queries: f.num_queries as u32, proof_of_work_bits: f.proof_of_work_bits as u32 }
}

/// Replacement for `prove(&config, &air, &mut challenger, trace, &publics)` at bin/src/main.rs:80-86: one call, device
/// resident end to end, transcript included (`HashChallenger` with empty initial state, main.rs:78).
pub fn gpu_prove<SC, M>(ctx: &GpuCtx, fri: &FriConfig<M>, cfgs: &[AirConfig], trace: RowMajorMatrix<Val>, publics: &[Val; 2]) -> Proof<SC>
where SC: StarkGenericConfig<Challenge = Val>, SC::Pcs: Pcs<Val, SC::Challenger, Commitment = Hash<Val, Val, 1>, Proof = FriProof<Val, GpuMmcs<'static>, Val, Vec<BatchOpening<Val, GpuMmcs<'static>>>>> {
    let (c, f, n, w) = (c_configs(cfgs), c_fri(fri), trace.height(), trace.width());
    let log_q = unsafe { sys::lsp_air_log_quotient_degree_cfg(c.lookups.as_ptr(), c.lookups.len() as i32, c.perms.as_ptr(), c.perms.len() as i32) } as usize;
    let words = unsafe { sys::lsp_proof_words(n.trailing_zeros(), w as u32, log_q as u32, &f) };
    let mut flat = zeros(words / 4);
    ctx.must(unsafe { sys::lsp_prove_air(ctx.raw, &f, limbs(&trace.values), n, w, c.lookups.as_ptr(), c.lookups.len() as i32, c.perms.as_ptr(),
                                         c.perms.len() as i32, limbs(publics), limbs_mut(&mut flat), words, ptr::null_mut()) });
    proof_from_flat::<SC>(&flat, n.trailing_zeros() as usize, w, 1 << log_q, &f)
}

/// The flat array `lsp_prove_*` writes -> `p3_uni_stark::Proof`.  FORK: field names as upstream of
/// the era (`Proof{commitments, opened_values, opening_proof, degree_bits}`, `FriProof{commit_phase_commits, query_proofs,
/// final_poly, pow_witness}`, `QueryProof{input_proof, commit_phase_openings}`).  FORK: upstream declares `Proof`'s fields
/// `pub(crate)`; unless the fork exposes them, build the struct through its `Deserialize` impl instead (serialise the parts
/// assembled here with the same serde format and read them back as `Proof<SC>`) -- the field values are the ones below.
pub fn proof_from_flat<SC: StarkGenericConfig<Challenge = Val>>(flat: &[Val], log_n: usize, w: usize, q: usize, f: &sys::lsp_fri_config) -> Proof<SC>
where SC::Pcs: Pcs<Val, SC::Challenger, Commitment = Hash<Val, Val, 1>, Proof = FriProof<Val, GpuMmcs<'static>, Val, Vec<BatchOpening<Val, GpuMmcs<'static>>>>> {
    let (log_l, rounds) = (log_n + f.log_blowup as usize, log_n - f.log_final_poly_len as usize);
    let n_final = 1usize << (f.log_blowup + f.log_final_poly_len);
    let mut at = 0;
    let mut take = |k: usize| { let s = &flat[at..at + k]; at += k; s };
    let (trace_commit, quot_commit) = (take(1)[0], take(1)[0]);
    let (local, next) = (take(w).to_vec(), take(w).to_vec());
    let chunks: Vec<Vec<Val>> = take(q).iter().map(|&v| vec![v]).collect();
    let commits: Vec<Hash<Val, Val, 1>> = take(rounds).iter().map(|&r| Hash::from([r])).collect();
    let final_poly = take(n_final)[..1usize << f.log_final_poly_len].to_vec();   // the observed tail beyond it is zero (A.10)
    let pow_witness = take(1)[0];
    let digests = |s: &[Val]| -> Vec<[Val; 1]> { s.iter().map(|&d| [d]).collect() };
    let query_proofs = (0..f.num_queries).map(|_| {
        let _index = take(1);                                      // redundant: the verifier samples it
        let trace_open = BatchOpening { opened_values: vec![take(w).to_vec()], opening_proof: digests(take(log_l)) };
        let quot_open = BatchOpening { opened_values: take(q).iter().map(|&v| vec![v]).collect(), opening_proof: digests(take(log_l)) };
        let commit_phase_openings = (0..rounds).map(|r| CommitPhaseProofStep { sibling_value: take(1)[0], opening_proof: digests(take(log_l - 1 - r)) }).collect();
        QueryProof { input_proof: vec![trace_open, quot_open], commit_phase_openings }
    }).collect();
    Proof {
        commitments: Commitments { trace: Hash::from([trace_commit]), quotient_chunks: Hash::from([quot_commit]) },
        opened_values: OpenedValues { trace_local: local, trace_next: next, quotient_chunks: chunks },
        opening_proof: FriProof { commit_phase_commits: commits, query_proofs, final_poly, pow_witness },
        degree_bits: log_n,
    }
}

/// `verify` (main.rs:88-96) on the device from the flat array: Ok, or the LSP_VERIFY_* code of the first failing check
/// (the `VerificationError` / `FriError` variant the Plonky3 verifier would return).  The unchanged Plonky3 `verify` on
/// the `Proof` rebuilt by `proof_from_flat` is the independent check.
pub fn gpu_verify<M>(ctx: &GpuCtx, fri: &FriConfig<M>, cfgs: &[AirConfig], log_n: usize, width: usize, publics: &[Val; 2], flat: &[Val]) -> Result<(), i32> {
    let (c, f) = (c_configs(cfgs), c_fri(fri));
    match unsafe { sys::lsp_verify_air(ctx.raw, &f, log_n as u32, width, c.lookups.as_ptr(), c.lookups.len() as i32, c.perms.as_ptr(), c.perms.len() as i32,
                                       limbs(publics), limbs(flat), flat.len() * 4, ptr::null_mut()) } {
        0 => Ok(()),
        rc if rc > 0 => Err(rc),
        rc => { ctx.must(rc); unreachable!() }
    }
}
