"""CPU oracle for the linea-stark-prover hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, with Python big integers, the algorithm the reference
(`/root/reference`, distributed-lab/linea-stark-prover) runs when it calls
`p3_uni_stark::prove` / `verify` on its permutation AIR
(reference `bin/src/main.rs:80-96`, `bin/src/config.rs:9-25`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import anything below `oracle/`.  The product
(`linea-stark-prover_b200/`) never does: it is CUDA or nothing.

PARITY UNPINNED.  The proving arithmetic of the reference lives in a git
dependency that is not vendored (`github.com/distributed-lab/Plonky3` rev
f888f9014203e321371b7088cac0569769b2fb77, `Cargo.lock:505`; ark-ff 0.5.0,
ark-bls12-377 0.5.0) and no Rust toolchain exists in this image, so the
restatement follows the published Plonky3 algorithms of that era
(SURVEY.md Appendix A).  The reference ships no tests, fixtures or golden
vectors.  What *is* pinned: the field constants (modulus, R, R^2, -1/r mod
2^64, 22*R, the 2^47-th root of unity) against the values published in the
arkworks `ark-bls12-377` sources -- see `oracle/field.py` and
`tests/test_oracle_field.py`.  Everything fork-specific (S-box degree, the
width-3 linear layers, `sample_bits`, transcript order in `Pcs::open`) is a
parameter here and in the CUDA library, never a baked-in constant.
"""
