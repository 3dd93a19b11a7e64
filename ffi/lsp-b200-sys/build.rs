// LSP_B200_LIB_DIR = directory holding liblsp_b200.so (built by `make -C linea-stark-prover_b200`).
fn main() {
    let dir = std::env::var("LSP_B200_LIB_DIR").expect("set LSP_B200_LIB_DIR to the directory of liblsp_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=lsp_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=LSP_B200_LIB_DIR");
}
