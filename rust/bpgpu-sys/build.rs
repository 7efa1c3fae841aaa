// Links libbpgpu.so, built in-tree by `python __graft_entry__.py` (nvcc, sm_100a).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("BPGPU_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../mpc_bulletproof_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=bpgpu");
    println!("cargo:rerun-if-env-changed=BPGPU_LIB_DIR");
}
