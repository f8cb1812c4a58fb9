// Links libdrprg_cuda.so (built in-tree by `python -m drprg_b200.build`).
// DRPRG_CUDA_LIB_DIR overrides the search directory.
fn main() {
    let dir = std::env::var("DRPRG_CUDA_LIB_DIR").unwrap_or_else(|_| "../../drprg_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=drprg_cuda");
    println!("cargo:rerun-if-env-changed=DRPRG_CUDA_LIB_DIR");
}
