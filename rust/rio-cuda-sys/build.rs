// Links against librio_cuda.so built by `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
fn main() {
    let dir = std::env::var("RIO_CUDA_LIB_DIR").unwrap_or_else(|_| "../../rio_rs_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=rio_cuda");
    println!("cargo:rerun-if-env-changed=RIO_CUDA_LIB_DIR");
}
