// Links against librio_client.so (g++ -shared rio_rs_b200/csrc/client.cpp; no CUDA).
fn main() {
    let dir = std::env::var("RIO_CLIENT_LIB_DIR").unwrap_or_else(|_| "../../rio_rs_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=rio_client");
    println!("cargo:rerun-if-env-changed=RIO_CLIENT_LIB_DIR");
}
