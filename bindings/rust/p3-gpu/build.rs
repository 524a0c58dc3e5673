// Links libp3gpu.so (built by plonky3_b200/csrc/build.sh).  P3GPU_LIB_DIR points at the directory that holds it.
fn main() {
    let dir = std::env::var("P3GPU_LIB_DIR").unwrap_or_else(|_| "../../../plonky3_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=p3gpu");
    println!("cargo:rerun-if-env-changed=P3GPU_LIB_DIR");
}
