// Points rustc at the directory that holds libfuse_gpu.so (FUSE_GPU_LIB_DIR, default: the in-tree build output).
fn main() {
    let dir = std::env::var("FUSE_GPU_LIB_DIR").unwrap_or_else(|_| format!("{}/../../fuse_query_b200", env!("CARGO_MANIFEST_DIR")));
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=fuse_gpu");
    println!("cargo:rerun-if-env-changed=FUSE_GPU_LIB_DIR");
}
