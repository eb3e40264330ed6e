// Links librt_b200.so. RT_B200_LIB_DIR must point at the directory holding it
// (ray-tracer-s8_b200/lib after `make -C ray-tracer-s8_b200/csrc`).
fn main() {
    let dir = std::env::var("RT_B200_LIB_DIR").unwrap_or_else(|_| "../../lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=rt_b200");
    println!("cargo:rerun-if-env-changed=RT_B200_LIB_DIR");
}
