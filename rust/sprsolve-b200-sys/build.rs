// Link against the C-ABI library built by `python -m sprsolve_b200.build`
// (sprsolve_b200/lib/libsprsolve_b200.so).  Set SPRSOLVE_B200_LIB_DIR to its directory.
fn main() {
    let dir = std::env::var("SPRSOLVE_B200_LIB_DIR")
        .unwrap_or_else(|_| "../../sprsolve_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=sprsolve_b200");
    println!("cargo:rerun-if-env-changed=SPRSOLVE_B200_LIB_DIR");
}
