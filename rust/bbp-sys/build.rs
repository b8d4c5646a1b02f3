// Links against the prebuilt shared library (built by build.sh with nvcc for sm_100a).
// BBP_B200_LIB_DIR points at the directory holding libbbp_b200.so (default: ../../dusk-blindbidproof_b200).
fn main() {
    let dir = std::env::var("BBP_B200_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../../dusk-blindbidproof_b200", manifest)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=bbp_b200");
    println!("cargo:rerun-if-env-changed=BBP_B200_LIB_DIR");
}
