// UNVERIFIED (no Rust toolchain in the build image).
fn main() {
    // BENDY2D_B200_LIB_DIR = directory holding libbendy2d_b200.so (bendy2d_b200/lib in this repo)
    if let Ok(dir) = std::env::var("BENDY2D_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=bendy2d_b200");
    println!("cargo:rerun-if-env-changed=BENDY2D_B200_LIB_DIR");
}
