// Links the shim against the in-tree CUDA library.  BROTLI_B200_LIB_DIR overrides the default location
// (<repo>/rust-brotli_b200, where `python __graft_entry__.py build` leaves libbrotli_b200.so).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("BROTLI_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..").join("rust-brotli_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=brotli_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=BROTLI_B200_LIB_DIR");
}
