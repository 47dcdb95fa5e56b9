// Links against the prebuilt libtokamak_b200.so (make -C tokamak-zk-evm_b200).
fn main() {
    let dir = std::env::var("TOKAMAK_B200_LIB_DIR").unwrap_or_else(|_| "../../../lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=tokamak_b200");
    println!("cargo:rerun-if-env-changed=TOKAMAK_B200_LIB_DIR");
}
