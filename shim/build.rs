// Links libzkb200.so.  ZKB200_LIB_DIR points at <repo>/zksnap-circuits-halo2_b200 (where `make` leaves the library).
fn main() {
    let dir = std::env::var("ZKB200_LIB_DIR").unwrap_or_else(|_| "/usr/local/lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zkb200");
    println!("cargo:rerun-if-env-changed=ZKB200_LIB_DIR");
}
