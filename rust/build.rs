// Links libtfhe_aes_b200.so.  TFHE_AES_B200_LIB_DIR points at the directory holding it (default: ../tfhe-aes_b200,
// where `make -C tfhe-aes_b200/csrc` puts it).
fn main() {
    let dir = std::env::var("TFHE_AES_B200_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{manifest}/../tfhe-aes_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=tfhe_aes_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=TFHE_AES_B200_LIB_DIR");
}
