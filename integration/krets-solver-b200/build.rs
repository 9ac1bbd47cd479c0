// Links libkrets_b200.so (built by `make -C krets_b200/csrc` in the krets_b200 repository).
// KRETS_B200_LIB_DIR points at the directory that holds it.
fn main() {
    if let Ok(dir) = std::env::var("KRETS_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=krets_b200");
    println!("cargo:rerun-if-env-changed=KRETS_B200_LIB_DIR");
}
