// Links against the in-tree shared library built by `python -m cheq_b200.build`.
fn main() {
    let dir = std::env::var("CHEQ_B200_LIB_DIR").unwrap_or_else(|_| "../../cheq_b200/lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=cheq_b200");
}
