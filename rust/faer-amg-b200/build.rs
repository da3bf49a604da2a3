// Links the prebuilt C-ABI library (built by `make -C faer_amg_b200/csrc`).
fn main() {
    let dir = std::env::var("FAMG_LIB_DIR").unwrap_or_else(|_| "../../faer_amg_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=famg");
}
