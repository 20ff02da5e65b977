// Links the CUDA engine built by `make lib` (or __graft_entry__.build()).
fn main() {
    let dir = std::env::var("ZKP_B200_LIB_DIR").unwrap_or_else(|_| "../zk-paillier_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=zkp_b200");
}
