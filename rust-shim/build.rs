// Links libb200promql.so (built by greptimedb_b200/csrc/build.sh).  B200PROMQL_LIB_DIR points at the directory that
// holds it; the library itself dlopens libnccl.so.2 on first use of the multi-GPU entry points.
fn main() {
    let dir = std::env::var("B200PROMQL_LIB_DIR").unwrap_or_else(|_| "../greptimedb_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=b200promql");
    println!("cargo:rerun-if-env-changed=B200PROMQL_LIB_DIR");
    println!("cargo:rerun-if-changed=../include/b200promql.h");
}
