// build.rs — link the reference against libkta_gpu.so (UNCOMPILED: no Rust toolchain in the build image)
fn main() {
    let dir = std::env::var("KTA_LIB_DIR").unwrap_or_else(|_| "../kafka_topic_analyzer_b200".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=kta_gpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=KTA_LIB_DIR");
}
