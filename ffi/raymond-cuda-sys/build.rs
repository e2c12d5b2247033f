fn main() {
    let dir = std::env::var("RAYMOND_CUDA_LIB_DIR").expect("set RAYMOND_CUDA_LIB_DIR to the directory that holds libraymond_cuda.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=raymond_cuda");
    println!("cargo:rerun-if-env-changed=RAYMOND_CUDA_LIB_DIR");
}
