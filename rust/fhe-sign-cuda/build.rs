fn main() {
    // point FHE_SIGN_CUDA_LIB at the directory holding libfhe_sign_cuda.so (fhe_sign_b200/lib)
    if let Ok(dir) = std::env::var("FHE_SIGN_CUDA_LIB") {
        println!("cargo:rustc-link-search=native={}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=fhe_sign_cuda");
}
