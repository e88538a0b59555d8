// Point BLAST_CUDA_LIB_DIR at the directory holding libblast_cuda.so (audio_decoder_b200/ in this repo).
fn main() {
    if let Ok(dir) = std::env::var("BLAST_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=blast_cuda");
}
