//! Builds librstsr_cuda.so from the CUDA sources with nvcc for sm_100a and links it.
//!
//! The sources live in `csrc/` of this crate (in the development repository: `rstsr_b200/csrc`, the header in
//! `include/rstsr_cuda.h`); the Makefile there is the single source of truth for the nvcc flags:
//!   -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -fmad=false
//! (`-fmad=false`: a + b * c must round twice, like the CPU devices, so elementwise results stay bit-exact).
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    println!("cargo:rerun-if-env-changed=RSTSR_CUDA_LIB_DIR");
    println!("cargo:rerun-if-env-changed=NVCC");
    if let Ok(dir) = env::var("RSTSR_CUDA_LIB_DIR") {
        // feature "prebuilt" or an explicit directory: link what is there
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=rstsr_cuda");
        return;
    }
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("csrc");
    assert!(csrc.join("Makefile").exists(), "csrc/Makefile not found: vendor rstsr_b200/csrc and include/ into the crate");
    println!("cargo:rerun-if-changed={}", csrc.display());
    let jobs = env::var("NUM_JOBS").unwrap_or_else(|_| "8".into());
    let mut make = Command::new("make");
    make.arg("-C").arg(&csrc).arg("-j").arg(&jobs);
    if let Ok(nvcc) = env::var("NVCC") {
        make.arg(format!("NVCC={nvcc}"));
    }
    let status = make.status().expect("failed to run make (is nvcc 12.8+ installed?)");
    assert!(status.success(), "nvcc build of librstsr_cuda.so failed");
    let lib_dir = csrc.join("..").join("lib");
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=rstsr_cuda");
    // the device code needs the CUDA runtime at run time; it is linked into the .so by nvcc
    println!("cargo:rustc-env=RSTSR_CUDA_LIB_DIR={}", lib_dir.display());
}
