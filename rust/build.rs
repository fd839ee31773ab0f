// build.rs -- compile the CUDA engine with nvcc for sm_100a and link it.
// PMT_ROOT: the checkout of this repository (default: the parent directory of rust/).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let root = env::var("PMT_ROOT").map(PathBuf::from).unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join(".."));
    let src = root.join("plonky2_merkle_trees_b200/csrc/pmt_api.cu");
    let lib = out.join("libpmt.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-ldl", "-o"])
        .arg(&lib)
        .arg(&src)
        .status()
        .expect("nvcc not found (set NVCC)");
    assert!(status.success(), "nvcc failed on {}", src.display());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=pmt");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
    println!("cargo:rerun-if-changed={}", root.join("plonky2_merkle_trees_b200/csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include/pmt.h").display());
}
