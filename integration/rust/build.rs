//! Builds the CUDA static library for sm_100a with nvcc and links it (feature `cuda`).
//! H2V_B200_DIR = path to this repository (default: two levels up from this crate).
use std::{env, path::PathBuf, process::Command};

fn main() {
    if env::var_os("CARGO_FEATURE_CUDA").is_none() {
        return;
    }
    let root = env::var("H2V_B200_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..")
    });
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = root.join("halo2-verifier_b200/csrc");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let lib = out.join("libh2v_b200.a");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-lib"])
        .arg("-I").arg(root.join("include"))
        .arg("-I").arg(&csrc)
        .arg(csrc.join("kernels.cu"))
        .arg(csrc.join("plan_build.cpp"))
        .arg("-o").arg(&lib)
        .status()
        .expect("nvcc not found: the cuda feature has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    for f in ["kernels.cu", "plan_build.cpp", "stages.cuh", "field.cuh", "curve.cuh", "tower.cuh", "hash.cuh", "plan.h",
              "pairing_cta.cuh", "exchange.cuh", "timeline.cuh", "pairing_lin.inc"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/h2v.h").display());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=h2v_b200");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
