// Builds libcapgpu.so from the CUDA sources with nvcc (no `cc` crate: nothing can be downloaded
// in the target environment).  NOT BUILT in this repository's container (no Rust toolchain);
// the same nvcc command line is what cap_b200/build.py runs.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("cap_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let lib = out.join("libcapgpu.so");
    let mut cmd = Command::new(nvcc);
    cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--shared",
              "-Xcompiler", "-fPIC", "-o"]).arg(&lib);
    for f in ["capi.cu", "ntt.cu", "msm.cu", "poly.cu", "prover.cu", "formats.cu"] {
        cmd.arg(csrc.join(f));
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    let status = cmd.status().expect("nvcc not found (set NVCC)");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=capgpu");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=dl");
}
