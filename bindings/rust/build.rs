// Builds libslzw.so with nvcc for sm_100a (the same recipe as lzw_b200/csrc/Makefile) and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("lzw_b200/csrc");
    let status = Command::new("make").arg("-C").arg(&csrc).arg("-j4").status().expect("make");
    assert!(status.success(), "nvcc build of libslzw.so failed");
    println!("cargo:rustc-link-search=native={}", csrc.display());
    println!("cargo:rustc-link-lib=dylib=slzw");
    for f in ["slzw_api.cu", "encode_kernels.cu", "decode_kernels.cu", "sched_kernels.cu", "slzw_device.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/slzw.h").display());
}
