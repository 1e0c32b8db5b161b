// build.rs — UNCOMPILED HERE (no Rust toolchain in the build image).
//
// Extends the reference's compile hook (/root/reference/build.rs:1-8, today two macOS link flags) with the step the
// north star names: "a thin extern \"C\" launcher layer that build.rs compiles with nvcc for sm_100a".  Behind the
// `b200` cargo feature it
//   1. runs `make -C $SCIPORT_B200_DIR/csrc` (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo, one
//      translation unit per kernel family; the Makefile is sciport_rs_b200/csrc/Makefile of this repository), and
//   2. links the resulting libsciport_b200.so and the CUDA runtime.
use std::{env, path::PathBuf, process::Command};

pub fn main() {
    // ---- the reference's own lines, unchanged (build.rs:2-7) ----
    #[cfg(all(target_os = "macos", feature = "blas"))]
    println!("cargo:rustc-link-lib=framework=Accelerate");
    #[cfg(target_os = "macos")]
    println!("cargo:rustc-link-arg=-Wl,-rpath,/Library/Developer/CommandLineTools/Library/Frameworks");

    // ---- B200 path ----
    if env::var_os("CARGO_FEATURE_B200").is_none() {
        return;
    }
    let dir = PathBuf::from(env::var("SCIPORT_B200_DIR").expect(
        "SCIPORT_B200_DIR must point at the sciport_rs_b200/ directory (csrc/ with the CUDA sources inside)",
    ));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let jobs = env::var("NUM_JOBS").unwrap_or_else(|_| "8".into());
    let status = Command::new("make")
        .arg("-C")
        .arg(dir.join("csrc"))
        .arg(format!("-j{jobs}"))
        .arg(format!("NVCC={nvcc}"))
        .status()
        .expect("failed to run make for libsciport_b200");
    assert!(status.success(), "nvcc build of libsciport_b200 failed");

    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=sciport_b200");
    if let Ok(cuda) = env::var("CUDA_HOME") {
        println!("cargo:rustc-link-search=native={cuda}/lib64");
    }
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    for f in ["capi.cu", "k_pipeline.cu", "k_ipass.cu", "k_fused.cu", "k_debye.cu", "k_kpass.cu", "k_general.cu",
              "k_sequence.cu", "k_window.cu", "k_airy.cu", "k_cgamma.cu", "k_misc.cu"] {
        println!("cargo:rerun-if-changed={}", dir.join("csrc").join(f).display());
    }
    println!("cargo:rerun-if-env-changed=SCIPORT_B200_DIR");
}
