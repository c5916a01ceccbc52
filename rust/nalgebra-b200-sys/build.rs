// Builds libnalgebra_b200.so from the CUDA sources with nvcc for sm_100a (the command line of
// nalgebra_b200/build.py) and tells cargo to link it.  NALGEBRA_B200_LIB_DIR skips the build and links a prebuilt one.
use std::{env, path::PathBuf, process::Command};

fn main() {
    println!("cargo:rerun-if-env-changed=NALGEBRA_B200_LIB_DIR");
    if let Ok(dir) = env::var("NALGEBRA_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=nalgebra_b200");
        return;
    }
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("nalgebra_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = Vec::new();
    for entry in std::fs::read_dir(&csrc).expect("csrc") {
        let p = entry.unwrap().path();
        println!("cargo:rerun-if-changed={}", p.display());
        if p.extension().map_or(false, |e| e == "cu") {
            let obj = out.join(p.file_stem().unwrap()).with_extension("o");
            let ok = Command::new(&nvcc)
                .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"])
                .args(["-Xcompiler", "-fPIC,-fvisibility=hidden", "-c"])
                .arg(&p).arg("-o").arg(&obj)
                .status().expect("nvcc").success();
            assert!(ok, "nvcc failed on {}", p.display());
            objs.push(obj);
        }
    }
    let lib = out.join("libnalgebra_b200.so");
    let ok = Command::new(&nvcc).args(["-shared", "-cudart", "static", "-o"]).arg(&lib).args(&objs)
        .args(["-lpthread", "-ldl", "-lrt"]).status().expect("nvcc link").success();
    assert!(ok, "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=nalgebra_b200");
}
