// build.rs — build the CUDA core as a static library
use std::process::Command;

fn main() {
    let out = std::env::var("OUT_DIR").unwrap();
    let root = std::env::var("RT2025_ROOT").unwrap_or_else(|_| "../..".into()); // repo that holds include/ and csrc/
    let csrc = format!("{root}/raytracer-2025_b200/csrc");
    let mut objs = Vec::new();
    for s in ["api.cu", "kernels.cu", "lbvh.cu", "compile.cpp", "bvh_build.cpp"] {
        let o = format!("{out}/{s}.o");
        let ok = Command::new("nvcc")
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false"])
            .args(["-Xcompiler", "-fPIC,-ffp-contract=off,-fopenmp,-O3"])
            .arg(format!("-I{root}/include"))
            .arg(format!("-I{csrc}"))
            .args(["-c", &format!("{csrc}/{s}"), "-o", &o])
            .status()
            .expect("nvcc not found")
            .success();
        assert!(ok, "nvcc failed on {s}");
        objs.push(o);
        println!("cargo:rerun-if-changed={csrc}/{s}");
    }
    assert!(Command::new("ar").arg("crs").arg(format!("{out}/librt2025.a")).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={out}");
    println!("cargo:rustc-link-lib=static=rt2025");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=gomp");
}
