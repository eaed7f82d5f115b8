// Links libgl_b200.so.  GL_B200_LIB_DIR = directory holding the library (default: the in-tree build,
// ../../plonky2-lib_b200, produced by `python plonky2-lib_b200/build.py`: nvcc, sm_100a).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("GL_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../plonky2-lib_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=gl_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=GL_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/gl_b200.h");
}
