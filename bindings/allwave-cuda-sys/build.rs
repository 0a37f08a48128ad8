// Links liballwave_cuda.so.  ALLWAVE_CUDA_LIB_DIR points at the directory that holds it (allwave_b200/ of this repository
// after `python -c "import __graft_entry__ as g; g.build()"`); the library itself links cudart.
use std::env;
use std::path::PathBuf;

fn main() {
    println!("cargo:rerun-if-env-changed=ALLWAVE_CUDA_LIB_DIR");
    let dir = env::var("ALLWAVE_CUDA_LIB_DIR")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../allwave_b200"));
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=allwave_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
