// Links libdiffsol_b200.so, built in-tree by `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
// DIFFSOL_B200_LIB_DIR overrides the default location relative to this crate.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("DIFFSOL_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../diffsol_b200/_lib")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=diffsol_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=DIFFSOL_B200_LIB_DIR");
}
