// Links libb200zk.so.  The reference's own build.rs (/root/reference/build.rs:1-6) only builds the SP1 guest; this
// crate carries the link step so that the reference needs no build-script change beyond the dependency line.
//   B200ZK_DIR  directory holding libb200zk.so (default: ../../zksnark-finalproject_b200 next to this crate);
//               set B200ZK_BUILD=1 to (re)build it first with `python build.py` (needs nvcc, sm_100a).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let here = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let dir = env::var("B200ZK_DIR")
        .map(PathBuf::from)
        .unwrap_or_else(|_| here.join("../../zksnark-finalproject_b200"));
    if env::var("B200ZK_BUILD").map(|v| v == "1").unwrap_or(false) {
        let status = Command::new("python")
            .arg(dir.join("build.py"))
            .status()
            .expect("python build.py (nvcc -gencode arch=compute_100a,code=sm_100a) could not be started");
        assert!(status.success(), "building libb200zk.so failed");
    }
    assert!(
        dir.join("libb200zk.so").exists(),
        "libb200zk.so not found in {} (set B200ZK_DIR or B200ZK_BUILD=1)",
        dir.display()
    );
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=b200zk");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=B200ZK_DIR");
    println!("cargo:rerun-if-env-changed=B200ZK_BUILD");
    println!("cargo:rerun-if-changed=../../include/b200zk.h");
}
