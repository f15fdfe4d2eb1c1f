// build.rs for thomasantony/splat with the b200 feature.
// NOT compiled in this repository: the image has no Rust toolchain.  Kept verbatim with INTEGRATION.md.
use std::{env, path::PathBuf};

fn main() {
    // SPLAT_B200_DIR = checkout of this repository
    let root = PathBuf::from(env::var("SPLAT_B200_DIR").expect("set SPLAT_B200_DIR"));
    println!("cargo:rustc-link-search=native={}", root.join("splat_b200").display());
    println!("cargo:rustc-link-lib=dylib=splat_b200");
    println!("cargo:rerun-if-changed={}", root.join("include/splat.h").display());
    bindgen::Builder::default()
        .header(root.join("include/splat.h").to_str().unwrap())
        .allowlist_function("splat_.*")
        .allowlist_type("splat_.*")
        .allowlist_var("SPLAT_.*")
        .generate()
        .expect("bindgen on splat.h")
        .write_to_file(PathBuf::from(env::var("OUT_DIR").unwrap()).join("splat_sys.rs"))
        .unwrap();
}
