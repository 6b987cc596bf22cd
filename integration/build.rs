// integration/build.rs - cargo build script a maintainer adds next to Cargo.toml so that `#[link(name = "vors_b200")]` in
// src/core/track/b200.rs resolves.  UNBUILT here (no cargo in the build image).  VORS_B200_LIB = directory holding
// libvors_b200.so (this repository: visual-odometry-rs_b200/lib).
fn main() {
    let dir = std::env::var("VORS_B200_LIB").unwrap_or_else(|_| "../visual-odometry-rs_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=vors_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=VORS_B200_LIB");
}
