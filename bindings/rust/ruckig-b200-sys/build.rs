// RUCKIG_B200_LIB_DIR = the directory that holds libruckig_b200.so (built by `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    println!("cargo:rerun-if-env-changed=RUCKIG_B200_LIB_DIR");
    if let Ok(dir) = std::env::var("RUCKIG_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=ruckig_b200");
}
