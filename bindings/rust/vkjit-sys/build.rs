// Points rustc at libvkjit_b200.so; set VKJIT_B200_LIB_DIR to the directory holding it
// (vkjit_b200/ inside the repository after `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    if let Ok(dir) = std::env::var("VKJIT_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=vkjit_b200");
    println!("cargo:rerun-if-env-changed=VKJIT_B200_LIB_DIR");
}
