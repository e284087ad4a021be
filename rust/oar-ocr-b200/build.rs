// Links liboar_b200.so (built by `python -m oar_ocr_b200.build` or `__graft_entry__.build()` of the B200 repository:
// nvcc -gencode arch=compute_100a,code=sm_100a, static cudart, no other dependency).
// OAR_B200_LIB_DIR = directory that holds liboar_b200.so.
fn main() {
    println!("cargo:rerun-if-env-changed=OAR_B200_LIB_DIR");
    if let Ok(dir) = std::env::var("OAR_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=oar_b200");
}
