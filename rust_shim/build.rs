fn main() {
    // libaleo_b200.so is built by `make` in the aleo_b200 repository
    let dir = std::env::var("ALEO_B200_LIB_DIR").expect("set ALEO_B200_LIB_DIR to the directory holding libaleo_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=aleo_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
