// Links libr4w_b200.so.  R4W_B200_LIB_DIR = directory that holds it (the repository's r4w_b200/ after `make -C r4w_b200/csrc`).
fn main() {
    if let Ok(dir) = std::env::var("R4W_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=r4w_b200");
    println!("cargo:rerun-if-env-changed=R4W_B200_LIB_DIR");
}
