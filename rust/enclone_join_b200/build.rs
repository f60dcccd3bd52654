fn main() {
    // point ENCLONE_JOIN_B200_LIB_DIR at the directory holding libenclone_join_b200.so
    if let Ok(dir) = std::env::var("ENCLONE_JOIN_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=enclone_join_b200");
}
