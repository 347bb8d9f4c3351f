// Link against the prebuilt shared library (built by nobs_whisper_b200/csrc/Makefile with nvcc for sm_100a).
// NOBS_WHISPER_B200_LIB_DIR must point at the directory that holds libnobswhisper_b200.so.
fn main() {
    let dir = std::env::var("NOBS_WHISPER_B200_LIB_DIR").unwrap_or_else(|_| "../../nobs_whisper_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=nobswhisper_b200");
    println!("cargo:rerun-if-env-changed=NOBS_WHISPER_B200_LIB_DIR");
}
