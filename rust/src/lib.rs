//! zk-paillier-b200: the `zkproofs::*` call surface of zk-paillier 0.4.4 (reference src/zkproofs/mod.rs:29-43) with
//! the big-integer loops delegated to the CUDA engine `libzkp_b200.so` over `extern "C"` (include/zkp_b200.h).
//!
//! SOURCE ONLY in this repository: the build image has no Rust toolchain and the dependency crates are not vendored,
//! so this crate has never been compiled here.  What keeps it honest: `ffi.rs` is generated from the header
//! (scripts/gen_rust_ffi.py) and tests/test_rust_face.py checks, on the CPU, that every function this crate calls
//! exists in the header with the arity used here, and that every public proof of the reference has its struct,
//! `prove` / `verify` and `*_batch` form.  The C++ mirror under zk-paillier_b200/host/ implements the same
//! conventions and IS exercised against the oracle on the GPU.
mod engine;
pub mod ffi;
mod serialize;
pub mod zkproofs;

pub use engine::Engine;
