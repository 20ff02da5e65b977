//! zk-paillier-b200: the `zkproofs::*` call surface of zk-paillier 0.4.4 (src/zkproofs/mod.rs:29-43) with the
//! big-integer loops delegated to the CUDA engine.  SOURCE ONLY in this repository (no Rust toolchain in the
//! build image); the C++ mirror under zk-paillier_b200/host/ is what the tests exercise.
pub mod ffi;
pub mod zkproofs;
