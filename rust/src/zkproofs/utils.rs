//! `compute_digest` (reference src/zkproofs/utils.rs:9-22) on the device: SHA-256 over the concatenation of
//! `BigInt::to_bytes()` of every item (minimal big-endian, zero -> 0x00), result as a BigInt.
use std::borrow::Borrow;

use curv::arithmetic::traits::*;
use curv::BigInt;

use crate::engine::{limbs_for_bits, pack, Engine};
use crate::ffi;

pub fn compute_digest<IT>(it: IT) -> BigInt
where
    IT: Iterator,
    IT::Item: Borrow<BigInt>,
{
    let items: Vec<BigInt> = it.map(|x| x.borrow().clone()).collect();
    let limbs = limbs_for_bits(items.iter().map(|x| x.bit_length()).max().unwrap_or(1));
    let mut digest = [0u8; 32];
    Engine::with(|eng| {
        let rows = pack(items.iter(), limbs);
        // count = 0 hashes the empty string; the kernel takes it as an empty transcript
        eng.check(unsafe { ffi::zkp_sha256_transcript(eng.h, rows.as_ptr(), limbs as i32, items.len() as i32, 1, digest.as_mut_ptr()) });
    });
    BigInt::from_bytes(&digest)
}
