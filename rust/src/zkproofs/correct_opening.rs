//! `CorrectOpening` (reference src/zkproofs/correct_opening.rs:17-30): c == Enc(m, r), over zkp_verify_opening.
use curv::arithmetic::traits::*;
use curv::BigInt;
use paillier::{EncryptionKey, Paillier, Randomness, RawCiphertext, RawPlaintext};

use crate::engine::{fits, pack, Engine};
use crate::ffi;

pub trait CorrectOpening<R, CT> {
    fn verify_opening(ek: &EncryptionKey, m: RawPlaintext, r: &R, c: &CT) -> bool;
}

/// The reference implements the trait for every (R, CT) kzen-paillier can encrypt into; the engine has one ciphertext type,
/// so the shim implements the instance the reference's own test uses (correct_opening.rs:47-56).
impl<'c> CorrectOpening<Randomness, RawCiphertext<'c>> for Paillier {
    fn verify_opening(ek: &EncryptionKey, m: RawPlaintext, r: &Randomness, c: &RawCiphertext<'c>) -> bool {
        verify_opening_batch(ek, &[m.0.into_owned()], &[r.0.clone()], &[c.0.as_ref().clone()])[0]
    }
}

/// Many openings under one key in one device call.
pub fn verify_opening_batch(ek: &EncryptionKey, m: &[BigInt], r: &[BigInt], c: &[BigInt]) -> Vec<bool> {
    assert!(m.len() == r.len() && r.len() == c.len());
    if m.is_empty() {
        return Vec::new();
    }
    Engine::with(|eng| {
        eng.use_key(ek);
        let (nl, nnl) = (eng.nl(), eng.nnl());
        // (m n + 1) % nn depends on m mod n only, r^n mod nn on r mod n only; a c wider than n^2 is never a ciphertext
        let mr: Vec<BigInt> = m.iter().map(|x| x % &ek.n).collect();
        let rr: Vec<BigInt> = r.iter().map(|x| x % &ek.n).collect();
        let zero = BigInt::zero();
        let wide: Vec<bool> = c.iter().map(|x| !fits(x, nnl)).collect();
        let cc = pack(c.iter().zip(&wide).map(|(x, &w)| if w { &zero } else { x }), nnl);
        let mut ok = vec![0u8; m.len()];
        eng.check(unsafe { ffi::zkp_verify_opening(eng.h, m.len() as i32, nl as i32, pack(mr.iter(), nl).as_ptr(), pack(rr.iter(), nl).as_ptr(), cc.as_ptr(), ok.as_mut_ptr()) });
        ok.iter().zip(&wide).map(|(&o, &w)| o == 1 && !w).collect()
    })
}
