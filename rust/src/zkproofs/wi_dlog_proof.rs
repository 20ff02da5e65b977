//! CompositeDLogProof (reference src/zkproofs/wi_dlog_proof.rs:23-107): Girault / Pointcheval proof of knowledge of a
//! discrete log modulo a composite N, over zkp_dlog_prove / zkp_dlog_verify (one modulus per statement).
use curv::arithmetic::traits::*;
use curv::BigInt;
use serde::{Deserialize, Serialize};

use super::errors::IncorrectProof;
use crate::engine::{fits, limbs_for_bits, pack, unpack, Engine, Verdict};
use crate::ffi;

const K: usize = 128;
const K_PRIME: usize = 128;
const SAMPLE_S: usize = 256;

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct CompositeDLogProof {
    pub x: BigInt,
    pub y: BigInt,
}

#[allow(non_snake_case)]
#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct DLogStatement {
    pub N: BigInt,
    pub g: BigInt,
    pub ni: BigInt,
}

const ASSERT_PANIC: &str = "assertion failed: N > 2^K, gcd(g, N) == 1, gcd(ni, N) == 1"; // wi_dlog_proof.rs:69-73

impl CompositeDLogProof {
    /// wi_dlog_proof.rs:46-65
    pub fn prove(statement: &DLogStatement, secret: &BigInt) -> CompositeDLogProof {
        Self::prove_batch(std::slice::from_ref(statement), std::slice::from_ref(secret)).pop().unwrap()
    }
    /// wi_dlog_proof.rs:67-91
    pub fn verify(&self, statement: &DLogStatement) -> Result<(), IncorrectProof> {
        Self::verify_batch(&[self], std::slice::from_ref(statement))[0].into_result(ASSERT_PANIC)
    }

    pub fn prove_batch(statement: &[DLogStatement], secret: &[BigInt]) -> Vec<CompositeDLogProof> {
        assert_eq!(statement.len(), secret.len());
        let b = statement.len();
        if b == 0 {
            return Vec::new();
        }
        let bound = BigInt::from(2).pow((K + K_PRIME + SAMPLE_S) as u32);
        let r: Vec<BigInt> = (0..b).map(|_| BigInt::sample_below(&bound)).collect();
        let nl = limbs_for_bits(statement.iter().map(|s| s.N.bit_length().max(s.g.bit_length()).max(s.ni.bit_length())).max().unwrap());
        let sl = limbs_for_bits(secret.iter().map(|s| s.bit_length()).max().unwrap());
        let rl = limbs_for_bits(K + K_PRIME + SAMPLE_S);
        let yl = limbs_for_bits(std::cmp::max(32 * rl, 32 * sl + 256) + 1); // y = r + e * secret, unreduced
        Engine::with(|eng| {
            let (mut x, mut y, mut fault) = (vec![0u32; b * nl], vec![0u32; b * yl], vec![0u8; b]);
            eng.check(unsafe {
                ffi::zkp_dlog_prove(
                    eng.h, b as i32, nl as i32, pack(statement.iter().map(|s| &s.N), nl).as_ptr(), pack(statement.iter().map(|s| &s.g), nl).as_ptr(),
                    pack(statement.iter().map(|s| &s.ni), nl).as_ptr(), pack(secret.iter(), sl).as_ptr(), sl as i32, pack(r.iter(), rl).as_ptr(), rl as i32,
                    yl as i32, x.as_mut_ptr(), y.as_mut_ptr(), fault.as_mut_ptr(),
                )
            });
            assert!(fault.iter().all(|&f| f == 0), "y = r + e * secret does not fit its row");
            unpack(&x, nl).into_iter().zip(unpack(&y, yl)).map(|(x, y)| CompositeDLogProof { x, y }).collect()
        })
    }

    /// `Verdict::Panic` where one of the reference's three asserts fires for that statement.
    pub fn verify_batch(proofs: &[&CompositeDLogProof], statement: &[DLogStatement]) -> Vec<Verdict> {
        assert_eq!(proofs.len(), statement.len());
        let b = proofs.len();
        if b == 0 {
            return Vec::new();
        }
        let nl = limbs_for_bits(statement.iter().map(|s| s.N.bit_length().max(s.g.bit_length()).max(s.ni.bit_length())).max().unwrap());
        let yl = limbs_for_bits(proofs.iter().map(|p| p.y.bit_length()).max().unwrap());
        let zero = BigInt::zero();
        // x is compared with a canonical residue mod N: wider than the row, it never matches
        let wide: Vec<bool> = proofs.iter().map(|p| !fits(&p.x, nl)).collect();
        Engine::with(|eng| {
            let (mut accept, mut fault) = (vec![0u8; b], vec![0u8; b]);
            eng.check(unsafe {
                ffi::zkp_dlog_verify(
                    eng.h, b as i32, nl as i32, pack(statement.iter().map(|s| &s.N), nl).as_ptr(), pack(statement.iter().map(|s| &s.g), nl).as_ptr(),
                    pack(statement.iter().map(|s| &s.ni), nl).as_ptr(), pack(proofs.iter().zip(&wide).map(|(p, &w)| if w { &zero } else { &p.x }), nl).as_ptr(),
                    pack(proofs.iter().map(|p| &p.y), yl).as_ptr(), yl as i32, accept.as_mut_ptr(), fault.as_mut_ptr(),
                )
            });
            (0..b).map(|k| if fault[k] != 0 { Verdict::Panic } else if wide[k] { Verdict::Reject } else { Verdict::from_flags(accept[k], 0) }).collect()
        })
    }
}

/// wi_dlog_proof.rs:94-107 (host-side helper of the reference's tests; not on the device path): a^((p-1)/2) mod p with the
/// exponent taken as (p - 1) * 2^-1 mod p, and every outcome but 1 - a multiple of p included - reported as -1, as there.
pub fn legendre_symbol(a: &BigInt, p: &BigInt) -> i32 {
    let half = BigInt::mod_mul(&(p - BigInt::one()), &BigInt::mod_inv(&BigInt::from(2), p).unwrap(), p);
    if BigInt::mod_pow(a, &half, p) == BigInt::one() {
        1
    } else {
        -1
    }
}
