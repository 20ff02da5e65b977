//! RangeProofNi (reference src/zkproofs/range_proof_ni.rs:23-129 over range_proof.rs:128-355) through
//! zkp_rangeproof_ni_prove / zkp_rangeproof_ni_verify: the 256 + 128..256 Paillier encryptions of one proof, the
//! Fiat-Shamir hash between them and the accept predicates all run on the device.
use curv::arithmetic::traits::*;
use curv::BigInt;
use paillier::EncryptionKey;
use rand::RngCore;
use serde::{Deserialize, Serialize};

use super::errors::IncorrectProof;
use crate::engine::{fits, from_limbs, limbs_for_bits, pack, to_limbs, unpack, Engine, Verdict};
use crate::ffi;

const SECURITY_PARAMETER: usize = 128; // range_proof_ni.rs:23
const MAX_RANGE_BITS: usize = 64 * 32 - 1; // device rows of range / w1 / w2 / masked_x hold at most 64 limbs

/// range_proof.rs:32-39
#[derive(Default, Debug, Serialize, Deserialize, Clone)]
pub struct EncryptedPairs {
    #[serde(with = "crate::serialize::vecbigint")]
    pub c1: Vec<BigInt>,
    #[serde(with = "crate::serialize::vecbigint")]
    pub c2: Vec<BigInt>,
}

/// range_proof.rs:53-78
#[derive(Debug, Serialize, Deserialize, Clone)]
pub enum Response {
    Open {
        #[serde(with = "crate::serialize::bigint")]
        w1: BigInt,
        #[serde(with = "crate::serialize::bigint")]
        r1: BigInt,
        #[serde(with = "crate::serialize::bigint")]
        w2: BigInt,
        #[serde(with = "crate::serialize::bigint")]
        r2: BigInt,
    },
    Mask {
        j: u8,
        #[serde(with = "crate::serialize::bigint")]
        masked_x: BigInt,
        #[serde(with = "crate::serialize::bigint")]
        masked_r: BigInt,
    },
}

/// range_proof.rs:80-81
#[derive(Debug, Serialize, Deserialize, Clone)]
pub struct Proof(Vec<Response>);

/// range_proof_ni.rs:36-44 (all fields private, as there)
#[derive(Debug, Serialize, Deserialize, Clone)]
pub struct RangeProofNi {
    ek: EncryptionKey,
    range: BigInt,
    ciphertext: BigInt,
    encrypted_pairs: EncryptedPairs,
    proof: Proof,
    error_factor: usize,
}

/// One statement of a proving batch: the arguments of `RangeProofNi::prove` after the key.
pub struct RangeStatement<'a> {
    pub range: &'a BigInt,
    pub ciphertext: &'a BigInt,
    pub secret_x: &'a BigInt,
    pub secret_r: &'a BigInt,
}

impl RangeProofNi {
    /// range_proof_ni.rs:47-82
    pub fn prove(ek: &EncryptionKey, range: &BigInt, ciphertext: &BigInt, secret_x: &BigInt, secret_r: &BigInt) -> RangeProofNi {
        Self::prove_batch(ek, &[RangeStatement { range, ciphertext, secret_x, secret_r }]).pop().unwrap()
    }

    /// Many statements under one key in ONE device call (2 * 128 * batch encryptions in one launch).  The randomness
    /// is drawn here, per statement, in the order generate_encrypted_pairs draws it (range_proof.rs:136-159): w1[128]
    /// in [q/3, 2q/3), the 128 coins, r1[128], r2[128] below n.
    pub fn prove_batch(ek: &EncryptionKey, statements: &[RangeStatement]) -> Vec<RangeProofNi> {
        let (b, ef) = (statements.len(), SECURITY_PARAMETER);
        if b == 0 {
            return Vec::new();
        }
        let wl = limbs_for_bits(statements.iter().map(|s| std::cmp::max(s.range.bit_length(), s.secret_x.bit_length()) + 2).max().unwrap());
        assert!(wl <= 64, "range / secret_x wider than 2047 bits");
        let (mut w1, mut swap, mut r1, mut r2) = (Vec::new(), vec![0u8; b * ef], Vec::new(), Vec::new());
        let mut rng = rand::thread_rng();
        for (k, s) in statements.iter().enumerate() {
            let third = s.range.div_floor(&BigInt::from(3));
            let two_thirds = BigInt::from(2) * &third;
            for _ in 0..ef {
                w1.push(BigInt::sample_range(&third, &two_thirds));
            }
            rng.fill_bytes(&mut swap[k * ef..(k + 1) * ef]);
            for _ in 0..ef {
                r1.push(BigInt::sample_below(&ek.n));
            }
            for _ in 0..ef {
                r2.push(BigInt::sample_below(&ek.n));
            }
        }
        for s in swap.iter_mut() {
            *s &= 1; // rand::random::<bool>() (range_proof.rs:146)
        }
        Engine::with(|eng| {
            eng.use_key(ek);
            let (nl, nnl) = (eng.nl(), eng.nnl());
            let (mut c1, mut c2) = (vec![0u32; b * ef * nnl], vec![0u32; b * ef * nnl]);
            let (mut kind, mut digest) = (vec![0u8; b * ef], vec![0u8; b * 32]);
            let (mut resp_w, mut resp_r) = (vec![0u32; b * ef * 2 * wl], vec![0u32; b * ef * 2 * nl]);
            eng.check(unsafe {
                ffi::zkp_rangeproof_ni_prove(
                    eng.h, b as i32, ef as i32, wl as i32,
                    pack(statements.iter().map(|s| s.range), wl).as_ptr(), pack(statements.iter().map(|s| s.secret_x), wl).as_ptr(),
                    pack(statements.iter().map(|s| s.secret_r), nl).as_ptr(),
                    pack(w1.iter(), wl).as_ptr(), swap.as_ptr(), pack(r1.iter(), nl).as_ptr(), pack(r2.iter(), nl).as_ptr(),
                    c1.as_mut_ptr(), c2.as_mut_ptr(), digest.as_mut_ptr(), kind.as_mut_ptr(), resp_w.as_mut_ptr(), resp_r.as_mut_ptr(),
                )
            });
            (0..b)
                .map(|k| {
                    let responses = (0..ef)
                        .map(|i| {
                            let t = k * ef + i;
                            let w = &resp_w[t * 2 * wl..(t + 1) * 2 * wl];
                            let r = &resp_r[t * 2 * nl..(t + 1) * 2 * nl];
                            if kind[t] as i32 == ffi::ZKP_RP_OPEN {
                                Response::Open { w1: from_limbs(&w[..wl]), r1: from_limbs(&r[..nl]), w2: from_limbs(&w[wl..]), r2: from_limbs(&r[nl..]) }
                            } else {
                                Response::Mask { j: kind[t], masked_x: from_limbs(&w[..wl]), masked_r: from_limbs(&r[..nl]) }
                            }
                        })
                        .collect();
                    RangeProofNi {
                        ek: ek.clone(),
                        range: statements[k].range.clone(),
                        ciphertext: statements[k].ciphertext.clone(),
                        encrypted_pairs: EncryptedPairs {
                            c1: unpack(&c1[k * ef * nnl..(k + 1) * ef * nnl], nnl),
                            c2: unpack(&c2[k * ef * nnl..(k + 1) * ef * nnl], nnl),
                        },
                        proof: Proof(responses),
                        error_factor: ef,
                    }
                })
                .collect()
        })
    }

    /// range_proof_ni.rs:84-107.  Precondition failures panic exactly where the reference does (:86, :88).
    pub fn verify(&self, ek: &EncryptionKey, ciphertext: &BigInt) -> Result<(), IncorrectProof> {
        assert_eq!(ek, &self.ek);
        assert_eq!(ciphertext, &self.ciphertext);
        self.verify_self()
    }

    /// range_proof_ni.rs:109-128
    pub fn verify_self(&self) -> Result<(), IncorrectProof> {
        Self::verify_batch(&[self])[0].into_result("index out of bounds")
    }

    /// Proofs under one key and error factor (checked) in one device call; one verdict per proof.  A response value
    /// wider than its proof's range can never pass the interval predicates (range_proof.rs:301-307,341) and a
    /// ciphertext wider than n^2 is never equal to a canonical one: such a proof is rejected on its own.  Randomness
    /// enters only as r^n mod n^2 and is reduced mod n, as the reference's mod_pow does.
    pub fn verify_batch(proofs: &[&RangeProofNi]) -> Vec<Verdict> {
        let b = proofs.len();
        if b == 0 {
            return Vec::new();
        }
        let (ek, ef) = (&proofs[0].ek, proofs[0].error_factor);
        assert!(proofs.iter().all(|p| p.ek.n == ek.n && p.error_factor == ef), "verify_batch: proofs must share key and error factor");
        if ef == 0 {
            return vec![Verdict::Accept; b]; // (0..0).all(..): the reference returns Ok(())
        }
        let mut out = vec![Verdict::Reject; b];
        // responses[i] / bits_of_e[i] index out of range in the reference (range_proof.rs:273-274)
        for (k, p) in proofs.iter().enumerate() {
            if p.proof.0.len() < ef || p.encrypted_pairs.c1.len() < ef || p.encrypted_pairs.c2.len() < ef || ef > 256 {
                out[k] = Verdict::Panic;
            }
        }
        let wl = limbs_for_bits(proofs.iter().map(|p| p.range.bit_length()).filter(|&bits| bits <= MAX_RANGE_BITS).max().unwrap_or(1) + 1);
        Engine::with(|eng| {
            eng.use_key(ek);
            let (nl, nnl) = (eng.nl(), eng.nnl());
            let (mut range, mut cx) = (vec![0u32; b * wl], vec![0u32; b * nnl]);
            let (mut c1, mut c2, mut kind) = (vec![0u32; b * ef * nnl], vec![0u32; b * ef * nnl], vec![0u8; b * ef]);
            let (mut resp_w, mut resp_r) = (vec![0u32; b * ef * 2 * wl], vec![0u32; b * ef * 2 * nl]);
            let mut laid_out = vec![false; b];
            let red = |r: &BigInt| if fits(r, nl) { r.clone() } else { r % &ek.n };
            for (k, p) in proofs.iter().enumerate() {
                if out[k] == Verdict::Panic {
                    continue;
                }
                let representable = p.range.bit_length() <= MAX_RANGE_BITS
                    && fits(&p.ciphertext, nnl)
                    && (0..ef).all(|i| {
                        fits(&p.encrypted_pairs.c1[i], nnl)
                            && fits(&p.encrypted_pairs.c2[i], nnl)
                            && match &p.proof.0[i] {
                                Response::Open { w1, w2, .. } => fits(w1, wl) && fits(w2, wl),
                                Response::Mask { masked_x, .. } => fits(masked_x, wl),
                            }
                    });
                if !representable {
                    continue; // stays Reject; its rows stay zero
                }
                laid_out[k] = true;
                range[k * wl..(k + 1) * wl].copy_from_slice(&to_limbs(&p.range, wl));
                cx[k * nnl..(k + 1) * nnl].copy_from_slice(&to_limbs(&p.ciphertext, nnl));
                for i in 0..ef {
                    let t = k * ef + i;
                    c1[t * nnl..(t + 1) * nnl].copy_from_slice(&to_limbs(&p.encrypted_pairs.c1[i], nnl));
                    c2[t * nnl..(t + 1) * nnl].copy_from_slice(&to_limbs(&p.encrypted_pairs.c2[i], nnl));
                    let (w, r) = (&mut resp_w[t * 2 * wl..(t + 1) * 2 * wl], &mut resp_r[t * 2 * nl..(t + 1) * 2 * nl]);
                    match &p.proof.0[i] {
                        Response::Open { w1, r1, w2, r2 } => {
                            kind[t] = ffi::ZKP_RP_OPEN as u8;
                            w[..wl].copy_from_slice(&to_limbs(w1, wl));
                            w[wl..].copy_from_slice(&to_limbs(w2, wl));
                            r[..nl].copy_from_slice(&to_limbs(&red(r1), nl));
                            r[nl..].copy_from_slice(&to_limbs(&red(r2), nl));
                        }
                        Response::Mask { j, masked_x, masked_r } => {
                            // `if *j == 1 { c1 } else { c2 }` (range_proof.rs:321-325)
                            kind[t] = if *j == 1 { ffi::ZKP_RP_MASK1 as u8 } else { ffi::ZKP_RP_MASK2 as u8 };
                            w[..wl].copy_from_slice(&to_limbs(masked_x, wl));
                            r[..nl].copy_from_slice(&to_limbs(&red(masked_r), nl));
                        }
                    }
                }
            }
            let (mut accept, mut fault) = (vec![0u8; b], vec![0u8; b]);
            eng.check(unsafe {
                ffi::zkp_rangeproof_ni_verify(
                    eng.h, b as i32, ef as i32, wl as i32, range.as_ptr(), cx.as_ptr(), c1.as_ptr(), c2.as_ptr(), kind.as_ptr(),
                    resp_w.as_ptr(), resp_r.as_ptr(), accept.as_mut_ptr(), fault.as_mut_ptr(), std::ptr::null_mut(),
                )
            });
            for k in 0..b {
                if laid_out[k] {
                    out[k] = Verdict::from_flags(accept[k], fault[k]);
                }
            }
        });
        out
    }
}
