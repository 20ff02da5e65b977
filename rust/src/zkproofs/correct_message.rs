//! CorrectMessageProof (reference src/zkproofs/correct_message.rs:19-163): ring proof that a ciphertext encrypts one
//! of a few valid messages, over zkp_correct_message_prove / zkp_correct_message_verify.
use curv::arithmetic::traits::*;
use curv::BigInt;
use paillier::EncryptionKey;

use super::errors::IncorrectProof;
use crate::engine::{fits, limbs_for_bits, pack, unpack, Engine, Verdict};
use crate::ffi;

const B: usize = 256; // correct_message.rs:19
const E_LIMBS: usize = B / 32;

/// correct_message.rs:25-32 (fields private and not `Serialize`, as there)
pub struct CorrectMessageProof {
    e_vec: Vec<BigInt>,
    z_vec: Vec<BigInt>,
    a_vec: Vec<BigInt>,
    ciphertext: BigInt,
    valid_messages: Vec<BigInt>,
    ek: EncryptionKey,
}

const CHAL_PANIC: &str = "assertion failed: `(left == right)` (chal, ei_sum)"; // correct_message.rs:133

impl CorrectMessageProof {
    /// correct_message.rs:35-128
    pub fn prove(ek: &EncryptionKey, valid_messages: &[BigInt], message_to_encrypt: &BigInt) -> CorrectMessageProof {
        Self::prove_batch(ek, valid_messages, std::slice::from_ref(message_to_encrypt)).pop().unwrap()
    }
    /// correct_message.rs:129-162
    pub fn verify(&self) -> Result<(), IncorrectProof> {
        Self::verify_batch(&[self])[0].into_result(CHAL_PANIC)
    }

    /// One proof per message, all over the same key and message list.
    pub fn prove_batch(ek: &EncryptionKey, valid_messages: &[BigInt], messages: &[BigInt]) -> Vec<CorrectMessageProof> {
        let (b, m) = (messages.len(), valid_messages.len());
        if b == 0 {
            return Vec::new();
        }
        assert!(m >= 1, "attempt to subtract with overflow"); // 0..num_of_message - 1 (:59)
        // r, e_i (M - 1 samples of B bits), z_i (M - 1 below n), w - per proof, in the reference's order (:41,59-66)
        let (mut r, mut e_rand, mut z_rand, mut w) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for _ in 0..b {
            r.push(BigInt::sample_below(&ek.n));
            for _ in 0..m - 1 {
                e_rand.push(BigInt::sample(B));
            }
            for _ in 0..m - 1 {
                z_rand.push(BigInt::sample_below(&ek.n));
            }
            w.push(BigInt::sample_below(&ek.n));
        }
        Engine::with(|eng| {
            eng.use_key(ek);
            let (nl, nnl) = (eng.nl(), eng.nnl());
            let ml = nl; // message rows: reduced mod n (only m mod n enters (m n + 1) % nn)
            let reduced: Vec<BigInt> = valid_messages.iter().map(|v| v % &ek.n).collect();
            let valid: Vec<BigInt> = (0..b).flat_map(|_| reduced.iter().cloned()).collect();
            let msgs: Vec<BigInt> = messages.iter().map(|v| v % &ek.n).collect();
            let (mut c, mut e, mut z, mut a, mut fault) = (vec![0u32; b * nnl], vec![0u32; b * m * E_LIMBS], vec![0u32; b * m * nl], vec![0u32; b * m * nnl], vec![0u8; b]);
            eng.check(unsafe {
                ffi::zkp_correct_message_prove(
                    eng.h, b as i32, m as i32, ml as i32, pack(valid.iter(), ml).as_ptr(), pack(msgs.iter(), ml).as_ptr(), pack(r.iter(), nl).as_ptr(),
                    pack(e_rand.iter(), E_LIMBS).as_ptr(), pack(z_rand.iter(), nl).as_ptr(), pack(w.iter(), nl).as_ptr(),
                    c.as_mut_ptr(), e.as_mut_ptr(), z.as_mut_ptr(), a.as_mut_ptr(), fault.as_mut_ptr(),
                )
            });
            // the message is not among the valid ones (index past the random vectors) or a non-invertible value under unwrap()
            assert!(fault.iter().all(|&f| f == 0), "index out of bounds / called `Option::unwrap()` on a `None` value");
            let (c, e, z, a) = (unpack(&c, nnl), unpack(&e, E_LIMBS), unpack(&z, nl), unpack(&a, nnl));
            (0..b)
                .map(|k| CorrectMessageProof {
                    e_vec: e[k * m..(k + 1) * m].to_vec(),
                    z_vec: z[k * m..(k + 1) * m].to_vec(),
                    a_vec: a[k * m..(k + 1) * m].to_vec(),
                    ciphertext: c[k].clone(),
                    valid_messages: valid_messages.to_vec(),
                    ek: ek.clone(),
                })
                .collect()
        })
    }

    /// Proofs of one batch share the key and the number of valid messages (checked).  `Verdict::Panic` where
    /// `assert_eq!(chal, ei_sum)` fires (:133) or a proof vector is shorter than the message list (index out of range).
    /// The reference hashes ALL of a_vec and folds ALL of e_vec: vectors longer than the message list are refused
    /// here (usage error) rather than truncated.
    pub fn verify_batch(proofs: &[&CorrectMessageProof]) -> Vec<Verdict> {
        let b = proofs.len();
        if b == 0 {
            return Vec::new();
        }
        let (ek, m) = (&proofs[0].ek, proofs[0].valid_messages.len());
        assert!(proofs.iter().all(|p| p.ek.n == ek.n && p.valid_messages.len() == m), "verify_batch: proofs must share the key and the number of messages");
        assert!(proofs.iter().all(|p| p.e_vec.len() <= m && p.a_vec.len() <= m), "e_vec / a_vec longer than the message list are not supported");
        let short: Vec<bool> = proofs.iter().map(|p| p.e_vec.len() < m || p.z_vec.len() < m || p.a_vec.len() < m).collect();
        let el = limbs_for_bits(proofs.iter().flat_map(|p| p.e_vec.iter().map(|e| e.bit_length())).max().unwrap_or(B).max(B));
        Engine::with(|eng| {
            eng.use_key(ek);
            let (nl, nnl) = (eng.nl(), eng.nnl());
            let zero = BigInt::zero();
            // a_vec is hashed as given: a row wider than n^2 cannot be laid out and rejects that proof
            let wide: Vec<bool> = proofs.iter().map(|p| p.a_vec.iter().any(|a| !fits(a, nnl))).collect();
            let skip = |k: usize| short[k] || wide[k];
            let row = |k: usize, v: &Vec<BigInt>, i: usize| -> BigInt { if skip(k) { zero.clone() } else { v[i].clone() } };
            let (mut c, mut valid, mut e, mut z, mut a) = (Vec::new(), Vec::new(), Vec::new(), Vec::new(), Vec::new());
            for (k, p) in proofs.iter().enumerate() {
                c.push(&p.ciphertext % &ek.nn);
                for i in 0..m {
                    valid.push(&p.valid_messages[i] % &ek.n);
                    e.push(row(k, &p.e_vec, i));
                    z.push(&row(k, &p.z_vec, i) % &ek.n); // z^n mod nn depends on z mod n only
                    a.push(row(k, &p.a_vec, i));
                }
            }
            let (mut accept, mut fault) = (vec![0u8; b], vec![0u8; b]);
            eng.check(unsafe {
                ffi::zkp_correct_message_verify(
                    eng.h, b as i32, m as i32, nl as i32, el as i32, pack(c.iter(), nnl).as_ptr(), pack(valid.iter(), nl).as_ptr(), pack(e.iter(), el).as_ptr(),
                    pack(z.iter(), nl).as_ptr(), pack(a.iter(), nnl).as_ptr(), accept.as_mut_ptr(), fault.as_mut_ptr(),
                )
            });
            (0..b).map(|k| if short[k] { Verdict::Panic } else if wide[k] { Verdict::Reject } else { Verdict::from_flags(accept[k], fault[k]) }).collect()
        })
    }
}
