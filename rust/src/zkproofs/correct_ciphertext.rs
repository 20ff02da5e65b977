//! CiphertextProof: knowledge of the plaintext of c (reference src/zkproofs/correct_ciphertext.rs:22-98) over
//! zkp_ciphertext_prove / zkp_ciphertext_verify.
use curv::arithmetic::traits::*;
use curv::BigInt;
use paillier::EncryptionKey;
use serde::{Deserialize, Serialize};

use super::errors::IncorrectProof;
use crate::engine::{fits, group_by_key, pack, require_one_key, unpack, Engine, Verdict};
use crate::ffi;

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct CiphertextProof {
    pub z1: BigInt,
    pub z2: BigInt,
    pub c_prime: BigInt,
}

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct CiphertextWitness {
    pub x: BigInt,
    pub r: BigInt,
}

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct CiphertextStatement {
    pub ek: EncryptionKey,
    pub c: BigInt,
}

impl CiphertextProof {
    /// correct_ciphertext.rs:42-64
    pub fn prove(witness: &CiphertextWitness, statement: &CiphertextStatement) -> Self {
        Self::prove_batch(std::slice::from_ref(witness), std::slice::from_ref(statement)).pop().unwrap()
    }
    /// correct_ciphertext.rs:66-98
    pub fn verify(&self, statement: &CiphertextStatement) -> Result<(), IncorrectProof> {
        Self::verify_batch(&[self], std::slice::from_ref(statement))[0].into_result("unreachable")
    }

    pub fn prove_batch(witness: &[CiphertextWitness], statement: &[CiphertextStatement]) -> Vec<CiphertextProof> {
        assert_eq!(witness.len(), statement.len());
        if statement.is_empty() {
            return Vec::new();
        }
        require_one_key(statement.iter().map(|s| &s.ek), "CiphertextProof::prove_batch");
        // x' then r' per statement, in the reference's draw order (:43-44)
        let (mut x_prime, mut r_prime) = (Vec::new(), Vec::new());
        for s in statement {
            x_prime.push(BigInt::sample_below(&s.ek.n));
            r_prime.push(BigInt::sample_below(&s.ek.n));
        }
        Engine::with(|eng| {
            eng.use_key(&statement[0].ek);
            let (nl, nnl, zl, b) = (eng.nl(), eng.nnl(), eng.zl(), statement.len());
            let (mut z1, mut z2, mut cp) = (vec![0u32; b * zl], vec![0u32; b * nnl], vec![0u32; b * nnl]);
            eng.check(unsafe {
                ffi::zkp_ciphertext_prove(
                    eng.h, b as i32, zl as i32, pack(witness.iter().map(|w| &w.x), nl).as_ptr(), pack(witness.iter().map(|w| &w.r), nl).as_ptr(),
                    pack(statement.iter().map(|s| &s.c), nnl).as_ptr(), pack(x_prime.iter(), nl).as_ptr(), pack(r_prime.iter(), nl).as_ptr(),
                    z1.as_mut_ptr(), z2.as_mut_ptr(), cp.as_mut_ptr(),
                )
            });
            let (z1, z2, cp) = (unpack(&z1, zl), unpack(&z2, nnl), unpack(&cp, nnl));
            (0..b).map(|i| CiphertextProof { z1: z1[i].clone(), z2: z2[i].clone(), c_prime: cp[i].clone() }).collect()
        })
    }

    /// `c` and `c_prime` are hashed as given (wider than n^2: that proof is rejected); `z1` only matters mod n and
    /// `z2` mod n^2 (Enc(z1, z2), :73-79).
    pub fn verify_batch(proofs: &[&CiphertextProof], statement: &[CiphertextStatement]) -> Vec<Verdict> {
        assert_eq!(proofs.len(), statement.len());
        let mut out = vec![Verdict::Reject; proofs.len()];
        for (ek, idx) in group_by_key(statement.iter().map(|s| &s.ek)) {
            Engine::with(|eng| {
                eng.use_key(&ek);
                let (nnl, zl) = (eng.nnl(), eng.zl());
                let zero = BigInt::zero();
                let ok: Vec<bool> = idx.iter().map(|&i| fits(&statement[i].c, nnl) && fits(&proofs[i].c_prime, nnl)).collect();
                let c = pack(idx.iter().zip(&ok).map(|(&i, &k)| if k { &statement[i].c } else { &zero }), nnl);
                let cp = pack(idx.iter().zip(&ok).map(|(&i, &k)| if k { &proofs[i].c_prime } else { &zero }), nnl);
                let z1: Vec<BigInt> = idx.iter().map(|&i| if fits(&proofs[i].z1, zl) { proofs[i].z1.clone() } else { &proofs[i].z1 % &ek.n }).collect();
                let z2: Vec<BigInt> = idx.iter().map(|&i| &proofs[i].z2 % &ek.nn).collect();
                let mut accept = vec![0u8; idx.len()];
                eng.check(unsafe {
                    ffi::zkp_ciphertext_verify(eng.h, idx.len() as i32, zl as i32, c.as_ptr(), pack(z1.iter(), zl).as_ptr(), pack(z2.iter(), nnl).as_ptr(),
                                               cp.as_ptr(), accept.as_mut_ptr())
                });
                for (k, &i) in idx.iter().enumerate() {
                    out[i] = if ok[k] { Verdict::from_flags(accept[k], 0) } else { Verdict::Reject };
                }
            });
        }
        out
    }
}
