//! ZeroProof: c = r^n mod n^2 encrypts zero (reference src/zkproofs/zero_enc_proof.rs:26-94) over
//! zkp_zero_prove / zkp_zero_verify.
use curv::arithmetic::traits::*;
use curv::BigInt;
use paillier::EncryptionKey;
use serde::{Deserialize, Serialize};

use super::errors::IncorrectProof;
use crate::engine::{fits, group_by_key, pack, require_one_key, unpack, Engine, Verdict};
use crate::ffi;

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct ZeroProof {
    pub z: BigInt,
    pub a: BigInt,
}

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct ZeroWitness {
    pub r: BigInt,
}

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct ZeroStatement {
    pub ek: EncryptionKey,
    pub c: BigInt,
}

impl ZeroProof {
    /// zero_enc_proof.rs:44-64
    pub fn prove(witness: &ZeroWitness, statement: &ZeroStatement) -> Self {
        Self::prove_batch(std::slice::from_ref(witness), std::slice::from_ref(statement)).pop().unwrap()
    }
    /// zero_enc_proof.rs:66-94
    pub fn verify(&self, statement: &ZeroStatement) -> Result<(), IncorrectProof> {
        Self::verify_batch(&[self], std::slice::from_ref(statement))[0].into_result("unreachable")
    }

    /// Many statements under ONE key.  r' is drawn per statement as the reference does (:45); everything else is one
    /// device call.
    pub fn prove_batch(witness: &[ZeroWitness], statement: &[ZeroStatement]) -> Vec<ZeroProof> {
        assert_eq!(witness.len(), statement.len());
        if statement.is_empty() {
            return Vec::new();
        }
        require_one_key(statement.iter().map(|s| &s.ek), "ZeroProof::prove_batch");
        let r_prime: Vec<BigInt> = statement.iter().map(|s| BigInt::sample_below(&s.ek.n)).collect();
        Engine::with(|eng| {
            eng.use_key(&statement[0].ek);
            let (nl, nnl, b) = (eng.nl(), eng.nnl(), statement.len());
            let (mut z, mut a) = (vec![0u32; b * nnl], vec![0u32; b * nnl]);
            eng.check(unsafe {
                ffi::zkp_zero_prove(
                    eng.h, b as i32, pack(witness.iter().map(|w| &w.r), nl).as_ptr(), pack(statement.iter().map(|s| &s.c), nnl).as_ptr(),
                    pack(r_prime.iter(), nl).as_ptr(), z.as_mut_ptr(), a.as_mut_ptr(),
                )
            });
            unpack(&z, nnl).into_iter().zip(unpack(&a, nnl)).map(|(z, a)| ZeroProof { z, a }).collect()
        })
    }

    /// One verdict per proof; statements may be under different keys (grouped, one device call per key).  `c` and `a`
    /// enter the transcript hash as given, so a value wider than n^2 rejects that proof; `z` is only ever used reduced.
    pub fn verify_batch(proofs: &[&ZeroProof], statement: &[ZeroStatement]) -> Vec<Verdict> {
        assert_eq!(proofs.len(), statement.len());
        let mut out = vec![Verdict::Reject; proofs.len()];
        for (ek, idx) in group_by_key(statement.iter().map(|s| &s.ek)) {
            Engine::with(|eng| {
                eng.use_key(&ek);
                let nnl = eng.nnl();
                let zero = BigInt::zero();
                let ok: Vec<bool> = idx.iter().map(|&i| fits(&statement[i].c, nnl) && fits(&proofs[i].a, nnl)).collect();
                let c = pack(idx.iter().zip(&ok).map(|(&i, &k)| if k { &statement[i].c } else { &zero }), nnl);
                let a = pack(idx.iter().zip(&ok).map(|(&i, &k)| if k { &proofs[i].a } else { &zero }), nnl);
                let z: Vec<BigInt> = idx.iter().map(|&i| &proofs[i].z % &ek.nn).collect();
                let mut accept = vec![0u8; idx.len()];
                eng.check(unsafe { ffi::zkp_zero_verify(eng.h, idx.len() as i32, c.as_ptr(), pack(z.iter(), nnl).as_ptr(), a.as_ptr(), accept.as_mut_ptr()) });
                for (k, &i) in idx.iter().enumerate() {
                    out[i] = if ok[k] { Verdict::from_flags(accept[k], 0) } else { Verdict::Reject };
                }
            });
        }
        out
    }
}
