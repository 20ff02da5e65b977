//! MulProof: c = a b mod n under encryption (reference src/zkproofs/multiplication_proof.rs:32-154) over
//! zkp_mul_prove / zkp_mul_verify.
use curv::arithmetic::traits::*;
use curv::BigInt;
use paillier::EncryptionKey;
use serde::{Deserialize, Serialize};

use super::errors::IncorrectProof;
use crate::engine::{fits, group_by_key, pack, require_one_key, unpack, Engine, Verdict};
use crate::ffi;

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct MulProof {
    pub f: BigInt,
    pub z1: BigInt,
    pub z2: BigInt,
    pub e_d: BigInt,
    pub e_db: BigInt,
}

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct MulWitness {
    pub a: BigInt,
    pub b: BigInt,
    pub c: BigInt,
    pub r_a: BigInt,
    pub r_b: BigInt,
    pub r_c: BigInt,
}

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct MulStatement {
    pub ek: EncryptionKey,
    pub e_a: BigInt,
    pub e_b: BigInt,
    pub e_c: BigInt,
}

const INV_PANIC: &str = "called `Option::unwrap()` on a `None` value"; // mod_inv(..).unwrap(), multiplication_proof.rs:96,137

impl MulProof {
    /// multiplication_proof.rs:60-106
    pub fn prove(witness: &MulWitness, statement: &MulStatement) -> Self {
        Self::prove_batch(std::slice::from_ref(witness), std::slice::from_ref(statement)).pop().unwrap()
    }
    /// multiplication_proof.rs:108-145
    pub fn verify(&self, statement: &MulStatement) -> Result<(), IncorrectProof> {
        Self::verify_batch(&[self], std::slice::from_ref(statement))[0].into_result(INV_PANIC)
    }

    pub fn prove_batch(witness: &[MulWitness], statement: &[MulStatement]) -> Vec<MulProof> {
        assert_eq!(witness.len(), statement.len());
        if statement.is_empty() {
            return Vec::new();
        }
        require_one_key(statement.iter().map(|s| &s.ek), "MulProof::prove_batch");
        // d, then r_d coprime to n, per statement (:61-62)
        let (mut d, mut r_d) = (Vec::new(), Vec::new());
        for s in statement {
            d.push(BigInt::sample_below(&s.ek.n));
            r_d.push(sample_paillier_random(&s.ek.n));
        }
        Engine::with(|eng| {
            eng.use_key(&statement[0].ek);
            let (nl, nnl, b) = (eng.nl(), eng.nnl(), statement.len());
            let (mut f, mut z1, mut z2) = (vec![0u32; b * nl], vec![0u32; b * nnl], vec![0u32; b * nnl]);
            let (mut e_d, mut e_db, mut fault) = (vec![0u32; b * nnl], vec![0u32; b * nnl], vec![0u8; b]);
            eng.check(unsafe {
                ffi::zkp_mul_prove(
                    eng.h, b as i32,
                    pack(witness.iter().map(|w| &w.a), nl).as_ptr(), pack(witness.iter().map(|w| &w.b), nl).as_ptr(),
                    pack(witness.iter().map(|w| &w.r_a), nl).as_ptr(), pack(witness.iter().map(|w| &w.r_b), nl).as_ptr(),
                    pack(witness.iter().map(|w| &w.r_c), nl).as_ptr(),
                    pack(statement.iter().map(|s| &s.e_a), nnl).as_ptr(), pack(statement.iter().map(|s| &s.e_b), nnl).as_ptr(),
                    pack(statement.iter().map(|s| &s.e_c), nnl).as_ptr(), pack(d.iter(), nl).as_ptr(), pack(r_d.iter(), nl).as_ptr(),
                    f.as_mut_ptr(), z1.as_mut_ptr(), z2.as_mut_ptr(), e_d.as_mut_ptr(), e_db.as_mut_ptr(), fault.as_mut_ptr(),
                )
            });
            assert!(fault.iter().all(|&x| x == 0), "{}", INV_PANIC);
            let (f, z1, z2, e_d, e_db) = (unpack(&f, nl), unpack(&z1, nnl), unpack(&z2, nnl), unpack(&e_d, nnl), unpack(&e_db, nnl));
            (0..b).map(|i| MulProof { f: f[i].clone(), z1: z1[i].clone(), z2: z2[i].clone(), e_d: e_d[i].clone(), e_db: e_db[i].clone() }).collect()
        })
    }

    /// `Verdict::Panic` where `mod_inv(..).unwrap()` fails for that proof (:137).  The five ciphertexts are hashed as
    /// given and `f` is an exponent (:138): wider than their rows, that proof is rejected; z1 / z2 are randomness.
    pub fn verify_batch(proofs: &[&MulProof], statement: &[MulStatement]) -> Vec<Verdict> {
        assert_eq!(proofs.len(), statement.len());
        let mut out = vec![Verdict::Reject; proofs.len()];
        for (ek, idx) in group_by_key(statement.iter().map(|s| &s.ek)) {
            Engine::with(|eng| {
                eng.use_key(&ek);
                let (nl, nnl) = (eng.nl(), eng.nnl());
                let zero = BigInt::zero();
                let ok: Vec<bool> = idx
                    .iter()
                    .map(|&i| {
                        let (s, p) = (&statement[i], proofs[i]);
                        fits(&s.e_a, nnl) && fits(&s.e_b, nnl) && fits(&s.e_c, nnl) && fits(&p.e_d, nnl) && fits(&p.e_db, nnl) && fits(&p.f, nl)
                    })
                    .collect();
                // rows of the proofs that can be laid out; zero rows for the others (their verdict is already Reject)
                let pick = |rows: Vec<&BigInt>, limbs: usize| -> Vec<u32> {
                    pack(rows.into_iter().zip(&ok).map(|(x, &k)| if k { x } else { &zero }), limbs)
                };
                let z1: Vec<BigInt> = idx.iter().map(|&i| &proofs[i].z1 % &ek.nn).collect();
                let z2: Vec<BigInt> = idx.iter().map(|&i| &proofs[i].z2 % &ek.nn).collect();
                let (mut accept, mut fault) = (vec![0u8; idx.len()], vec![0u8; idx.len()]);
                eng.check(unsafe {
                    ffi::zkp_mul_verify(
                        eng.h, idx.len() as i32,
                        pick(idx.iter().map(|&i| &statement[i].e_a).collect(), nnl).as_ptr(), pick(idx.iter().map(|&i| &statement[i].e_b).collect(), nnl).as_ptr(), pick(idx.iter().map(|&i| &statement[i].e_c).collect(), nnl).as_ptr(),
                        pick(idx.iter().map(|&i| &proofs[i].f).collect(), nl).as_ptr(), pack(z1.iter(), nnl).as_ptr(), pack(z2.iter(), nnl).as_ptr(),
                        pick(idx.iter().map(|&i| &proofs[i].e_d).collect(), nnl).as_ptr(), pick(idx.iter().map(|&i| &proofs[i].e_db).collect(), nnl).as_ptr(), accept.as_mut_ptr(), fault.as_mut_ptr(),
                    )
                });
                for (k, &i) in idx.iter().enumerate() {
                    out[i] = if ok[k] { Verdict::from_flags(accept[k], fault[k]) } else { Verdict::Reject };
                }
            });
        }
        out
    }
}

/// multiplication_proof.rs:148-154
pub(crate) fn sample_paillier_random(modulo: &BigInt) -> BigInt {
    loop {
        let r = BigInt::sample_below(modulo);
        if BigInt::gcd(&r, modulo) == BigInt::one() {
            return r;
        }
    }
}
