//! VerlinProof: phi_x = c^x c'^x' Enc(x'', r_x) (reference src/zkproofs/verlin_proof.rs:34-165) over
//! zkp_verlin_prove / zkp_verlin_verify.
use curv::arithmetic::traits::*;
use curv::BigInt;
use paillier::EncryptionKey;
use serde::{Deserialize, Serialize};

use super::errors::IncorrectProof;
use super::multiplication_proof::sample_paillier_random;
use crate::engine::{fits, group_by_key, pack, require_one_key, unpack, Engine, Verdict};
use crate::ffi;

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct VerlinProof {
    pub phi_a: BigInt,
    pub z: BigInt,
    pub z_prime: BigInt,
    pub z_double_prime: BigInt,
    pub r_z: BigInt,
}

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct VerlinWitness {
    pub x: BigInt,
    pub x_prime: BigInt,
    pub x_double_prime: BigInt,
    pub r_x: BigInt,
}

#[derive(Clone, PartialEq, Debug, Serialize, Deserialize)]
pub struct VerlinStatement {
    pub ek: EncryptionKey,
    pub c: BigInt,
    pub c_prime: BigInt,
    pub phi_x: BigInt,
}

impl VerlinProof {
    /// verlin_proof.rs:60-99
    pub fn prove(witness: &VerlinWitness, statement: &VerlinStatement) -> Self {
        Self::prove_batch(std::slice::from_ref(witness), std::slice::from_ref(statement)).pop().unwrap()
    }
    /// verlin_proof.rs:101-134
    pub fn verify(&self, statement: &VerlinStatement) -> Result<(), IncorrectProof> {
        Self::verify_batch(&[self], std::slice::from_ref(statement))[0].into_result("unreachable")
    }

    pub fn prove_batch(witness: &[VerlinWitness], statement: &[VerlinStatement]) -> Vec<VerlinProof> {
        assert_eq!(witness.len(), statement.len());
        if statement.is_empty() {
            return Vec::new();
        }
        require_one_key(statement.iter().map(|s| &s.ek), "VerlinProof::prove_batch");
        // a, a', a'', r_a per statement (:61-67)
        let (mut a, mut a_prime, mut a_dp, mut r_a) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for s in statement {
            a.push(BigInt::sample_below(&s.ek.n));
            a_prime.push(BigInt::sample_below(&s.ek.n));
            a_dp.push(BigInt::sample_below(&s.ek.n));
            r_a.push(sample_paillier_random(&s.ek.n));
        }
        Engine::with(|eng| {
            eng.use_key(&statement[0].ek);
            let (nl, nnl, zl, b) = (eng.nl(), eng.nnl(), eng.zl(), statement.len());
            let (mut phi_a, mut z, mut zp, mut zdp, mut r_z) = (vec![0u32; b * nnl], vec![0u32; b * zl], vec![0u32; b * zl], vec![0u32; b * zl], vec![0u32; b * nnl]);
            eng.check(unsafe {
                ffi::zkp_verlin_prove(
                    eng.h, b as i32, zl as i32,
                    pack(witness.iter().map(|w| &w.x), nl).as_ptr(), pack(witness.iter().map(|w| &w.x_prime), nl).as_ptr(),
                    pack(witness.iter().map(|w| &w.x_double_prime), nl).as_ptr(), pack(witness.iter().map(|w| &w.r_x), nl).as_ptr(),
                    pack(statement.iter().map(|s| &s.c), nnl).as_ptr(), pack(statement.iter().map(|s| &s.c_prime), nnl).as_ptr(),
                    pack(statement.iter().map(|s| &s.phi_x), nnl).as_ptr(),
                    pack(a.iter(), nl).as_ptr(), pack(a_prime.iter(), nl).as_ptr(), pack(a_dp.iter(), nl).as_ptr(), pack(r_a.iter(), nl).as_ptr(),
                    phi_a.as_mut_ptr(), z.as_mut_ptr(), zp.as_mut_ptr(), zdp.as_mut_ptr(), r_z.as_mut_ptr(),
                )
            });
            let (phi_a, z, zp, zdp, r_z) = (unpack(&phi_a, nnl), unpack(&z, zl), unpack(&zp, zl), unpack(&zdp, zl), unpack(&r_z, nnl));
            (0..b)
                .map(|i| VerlinProof { phi_a: phi_a[i].clone(), z: z[i].clone(), z_prime: zp[i].clone(), z_double_prime: zdp[i].clone(), r_z: r_z[i].clone() })
                .collect()
        })
    }

    /// c, c', phi_x, phi_a are hashed as given and z, z' are exponents (:147-155): wider than their rows, that proof is
    /// rejected; z'' only matters mod n and r_z mod n^2 (:157-163).
    pub fn verify_batch(proofs: &[&VerlinProof], statement: &[VerlinStatement]) -> Vec<Verdict> {
        assert_eq!(proofs.len(), statement.len());
        let mut out = vec![Verdict::Reject; proofs.len()];
        for (ek, idx) in group_by_key(statement.iter().map(|s| &s.ek)) {
            Engine::with(|eng| {
                eng.use_key(&ek);
                let (nnl, zl) = (eng.nnl(), eng.zl());
                let zero = BigInt::zero();
                let ok: Vec<bool> = idx
                    .iter()
                    .map(|&i| {
                        let (s, p) = (&statement[i], proofs[i]);
                        fits(&s.c, nnl) && fits(&s.c_prime, nnl) && fits(&s.phi_x, nnl) && fits(&p.phi_a, nnl) && fits(&p.z, zl) && fits(&p.z_prime, zl)
                    })
                    .collect();
                // rows of the proofs that can be laid out; zero rows for the others (their verdict is already Reject)
                let pick = |rows: Vec<&BigInt>, limbs: usize| -> Vec<u32> {
                    pack(rows.into_iter().zip(&ok).map(|(x, &k)| if k { x } else { &zero }), limbs)
                };
                let zdp: Vec<BigInt> = idx
                    .iter()
                    .map(|&i| if fits(&proofs[i].z_double_prime, zl) { proofs[i].z_double_prime.clone() } else { &proofs[i].z_double_prime % &ek.n })
                    .collect();
                let r_z: Vec<BigInt> = idx.iter().map(|&i| &proofs[i].r_z % &ek.nn).collect();
                let mut accept = vec![0u8; idx.len()];
                eng.check(unsafe {
                    ffi::zkp_verlin_verify(
                        eng.h, idx.len() as i32, zl as i32,
                        pick(idx.iter().map(|&i| &statement[i].c).collect(), nnl).as_ptr(), pick(idx.iter().map(|&i| &statement[i].c_prime).collect(), nnl).as_ptr(), pick(idx.iter().map(|&i| &statement[i].phi_x).collect(), nnl).as_ptr(),
                        pick(idx.iter().map(|&i| &proofs[i].phi_a).collect(), nnl).as_ptr(), pick(idx.iter().map(|&i| &proofs[i].z).collect(), zl).as_ptr(), pick(idx.iter().map(|&i| &proofs[i].z_prime).collect(), zl).as_ptr(),
                        pack(zdp.iter(), zl).as_ptr(), pack(r_z.iter(), nnl).as_ptr(), accept.as_mut_ptr(),
                    )
                });
                for (k, &i) in idx.iter().enumerate() {
                    out[i] = if ok[k] { Verdict::from_flags(accept[k], 0) } else { Verdict::Reject };
                }
            });
        }
        out
    }
}
