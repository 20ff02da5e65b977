//! NiCorrectKeyProof (reference src/zkproofs/correct_key_ni.rs:26-117).  `verify`: the 11 sigma_i^N mod N, the
//! rho derivation (salt hash, seed, mask_generation, % N) and gcd(P, N) == 1 run on the device
//! (zkp_correct_key_ni_verify, one modulus per proof).  `proof`: rho from the device (zkp_correct_key_ni_rho), the
//! N-th roots as two half-width zkp_modexp_var launches (mod p, mod q) and a CRT recombination here.
use curv::arithmetic::traits::*;
use curv::BigInt;
use paillier::{DecryptionKey, EncryptionKey};
use serde::{Deserialize, Serialize};

use super::errors::IncorrectProof;
use crate::engine::{fits, limbs_for_bits, pack, unpack, Engine, Verdict};
use crate::ffi;

pub const SALT_STRING: &[u8] = &[75, 90, 101, 110]; // correct_key_ni.rs:28
const M2: usize = ffi::ZKP_CK_M2 as usize; // correct_key_ni.rs:29

#[derive(Clone, Debug, Serialize, Deserialize)]
pub struct NiCorrectKeyProof {
    #[serde(with = "crate::serialize::vecbigint")]
    pub sigma_vec: Vec<BigInt>,
}

impl NiCorrectKeyProof {
    /// correct_key_ni.rs:42-71 (note the name: `proof`, not `prove`)
    pub fn proof(dk: &DecryptionKey, salt_str: Option<&'static [u8]>) -> NiCorrectKeyProof {
        Self::proof_batch(std::slice::from_ref(dk), salt_str).pop().unwrap()
    }

    /// One proof per decryption key; keys of one batch share a width class (rows are as wide as the widest N).
    pub fn proof_batch(dks: &[DecryptionKey], salt_str: Option<&'static [u8]>) -> Vec<NiCorrectKeyProof> {
        let salt = salt_str.unwrap_or(SALT_STRING);
        let b = dks.len();
        if b == 0 {
            return Vec::new();
        }
        let n: Vec<BigInt> = dks.iter().map(|dk| &dk.q * &dk.p).collect();
        let nl = limbs_for_bits(n.iter().map(|x| x.bit_length()).max().unwrap());
        let hl = limbs_for_bits(dks.iter().map(|dk| std::cmp::max(dk.p.bit_length(), dk.q.bit_length())).max().unwrap());
        Engine::with(|eng| {
            let mut rho = vec![0u32; b * M2 * nl];
            eng.check(unsafe { ffi::zkp_correct_key_ni_rho(eng.h, b as i32, nl as i32, pack(n.iter(), nl).as_ptr(), salt.as_ptr(), salt.len() as i32, rho.as_mut_ptr()) });
            let rho = unpack(&rho, nl);
            // extract_nroot (kzen-paillier): sigma = rho^(N^-1 mod phi) by CRT - rho^(N^-1 mod p-1) mod p, the same mod q
            let half = |prime: &dyn Fn(&DecryptionKey) -> &BigInt| -> Vec<BigInt> {
                let mods: Vec<BigInt> = dks.iter().map(|dk| prime(dk).clone()).collect();
                let exps: Vec<BigInt> = dks
                    .iter()
                    .zip(&n)
                    .map(|(dk, n)| {
                        let pm1 = prime(dk) - BigInt::one();
                        BigInt::mod_inv(&(n % &pm1), &pm1).expect("extract_nroot: N is not invertible mod phi(N)")
                    })
                    .collect();
                let bases: Vec<BigInt> = (0..b * M2).map(|t| &rho[t] % &mods[t / M2]).collect();
                let mut out = vec![0u32; b * M2 * hl];
                eng.check(unsafe {
                    ffi::zkp_modexp_var(
                        eng.h, pack(bases.iter(), hl).as_ptr(), pack(exps.iter(), hl).as_ptr(), hl as i32, (32 * hl) as i32, M2 as i32,
                        pack(mods.iter(), hl).as_ptr(), hl as i32, M2 as i32, (b * M2) as i32, out.as_mut_ptr(),
                    )
                });
                unpack(&out, hl)
            };
            let sp = half(&|dk| &dk.p);
            let sq = half(&|dk| &dk.q);
            (0..b)
                .map(|k| {
                    let (p, q) = (&dks[k].p, &dks[k].q);
                    let pinv = BigInt::mod_inv(&(p % q), q).expect("p is not invertible mod q");
                    let sigma_vec = (0..M2)
                        .map(|i| {
                            let (a, c) = (&sp[k * M2 + i], &sq[k * M2 + i]);
                            // a + p * ((c - a) * p^-1 mod q)
                            let h = BigInt::mod_mul(&BigInt::mod_sub(c, a, q), &pinv, q);
                            let ph = p * &h;
                            a + &ph
                        })
                        .collect();
                    NiCorrectKeyProof { sigma_vec }
                })
                .collect()
        })
    }

    /// correct_key_ni.rs:73-100
    pub fn verify(&self, ek: &EncryptionKey, salt_str: &[u8]) -> Result<(), IncorrectProof> {
        Self::verify_batch(&[self], std::slice::from_ref(ek), salt_str)[0].into_result("index out of bounds: sigma_vec shorter than M2")
    }

    /// One modulus per proof, one device call: 11 modexps with exponent = modulus = N_b each.  `Verdict::Panic` where
    /// `sigma_vec[i]` indexes out of range in the reference (:92).  mod_pow reduces its base, so a sigma wider than its
    /// row is reduced mod N here and decides nothing but itself.
    pub fn verify_batch(proofs: &[&NiCorrectKeyProof], eks: &[EncryptionKey], salt_str: &[u8]) -> Vec<Verdict> {
        assert_eq!(proofs.len(), eks.len());
        let b = proofs.len();
        if b == 0 {
            return Vec::new();
        }
        let nl = limbs_for_bits(eks.iter().map(|ek| ek.n.bit_length()).max().unwrap());
        let short: Vec<bool> = proofs.iter().map(|p| p.sigma_vec.len() < M2).collect();
        let zero = BigInt::zero();
        let mut sigma: Vec<BigInt> = Vec::with_capacity(b * M2);
        for (k, p) in proofs.iter().enumerate() {
            for i in 0..M2 {
                sigma.push(if short[k] { zero.clone() } else if fits(&p.sigma_vec[i], nl) { p.sigma_vec[i].clone() } else { &p.sigma_vec[i] % &eks[k].n });
            }
        }
        let mut accept = vec![0u8; b];
        Engine::with(|eng| {
            eng.check(unsafe {
                ffi::zkp_correct_key_ni_verify(
                    eng.h, b as i32, nl as i32, pack(eks.iter().map(|ek| &ek.n), nl).as_ptr(), pack(sigma.iter(), nl).as_ptr(), salt_str.as_ptr(),
                    salt_str.len() as i32, accept.as_mut_ptr(), std::ptr::null_mut(),
                )
            });
        });
        (0..b).map(|k| if short[k] { Verdict::Panic } else { Verdict::from_flags(accept[k], 0) }).collect()
    }
}
