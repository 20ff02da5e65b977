//! RangeProofNi / NiCorrectKeyProof over the C ABI.  Field names, visibility, serde attributes and error
//! conventions follow the reference (range_proof_ni.rs:35-44, range_proof.rs:31-81, correct_key_ni.rs:34-39,
//! errors.rs:5-13); CompositeDLogProof and verify_opening follow; the sigma-protocol structs and CorrectMessageProof bind
//! zkp_{zero,ciphertext,mul,verlin,correct_message}_{prove,verify} the same way.
use std::fmt;
use std::ptr;

use curv::arithmetic::traits::*;
use curv::BigInt;
use paillier::EncryptionKey;
use rand::RngCore;
use serde::{Deserialize, Serialize};

use crate::ffi;

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub struct IncorrectProof;
impl fmt::Display for IncorrectProof {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        write!(f, "given proof doesn't match a statement")
    }
}
impl std::error::Error for IncorrectProof {}

const SECURITY_PARAMETER: usize = 128; // range_proof_ni.rs:23
const W_LIMBS: usize = 12;

/// One engine context per thread (a zkp_ctx is single-threaded; this replaces the rayon fan-out).
pub struct Engine(*mut ffi::zkp_ctx);
impl Engine {
    pub fn new(device: i32) -> Engine {
        let mut h = ptr::null_mut();
        let rc = unsafe { ffi::zkp_ctx_create(device, ptr::null_mut(), &mut h) };
        assert_eq!(rc, ffi::ZKP_OK, "no CUDA device: the engine has no CPU fallback");
        Engine(h)
    }
    fn check(&self, rc: i32) {
        if rc != ffi::ZKP_OK {
            let msg = unsafe { std::ffi::CStr::from_ptr(ffi::zkp_last_error(self.0)) };
            panic!("zkp_b200 error {}: {}", rc, msg.to_string_lossy());
        }
    }
}
impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { ffi::zkp_ctx_destroy(self.0) }
    }
}

/// BigInt::to_bytes() (big-endian, minimal) -> fixed-width little-endian u32 limbs.
pub fn to_limbs(x: &BigInt, limbs: usize) -> Vec<u32> {
    let be = BigInt::to_bytes(x);
    assert!(be.len() <= 4 * limbs, "value wider than the row");
    let mut out = vec![0u32; limbs];
    for (i, b) in be.iter().rev().enumerate() {
        out[i / 4] |= (*b as u32) << (8 * (i % 4));
    }
    out
}
pub fn from_limbs(l: &[u32]) -> BigInt {
    let mut be = Vec::with_capacity(4 * l.len());
    for w in l.iter().rev() {
        be.extend_from_slice(&w.to_be_bytes());
    }
    BigInt::from_bytes(&be)
}

#[derive(Default, Debug, Serialize, Deserialize, Clone)]
pub struct EncryptedPairs {
    #[serde(with = "crate::zkproofs::vecbigint")]
    pub c1: Vec<BigInt>,
    #[serde(with = "crate::zkproofs::vecbigint")]
    pub c2: Vec<BigInt>,
}

#[derive(Debug, Serialize, Deserialize, Clone)]
pub enum Response {
    Open {
        #[serde(with = "crate::zkproofs::bigint")]
        w1: BigInt,
        #[serde(with = "crate::zkproofs::bigint")]
        r1: BigInt,
        #[serde(with = "crate::zkproofs::bigint")]
        w2: BigInt,
        #[serde(with = "crate::zkproofs::bigint")]
        r2: BigInt,
    },
    Mask {
        j: u8,
        #[serde(with = "crate::zkproofs::bigint")]
        masked_x: BigInt,
        #[serde(with = "crate::zkproofs::bigint")]
        masked_r: BigInt,
    },
}

#[derive(Debug, Serialize, Deserialize, Clone)]
pub struct Proof(Vec<Response>);

#[derive(Debug, Serialize, Deserialize, Clone)]
pub struct RangeProofNi {
    ek: EncryptionKey,
    range: BigInt,
    ciphertext: BigInt,
    encrypted_pairs: EncryptedPairs,
    proof: Proof,
    error_factor: usize,
}

impl RangeProofNi {
    /// range_proof_ni.rs:47-82.  Randomness is drawn here (as the reference does inside
    /// generate_encrypted_pairs, range_proof.rs:136-159) and handed to the engine as arrays.
    pub fn prove(ek: &EncryptionKey, range: &BigInt, ciphertext: &BigInt, secret_x: &BigInt, secret_r: &BigInt) -> RangeProofNi {
        let eng = Engine::new(0);
        let ef = SECURITY_PARAMETER;
        let nl = (ek.n.bit_length() + 127) / 128 * 4;
        let nnl = 2 * nl;
        let third = range.div_floor(&BigInt::from(3));
        let two_thirds = BigInt::from(2) * &third;
        let mut w1 = Vec::with_capacity(ef * W_LIMBS);
        let mut r1 = Vec::with_capacity(ef * nl);
        let mut r2 = Vec::with_capacity(ef * nl);
        let mut swap = vec![0u8; ef];
        rand::thread_rng().fill_bytes(&mut swap);
        for i in 0..ef {
            swap[i] &= 1;
            w1.extend(to_limbs(&BigInt::sample_range(&third, &two_thirds), W_LIMBS));
            r1.extend(to_limbs(&BigInt::sample_below(&ek.n), nl));
            r2.extend(to_limbs(&BigInt::sample_below(&ek.n), nl));
        }
        let (mut c1, mut c2) = (vec![0u32; ef * nnl], vec![0u32; ef * nnl]);
        let (mut kind, mut digest) = (vec![0u8; ef], vec![0u8; 32]);
        let (mut resp_w, mut resp_r) = (vec![0u32; ef * 2 * W_LIMBS], vec![0u32; ef * 2 * nl]);
        unsafe {
            eng.check(ffi::zkp_set_key(eng.0, to_limbs(&ek.n, nl).as_ptr(), nl as i32));
            eng.check(ffi::zkp_rangeproof_ni_prove(
                eng.0, 1, ef as i32, W_LIMBS as i32,
                to_limbs(range, W_LIMBS).as_ptr(), to_limbs(secret_x, W_LIMBS).as_ptr(), to_limbs(secret_r, nl).as_ptr(),
                w1.as_ptr(), swap.as_ptr(), r1.as_ptr(), r2.as_ptr(),
                c1.as_mut_ptr(), c2.as_mut_ptr(), digest.as_mut_ptr(), kind.as_mut_ptr(), resp_w.as_mut_ptr(), resp_r.as_mut_ptr(),
            ));
        }
        let rows = |v: &[u32], w: usize| -> Vec<BigInt> { v.chunks(w).map(from_limbs).collect() };
        let responses = (0..ef)
            .map(|i| {
                let w = &resp_w[i * 2 * W_LIMBS..(i + 1) * 2 * W_LIMBS];
                let r = &resp_r[i * 2 * nl..(i + 1) * 2 * nl];
                match kind[i] {
                    ffi::ZKP_RP_OPEN => Response::Open {
                        w1: from_limbs(&w[..W_LIMBS]),
                        r1: from_limbs(&r[..nl]),
                        w2: from_limbs(&w[W_LIMBS..]),
                        r2: from_limbs(&r[nl..]),
                    },
                    j => Response::Mask { j, masked_x: from_limbs(&w[..W_LIMBS]), masked_r: from_limbs(&r[..nl]) },
                }
            })
            .collect();
        RangeProofNi {
            ek: ek.clone(),
            range: range.clone(),
            ciphertext: ciphertext.clone(),
            encrypted_pairs: EncryptedPairs { c1: rows(&c1, nnl), c2: rows(&c2, nnl) },
            proof: Proof(responses),
            error_factor: ef,
        }
    }

    /// range_proof_ni.rs:84-107.  Precondition failures panic exactly where the reference does.
    pub fn verify(&self, ek: &EncryptionKey, ciphertext: &BigInt) -> Result<(), IncorrectProof> {
        assert_eq!(ek, &self.ek);
        assert_eq!(ciphertext, &self.ciphertext);
        self.verify_self()
    }

    pub fn verify_self(&self) -> Result<(), IncorrectProof> {
        let eng = Engine::new(0);
        let ef = self.error_factor;
        let nl = (self.ek.n.bit_length() + 127) / 128 * 4;
        let nnl = 2 * nl;
        // responses[i] indexes out of range in the reference when the proof is short (range_proof.rs:274)
        assert!(self.proof.0.len() >= ef && self.encrypted_pairs.c1.len() >= ef && self.encrypted_pairs.c2.len() >= ef);
        let flat = |v: &[BigInt], w: usize| -> Vec<u32> { v.iter().take(ef).flat_map(|x| to_limbs(x, w)).collect() };
        let (mut kind, mut resp_w, mut resp_r) = (vec![0u8; ef], vec![0u32; ef * 2 * W_LIMBS], vec![0u32; ef * 2 * nl]);
        for (i, resp) in self.proof.0.iter().take(ef).enumerate() {
            let (k, wa, ra, wb, rb) = match resp {
                Response::Open { w1, r1, w2, r2 } => (ffi::ZKP_RP_OPEN, w1.clone(), r1.clone(), w2.clone(), r2.clone()),
                Response::Mask { j, masked_x, masked_r } => (
                    if *j == 1 { ffi::ZKP_RP_MASK1 } else { ffi::ZKP_RP_MASK2 }, // `if *j == 1 {c1} else {c2}` (range_proof.rs:321)
                    masked_x.clone(), masked_r.clone(), BigInt::zero(), BigInt::zero(),
                ),
            };
            kind[i] = k;
            resp_w[i * 2 * W_LIMBS..i * 2 * W_LIMBS + W_LIMBS].copy_from_slice(&to_limbs(&wa, W_LIMBS));
            resp_w[i * 2 * W_LIMBS + W_LIMBS..(i + 1) * 2 * W_LIMBS].copy_from_slice(&to_limbs(&wb, W_LIMBS));
            resp_r[i * 2 * nl..i * 2 * nl + nl].copy_from_slice(&to_limbs(&ra, nl));
            resp_r[i * 2 * nl + nl..(i + 1) * 2 * nl].copy_from_slice(&to_limbs(&rb, nl));
        }
        let (mut accept, mut fault) = ([0u8; 1], [0u8; 1]);
        unsafe {
            eng.check(ffi::zkp_set_key(eng.0, to_limbs(&self.ek.n, nl).as_ptr(), nl as i32));
            eng.check(ffi::zkp_rangeproof_ni_verify(
                eng.0, 1, ef as i32, W_LIMBS as i32, to_limbs(&self.range, W_LIMBS).as_ptr(), to_limbs(&self.ciphertext, nnl).as_ptr(),
                flat(&self.encrypted_pairs.c1, nnl).as_ptr(), flat(&self.encrypted_pairs.c2, nnl).as_ptr(), kind.as_ptr(),
                resp_w.as_ptr(), resp_r.as_ptr(), accept.as_mut_ptr(), fault.as_mut_ptr(), ptr::null_mut(),
            ));
        }
        assert_eq!(fault[0], 0, "index out of bounds"); // the reference panics here
        if accept[0] == 1 { Ok(()) } else { Err(IncorrectProof) }
    }
}

#[derive(Clone, Debug, Serialize, Deserialize)]
pub struct NiCorrectKeyProof {
    #[serde(with = "crate::zkproofs::vecbigint")]
    pub sigma_vec: Vec<BigInt>,
}

impl NiCorrectKeyProof {
    /// correct_key_ni.rs:73-100
    pub fn verify(&self, ek: &EncryptionKey, salt_str: &[u8]) -> Result<(), IncorrectProof> {
        let eng = Engine::new(0);
        let nl = (ek.n.bit_length() + 127) / 128 * 4;
        let sigma: Vec<u32> = (0..ffi::ZKP_CK_M2).flat_map(|i| to_limbs(&self.sigma_vec[i], nl)).collect(); // [i] panics if short, as :92
        let mut accept = [0u8; 1];
        unsafe {
            eng.check(ffi::zkp_correct_key_ni_verify(eng.0, 1, nl as i32, to_limbs(&ek.n, nl).as_ptr(), sigma.as_ptr(), salt_str.as_ptr(),
                                                     salt_str.len() as i32, accept.as_mut_ptr(), ptr::null_mut()));
        }
        if accept[0] == 1 { Ok(()) } else { Err(IncorrectProof) }
    }
}

/// wi_dlog_proof.rs:28-39 — fields keep curv's native BigInt serde (no `with =` attribute in the reference)
#[derive(Debug, Serialize, Deserialize, Clone)]
pub struct CompositeDLogProof {
    pub x: BigInt,
    pub y: BigInt,
}
#[allow(non_snake_case)]
#[derive(Debug, Serialize, Deserialize, Clone)]
pub struct DLogStatement {
    pub N: BigInt,
    pub g: BigInt,
    pub ni: BigInt,
}

impl CompositeDLogProof {
    const K: usize = 128;
    const K_PRIME: usize = 128;
    const SAMPLE_S: usize = 256;

    /// wi_dlog_proof.rs:46-65: r is drawn here, x = g^r mod N, the hash and y = r + e * secret come from the engine
    pub fn prove(statement: &DLogStatement, secret: &BigInt) -> CompositeDLogProof {
        let eng = Engine::new(0);
        let nl = (statement.N.bit_length() + 127) / 128 * 4;
        let r = BigInt::sample_below(&BigInt::from(2).pow((Self::K + Self::K_PRIME + Self::SAMPLE_S) as u32));
        let (sl, rl) = ((secret.bit_length() + 127) / 128 * 4, 16usize);
        let yl = (std::cmp::max(rl, sl + 8) + 4) / 4 * 4;
        let (mut x, mut y, mut fault) = (vec![0u32; nl], vec![0u32; yl], [0u8; 1]);
        unsafe {
            eng.check(ffi::zkp_dlog_prove(
                eng.0, 1, nl as i32, to_limbs(&statement.N, nl).as_ptr(), to_limbs(&statement.g, nl).as_ptr(),
                to_limbs(&statement.ni, nl).as_ptr(), to_limbs(secret, sl).as_ptr(), sl as i32, to_limbs(&r, rl).as_ptr(), rl as i32,
                yl as i32, x.as_mut_ptr(), y.as_mut_ptr(), fault.as_mut_ptr(),
            ));
        }
        assert_eq!(fault[0], 0);
        CompositeDLogProof { x: from_limbs(&x), y: from_limbs(&y) }
    }

    /// wi_dlog_proof.rs:66-91: the three asserts of the reference come back as `fault` and panic here as they do there
    pub fn verify(&self, statement: &DLogStatement) -> Result<(), IncorrectProof> {
        let eng = Engine::new(0);
        let nl = (statement.N.bit_length() + 127) / 128 * 4;
        let yl = (self.y.bit_length() + 127) / 128 * 4;
        let (mut accept, mut fault) = ([0u8; 1], [0u8; 1]);
        unsafe {
            eng.check(ffi::zkp_dlog_verify(
                eng.0, 1, nl as i32, to_limbs(&statement.N, nl).as_ptr(), to_limbs(&statement.g, nl).as_ptr(),
                to_limbs(&statement.ni, nl).as_ptr(), to_limbs(&self.x, nl).as_ptr(), to_limbs(&self.y, yl).as_ptr(), yl as i32,
                accept.as_mut_ptr(), fault.as_mut_ptr(),
            ));
        }
        assert_eq!(fault[0], 0, "assertion failed: N > 2^K, gcd(g, N) == 1, gcd(ni, N) == 1");
        if accept[0] == 1 { Ok(()) } else { Err(IncorrectProof) }
    }
}

/// correct_opening.rs:17-30 — `Paillier::verify_opening(ek, m, r, c)`
pub fn verify_opening(ek: &EncryptionKey, m: &BigInt, r: &BigInt, c: &BigInt) -> bool {
    let eng = Engine::new(0);
    let nl = (ek.n.bit_length() + 127) / 128 * 4;
    let mut ok = [0u8; 1];
    unsafe {
        eng.check(ffi::zkp_set_key(eng.0, to_limbs(&ek.n, nl).as_ptr(), nl as i32));
        eng.check(ffi::zkp_verify_opening(eng.0, 1, nl as i32, to_limbs(&(m % &ek.n), nl).as_ptr(), to_limbs(&(r % &ek.n), nl).as_ptr(),
                                          to_limbs(c, 2 * nl).as_ptr(), ok.as_mut_ptr()));
    }
    ok[0] == 1
}

/// serialize.rs:1-31 — decimal-string BigInt codec
pub mod bigint {
    use curv::arithmetic::traits::*;
    use curv::BigInt;
    use serde::{de, ser, Deserialize};
    pub fn serialize<S: ser::Serializer>(x: &BigInt, s: S) -> Result<S::Ok, S::Error> {
        s.serialize_str(&x.to_str_radix(10))
    }
    pub fn deserialize<'de, D: de::Deserializer<'de>>(d: D) -> Result<BigInt, D::Error> {
        let s = String::deserialize(d)?;
        BigInt::from_str_radix(&s, 10).map_err(de::Error::custom)
    }
}
/// serialize.rs:33-78 — sequence of decimal strings
pub mod vecbigint {
    use curv::arithmetic::traits::*;
    use curv::BigInt;
    use serde::ser::SerializeSeq;
    use serde::{de, ser, Deserialize};
    pub fn serialize<S: ser::Serializer>(x: &[BigInt], s: S) -> Result<S::Ok, S::Error> {
        let mut seq = s.serialize_seq(Some(x.len()))?;
        for e in x {
            seq.serialize_element(&e.to_str_radix(10))?;
        }
        seq.end()
    }
    pub fn deserialize<'de, D: de::Deserializer<'de>>(d: D) -> Result<Vec<BigInt>, D::Error> {
        let v = Vec::<String>::deserialize(d)?;
        Ok(v.iter().map(|s| BigInt::from_str_radix(s, 10).unwrap()).collect()) // unwrap: serialize.rs:69
    }
}
