//! One engine context per calling thread, limb conversions, and the batch plumbing shared by the proofs.
//!
//! The reference's functions borrow `&` data and may be called from any thread (they fan out on rayon's pool); a
//! `zkp_ctx` is single-threaded.  So the shim keeps ONE context per thread in a thread-local, created on first use and
//! reused by every `prove` / `verify` of that thread (creating a context allocates device scratch: never per call).
use std::cell::RefCell;
use std::ptr;

use curv::arithmetic::traits::*;
use curv::BigInt;
use paillier::EncryptionKey;

use crate::ffi;

pub struct Engine {
    pub(crate) h: *mut ffi::zkp_ctx,
    key: Option<BigInt>,
    nl: usize,
}

thread_local! {
    static ENGINE: RefCell<Option<Engine>> = RefCell::new(None);
}

impl Engine {
    /// `ZKP_B200_DEVICE` picks the GPU of this process (default 0): one process per GPU, as the engine is sharded.
    fn create() -> Engine {
        let device = std::env::var("ZKP_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut h = ptr::null_mut();
        let rc = unsafe { ffi::zkp_ctx_create(device, ptr::null_mut(), &mut h) };
        assert_eq!(rc, ffi::ZKP_OK, "zkp_ctx_create failed: no usable CUDA device (the engine has no CPU fallback)");
        Engine { h, key: None, nl: 0 }
    }

    /// Runs `f` with this thread's engine.
    pub fn with<R>(f: impl FnOnce(&mut Engine) -> R) -> R {
        ENGINE.with(|cell| {
            let mut slot = cell.borrow_mut();
            if slot.is_none() {
                *slot = Some(Engine::create());
            }
            f(slot.as_mut().unwrap())
        })
    }

    pub(crate) fn check(&self, rc: i32) {
        if rc != ffi::ZKP_OK {
            let msg = unsafe { std::ffi::CStr::from_ptr(ffi::zkp_last_error(self.h)) };
            panic!("zkp_b200 error {}: {}", rc, msg.to_string_lossy());
        }
    }

    /// `zkp_set_key` when the key differs from the one resident on the device (it derives n^2, the Montgomery constants
    /// and the recoded exponent: once per key, not per call).
    pub(crate) fn use_key(&mut self, ek: &EncryptionKey) {
        if self.key.as_ref() != Some(&ek.n) {
            let nl = limbs_for_bits(ek.n.bit_length());
            self.check(unsafe { ffi::zkp_set_key(self.h, to_limbs(&ek.n, nl).as_ptr(), nl as i32) });
            self.key = Some(ek.n.clone());
            self.nl = nl;
        }
    }
    /// limbs of n / of n^2 / of the unreduced responses x' + x e under the current key
    pub(crate) fn nl(&self) -> usize {
        self.nl
    }
    pub(crate) fn nnl(&self) -> usize {
        2 * self.nl
    }
    pub(crate) fn zl(&self) -> usize {
        self.nl + 12
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { ffi::zkp_ctx_destroy(self.h) }
    }
}

/// Row widths are multiples of 4 limbs (16-byte rows, include/zkp_b200.h).
pub(crate) fn limbs_for_bits(bits: usize) -> usize {
    (std::cmp::max(bits, 1) + 127) / 128 * 4
}

/// `BigInt::to_bytes()` (big-endian, minimal) -> fixed-width little-endian u32 limbs.
pub(crate) fn to_limbs(x: &BigInt, limbs: usize) -> Vec<u32> {
    let be = BigInt::to_bytes(x);
    assert!(be.len() <= 4 * limbs, "value wider than its row");
    let mut out = vec![0u32; limbs];
    for (i, b) in be.iter().rev().enumerate() {
        out[i / 4] |= (*b as u32) << (8 * (i % 4));
    }
    out
}
pub(crate) fn from_limbs(l: &[u32]) -> BigInt {
    let mut be = Vec::with_capacity(4 * l.len());
    for w in l.iter().rev() {
        be.extend_from_slice(&w.to_be_bytes());
    }
    BigInt::from_bytes(&be)
}
pub(crate) fn fits(x: &BigInt, limbs: usize) -> bool {
    x.bit_length() <= 32 * limbs
}
/// `[rows][limbs]`, dense
pub(crate) fn pack<'a>(xs: impl IntoIterator<Item = &'a BigInt>, limbs: usize) -> Vec<u32> {
    xs.into_iter().flat_map(|x| to_limbs(x, limbs)).collect()
}
pub(crate) fn unpack(v: &[u32], limbs: usize) -> Vec<BigInt> {
    v.chunks(limbs).map(from_limbs).collect()
}

/// Verdict of one proof of a verification batch: what the reference's `verify` would have done with it.
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum Verdict {
    /// `Ok(())`
    Accept,
    /// `Err(IncorrectProof)`
    Reject,
    /// the reference panics on this input (`unwrap()` on a non-invertible value, `assert_eq!`, index out of range)
    Panic,
}
impl Verdict {
    pub(crate) fn from_flags(accept: u8, fault: u8) -> Verdict {
        if fault != 0 {
            Verdict::Panic
        } else if accept == 1 {
            Verdict::Accept
        } else {
            Verdict::Reject
        }
    }
    /// the single-proof form: panic where the reference panics, else its `Result`
    pub(crate) fn into_result(self, panic_text: &str) -> Result<(), crate::zkproofs::IncorrectProof> {
        match self {
            Verdict::Accept => Ok(()),
            Verdict::Reject => Err(crate::zkproofs::IncorrectProof),
            Verdict::Panic => panic!("{}", panic_text),
        }
    }
}

/// A verification batch is untrusted input and every statement carries its own key, while the device verifies one
/// key per launch: group the indices by key (first-appearance order) and let the caller run one device call per group.
pub(crate) fn group_by_key<'a>(keys: impl IntoIterator<Item = &'a EncryptionKey>) -> Vec<(EncryptionKey, Vec<usize>)> {
    let mut groups: Vec<(EncryptionKey, Vec<usize>)> = Vec::new();
    for (i, ek) in keys.into_iter().enumerate() {
        match groups.iter_mut().find(|(k, _)| k.n == ek.n) {
            Some((_, idx)) => idx.push(i),
            None => groups.push((ek.clone(), vec![i])),
        }
    }
    groups
}
/// A proving batch is the prover's own: one key, or it is a usage error.
pub(crate) fn require_one_key<'a>(mut keys: impl Iterator<Item = &'a EncryptionKey>, who: &str) {
    if let Some(first) = keys.next() {
        assert!(keys.all(|k| k.n == first.n), "{}: the statements of one proving batch must share the key", who);
    }
}
