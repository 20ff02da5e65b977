//! The public names of `zk_paillier::zkproofs` (reference src/zkproofs/mod.rs:29-43), one module per reference file.
//! Each proof keeps the reference's struct (field names, visibility, serde attributes), its `prove` / `verify`
//! signatures and its error conventions (`Err(IncorrectProof)` vs panic), and adds the engine's one API extension:
//! `*_batch` associated functions, because one statement per call leaves a B200 97 % idle.
mod correct_ciphertext;
mod correct_key_ni;
mod correct_message;
mod correct_opening;
mod errors;
mod multiplication_proof;
mod range_proof_ni;
mod utils;
mod verlin_proof;
mod wi_dlog_proof;
mod zero_enc_proof;

pub use self::{
    correct_ciphertext::*,
    correct_key_ni::{NiCorrectKeyProof, SALT_STRING},
    correct_message::CorrectMessageProof,
    correct_opening::CorrectOpening,
    multiplication_proof::*,
    range_proof_ni::{EncryptedPairs, Proof, RangeProofNi, RangeStatement, Response},
    verlin_proof::*,
    wi_dlog_proof::*,
    zero_enc_proof::*,
};
pub use self::{errors::IncorrectProof, utils::compute_digest};
pub use crate::engine::Verdict;
