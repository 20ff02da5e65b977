//! reference src/zkproofs/errors.rs:5-13
use std::fmt;

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub struct IncorrectProof;

impl fmt::Display for IncorrectProof {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        write!(f, "given proof doesn't match a statement")
    }
}
impl std::error::Error for IncorrectProof {}
