//! `extern "C"` declarations of include/zkp_b200.h (one-to-one; keep in sync with the header).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_longlong, c_void};

#[repr(C)]
pub struct zkp_ctx {
    _private: [u8; 0],
}

pub const ZKP_OK: c_int = 0;
pub const ZKP_RP_OPEN: u8 = 0;
pub const ZKP_RP_MASK1: u8 = 1;
pub const ZKP_RP_MASK2: u8 = 2;
pub const ZKP_CK_M2: usize = 11;

extern "C" {
    pub fn zkp_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut zkp_ctx) -> c_int;
    pub fn zkp_ctx_destroy(ctx: *mut zkp_ctx);
    pub fn zkp_last_error(ctx: *const zkp_ctx) -> *const c_char;
    pub fn zkp_version() -> c_int;
    pub fn zkp_sm_count(ctx: *const zkp_ctx) -> c_int;
    pub fn zkp_sync(ctx: *mut zkp_ctx) -> c_int;
    pub fn zkp_profile_enable(ctx: *mut zkp_ctx, on: c_int) -> c_int;
    pub fn zkp_profile_reset(ctx: *mut zkp_ctx) -> c_int;
    pub fn zkp_profile_get(ctx: *mut zkp_ctx, kernel: c_int, ms: *mut c_double, launches: *mut c_longlong, units: *mut c_double) -> c_int;
    pub fn zkp_set_key(ctx: *mut zkp_ctx, n: *const u32, n_limbs: c_int) -> c_int;
    pub fn zkp_set_modulus(ctx: *mut zkp_ctx, m: *const u32, m_limbs: c_int, e: *const u32, e_limbs: c_int) -> c_int;
    pub fn zkp_nn_limbs(ctx: *const zkp_ctx) -> c_int;
    pub fn zkp_modexp_shared(ctx: *mut zkp_ctx, bases: *const u32, base_limbs: c_int, batch: c_int, out: *mut u32) -> c_int;
    pub fn zkp_paillier_enc(ctx: *mut zkp_ctx, m: *const u32, m_limbs: c_int, r: *const u32, r_limbs: c_int, batch: c_int, out: *mut u32) -> c_int;
    pub fn zkp_modexp_var(ctx: *mut zkp_ctx, bases: *const u32, exps: *const u32, exp_limbs: c_int, exp_bits: c_int, exp_per: c_int,
                          mods: *const u32, mod_limbs: c_int, mod_per: c_int, batch: c_int, out: *mut u32) -> c_int;
    pub fn zkp_modmul(ctx: *mut zkp_ctx, which_nn: c_int, a: *const u32, b: *const u32, b_per: c_int, batch: c_int, out: *mut u32) -> c_int;
    pub fn zkp_sha256_transcript(ctx: *mut zkp_ctx, items: *const u32, limbs: c_int, count: c_int, batch: c_int, digest: *mut u8) -> c_int;
    pub fn zkp_rangeproof_ni_prove(ctx: *mut zkp_ctx, batch: c_int, ef: c_int, w_limbs: c_int, range: *const u32, x: *const u32, r: *const u32,
                                   w1: *const u32, swap: *const u8, r1: *const u32, r2: *const u32, c1: *mut u32, c2: *mut u32,
                                   digest: *mut u8, kind: *mut u8, resp_w: *mut u32, resp_r: *mut u32) -> c_int;
    pub fn zkp_rangeproof_ni_verify(ctx: *mut zkp_ctx, batch: c_int, ef: c_int, w_limbs: c_int, range: *const u32, cipher_x: *const u32,
                                    c1: *const u32, c2: *const u32, kind: *const u8, resp_w: *const u32, resp_r: *const u32,
                                    accept: *mut u8, fault: *mut u8, digest: *mut u8) -> c_int;
    pub fn zkp_correct_key_ni_verify(ctx: *mut zkp_ctx, batch: c_int, n_limbs: c_int, n: *const u32, sigma: *const u32, salt: *const u8,
                                     salt_len: c_int, accept: *mut u8, rho: *mut u32) -> c_int;
    pub fn zkp_zero_prove(ctx: *mut zkp_ctx, batch: c_int, r: *const u32, c: *const u32, r_prime: *const u32, z: *mut u32, a: *mut u32) -> c_int;
    pub fn zkp_zero_verify(ctx: *mut zkp_ctx, batch: c_int, c: *const u32, z: *const u32, a: *const u32, accept: *mut u8) -> c_int;
    pub fn zkp_ciphertext_prove(ctx: *mut zkp_ctx, batch: c_int, z_limbs: c_int, x: *const u32, r: *const u32, c: *const u32,
                                x_prime: *const u32, r_prime: *const u32, z1: *mut u32, z2: *mut u32, c_prime: *mut u32) -> c_int;
    pub fn zkp_ciphertext_verify(ctx: *mut zkp_ctx, batch: c_int, z_limbs: c_int, c: *const u32, z1: *const u32, z2: *const u32,
                                 c_prime: *const u32, accept: *mut u8) -> c_int;
    pub fn zkp_mul_prove(ctx: *mut zkp_ctx, batch: c_int, a: *const u32, b: *const u32, r_a: *const u32, r_b: *const u32, r_c: *const u32,
                         e_a: *const u32, e_b: *const u32, e_c: *const u32, d: *const u32, r_d: *const u32, f: *mut u32, z1: *mut u32,
                         z2: *mut u32, e_d: *mut u32, e_db: *mut u32, fault: *mut u8) -> c_int;
    pub fn zkp_mul_verify(ctx: *mut zkp_ctx, batch: c_int, e_a: *const u32, e_b: *const u32, e_c: *const u32, f: *const u32, z1: *const u32,
                          z2: *const u32, e_d: *const u32, e_db: *const u32, accept: *mut u8, fault: *mut u8) -> c_int;
    pub fn zkp_verlin_prove(ctx: *mut zkp_ctx, batch: c_int, z_limbs: c_int, x: *const u32, x_prime: *const u32, x_dp: *const u32,
                            r_x: *const u32, c: *const u32, c_prime: *const u32, phi_x: *const u32, a: *const u32, a_prime: *const u32,
                            a_dp: *const u32, r_a: *const u32, phi_a: *mut u32, z: *mut u32, z_prime: *mut u32, z_dp: *mut u32,
                            r_z: *mut u32) -> c_int;
    pub fn zkp_verlin_verify(ctx: *mut zkp_ctx, batch: c_int, z_limbs: c_int, c: *const u32, c_prime: *const u32, phi_x: *const u32,
                             phi_a: *const u32, z: *const u32, z_prime: *const u32, z_dp: *const u32, r_z: *const u32, accept: *mut u8) -> c_int;
    // the remaining public proofs: CorrectOpening, CompositeDLogProof, CorrectMessageProof
    pub fn zkp_verify_opening(ctx: *mut zkp_ctx, batch: c_int, m_limbs: c_int, m: *const u32, r: *const u32, c: *const u32, ok: *mut u8) -> c_int;
    pub fn zkp_dlog_prove(ctx: *mut zkp_ctx, batch: c_int, n_limbs: c_int, n: *const u32, g: *const u32, ni: *const u32, secret: *const u32,
                          secret_limbs: c_int, r: *const u32, r_limbs: c_int, y_limbs: c_int, x: *mut u32, y: *mut u32, fault: *mut u8) -> c_int;
    pub fn zkp_dlog_verify(ctx: *mut zkp_ctx, batch: c_int, n_limbs: c_int, n: *const u32, g: *const u32, ni: *const u32, x: *const u32,
                           y: *const u32, y_limbs: c_int, accept: *mut u8, fault: *mut u8) -> c_int;
    pub fn zkp_correct_message_prove(ctx: *mut zkp_ctx, batch: c_int, m_count: c_int, m_limbs: c_int, valid: *const u32, msg: *const u32,
                                     r: *const u32, e_rand: *const u32, z_rand: *const u32, w: *const u32, ciphertext: *mut u32,
                                     e_vec: *mut u32, z_vec: *mut u32, a_vec: *mut u32, fault: *mut u8) -> c_int;
    pub fn zkp_correct_message_verify(ctx: *mut zkp_ctx, batch: c_int, m_count: c_int, m_limbs: c_int, e_limbs: c_int, ciphertext: *const u32,
                                      valid: *const u32, e_vec: *const u32, z_vec: *const u32, a_vec: *const u32, accept: *mut u8,
                                      fault: *mut u8) -> c_int;
    pub fn zkp_imad_peak(ctx: *mut zkp_ctx, variant: c_int, mads_per_s: *mut c_double) -> c_int;
}
