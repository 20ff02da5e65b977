//! Emits tests/golden/reference_vectors.json from the UNMODIFIED reference crate (see Cargo.toml).
//!
//! What it pins (SURVEY.md section 8c lists these as RECALLED rules, unverifiable without a Rust toolchain):
//!   to_bytes           BigInt::to_bytes of 0, one-byte, leading-zero-limb and wide values (rule H of compute_digest)
//!   compute_digest     zkproofs::compute_digest over item lists that contain 0 and values with leading zero bytes
//!   bigint_serde       serde_json of a bare BigInt (curv's native serde: the fields without `with =`)
//!   encryption_key     serde_json of kzen-paillier's EncryptionKey
//!   paillier_enc       Paillier::encrypt_with_chosen_randomness on the fixed test key (range_proof_ni.rs:141-145)
//!   ni_correct_key     NiCorrectKeyProof::proof(dk, salt) - deterministic: the oracle and the engine must reproduce
//!                      sigma_vec bit for bit - and its JSON
//!   range_proof_ni / zero / ciphertext / mul / verlin / dlog
//!                      one honest and one dishonest proof each, as JSON, with the statement and the reference's own
//!                      verdict: the oracle and the engine must parse the JSON, re-serialize it byte-identically,
//!                      recompute the Fiat-Shamir challenge they imply and return the same verdict
//! Randomness inside prove() is not seedable in the reference, so proofs are pinned through verification and
//! re-serialization, not through re-proving.
use std::env;
use std::fs;

use curv::arithmetic::traits::*;
use curv::BigInt;
use paillier::{EncryptWithChosenRandomness, EncryptionKey, Keypair, Paillier, Randomness, RawPlaintext};
use serde_json::{json, Value};
use zk_paillier::zkproofs::*;

const P: &str = "148677972634832330983979593310074301486537017973460461278300587514468301043894574906886127642530475786889672304776052879927627556769456140664043088700743909632312483413393134504352834240399191134336344285483935856491230340093391784574980688823380828143810804684752914935441384845195613674104960646037368551517";
const Q: &str = "158741574437007245654463598139927898730476924736461654463975966787719309357536545869203069369466212089132653564188443272208127277664424448947476335413293018778018615899291704693105620242763173357203898195318179150836424196645745308205164116144020613415407736216097185962171301808761138424668335445923774195463";

fn dec(x: &BigInt) -> String {
    x.to_str_radix(10)
}
fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}
fn enc(ek: &EncryptionKey, m: &BigInt, r: &BigInt) -> BigInt {
    Paillier::encrypt_with_chosen_randomness(ek, RawPlaintext::from(m), &Randomness(r.clone())).0.into_owned()
}
fn verdict(r: Result<(), IncorrectProof>) -> &'static str {
    if r.is_ok() { "ok" } else { "incorrect" }
}

fn main() {
    let out_path = env::args().nth(1).unwrap_or_else(|| "reference_vectors.json".to_string());
    let kp = Keypair { p: BigInt::from_str_radix(P, 10).unwrap(), q: BigInt::from_str_radix(Q, 10).unwrap() };
    let (ek, dk) = kp.keys();
    let n = ek.n.clone();

    // ---- BigInt::to_bytes and the transcript hash built on it
    let two = BigInt::from(2);
    let probes: Vec<BigInt> = vec![
        BigInt::zero(), BigInt::one(), BigInt::from(255), BigInt::from(256), two.pow(32), two.pow(32) - BigInt::one(), two.pow(64),
        two.pow(255), two.pow(256) - BigInt::one(), two.pow(2047) + BigInt::from(12345), n.clone(), &n * &n,
    ];
    let to_bytes: Vec<Value> = probes.iter().map(|x| json!({"dec": dec(x), "hex": hex(&BigInt::to_bytes(x))})).collect();
    let digest_lists: Vec<Vec<BigInt>> = vec![
        vec![BigInt::zero()],
        vec![n.clone(), BigInt::zero(), BigInt::one()],
        vec![two.pow(64), BigInt::from(255), &n * &n],
        vec![BigInt::from_bytes(&[75, 90, 101, 110])],
        vec![BigInt::from_bytes(&[0, 0, 7, 9])], // from_bytes drops the leading zero bytes again on to_bytes
        (0u32..40).map(BigInt::from).collect(),
    ];
    let digests: Vec<Value> = digest_lists
        .iter()
        .map(|l| json!({"items": l.iter().map(dec).collect::<Vec<_>>(), "digest": dec(&compute_digest(l.iter()))}))
        .collect();

    // ---- serde of a bare BigInt and of the encryption key
    let bigint_serde: Vec<Value> = probes.iter().map(|x| json!({"dec": dec(x), "json": serde_json::to_string(x).unwrap()})).collect();
    let ek_json = serde_json::to_string(&ek).unwrap();

    // ---- Paillier::encrypt_with_chosen_randomness
    let mut encs = Vec::new();
    for (m, r) in vec![
        (BigInt::zero(), BigInt::one()), (BigInt::one(), BigInt::from(2)), (&n - BigInt::one(), &n - BigInt::one()),
        (BigInt::sample_below(&n), BigInt::sample_below(&n)), (BigInt::sample(256), BigInt::sample_below(&n)),
    ] {
        encs.push(json!({"m": dec(&m), "r": dec(&r), "c": dec(&enc(&ek, &m, &r))}));
    }

    // ---- NiCorrectKeyProof (deterministic)
    let salts: Vec<&'static [u8]> = vec![SALT_STRING, &[90, 101, 110, 32, 71, 111, 32, 88], &[0, 0, 1]];
    let ni: Vec<Value> = salts
        .iter()
        .map(|salt| {
            let proof = NiCorrectKeyProof::proof(&dk, Some(*salt));
            json!({"p": P, "q": Q, "salt_hex": hex(salt), "json": serde_json::to_string(&proof).unwrap(),
                   "verdict": verdict(proof.verify(&ek, salt)), "verdict_wrong_salt": verdict(proof.verify(&ek, &[1, 2, 3]))})
        })
        .collect();

    // ---- RangeProofNi: honest, and a secret far outside the range (range_proof_ni.rs:181-199)
    let mut rps = Vec::new();
    for honest in &[true, false] {
        let range = BigInt::sample(256);
        let r = BigInt::sample_below(&n);
        let x = if *honest {
            BigInt::sample_below(&range.div_floor(&BigInt::from(3)))
        } else {
            BigInt::sample_range(&(BigInt::from(100) * &range), &(BigInt::from(10000) * &range))
        };
        let c = enc(&ek, &x, &r);
        let proof = RangeProofNi::prove(&ek, &range, &c, &x, &r);
        rps.push(json!({"range": dec(&range), "x": dec(&x), "r": dec(&r), "ciphertext": dec(&c),
                        "json": serde_json::to_string(&proof).unwrap(), "verdict": verdict(proof.verify(&ek, &c))}));
    }

    // ---- the sigma protocols: an honest statement and the dishonest one of the reference's own tests
    let mut zero = Vec::new();
    for m in &[0u32, 1] {
        let r = BigInt::sample_below(&n);
        let c = enc(&ek, &BigInt::from(*m), &r);
        let st = ZeroStatement { ek: ek.clone(), c: c.clone() };
        let proof = ZeroProof::prove(&ZeroWitness { r }, &st);
        zero.push(json!({"c": dec(&c), "json": serde_json::to_string(&proof).unwrap(), "verdict": verdict(proof.verify(&st))}));
    }
    let mut ciphertext = Vec::new();
    for bad in &[0u32, 1] {
        let (x, r) = (BigInt::sample_below(&n), BigInt::sample_below(&n));
        let c = enc(&ek, &x, &r);
        let st = CiphertextStatement { ek: ek.clone(), c: c.clone() };
        let proof = CiphertextProof::prove(&CiphertextWitness { x, r: &r + BigInt::from(*bad) }, &st);
        ciphertext.push(json!({"c": dec(&c), "json": serde_json::to_string(&proof).unwrap(), "verdict": verdict(proof.verify(&st))}));
    }
    let mut mul = Vec::new();
    for bad in &[0u32, 1] {
        let (a, b) = (BigInt::sample_below(&n), BigInt::sample_below(&n));
        let c = (&a * &b + BigInt::from(*bad)) % &n;
        let (r_a, r_b, r_c) = (BigInt::sample_below(&n), BigInt::sample_below(&n), BigInt::sample_below(&n));
        let (e_a, e_b, e_c) = (enc(&ek, &a, &r_a), enc(&ek, &b, &r_b), enc(&ek, &c, &r_c));
        let st = MulStatement { ek: ek.clone(), e_a: e_a.clone(), e_b: e_b.clone(), e_c: e_c.clone() };
        let proof = MulProof::prove(&MulWitness { a, b, c, r_a, r_b, r_c }, &st);
        mul.push(json!({"e_a": dec(&e_a), "e_b": dec(&e_b), "e_c": dec(&e_c), "json": serde_json::to_string(&proof).unwrap(),
                        "verdict": verdict(proof.verify(&st))}));
    }
    let mut verlin = Vec::new();
    for bad in &[0u32, 1] {
        let (x, xp, xdp, r_x) = (BigInt::sample_below(&n), BigInt::sample_below(&n), BigInt::sample_below(&n), BigInt::sample_below(&n));
        let c = enc(&ek, &BigInt::sample_below(&n), &BigInt::sample_below(&n));
        let cp = enc(&ek, &BigInt::sample_below(&n), &BigInt::sample_below(&n));
        // phi_x = c^x c'^x' Enc(x'', r_x) mod nn (verlin_proof.rs:138-165)
        let phi_x = BigInt::mod_mul(
            &BigInt::mod_mul(&BigInt::mod_pow(&c, &x, &ek.nn), &BigInt::mod_pow(&cp, &xp, &ek.nn), &ek.nn),
            &enc(&ek, &xdp, &r_x),
            &ek.nn,
        );
        let st = VerlinStatement { ek: ek.clone(), c: c.clone(), c_prime: cp.clone(), phi_x: phi_x.clone() };
        let proof = VerlinProof::prove(&VerlinWitness { x: &x + BigInt::from(*bad), x_prime: xp, x_double_prime: xdp, r_x }, &st);
        verlin.push(json!({"c": dec(&c), "c_prime": dec(&cp), "phi_x": dec(&phi_x), "json": serde_json::to_string(&proof).unwrap(),
                           "verdict": verdict(proof.verify(&st))}));
    }
    let mut dlog = Vec::new();
    for bad in &[0u32, 1] {
        let g = BigInt::sample_below(&n);
        let secret = BigInt::sample(256);
        let ni_val = BigInt::mod_pow(&BigInt::mod_inv(&g, &n).unwrap(), &secret, &n);
        let st = DLogStatement { N: n.clone(), g: g.clone(), ni: ni_val.clone() };
        let proof = CompositeDLogProof::prove(&st, &(&secret + BigInt::from(*bad)));
        dlog.push(json!({"N": dec(&n), "g": dec(&g), "ni": dec(&ni_val), "json": serde_json::to_string(&proof).unwrap(),
                         "statement_json": serde_json::to_string(&st).unwrap(), "verdict": verdict(proof.verify(&st))}));
    }

    let doc = json!({
        "generator": "rust/gen_vectors (unmodified reference crate)",
        "crates": {"zk-paillier": "0.4.4", "curv-kzen": "0.10", "kzen-paillier": "0.4.3"},
        "key": {"p": P, "q": Q, "n": dec(&n)},
        "to_bytes": to_bytes,
        "compute_digest": digests,
        "bigint_serde": bigint_serde,
        "encryption_key": {"n": dec(&n), "json": ek_json},
        "paillier_enc": encs,
        "ni_correct_key": ni,
        "range_proof_ni": rps,
        "zero": zero,
        "ciphertext": ciphertext,
        "mul": mul,
        "verlin": verlin,
        "dlog": dlog,
    });
    fs::write(&out_path, serde_json::to_string_pretty(&doc).unwrap()).expect("cannot write the vectors file");
    println!("wrote {}", out_path);
}
