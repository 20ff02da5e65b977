"""CPU suite: the Rust face (rust/, source only - no cargo in this image) against the C header it binds.

The shim cannot be compiled here, so it is DIFFED: rust/src/ffi.rs is generated from include/zkp_b200.h and must be
current; both files are then parsed independently and compared symbol by symbol (names, arity, pointer constness,
return type); every `ffi::zkp_*` call in the shim must pass exactly as many arguments as the header declares; and every
public proof of the reference (src/zkproofs/mod.rs:29-43) must have its struct, `prove` / `verify` with the
reference's names, and the `*_batch` forms."""
import os
import re
import subprocess
import sys

from util import ROOT

RUST = os.path.join(ROOT, "rust", "src")


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", "", s, flags=re.S)


def _header_protos():
    src = _strip_c_comments(open(os.path.join(ROOT, "include", "zkp_b200.h")).read())
    out = {}
    for m in re.finditer(r"\b(const char\*|long long|int|void)\s+(zkp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        args = " ".join(m.group(3).split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        kinds = []
        for a in params:
            stars = a.count("*")
            kinds.append(("ptr" + str(stars), "const" if a.startswith("const ") else "mut") if stars else ("val", ""))
        out[m.group(2)] = (m.group(1), kinds)
    return out


def _rust_decls():
    src = open(os.path.join(RUST, "ffi.rs")).read()
    out = {}
    for m in re.finditer(r"pub fn (zkp_[a-z0-9_]+)\((.*?)\)( -> ([^;]+))?;", src, flags=re.S):
        params = [p.strip() for p in m.group(2).split(",") if p.strip()]
        kinds = []
        for p in params:
            ty = p.split(":", 1)[1].strip()
            stars = ty.count("*")
            if stars:
                kinds.append(("ptr" + str(stars), "const" if ty.startswith("*const") else "mut"))
            else:
                kinds.append(("val", ""))
        out[m.group(1)] = ((m.group(4) or "").strip(), kinds)
    return out


def test_ffi_rs_is_generated_from_the_header_and_current():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_rust_ffi.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_ffi_rs_matches_the_header_symbol_by_symbol():
    hdr, rs = _header_protos(), _rust_decls()
    assert len(hdr) >= 50
    assert sorted(hdr) == sorted(rs), f"missing in ffi.rs: {sorted(set(hdr) - set(rs))}; extra: {sorted(set(rs) - set(hdr))}"
    ret_map = {"int": "c_int", "void": "", "const char*": "*const c_char", "long long": "c_longlong"}
    for name, (ret, kinds) in hdr.items():
        rret, rkinds = rs[name]
        assert ret_map[ret] == rret, (name, ret, rret)
        assert len(kinds) == len(rkinds), f"{name}: {len(kinds)} parameters in the header, {len(rkinds)} in ffi.rs"
        for i, (a, b) in enumerate(zip(kinds, rkinds)):
            assert a == b, f"{name} parameter {i}: header {a}, ffi.rs {b}"


def _rust_sources():
    for base, _, files in os.walk(RUST):
        for f in files:
            if f.endswith(".rs") and f != "ffi.rs":
                yield os.path.join(base, f)
    ex = os.path.join(ROOT, "rust", "examples")
    for f in os.listdir(ex):
        if f.endswith(".rs"):
            yield os.path.join(ex, f)


def _call_args(src, start):
    """arguments of the call whose '(' is at src[start]: top-level commas only"""
    depth, i, args, cur = 0, start, [], ""
    while i < len(src):
        ch = src[i]
        if ch in "([{":
            depth += 1
            if depth > 1:
                cur += ch
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                if cur.strip():
                    args.append(cur.strip())
                return args
            cur += ch
        elif ch == "," and depth == 1:
            args.append(cur.strip())
            cur = ""
        else:
            cur += ch
        i += 1
    raise AssertionError("unbalanced call")


def test_every_ffi_call_in_the_shim_has_the_declared_arity():
    hdr = _header_protos()
    calls = 0
    used = set()
    for path in _rust_sources():
        src = re.sub(r"//[^\n]*", "", open(path).read())
        for m in re.finditer(r"ffi::(zkp_[a-z0-9_]+)\s*\(", src):
            name = m.group(1)
            assert name in hdr, f"{path}: {name} is not declared in include/zkp_b200.h"
            args = _call_args(src, m.end() - 1)
            assert len(args) == len(hdr[name][1]), f"{os.path.relpath(path, ROOT)}: {name} called with {len(args)} arguments, header declares {len(hdr[name][1])}"
            calls += 1
            used.add(name)
    assert calls >= 20
    # the entry points of the six north-star proofs and the remaining public proofs are all reached from the shim
    for name in ("zkp_rangeproof_ni_prove", "zkp_rangeproof_ni_verify", "zkp_correct_key_ni_verify", "zkp_correct_key_ni_rho", "zkp_zero_prove",
                 "zkp_zero_verify", "zkp_ciphertext_prove", "zkp_ciphertext_verify", "zkp_mul_prove", "zkp_mul_verify", "zkp_verlin_prove",
                 "zkp_verlin_verify", "zkp_correct_message_prove", "zkp_correct_message_verify", "zkp_dlog_prove", "zkp_dlog_verify",
                 "zkp_verify_opening", "zkp_sha256_transcript", "zkp_modexp_var", "zkp_set_key"):
        assert name in used, f"{name} is never called from rust/src"


# reference file -> (struct names, associated functions) the shim must define (reference src/zkproofs/*.rs)
SURFACE = {
    "range_proof_ni.rs": (["RangeProofNi", "EncryptedPairs", "Proof", "Response"], ["prove", "verify", "verify_self", "prove_batch", "verify_batch"]),
    "correct_key_ni.rs": (["NiCorrectKeyProof"], ["proof", "verify", "proof_batch", "verify_batch"]),
    "zero_enc_proof.rs": (["ZeroProof", "ZeroWitness", "ZeroStatement"], ["prove", "verify", "prove_batch", "verify_batch"]),
    "correct_ciphertext.rs": (["CiphertextProof", "CiphertextWitness", "CiphertextStatement"], ["prove", "verify", "prove_batch", "verify_batch"]),
    "multiplication_proof.rs": (["MulProof", "MulWitness", "MulStatement"], ["prove", "verify", "prove_batch", "verify_batch"]),
    "verlin_proof.rs": (["VerlinProof", "VerlinWitness", "VerlinStatement"], ["prove", "verify", "prove_batch", "verify_batch"]),
    "correct_message.rs": (["CorrectMessageProof"], ["prove", "verify", "prove_batch", "verify_batch"]),
    "wi_dlog_proof.rs": (["CompositeDLogProof", "DLogStatement"], ["prove", "verify", "prove_batch", "verify_batch"]),
}


def test_shim_has_the_reference_public_surface():
    for fname, (structs, fns) in SURFACE.items():
        src = open(os.path.join(RUST, "zkproofs", fname)).read()
        for s in structs:
            assert re.search(rf"pub (struct|enum) {s}\b", src), f"{fname}: {s} missing"
        for fn in fns:
            assert re.search(rf"pub fn {fn}\b", src), f"{fname}: fn {fn} missing"
    mod = open(os.path.join(RUST, "zkproofs", "mod.rs")).read()
    for name in ("RangeProofNi", "NiCorrectKeyProof", "SALT_STRING", "CorrectMessageProof", "CorrectOpening", "IncorrectProof", "compute_digest"):
        assert name in mod
    assert "pub trait CorrectOpening" in open(os.path.join(RUST, "zkproofs", "correct_opening.rs")).read()
    # one engine context per thread, never one per call
    eng = open(os.path.join(RUST, "engine.rs")).read()
    assert "thread_local!" in eng
    for path in _rust_sources():
        if path.endswith("engine.rs"):
            continue
        assert "zkp_ctx_create" not in open(path).read(), f"{path} creates a context per call"


def test_wire_structs_keep_the_reference_field_order_and_serde_attributes():
    """Field names and order are the wire schema (serde derives follow declaration order)."""
    want = {
        "zero_enc_proof.rs": {"ZeroProof": ["z", "a"], "ZeroWitness": ["r"], "ZeroStatement": ["ek", "c"]},
        "correct_ciphertext.rs": {"CiphertextProof": ["z1", "z2", "c_prime"], "CiphertextWitness": ["x", "r"], "CiphertextStatement": ["ek", "c"]},
        "multiplication_proof.rs": {"MulProof": ["f", "z1", "z2", "e_d", "e_db"], "MulWitness": ["a", "b", "c", "r_a", "r_b", "r_c"],
                                    "MulStatement": ["ek", "e_a", "e_b", "e_c"]},
        "verlin_proof.rs": {"VerlinProof": ["phi_a", "z", "z_prime", "z_double_prime", "r_z"],
                            "VerlinWitness": ["x", "x_prime", "x_double_prime", "r_x"], "VerlinStatement": ["ek", "c", "c_prime", "phi_x"]},
        "range_proof_ni.rs": {"RangeProofNi": ["ek", "range", "ciphertext", "encrypted_pairs", "proof", "error_factor"], "EncryptedPairs": ["c1", "c2"]},
        "correct_key_ni.rs": {"NiCorrectKeyProof": ["sigma_vec"]},
        "wi_dlog_proof.rs": {"CompositeDLogProof": ["x", "y"], "DLogStatement": ["N", "g", "ni"]},
    }
    for fname, structs in want.items():
        src = open(os.path.join(RUST, "zkproofs", fname)).read()
        for name, fields in structs.items():
            body = re.search(rf"pub struct {name} \{{(.*?)\n\}}", src, flags=re.S).group(1)
            got = re.findall(r"^\s*(?:pub )?([A-Za-z_][A-Za-z0-9_]*):", body, flags=re.M)
            assert got == fields, (fname, name, got)
    rp = open(os.path.join(RUST, "zkproofs", "range_proof_ni.rs")).read()
    assert rp.count('#[serde(with = "crate::serialize::vecbigint")]') == 2 and rp.count('#[serde(with = "crate::serialize::bigint")]') == 6
    assert 'serde(with = "crate::serialize::vecbigint")' in open(os.path.join(RUST, "zkproofs", "correct_key_ni.rs")).read()
