// The reference's own range-proof test (range_proof_ni.rs:130-179), written against the C++ mirror of its interface:
// generate a Paillier key, encrypt a secret below range/3, prove that it lies in the range, serialise the proof with the
// reference's serde wire format, parse it back and verify it -- every modexp on the B200.
//   make examples && ./examples/range_proof_ni          (needs a CUDA device: the engine has no CPU fallback)
#include <cstdio>

#include "../zk-paillier_b200/host/zkproofs.hpp"

using namespace zkproofs;

int main() {
  try {
    Engine eng(0);
    const ByteSource rng = os_rng();
    const DecryptionKey dk = Paillier::keypair_with_modulus_size(eng, 2048, rng);      // Paillier::keypair()
    const EncryptionKey ek(dk.p * dk.q);
    const BigInt range = BigInt::sample(rng, 256) | BigInt(1).shl(255);                          // range_proof_ni.rs:133
    const BigInt secret_x = BigInt::sample_below(rng, range / BigInt(3));              // :134
    auto enc = Paillier::encrypt_batch(eng, ek, {secret_x}, rng);                      // :135-141 (c, r)
    const RangeProofNi proof = RangeProofNi::prove(eng, ek, range, enc.first[0], secret_x, enc.second[0], rng);  // :143
    const std::string wire = proof.to_json();
    RangeProofNi::from_json(wire).verify(eng, ek, enc.first[0]);                       // :144 -> Ok(())
    std::printf("RangeProofNi: %zu bytes of JSON, verified\n", wire.size());

    const NiCorrectKeyProof ck = NiCorrectKeyProof::proof(eng, dk);                    // correct_key_ni.rs:125-131
    ck.verify(eng, ek, SALT_STRING, sizeof(SALT_STRING));
    std::printf("NiCorrectKeyProof: verified\n");

    const BigInt big_x = range * BigInt(1000);                                         // :181-199: x far outside the range
    auto enc2 = Paillier::encrypt_batch(eng, ek, {big_x}, rng);
    try {
      RangeProofNi::prove(eng, ek, range, enc2.first[0], big_x, enc2.second[0], rng).verify(eng, ek, enc2.first[0]);
      std::printf("out-of-range statement accepted: BUG\n");
      return 1;
    } catch (const IncorrectProof&) {
      std::printf("out-of-range statement rejected: Err(IncorrectProof)\n");
    }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 2;
  }
}
