"""Generate the committed key fixtures tests/golden/keys.json: RSA-style moduli n = p*q with |n| in
{1024, 2048, 3072, 4096} (SURVEY.md section 8d asks for offline-generated, committed keys for the configs
that do not use the reference's fixed 2048-bit test key).  Primes come from OpenSSL through the
`cryptography` package; gcd(n, phi(n)) = 1 is checked so NiCorrectKeyProof statements are honest."""
import json, math, os, sys
from cryptography.hazmat.primitives.asymmetric import rsa

counts = {1024: 4, 2048: 4, 3072: 16, 4096: 4}
out = {}
for bits, cnt in counts.items():
    ks = []
    while len(ks) < cnt:
        k = rsa.generate_private_key(65537, bits).private_numbers()
        p, q = k.p, k.q
        n = p * q
        if n.bit_length() != bits or math.gcd(n, (p - 1) * (q - 1)) != 1:
            continue
        ks.append({"p": str(p), "q": str(q)})
    out[str(bits)] = ks
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "keys.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path)
