"""Model of K4w's block fetch and schedule split (rangeproof.cu, sha256_transcript_warp_kernel), word by word.

A transcript is the concatenation of to_bytes() of its items (big-endian, minimal length, zero -> one 0x00 byte); the
items sit in memory as fixed-width little-endian 32-bit limb rows.  The kernel lays the byte offsets out once (off[]),
then lane j of a pass fetches block base + j on its own:
  * a word whose four bytes lie inside one item is bits [sh, sh + 32) of that item's limbs, sh = 8 * (bytes of the item
    below the word) - one funnel shift of two adjacent limbs;
  * a word that straddles items, the 0x80 terminator and the zero padding go byte by byte;
  * the last block carries the bit length in words 14 and 15;
each lane expands its block's message schedule W[0..63] (+ K), and the blocks are compressed in order with W taken from
the owning lane.  The model runs exactly those steps on Python integers and compares the digest with hashlib."""
import hashlib
import random

MASK = 0xFFFFFFFF
K = [
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3,
    0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2,
]


def rotr(x, n):
    return ((x >> n) | (x << (32 - n))) & MASK


def funnelshift_r(lo, hi, sh):  # low 32 bits of (hi:lo) >> sh, 0 <= sh < 32
    return (((hi << 32) | lo) >> sh) & MASK


def limbs_of(v, limbs):
    return [(v >> (32 * i)) & MASK for i in range(limbs)]


def item_len(row):  # minimal big-endian length; zero -> 1 (the kernel walks down from the top limb)
    top = len(row) - 1
    while top >= 0 and row[top] == 0:
        top -= 1
    if top < 0:
        return 1
    clz = 32 - row[top].bit_length()
    return 4 * top + 4 - (clz >> 3)


def fetch(rows, off, total, nblocks, blk):
    """the 16 words of block blk, as lane (blk mod 32) computes them"""
    m = [0] * 16
    p0 = blk * 64
    it, end = 0, 0
    if blk < nblocks and p0 < total:
        lo, hi = 0, len(rows) - 1
        while lo < hi:  # largest i with off[i] <= p0
            mid = (lo + hi + 1) >> 1
            if off[mid] <= p0:
                lo = mid
            else:
                hi = mid - 1
        it, end = lo, off[lo + 1]
    for k in range(16):
        pos = p0 + 4 * k
        word = 0
        if blk < nblocks:
            if pos < total:
                while pos >= end:
                    it += 1
                    end = off[it + 1]
                if pos + 4 <= end:
                    sh = 8 * (end - pos - 4)
                    lo32 = rows[it][sh >> 5]
                    hi32 = rows[it][(sh >> 5) + 1] if (sh & 31) else 0  # never read past the row when the word is limb-aligned
                    word = funnelshift_r(lo32, hi32, sh & 31)
                else:
                    for bq in range(4):
                        pb = pos + bq
                        if pb < total:
                            while pb >= end:
                                it += 1
                                end = off[it + 1]
                            le = end - 1 - pb
                            byte = (rows[it][le >> 2] >> (8 * (le & 3))) & 0xFF
                        else:
                            byte = 0x80 if pb == total else 0
                        word = ((word << 8) | byte) & MASK
            elif pos == total:
                word = 0x80000000
        m[k] = word
    if blk == nblocks - 1:
        bits = total * 8
        m[14], m[15] = (bits >> 32) & MASK, bits & MASK
    return m


def digest(items, limbs):
    rows = [limbs_of(v, limbs) for v in items]
    off, run = [], 0
    for r in rows:
        off.append(run)
        run += item_len(r)
    off.append(run)
    total = run
    nblocks = (total + 9 + 63) // 64
    h = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
    for base in range(0, nblocks, 32):
        kw = []
        for lane in range(32):  # every lane: fetch its block, expand W + K
            m = fetch(rows, off, total, nblocks, base + lane)
            w = list(m)
            out = [(w[i] + K[i]) & MASK for i in range(16)]
            for i in range(16, 64):
                w15, w2 = w[(i + 1) & 15], w[(i + 14) & 15]
                s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3)
                s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10)
                wi = (w[i & 15] + s0 + w[(i + 9) & 15] + s1) & MASK
                w[i & 15] = wi
                out.append((wi + K[i]) & MASK)
            kw.append(out)
        for j in range(min(32, nblocks - base)):  # the rounds, blocks in order, words from lane j
            a, b, c, d, e, f, g, hh = h
            for i in range(64):
                x = kw[j][i]
                t1 = (hh + x + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g & MASK))) & MASK
                t2 = ((rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c))) & MASK
                hh, g, f, e, d, c, b, a = g, f, e, (d + t1) & MASK, c, b, a, (t1 + t2) & MASK
            h = [(x + y) & MASK for x, y in zip(h, (a, b, c, d, e, f, g, hh))]
    return b"".join(x.to_bytes(4, "big") for x in h)


def to_bytes(v):
    return v.to_bytes(max(1, (v.bit_length() + 7) // 8), "big")


if __name__ == "__main__":
    rng = random.Random(11)
    cases = []
    for nbytes in (1, 3, 4, 5, 54, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 129, 2047, 2048, 2049):  # one item around every padding boundary
        cases.append(([int.from_bytes(b"\x01" + bytes(rng.randrange(256) for _ in range(nbytes - 1)), "big")], (nbytes + 3) // 4 + 1))
    cases.append(([0], 4))                                                    # the empty-looking transcript: one 0x00 byte
    cases.append(([0, 0, 0, 0, 0], 4))
    cases.append(([(i * 7) % 3 << (8 * (i % 3)) for i in range(300)], 4))     # tiny items: words spanning up to four items
    for _ in range(6):                                                        # mixed widths, leading zero bytes and limbs, > 32 blocks
        items = []
        for k in range(rng.randrange(1, 12)):
            kind = rng.randrange(5)
            v = [0, rng.getrandbits(rng.randrange(1, 4096)), rng.getrandbits(4096) >> (8 * rng.randrange(9)), (1 << 4096) - 1, 1 << (8 * rng.randrange(512))][kind]
            items.append(v)
        cases.append((items, 128))
    for items, limbs in cases:
        want = hashlib.sha256(b"".join(to_bytes(v) for v in items)).digest()
        assert digest(items, limbs) == want, (len(items), limbs)
    print(f"sha_fetch_model: ok ({len(cases)} transcripts)")
