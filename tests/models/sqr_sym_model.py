"""Executable model of the symmetric squaring of K1m (csrc/mp_coop.cuh: sqr_product / mont_redc_x): X0^2 / W mod n with the
product computed ONCE per unordered pair of lane blocks, then a reduction-only Montgomery loop.

Lane g of a group of T holds the block A_g (L limbs) of x.  x^2 = sum_g A_g^2 B^(2g) + 2 sum_{g<h} A_g A_h B^(g+h), B = 2^(32 L).
Round r = 0 .. T/2: lane g multiplies its own block by the block of lane (g + r) mod T (in-lane product scanning, L^2 limb
products).  Without wrap-around that is the pair of difference r at block position p = 2g + r; with wrap-around it is the
pair of difference T - r at p = 2g + r - T; r = T/2 produces every pair twice (weight 1 instead of 2).  Position p is
owned by lane p >> 1, slot p & 1 = r & 1: the slot is the same for the whole round, and a destination lane has at most two
sources, (d - h) and (d - h + T/2) with h = r >> 1.  (T/2 + 1) L^2 limb products per lane instead of T L^2.
The 2 S-limb square then sits two blocks per lane; it is made canonical, transposed into the low half (block g in lane g,
the initial accumulator) and the high half (fed into the top lane one limb per reduction row), and reduced by S rows of
q n only: the quotient digits and the final-subtraction flag come out exactly as from the CIOS multiplication."""
import random

M32 = 0xffffffff


def blocks(x, T, L):
    return [[(x >> (32 * (g * L + j))) & M32 for j in range(L)] for g in range(T)]


def val(arr):
    v = 0
    for k, limb in enumerate(arr):
        assert 0 <= limb <= M32
        v |= limb << (32 * k)
    return v


def put(v, cnt):
    out = [(v >> (32 * k)) & M32 for k in range(cnt)]
    assert v >> (32 * cnt) == 0, "overflow"
    return out


def sqr_product(x, T, L):
    """-> (plo, phi): per-lane L-limb blocks of the low and the high half of x^2 (canonical)."""
    S = T * L
    A = blocks(x, T, L)
    acc = [[[0] * (2 * L + 1) for _ in range(2)] for _ in range(T)]
    for r in range(T // 2 + 1):
        h, slot = r >> 1, r & 1
        weight = 1 if r in (0, T // 2) else 2
        prod = [put(val(A[g]) * val(A[(g + r) % T]), 2 * L) for g in range(T)]
        if r == 0:
            for g in range(T):
                acc[g][0] = put(val(acc[g][0]) + val(prod[g]), 2 * L + 1)
            continue
        for d in range(T):
            for src, wrapped in ((d - h, False), (d - h + T // 2, True)):
                if not 0 <= src < T:
                    continue
                # the source's own view of where its product goes
                nonwrapped = src + r <= T - 1
                dest = src + h if nonwrapped else src + h - T // 2
                valid = dest == d and (nonwrapped != wrapped)
                if valid:
                    acc[d][slot] = put(val(acc[d][slot]) + weight * val(prod[src]), 2 * L + 1)
    # pair layout: lane d holds limbs [2 L d, 2 L (d + 1)) of the square; the rest moves one lane up
    own, ovf = [], []
    for d in range(T):
        v = val(acc[d][0]) + (val(acc[d][1]) << (32 * L))
        own.append(v & ((1 << (64 * L)) - 1))
        ovf.append(v >> (64 * L))
        assert ovf[-1] < 1 << (32 * (L + 1))
    lanes = []
    carry_gen = []
    for d in range(T):
        v = own[d] + (ovf[d - 1] if d else 0)
        carry_gen.append(v >> (64 * L))          # 0 or 1: resolved across lanes by generate / propagate ballots
        assert carry_gen[-1] <= 1
        lanes.append(v & ((1 << (64 * L)) - 1))
    cin = 0
    for d in range(T):
        v = lanes[d] + cin
        cin = carry_gen[d] | (v >> (64 * L))
        lanes[d] = v & ((1 << (64 * L)) - 1)
    assert cin == 0 and ovf[T - 1] == 0
    assert sum(l << (64 * L * d) for d, l in enumerate(lanes)) == x * x
    V = [put(l, 2 * L) for l in lanes]
    plo = [[V[d >> 1][(d & 1) * L + j] for j in range(L)] for d in range(T)]
    phi = [[V[T // 2 + (d >> 1)][(d & 1) * L + j] for j in range(L)] for d in range(T)]
    assert val(sum(plo, [])) + (val(sum(phi, [])) << (32 * S)) == x * x
    return plo, phi


def mont_redc(plo, phi, n, T, L):
    """Reduction-only rows on the split accumulator of mp_coop.cuh.  -> (value < 2n, q digits)."""
    S = T * L
    N = blocks(n, T, L)
    n0inv = (-pow(n, -1, 1 << 32)) & M32
    E = [plo[g] + [0, 0] for g in range(T)]
    O = [[0] * (L + 2) for _ in range(T)]
    qs = []

    def v2(arr, lo, cnt):
        return val(arr[lo:lo + cnt])

    def step(X, Y, feed):
        ins = [Y[g + 1][0] if g < T - 1 else feed for g in range(T)]
        Zs = []
        for g in range(T):
            y, x = Y[g], X[g]
            y[L:L + 2] = put(v2(y, L, 2) + ins[g], 2)
            s = x[0] + y[1]
            x[0] = s & M32
            Zs.append(s >> 32)                      # carry into the odd chain
        q = (X[0][0] * n0inv) & M32
        qs.append(q)
        for g in range(T):
            y, x = Y[g], X[g]
            c = Zs[g]
            Z = [0] * (L + 2)
            for j in range(0, L, 2):                # Z = n_odd * q + (Y >> 2 limbs) + carry
                t = N[g][j + 1] * q + v2(y, j + 2, 2) + c
                Z[j], Z[j + 1], c = t & M32, (t >> 32) & M32, t >> 64
            Z[L], Z[L + 1] = c, 0
            assert c <= M32
            v = val(x)
            for j in range(0, L, 2):
                v += (N[g][j] * q) << (32 * j)
            x[:] = put(v, L + 2)
            Y[g][:] = Z

    flat_hi = sum(phi, [])
    prev = 0
    for i in range(0, S, 2):
        step(E, O, prev)                            # row i takes high limb i - 1 (it sits at limb S - 1 after the shift)
        step(O, E, flat_hi[i])
        prev = flat_hi[i + 1]
    tot = 0
    for g in range(T):
        inn = O[g + 1][0] if g < T - 1 else prev
        O[g][L:L + 2] = put(v2(O[g], L, 2) + inn, 2)
        E[g][:] = put(val(E[g]) + v2(O[g], 1, L + 1), L + 2)
        tot += val(E[g]) << (32 * g * L)
    assert O[0][0] == 0
    q = val(qs)
    return tot, q


def check(x, n, T, L):
    S = T * L
    W = 1 << (32 * S)
    plo, phi = sqr_product(x, T, L)
    z, q = mont_redc(plo, phi, n, T, L)
    assert z * W == x * x + q * n and z < 2 * n and q < W      # the identity the second digit relies on
    nprime = (-pow(n, -1, W)) % W
    assert q == (x * x * nprime) % W                            # the same digits as the CIOS multiplication produces
    return z


if __name__ == "__main__":
    random.seed(7)
    for T, L in [(8, 8), (4, 8), (8, 12), (16, 8), (8, 2), (4, 2), (2, 4)]:
        S = T * L
        for it in range(12):
            n = random.getrandbits(32 * S) | 1 | (1 << (32 * S - 1)) if it % 2 == 0 else (random.getrandbits(32 * S - random.randint(0, 40)) | 1)
            if it == 5:
                n = (1 << (32 * S)) - 5
            x = random.randrange(n)
            if it == 3:
                x = n - 1
            if it == 4:
                x = (1 << (32 * S)) - 1 if n > (1 << (32 * S - 1)) else n - 1   # any S-limb value with x^2 < n W
                if x * x >= n << (32 * S):
                    x = n - 1
            if it == 6:
                x = 0
            if it == 7:
                x = 1
            check(x, n, T, L)
        print(T, L, "ok")
