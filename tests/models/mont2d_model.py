"""Executable model of K1m (csrc/modexp2m.cu): Montgomery multiplication modulo n^2 in two-digit base-n form.

An element x of Z_{n^2} is the pair (X0, X1), x = X0 + X1 n, 0 <= X0, X1 < n; W = 2^(32 S) > n.
The CIOS Montgomery multiplication of X0 and Y0 modulo n yields, besides Z0 = X0 Y0 / W mod n, the quotient
digits q it added:          X0 Y0 = Z0' W - q n      (an identity between integers; Z0' < 2n)
so modulo n^2
    x y / W = Z0 + ((X0 Y1 + X1 Y0 - q_eff) / W mod n) n,     q_eff = q - delta W (delta: Z0 = Z0' - n was taken)
i.e. the second digit is ONE more Montgomery reduction modulo n whose accumulator starts at a non-negative
representative of -q_eff:  init = W + (delta ? 0 : K_lo) - q,  K_lo = -W mod n  (init < W + n, so the result is
< 2n + 2: two conditional subtractions).  No true quotient, no Barrett, only half-width CIOS rows:
a squaring costs 4 S^2 limb products and a multiplication 5-6 S^2, against 8 S^2 for CIOS modulo n^2 (K1).

Part 1 checks the algebra with Python integers (including unreduced bases, 2047-bit n, the Paillier epilogue and
the final assembly).  Part 2 re-runs the limb-level split-accumulator model of mp_coop.cuh with a non-zero initial
accumulator and an overflow limb in the top lane, which is what phase 2 feeds it."""
import os
import random
import runpy

HERE = os.path.dirname(os.path.abspath(__file__))


class Key:
    def __init__(self, n, S):
        self.n, self.S = n, S
        self.W = 1 << (32 * S)
        assert n % 2 == 1 and n < self.W - 4
        self.nprime = (-pow(n, -1, self.W)) % self.W
        self.klo = (-self.W) % n
        c = (self.W * self.W) % (n * n)  # host: 64 S pair doublings from (1, 0)
        self.C = (c % n, c // n)


def mont_q(k, a, b, init=0):
    """CIOS: returns (z, q) with z W = init + a b + q n exactly, 0 <= q < W"""
    t = init + a * b
    q = (t * k.nprime) % k.W
    z, rem = divmod(t + q * k.n, k.W)
    assert rem == 0
    return z, q


def phase1(k, a, b):
    assert a * b < k.W * k.n
    z, q = mont_q(k, a, b)
    assert z < 2 * k.n
    delta = 1 if z >= k.n else 0
    return z - delta * k.n, q, delta


def init_from_q(k, q, delta):
    m = 0 if delta else k.klo
    low = (m - q) % k.W
    top = 0 if m < q else 1
    init = low + top * k.W
    assert init == k.W + m - q and 0 <= init < k.W + k.n
    return init


def phase2(k, a, b, init):
    assert a * b < k.W * k.n
    z, _ = mont_q(k, a, b, init)
    assert z < 2 * k.n + 2
    if z >= k.n:
        z -= k.n
    if z >= k.n:
        z -= k.n
    assert z < k.n
    return z


def mul2d(k, X, Y):
    z0, q, d = phase1(k, X[0], Y[0])
    # phase 2 is ONE reduction of init + X0 Y1 + X1 Y0 (cios_step2: two product rows and one q n row per step)
    init = init_from_q(k, q, d)
    assert X[0] * Y[1] < k.W * k.n and X[1] * Y[0] < k.W * k.n
    t = init + X[0] * Y[1] + X[1] * Y[0]
    z1 = (t + ((t * k.nprime) % k.W) * k.n) // k.W
    assert z1 < 3 * k.n + 2
    for _ in range(3):
        if z1 >= k.n:
            z1 -= k.n
    assert z1 < k.n
    return (z0, z1)


def sqr2d(k, X):
    z0, q, d = phase1(k, X[0], X[0])
    b = (2 * X[1]) % k.n  # mod_double
    z1 = phase2(k, X[0], b, init_from_q(k, q, d))
    return (z0, z1)


def val(k, X):
    return (X[0] + X[1] * k.n) % (k.n * k.n)


def enc(k, r, m, e):
    """(1 + m n) r^e mod n^2 the way the kernel does it (binary ladder here; the kernel uses its window schedule)"""
    nn = k.n * k.n
    x = mul2d(k, (r, 0), k.C)  # r may be >= n (any value < W): only a b < W n is needed
    assert val(k, x) == (r * k.W) % nn
    acc = x
    for bit in bin(e)[3:]:
        acc = sqr2d(k, acc)
        if bit == "1":
            acc = mul2d(k, acc, x)
    assert val(k, acc) == (pow(r, e, nn) * k.W) % nn
    z = mul2d(k, acc, (1, m))  # leaves Montgomery form and applies the Paillier factor 1 + m n in one product
    c = z[0] + z[1] * k.n  # final assembly: plain product rows with Z0 as the initial accumulator
    assert c < nn
    return c


def part1():
    random.seed(11)
    for S, bits, trials in [(2, 64, 200), (2, 63, 200), (2, 40, 100), (64, 2048, 3), (64, 2047, 3), (96, 3072, 1)]:
        for it in range(trials):
            n = random.getrandbits(bits) | 1 | (1 << (bits - 1))
            k = Key(n, S)
            nn = n * n
            for _ in range(4):
                x, y = random.randrange(nn), random.randrange(nn)
                X, Y = (x % n, x // n), (y % n, y // n)
                winv = pow(k.W, -1, nn)
                assert val(k, mul2d(k, X, Y)) == (x * y * winv) % nn
                assert val(k, sqr2d(k, X)) == (x * x * winv) % nn
            # edge operands
            for x in (0, 1, n - 1, n, nn - 1):
                X = (x % n, x // n)
                assert val(k, sqr2d(k, X)) == (x * x * pow(k.W, -1, nn)) % nn
            r = random.randrange(k.W) if it % 2 else random.randrange(n)
            m = random.randrange(n)
            e = n if S <= 2 else random.getrandbits(24) | 1
            assert enc(k, r, m, e) == ((1 + m * n) * pow(r, e, nn)) % nn
            assert enc(k, r, 0, e) == pow(r, e, nn)
        print("two-digit Montgomery", S, bits, "ok")


def part2():
    """limb-level: split-accumulator CIOS rows with a non-zero initial accumulator (top-lane overflow limb set)"""
    ns = runpy.run_path(os.path.join(HERE, "cios_model.py"), run_name="model")
    model = ns["model_montmul_init"]
    random.seed(5)
    for (T, L) in [(8, 8), (4, 8), (8, 12), (16, 8)]:
        S = T * L
        W = 1 << (32 * S)
        for it in range(12):
            bits = 32 * S - (it % 3)
            n = random.getrandbits(bits) | 1 | (1 << (bits - 1))
            if it == 4:
                n = W - 5
            a = random.randrange(n) if it % 2 else random.randrange(W)
            b = random.randrange(n)
            init = random.randrange(W + n) if it else W + n - 1
            model(a, b, n, T, L, init)
            model(random.randrange(n), b, n, T, L, init, random.randrange(n), random.randrange(n))  # fused multiply: two rows
            if it == 2:
                model(n - 1, n - 1, n, T, L, W + n - 1, n - 1, n - 1)
        print("cios rows with initial accumulator", T, L, "ok")


if __name__ == "__main__":
    part1()
    part2()
