"""Limb-level model of the PAIR-ROW CIOS of mp_coop.cuh (Mp::cios_pair): two multiplier limbs and a two-limb quotient per
step, so that the serial quotient chain (broadcast, multiply, first products) runs once per TWO rows.

Per lane (L even limbs; lane g owns limbs [g L, g L + L)):
    E[w] sits at lane-local limb w      (w = 0 .. L + 2; E[L..] hang over into the next lane's range)
    O[w] sits at lane-local limb w + 1  (w = 0 .. L + 1)
A step computes  acc = (acc + a (b0 + b1 B) + (q0 + q1 B) n) / B^2,  B = 2^32, with
    (q0 + q1 B) = (low 64 bits of acc + a (b0 + b1 B)) * n' mod B^2,   n' = -n^-1 mod B^2
which are exactly the two quotient digits two single CIOS rows would have produced.  Products go to the array whose pairs are
aligned with their position:  a[2m] b0 -> E pair m,  a[2m+1] b0 -> O pair m,  a[2m] b1 -> O pair m,  a[2m+1] b1 -> E pair m + 1.
Dividing by B^2 is a shift by one aligned pair in BOTH arrays (done by the first chain of each array writing to the shifted
destination), plus one bridge: limb 2 of O (O[1]) lands on limb 0 and is added into E there, together with the carry of the
vanishing limb 1; the bridge's own carry enters the first odd chain of the next step.  The two lowest limbs of every lane but
lane 0 travel to the lane below (2 shuffles per step)."""
import random

M32 = 0xFFFFFFFF
B = 1 << 32


def model_montmul_pair(a, b, n, T, L, init=0, a2=None, b2=None):
    S = T * L
    R = 1 << (32 * S)
    npr = (-pow(n, -1, 1 << 64)) % (1 << 64)
    np0, np1 = npr & M32, npr >> 32
    limbs = lambda x: [[(x >> (32 * (g * L + j))) & M32 for j in range(L)] for g in range(T)]
    A, Bm, N = limbs(a), limbs(b), limbs(n)
    A2 = limbs(a2) if a2 is not None else None
    B2 = limbs(b2) if b2 is not None else None
    assert init >> (32 * (S + 2)) == 0
    E = [[(init >> (32 * (g * L + k))) & M32 if (k < L or g == T - 1) else 0 for k in range(L + 3)] for g in range(T)]
    O = [[0] * (L + 2) for _ in range(T)]
    qs = []

    def value():
        tot = 0
        for g in range(T):
            for w in range(L + 3):
                tot += E[g][w] << (32 * (g * L + w))
            for w in range(L + 2):
                tot += O[g][w] << (32 * (g * L + w + 1))
        return tot

    def add_at(arr, pos, v, size):
        """arr[pos..] += v with ripple; every entry stays < 2^32; must not overflow the array"""
        k = pos
        while v:
            assert k < size, "overflow out of the array"
            t = arr[k] + (v & M32)
            arr[k] = t & M32
            v = (v >> 32) + (t >> 32)
            k += 1

    def products(x, y0, y1):
        for g in range(T):
            for m in range(L // 2):
                add_at(E[g], 2 * m, x[g][2 * m] * y0, L + 3)          # even chain A
                add_at(O[g], 2 * m, x[g][2 * m + 1] * y0, L + 2)      # odd chain A
                add_at(O[g], 2 * m, x[g][2 * m] * y1, L + 2)          # odd chain B
                add_at(E[g], 2 * m + 2, x[g][2 * m + 1] * y1, L + 3)  # even chain B

    for i in range(0, S, 2):
        owner, j = divmod(i, L)
        before = value()
        products(A, Bm[owner][j], Bm[owner][j + 1])
        if A2 is not None:
            products(A2, B2[owner][j], B2[owner][j + 1])
        t0 = E[0][0]
        t1 = (E[0][1] + O[0][0]) & M32
        q0 = (t0 * np0) & M32
        q1 = (((t0 * np0) >> 32) + t0 * np1 + t1 * np0) & M32
        qs += [q0, q1]
        products(N, q0, q1)
        assert E[0][0] == 0 and (E[0][1] + O[0][0]) & M32 == 0
        # ---- shift by B^2
        s0 = [E[g][0] for g in range(T)]
        s1s = [E[g][1] + O[g][0] for g in range(T)]
        s1 = [v & M32 for v in s1s]
        c1 = [v >> 32 for v in s1s]
        for g in range(T):
            r0, r1 = (s0[g + 1], s1[g + 1]) if g < T - 1 else (0, 0)
            add_at(E[g], L, r0 | (r1 << 32), L + 3)                   # the upper lane's two lowest limbs
        for g in range(T):
            e, o = E[g], O[g]
            bridge = e[2] + o[1] + c1[g]                              # limb 2: the new limb 0
            ne = [bridge & M32] + e[3:] + [0, 0]
            no = o[2:] + [0, 0]
            add_at(no, 0, bridge >> 32, L + 2)                        # the bridge's carry enters the odd array at the new limb 1
            E[g], O[g] = ne, no
            assert len(ne) == L + 3 and len(no) == L + 2
        after = value()
        prods = a * (Bm[owner][j] + (Bm[owner][j + 1] << 32)) + ((a2 * (B2[owner][j] + (B2[owner][j + 1] << 32))) if a2 is not None else 0)
        assert after * (1 << 64) == before + prods + (q0 + (q1 << 32)) * n, "step is not exact"
        # bounds the kernel relies on: the hanging entries hold carries / one product only
        for g in range(T):
            assert E[g][L + 1] == 0 and E[g][L + 2] == 0 and O[g][L] == 0 and O[g][L + 1] == 0, ("hang", g, E[g][L:], O[g][L - 1:])
    tot = value()
    # the merge the kernel does before finish_x: per lane u[0 .. L + 1]; finish_x reads u[L] as the lane's whole overflow
    for g in range(T):
        lane = sum(E[g][w] << (32 * w) for w in range(L + 3)) + sum(O[g][w] << (32 * (w + 1)) for w in range(L + 2))
        assert lane < 1 << (32 * (L + 1)), ("lane value needs more than one overflow limb", g, lane >> (32 * (L + 1)))
    ab = a * b + (a2 * b2 if a2 is not None else 0)
    assert tot * R == init + ab + sum(q << (32 * k) for k, q in enumerate(qs)) * n
    assert tot % n == ((init + ab) * pow(R, -1, n)) % n
    assert tot < (3 if a2 is not None else 2) * n + (2 if init else 0)
    # the same quotient digits as row-by-row CIOS
    Q = ((init + ab) * ((-pow(n, -1, R)) % R)) % R
    assert Q == sum(q << (32 * k) for k, q in enumerate(qs))
    return tot


if __name__ == "__main__":
    random.seed(2)
    for (T, L) in [(32, 2), (32, 4), (16, 4), (16, 6), (16, 2), (8, 4), (8, 8), (16, 8), (4, 8), (8, 12)]:
        S = T * L
        W = 1 << (32 * S)
        for it in range(12):
            n = random.getrandbits(32 * S) | 1 | (1 << (32 * S - 1)) if it % 2 == 0 else (random.getrandbits(32 * S - random.randint(0, 40)) | 1)
            if it == 5:
                n = W - 5
            a, b = random.randrange(n), random.randrange(n)
            if it == 3:
                a = b = n - 1
            model_montmul_pair(a, b, n, T, L)
            init = random.randrange(W + n) if it % 3 else W + n - 1
            model_montmul_pair(a, b, n, T, L, init)
            a2, b2 = random.randrange(n), random.randrange(n)
            if it == 4:
                a = b = a2 = b2 = n - 1
            model_montmul_pair(a, b, n, T, L, init, a2, b2)
            if (W - 1) * b < W * n:
                model_montmul_pair(W - 1, b, n, T, L, init)
        print("pair rows", T, L, "ok")
