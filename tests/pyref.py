"""Independent big-integer model of edwards25519 / ristretto255 (affine arithmetic, RFC 8032 /
RFC 9496 formulas).  Used to cross-check the C oracle and to derive expected values at sizes
where the reference holds no known answer."""
import hashlib

p = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
d = (-121665 * pow(121666, p - 2, p)) % p
SQRT_M1 = pow(2, (p - 1) // 4, p)


def inv(x):
    return pow(x, p - 2, p)


def add(P, Q):
    x1, y1 = P
    x2, y2 = Q
    k = d * x1 * x2 * y1 * y2 % p
    return ((x1 * y2 + x2 * y1) * inv(1 + k) % p, (y1 * y2 + x1 * x2) * inv(1 - k) % p)


def neg(P):
    return ((-P[0]) % p, P[1])


def mul(s, P):
    Q = (0, 1)
    while s > 0:
        if s & 1:
            Q = add(Q, P)
        P = add(P, P)
        s >>= 1
    return Q


def recover_x(y, sign):
    u = (y * y - 1) % p
    v = (d * y * y + 1) % p
    x2 = u * inv(v) % p
    x = pow(x2, (p + 3) // 8, p)
    if (x * x - x2) % p != 0:
        x = x * SQRT_M1 % p
    if (x * x - x2) % p != 0:
        return None
    if x & 1:
        x = p - x
    if sign:
        x = (p - x) % p
    return x


By = 4 * inv(5) % p
B = (recover_x(By, 0), By)


def compress(P):
    x, y = P
    return (y | ((x & 1) << 255)).to_bytes(32, "little")


def decompress(b):
    """dalek semantics: y taken mod 2^255 without range check (C/edwards.rs:211-257)."""
    v = int.from_bytes(b, "little")
    y = (v & ((1 << 255) - 1)) % p
    x = recover_x(y, v >> 255)
    if x is None:
        return None
    return (x, y)


def msm(scalars, points):
    acc = (0, 1)
    for s, P in zip(scalars, points):
        acc = add(acc, mul(s, P))
    return acc


def sc(b):
    return int.from_bytes(b, "little")


def sc_bytes(x):
    return (x % L).to_bytes(32, "little")


def labelled_scalar(label, seed, i):
    """SURVEY 8(d) config-1 generator: from_bytes_mod_order_wide(SHA-512(label || seed_le64 || i_le64))."""
    h = hashlib.sha512(label + seed.to_bytes(8, "little") + i.to_bytes(8, "little")).digest()
    return int.from_bytes(h, "little") % L
