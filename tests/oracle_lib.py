"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
_lib = None

u8p = C.POINTER(C.c_uint8)


class Fe(C.Structure):
    _fields_ = [("v", C.c_uint64 * 5)]


class P3(C.Structure):
    _fields_ = [("X", Fe), ("Y", Fe), ("Z", Fe), ("T", Fe)]


def build():
    subprocess.check_call(["make", "-s", "-C", ODIR])


def load():
    global _lib
    if _lib is not None:
        return _lib
    so = os.path.join(ODIR, "liboracle.so")
    srcs = [os.path.join(ODIR, f) for f in os.listdir(ODIR) if f.endswith((".c", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        build()
    _lib = Oracle(C.CDLL(so))
    return _lib


def buf(b):
    return (C.c_uint8 * len(b)).from_buffer_copy(bytes(b)) if len(b) else (C.c_uint8 * 1)()


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.scalar_to_radix_2w_size_hint.restype = C.c_size_t

    # ---- field ----
    def fe_from_bytes(self, b):
        f = Fe(); self.lib.fe_from_bytes(C.byref(f), buf(b)); return f

    def fe_to_bytes(self, f):
        o = (C.c_uint8 * 32)(); self.lib.fe_to_bytes(o, C.byref(f)); return bytes(o)

    def fe_op2(self, name, a, b):
        o = Fe(); getattr(self.lib, name)(C.byref(o), C.byref(a), C.byref(b)); return o

    def fe_op1(self, name, a):
        o = Fe(); getattr(self.lib, name)(C.byref(o), C.byref(a)); return o

    def fe_sqrt_ratio_i(self, u, v):
        o = Fe(); ok = self.lib.fe_sqrt_ratio_i(C.byref(o), C.byref(u), C.byref(v)); return ok, o

    # ---- scalars ----
    def sc_op2(self, name, a, b):
        o = (C.c_uint8 * 32)(); getattr(self.lib, name)(o, buf(a), buf(b)); return bytes(o)

    def sc_op1(self, name, a):
        o = (C.c_uint8 * 32)(); getattr(self.lib, name)(o, buf(a)); return bytes(o)

    def scalar_from_wide(self, a):
        o = (C.c_uint8 * 32)(); self.lib.scalar_from_bytes_mod_order_wide(o, buf(a)); return bytes(o)

    def scalar_is_canonical(self, a):
        return bool(self.lib.scalar_is_canonical(buf(a)))

    def naf(self, a, w):
        o = (C.c_int8 * 256)(); self.lib.scalar_non_adjacent_form(o, buf(a), C.c_uint(w)); return list(o)

    def radix16(self, a):
        o = (C.c_int8 * 64)(); self.lib.scalar_as_radix_16(o, buf(a)); return list(o)

    def radix2w(self, a, w):
        o = (C.c_int8 * 64)(); self.lib.scalar_as_radix_2w(o, buf(a), C.c_uint(w)); return list(o)

    def radix_size_hint(self, w):
        return self.lib.scalar_to_radix_2w_size_hint(C.c_uint(w))

    # ---- points ----
    def basepoint(self):
        p = P3(); self.lib.ge_basepoint(C.byref(p)); return p

    def identity(self):
        p = P3(); self.lib.ge_identity(C.byref(p)); return p

    def decompress(self, b):
        p = P3(); ok = self.lib.ge_decompress(C.byref(p), buf(b)); return p if ok else None

    def compress(self, p):
        o = (C.c_uint8 * 32)(); self.lib.ge_compress(o, C.byref(p)); return bytes(o)

    def scalarmul(self, s, p):
        o = P3(); self.lib.ge_scalarmul(C.byref(o), buf(s), C.byref(p)); return o

    def add(self, p, q):
        o = P3(); self.lib.ge_p3_add(C.byref(o), C.byref(p), C.byref(q)); return o

    def sub(self, p, q):
        o = P3(); self.lib.ge_p3_sub(C.byref(o), C.byref(p), C.byref(q)); return o

    def double(self, p):
        o = P3(); self.lib.ge_p3_double(C.byref(o), C.byref(p)); return o

    def mul_by_pow_2(self, p, k):
        o = P3(); self.lib.ge_mul_by_pow_2(C.byref(o), C.byref(p), C.c_uint32(k)); return o

    def ct_eq(self, p, q):
        return bool(self.lib.ge_p3_ct_eq(C.byref(p), C.byref(q)))

    def is_identity(self, p):
        return bool(self.lib.ge_is_identity(C.byref(p)))

    def p3_limbs(self, p):
        o = (C.c_uint64 * 20)(); self.lib.ge_p3_to_limbs(o, C.byref(p)); return list(o)

    def p3_from_limbs(self, limbs):
        p = P3(); self.lib.ge_p3_from_limbs(C.byref(p), (C.c_uint64 * 20)(*limbs)); return p

    def _pts(self, points):
        arr = (P3 * max(1, len(points)))()
        pres = (C.c_uint8 * max(1, len(points)))()
        for i, p in enumerate(points):
            if p is None:
                pres[i] = 0; self.lib.ge_identity(C.byref(arr[i]))
            else:
                pres[i] = 1; arr[i] = p
        return arr, pres

    def msm(self, which, scalars, points):
        """which in {'pippenger','straus_vartime','optional'}; points may hold None."""
        fn = {"pippenger": self.lib.msm_pippenger, "straus_vartime": self.lib.msm_straus_vartime,
              "optional": self.lib.edwards_optional_multiscalar_mul}[which]
        arr, pres = self._pts(points)
        o = P3()
        ok = fn(C.byref(o), buf(b"".join(scalars)), arr, pres, C.c_size_t(len(points)))
        return o if ok else None

    def precomputed_straus(self, static_scalars, static_points, dynamic_scalars, dynamic_points):
        """VartimePrecomputedStraus::optional_mixed_multiscalar_mul (precomputed_straus.rs:57-126); dynamic points may
        hold None.  Returns the point, None, or raises if there are more static scalars than static points."""
        sarr, _ = self._pts(static_points)
        darr, pres = self._pts(dynamic_points)
        o = P3()
        rc = self.lib.msm_precomputed_straus(C.byref(o), buf(b"".join(static_scalars)), C.c_size_t(len(static_scalars)), sarr,
                                             C.c_size_t(len(static_points)), buf(b"".join(dynamic_scalars)), darr, pres,
                                             C.c_size_t(len(dynamic_points)))
        if rc < 0:
            raise ValueError("more static scalars than static points")
        return o if rc else None

    def scalar_invert_batch(self, scalars):
        """Scalar::invert_batch (C/scalar.rs:779-853): (list of inverses, product of all inverses)."""
        n = len(scalars)
        b = (C.c_uint8 * (32 * max(n, 1))).from_buffer_copy(b"".join(scalars) + bytes(32 * (1 if n == 0 else 0)))
        ret = (C.c_uint8 * 32)()
        self.lib.scalar_invert_batch(b, C.c_size_t(n), ret)
        raw = bytes(b)
        return [raw[32 * i:32 * i + 32] for i in range(n)], bytes(ret)

    def compress_batch(self, points):
        """EdwardsPoint::compress_batch (C/edwards.rs:619-647)."""
        arr, _ = self._pts(points)
        out = (C.c_uint8 * (32 * max(len(points), 1)))()
        self.lib.ge_compress_batch(out, arr, C.c_size_t(len(points)))
        return bytes(out)[:32 * len(points)]

    def ristretto_double_and_compress_batch(self, points):
        """RistrettoPoint::double_and_compress_batch (C/ristretto.rs:564-646)."""
        arr, _ = self._pts(points)
        out = (C.c_uint8 * (32 * max(len(points), 1)))()
        self.lib.ristretto_double_and_compress_batch(out, arr, C.c_size_t(len(points)))
        return bytes(out)[:32 * len(points)]

    def msm_ct(self, scalars, points):
        arr, _ = self._pts(points)
        o = P3()
        self.lib.edwards_multiscalar_mul(C.byref(o), buf(b"".join(scalars)), arr, C.c_size_t(len(points)))
        return o

    # ---- ristretto ----
    def ristretto_decompress(self, b):
        p = P3(); ok = self.lib.ristretto_decompress(C.byref(p), buf(b)); return p if ok else None

    def ristretto_compress(self, p):
        o = (C.c_uint8 * 32)(); self.lib.ristretto_compress(o, C.byref(p)); return bytes(o)

    def ristretto_ct_eq(self, p, q):
        return bool(self.lib.ristretto_ct_eq(C.byref(p), C.byref(q)))

    def ristretto_double_base_batch(self, a, b, G, H):
        n = len(a) // 32
        o = (C.c_uint8 * (32 * max(n, 1)))()
        rc = self.lib.ristretto_double_base_batch(o, buf(a), buf(b), buf(G), buf(H), C.c_size_t(n))
        return rc, bytes(o)[:32 * n]

    # ---- hashing ----
    def sha512(self, m):
        o = (C.c_uint8 * 64)(); self.lib.sha512(o, buf(m), C.c_size_t(len(m))); return bytes(o)

    def keccak_f1600(self, lanes):
        a = (C.c_uint64 * 25)(*lanes); self.lib.keccak_f1600(a); return list(a)

    # ---- ed25519 ----
    def _msgs(self, msgs):
        n = len(msgs)
        bufs = [buf(m) for m in msgs]
        ptrs = (C.POINTER(C.c_uint8) * max(n, 1))(*[C.cast(b, C.POINTER(C.c_uint8)) for b in bufs])
        lens = (C.c_size_t * max(n, 1))(*[len(m) for m in msgs])
        return bufs, ptrs, lens

    def verify_batch(self, msgs, sigs, pks, chunk=0, want_zs=False):
        n = len(msgs)
        bufs, ptrs, lens = self._msgs(msgs)
        zs = (C.c_uint8 * (16 * max(n, 1)))()
        rc = self.lib.ed25519_verify_batch_chunked(ptrs, lens, buf(b"".join(sigs)), buf(b"".join(pks)),
                                                   C.c_size_t(n), C.c_size_t(chunk), zs)
        return (rc, bytes(zs)[:16 * n]) if want_zs else rc

    def verify(self, msg, sig, pk, strict=False):
        fn = self.lib.ed25519_verify_strict if strict else self.lib.ed25519_verify
        return fn(buf(msg), C.c_size_t(len(msg)), buf(sig), buf(pk))

    def public_key(self, seed):
        o = (C.c_uint8 * 32)(); self.lib.ed25519_public_key(o, buf(seed)); return bytes(o)

    def sign(self, msg, seed):
        o = (C.c_uint8 * 64)(); self.lib.ed25519_sign(o, buf(msg), C.c_size_t(len(msg)), buf(seed)); return bytes(o)
