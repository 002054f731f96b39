#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors for the hot path into JSON fixtures.

Run in the build container (needs /root/reference, which does NOT exist on the GPU box):
    python tests/golden/make_golden.py
Outputs (committed): tests/golden/kat.json, ed25519_testvectors.json, ed25519_validation.json

Only DATA (byte strings of test vectors) is extracted; each entry records the reference
file:line it came from.
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def read(path):
    with open(os.path.join(REF, path)) as f:
        return f.read()


def ints_in(text):
    return [int(t, 0) for t in re.findall(r"-?(?:0x[0-9a-fA-F]+|\d+)", text)]


def array_after(src, name, path):
    """Return (list of ints, line) for `static NAME ... = ... [ ... ]`."""
    m = re.search(r"\bstatic\s+%s\b[^=]*=" % re.escape(name), src)
    if m is None:
        m = re.search(r"\bconst\s+%s\b[^=]*=" % re.escape(name), src)
    assert m, (name, path)
    i = src.index("[", m.end())
    depth, j = 0, i
    while True:
        if src[j] == "[":
            depth += 1
        elif src[j] == "]":
            depth -= 1
            if depth == 0:
                break
        j += 1
    line = src.count("\n", 0, m.start()) + 1
    return ints_in(src[i + 1:j]), line


def hexb(v):
    return bytes((x + 256) % 256 for x in v).hex()


kat = {}


def grab(path, names, group):
    src = read(path)
    for n in names:
        v, line = array_after(src, n, path)
        kat.setdefault(group, {})[n] = {"src": "%s:%d" % (path, line),
                                        "hex" if n != "A_NAF" else "ints": hexb(v) if n != "A_NAF" else v}


grab("curve25519-dalek/src/field.rs", ["A_BYTES", "ASQ_BYTES", "AINV_BYTES", "AP58_BYTES", "B_BYTES"], "field")
grab("curve25519-dalek/src/scalar.rs",
     ["X", "XINV", "Y", "X_TIMES_Y", "CANONICAL_2_256_MINUS_1", "A_SCALAR", "A_NAF",
      "LARGEST_UNREDUCED_SCALAR"], "scalar")
grab("curve25519-dalek/src/edwards.rs",
     ["BASE_X_COORD_BYTES", "BASE2_CMPRSSD", "BASE16_CMPRSSD", "A_SCALAR", "B_SCALAR",
      "A_TIMES_BASEPOINT", "DOUBLE_SCALAR_MULT_RESULT"], "edwards")
grab("curve25519-dalek/src/constants.rs",
     ["ED25519_BASEPOINT_COMPRESSED", "RISTRETTO_BASEPOINT_COMPRESSED"], "constants")

# scalar.rs from_bytes_mod_order_wide KAT: x + 2^256 x mod l
src = read("curve25519-dalek/src/scalar.rs")
m = re.search(r"fn from_bytes_mod_order_wide\(\) \{.*?let reduced = Scalar \{\s*bytes: \[(.*?)\]", src, re.S)
kat["scalar"]["X_PLUS_2_256_X_REDUCED"] = {
    "src": "curve25519-dalek/src/scalar.rs:%d" % (src.count("\n", 0, m.start()) + 1),
    "hex": hexb(ints_in(m.group(1)))}

# u64 limb-level constants cross-check (constants.rs): values only
src = read("curve25519-dalek/src/backend/serial/u64/constants.rs")
for n in ["EDWARDS_D", "EDWARDS_D2", "SQRT_M1", "INVSQRT_A_MINUS_D", "L", "R", "RR"]:
    v, line = array_after(src, n, "u64/constants.rs")
    kat.setdefault("u64_constants", {})[n] = {
        "src": "curve25519-dalek/src/backend/serial/u64/constants.rs:%d" % line, "limbs": v}
m = re.search(r"const LFACTOR: u64 = (0x[0-9a-f]+);", src)
kat["u64_constants"]["LFACTOR"] = {"src": "u64/constants.rs", "limbs": [int(m.group(1), 16)]}
m = re.search(r"ED25519_BASEPOINT_POINT: EdwardsPoint = EdwardsPoint \{(.*?)\n\};", src, re.S)
kat["u64_constants"]["ED25519_BASEPOINT_POINT_XYZT"] = {
    "src": "curve25519-dalek/src/backend/serial/u64/constants.rs:%d" % (src.count("\n", 0, m.start()) + 1),
    "limbs": ints_in(re.sub(r"FieldElement51", "", m.group(1)))}
# eight-torsion points (edge-case inputs)
m = re.search(r"EIGHT_TORSION_INNER_DOC_HIDDEN: \[EdwardsPoint; 8\] = \[(.*?)\n\];", src, re.S)
kat["u64_constants"]["EIGHT_TORSION_XYZT"] = {
    "src": "curve25519-dalek/src/backend/serial/u64/constants.rs:%d" % (src.count("\n", 0, m.start()) + 1),
    "limbs": ints_in(re.sub(r"FieldElement51|EdwardsPoint", "", m.group(1)))}

# ristretto: encodings of 0..15 * basepoint, and the bad-encoding lists
src = read("curve25519-dalek/src/ristretto.rs")
m = re.search(r"fn encodings_of_small_multiples_of_basepoint\(\) \{.*?let compressed = \[(.*?)\n        \];", src, re.S)
encs = re.findall(r"CompressedRistretto\(\[(.*?)\]\)", m.group(1), re.S)
kat["ristretto"] = {"SMALL_MULTIPLES": {
    "src": "curve25519-dalek/src/ristretto.rs:%d" % (src.count("\n", 0, m.start()) + 1),
    "hex": [hexb(ints_in(e)) for e in encs]}}
assert len(encs) == 16

with open(os.path.join(HERE, "kat.json"), "w") as f:
    json.dump(kat, f, indent=1, sort_keys=True)

# ed25519 TESTVECTORS (ed25519.cr.yp.to sign.input): sk||pk : pk : msg : sig||msg :
tv = []
with open(os.path.join(REF, "ed25519-dalek/TESTVECTORS")) as f:
    for ln, line in enumerate(f, 1):
        parts = line.strip().split(":")
        if len(parts) < 4:
            continue
        tv.append({"line": ln, "seed": parts[0][:64], "pk": parts[1], "msg": parts[2], "sig": parts[3][:128]})
with open(os.path.join(HERE, "ed25519_testvectors.json"), "w") as f:
    json.dump({"src": "ed25519-dalek/TESTVECTORS", "vectors": tv}, f)

# VALIDATIONVECTORS (C2SP/CCTV ed25519vectors): keep number,key,sig,msg,flags
with open(os.path.join(REF, "ed25519-dalek/VALIDATIONVECTORS")) as f:
    vv = json.load(f)
vv = [{"number": v["number"], "key": v["key"], "sig": v["sig"], "msg": v["msg"],
       "flags": v.get("flags") or []} for v in vv]
with open(os.path.join(HERE, "ed25519_validation.json"), "w") as f:
    json.dump({"src": "ed25519-dalek/VALIDATIONVECTORS", "vectors": vv}, f)
print("kat groups:", {k: len(v) for k, v in kat.items()}, "testvectors:", len(tv), "validation:", len(vv))
