#!/usr/bin/env python3
"""Extract the reference's own known-answer data into small fixtures.

Run HERE (the build container), where the read-only reference checkout is mounted:

    python tests/golden/make_golden.py [/root/reference]

It only READS literal test data / constants out of the reference's test functions
(decimal strings, hex strings, Montgomery limb literals) and writes them as JSON /
text next to this script.  No reference source code is copied.  The fixtures travel
to the GPU box; the reference checkout does not.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

N_G1 = 2000   # of 10000
N_G2 = 400    # of 10000
N_FR = 10000  # all


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def fn_body(src, name):
    """Text of `fn name(...) { ... }` (brace matched)."""
    m = re.search(r"fn\s+%s\s*[<(]" % re.escape(name), src)
    assert m, name
    i = src.index("{", m.end())
    depth, j = 0, i
    while True:
        if src[j] == "{":
            depth += 1
        elif src[j] == "}":
            depth -= 1
            if depth == 0:
                return src[i:j + 1]
        j += 1


def from_strs(text):
    return re.findall(r'from_str\("(\d+)"\)', text)


def limb_arrays(text):
    """All `[a, b, c, d]` u64 literals (hex or decimal) -> python int (little-endian limbs)."""
    out = []
    for m in re.finditer(r"\[\s*((?:0x[0-9a-fA-F]+|\d+)\s*,\s*(?:0x[0-9a-fA-F]+|\d+)\s*,\s*"
                         r"(?:0x[0-9a-fA-F]+|\d+)\s*,\s*(?:0x[0-9a-fA-F]+|\d+))\s*\]", text):
        limbs = [int(x.strip(), 0) for x in m.group(1).split(",")]
        out.append(sum(l << (64 * i) for i, l in enumerate(limbs)))
    return out


def main():
    groups = read("src/groups/mod.rs")
    fields = read("src/fields/mod.rs")
    fp = read("src/fields/fp.rs")
    fq2 = read("src/fields/fq2.rs")
    fq6 = read("src/fields/fq6.rs")
    fq12 = read("src/fields/fq12.rs")
    ser = read("tests/serialization.rs")

    # ---- pairing KATs (src/groups/mod.rs:522-547, 637-762, 773-796) ----
    ml = from_strs(fn_body(groups, "test_miller_loop"))
    rp = from_strs(fn_body(groups, "test_reduced_pairing"))
    pg = from_strs(fn_body(groups, "test_prepared_g2"))
    assert len(ml) == 14 and len(rp) == 14 and len(pg) == 1 + 4 + 102 * 6, (len(ml), len(rp), len(pg))
    assert ml[:2] == rp[:2] and pg[0] == ml[1]
    kat = {
        "source": "src/groups/mod.rs test_miller_loop / test_reduced_pairing / test_prepared_g2",
        "k1": ml[0], "k2": ml[1],
        "miller_loop": ml[2:],
        "reduced_pairing": rp[2:],
        "prepared_q_affine": pg[1:5],
        # field order inside each literal: ell_0(c0,c1), ell_vw(c0,c1), ell_vv(c0,c1)
        "prepared_coeffs": [pg[5 + 6 * i: 11 + 6 * i] for i in range(102)],
    }
    with open(os.path.join(OUT, "pairing_kat.json"), "w") as f:
        json.dump(kat, f, indent=0)

    # ---- Fq12 KATs (src/fields/mod.rs:83-201) ----
    tv = from_strs(fn_body(fields, "fq12_test_vector"))
    ce = from_strs(fn_body(fields, "test_cyclotomic_exp"))
    ts = from_strs(fn_body(fields, "test_str"))
    assert len(tv) == 24 and len(ce) == 24 and len(ts) == 2
    with open(os.path.join(OUT, "fq12_kat.json"), "w") as f:
        json.dump({
            "source": "src/fields/mod.rs fq12_test_vector / test_cyclotomic_exp / test_str",
            "vector_start": tv[:12], "vector_final": tv[12:],
            "cyclotomic_orig": ce[:12], "cyclotomic_expected": ce[12:],
            "minus_one_fr": ts[0], "minus_one_fq": ts[1],
        }, f, indent=0)

    # ---- constants (Montgomery limb literals), to pin the oracle's derived constants ----
    fr_m = re.search(r"field_impl!\(\s*Fr,(.*?)\);", fp, re.S).group(1)
    fq_m = re.search(r"field_impl!\(\s*Fq,(.*?)\);", fp, re.S).group(1)

    def params(txt):
        arr = limb_arrays(txt)
        inv = int(re.findall(r"(0x[0-9a-fA-F]+)\s*$", txt.strip())[0], 16)
        return {"modulus": arr[0], "rsquared": arr[1], "rcubed": arr[2], "one": arr[3], "inv": inv}

    consts = {
        "source": "Montgomery-form limb literals from src/fields/*.rs and src/groups/mod.rs",
        "fr": params(fr_m), "fq": params(fq_m),
        "fq_non_residue": limb_arrays(fn_body(fq2, "fq_non_residue"))[0],
        "fq2_nonresidue": limb_arrays(fn_body(fq2, "fq2_nonresidue")),
        "fq6_frob_c1": limb_arrays(fn_body(fq6, "frobenius_coeffs_c1")),   # powers 1,2,3: (c0,c1),(c0),(c0,c1)
        "fq6_frob_c2": limb_arrays(fn_body(fq6, "frobenius_coeffs_c2")),
        "fq12_frob_c1": limb_arrays(fn_body(fq12, "frobenius_coeffs_c1")),
        "g1_one_y": limb_arrays(fn_body(groups[groups.index("impl GroupParams for G1Params"):], "one"))[0],
        "g1_coeff_b": limb_arrays(fn_body(groups[groups.index("impl GroupParams for G1Params"):], "coeff_b"))[0],
        "g2_one": limb_arrays(fn_body(groups[groups.index("impl GroupParams for G2Params"):], "one")),
        "g2_coeff_b": limb_arrays(fn_body(groups[groups.index("impl GroupParams for G2Params"):], "coeff_b")),
        "two_inv": limb_arrays(fn_body(groups, "two_inv"))[0],
        "ate_loop_count": limb_arrays(fn_body(groups, "ate_loop_count"))[0],
        "twist_mul_by_q_x": limb_arrays(fn_body(groups, "twist_mul_by_q_x")),
        "twist_mul_by_q_y": limb_arrays(fn_body(groups, "twist_mul_by_q_y")),
        "exp_by_neg_z_u": int(re.search(r"U256\(\[(\d+), 0, 0, 0\]\)", fn_body(fq12, "exp_by_neg_z")).group(1)),
    }
    consts = json.loads(json.dumps(consts, default=str))

    def stringify(o):
        if isinstance(o, int) and not isinstance(o, bool):
            return str(o)
        if isinstance(o, list):
            return [stringify(x) for x in o]
        if isinstance(o, dict):
            return {k: stringify(v) for k, v in o.items()}
        return o

    with open(os.path.join(OUT, "constants.json"), "w") as f:
        json.dump(stringify(consts), f, indent=0)

    # ---- serialization golden vectors (tests/serialization.rs:74-30143) ----
    def vectors(name):
        body = fn_body(ser, name)
        return re.findall(r'"([0-9a-f]+)"', body[:body.index("];")])

    g1v, g2v, frv = vectors("g1_vectors"), vectors("g2_vectors"), vectors("fr_vectors")
    assert len(g1v) == len(g2v) == len(frv) == 10000
    for fname, vec, n in (("g1_vectors.txt", g1v, N_G1), ("g2_vectors.txt", g2v, N_G2), ("fr_vectors.txt", frv, N_FR)):
        with open(os.path.join(OUT, fname), "w") as f:
            f.write("# first %d of 10000 vectors from tests/serialization.rs %s (hex wire format)\n" % (n, fname[:-4]))
            f.write("\n".join(vec[:n]) + "\n")
    # last vector of each, so a full 10000-step replay can still be pinned cheaply
    with open(os.path.join(OUT, "last_vectors.json"), "w") as f:
        json.dump({"g1_9999": g1v[-1], "g2_9999": g2v[-1], "fr_9999": frv[-1],
                   "scalar": re.search(r'from_str\("(\d+)"\)', fn_body(ser, "g1_vectors")).group(1)}, f, indent=0)

    edge = fn_body(ser, "group_serialization_edge_cases")
    hexes = re.findall(r'from_hex::<(G1|G2)>\("([0-9a-f]+)"\)', edge)
    with open(os.path.join(OUT, "wire_edge_cases.json"), "w") as f:
        json.dump({"source": "tests/serialization.rs group_serialization_edge_cases (all but '00' must be rejected)",
                   "cases": hexes}, f, indent=0)
    print("wrote fixtures to", OUT)


if __name__ == "__main__":
    main()
