#!/usr/bin/env python3
"""Extract numeric parameters and known-answer vectors from the reference tree.

Run once in the build container (needs /root/reference); outputs are committed:
  plonky3_b200/p2_constants.json   Poseidon2 round constants (canonical form) for
                                   BabyBear/KoalaBear widths 16 and 24
                                   (baby-bear/src/poseidon2.rs:111-281, koala-bear/src/poseidon2.rs:119-285)
  tests/golden/poseidon2_kat.json  known-answer vectors of the default-constant permutations
                                   (koala-bear/src/poseidon2.rs:614-653, baby-bear/src/poseidon2.rs:599-639)
  tests/golden/two_adic_generators.json  (baby_bear.rs:48-53, koala_bear.rs:73-78)
Only numbers are extracted, no code.
"""
import json, re, sys, pathlib

REF = pathlib.Path("/root/reference")
OUT = pathlib.Path(__file__).resolve().parent.parent


def ints(s):
    return [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\b\d+\b", s)]


def const_block(src, name):
    i = src.index("pub const " + name)
    j = src.index(";\n", src.index("=", i))
    body = src[src.index("=", i) + 1 : j]
    body = body[body.index("(") :]          # drop `KoalaBear::new_2d_array`
    return ints(body)


def kat(src, fn):
    i = src.index("fn " + fn)
    blk = src[i : src.index("assert_eq!", i)]
    a = blk.index("new_array(")
    b = blk.index("new_array(", a + 1)
    inp = ints(blk[a : blk.index("]);", a)][0:])
    exp = ints(blk[b : blk.index("]);", b)][0:])
    return inp, exp


def main():
    consts, kats, gens = {}, {}, {}
    for fld, d, pfx in (("baby_bear", "baby-bear", "BABYBEAR"), ("koala_bear", "koala-bear", "KOALABEAR")):
        src = (REF / d / "src" / "poseidon2.rs").read_text()
        for w in (16, 24):
            ini = const_block(src, f"{pfx}_POSEIDON2_RC_{w}_EXTERNAL_INITIAL")
            fin = const_block(src, f"{pfx}_POSEIDON2_RC_{w}_EXTERNAL_FINAL")
            itl = const_block(src, f"{pfx}_POSEIDON2_RC_{w}_INTERNAL")
            # first two ints of each block come from the type annotation "[[F; w]; 4]" / "[F; n]"
            ini = ini[-4 * w :]; fin = fin[-4 * w :]
            rp = int(re.search(rf"{pfx}_POSEIDON2_PARTIAL_ROUNDS_{w}: usize = (\d+)", src).group(1))
            itl = itl[-rp:]
            assert len(ini) == 4 * w and len(fin) == 4 * w and len(itl) == rp
            consts[f"{fld}_{w}"] = {"external_initial": ini, "external_final": fin, "internal": itl}
            name = "babybear" if fld == "baby_bear" else "koalabear"
            inp, exp = kat(src, f"test_default_{name}_poseidon2_width_{w}")
            assert len(inp) == w and len(exp) == w, (len(inp), len(exp))
            kats[f"{fld}_{w}"] = {"input": inp, "expected": exp}
        fsrc = (REF / d / "src" / f"{fld}.rs").read_text()
        i = fsrc.index("const TWO_ADIC_GENERATORS")
        blk = fsrc[i : fsrc.index("]);", i)]
        gens[fld] = ints(blk[blk.index("new_array(") :])
    (OUT / "plonky3_b200" / "p2_constants.json").write_text(json.dumps(consts))
    (OUT / "tests" / "golden" / "poseidon2_kat.json").write_text(json.dumps(kats, indent=0))
    (OUT / "tests" / "golden" / "two_adic_generators.json").write_text(json.dumps(gens))
    print({k: (len(v["external_initial"]), len(v["internal"])) for k, v in consts.items()}, {k: len(v) for k, v in gens.items()})


if __name__ == "__main__":
    main()
