#!/usr/bin/env python3
"""Extract the reference's RNG-constant Poseidon2 known-answer tests (test_poseidon2_width_{16,24}_random in
koala-bear/src/poseidon2.rs:527-573 and baby-bear/src/poseidon2.rs:512-558: constants from Xoroshiro128Plus::seed_from_u64(1) through
Poseidon2::new_from_rng_128) into tests/golden/poseidon2_rng_kat.json.  Run in the build container (needs /root/reference)."""
import json, pathlib, re

OUT = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden" / "poseidon2_rng_kat.json"
g = {}
for name, path in (("koala_bear", "/root/reference/koala-bear/src/poseidon2.rs"), ("baby_bear", "/root/reference/baby-bear/src/poseidon2.rs")):
    src = pathlib.Path(path).read_text()
    for w in (16, 24):
        body = src[src.index(f"fn test_poseidon2_width_{w}_random"):]
        body = body[:body.index("assert_eq!")]
        arrays = re.findall(r"new_array\(\[(.*?)\]\)", body, re.S)
        vals = [[int(x) for x in re.findall(r"\d+", a)] for a in arrays]
        assert len(vals) == 2 and len(vals[0]) == len(vals[1]) == w
        g[f"{name}_{w}"] = {"input": vals[0], "expected": vals[1]}
g["source"] = "test_poseidon2_width_{16,24}_random: koala-bear/src/poseidon2.rs:527-573, baby-bear/src/poseidon2.rs:512-558 (canonical integers)"
OUT.write_text(json.dumps(g, indent=0))
print({k: (len(v["input"]) if isinstance(v, dict) else v) for k, v in g.items()})
