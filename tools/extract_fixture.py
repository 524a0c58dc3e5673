#!/usr/bin/env python3
"""Extract the reference's committed proof fixture uni-stark/tests/fixtures/uni_stark_two_adic_v1.postcard (verified by
uni-stark/tests/fib_air.rs:414-422) into tests/golden/uni_stark_two_adic_v1.json — EVERY field of the proof, plus the raw bytes
(`postcard_hex`) so that the proof serialiser (plonky3_b200/proof_io.py) can be checked byte for byte.  Run in the build container
(needs /root/reference).

Wire format (postcard): struct fields in declaration order (uni-stark/src/proof.rs:19-62, fri/src/proof.rs:12-75,
merkle-tree/src/pruning.rs:83-89), Vec = varint length + items, Option = 0/1 tag, u8 = one byte, usize = varint, every field
element = 4 bytes LE of the MONTGOMERY representation (monty-31/src/monty_31.rs:167-179), EF = 4 consecutive F, a digest = 8
consecutive F (arrays carry no length).  All numbers in the JSON are those raw Montgomery u32 values.
"""
import json, pathlib, struct

SRC = pathlib.Path("/root/reference/uni-stark/tests/fixtures/uni_stark_two_adic_v1.postcard")
OUT = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden" / "uni_stark_two_adic_v1.json"
b = SRC.read_bytes()
pos = 0


def varint():
    global pos
    r = s = 0
    while True:
        c = b[pos]; pos += 1
        r |= (c & 0x7F) << s; s += 7
        if c < 0x80:
            return r


def felts(n):
    global pos
    v = list(struct.unpack_from("<%dI" % n, b, pos)); pos += 4 * n
    return v


def digests(): return [felts(8) for _ in range(varint())]
def ef_vec(): return [felts(4) for _ in range(varint())]


g = {}
g["trace_cap"] = digests()
g["quotient_cap"] = digests()
assert varint() == 0                       # commitments.random = None
g["trace_local"] = ef_vec()
assert varint() == 1                       # trace_next = Some
g["trace_next"] = ef_vec()
assert varint() == 0 and varint() == 0     # preprocessed_* = None
g["quotient_chunks"] = [ef_vec() for _ in range(varint())]
assert varint() == 0                       # opened_values.random = None
g["commit_phase_commits"] = [digests() for _ in range(varint())]
g["commit_pow_witnesses"] = felts(varint())
g["input_openings"] = []                   # Vec<BatchMultiOpening>: opened_values[query][matrix][col], one pruned multiproof
for _ in range(varint()):
    ov = [[felts(varint()) for _ in range(varint())] for _ in range(varint())]
    g["input_openings"].append({"opened_values": ov, "proof": digests()})
g["commit_phase_openings"] = []            # Vec<CommitPhaseMultiStep>
for _ in range(varint()):
    la = b[pos]; pos += 1
    sv = [ef_vec() for _ in range(varint())]
    g["commit_phase_openings"].append({"log_arity": la, "sibling_values": sv, "proof": digests()})
g["final_poly"] = ef_vec()
g["query_pow_witness"] = felts(1)[0]
g["degree_bits"] = varint()
assert pos == len(b), "trailing bytes"
g["postcard_hex"] = b.hex()
g["source"] = "uni-stark/tests/fixtures/uni_stark_two_adic_v1.postcard (%d bytes)" % len(b)
OUT.write_text(json.dumps(g, indent=0))
print({k: (len(v) if isinstance(v, (list, str)) else v) for k, v in g.items()})
