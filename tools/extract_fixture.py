#!/usr/bin/env python3
"""Extract the values our hot path must reproduce from the reference's committed proof fixture
uni-stark/tests/fixtures/uni_stark_two_adic_v1.postcard (verified by uni-stark/tests/fib_air.rs:414-422)
into tests/golden/uni_stark_two_adic_v1.json.  Run in the build container (needs /root/reference).

Wire format (postcard): struct fields in order (uni-stark/src/proof.rs:19-62, fri/src/proof.rs:12-24),
Vec = varint length + items, Option = 0/1 tag, every field element = 4 bytes LE of the MONTGOMERY
representation (monty-31/src/monty_31.rs:167-179), EF = 4 consecutive F.  All numbers in the JSON are
those raw Montgomery u32 values.
"""
import json, pathlib, struct

SRC = pathlib.Path("/root/reference/uni-stark/tests/fixtures/uni_stark_two_adic_v1.postcard")
OUT = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden" / "uni_stark_two_adic_v1.json"
b = SRC.read_bytes()
pos = 0

def byte():
    global pos
    pos += 1
    return b[pos - 1]

def felts(n):
    global pos
    v = list(struct.unpack_from("<%dI" % n, b, pos)); pos += 4 * n
    return v

def cap():
    n = byte()
    return [felts(8) for _ in range(n)]

def ef_vec():
    n = byte()
    return [felts(4) for _ in range(n)]

g = {}
g["trace_cap"] = cap()
g["quotient_cap"] = cap()
assert byte() == 0                       # commitments.random = None
g["trace_local"] = ef_vec()
assert byte() == 1                       # trace_next = Some
g["trace_next"] = ef_vec()
assert byte() == 0 and byte() == 0       # preprocessed_* = None
g["quotient_chunks"] = [ef_vec() for _ in range(byte())]
assert byte() == 0                       # opened_values.random = None
g["commit_phase_commits"] = [cap() for _ in range(byte())]
g["commit_pow_witnesses"] = felts(byte())
e = len(b)
g["degree_bits"] = b[e - 1]
g["query_pow_witness"] = struct.unpack_from("<I", b, e - 5)[0]
assert b[e - 70] == 4
g["final_poly"] = [list(struct.unpack_from("<4I", b, e - 69 + 16 * i)) for i in range(4)]
g["source"] = "uni-stark/tests/fixtures/uni_stark_two_adic_v1.postcard (%d bytes)" % e
OUT.write_text(json.dumps(g, indent=0))
print({k: (len(v) if isinstance(v, list) else v) for k, v in g.items()})
