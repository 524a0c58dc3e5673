#!/usr/bin/env python3
"""Extract the commitments of the reference's second committed proof fixture, batch-stark/tests/fixtures/
batch_stark_two_adic_v1.postcard (verified by batch-stark/tests/simple.rs:1792-1802), into tests/golden/batch_stark_two_adic_v1.json.
Run in the build container (needs /root/reference).

Wire format (postcard): BatchProof { commitments: BatchCommitments { main, permutation: Option, quotient_chunks, random: Option }, .. }
(batch-stark/src/proof.rs:9-26); a commitment is MerkleCap = Vec<[F; 8]> (varint length, then 8 x 4-byte LE Montgomery words per
digest).  The main commitment is ONE MMCS commitment over the LDEs of BOTH instance traces (batch-stark/src/prover.rs:225-231): the
only reference-held pin for "several matrices in one Merkle tree" (row-wise concatenation in input order, merkle_tree.rs:312-316)."""
import json
import pathlib
import struct

SRC = pathlib.Path("/root/reference/batch-stark/tests/fixtures/batch_stark_two_adic_v1.postcard")
OUT = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden" / "batch_stark_two_adic_v1.json"
b = SRC.read_bytes()
pos = 0


def byte():
    global pos
    pos += 1
    return b[pos - 1]


def cap():
    global pos
    n = byte()
    out = []
    for _ in range(n):
        out.append(list(struct.unpack_from("<8I", b, pos))); pos += 32
    return out


g = {"main_cap": cap()}
assert byte() == 1                          # permutation = Some (the case uses global lookups)
g["permutation_cap"] = cap()
g["quotient_chunks_cap"] = cap()
assert byte() == 0                          # random = None
g["source"] = "batch-stark/tests/fixtures/batch_stark_two_adic_v1.postcard (%d bytes)" % len(b)
g["case"] = "two_adic_compat_case (simple.rs:1693-1732): BabyBear, Perm = Poseidon2BabyBear<16>::new_from_rng_128(SmallRng(777)), " \
            "cap_height 1, log_blowup 2; instances: mul_trace(32, reps 2) 32 x 7 and fib_trace(0, 1, 32) 32 x 2"
P = 0x78000001
assert all(v < P for c in (g["main_cap"], g["permutation_cap"], g["quotient_chunks_cap"]) for d in c for v in d)
assert len(g["main_cap"]) == 2
OUT.write_text(json.dumps(g, indent=0))
print({k: (len(v) if isinstance(v, list) else v) for k, v in g.items()})
