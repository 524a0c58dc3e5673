"""Wire form of a uni-stark proof: what `postcard::to_allocvec(&Proof<SC>)` writes for
SC = StarkConfig<TwoAdicFriPcs<Val, Dft, MerkleTreeMmcs<.., 2, 8>, ExtensionMmcs<..>>, BinomialExtensionField<Val, 4>, ..>
(uni-stark/tests/fib_air.rs:401-412), so that a proof made on the GPU is read by the reference's verifier with
`postcard::from_bytes` and vice versa.  Host-side re-encoding only; no field arithmetic happens here.

Layout = serde declaration order, postcard rules:
  Proof            { commitments, opened_values, opening_proof, degree_bits: usize }                 uni-stark/src/proof.rs:19-26
  Commitments      { trace: MerkleCap, quotient_chunks: MerkleCap, random: Option<..> }              proof.rs:44-49
  OpenedValues     { trace_local: Vec<EF>, trace_next: Option<Vec<EF>>, preprocessed_local: Option, preprocessed_next: Option,
                     quotient_chunks: Vec<Vec<EF>>, random: Option }                                 proof.rs:51-62
  FriProof         { commit_phase_commits: Vec<MerkleCap>, commit_pow_witnesses: Vec<F>, input_openings: Vec<BatchMultiOpening>,
                     commit_phase_openings: Vec<CommitPhaseMultiStep>, final_poly: Vec<EF>, query_pow_witness: F }   fri/src/proof.rs:12-24
  BatchMultiOpening{ opened_values: Vec<Vec<Vec<F>>> (query, matrix, column), opening_proof: PrunedMerklePaths }      fri/src/proof.rs:68-75
  CommitPhaseMultiStep { log_arity: u8, sibling_values: Vec<Vec<EF>>, opening_proof: PrunedMerklePaths }              fri/src/proof.rs:33-44
  PrunedMerklePaths{ sibling_hashes: Vec<[F; 8]> }                                                   merkle-tree/src/pruning.rs:83-89
  MerkleCap = Vec<[F; 8]>;  Vec = varint length + items;  Option = tag byte;  usize = varint;  u8 = one byte;  arrays carry no length;
  F = the 4 little-endian bytes of the Montgomery word (monty-31/src/monty_31.rs:167-179);  EF = 4 F.
Digests are [F; 8] (the Poseidon2 MMCS of the example configurations); a Keccak MMCS commits to [u64; 4] digests, which postcard
writes as varints — that configuration's proofs are not covered by this module.
Pinned byte for byte against the reference's committed proof fixture (tests/golden/uni_stark_two_adic_v1.json `postcard_hex`)."""
from __future__ import annotations

import numpy as np

from .merkle_tree import prune_paths


def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _words(a) -> bytes:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint32)).astype("<u4").tobytes()


def _vec_of(a, width: int) -> bytes:
    """Vec<[F; width]> (width = 8: digests / caps, 4: extension elements)."""
    a = np.asarray(a, dtype=np.uint32).reshape(-1, width)
    return _varint(a.shape[0]) + _words(a)


def _option_vec_ef(a) -> bytes:
    return b"\x00" if a is None else b"\x01" + _vec_of(a, 4)


def proof_to_postcard(proof) -> bytes:
    """`proof`: plonky3_b200.uni_stark.Proof (non-ZK, no preprocessed trace)."""
    out = bytearray()
    out += _vec_of(proof.trace_commit, 8) + _vec_of(proof.quotient_commit, 8) + b"\x00"
    out += _vec_of(proof.trace_local, 4) + _option_vec_ef(proof.trace_next) + b"\x00\x00"
    out += _varint(len(proof.quotient_chunks)) + b"".join(_vec_of(c, 4) for c in proof.quotient_chunks) + b"\x00"
    out += _varint(len(proof.commit_phase_commits)) + b"".join(_vec_of(c, 8) for c in proof.commit_phase_commits)
    out += _varint(len(proof.commit_pow_witnesses)) + _words(np.array(proof.commit_pow_witnesses, dtype=np.uint32))
    assert len(proof.input_opening_indices) == len(proof.input_openings) and len(proof.commit_phase_indices) == len(proof.commit_phase_openings)
    out += _varint(len(proof.input_openings))
    for (rows, paths), idx in zip(proof.input_openings, proof.input_opening_indices):
        out += _varint(len(idx))
        if len(idx):                                              # all queries have the same byte layout: assemble them as one (n, bytes) block
            n = len(idx)
            parts = [np.tile(np.frombuffer(_varint(len(rows)), dtype=np.uint8), (n, 1))]
            for m in rows:
                m = np.ascontiguousarray(np.asarray(m, dtype=np.uint32).reshape(n, -1)).astype("<u4")
                parts.append(np.tile(np.frombuffer(_varint(m.shape[1]), dtype=np.uint8), (n, 1)))
                parts.append(m.view(np.uint8).reshape(n, -1))
            out += np.hstack(parts).tobytes()
        out += _vec_of(prune_paths(idx, paths), 8)
    out += _varint(len(proof.commit_phase_openings))
    for (log_arity, siblings, paths), idx in zip(proof.commit_phase_openings, proof.commit_phase_indices):
        out += bytes([log_arity]) + _varint(len(idx))
        if len(idx):
            sib = np.ascontiguousarray(np.asarray(siblings, dtype=np.uint32).reshape(len(idx), -1)).astype("<u4")      # (n, (arity - 1) * 4)
            pre = np.tile(np.frombuffer(_varint(sib.shape[1] // 4), dtype=np.uint8), (len(idx), 1))
            out += np.hstack([pre, sib.view(np.uint8).reshape(len(idx), -1)]).tobytes()
        out += _vec_of(prune_paths(idx, paths), 8)
    out += _vec_of(proof.final_poly, 4) + _words([proof.query_pow_witness]) + _varint(proof.degree_bits)
    return bytes(out)


class _Reader:
    def __init__(self, data: bytes, prime=None):
        self.b, self.pos, self.prime = data, 0, prime

    def varint(self) -> int:
        r = s = 0
        while True:
            if self.pos >= len(self.b):
                raise ValueError("truncated proof")
            c = self.b[self.pos]; self.pos += 1
            r |= (c & 0x7F) << s; s += 7
            if c < 0x80:
                return r

    def byte(self) -> int:
        if self.pos >= len(self.b):
            raise ValueError("truncated proof")
        self.pos += 1
        return self.b[self.pos - 1]

    def words(self, n: int) -> np.ndarray:
        if self.pos + 4 * n > len(self.b):
            raise ValueError("truncated proof")
        a = np.frombuffer(self.b, dtype="<u4", count=n, offset=self.pos).astype(np.uint32)
        self.pos += 4 * n
        if self.prime is not None and n and int(a.max()) >= self.prime:
            raise ValueError("Value is out of range")               # MontyField31::deserialize (monty-31/src/monty_31.rs:181-196)
        return a

    def vec_of(self, width: int) -> np.ndarray:
        n = self.varint()
        return self.words(n * width).reshape(n, width)

    def option_vec_ef(self):
        tag = self.byte()
        if tag > 1:
            raise ValueError("bad Option tag")
        return self.vec_of(4) if tag else None


def proof_from_postcard(data: bytes, prime=None) -> dict:
    """The inverse: every field of the wire proof as arrays of Montgomery words (pruned multiproofs are left pruned — the
    verifier consumes them against its own query indices).  With `prime` given, words >= prime are rejected as the reference's
    deserialiser rejects them (one field element has one encoding).  Raises ValueError on malformed input."""
    r = _Reader(data, prime)
    p = {"trace_commit": r.vec_of(8), "quotient_commit": r.vec_of(8)}
    if r.byte() != 0:
        raise ValueError("ZK (random) commitments are not supported")
    p["trace_local"] = r.vec_of(4)
    p["trace_next"] = r.option_vec_ef()
    if r.byte() != 0 or r.byte() != 0:
        raise ValueError("preprocessed openings are not supported")
    p["quotient_chunks"] = [r.vec_of(4) for _ in range(r.varint())]
    if r.byte() != 0:
        raise ValueError("ZK (random) openings are not supported")
    p["commit_phase_commits"] = [r.vec_of(8) for _ in range(r.varint())]
    p["commit_pow_witnesses"] = [int(v) for v in r.words(r.varint())]
    p["input_openings"] = []
    for _ in range(r.varint()):
        ov = [[r.words(r.varint()) for _ in range(r.varint())] for _ in range(r.varint())]
        p["input_openings"].append({"opened_values": ov, "proof": r.vec_of(8)})
    p["commit_phase_openings"] = []
    for _ in range(r.varint()):
        la = r.byte()
        sv = [r.vec_of(4) for _ in range(r.varint())]
        p["commit_phase_openings"].append({"log_arity": la, "sibling_values": sv, "proof": r.vec_of(8)})
    p["final_poly"] = r.vec_of(4)
    p["query_pow_witness"] = int(r.words(1)[0])
    p["degree_bits"] = r.varint()
    if r.pos != len(data):
        raise ValueError("trailing bytes after the proof")
    return p
