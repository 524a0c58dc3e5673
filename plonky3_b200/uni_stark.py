"""uni-stark `prove` for the Poseidon2 AIR with every data-parallel step on the GPU — the BASELINE config-5 benchmark
(`prove_prime_field_31 --field koala-bear --objective poseidon-2-permutations --log-trace-length L -d radix-2-dit-parallel
-m poseidon-2`).  Mirrors uni-stark/src/prover.rs:87-442 (prove_with_preprocessed) and fri/src/prover.rs:43-160 (prove_fri) with
the reference's names; host code is only the protocol sequencing (the transcript's sponge itself runs on the device,
challenger.py).  Non-ZK, no preprocessed columns, no public values — what the example binary proves
(examples/src/proofs.rs:120-170).

    trace (device)  --pcs.commit-->  trace cap ............................... p3gpu_coset_lde_batch_dev + p3gpu_merkle_commit_dev
    alpha <- transcript;  quotient values on GENERATOR * K ................... p3gpu_p2air_quotient_dev
    commit_quotient (2 chunks) ................................................ LDE + Merkle as above
    zeta <- transcript;  pcs.open: opened values + reduced openings .......... p3gpu_open_* / columnwise / rowwise dot kernels
    prove_fri: commit phase (fold + commit per round), grind, query openings . p3gpu_fri_fold_dev, p3gpu_challenger_grind,
                                                                               p3gpu_gather_rows_dev / p3gpu_merkle_paths_dev
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field as dc_field
from typing import List, Optional

import numpy as np

from . import _lib
from .challenger import DuplexChallenger
from .dft import Radix2DitParallel, _log2_strict
from .field import Field
from .fri import FriParameters, TwoAdicFriFolding, TwoAdicFriPcs, commit_phase
from .merkle_tree import MerkleTreeMmcs
from .poseidon2 import Poseidon2

VECTOR_LEN = 8           # examples/src/airs.rs: P2_VECTOR_LEN = 1 << 3


@dataclass
class RoundConstants:
    """poseidon2-air/src/constants.rs:28-57 (Montgomery form)."""
    beginning_full_round_constants: np.ndarray    # (4, 16)
    partial_round_constants: np.ndarray           # (rounds_p,)
    ending_full_round_constants: np.ndarray       # (4, 16)


class VectorizedPoseidon2Air:
    """VectorizedPoseidon2Air<KoalaBear, ..., WIDTH 16, SBOX_DEGREE 3, SBOX_REGISTERS 0, 4, 20, VECTOR_LEN 8> on the GPU."""

    def __init__(self, field: Field, constants: RoundConstants, gpu, vector_len: int = VECTOR_LEN):
        self.field, self.constants, self.gpu, self.vector_len = field, constants, gpu, vector_len
        self.rounds_p = int(np.asarray(constants.partial_round_constants).size)
        self._upload()

    def _upload(self):
        c = self.constants
        if self.gpu is None:                     # verifier-only use: the constraint folder below is host code
            return
        self.gpu.p2air_set_constants(self.field.id, c.beginning_full_round_constants, c.partial_round_constants, c.ending_full_round_constants)

    def width(self) -> int:                      # BaseAir::width (vectorized.rs)
        return self.vector_len * (144 + self.rounds_p)

    def max_constraint_degree(self) -> int:      # air.rs:151-160 for (3, 0)
        return 3

    def num_public_values(self) -> int: return 0
    def main_next_row_columns(self): return []   # no transition constraints: the next row is never opened (verifier.rs:431-440)

    def eval_folded_constraints(self, e, local, nxt, public_values, is_first_row, is_last_row, is_transition, alpha):
        """The verifier's constraint folder (uni-stark/src/folder.rs VerifierConstraintFolder: acc = acc * alpha + c per
        assert_zero) over the opened row at zeta, scalar host code on canonical EF4 values (`e`: verifier.Ext).  Constraint order =
        poseidon2-air/src/air.rs eval: per permutation the committed post-state of every full round (16 each) and the S-box output
        of every partial round; degree-3 S-box without registers."""
        f, P = self.field, self.field.P
        c = self.constants
        beg = [[f.from_monty(int(v)) for v in r] for r in np.asarray(c.beginning_full_round_constants).reshape(4, 16)]
        end = [[f.from_monty(int(v)) for v in r] for r in np.asarray(c.ending_full_round_constants).reshape(4, 16)]
        part = [f.from_monty(int(v)) for v in np.asarray(c.partial_round_constants).ravel()]
        ip = lambda k: pow(pow(2, k, P), P - 2, P)
        # internal diagonal of Poseidon2KoalaBear<16> (koala-bear/src/poseidon2.rs:410-428)
        v16 = [P - 2, 1, 2, ip(1), 3, 4, P - ip(1), P - 3, P - 4, ip(8), ip(3), ip(24), P - ip(8), P - ip(3), P - ip(4), P - ip(24)]
        add, sc, mul = e.add, e.scale, e.mul

        def mat4(x):
            a, b, cc, d = x
            return [add(add(sc(a, 2), sc(b, 3)), add(cc, d)), add(add(a, sc(b, 2)), add(sc(cc, 3), d)),
                    add(add(a, b), add(sc(cc, 2), sc(d, 3))), add(add(sc(a, 3), b), add(cc, sc(d, 2)))]

        def mds(s):
            s = sum((mat4(s[i:i + 4]) for i in range(0, 16, 4)), [])
            t = [[0, 0, 0, 0] for _ in range(4)]
            for i in range(16):
                t[i % 4] = add(t[i % 4], s[i])
            return [add(s[i], t[i % 4]) for i in range(16)]

        cube = lambda x: mul(mul(x, x), x)
        cols = 144 + self.rounds_p
        acc = [0, 0, 0, 0]
        for v in range(self.vector_len):
            col = local[v * cols:(v + 1) * cols]
            s = mds(col[:16]); k = 16
            for rc in beg:
                s = mds([cube(add(s[i], e.base(rc[i]))) for i in range(16)])
                for i in range(16):
                    acc = add(mul(acc, alpha), e.sub(s[i], col[k + i])); s[i] = col[k + i]
                k += 16
            for r in range(self.rounds_p):
                x = cube(add(s[0], e.base(part[r])))
                acc = add(mul(acc, alpha), e.sub(x, col[k])); s[0] = col[k]; k += 1
                t = [0, 0, 0, 0]
                for i in range(16):
                    t = add(t, s[i])
                s = [add(sc(s[i], v16[i]), t) for i in range(16)]
            for rc in end:
                s = mds([cube(add(s[i], e.base(rc[i]))) for i in range(16)])
                for i in range(16):
                    acc = add(mul(acc, alpha), e.sub(s[i], col[k + i])); s[i] = col[k + i]
                k += 16
        return acc

    def generate_trace_rows(self, inputs_dev):
        """generate_vectorized_trace_rows (generation.rs:14-70): (n_perms, 16) device inputs -> (n_perms / 8, 1312) device trace."""
        if self.gpu is None:
            raise _lib.P3GpuError("trace generation needs a GPU context (no CPU fallback)")
        self._upload()
        return self.gpu.p2air_generate_trace(self.field.id, inputs_dev, self.vector_len)

    def quotient_values(self, trace_lde_dev, log_degree: int, alpha):
        """uni-stark/src/prover.rs:462-827 on the committed LDE (natural order over the quotient domain)."""
        if self.gpu is None:
            raise _lib.P3GpuError("quotient evaluation needs a GPU context (no CPU fallback)")
        self._upload()
        return self.gpu.p2air_quotient(self.field.id, trace_lde_dev, log_degree, alpha, self.vector_len)


@dataclass
class StarkConfig:
    """uni-stark/src/config.rs:47-87: the PCS and the initial challenger state."""
    pcs: TwoAdicFriPcs
    challenger_perm: Poseidon2                    # DuplexChallenger<_, Perm24, 24, 16>
    challenger_rate: int = 16

    def initialise_challenger(self) -> DuplexChallenger:
        return DuplexChallenger(self.pcs.dft.field, self.challenger_perm, self.challenger_rate, self.pcs.dft.gpu)


@dataclass
class Proof:
    """uni-stark/src/proof.rs:19-62 + fri/src/proof.rs:12-24 as plain arrays (Montgomery words)."""
    trace_commit: np.ndarray
    quotient_commit: np.ndarray
    trace_local: np.ndarray                       # (width, 4)
    quotient_chunks: List[np.ndarray]             # per chunk (4, 4)
    commit_phase_commits: List[np.ndarray]
    commit_pow_witnesses: List[int]
    final_poly: np.ndarray
    query_pow_witness: int
    query_indices: List[int]
    input_openings: list                          # per round: (opened rows per matrix, paths)
    commit_phase_openings: list                   # per FRI round: (log_arity, sibling values (n, arity-1, 4), paths)
    degree_bits: int
    timings_ms: dict = dc_field(default_factory=dict)
    trace_next: Optional[np.ndarray] = None       # (width, 4) when the AIR reads the next row (uni-stark/src/proof.rs:52-56)
    input_opening_indices: list = dc_field(default_factory=list)     # per input batch: the height-reduced query indices
    commit_phase_indices: list = dc_field(default_factory=list)      # per FRI round: the opened group index of every query

    def to_postcard(self) -> bytes:
        """The reference's wire form (`postcard::to_allocvec(&proof)`, uni-stark/tests/fib_air.rs:401-412)."""
        from .proof_io import proof_to_postcard
        return proof_to_postcard(self)


def get_log_num_quotient_chunks(air) -> int:
    """uni-stark/src/symbolic.rs get_log_num_quotient_chunks: log2_ceil(max(constraint_degree, 2) - 1) (non-ZK)."""
    d = max(air.max_constraint_degree(), 2)
    return max(d - 2, 0).bit_length()


def verify(config: StarkConfig, air, proof, public_values=()):
    """uni-stark/src/verifier.rs:282-295.  Raises verifier.VerificationError."""
    from .verifier import verify as _verify
    return _verify(config, air, proof, public_values)


def prove(config: StarkConfig, air: VectorizedPoseidon2Air, trace, public_values=()) -> Proof:
    """uni-stark/src/prover.rs:87-442.  `trace`: device (CUDA int32) matrix of height 2^n."""
    import torch
    pcs, f, gpu = config.pcs, config.pcs.dft.field, config.pcs.dft.gpu
    sync = torch.cuda.synchronize
    T = {}

    def span(name, t0):
        sync(); T[name] = (time.perf_counter() - t0) * 1e3

    assert len(public_values) == 0, "the Poseidon2 AIR has no public values"
    degree = int(trace.shape[0])
    log_degree = _log2_strict(degree)
    log_num_quotient_chunks = get_log_num_quotient_chunks(air)
    num_quotient_chunks = 1 << log_num_quotient_chunks
    assert log_num_quotient_chunks == pcs.fri.log_blowup, "quotient domain must equal the LDE domain (fast path of get_evaluations_on_domain)"
    challenger = config.initialise_challenger()
    trace_domain = pcs.natural_domain_for_degree(degree)

    t0 = time.perf_counter()
    trace_commit, trace_data = pcs.commit([(trace_domain, trace)])                       # prover.rs:215
    span("commit to trace data", t0)

    challenger.observe_canonical(log_degree)                                             # log_ext_degree (non-ZK)       :224
    challenger.observe_canonical(log_degree)                                             # log_degree                    :225
    challenger.observe_canonical(0)                                                      # preprocessed_width            :226
    challenger.observe_cap(trace_commit)                                                 # :230
    alpha = challenger.sample_algebra_element()                                          # :258

    t0 = time.perf_counter()
    quotient_domain = (f.mul(trace_domain[0], f.generator), log_degree + log_num_quotient_chunks)      # create_disjoint_domain
    trace_on_quotient_domain = pcs.get_evaluations_on_domain(trace_data, 0, quotient_domain).bit_reverse_rows()
    quotient_flat = air.quotient_values(trace_on_quotient_domain, log_degree, alpha)    # (2N, 4) natural order = flatten_to_base
    span("compute quotient polynomial", t0)

    t0 = time.perf_counter()
    quotient_commit, quotient_data = pcs.commit_quotient(quotient_domain, quotient_flat, num_quotient_chunks)     # :319
    span("commit to quotient poly chunks", t0)
    challenger.observe_cap(quotient_commit)

    zeta = challenger.sample_algebra_element()                                           # :365
    t0 = time.perf_counter()
    rounds = [(trace_data, [[zeta]]), (quotient_data, [[zeta]] * num_quotient_chunks)]   # main_next_row_columns() is empty: zeta only
    opened_values, fri_inputs = pcs.open_values_and_fri_inputs(rounds, challenger)
    span("open: evaluate + reduce", t0)

    t0 = time.perf_counter()
    fri = prove_fri(pcs, fri_inputs, challenger, rounds)
    span("open: FRI", t0)

    return Proof(trace_commit=trace_commit, quotient_commit=quotient_commit, trace_local=opened_values[0][0][0],
                 quotient_chunks=[v[0] for v in opened_values[1]], commit_phase_commits=fri["commits"],
                 commit_pow_witnesses=fri["pow_witnesses"], final_poly=fri["final_poly"], query_pow_witness=fri["query_pow_witness"],
                 query_indices=fri["indices"], input_openings=fri["input_openings"], commit_phase_openings=fri["commit_phase_openings"],
                 degree_bits=log_degree, timings_ms=T, input_opening_indices=fri["input_opening_indices"],
                 commit_phase_indices=fri["commit_phase_indices"])


def prove_fri(pcs: TwoAdicFriPcs, inputs: list, challenger: DuplexChallenger, prover_data_with_opening_points: list) -> dict:
    """fri/src/prover.rs:43-160."""
    import torch
    params: FriParameters = pcs.fri
    f, gpu = pcs.dft.field, pcs.dft.gpu
    assert inputs and params.num_queries > 0
    log_global_max_height = _log2_strict(int(inputs[0].shape[0]))
    T = {}
    t0 = time.perf_counter()
    res = commit_phase(TwoAdicFriFolding(f, gpu), params, inputs, challenger, pcs.dft)
    torch.cuda.synchronize(); T["commit phase"] = (time.perf_counter() - t0) * 1e3
    for la in res.log_arities:
        challenger.observe_canonical(la)                                                 # :108-110
    t0 = time.perf_counter()
    pow_witness = challenger.grind(params.query_proof_of_work_bits)                      # :112
    T["grind"] = (time.perf_counter() - t0) * 1e3
    indices = [challenger.sample_bits(log_global_max_height) for _ in range(params.num_queries)]     # extra_query_index_bits = 0
    t0 = time.perf_counter()
    # open_inputs (:380-417): every committed batch at the (height-reduced) query indices
    input_openings, input_opening_indices, commit_phase_indices = [], [], []
    for data, _ in prover_data_with_opening_points:
        log_max_height = _log2_strict(pcs.mmcs.get_max_height(data))
        reduced = [i >> (log_global_max_height - log_max_height) for i in indices]
        input_openings.append(pcs.mmcs.open_multi_batch(reduced, data))
        input_opening_indices.append(reduced)
    # answer_queries (:308-378)
    commit_phase_openings, cur = [], list(indices)
    for la, data in zip(res.log_arities, res.data):
        group = [i >> la for i in cur]
        commit_phase_indices.append(group)
        rows, paths = params.mmcs.open_multi_batch(group, data)
        opened = rows[0].reshape(len(cur), 1 << la, 4)
        keep = np.array([[j for j in range(1 << la) if j != (i & ((1 << la) - 1))] for i in cur], dtype=np.int64)
        siblings = np.take_along_axis(opened, keep[:, :, None], axis=1)
        commit_phase_openings.append((la, siblings, paths))
        cur = group
    torch.cuda.synchronize(); T["query phase"] = (time.perf_counter() - t0) * 1e3
    return {"commits": res.commits, "pow_witnesses": res.pow_witnesses, "final_poly": res.final_poly, "query_pow_witness": pow_witness,
            "indices": indices, "input_openings": input_openings, "commit_phase_openings": commit_phase_openings, "log_arities": res.log_arities,
            "input_opening_indices": input_opening_indices, "commit_phase_indices": commit_phase_indices, "timings_ms": T}
