"""uni-stark `prove` of the vectorised Poseidon2 AIR on the GPU (plonky3_b200.uni_stark, BASELINE config 5 at small sizes) against the
CPU replay built from the oracle (tests/p2_prove_replay.py): every transcript-visible value — trace cap, quotient cap, opened
values, FRI round caps, final polynomial, proof-of-work witness (smallest), query indices, every opened row and authentication
path — bit for bit, with the example binary's constants (SmallRng seed 1: AIR round constants, Perm16, Perm24, in that order,
examples/examples/prove_prime_field_31.rs:115,150,190-191) and its FRI parameters (new_benchmark_high_arity, cap_height 3)."""
import numpy as np
import pytest
import torch

from oracle import p3_oracle as O
import p2_prove_replay as R

from plonky3_b200 import _lib
from plonky3_b200.dft import Radix2DitParallel
from plonky3_b200.field import KoalaBear
from plonky3_b200.fri import FriParameters, TwoAdicFriPcs
from plonky3_b200.gpu import default_gpu
from plonky3_b200.merkle_tree import MerkleTreeMmcs
from plonky3_b200.poseidon2 import Poseidon2
from plonky3_b200.uni_stark import RoundConstants, StarkConfig, VectorizedPoseidon2Air, prove

pytestmark = pytest.mark.gpu
f = KoalaBear


def _gpu_perm(pm):
    w = pm.width
    return Poseidon2.new(f, w, np.array(pm.rc_init)[: 4 * w].reshape(4, w), np.array(pm.rc_term)[: 4 * w].reshape(4, w),
                         np.array(pm.rc_int)[: pm.rounds_p], monty=True)


@pytest.fixture(scope="module")
def setup():
    assert torch.cuda.is_available() and _lib.LIB_PATH.exists()
    gpu = default_gpu(0)
    rng = O.SmallRng(1)
    oair = O.air_from_rng(f.id, rng)
    o16 = O.perm_from_rng(f.id, 16, rng); o24 = O.perm_from_rng(f.id, 24, rng)
    p16, p24 = _gpu_perm(o16), _gpu_perm(o24)
    return gpu, oair, o16, o24, p16, p24


def _config(gpu, p16, p24, num_queries, pow_bits):
    mmcs = MerkleTreeMmcs.poseidon2(p16, p24, cap_height=3, gpu=gpu)                       # get_poseidon2_mmcs(perm16, perm24, 3)
    fri = FriParameters(1, 0, 3, num_queries, 0, pow_bits, mmcs)                           # new_benchmark_high_arity
    pcs = TwoAdicFriPcs(Radix2DitParallel(f, gpu), mmcs, fri)
    return StarkConfig(pcs, p24, 16)


@pytest.mark.parametrize("log_n,num_queries,pow_bits", [(4, 7, 5), (6, 100, 16), (9, 100, 8)])
def test_prove_matches_cpu_replay(setup, log_n, num_queries, pow_bits):
    gpu, oair, o16, o24, p16, p24 = setup
    inputs = O.SmallRng(1).field(f.id, (8 << log_n) * 16).reshape(-1, 16)                  # generate_random_trace_rows' inputs
    exp = R.prove(oair, o16, o24, inputs, num_queries=num_queries, query_pow_bits=pow_bits)
    assert R.verify_constraints_at_zeta(oair, exp)
    config = _config(gpu, p16, p24, num_queries, pow_bits)
    air = VectorizedPoseidon2Air(f, RoundConstants(np.array(oair.beg).reshape(4, 16), np.array(oair.part)[: oair.rounds_p], np.array(oair.end).reshape(4, 16)), gpu)
    trace = air.generate_trace_rows(torch.from_numpy(inputs.view(np.int32)).cuda())
    proof = prove(config, air, trace)
    assert proof.degree_bits == log_n
    assert np.array_equal(proof.trace_commit, exp["trace_cap"])
    assert np.array_equal(proof.quotient_commit, exp["quotient_cap"])
    assert np.array_equal(proof.trace_local, exp["trace_local"])
    for a, b in zip(proof.quotient_chunks, exp["quotient_chunks"]):
        assert np.array_equal(a, b)
    assert len(proof.commit_phase_commits) == len(exp["commit_phase_commits"])
    for a, b in zip(proof.commit_phase_commits, exp["commit_phase_commits"]):
        assert np.array_equal(a, b)
    assert np.array_equal(proof.final_poly, exp["final_poly"])
    assert proof.query_pow_witness == exp["query_pow_witness"]
    assert proof.query_indices == exp["indices"]
    for (rows, paths), (erows, epaths) in zip(proof.input_openings, exp["input_openings"]):
        assert len(rows) == len(erows)
        for a, b in zip(rows, erows):
            assert np.array_equal(a, b)
        assert np.array_equal(paths, epaths)
    for (la, sib, paths), (ela, esib, epaths) in zip(proof.commit_phase_openings, exp["commit_phase_openings"]):
        assert la == ela and np.array_equal(sib, esib) and np.array_equal(paths, epaths)
    # the GPU proof satisfies the verifier's identity as well (same check on the GPU's own opened values)
    mine = dict(exp, trace_local=proof.trace_local, quotient_chunks=proof.quotient_chunks)
    assert R.verify_constraints_at_zeta(oair, mine)
    # wire form (postcard, pruned multiproofs): same bytes as the replay's, and the restated reference verifier — the one that
    # accepts the reference's own proof fixture (tests/test_oracle.py) — accepts what the GPU wrote
    import stark_verify as V
    from plonky3_b200.proof_io import proof_from_postcard
    raw = proof.to_postcard()
    assert raw == R.to_wire_proof(exp).to_postcard()
    V.verify(V.Fld(f.id), R.verifier_config(o16, o24, num_queries=num_queries, query_pow_bits=pow_bits), V.poseidon2_air(oair), proof_from_postcard(raw))
    # and the product's verifier with the batch hashing and the transcript on the GPU (plonky3_b200.verifier)
    from plonky3_b200.uni_stark import verify
    from plonky3_b200.verifier import VerificationError
    verify(config, air, proof)
    bad = bytearray(raw); bad[len(raw) // 2] ^= 4
    with pytest.raises(VerificationError):
        verify(config, air, bytes(bad))


def test_gpu_verifier_on_the_reference_fixture(setup):
    """The reference's committed proof (uni-stark/tests/fixtures/uni_stark_two_adic_v1.postcard, bytes in the golden JSON) is accepted
    by plonky3_b200.verifier with Poseidon2 leaf hashing, node compression and the duplex challenger on the GPU; a wrong public
    value and corrupted bytes are rejected."""
    import json
    import pathlib
    import fixture_replay as FR
    import stark_verify as V
    from plonky3_b200.field import BabyBear
    from plonky3_b200.uni_stark import verify
    from plonky3_b200.verifier import VerificationError
    gpu = setup[0]
    gold = json.loads((pathlib.Path(__file__).resolve().parent / "golden" / "uni_stark_two_adic_v1.json").read_text())
    raw = bytes.fromhex(gold["postcard_hex"])
    rc_i, rc_t, rc_p = FR.fixture_constants()
    pm = Poseidon2.new(BabyBear, 16, rc_i, rc_t, rc_p, monty=True)
    mmcs = MerkleTreeMmcs.poseidon2(pm, None, 0, gpu)
    pcs = TwoAdicFriPcs(Radix2DitParallel(BabyBear, gpu), mmcs, FriParameters(2, 2, 1, 2, 1, 1, mmcs))       # fib_air.rs:134-155
    config = StarkConfig(pcs, pm, 8)                                                                         # DuplexChallenger<Val, Perm, 16, 8>
    verify(config, V.FibonacciAir(), raw, [0, 1, 21])
    with pytest.raises(VerificationError):
        verify(config, V.FibonacciAir(), raw, [0, 1, 22])
    for pos in range(2, len(raw) - 1, 97):
        bad = bytearray(raw); bad[pos] ^= 1
        with pytest.raises(VerificationError):
            verify(config, V.FibonacciAir(), bytes(bad), [0, 1, 21])


def test_challenger_matches_oracle(setup):
    """DuplexChallenger on the device vs the oracle restatement: observe / sample interleavings, clone, grind."""
    from plonky3_b200.challenger import DuplexChallenger
    gpu, _, _, o24, _, p24 = setup
    ch = DuplexChallenger(f, p24, 16, gpu)
    oc = R.OracleChallenger(o24)
    vals = O.random_matrix(f.id, 1, 200, seed=13)[0]
    k = 0
    for n_obs, n_smp in [(1, 1), (15, 2), (16, 0), (17, 5), (0, 20), (40, 3), (3, 0), (0, 1)]:
        ch.observe_slice(vals[k:k + n_obs]); oc.observe_slice(vals[k:k + n_obs]); k += n_obs
        got = ch.sample_many(n_smp) if n_smp else np.zeros(0, dtype=np.uint32)
        assert list(got) == [oc.sample() for _ in range(n_smp)]
    dv = torch.from_numpy(vals[100:150].view(np.int32)).cuda()
    ch.observe_slice(dv); oc.observe_slice(vals[100:150])                              # device-resident observe
    c2, o2 = ch.clone(), oc.clone()
    for bits in (1, 7, 12):
        assert ch.grind(bits) == oc.grind(bits)
    assert ch.sample_bits(20) == oc.sample_bits(20)
    assert c2.sample_bits(9) == o2.sample_bits(9)                                      # the clone kept the pre-grind state


def test_open_multi_batch_matches_open_batch(setup):
    gpu, _, _, _, p16, p24 = setup
    mmcs = MerkleTreeMmcs.poseidon2(p16, p24, cap_height=2, gpu=gpu)
    a = O.random_matrix(f.id, 256, 11, seed=1); b = O.random_matrix(f.id, 64, 5, seed=2)
    dev = lambda m: torch.from_numpy(m.view(np.int32)).cuda()
    _, tree = mmcs.commit([dev(a), dev(b)])
    idx = [0, 1, 77, 255, 128, 77]
    rows, paths = mmcs.open_multi_batch(idx, tree)
    for q, i in enumerate(idx):
        op, proof = mmcs.open_batch(i, tree)
        assert np.array_equal(rows[0][q], op[0]) and np.array_equal(rows[1][q], op[1])
        assert np.array_equal(paths[q], np.array(proof))
    _, htree = mmcs.commit([a, b])                                                      # host-resident prover data
    hrows, hpaths = mmcs.open_multi_batch(idx, htree)
    assert all(np.array_equal(x, y) for x, y in zip(rows, hrows)) and np.array_equal(paths, hpaths)


@pytest.mark.parametrize("log_blowup,log_final_poly_len,max_log_arity,cap_height", [(1, 0, 1, 0), (1, 2, 2, 2), (2, 1, 3, 1), (1, 3, 4, 0)])
def test_pcs_commit_open_verify_mixed_heights(setup, log_blowup, log_final_poly_len, max_log_arity, cap_height):
    """fri/tests/pcs.rs (the macro suite: several rounds, matrices of different heights in one batch, one or two opening points,
    blowup 1/2, arity 2..16, final polynomial 1..8 coefficients): Pcs::commit -> Pcs::open on the GPU, Pcs::verify by the product
    verifier (hashing on the GPU); a tampered opened value, opened row, multiproof digest, sibling value and final polynomial are
    rejected."""
    from plonky3_b200.challenger import DuplexChallenger
    from plonky3_b200.verifier import VerificationError
    gpu, oair, o16, o24, p16, p24 = setup
    mmcs = MerkleTreeMmcs.poseidon2(p16, p24, cap_height=cap_height, gpu=gpu)
    pcs = TwoAdicFriPcs(Radix2DitParallel(f, gpu), mmcs, FriParameters(log_blowup, log_final_poly_len, max_log_arity, 9, 1, 3, mmcs))
    g = torch.Generator(device="cuda"); g.manual_seed(11 * log_blowup + max_log_arity)
    shapes = [[(5, 3), (8, 2)], [(8, 1), (7, 4), (5, 2)]]                  # per round: (log degree, width)

    def transcript_head(ch, commits):
        for c in commits:
            ch.observe_cap(c)
        return [ch.sample_algebra_element() for _ in range(2)]
    commits, datas = [], []
    for rnd in shapes:
        evals = [((f.ONE, d), torch.randint(0, f.P, (1 << d, w), device="cuda", dtype=torch.int32, generator=g)) for d, w in rnd]
        c, pd = pcs.commit(evals)
        commits.append(c); datas.append(pd)
    ch = DuplexChallenger(f, p24, 16, gpu)
    zeta, zeta2 = transcript_head(ch, commits)
    points = [[[zeta], [zeta, zeta2]], [[zeta2], [zeta], [zeta, zeta2]]]
    opened, proof = pcs.open(list(zip(datas, points)), ch)

    def claims(opened_values):
        return [(c, [((f.ONE, d), [(z, ys) for z, ys in zip(pts, ov)]) for (d, _), pts, ov in zip(rnd, rpts, rov)])
                for c, rnd, rpts, rov in zip(commits, shapes, points, opened_values)]

    def check(opened_values, prf):
        chv = DuplexChallenger(f, p24, 16, gpu)
        z1, z2 = transcript_head(chv, commits)
        assert np.array_equal(z1, zeta) and np.array_equal(z2, zeta2)
        pcs.verify(claims(opened_values), prf, chv)
    check(opened, proof)
    import copy
    bad = copy.deepcopy(opened); bad[1][1][0][2][1] ^= 1
    with pytest.raises(VerificationError):
        check(bad, proof)
    for mutate in (lambda p: p["input_openings"][0]["opened_values"][3][1].__setitem__(0, int(p["input_openings"][0]["opened_values"][3][1][0]) ^ 1),
                   lambda p: p["input_openings"][1]["proof"].__setitem__((0, 0), int(p["input_openings"][1]["proof"][0, 0]) ^ 1),
                   lambda p: p["commit_phase_openings"][0]["sibling_values"][2].__setitem__((0, 0), int(p["commit_phase_openings"][0]["sibling_values"][2][0, 0]) ^ 1),
                   lambda p: p["final_poly"].__setitem__((0, 0), int(p["final_poly"][0, 0]) ^ 1),
                   lambda p: p["commit_phase_openings"].pop()):
        prf = copy.deepcopy(proof)
        mutate(prf)
        with pytest.raises(VerificationError):
            check(opened, prf)
