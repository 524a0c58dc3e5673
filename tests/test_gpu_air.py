"""Poseidon2 AIR on the GPU: trace generation and quotient evaluation of VectorizedPoseidon2Air
(KoalaBear, width 16, degree-3 S-box) against the oracle's restatement of poseidon2-air/src/{generation,air,vectorized}.rs and
uni-stark/src/prover.rs:462-827, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import p3_oracle as O

from plonky3_b200 import _lib
from plonky3_b200.field import KoalaBear
from plonky3_b200.gpu import default_gpu

pytestmark = pytest.mark.gpu
f = KoalaBear


@pytest.fixture(scope="module")
def setup():
    assert torch.cuda.is_available() and _lib.LIB_PATH.exists()
    gpu = default_gpu(0)
    rng = O.SmallRng(1)
    air = O.air_from_rng(f.id, rng)                      # RoundConstants::from_rng(&mut SmallRng::seed_from_u64(1)), as the example binary
    gpu.p2air_set_constants(f.id, np.array(air.beg), np.array(air.part)[: air.rounds_p], np.array(air.end))
    return gpu, air


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("n_perms,vec", [(8, 8), (64, 8), (4096, 8), (1 << 15, 8), (16, 1), (32, 2)])
def test_trace_generation_matches_oracle(setup, n_perms, vec):
    gpu, air = setup
    inputs = O.SmallRng(1).field(f.id, n_perms * 16).reshape(n_perms, 16)          # generate_random_trace_rows' fixed-seed inputs
    exp = O.p2air_generate(air, inputs, vec)
    got = host(gpu.p2air_generate_trace(f.id, dev(inputs), vec))
    assert got.shape == exp.shape == (n_perms // vec, vec * 164)
    assert np.array_equal(got, exp)
    assert O.p2air_check(air, got, vec) == 0                                         # every constraint vanishes on the trace


def test_trace_edge_inputs(setup):
    gpu, air = setup
    inputs = np.zeros((16, 16), dtype=np.uint32)
    inputs[1] = f.P - 1
    inputs[2, 0] = f.ONE
    assert np.array_equal(host(gpu.p2air_generate_trace(f.id, dev(inputs), 8)), O.p2air_generate(air, inputs, 8))


@pytest.mark.parametrize("log_n,vec", [(3, 8), (6, 8), (10, 8), (12, 8), (5, 1), (7, 4)])
def test_quotient_matches_oracle(setup, log_n, vec):
    gpu, air = setup
    n_perms = vec << log_n
    inputs = O.random_matrix(f.id, n_perms, 16, seed=log_n)
    trace = O.p2air_generate(air, inputs, vec)
    lde = O.coset_lde_batch(f.id, trace, 1, f.generator, bitrev_out=True)
    alpha = O.random_matrix(f.id, 1, 4, seed=77)[0]
    exp = O.p2air_quotient(air, lde, log_n, alpha, vec)
    got = host(gpu.p2air_quotient(f.id, dev(lde), log_n, alpha, vec))
    assert np.array_equal(got, exp)
    # the quotient is a polynomial of degree < 2N: its coefficients over the coset vanish from 2N - 2 on (uni-stark/src/prover.rs:270-280)
    co = O.coset_idft_batch(f.id, got, f.generator)
    assert not co[(2 << log_n) - 2:].any()


def test_quotient_detects_a_broken_trace(setup):
    """A trace that violates a constraint does not give a low-degree quotient (sanity of the test above)."""
    gpu, air = setup
    trace = O.p2air_generate(air, O.random_matrix(f.id, 8 << 5, 16, seed=1), 8)
    trace[3, 200] ^= 1
    lde = O.coset_lde_batch(f.id, trace, 1, f.generator, bitrev_out=True)
    alpha = O.random_matrix(f.id, 1, 4, seed=78)[0]
    got = host(gpu.p2air_quotient(f.id, dev(lde), 5, alpha, 8))
    assert np.array_equal(got, O.p2air_quotient(air, lde, 5, alpha, 8))
    assert O.coset_idft_batch(f.id, got, f.generator)[62:].any()


def test_air_errors(setup):
    gpu, _ = setup
    with pytest.raises(_lib.P3GpuError):
        gpu.L.p3gpu_p2air_set_constants.restype  # noqa: B018 (touch)
        _lib.check(gpu.L.p3gpu_p2air_generate_trace_dev(gpu.h, 0, 0, 0, 0))          # BabyBear instance is not built
