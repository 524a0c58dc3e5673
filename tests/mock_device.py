"""Test double of plonky3_b200.gpu.Gpu backed by the CPU oracle — TEST INFRASTRUCTURE, never importable from the product.

It lets the CPU suite (`-m "not gpu"`) execute the host-side driver code of the product — plonky3_b200.uni_stark.prove / prove_fri,
fri.commit_phase (arity schedule, roll-in), TwoAdicFriPcs.commit / commit_quotient / get_evaluations_on_domain /
open_values_and_fri_inputs, MerkleTreeMmcs.commit / open_multi_batch, the wire serialiser — with every device call answered by the
oracle on CPU tensors.  What it checks is the SEQUENCING above the C ABI (which the GPU tests check again with the real kernels);
it says nothing about the kernels.  "Device" matrices are CPU int32 torch tensors."""
import numpy as np
import torch

from oracle import p3_oracle as O
from plonky3_b200 import _lib


def _n(t):
    if isinstance(t, torch.Tensor):
        return np.ascontiguousarray(t.contiguous().numpy().view(np.uint32))
    return np.ascontiguousarray(t, dtype=np.uint32)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32))


class MockGpu:
    """The methods of plonky3_b200.gpu.Gpu the prove driver reaches (see `grep gpu\\. plonky3_b200/*.py`)."""
    device = "mock"
    h = None
    L = None

    def __init__(self):
        self.perms, self.air, self.calls = {}, None, []

    def _note(self, name): self.calls.append(name)

    # ---- constants
    def poseidon2_set_constants(self, field, width, rc_initial, rc_terminal, rc_internal):
        self.perms[(field, width)] = O.make_perm(field, width, _n(rc_initial), _n(rc_terminal), _n(rc_internal), monty=True)

    def _hasher(self, field, hash_kind):
        if hash_kind == _lib.HASH_KECCAK:
            return O.keccak_hasher()
        leaf = self.perms[(field, 24 if hash_kind == _lib.HASH_POSEIDON2_W24 else 16)]
        return O.poseidon2_hasher(leaf, self.perms[(field, 16)])

    # ---- TwoAdicSubgroupDft
    def dft_batch(self, field, kind, mat, shift=0):
        self._note("dft_batch")
        m = _n(mat)
        out = {_lib.DFT: lambda: O.dft_batch(field, m), _lib.IDFT: lambda: O.idft_batch(field, m),
               _lib.COSET_DFT: lambda: O.coset_dft_batch(field, m, shift), _lib.COSET_IDFT: lambda: O.coset_idft_batch(field, m, shift)}[kind]()
        return _t(out) if isinstance(mat, torch.Tensor) else out

    def coset_lde_batch(self, field, mat, added_bits, shift, bitrev_rows=True):
        self._note("coset_lde_batch")
        out = O.coset_lde_batch(field, _n(mat), added_bits, shift, bitrev_out=bool(bitrev_rows))
        return _t(out) if isinstance(mat, torch.Tensor) else out

    # ---- Mmcs
    def merkle_commit(self, field, hash_kind, mats):
        self._note("merkle_commit")
        return [np.asarray(l, dtype=np.uint32) for l in O.merkle_tree(self._hasher(field, hash_kind), [_n(m) for m in mats])]

    # ---- FRI
    def fri_fold(self, field, vec_ef, log_arity, beta):
        self._note("fri_fold")
        return _t(O.fold_matrix(field, _n(vec_ef).reshape(-1, 4), log_arity, _n(beta)))

    def ef_axpy(self, field, acc, x, s):
        self._note("ef_axpy")
        a, b = _n(acc).reshape(-1, 4), _n(x).reshape(-1, 4)
        out = np.array([O.ef_add(field, a[i], O.ef_mul(field, s, b[i])) for i in range(a.shape[0])], dtype=np.uint32)
        acc.copy_(_t(out).reshape(acc.shape))
        return acc

    # ---- Pcs::open, pre-FRI part
    def open_inv_denoms(self, field, log_height, z, zinv=None):
        self._note("open_inv_denoms")
        inv = O.open_inv_denoms(field, log_height, _n(z))
        adj = None if zinv is None else np.array([O.ef_sub(field, row, _n(zinv)) for row in inv], dtype=np.uint32)
        return _t(inv), (None if adj is None else _t(adj))

    def columnwise_dot(self, field, mat, vec_ef, scale=None):
        self._note("columnwise_dot")
        m = _n(mat)
        out = O.columnwise_dot(field, m, _n(vec_ef)[: m.shape[0]])
        if scale is not None:
            out = np.array([O.ef_mul(field, _n(scale), e) for e in out], dtype=np.uint32)
        return _t(out)

    def rowwise_dot(self, field, mat, alpha):
        self._note("rowwise_dot")
        return _t(O.rowwise_dot(field, _n(mat), _n(alpha)))

    def open_reduce(self, field, ro, r, inv_denoms, coeff, yred):
        self._note("open_reduce")
        h = int(ro.shape[0])
        ro.copy_(_t(O.open_reduce(field, _n(ro), _n(r), _n(inv_denoms)[:h], _n(coeff), _n(yred))))
        return ro

    # ---- Poseidon2 AIR
    def p2air_set_constants(self, field, beginning_full, partial, ending_full):
        self.air = O.make_air(field, _n(beginning_full), _n(partial), _n(ending_full))

    def p2air_generate_trace(self, field, inputs, vector_len=8):
        self._note("p2air_generate_trace")
        return _t(O.p2air_generate(self.air, _n(inputs), vector_len))

    def p2air_quotient(self, field, lde, log_trace_height, alpha, vector_len=8):
        self._note("p2air_quotient")
        return _t(O.p2air_quotient(self.air, _n(lde), log_trace_height, _n(alpha), vector_len))


class MockChallenger:
    """The surface of plonky3_b200.challenger.DuplexChallenger on the replay's oracle challenger (KoalaBear, Montgomery words)."""

    def __init__(self, perm24):
        import p2_prove_replay as R
        self.c = R.OracleChallenger(perm24)

    def observe(self, word): self.c.observe(int(word))
    def observe_canonical(self, x): self.c.observe_canonical(int(x))
    def observe_slice(self, values): self.c.observe_slice(_n(values))
    def observe_cap(self, cap): self.c.observe_slice(_n(cap))
    def observe_algebra_slice(self, ys): self.c.observe_slice(_n(ys))
    def sample_algebra_element(self): return self.c.sample_ef()
    def sample_bits(self, bits): return self.c.sample_bits(bits)
    def grind(self, bits): return self.c.grind(bits)
