"""Replay of the reference's committed proof fixture with a pluggable hot-path backend.

Re-derives, from first principles, every value of
uni-stark/tests/fixtures/uni_stark_two_adic_v1.postcard that the NTT -> Merkle -> FRI hot path
determines (golden copy: tests/golden/uni_stark_two_adic_v1.json, made by tools/extract_fixture.py).
Recipe (uni-stark/tests/fib_air.rs:134-155,193-198,435-443): BabyBear, Fibonacci AIR on an 8x2 trace,
pis [0,1,21], Poseidon2-16 constants from SmallRng::seed_from_u64(1), PaddingFreeSponge<16,8,8>,
TruncatedPermutation<2,8,16>, cap_height 0, DuplexChallenger<16,8>, log_blowup 2, arity 2, PoW 1 bit,
log_final_poly_len 2.

The three hot-path steps are delegated to `backend`:
    backend.lde(mat, added_bits, shift)  -> bit-reversed-row LDE   (Radix2DitParallel::coset_lde_batch)
    backend.commit(mats)                 -> cap digests            (MerkleTreeMmcs::commit, cap_height 0)
    backend.fold(vec_ef, log_arity, beta)-> folded vector          (TwoAdicFriFolding::fold_matrix)
Everything else (transcript, constraint evaluation, openings) is host-side scalar code written here in
plain Python integers; it is the caller of the hot path, not part of it.
"""
import copy

import numpy as np

P = 0x78000001
M64 = (1 << 64) - 1
W = 11
GEN = 31
TOP27 = 0x1A427A41
R = (1 << 32) % P
RINV = pow(R, P - 2, P)


def inv(v): return pow(v, P - 2, P)
def br(i, b): return int(format(i, "0%db" % b)[::-1], 2) if b else 0
def gen(bits): return pow(TOP27, 1 << (27 - bits), P)
def to_m(x): return x * R % P
def from_m(x): return x * RINV % P


class SmallRng:
    """rand 0.10 SmallRng on 64-bit targets = xoshiro256++ seeded through SplitMix64."""

    def __init__(self, seed):
        self.s = []
        x = seed
        for _ in range(4):
            x = (x + 0x9E3779B97F4A7C15) & M64
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
            self.s.append(z ^ (z >> 31))

    def u32(self):
        a = self.s
        rotl = lambda x, k: ((x << k) | (x >> (64 - k))) & M64
        r = (rotl((a[0] + a[3]) & M64, 23) + a[0]) & M64
        t = (a[1] << 17) & M64
        a[2] ^= a[0]; a[3] ^= a[1]; a[1] ^= a[2]; a[0] ^= a[3]; a[2] ^= t; a[3] = rotl(a[3], 45)
        return r >> 32

    def field_monty(self):
        """monty_31.rs:154-165: the accepted 31-bit value IS the Montgomery representation."""
        while True:
            v = self.u32() >> 1
            if v < P:
                return v


def fixture_constants():
    """(rc_init 4x16, rc_term 4x16, rc_int 13) in Montgomery form (poseidon2/src/lib.rs:92-107)."""
    rng = SmallRng(1)
    rc_i = [[rng.field_monty() for _ in range(16)] for _ in range(4)]
    rc_t = [[rng.field_monty() for _ in range(16)] for _ in range(4)]
    rc_p = [rng.field_monty() for _ in range(13)]
    return rc_i, rc_t, rc_p


# ---- canonical-integer Poseidon2-16 (BabyBear, x^7) for the challenger: independent of the C oracle
def _ip(k): return inv(pow(2, k, P))
V16 = [P - 2, 1, 2, _ip(1), 3, 4, P - _ip(1), P - 3, P - 4, _ip(8), _ip(2), _ip(3), _ip(27), P - _ip(8), P - _ip(4), P - _ip(27)]


def _mat4(x):
    a, b, c, d = x
    return [(2 * a + 3 * b + c + d) % P, (a + 2 * b + 3 * c + d) % P, (a + b + 2 * c + 3 * d) % P, (3 * a + b + c + 2 * d) % P]


def _mds(s):
    s = sum((_mat4(s[i:i + 4]) for i in range(0, 16, 4)), [])
    t = [sum(s[j + k] for j in range(0, 16, 4)) % P for k in range(4)]
    return [(s[i] + t[i % 4]) % P for i in range(16)]


class PyPerm:
    def __init__(self, rc_i, rc_t, rc_p):  # Montgomery inputs -> canonical
        self.i = [[from_m(v) for v in r] for r in rc_i]
        self.t = [[from_m(v) for v in r] for r in rc_t]
        self.p = [from_m(v) for v in rc_p]

    def __call__(self, s):
        s = _mds(s)
        for rc in self.i: s = _mds([pow((s[i] + rc[i]) % P, 7, P) for i in range(16)])
        for rc in self.p:
            s = s[:]; s[0] = pow((s[0] + rc) % P, 7, P); t = sum(s) % P
            s = [(V16[i] * s[i] + t) % P for i in range(16)]
        for rc in self.t: s = _mds([pow((s[i] + rc[i]) % P, 7, P) for i in range(16)])
        return s


class Duplex:
    """challenger/src/duplex_challenger.rs:88-114,168-268 (canonical integers)."""

    def __init__(self, perm): self.perm = perm; self.st = [0] * 16; self.inb = []; self.out = []

    def duplex(self):
        n = len(self.inb)
        for i, v in enumerate(self.inb): self.st[i] = v
        self.inb = []
        if n:
            self.st[n:8] = [0] * (8 - n)
            self.st[8] = (self.st[8] + n) % P
        self.st = self.perm(self.st); self.out = self.st[:8]

    def observe(self, v):
        self.out = []; self.inb.append(v)
        if len(self.inb) == 8: self.duplex()

    def sample(self):
        if self.inb or not self.out: self.duplex()
        return self.out.pop()

    def sample_ef(self): return [self.sample() for _ in range(4)]


def emul(a, b):
    r = [sum(a[i] * b[k - i] for i in range(4) if 0 <= k - i < 4) for k in range(7)] + [0]
    return [(r[i] + W * r[i + 4]) % P if i < 3 else r[i] % P for i in range(4)]


def eadd(a, b): return [(x + y) % P for x, y in zip(a, b)]
def esub(a, b): return [(x - y) % P for x, y in zip(a, b)]
def escal(a, s): return [x * s % P for x in a]
def eF(x): return [x % P, 0, 0, 0]


def einv(a):
    e = P ** 4 - 2; r = [1, 0, 0, 0]
    while e:
        if e & 1: r = emul(r, a)
        a = emul(a, a); e >>= 1
    return r


def evalp_e(co, z):
    acc = [0, 0, 0, 0]
    for c in reversed(co): acc = eadd(emul(acc, z), eF(c))
    return acc


def _m_arr(rows):  # canonical python rows -> Montgomery uint32 matrix
    return np.array([[to_m(v) for v in r] for r in rows], dtype=np.uint32)


def _c_rows(arr):  # Montgomery matrix -> canonical python rows
    return [[from_m(int(v)) for v in r] for r in np.asarray(arr)]


def _commit(backend, mats):
    """(cap, prover data or None).  Backends with `commit_data`/`open_multi` also replay the query phase."""
    if hasattr(backend, "commit_data"):
        return backend.commit_data(mats)
    return backend.commit(mats), None


def replay(backend):
    """Returns a dict with the same keys/encoding (Montgomery u32) as the golden JSON.  With a backend that keeps prover data
    (`commit_data(mats) -> (cap, data)`, `open_multi(data, indices) -> (rows per matrix (n, w), full sibling paths (n, L, 8))`)
    the query phase is replayed too (fri/src/prover.rs:103-160,308-417) and the whole proof is returned in wire form under
    "postcard_hex" (plonky3_b200.proof_io on a plonky3_b200.uni_stark.Proof)."""
    rc_i, rc_t, rc_p = fixture_constants()
    perm = PyPerm(rc_i, rc_t, rc_p)
    w8 = gen(3); w32 = gen(5); w8i = inv(w8)
    rows = [(0, 1)]
    for _ in range(7): rows.append((rows[-1][1], (rows[-1][0] + rows[-1][1]) % P))

    # --- pcs.commit(trace): coset LDE onto GENERATOR*K (blowup 4), bit-reversed rows, Merkle commit
    trace_lde_m = backend.lde(_m_arr(rows), 2, to_m(GEN))           # two_adic_pcs.rs:300-324
    trace_cap, trace_data = _commit(backend, [trace_lde_m])
    lde = _c_rows(trace_lde_m)
    trace_root = [from_m(int(v)) for v in np.asarray(trace_cap)[0]]

    ch = Duplex(perm)
    for v in [3, 3, 0] + trace_root + [0, 1, 21]: ch.observe(v)      # uni-stark/src/prover.rs:224-236
    alpha = ch.sample_ef()

    # --- quotient on the coset GENERATOR*H (host-side constraint evaluation, not hot path)
    idft8 = lambda ev: [sum(ev[j] * pow(w8i, i * j, P) for j in range(8)) * inv(8) % P for i in range(8)]
    tcoef = [idft8([r[c] for r in rows]) for c in range(2)]
    # rows of the bit-reversed LDE with index < 8 are exactly the evaluations on GENERATOR*H (bit-reversed)
    T = [lde[br(i, 3)] for i in range(8)]
    xs = [GEN * pow(w8, i, P) % P for i in range(8)]
    Z = [(pow(x, 8, P) - 1) % P for x in xs]
    first = [Z[i] * inv((xs[i] - 1) % P) % P for i in range(8)]
    last = [Z[i] * inv((xs[i] - w8i) % P) % P for i in range(8)]
    trans = [(xs[i] - w8i) % P for i in range(8)]
    ap = [[1, 0, 0, 0]]
    for _ in range(4): ap.append(emul(ap[-1], alpha))
    Q = []
    for i in range(8):
        (l, r), (nl, nr) = T[i], T[(i + 1) % 8]
        cs = [first[i] * (l - 0) % P, first[i] * (r - 1) % P, trans[i] * (r - nl) % P, trans[i] * (l + r - nr) % P, last[i] * (r - 21) % P]
        acc = [0, 0, 0, 0]
        for k, c in enumerate(cs): acc = eadd(acc, escal(ap[4 - k], c))
        Q.append(escal(acc, inv(Z[i])))

    # --- pcs.commit_quotient: evals on GENERATOR*H -> coset_lde_batch(shift = 1) -> GENERATOR*K
    quot_lde_m = backend.lde(_m_arr(Q), 2, to_m(1))                  # two_adic_pcs.rs:326-345
    quot_cap, quot_data = _commit(backend, [quot_lde_m])
    qlde = _c_rows(quot_lde_m)
    quot_root = [from_m(int(v)) for v in np.asarray(quot_cap)[0]]
    for v in quot_root: ch.observe(v)
    zeta = ch.sample_ef(); zeta_n = escal(zeta, w8)

    # --- openings.  With backend.open (TwoAdicFriPcs::open's pre-FRI part on the hot-path backend: inverse denominators,
    #     barycentric interpolation of the low coset, alpha-compression + quotient accumulation) the values below come from
    #     the backend; otherwise they are evaluated here from the polynomial coefficients (pure Python).
    if hasattr(backend, "open"):
        class _Adapter:                                                # challenger protocol of plonky3_b200.fri
            def observe_algebra_slice(self, ys):
                if hasattr(ys, "cpu"):                                 # device-resident opened values
                    ys = ys.cpu().numpy().view(np.uint32)
                for y in np.asarray(ys).reshape(-1, 4):
                    for v in y: ch.observe(from_m(int(v)))
            def sample_algebra_element(self): return [to_m(v) for v in ch.sample_ef()]
        zm = np.array([to_m(v) for v in zeta], dtype=np.uint32); znm = np.array([to_m(v) for v in zeta_n], dtype=np.uint32)
        opened, fri_inputs = backend.open([([trace_lde_m], [[zm, znm]]), ([quot_lde_m], [[zm]])], _Adapter(), 2)
        y1, y2, y3 = _c_rows(opened[0][0][0]), _c_rows(opened[0][0][1]), _c_rows(opened[1][0][0])
        ro = _c_rows(fri_inputs[0])
    else:
        qcoef = [[c * pow(inv(GEN), i, P) % P for i, c in enumerate(idft8([Q[i][k] for i in range(8)]))] for k in range(4)]
        y1 = [evalp_e(c, zeta) for c in tcoef]; y2 = [evalp_e(c, zeta_n) for c in tcoef]; y3 = [evalp_e(c, zeta) for c in qcoef]
        for ys in (y1, y2, y3):
            for y in ys:
                for v in y: ch.observe(v)
        al = ch.sample_ef(); alp = [[1, 0, 0, 0]]
        for _ in range(8): alp.append(emul(alp[-1], al))
        xb = [GEN * pow(w32, br(i, 5), P) % P for i in range(32)]
        ro = [[0] * 4 for _ in range(32)]; nred = 0
        for mat, wd, z, ys in ((lde, 2, zeta, y1), (lde, 2, zeta_n, y2), (qlde, 4, zeta, y3)):   # two_adic_pcs.rs:606-661
            yred = [0] * 4
            for i in range(wd): yred = eadd(yred, emul(alp[i], ys[i]))
            for r in range(32):
                rr = [0] * 4
                for i in range(wd): rr = eadd(rr, escal(alp[i], mat[r][i]))
                ro[r] = eadd(ro[r], emul(alp[nred], emul(esub(yred, rr), einv(esub(z, eF(xb[r]))))))
            nred += wd

    # --- FRI commit phase round 0: commit (16 x 2 EF = 16 x 8 base), grind, beta, fold   (fri/src/prover.rs:219-266)
    ro_m = _m_arr(ro)                                                 # (32, 4)
    fri_cap, fri_data = _commit(backend, [ro_m.reshape(16, 8)])
    fri_root = [from_m(int(v)) for v in np.asarray(fri_cap)[0]]
    for v in fri_root: ch.observe(v)
    for cand in range(P):                                             # grind(1): smallest witness (serial build)
        c2 = copy.deepcopy(ch); c2.observe(cand)
        if c2.sample() & 1 == 0:
            wit = cand; ch = c2; break
    beta = ch.sample_ef()
    fold_m = backend.fold(ro_m, 1, np.array([to_m(v) for v in beta], dtype=np.uint32))
    fold = _c_rows(fold_m)
    f4 = [fold[br(i, 2)] for i in range(4)]; w4i = inv(gen(2))
    final = [[sum(f4[j][c] * pow(w4i, i * j, P) for j in range(4)) * inv(4) % P for c in range(4)] for i in range(4)]

    mm = lambda rows_: [[to_m(v) for v in r] for r in rows_]
    query = {}
    if trace_data is not None:
        # --- query phase: bind final poly + arity schedule, grind, sample the indices, open everything   (fri/src/prover.rs:103-160)
        for co in final:
            for v in co: ch.observe(v)
        ch.observe(1)                                                 # log_arity of the single round
        for cand in range(P):                                         # grind(query_proof_of_work_bits = 1)
            c2 = copy.deepcopy(ch); c2.observe(cand)
            if c2.sample() & 1 == 0:
                qwit = cand; ch = c2; break
        indices = [ch.sample() & 31 for _ in range(2)]                # sample_bits(log_global_max_height = 5)
        from plonky3_b200.merkle_tree import prune_paths
        from plonky3_b200.uni_stark import Proof
        input_openings = [backend.open_multi(trace_data, indices), backend.open_multi(quot_data, indices)]
        group = [i >> 1 for i in indices]
        rows, paths = backend.open_multi(fri_data, group)
        opened = np.asarray(rows[0], dtype=np.uint32).reshape(2, 2, 4)
        siblings = np.array([[opened[q][(i & 1) ^ 1]] for q, i in enumerate(indices)], dtype=np.uint32)     # (2, arity - 1, 4)
        proof = Proof(trace_commit=np.asarray(trace_cap), quotient_commit=np.asarray(quot_cap), trace_local=_m_arr(y1),
                      quotient_chunks=[_m_arr(y3)], commit_phase_commits=[np.asarray(fri_cap)], commit_pow_witnesses=[to_m(wit)],
                      final_poly=_m_arr(final), query_pow_witness=to_m(qwit), query_indices=indices, input_openings=input_openings,
                      commit_phase_openings=[(1, siblings, paths)], degree_bits=3, trace_next=_m_arr(y2),
                      input_opening_indices=[indices, indices], commit_phase_indices=[group])
        ints = lambda a: [[int(v) for v in r] for r in np.asarray(a).reshape(-1, np.asarray(a).shape[-1])]
        query = {
            "query_pow_witness": to_m(qwit),
            "input_openings": [{"opened_values": [[[int(v) for v in m[q]] for m in r] for q in range(2)], "proof": ints(prune_paths(indices, pth))}
                               for r, pth in input_openings],
            "commit_phase_openings": [{"log_arity": 1, "sibling_values": [ints(siblings[q]) for q in range(2)], "proof": ints(prune_paths(group, paths))}],
            "postcard_hex": proof.to_postcard().hex(),
        }
    return {
        **query,
        "trace_cap": [[to_m(v) for v in trace_root]],
        "quotient_cap": [[to_m(v) for v in quot_root]],
        "trace_local": mm(y1), "trace_next": mm(y2), "quotient_chunks": [mm(y3)],
        "commit_phase_commits": [[[to_m(v) for v in fri_root]]],
        "commit_pow_witnesses": [to_m(wit)],
        "final_poly": mm(final),
    }
