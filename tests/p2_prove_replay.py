"""CPU replay of uni-stark `prove` for the vectorised Poseidon2 AIR (the BASELINE config-5 statement at small sizes), built only
from the CPU oracle (oracle/p3_oracle.c) — test infrastructure.  It restates, step by step and with the reference's transcript
order, what plonky3_b200.uni_stark.prove does on the GPU:

    uni-stark/src/prover.rs:87-442 (prove_with_preprocessed), fri/src/two_adic_pcs.rs:413-662 (open), fri/src/prover.rs:43-417
    (prove_fri, commit_phase, answer_queries, open_inputs), challenger/src/duplex_challenger.rs:60-300,
    challenger/src/grinding_challenger.rs:100-232 (serial semantics: the smallest witness).

Parity anchor: the generic flow (LDE layout, MMCS, challenger, alpha ordering, open, FRI round, PoW, final poly) is the one pinned by
the reference's committed proof fixture (tests/fixture_replay.py); the AIR-specific parts (trace columns, constraint order) follow
poseidon2-air/src/{columns,generation,air,vectorized}.rs and are additionally checked by `verify_constraints_at_zeta`, a restatement
of the verifier's identity C(zeta) / Z_H(zeta) = Q(zeta) (uni-stark/src/verifier.rs:98-220) in plain Python integers.
PARITY UNPINNED by a reference artifact for this AIR: the reference holds no proof fixture of prove_prime_field_31.  The proof in
wire form (`to_wire_proof(...).to_postcard()`) is accepted by tests/stark_verify.py, the restated verifier that accepts the
reference's own committed proof.
"""
import numpy as np

from oracle import p3_oracle as O

F = 1                      # KoalaBear
P = 0x7F000001
VEC = 8


def _log2(n):
    assert n & (n - 1) == 0
    return n.bit_length() - 1


class OracleChallenger:
    """DuplexChallenger<F, Perm24, 24, 16> on Montgomery words (duplex_challenger.rs:60-114,168-283)."""

    def __init__(self, perm, width=24, rate=16):
        self.perm, self.w, self.rate = perm, width, rate
        self.state = np.zeros(width, dtype=np.uint32)
        self.inb, self.out = [], []

    def clone(self):
        c = OracleChallenger(self.perm, self.w, self.rate)
        c.state, c.inb, c.out = self.state.copy(), list(self.inb), list(self.out)
        return c

    def duplexing(self):
        n = len(self.inb)
        for i, v in enumerate(self.inb):
            self.state[i] = v
        self.inb = []
        if n:
            self.state[n:self.rate] = 0
            self.state[self.rate] = O.add(F, int(self.state[self.rate]), O.to_monty(F, n))
        self.state = O.poseidon2_permute(self.perm, self.state)
        self.out = [int(v) for v in self.state[:self.rate]]

    def observe(self, v):
        self.out = []
        self.inb.append(int(v))
        if len(self.inb) == self.rate:
            self.duplexing()

    def observe_slice(self, vals):
        for v in np.asarray(vals, dtype=np.uint32).ravel():
            self.observe(v)

    def observe_canonical(self, x): self.observe(O.to_monty(F, x))

    def sample(self):
        if self.inb or not self.out:
            self.duplexing()
        return self.out.pop()

    def sample_ef(self): return np.array([self.sample() for _ in range(4)], dtype=np.uint32)
    def sample_bits(self, bits): return O.from_monty(F, self.sample()) & ((1 << bits) - 1)

    def grind(self, bits):
        """serial semantics: the smallest canonical witness whose check passes (grinding_challenger.rs:226-229)."""
        if bits == 0:
            return 0
        mask = (1 << bits) - 1
        widx = len(self.inb)
        base = self.state.copy()
        for i in range(self.rate):
            base[i] = self.inb[i] if i < widx else 0
        base[self.rate] = O.add(F, int(base[self.rate]), O.to_monty(F, widx + 1))
        for cand in range(P):
            s = base.copy()
            s[widx] = O.to_monty(F, cand)
            s = O.poseidon2_permute(self.perm, s)
            if O.from_monty(F, int(s[self.rate - 1])) & mask == 0:
                w = O.to_monty(F, cand)
                self.observe(w)
                assert self.sample_bits(bits) == 0
                return w
        raise AssertionError("no witness")


def merkle_open(layers, cap_height, mats, indices):
    """open_batch for many indices: rows per matrix + sibling paths up to the cap (mmcs/batch.rs:75-121)."""
    max_h = max(m.shape[0] for m in mats)
    log_max = _log2(max_h)
    rows = [np.array([m[i >> (log_max - _log2(m.shape[0]))] for i in indices], dtype=np.uint32) for m in mats]
    nl = len(layers)
    eff = min(cap_height, nl - 1)
    paths = np.array([[layers[l][(i >> l) ^ 1] for l in range(nl - 1 - eff)] for i in indices], dtype=np.uint32).reshape(len(indices), nl - 1 - eff, 8)
    return rows, paths


def prove(air, perm16, perm24, inputs, cap_height=3, log_blowup=1, max_log_arity=3, num_queries=100, query_pow_bits=16):
    """Returns a dict with every transcript-visible value of the proof (Montgomery words)."""
    hs = O.poseidon2_hasher(perm24, perm16)
    ch = OracleChallenger(perm24)
    g = O.generator(F)
    trace = O.p2air_generate(air, inputs, VEC)
    n = trace.shape[0]
    log_n = _log2(n)
    # pcs.commit([(H, trace)])
    trace_lde = O.coset_lde_batch(F, trace, log_blowup, g, bitrev_out=True)
    trace_layers = O.merkle_tree(hs, [trace_lde])
    trace_cap = O.merkle_cap(trace_layers, cap_height)
    ch.observe_canonical(log_n); ch.observe_canonical(log_n); ch.observe_canonical(0)
    ch.observe_slice(trace_cap)
    alpha = ch.sample_ef()
    # quotient over GENERATOR * K, |K| = 2N; commit_quotient in 2 chunks
    q = O.p2air_quotient(air, trace_lde, log_n, alpha, VEC)
    log_q = log_n + 1
    h = O.two_adic_generator(F, log_q)
    q_ldes = []
    for i in range(2):
        sub = np.ascontiguousarray(q[i::2])
        dshift = O.mul(F, g, O.fpow(F, h, i))
        q_ldes.append(O.coset_lde_batch(F, sub, log_blowup, O.mul(F, g, O.inv(F, dshift)), bitrev_out=True))
    q_layers = O.merkle_tree(hs, q_ldes)
    q_cap = O.merkle_cap(q_layers, cap_height)
    ch.observe_slice(q_cap)
    zeta = ch.sample_ef()
    # pcs.open: trace at zeta; both quotient chunks at zeta
    H = trace_lde.shape[0]
    log_H = _log2(H)
    inv_d = O.open_inv_denoms(F, log_H, zeta)
    opened = []
    for mats in ([trace_lde], q_ldes):
        per = []
        for m in mats:
            ys = O.interpolate_coset(F, m[: m.shape[0] >> log_blowup], zeta, inv_d)
            ch.observe_slice(ys)
            per.append(ys)
        opened.append(per)
    al = ch.sample_ef()
    ro = np.zeros((H, 4), dtype=np.uint32)
    nred = 0
    for mats, per in zip(([trace_lde], q_ldes), opened):
        for m, ys in zip(mats, per):
            r = O.rowwise_dot(F, m, al)
            yred = np.zeros(4, dtype=np.uint32)
            ap = O.ef_from_base(F, O.to_monty(F, 1))
            for y in ys:
                yred = O.ef_add(F, yred, O.ef_mul(F, ap, y)); ap = O.ef_mul(F, ap, al)
            ro = O.open_reduce(F, ro, r, inv_d, O.ef_pow(F, al, nred), yred)
            nred += m.shape[1]
    # prove_fri: commit phase
    log_final = log_blowup
    folded = ro
    commits, fri_data, arities = [], [], []
    while folded.shape[0] > (1 << log_final):
        la = O.compute_log_arity_for_round(_log2(folded.shape[0]), None, log_final, max_log_arity)
        arities.append(la)
        leaves = folded.reshape(folded.shape[0] >> la, 4 << la)
        layers = O.merkle_tree(hs, [leaves])
        cap = O.merkle_cap(layers, cap_height)
        ch.observe_slice(cap)
        commits.append(cap)
        beta = ch.sample_ef()                      # commit_proof_of_work_bits = 0: grind returns 0 without touching the transcript
        fri_data.append((leaves, layers))
        folded = O.fold_matrix(F, folded, la, beta)
    final_poly = folded[:1].copy()
    ch.observe_slice(final_poly)
    for la in arities:
        ch.observe_canonical(la)
    pow_witness = ch.grind(query_pow_bits)
    indices = [ch.sample_bits(log_H) for _ in range(num_queries)]
    input_openings = [merkle_open(trace_layers, cap_height, [trace_lde], indices), merkle_open(q_layers, cap_height, q_ldes, indices)]
    cpo, cur = [], list(indices)
    for la, (leaves, layers) in zip(arities, fri_data):
        group = [i >> la for i in cur]
        rows, paths = merkle_open(layers, cap_height, [leaves], group)
        opened_rows = rows[0].reshape(len(cur), 1 << la, 4)
        sib = np.array([[opened_rows[k][j] for j in range(1 << la) if j != (i & ((1 << la) - 1))] for k, i in enumerate(cur)], dtype=np.uint32)
        cpo.append((la, sib.reshape(len(cur), (1 << la) - 1, 4), paths))
        cur = group
    return {"trace_cap": trace_cap, "quotient_cap": q_cap, "alpha": alpha, "zeta": zeta, "trace_local": opened[0][0],
            "quotient_chunks": opened[1], "commit_phase_commits": commits, "log_arities": arities, "final_poly": final_poly,
            "query_pow_witness": pow_witness, "indices": indices, "input_openings": input_openings, "commit_phase_openings": cpo,
            "log_n": log_n}


def to_wire_proof(pr):
    """The replay's proof as a plonky3_b200.uni_stark.Proof (host-side container; `.to_postcard()` is the reference's wire form)."""
    from plonky3_b200.uni_stark import Proof
    cp_idx, cur = [], list(pr["indices"])
    for la in pr["log_arities"]:
        cur = [i >> la for i in cur]
        cp_idx.append(cur)
    return Proof(trace_commit=pr["trace_cap"], quotient_commit=pr["quotient_cap"], trace_local=pr["trace_local"],
                 quotient_chunks=pr["quotient_chunks"], commit_phase_commits=pr["commit_phase_commits"],
                 commit_pow_witnesses=[0] * len(pr["commit_phase_commits"]), final_poly=pr["final_poly"],
                 query_pow_witness=pr["query_pow_witness"], query_indices=pr["indices"], input_openings=pr["input_openings"],
                 commit_phase_openings=pr["commit_phase_openings"], degree_bits=pr["log_n"],
                 input_opening_indices=[pr["indices"]] * len(pr["input_openings"]), commit_phase_indices=cp_idx)


def verifier_config(perm16, perm24, cap_height=3, log_blowup=1, max_log_arity=3, num_queries=100, query_pow_bits=16):
    """tests/stark_verify.py configuration of the example binary's StarkConfig (examples/src/proofs.rs, new_benchmark_high_arity)."""
    return dict(hasher=O.poseidon2_hasher(perm24, perm16), challenger_perm=perm24, challenger_width=24, challenger_rate=16,
                log_blowup=log_blowup, log_final_poly_len=0, max_log_arity=max_log_arity, num_queries=num_queries,
                commit_pow_bits=0, query_pow_bits=query_pow_bits)


# ------------------------------------------------------------------------------------------------ verifier identity (plain integers)
RINV = pow(1 << 32, P - 2, P)
W = 3


def _c(x): return int(x) * RINV % P
def emul(a, b):
    r = [sum(a[i] * b[k - i] for i in range(4) if 0 <= k - i < 4) for k in range(7)] + [0]
    return [(r[i] + W * r[i + 4]) % P if i < 3 else r[i] % P for i in range(4)]
def eadd(a, b): return [(x + y) % P for x, y in zip(a, b)]
def esub(a, b): return [(x - y) % P for x, y in zip(a, b)]
def escal(a, s): return [x * s % P for x in a]
def ebase(x): return [x % P, 0, 0, 0]
def epow(a, e):
    r = [1, 0, 0, 0]
    while e:
        if e & 1: r = emul(r, a)
        a = emul(a, a); e >>= 1
    return r
def einv(a):
    # a^(p^4 - 2) would be slow; use the norm through the conjugates like the device code: solve a * x = 1 by linear algebra over F_p
    import itertools
    M = [[0] * 4 for _ in range(4)]
    for j in range(4):
        e = [0] * 4; e[j] = 1
        col = emul(a, e)
        for i in range(4): M[i][j] = col[i]
    rhs = [1, 0, 0, 0]
    for c in range(4):                                   # Gaussian elimination mod p
        piv = next(r for r in range(c, 4) if M[r][c])
        M[c], M[piv] = M[piv], M[c]; rhs[c], rhs[piv] = rhs[piv], rhs[c]
        iv = pow(M[c][c], P - 2, P)
        M[c] = [v * iv % P for v in M[c]]; rhs[c] = rhs[c] * iv % P
        for r in range(4):
            if r != c and M[r][c]:
                fct = M[r][c]
                M[r] = [(v - fct * w_) % P for v, w_ in zip(M[r], M[c])]; rhs[r] = (rhs[r] - fct * rhs[c]) % P
    return rhs


def _ip(k): return pow(pow(2, k, P), P - 2, P)
V16 = [P - 2, 1, 2, _ip(1), 3, 4, P - _ip(1), P - 3, P - 4, _ip(8), _ip(3), _ip(24), P - _ip(8), P - _ip(3), P - _ip(4), P - _ip(24)]   # koala-bear/src/poseidon2.rs:410-428


def _mat4(x):
    a, b, c, d = x
    return [eadd(eadd(escal(a, 2), escal(b, 3)), eadd(c, d)), eadd(eadd(a, escal(b, 2)), eadd(escal(c, 3), d)),
            eadd(eadd(a, b), eadd(escal(c, 2), escal(d, 3))), eadd(eadd(escal(a, 3), b), eadd(c, escal(d, 2)))]


def _mds(s):
    s = sum((_mat4(s[i:i + 4]) for i in range(0, 16, 4)), [])
    t = [[0, 0, 0, 0] for _ in range(4)]
    for i in range(16): t[i % 4] = eadd(t[i % 4], s[i])
    return [eadd(s[i], t[i % 4]) for i in range(16)]


def _cube(x): return emul(emul(x, x), x)


def verify_constraints_at_zeta(air, proof):
    """folded_constraints(zeta) * inv_vanishing(zeta) == recomposed quotient(zeta): uni-stark/src/verifier.rs:98-220 for this AIR,
    evaluated over EF4 with plain integers (independent of the C oracle's AIR code)."""
    log_n = proof["log_n"]
    alpha = [_c(v) for v in proof["alpha"]]; zeta = [_c(v) for v in proof["zeta"]]
    loc = [[_c(v) for v in row] for row in proof["trace_local"]]
    beg = [[_c(air.beg[r * 16 + i]) for i in range(16)] for r in range(4)]
    end = [[_c(air.end[r * 16 + i]) for i in range(16)] for r in range(4)]
    part = [_c(air.part[r]) for r in range(air.rounds_p)]
    cols = 144 + air.rounds_p
    acc = [0, 0, 0, 0]
    for v in range(VEC):
        c = loc[v * cols:(v + 1) * cols]
        s = _mds(c[:16]); k = 16
        for r in range(4):
            s = _mds([_cube(eadd(s[i], ebase(beg[r][i]))) for i in range(16)])
            for i in range(16):
                acc = eadd(emul(acc, alpha), esub(s[i], c[k + i])); s[i] = c[k + i]
            k += 16
        for r in range(air.rounds_p):
            x = _cube(eadd(s[0], ebase(part[r])))
            acc = eadd(emul(acc, alpha), esub(x, c[k])); s[0] = c[k]; k += 1
            t = [0, 0, 0, 0]
            for i in range(16): t = eadd(t, s[i])
            s = [eadd(escal(s[i], V16[i]), t) for i in range(16)]
        for r in range(4):
            s = _mds([_cube(eadd(s[i], ebase(end[r][i]))) for i in range(16)])
            for i in range(16):
                acc = eadd(emul(acc, alpha), esub(s[i], c[k + i])); s[i] = c[k + i]
            k += 16
    z_h = esub(epow(zeta, 1 << log_n), [1, 0, 0, 0])                              # trace domain H: shift 1
    lhs = emul(acc, einv(z_h))
    # quotient chunk domains: GENERATOR * h^i * K', |K'| = N, h = generator of the size-2N subgroup (domain.rs:243-255)
    g, h = 3, pow(_c(O.two_adic_generator(F, log_n + 1)), 1, P)
    shifts = [g * pow(h, i, P) % P for i in range(2)]
    def van(shift, x): return esub(epow(escal(x, pow(shift, P - 2, P)), 1 << log_n), [1, 0, 0, 0])
    rhs = [0, 0, 0, 0]
    for i in range(2):
        j = 1 - i
        zp = emul(van(shifts[j], zeta), einv(van(shifts[j], ebase(shifts[i]))))
        chv = [0, 0, 0, 0]
        for kk in range(4):                                                         # from_ext_basis_coefficients: sum_k X^k * ch[k]
            e = [0] * 4; e[kk] = 1
            chv = eadd(chv, emul(e, [_c(v) for v in proof["quotient_chunks"][i][kk]]))
        rhs = eadd(rhs, emul(zp, chv))
    return lhs == rhs
