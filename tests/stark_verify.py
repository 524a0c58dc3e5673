"""CPU restatement of the reference VERIFIER for the proofs this repo produces — test infrastructure (uses the CPU oracle for
hashing).  It reads a proof in the reference's wire form (plonky3_b200.proof_io.proof_from_postcard) and accepts or rejects it:

    uni-stark/src/verifier.rs:282-561   verify_with_preprocessed: shape checks, transcript, zeta, rounds, quotient identity
    fri/src/two_adic_pcs.rs:684-715     TwoAdicFriPcs::verify: observe the opened values
    fri/src/verifier.rs:158-436         verify_fri: alpha, betas (+ commit PoW), final poly, arities, query PoW, indices, fold chains,
                                        one multi-opening check per round
    fri/src/verifier.rs:471-606         fold_query;   :617-833  open_inputs (reduced openings from the authenticated rows)
    fri/src/two_adic_pcs.rs:108-131     fold_row (Lagrange interpolation over the arity-point coset — NOT the butterfly form the
                                        prover's fold_matrix uses, so prover and verifier are checked against each other)
    merkle-tree/src/mmcs/mod.rs:430-    verify_batch_pruned: amortised frontier walk over the pruned multiproof
    challenger/src/{duplex_challenger,grinding_challenger}.rs

All field arithmetic is on canonical Python integers, independent of the prover-side code paths (C oracle and GPU).
Pinned: it accepts the reference's own committed proof fixture (tests/golden/uni_stark_two_adic_v1.json, produced and verified by
the reference, uni-stark/tests/fib_air.rs:401-422) and rejects every single-field corruption of it (tests/test_oracle.py)."""
import numpy as np

from oracle import p3_oracle as O


class VerifyError(Exception):
    pass


def _need(cond, msg):
    if not cond:
        raise VerifyError(msg)


class Fld:
    """One 31-bit field + its quartic extension F[X]/(X^4 - W), canonical integers."""

    def __init__(self, fid):
        self.id, self.P = fid, O.prime(fid)
        self.W = 11 if fid == 0 else 3                         # baby_bear.rs:68, koala_bear.rs:94
        self.GEN = O.from_monty(fid, O.generator(fid))
        self.RINV = pow(1 << 32, self.P - 2, self.P)
        self.two_adicity = 27 if fid == 0 else 24

    def c(self, m): return int(m) * self.RINV % self.P        # Montgomery word -> canonical
    def m(self, x): return (int(x) << 32) % self.P
    def inv(self, x): return pow(x % self.P, self.P - 2, self.P)
    def root(self, bits): return O.from_monty(self.id, O.two_adic_generator(self.id, bits))

    def emul(self, a, b):
        r = [0] * 7
        for i in range(4):
            for j in range(4):
                r[i + j] += a[i] * b[j]
        return [(r[0] + self.W * r[4]) % self.P, (r[1] + self.W * r[5]) % self.P, (r[2] + self.W * r[6]) % self.P, r[3] % self.P]

    def eadd(self, a, b): return [(x + y) % self.P for x, y in zip(a, b)]
    def esub(self, a, b): return [(x - y) % self.P for x, y in zip(a, b)]
    def escal(self, a, s): return [x * s % self.P for x in a]
    def ebase(self, x): return [x % self.P, 0, 0, 0]

    def epow(self, a, e):
        r = [1, 0, 0, 0]
        while e:
            if e & 1:
                r = self.emul(r, a)
            a = self.emul(a, a); e >>= 1
        return r

    def einv(self, a):
        """through the Frobenius conjugates: a^-1 = (conj1 conj2 conj3) / Norm(a)."""
        P = self.P
        zeta = pow(self.W, (P - 1) // 4, P)
        conj = lambda k: [a[i] * pow(zeta, i * k, P) % P for i in range(4)]
        b = self.emul(self.emul(conj(1), conj(2)), conj(3))
        n = self.emul(a, b)
        _need(n[1] == n[2] == n[3] == 0 and n[0] != 0, "inverse of zero")
        return self.escal(b, self.inv(n[0]))

    def from_basis(self, coeffs):
        """sum_k X^k * coeffs[k] (from_ext_basis_coefficients): the flattened extension columns of the quotient chunks."""
        acc = [0, 0, 0, 0]
        for k, v in enumerate(coeffs):
            e = [0] * 4; e[k] = 1
            acc = self.eadd(acc, self.emul(e, v))
        return acc


def _rev(i, bits): return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0


class Challenger:
    """DuplexChallenger<F, Perm, WIDTH, RATE> + GrindingChallenger::check_witness, canonical integers in and out."""

    def __init__(self, fld, perm, width, rate):
        self.f, self.perm, self.w, self.rate = fld, perm, width, rate
        self.state, self.inb, self.out = [0] * width, [], []

    def _duplex(self):
        n = len(self.inb)
        for i, v in enumerate(self.inb):
            self.state[i] = v
        self.inb = []
        if n:                                                    # duplex_challenger.rs:88-114: zero the unused rate, bind the count
            for i in range(n, self.rate):
                self.state[i] = 0
            self.state[self.rate] = (self.state[self.rate] + n) % self.f.P
        st = np.array([self.f.m(v) for v in self.state], dtype=np.uint32)
        self.state = [self.f.c(v) for v in O.poseidon2_permute(self.perm, st)]
        self.out = list(self.state[:self.rate])

    def observe(self, v):
        self.out = []
        self.inb.append(v % self.f.P)
        if len(self.inb) == self.rate:
            self._duplex()

    def observe_words(self, words):                              # Montgomery words (digests, opened values)
        for v in np.asarray(words, dtype=np.uint32).ravel():
            self.observe(self.f.c(v))

    def sample(self):
        if self.inb or not self.out:
            self._duplex()
        return self.out.pop()

    def sample_ef(self): return [self.sample() for _ in range(4)]
    def sample_bits(self, bits): return self.sample() & ((1 << bits) - 1)

    def check_witness(self, bits, witness):
        if bits == 0:
            return True
        self.observe(witness)
        return self.sample_bits(bits) == 0


def verify_multi_batch(hs, cap, dims, indices, opened_values, pruned):
    """verify_batch_pruned for matrices of ONE height (all this repo's batches): hash every distinct opened leaf, then fold the
    frontier upwards, taking a sibling from the proof only where no queried leaf covers it; every digest must be consumed and
    every top node must equal its cap entry.  `dims`: [(width, height)]; opened_values[q][m] = row (Montgomery words)."""
    heights = {h for _, h in dims}
    _need(len(heights) == 1, "mixed-height batches are not used by this prover")
    height = heights.pop()
    log_h = height.bit_length() - 1
    _need(1 << log_h == height, "height must be a power of two")
    cap = np.asarray(cap, dtype=np.uint32).reshape(-1, 8)
    log_cap = cap.shape[0].bit_length() - 1
    _need(1 << log_cap == cap.shape[0] and log_cap <= log_h, "bad cap")
    _need(len(opened_values) == len(indices), "one opened row set per query")
    nodes = {}
    for i, rows in zip(indices, opened_values):
        _need(0 <= i < height, "index out of range")
        _need(len(rows) == len(dims) and all(len(r) == w for r, (w, _) in zip(rows, dims)), "opened row widths do not match the claimed dimensions")
        leaf = np.concatenate([np.asarray(r, dtype=np.uint32) for r in rows])
        d = O.hash_row(hs, leaf)
        if i in nodes:
            _need(np.array_equal(nodes[i], d), "two openings of one leaf disagree")
        nodes[i] = d
    pruned = np.asarray(pruned, dtype=np.uint32).reshape(-1, 8)
    k = 0
    for _ in range(log_h - log_cap):
        parents = {}
        for idx in sorted(nodes):
            if (idx >> 1) in parents:
                continue
            if (idx ^ 1) in nodes:
                sib = nodes[idx ^ 1]
            else:
                _need(k < pruned.shape[0], "multiproof too short")
                sib = pruned[k]; k += 1
            left, right = (nodes[idx], sib) if idx & 1 == 0 else (sib, nodes[idx])
            parents[idx >> 1] = O.compress(hs, left, right)
        nodes = parents
    _need(k == pruned.shape[0], "multiproof has unused digests")
    for idx, d in nodes.items():
        _need(np.array_equal(cap[idx], d), "cap mismatch")


def fold_row(f, index, log_height, log_arity, beta, evals):
    """two_adic_pcs.rs:108-131: interpolate the arity evaluations over their coset and evaluate at beta."""
    arity = 1 << log_arity
    start = pow(f.root(log_height + log_arity), _rev(index, log_height), f.P)
    w = f.root(log_arity)
    xs = [start * pow(w, k, f.P) % f.P for k in range(arity)]
    xs = [xs[_rev(k, log_arity)] for k in range(arity)]
    acc = [0, 0, 0, 0]
    for j in range(arity):                                       # plain Lagrange form
        num, den = [1, 0, 0, 0], 1
        for k in range(arity):
            if k != j:
                num = f.emul(num, f.esub(beta, f.ebase(xs[k])))
                den = den * (xs[j] - xs[k]) % f.P
        acc = f.eadd(acc, f.emul(evals[j], f.escal(num, f.inv(den))))
    return acc


def verify_fri(f, cfg, proof, ch, rounds):
    """`rounds`: [(cap, [(log_domain_size, [(z, values_at_z)])])] in commitment order; values canonical EF lists."""
    hs = cfg["hasher"]
    _need(cfg["num_queries"] > 0, "zero queries")
    alpha = ch.sample_ef()
    cpo = proof["commit_phase_openings"]
    _need(len(cpo) == len(proof["commit_phase_commits"]), "commit phase opening count")
    log_arities = []
    for o in cpo:
        _need(1 <= o["log_arity"] <= cfg["max_log_arity"], "invalid log arity")
        _need(len(o["sibling_values"]) == cfg["num_queries"], "commit phase query count")
        _need(all(len(s) == (1 << o["log_arity"]) - 1 for s in o["sibling_values"]), "sibling count")
        log_arities.append(o["log_arity"])
    log_final_height = cfg["log_blowup"] + cfg["log_final_poly_len"]
    log_max = sum(log_arities) + log_final_height
    _need(log_max <= f.two_adicity, "global height exceeds the two-adicity")
    expected = max(ld + cfg["log_blowup"] for _, mats in rounds for ld, _ in mats)
    _need(expected == log_max, "global max height mismatch")
    _need(len(proof["commit_pow_witnesses"]) == len(proof["commit_phase_commits"]), "commit PoW witness count")
    betas = []
    for cap, wit in zip(proof["commit_phase_commits"], proof["commit_pow_witnesses"]):
        ch.observe_words(cap)
        _need(ch.check_witness(cfg["commit_pow_bits"], f.c(wit)), "invalid commit PoW witness")
        betas.append(ch.sample_ef())
    final_poly = [[f.c(v) for v in co] for co in proof["final_poly"]]
    _need(len(final_poly) == 1 << cfg["log_final_poly_len"], "final polynomial length")
    ch.observe_words(proof["final_poly"])
    for la in log_arities:
        ch.observe(la)
    _need(ch.check_witness(cfg["query_pow_bits"], f.c(proof["query_pow_witness"])), "invalid query PoW witness")
    indices = [ch.sample_bits(log_max) for _ in range(cfg["num_queries"])]

    # open_inputs: authenticate the rows, then the reduced openings per height
    io = proof["input_openings"]
    _need(len(io) == len(rounds), "input batch count")
    for b, (cap, mats) in zip(io, rounds):
        dims = [(len(pts[0][1]), 1 << (ld + cfg["log_blowup"])) for ld, pts in mats]
        lh = max(h for _, h in dims).bit_length() - 1
        verify_multi_batch(hs, cap, dims, [i >> (log_max - lh) for i in indices], b["opened_values"], b["proof"])
    reduced = []
    for q, index in enumerate(indices):
        ro = {}
        for b, (_, mats) in zip(io, rounds):
            for row, (ld, pts) in zip(b["opened_values"][q], mats):
                lh = ld + cfg["log_blowup"]
                x = f.GEN * pow(f.root(lh), _rev(index >> (log_max - lh), lh), f.P) % f.P
                apow, acc = ro.get(lh, ([1, 0, 0, 0], [0, 0, 0, 0]))
                for z, ys in pts:
                    _need(len(ys) == len(row), "evaluation count")
                    den = f.esub(z, f.ebase(x))
                    _need(any(den), "query point equals the opening point")
                    quot = f.einv(den)
                    for px, pz in zip(row, ys):
                        acc = f.eadd(acc, f.emul(f.emul(apow, f.esub(pz, f.ebase(f.c(px)))), quot))
                        apow = f.emul(apow, alpha)
                ro[lh] = (apow, acc)
        if cfg["log_blowup"] in ro:
            _need(not any(ro[cfg["log_blowup"]][1]), "constant matrix quotient must vanish")
        reduced.append(sorted(((lh, v[1]) for lh, v in ro.items()), reverse=True))

    # fold chains
    groups = [[] for _ in cpo]
    rows_by_round = [[] for _ in cpo]
    for q, (index, ro) in enumerate(zip(indices, reduced)):
        _need(ro and ro[0][0] == log_max, "missing initial reduced opening")
        ro = list(ro)
        folded = ro.pop(0)[1]
        cur, idx = log_max, index
        for r, (beta, la, o) in enumerate(zip(betas, log_arities, cpo)):
            arity = 1 << la
            pos = idx % arity
            sib = [[f.c(v) for v in e] for e in o["sibling_values"][q]]
            evals = sib[:pos] + [folded] + sib[pos:]
            cur -= la
            idx >>= la
            folded = fold_row(f, idx, cur, la, beta, evals)
            groups[r].append(idx)
            rows_by_round[r].append([np.array([f.m(v) for e in evals for v in e], dtype=np.uint32)])
            if ro and ro[0][0] == cur:
                folded = f.eadd(folded, f.emul(f.epow(beta, arity), ro.pop(0)[1]))
        _need(cur == log_final_height, "final fold height")
        _need(not ro, "unconsumed reduced openings")
        x = pow(f.root(log_max), _rev(idx, log_max), f.P)
        ev = [0, 0, 0, 0]
        for co in reversed(final_poly):
            ev = f.eadd(f.escal(ev, x), co)
        _need(ev == folded, "final polynomial mismatch")

    cur = log_max
    for r, (cap, o, la) in enumerate(zip(proof["commit_phase_commits"], cpo, log_arities)):
        cur -= la
        verify_multi_batch(hs, cap, [(4 << la, 1 << cur)], groups[r], rows_by_round[r], o["proof"])


def verify(f, cfg, air, proof, public_values=()):
    """uni-stark verify.  `cfg`: dict(hasher, challenger_perm, challenger_width, challenger_rate, log_blowup, log_final_poly_len,
    max_log_arity, num_queries, commit_pow_bits, query_pow_bits).  `air`: dict(width, main_next, log_quotient_chunks,
    num_public_values, constraints(f, local, nxt, public_values, is_first, is_last, is_transition, alpha) -> folded EF).
    `proof`: proof_from_postcard(...).  Raises VerifyError."""
    db = proof["degree_bits"]
    _need(0 <= db and db + cfg["log_blowup"] <= f.two_adicity, "degree bits out of range")
    n = 1 << db
    nchunks = 1 << air["log_quotient_chunks"]
    _need(len(public_values) == air["num_public_values"], "public values length")
    _need(len(proof["trace_local"]) == air["width"], "trace_local width")
    if air["main_next"]:
        _need(proof["trace_next"] is not None and len(proof["trace_next"]) == air["width"], "trace_next width")
    else:
        _need(proof["trace_next"] is None, "unexpected trace_next")
    _need(len(proof["quotient_chunks"]) == nchunks and all(len(c) == 4 for c in proof["quotient_chunks"]), "quotient chunk shape")
    ch = Challenger(f, cfg["challenger_perm"], cfg["challenger_width"], cfg["challenger_rate"])
    ch.observe(db); ch.observe(db); ch.observe(0)
    ch.observe_words(proof["trace_commit"])
    for v in public_values:
        ch.observe(v)
    alpha = ch.sample_ef()
    ch.observe_words(proof["quotient_commit"])
    zeta = ch.sample_ef()
    z_h = f.esub(f.epow(zeta, n), [1, 0, 0, 0])
    _need(any(z_h), "out-of-domain point lies in the trace domain")
    g = f.root(db)
    zeta_next = f.escal(zeta, g)
    can = lambda rows: [[f.c(v) for v in e] for e in rows]
    local, nxt = can(proof["trace_local"]), (can(proof["trace_next"]) if air["main_next"] else [[0, 0, 0, 0]] * air["width"])
    chunks = [can(c) for c in proof["quotient_chunks"]]
    trace_pts = [(zeta, local)] + ([(zeta_next, nxt)] if air["main_next"] else [])
    rounds = [(proof["trace_commit"], [(db, trace_pts)]), (proof["quotient_commit"], [(db, [(zeta, c)]) for c in chunks])]
    # TwoAdicFriPcs::verify: all opened values enter the transcript first
    for _, mats in rounds:
        for _, pts in mats:
            for _, ys in pts:
                for e in ys:
                    for v in e:
                        ch.observe(v)
    verify_fri(f, cfg, proof, ch, rounds)
    # quotient(zeta) from the chunks (verifier.rs recompose_quotient_from_chunks; chunk i lives on GENERATOR * h^i * K, |K| = N)
    h = f.root(db + air["log_quotient_chunks"])
    shifts = [f.GEN * pow(h, i, f.P) % f.P for i in range(nchunks)]
    van = lambda s, x: f.esub(f.epow(f.escal(x, f.inv(s)), n), [1, 0, 0, 0])
    quotient = [0, 0, 0, 0]
    for i in range(nchunks):
        zp = [1, 0, 0, 0]
        for j in range(nchunks):
            if j != i:
                zp = f.emul(zp, f.emul(van(shifts[j], zeta), f.einv(van(shifts[j], f.ebase(shifts[i])))))
        quotient = f.eadd(quotient, f.emul(zp, f.from_basis(chunks[i])))
    # selectors at zeta on the trace domain (field/src/coset.rs selectors_at_point, shift 1)
    ginv = f.inv(g)
    is_first = f.emul(z_h, f.einv(f.esub(zeta, [1, 0, 0, 0])))
    is_last = f.emul(z_h, f.einv(f.esub(zeta, f.ebase(ginv))))
    is_trans = f.esub(zeta, f.ebase(ginv))
    folded = air["constraints"](f, local, nxt, list(public_values), is_first, is_last, is_trans, alpha)
    _need(f.emul(folded, f.einv(z_h)) == quotient, "out-of-domain evaluation mismatch")


# ------------------------------------------------------------------------------------------------------------------ the two AIRs
def fibonacci_air():
    """uni-stark/tests/fib_air.rs:33-75: columns (left, right); public values (a, b, x)."""
    def constraints(f, loc, nxt, pis, is_first, is_last, is_trans, alpha):
        l, r, nl, nr = loc[0], loc[1], nxt[0], nxt[1]
        cs = [f.emul(is_first, f.esub(l, f.ebase(pis[0]))), f.emul(is_first, f.esub(r, f.ebase(pis[1]))),
              f.emul(is_trans, f.esub(r, nl)), f.emul(is_trans, f.esub(f.eadd(l, r), nr)),
              f.emul(is_last, f.esub(r, f.ebase(pis[2])))]
        acc = [0, 0, 0, 0]
        for c in cs:
            acc = f.eadd(f.emul(acc, alpha), c)
        return acc
    return {"width": 2, "main_next": True, "log_quotient_chunks": 0, "num_public_values": 3, "constraints": constraints}


def poseidon2_air(air, vec=8):
    """VectorizedPoseidon2Air<KoalaBear, .., 16, 3, 1, 4, rounds_p, 8> (poseidon2-air/src/{air,columns,vectorized}.rs): S-box degree 3
    with no intermediate registers, so each round commits its post-state; no transition or boundary constraints."""
    def ip(f, k): return f.inv(pow(2, k, f.P))

    def constraints(f, loc, nxt, pis, is_first, is_last, is_trans, alpha):
        P = f.P
        v16 = [P - 2, 1, 2, ip(f, 1), 3, 4, P - ip(f, 1), P - 3, P - 4, ip(f, 8), ip(f, 3), ip(f, 24), P - ip(f, 8), P - ip(f, 3), P - ip(f, 4), P - ip(f, 24)]
        add, sc = f.eadd, f.escal

        def mat4(x):
            a, b, c, d = x
            return [add(add(sc(a, 2), sc(b, 3)), add(c, d)), add(add(a, sc(b, 2)), add(sc(c, 3), d)),
                    add(add(a, b), add(sc(c, 2), sc(d, 3))), add(add(sc(a, 3), b), add(c, sc(d, 2)))]

        def mds(s):
            s = sum((mat4(s[i:i + 4]) for i in range(0, 16, 4)), [])
            t = [[0, 0, 0, 0] for _ in range(4)]
            for i in range(16):
                t[i % 4] = add(t[i % 4], s[i])
            return [add(s[i], t[i % 4]) for i in range(16)]

        cube = lambda x: f.emul(f.emul(x, x), x)
        beg = [[f.c(air.beg[r * 16 + i]) for i in range(16)] for r in range(4)]
        end = [[f.c(air.end[r * 16 + i]) for i in range(16)] for r in range(4)]
        part = [f.c(air.part[r]) for r in range(air.rounds_p)]
        cols = 144 + air.rounds_p
        acc = [0, 0, 0, 0]
        for v in range(vec):
            c = loc[v * cols:(v + 1) * cols]
            s = mds(c[:16]); k = 16
            for rc in beg:
                s = mds([cube(add(s[i], f.ebase(rc[i]))) for i in range(16)])
                for i in range(16):
                    acc = add(f.emul(acc, alpha), f.esub(s[i], c[k + i])); s[i] = c[k + i]
                k += 16
            for r in range(air.rounds_p):
                x = cube(add(s[0], f.ebase(part[r])))
                acc = add(f.emul(acc, alpha), f.esub(x, c[k])); s[0] = c[k]; k += 1
                t = [0, 0, 0, 0]
                for i in range(16):
                    t = add(t, s[i])
                s = [add(sc(s[i], v16[i]), t) for i in range(16)]
            for rc in end:
                s = mds([cube(add(s[i], f.ebase(rc[i]))) for i in range(16)])
                for i in range(16):
                    acc = add(f.emul(acc, alpha), f.esub(s[i], c[k + i])); s[i] = c[k + i]
                k += 16
        return acc
    return {"width": vec * (144 + air.rounds_p), "main_next": False, "log_quotient_chunks": 1, "num_public_values": 0, "constraints": constraints}


# ------------------------------------------------------------------------------------------------------------------
# Oracle-backed stand-ins that let the PRODUCT verifier (plonky3_b200/verifier.py) run on the CPU: same verifier code as on the
# GPU, with hashing and the transcript permutation done by the C oracle instead of the device.  Test infrastructure.
class OracleMmcs:
    """The verifier-side surface of plonky3_b200.merkle_tree.MerkleTreeMmcs (hash_rows / compress_pairs / verify_multi_batch)."""

    def __init__(self, hasher):
        self.hs = hasher

    def hash_rows(self, rows): return np.array([O.hash_row(self.hs, r) for r in np.asarray(rows, dtype=np.uint32)], dtype=np.uint32)

    def compress_pairs(self, left, right):
        return np.array([O.compress(self.hs, l, r) for l, r in zip(np.asarray(left, dtype=np.uint32), np.asarray(right, dtype=np.uint32))], dtype=np.uint32)

    def verify_multi_batch(self, commit, dims, indices, opened_values, proof):
        from plonky3_b200.merkle_tree import verify_multi_batch_with
        verify_multi_batch_with(self.hash_rows, self.compress_pairs, commit, dims, indices, opened_values, proof)


class OracleDuplexChallenger:
    """The surface of plonky3_b200.challenger.DuplexChallenger the verifier uses, on the Challenger above (Montgomery words in/out)."""

    def __init__(self, fld, perm, width, rate):
        self.f, self.ch = fld, Challenger(fld, perm, width, rate)

    def observe(self, word): self.ch.observe(self.f.c(word))
    def observe_canonical(self, x): self.ch.observe(x)
    def observe_slice(self, words): self.ch.observe_words(words)
    def sample_algebra_element(self): return np.array([self.f.m(v) for v in self.ch.sample_ef()], dtype=np.uint32)
    def sample_bits(self, bits): return self.ch.sample_bits(bits)


def product_config(field, cfg):
    """A StarkConfig-shaped object for plonky3_b200.verifier.verify from the dict configuration used above."""
    from types import SimpleNamespace
    from plonky3_b200.fri import FriParameters
    mmcs = OracleMmcs(cfg["hasher"])
    fri = FriParameters(cfg["log_blowup"], cfg["log_final_poly_len"], cfg["max_log_arity"], cfg["num_queries"], cfg["commit_pow_bits"],
                        cfg["query_pow_bits"], mmcs)
    fld = Fld(field.id)
    return SimpleNamespace(pcs=SimpleNamespace(fri=fri, mmcs=mmcs, dft=SimpleNamespace(field=field)),
                           initialise_challenger=lambda: OracleDuplexChallenger(fld, cfg["challenger_perm"], cfg["challenger_width"], cfg["challenger_rate"]))


class FibonacciAir:
    """uni-stark/tests/fib_air.rs:33-75 in the AIR surface plonky3_b200.verifier.verify expects."""
    def width(self): return 2
    def num_public_values(self): return 3
    def main_next_row_columns(self): return [0, 1]
    def max_constraint_degree(self): return 2

    def eval_folded_constraints(self, e, loc, nxt, pis, is_first, is_last, is_trans, alpha):
        l, r, nl, nr = loc[0], loc[1], nxt[0], nxt[1]
        cs = [e.mul(is_first, e.sub(l, e.base(pis[0]))), e.mul(is_first, e.sub(r, e.base(pis[1]))),
              e.mul(is_trans, e.sub(r, nl)), e.mul(is_trans, e.sub(e.add(l, r), nr)), e.mul(is_last, e.sub(r, e.base(pis[2])))]
        acc = [0, 0, 0, 0]
        for c in cs:
            acc = e.add(e.mul(acc, alpha), c)
        return acc
