"""Host-side scalar arithmetic in EF4 = F[X]/(X^4 - W) (field/src/extension/binomial_extension.rs) for the handful of
transcript-level values `open` needs (1/z, z^N, alpha powers, Mred(z)).  Elements are length-4 sequences of Montgomery u32
words, exactly the wire/ABI representation; the arithmetic is done on canonical Python integers.  Not a compute path."""
from __future__ import annotations

import numpy as np

from .field import Field


def _c(field: Field, a): return [field.from_monty(int(v)) for v in a]
def _m(field: Field, a): return np.array([field.to_monty(int(v)) for v in a], dtype=np.uint32)


def ef_mul(field: Field, a, b):
    p, x, y = field.P, _c(field, a), _c(field, b)
    r = [0] * 7
    for i in range(4):
        for j in range(4):
            r[i + j] = (r[i + j] + x[i] * y[j]) % p
    return _m(field, [(r[i] + field.EXT_W * r[i + 4]) % p for i in range(3)] + [r[3]])


def ef_add(field: Field, a, b): return _m(field, [(x + y) % field.P for x, y in zip(_c(field, a), _c(field, b))])
def ef_sub(field: Field, a, b): return _m(field, [(x - y) % field.P for x, y in zip(_c(field, a), _c(field, b))])
def ef_scale(field: Field, a, s_monty: int): return _m(field, [x * field.from_monty(s_monty) % field.P for x in _c(field, a)])
def ef_from_base(field: Field, x_monty: int): return np.array([x_monty, 0, 0, 0], dtype=np.uint32)
def ef_one(field: Field): return ef_from_base(field, field.ONE)


def ef_pow(field: Field, a, e: int):
    r, a = ef_one(field), np.asarray(a, dtype=np.uint32)
    while e:
        if e & 1:
            r = ef_mul(field, r, a)
        a = ef_mul(field, a, a); e >>= 1
    return r


def ef_inv(field: Field, a):
    """a^-1 through the Frobenius conjugates: phi(X) = zeta X with zeta = W^((p-1)/4), so the conjugates of a are
    (a0, a1 zeta^k, a2 zeta^2k, a3 zeta^3k); b = conj1 conj2 conj3, N = a b lies in F, a^-1 = b / N."""
    p = field.P
    zeta = pow(field.EXT_W, (p - 1) // 4, p)
    x = _c(field, a)
    conj = lambda k: _m(field, [x[i] * pow(zeta, i * k, p) % p for i in range(4)])
    b = ef_mul(field, ef_mul(field, conj(1), conj(2)), conj(3))
    n = _c(field, ef_mul(field, a, b))
    assert n[1] == n[2] == n[3] == 0 and n[0] != 0, "norm must be a non-zero base-field element"
    return _m(field, [v * pow(n[0], p - 2, p) % p for v in _c(field, b)])


def ef_dot_powers(field: Field, alpha, ys):
    """sum_i alpha^i * ys[i] (dot_product(alpha.powers(), openings), two_adic_pcs.rs:636-637): Horner on canonical integers."""
    p, w = field.P, field.EXT_W
    a = _c(field, alpha)
    acc = [0, 0, 0, 0]
    for y in reversed([_c(field, v) for v in np.asarray(ys, dtype=np.uint32).reshape(-1, 4)]):
        r = [0] * 7
        for i in range(4):
            for j in range(4):
                r[i + j] += acc[i] * a[j]
        acc = [(r[0] + w * r[4] + y[0]) % p, (r[1] + w * r[5] + y[1]) % p, (r[2] + w * r[6] + y[2]) % p, (r[3] + y[3]) % p]
    return _m(field, acc)


def ef_from_basis_rows(field: Field, rows):
    """sum_k X^k * rows[k] for 4 EF4 values (from_ext_basis_coefficients): X * (a0, a1, a2, a3) = (W a3, a0, a1, a2)."""
    p, w = field.P, field.EXT_W
    acc = [0, 0, 0, 0]
    for k, r in enumerate(np.asarray(rows, dtype=np.uint32).reshape(4, 4)):
        v = _c(field, r)
        for _ in range(k):
            v = [w * v[3] % p, v[0], v[1], v[2]]
        acc = [(x + y) % p for x, y in zip(acc, v)]
    return _m(field, acc)
