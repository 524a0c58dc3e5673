"""Host-side descriptions of the two 31-bit Montgomery fields (baby-bear/src/baby_bear.rs:14-65,
koala-bear/src/koala_bear.rs:14-91).  Scalars crossing the C ABI are Montgomery u32 (MontyField31.value)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


_RINV = {}


def _rinv(p: int) -> int:
    """2^-32 mod p, cached (every Montgomery <-> canonical conversion of the host-side scalar code needs it)."""
    r = _RINV.get(p)
    if r is None:
        r = _RINV[p] = pow(1 << 32, p - 2, p)
    return r


@dataclass(frozen=True)
class Field:
    id: int            # P3GPU_BABY_BEAR / P3GPU_KOALA_BEAR
    name: str
    P: int
    GENERATOR: int     # canonical
    TWO_ADICITY: int
    TOP_ROOT: int      # canonical generator of the 2^TWO_ADICITY subgroup
    EXT_W: int         # EF4 = F[X]/(X^4 - W)
    SBOX_D: int

    # ---- canonical <-> Montgomery
    def to_monty(self, x: int) -> int: return (x % self.P) * (1 << 32) % self.P
    def from_monty(self, m: int) -> int: return m * _rinv(self.P) % self.P
    @property
    def ONE(self) -> int: return self.to_monty(1)

    def to_monty_array(self, a) -> np.ndarray:
        a = np.asarray(a, dtype=np.uint64) % np.uint64(self.P)
        return ((a << np.uint64(32)) % np.uint64(self.P)).astype(np.uint32)

    def from_monty_array(self, a) -> np.ndarray:
        rinv = _rinv(self.P)
        a = np.asarray(a, dtype=np.uint64)
        return (a * np.uint64(rinv) % np.uint64(self.P)).astype(np.uint32)

    # ---- Montgomery-domain scalar arithmetic for host logic (shifts, domains)
    def mul(self, a: int, b: int) -> int: return self.to_monty(self.from_monty(a) * self.from_monty(b))
    def inv(self, a: int) -> int: return self.to_monty(pow(self.from_monty(a), self.P - 2, self.P))
    def pow(self, a: int, e: int) -> int: return self.to_monty(pow(self.from_monty(a), e, self.P))
    def div(self, a: int, b: int) -> int: return self.mul(a, self.inv(b))

    def two_adic_generator(self, bits: int) -> int:
        """monty_31.rs:709-726 (Montgomery form)."""
        if bits > self.TWO_ADICITY:
            raise ValueError(f"bits {bits} exceeds two-adicity {self.TWO_ADICITY}")
        return self.to_monty(pow(self.TOP_ROOT, 1 << (self.TWO_ADICITY - bits), self.P))

    @property
    def generator(self) -> int: return self.to_monty(self.GENERATOR)


BabyBear = Field(0, "baby_bear", 0x78000001, 31, 27, 0x1A427A41, 11, 7)
KoalaBear = Field(1, "koala_bear", 0x7F000001, 3, 24, 0x6AC49F88, 3, 3)
FIELDS = {0: BabyBear, 1: KoalaBear, "baby_bear": BabyBear, "koala_bear": KoalaBear}
