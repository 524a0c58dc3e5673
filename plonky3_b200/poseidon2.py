"""Poseidon2 configuration objects (poseidon2/src/lib.rs:31-147).  Round constants are runtime inputs exactly as in
the reference (Poseidon2::new / new_from_rng); they are uploaded to a Gpu context in Montgomery form."""
from __future__ import annotations

import json
import pathlib
from dataclasses import dataclass

import numpy as np

from .field import Field

_CONSTS = None


def _defaults():
    global _CONSTS
    if _CONSTS is None:
        _CONSTS = json.loads((pathlib.Path(__file__).resolve().parent / "p2_constants.json").read_text())
    return _CONSTS


@dataclass
class Poseidon2:
    """Poseidon2<F, ..., WIDTH, D>: external constants 4+4 rounds x WIDTH, internal constants R_P scalars (Montgomery)."""
    field: Field
    width: int
    rc_initial: np.ndarray   # (4, width) Montgomery
    rc_terminal: np.ndarray  # (4, width)
    rc_internal: np.ndarray  # (rounds_p,)

    @classmethod
    def new(cls, field: Field, width: int, rc_initial, rc_terminal, rc_internal, monty: bool = True):
        conv = (lambda x: np.ascontiguousarray(x, dtype=np.uint32)) if monty else field.to_monty_array
        a, b, c = conv(rc_initial).reshape(4, width), conv(rc_terminal).reshape(4, width), conv(rc_internal).ravel()
        if width not in (16, 24):
            raise ValueError("Unsupported width (GPU backend supports 16 and 24)")
        return cls(field, width, a, b, c)

    def upload(self, gpu):
        gpu.poseidon2_set_constants(self.field.id, self.width, self.rc_initial, self.rc_terminal, self.rc_internal)

    def permute(self, gpu, states):
        self.upload(gpu)
        return gpu.poseidon2_permute(self.field.id, self.width, states)


def default_poseidon2(field: Field, width: int) -> Poseidon2:
    """default_{babybear,koalabear}_poseidon2_{16,24} (koala-bear/src/poseidon2.rs:190-198,287-295; baby-bear :180-188,276-284)."""
    k = _defaults()[f"{field.name}_{width}"]
    return Poseidon2.new(field, width, k["external_initial"], k["external_final"], k["internal"], monty=False)
