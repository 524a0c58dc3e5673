"""DuplexChallenger (challenger/src/duplex_challenger.rs:60-300) + GrindingChallenger::grind (grinding_challenger.rs:100-232) with the
sponge resident on the GPU (csrc/challenger.cu): caps and opened values produced on the device are absorbed there; only sampled
challenges come back.  Protocol plumbing of the prove driver (uni_stark.py), mirroring the reference's method names."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check
from .field import Field
from .gpu import _is_torch
from .poseidon2 import Poseidon2


class DuplexChallenger:
    """DuplexChallenger<F, Poseidon2<width>, width, rate>: examples use (Perm24, 24, 16) (examples/src/types.rs:56-60)."""

    def __init__(self, field: Field, perm: Poseidon2, rate: int, gpu):
        assert perm.field is field or perm.field == field
        self.field, self.perm, self.rate, self.gpu = field, perm, rate, gpu
        perm.upload(gpu)
        h = C.c_void_p()
        gpu._use_torch_stream()
        check(gpu.L.p3gpu_challenger_new(gpu.h, field.id, perm.width, rate, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.gpu.h:
                self.gpu.L.p3gpu_challenger_free(self.gpu.h, self.h)
                self.h = None
        except Exception:
            pass

    def clone(self):
        c = object.__new__(DuplexChallenger)
        c.field, c.perm, c.rate, c.gpu = self.field, self.perm, self.rate, self.gpu
        h = C.c_void_p()
        self.gpu._use_torch_stream()
        check(self.gpu.L.p3gpu_challenger_clone(self.gpu.h, self.h, C.byref(h)))
        c.h = h
        return c

    # ---- CanObserve
    def observe_slice(self, values):
        """Montgomery words; a CUDA int32 tensor is absorbed on the device without a copy."""
        self.gpu._use_torch_stream()
        if _is_torch(values) and values.is_cuda:
            v = values.contiguous()
            check(self.gpu.L.p3gpu_challenger_observe_dev(self.gpu.h, self.h, v.data_ptr(), v.numel()))
            self._keep = v
            return
        v = np.ascontiguousarray(values.cpu().numpy().view(np.uint32) if _is_torch(values) else values, dtype=np.uint32).ravel()
        check(self.gpu.L.p3gpu_challenger_observe(self.gpu.h, self.h, v.ctypes.data, v.size))

    def observe(self, value: int): self.observe_slice(np.array([value], dtype=np.uint32))
    def observe_canonical(self, x: int): self.observe(self.field.to_monty(x))           # Val::from_u8 / from_usize
    def observe_cap(self, cap): self.observe_slice(cap)                                # every digest, element by element
    def observe_algebra_slice(self, ys): self.observe_slice(ys)                        # EF4 = 4 base coefficients in order

    # ---- CanSample
    def sample_many(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint32)
        self.gpu._use_torch_stream()
        check(self.gpu.L.p3gpu_challenger_sample(self.gpu.h, self.h, out.ctypes.data, n))
        return out

    def sample(self) -> int: return int(self.sample_many(1)[0])
    def sample_algebra_element(self) -> np.ndarray: return self.sample_many(4)

    def sample_bits(self, bits: int) -> int:
        """CanSampleBits (duplex_challenger.rs:270-283): canonical value of one sample, masked."""
        assert (1 << bits) < self.field.P
        return self.field.from_monty(self.sample()) & ((1 << bits) - 1)

    # ---- GrindingChallenger
    def grind(self, bits: int) -> int:
        w = C.c_uint32()
        self.gpu._use_torch_stream()
        check(self.gpu.L.p3gpu_challenger_grind(self.gpu.h, self.h, bits, C.byref(w)))
        return int(w.value)
