// Device layout of the Poseidon2 round constants consumed by the hash kernels (set through p3gpu_poseidon2_set_constants).
#pragma once
#include "field.cuh"

namespace p3 {

struct Poseidon2Consts {          // device layout consumed by the hash kernels
    u32 rc_ext[8 * 24];  // external round r (0-3 initial, 4-7 terminal), element i at r * width + i
    u32 rc_int[32];
    int rounds_p;
    int width;
    int set;
};

}  // namespace p3
