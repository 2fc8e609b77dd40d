// at_noise.cuh — counter-based noise shared by the kernels of this package.
// SplitMix64 output number ((step << 40) | agent) of the stream seeded `seed`
// (numpy restatement: oracle/noise.py, torch restatement: agenttorch_b200/noise.py).
#pragma once
#include <stdint.h>

__host__ __device__ __forceinline__ uint64_t at_noise_u64(uint64_t seed, uint32_t step, uint64_t agent) {
    uint64_t z = (((uint64_t)step << 40) | agent) + 1ull;
    z = z * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
