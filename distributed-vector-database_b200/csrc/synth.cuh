// synth.cuh -- counter-based synthetic unit-norm rows; same definition as
// oracle/cpu_ref.py synth_rows (integer arithmetic + correctly rounded fp64 sqrt/divide,
// so host and device agree bit for bit).
#pragma once
#include <stdint.h>

namespace vdbk {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t synth_row_base(uint64_t seed, uint64_t row) {
    return splitmix64(seed * 0x9E3779B97F4A7C15ull + row);
}
// Irwin-Hall(4) over the four 16-bit fields of the hash, centred: integer in [-131070, 131070]
__host__ __device__ __forceinline__ long long synth_int(uint64_t base, int col) {
    const uint64_t h = splitmix64(base + (uint64_t)col);
    return (long long)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48)) - 131070;
}

}  // namespace vdbk
