// common.cuh — shared device/host definitions of the DMM integrator kernels (sm_100a).
//
// Compiled with -fmad=false: the reference (Rust, src/system.rs) never contracts a*b+c into
// an FMA, and bit-for-bit agreement with its arithmetic is the parity bar.  Where an FMA is
// provably exact (q = ±1) it is written explicitly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace odesat {

// system.rs:19-23
template <typename T> struct Kc {
    static constexpr T ALPHA = T(5.0);
    static constexpr T BETA = T(20.0);
    static constexpr T GAMMA = T(0.25);
    static constexpr T DELTA = T(0.05);
    static constexpr T EPSILON = T(0.001);
};

template <typename T> __host__ __device__ inline T inf_v();
template <> __host__ __device__ inline float inf_v<float>() { return __builtin_huge_valf(); }
template <> __host__ __device__ inline double inf_v<double>() { return __builtin_huge_val(); }

// Rust f64::max/min ignore a NaN operand; CUDA fmax/fmin have the same contract.
__device__ __forceinline__ float rmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double rmax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float rmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double rmin(double a, double b) { return fmin(a, b); }

// system.rs:94-96: (y + dt*dy).max(lo).min(hi)
template <typename T> __device__ __forceinline__ T euler_clamp(T y, T dy, T dt, T lo, T hi) {
    return rmin(rmax(y + dt * dy, lo), hi);
}

// Memories on which the fast arithmetic is bit-identical to the literal statements: finite, and zero or of a
// magnitude that keeps 0.5·xl·xs and its products with the clause minima normal (no underflow, no overflow).
// Every memory the integrator itself produces qualifies (xs ∈ [ε, 1−ε] or ±1, xl ∈ [1, 1e4·M]); a caller-uploaded
// inf / NaN / denormal sends the first step through the STRICT kernel.
template <typename T> __device__ __forceinline__ bool mem_in_fast_domain(T x) {
    const T a = fabs(x);
    return a == T(0) || (a >= T(1e-12) && a <= T(1e12));   // false for NaN and inf
}

// Device view of a formula: clause CSR + variable→clause transpose.
struct FormulaDev {
    int64_t N = 0, M = 0, L = 0;
    int K = 0;                         // uniform clause length, 0 = ragged
    const int32_t* coff = nullptr;     // [M+1] literal offsets
    const int32_t* lits = nullptr;     // [L]   ±(var+1)
    const int32_t* voff = nullptr;     // [N+1] occurrence offsets
    const int32_t* occ_clause = nullptr;   // [L] clause of each occurrence, sorted (clause, pos)
    const int32_t* occ_slot = nullptr;     // [L] literal slot (index into lits) of the occurrence
    const int8_t* xs0 = nullptr;       // [M]  +1 if the clause has a negated literal else −1
};

// SplitMix64 counter-based generator for v0 (stand-in for main.rs:170-174's OS-seeded
// thread_rng; the oracle implements the same function).
__host__ __device__ inline uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline uint64_t v0_key(uint64_t seed, uint64_t replica) {
    return sm64(seed ^ sm64(replica));
}
__host__ __device__ inline uint64_t v0_bits(uint64_t key, uint64_t var) {
    return sm64(key ^ (var * 0xD1342543DE82EF95ull));
}
template <typename T> __host__ __device__ inline T v0_value(uint64_t bits);
template <> __host__ __device__ inline double v0_value<double>(uint64_t bits) {
    return double(bits >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}
template <> __host__ __device__ inline float v0_value<float>(uint64_t bits) {
    return float(bits >> 40) * (1.0f / 16777216.0f) * 2.0f - 1.0f;
}

// Non-negative floats order like their bit patterns: max-reduce |a-b| with integer atomics.
// Stored value = bits + 1; 0 marks "no non-NaN element seen" (max_error's folds start at NaN,
// system.rs:103, and f64::max ignores NaN operands).
template <typename T> struct ErrBits;
template <> struct ErrBits<float> {
    using U = unsigned int;
    static constexpr U NONE = 0u;
    __device__ static U enc(float x) { return __float_as_uint(x) + 1u; }
    __device__ static float dec(U u) { return u == NONE ? __uint_as_float(0x7FC00000u) : __uint_as_float(u - 1u); }
};
template <> struct ErrBits<double> {
    using U = unsigned long long;
    static constexpr U NONE = 0ull;
    __device__ static U enc(double x) { return (U)__double_as_longlong(x) + 1ull; }
    __device__ static double dec(U u) {
        return u == NONE ? __longlong_as_double(0x7FF8000000000000ll) : __longlong_as_double((long long)(u - 1ull));
    }
};

}  // namespace odesat
