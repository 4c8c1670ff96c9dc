// packed_f32x2.cuh — the clause arithmetic of the fast path for TWO f32 replicas at once (sm_100a packed f32x2
// instructions), shared by the tile kernels, the slab kernel and the streaming clause kernel of the gather engine.
#pragma once
#include "common.cuh"

namespace odesat {

// ---- packed f32x2 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2) --------------------------------
// The two f32 replicas of a tile sit in adjacent registers (rows are {v0, v1, dv0, dv1}), so every
// add / mul / fma of the clause arithmetic is issued once for both.  Each lane is an IEEE
// round-to-nearest operation without flush-to-zero, i.e. bit-identical to the scalar instruction.
__device__ __forceinline__ unsigned long long pk2(float2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 up2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
    return up2(r);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
    return up2(r);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
    return up2(r);
}
__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }

// Fast path of clause_math for the two f32 replicas of a tile at once (same operations in the same
// order as clause_math<float, false>, the add / mul / fma ones packed).
__device__ __forceinline__ void clause_math_f32x2(const float2 (&v)[3], float2 (&d)[3], const float (&q)[3], float2& xs, float2& xl,
                                                  float (&mx)[2], float2 dt, float xl_max) {
    const float hi_s = 1.0f - Kc<float>::EPSILON;
    float2 a[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) a[j] = fma2(bc2(-q[j]), v[j], bc2(1.0f));
    float2 mn, sm;
    {
        const float lo = rmin(a[0].x, a[1].x), hi = rmax(a[0].x, a[1].x);
        mn.x = rmin(lo, a[2].x);
        sm.x = rmax(lo, rmin(hi, a[2].x));
    }
    {
        const float lo = rmin(a[0].y, a[1].y), hi = rmax(a[0].y, a[1].y);
        mn.y = rmin(lo, a[2].y);
        sm.y = rmax(lo, rmin(hi, a[2].y));
    }
    const float2 cm = mul2(bc2(0.5f), mn);                                  // :60
    const float2 h = mul2(bc2(0.5f), mul2(xl, xs));
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float2 sel = make_float2((a[j].x != mn.x) ? mn.x : sm.x, (a[j].y != mn.y) ? mn.y : sm.y);
        d[j] = fma2(mul2(h, sel), bc2(q[j]), d[j]);                         // :64-70, :80
    }
    const float2 dxs = mul2(mul2(bc2(Kc<float>::BETA), add2(xs, bc2(Kc<float>::EPSILON))), add2(cm, bc2(-Kc<float>::GAMMA)));   // :84
    const float2 dxl = mul2(bc2(Kc<float>::ALPHA), add2(cm, bc2(-Kc<float>::DELTA)));                                           // :85
    // :88 as a running maximum: C_m = 0.5·min exactly (the minimum is 0 or a multiple of 2^-24 ≥ 2^-24), so
    // "some C_m ≥ 0.25" ⇔ "max over clauses of min ≥ 0.5"; no NaN can occur on this path
    mx[0] = rmax(mx[0], mn.x);
    mx[1] = rmax(mx[1], mn.y);
    // y + dt·dy: the products are packed, the additions stay scalar — ptxas (12.9) contracts
    // mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false, which would round once instead
    // of twice.  (dt = 0 freezes a replica.)
    const float2 pxs = mul2(dt, dxs), pxl = mul2(dt, dxl);
    xs = make_float2(rmin(rmax(__fadd_rn(xs.x, pxs.x), Kc<float>::EPSILON), hi_s), rmin(rmax(__fadd_rn(xs.y, pxs.y), Kc<float>::EPSILON), hi_s));   // :94
    xl = make_float2(rmin(rmax(__fadd_rn(xl.x, pxl.x), 1.0f), xl_max), rmin(rmax(__fadd_rn(xl.y, pxl.y), 1.0f), xl_max));                          // :95
}

}  // namespace odesat
