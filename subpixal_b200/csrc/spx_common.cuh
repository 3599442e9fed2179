// Shared helpers for the subpixal_b200 kernels (sm_100a).
//
// Everything that does per-thread work between two block barriers is written
// as an SPX_HD function of (context, tid) so that the SAME source can be
// executed sequentially on the host (tools/host_emu.cu, built with
// -DSPX_HOST_EMU) to debug index arithmetic without a GPU.  The product never
// runs that emulation: it exists only as a developer tool.
#pragma once
#ifdef SPX_HOST_EMU          // plain g++ build of tools/host_emu.cu
#define __host__
#define __device__
#define __forceinline__ inline
#define __align__(n) __attribute__((aligned(n)))
#else
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdint.h>
#include <type_traits>

#define SPX_HD __host__ __device__ __forceinline__

namespace spx {

// ---- compile-time loop --------------------------------------------------
template <int B, int E, class F>
SPX_HD void static_for(F &&f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// ---- constexpr trigonometry (exact rational angles 2*pi*k/n) -------------
constexpr double kPi = 3.14159265358979323846264338327950288;

__host__ __device__ constexpr double poly_sin(double x) {   // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 12; ++i) {
        term *= -x2 / double((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
__host__ __device__ constexpr double poly_cos(double x) {   // |x| <= pi/4
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 12; ++i) {
        term *= -x2 / double((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}
// cos(2*pi*k/n), sin(2*pi*k/n) with octant reduction on the integers, so the
// special angles are exact (0, +-1, +-sqrt(1/2) to double rounding).
__host__ __device__ constexpr double cos2pi(long long k, long long n) {
    k %= n; if (k < 0) k += n;
    if (2 * k > n) k = n - k;                       // cos is even about pi
    // now 0 <= k/n <= 1/2
    if (4 * k > n) return -cos2pi(n - 2 * k, 2 * n); // cos(x) = -cos(pi-x)
    // 0 <= k/n <= 1/4
    if (8 * k > n) {                                 // cos(x) = sin(pi/2 - x)
        long long kk = n - 4 * k;                    // angle = 2*pi*kk/(4n)
        return poly_sin(2.0 * kPi * double(kk) / double(4 * n));
    }
    return poly_cos(2.0 * kPi * double(k) / double(n));
}
__host__ __device__ constexpr double sin2pi(long long k, long long n) {
    return cos2pi(4 * k - n, 4 * n);                 // sin(x) = cos(x - pi/2)
}

// ---- complex helpers ------------------------------------------------------
struct __align__(8) cf { float x, y; };   // plain complex float (float2 layout)

SPX_HD cf cmake(float a, float b) { cf r; r.x = a; r.y = b; return r; }
// two consecutive complex values moved as one 128-bit word (slab copies)
#ifdef SPX_HOST_EMU
struct alignas(16) cf2 { cf a, b; };
#else
struct __align__(16) cf2 { cf a, b; };
#endif

// ---- packed fp32x2 arithmetic (sm_100a) ------------------------------------
// Blackwell issues two fp32 operations per lane with ONE instruction (FADD2 /
// FMUL2 / FFMA2 on 64-bit register pairs).  A complex number is exactly such a
// pair, and the operand modifiers of these instructions (half swap, per-half
// negation, scalar broadcast) absorb the multiplications by +-i, conjugations and
// the (a.x, a.x) / (a.y, a.y) operands of a complex product, so a complex add is
// one instruction instead of two and a complex multiply two instead of four.  The
// kernels are bound by instruction issue, not by the fp32 pipes, which is what
// makes this pay.  tools/host_emu.cu and the host side use the scalar forms.
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU) && !defined(SPX_NO_F32X2)
#define SPX_F32X2 1
__device__ __forceinline__ cf pk_add(cf a, cf b) {
    cf r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; "
        "add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ cf pk_sub(cf a, cf b) {
    cf r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; "
        "sub.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ cf pk_mul(cf a, cf b) {
    cf r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; "
        "mul.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ cf pk_fma(cf a, cf b, cf c) {
    cf r;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; "
        "mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
#endif

SPX_HD cf cadd(cf a, cf b) {
#ifdef SPX_F32X2
    return pk_add(a, b);
#else
    return cmake(a.x + b.x, a.y + b.y);
#endif
}
SPX_HD cf csub(cf a, cf b) {
#ifdef SPX_F32X2
    return pk_sub(a, b);
#else
    return cmake(a.x - b.x, a.y - b.y);
#endif
}
SPX_HD cf cmul(cf a, cf b) {
#ifdef SPX_F32X2
    return pk_fma(cmake(a.x, a.x), b, pk_mul(cmake(a.y, a.y), cmake(-b.y, b.x)));
#else
    return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
#endif
}
// a * (s, s) and a + b * (s, s) with a real scalar s
SPX_HD cf cscale(cf a, float s) {
#ifdef SPX_F32X2
    return pk_mul(a, cmake(s, s));
#else
    return cmake(a.x * s, a.y * s);
#endif
}
SPX_HD cf caxpy(cf b, float s, cf a) {
#ifdef SPX_F32X2
    return pk_fma(b, cmake(s, s), a);
#else
    return cmake(a.x + s * b.x, a.y + s * b.y);
#endif
}
SPX_HD cf cconj(cf a) { return cmake(a.x, -a.y); }

// Inter-stage twiddles of the two-step transform P = E1*E2, laid out so that the
// runtime factor selects a base and the compile-time factor an immediate
// offset: w[n2*E1 + k1] = W_P^(n2*k1), n2 in [0, E2), k1 in [0, E1).
// INV: the conjugate table (inverse transforms) - a plain complex product with
// the stored value is two packed instructions, a product with the conjugate is not.
template <int P, int E1, bool INV = false>
struct TwTable2 {
    cf w[P];
    constexpr TwTable2() : w() {
        for (int n2 = 0; n2 < P / E1; ++n2)
            for (int k1 = 0; k1 < E1; ++k1) {
                w[n2 * E1 + k1].x = float(cos2pi((long long)n2 * k1, P));
                w[n2 * E1 + k1].y = float(INV ? sin2pi((long long)n2 * k1, P)
                                              : -sin2pi((long long)n2 * k1, P));
            }
    }
};

SPX_HD bool finitef(float v) {
    // works under any fast-math setting: exponent bits all ones <=> inf/nan
    union { float f; uint32_t u; } c; c.f = v;
    return (c.u & 0x7f800000u) != 0x7f800000u;
}

}  // namespace spx
