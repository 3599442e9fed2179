// In-register complex FFT codelets of small composite size (2,3,4,6,8,12,16,24).
// Fully unrolled at compile time: every index and every twiddle is a constant
// expression, so the compiler keeps the data in registers and folds the
// twiddles into FMUL/FFMA immediates.
#pragma once
#include "spx_common.cuh"

namespace spx {

// x * exp(-/+ 2*pi*i * K/N)  (forward: minus sign), trivial angles special-cased
template <int K, int N, bool INV>
SPX_HD cf twmul(cf a) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) {
        return a;
    } else if constexpr (4 * k == N) {            // -i (fwd) / +i (inv)
        return INV ? cmake(-a.y, a.x) : cmake(a.y, -a.x);
    } else if constexpr (2 * k == N) {            // -1
        return cmake(-a.x, -a.y);
    } else if constexpr (4 * k == 3 * N) {        // +i (fwd) / -i (inv)
        return INV ? cmake(a.y, -a.x) : cmake(-a.y, a.x);
    } else if constexpr ((8 * k) % N == 0) {      // odd multiples of pi/4
        constexpr float h = 0.70710678118654752440f;
        constexpr int o = (8 * k) / N;            // 1, 3, 5, 7
        // w = (cr + i*ci) * h with cr, ci in {+1,-1}; forward angle = -2pi*k/N
        constexpr float cr = (o == 1 || o == 7) ? 1.f : -1.f;
        constexpr float sfw = (o == 1 || o == 3) ? -1.f : 1.f;   // forward sin sign
        constexpr float ci = INV ? -sfw : sfw;
        // (cr + i ci) * a * h = (cr * a + ci * (i a)) * h  -  one packed add (the +-i
        // and the signs are operand modifiers) and one packed multiply
        return cscale(cadd(cmake(cr * a.x, cr * a.y), cmake(-ci * a.y, ci * a.x)), h);
    } else {
        constexpr float wr = float(cos2pi(k, N));
        constexpr float wi = float(INV ? sin2pi(k, N) : -sin2pi(k, N));
        return cmul(a, cmake(wr, wi));
    }
}

// ---- radix butterflies (in place on R values) ---------------------------
template <bool INV> SPX_HD void bfly2(cf &a, cf &b) {
    cf t = csub(a, b); a = cadd(a, b); b = t;
}
template <bool INV> SPX_HD void bfly3(cf &a, cf &b, cf &c) {
    constexpr float s = INV ? 0.86602540378443864676f : -0.86602540378443864676f;
    cf t1 = cadd(b, c);
    cf t2 = caxpy(t1, -0.5f, a);                        // a - t1/2
    cf t3 = cscale(csub(b, c), s);                      // s*(b-c)
    a = cadd(a, t1);
    // X1 = t2 + i*t3 , X2 = t2 - i*t3     (i*(u+iv) = -v + iu)
    cf it3 = cmake(-t3.y, t3.x);
    b = cadd(t2, it3);
    c = csub(t2, it3);
}
template <bool INV> SPX_HD void bfly4(cf &a, cf &b, cf &c, cf &d) {
    cf t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
    // forward: X1 = t1 - i*t3, X3 = t1 + i*t3 ; inverse swaps
    cf it3 = INV ? cmake(-t3.y, t3.x) : cmake(t3.y, -t3.x);  // (-/+ i)*t3 -> fwd: -i*t3
    a = cadd(t0, t2);
    c = csub(t0, t2);
    b = cadd(t1, it3);
    d = csub(t1, it3);
}

constexpr int pick_radix(int n) {
    return (n % 4 == 0) ? 4 : (n % 2 == 0) ? 2 : (n % 3 == 0) ? 3 : n;
}

// Natural-order in, natural-order out, unnormalised.
template <int N, bool INV>
struct RegFFT {
    static constexpr int R = pick_radix(N);
    static constexpr int M = N / R;
    static_assert(R == 2 || R == 3 || R == 4, "size must be 2^a * 3^b");

    SPX_HD static void run(cf (&x)[N]) {
        if constexpr (N == 1) {
            return;
        } else if constexpr (M == 1) {
            if constexpr (R == 2) bfly2<INV>(x[0], x[1]);
            if constexpr (R == 3) bfly3<INV>(x[0], x[1], x[2]);
            if constexpr (R == 4) bfly4<INV>(x[0], x[1], x[2], x[3]);
        } else {
            cf y[R][M];
            static_for<0, R>([&](auto Q) {
                constexpr int q = decltype(Q)::value;
                static_for<0, M>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    y[q][j] = x[R * j + q];
                });
                RegFFT<M, INV>::run(y[q]);
            });
            static_for<0, M>([&](auto K) {
                constexpr int k = decltype(K)::value;
                cf t[R];
                static_for<0, R>([&](auto Q) {
                    constexpr int q = decltype(Q)::value;
                    t[q] = twmul<q * k, N, INV>(y[q][k]);
                });
                if constexpr (R == 2) bfly2<INV>(t[0], t[1]);
                if constexpr (R == 3) bfly3<INV>(t[0], t[1], t[2]);
                if constexpr (R == 4) bfly4<INV>(t[0], t[1], t[2], t[3]);
                static_for<0, R>([&](auto Pp) {
                    constexpr int p = decltype(Pp)::value;
                    x[k + M * p] = t[p];
                });
            });
        }
    }
};

template <bool INV> struct RegFFT<1, INV> {
    SPX_HD static void run(cf (&)[1]) {}
};

}  // namespace spx
