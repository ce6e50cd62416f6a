// Register-resident radix-2 DIF FFT of compile-time size R in {2,4,8,16,32}.
// Fully unrolled: every index is a compile-time constant after unrolling, so the array
// lives in registers on the GPU.  Forward transform (kernel exp(-2 pi i jk/R)); the result
// for frequency k is left at position kb_brev<R>(k) (bit-reversed order).
#pragma once
#include "kb_common.h"

// cos(2 pi t / 32), t = 0..16
KB_HD constexpr float kb_cos32_q(int t) {
    return t == 0 ? 1.0f
         : t == 1 ? 0.98078528040323044913f
         : t == 2 ? 0.92387953251128675613f
         : t == 3 ? 0.83146961230254523708f
         : t == 4 ? 0.70710678118654752440f
         : t == 5 ? 0.55557023301960222474f
         : t == 6 ? 0.38268343236508977173f
         : t == 7 ? 0.19509032201612826785f
         : 0.0f;
}
KB_HD constexpr float kb_cos32(int t) { return t <= 8 ? kb_cos32_q(t) : -kb_cos32_q(16 - t); }
KB_HD constexpr float kb_sin32(int t) { return t <= 8 ? kb_cos32_q(8 - t) : kb_cos32_q(t - 8); }

// d * exp(-2 pi i t / 32), t in [0, 16).  `t` is a compile-time constant after unrolling,
// so the branch chain folds away and the trivial rotations cost no multiplies.
KB_HD cpx kb_twmul32(cpx d, int t) {
    if (t == 0) return d;
    if (t == 8) return cmake(d.im, -d.re);
    if (t == 4) {   // (1 - i)/sqrt2 :  ((re + im), (im - re)) * s
        const float s = 0.70710678118654752440f;
        return cscale(cadd(d, cmake(d.im, -d.re)), s);
    }
    if (t == 12) {  // (-1 - i)/sqrt2 : ((im - re), -(re + im)) * s
        const float s = 0.70710678118654752440f;
        return cscale(csub(cmake(d.im, -d.re), d), s);
    }
    return cmul_tw(d, kb_cos32(t), kb_sin32(t));
}

template <int R>
KB_HD constexpr int kb_brev(int k) {
    int r = 0;
    for (int b = 1; b < R; b <<= 1) {
        r = (r << 1) | (k & 1);
        k >>= 1;
    }
    return r;
}

template <int R, int HALF>
KB_HD void kb_fft_stage(cpx* v) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
        if ((i & HALF) == 0) {
            const int j = i & (HALF - 1);
            const int t = j * (16 / HALF);  // W_{2*HALF}^j == W_32^{j*16/HALF}
            const cpx a = v[i], b = v[i + HALF];
            v[i] = cadd(a, b);
            v[i + HALF] = kb_twmul32(csub(a, b), t);
        }
    }
    if constexpr (HALF > 1) kb_fft_stage<R, HALF / 2>(v);
}

template <int R>
KB_HD void kb_fft_dif(cpx* v) {
    static_assert(R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "unsupported radix");
    kb_fft_stage<R, R / 2>(v);
}
