// Inverse STFT kernel body: complex-to-real FFT + dual window + overlap-add.
//
// Replaces tf.signal.inverse_stft as called at kapre/time_frequency.py:307-314
// (irfft(n_fft) -> keep the first win_length samples -> * inverse_stft_window_fn window
// (kapre/time_frequency.py:278-280) -> overlap_and_add), plus the layout transposes of
// kapre/time_frequency.py:304-305,316-317 via element strides.
//
// Gather formulation, no atomics: a tile owns `hops_out` output hops of one (batch, channel)
// signal and recomputes the R-1 halo frames that overlap into it.  Frames are processed in
// R = ceil(win/hop) classes (t mod R); frames of one class never overlap, so the
// shared-memory overlap-add is a plain read-modify-write, with a CTA barrier between classes.
//
// irfft of length N via one complex FFT of length P = N/2 (same register FFT as the forward
// kernel, run on conj(Z) to invert):  Z[k] = E[k] + i O[k],
//   E[k] = (X[k] + conj X[P-k]) / 2,  O[k] = exp(+2 pi i k / N) (X[k] - conj X[P-k]) / 2,
// z = IFFT_P(Z) gives x[2n] = Re z[n], x[2n+1] = Im z[n].  The 1/2, the 1/P and the sign of
// the odd samples are folded into the dual-window table.
#pragma once
#include "stft_core.cuh"

struct KbIstftSmem {
    int dual, twp, twn, ola, ex, total;
    int ola_len;
};

KB_HD KbIstftSmem kb_istft_smem_layout(int Q, int n_fft, int hop, int win, int TFc, int n_warps) {
    KbIstftSmem s;
    const int P = 32 * Q;
    int off = 0;
    s.dual = off; off += kb_align16((win + 2) * 4);
    s.twp = off;  off += kb_align16(Q * 33 * 8);
    s.twn = off;  off += kb_align16((P / 2) * 8);
    s.ola_len = TFc * hop + win + 2;
    s.ola = off;  off += kb_align16(s.ola_len * 4);
    s.ex = off;   off += n_warps * (32 * 33 * 8);
    s.total = off;
    (void)n_fft;
    return s;
}

KB_HD int kb_istft_tiles(int T, int hop, int win_length, int hops_out) {
    const int out_len = (T - 1) * hop + win_length;
    const int hops = (out_len + hop - 1) / hop;
    return (hops + hops_out - 1) / hops_out;
}


template <int Q>
#if defined(KB_HOST_EMU)
inline void kb_istft_cta(const KbIstftParams& p, char* smem, int cta, int n_cta)
#else
__device__ __forceinline__ void kb_istft_cta(const KbIstftParams& p, char* smem, int cta, int n_cta)
#endif
{
    constexpr int P = 32 * Q;
    constexpr int FPW = 32 / Q;
    constexpr int ZSTR = P + Q;
    constexpr int EXW = 32 * 33;
    const int NW = p.n_warps;
    const int kb_nt = NW * 32;
    (void)kb_nt;
    const int H = p.hop, TFc = p.TFc, R_ = p.R, win = p.win;
    const KbIstftSmem L = kb_istft_smem_layout(Q, p.n_fft, H, win, TFc, NW);
    float* dual_s = reinterpret_cast<float*>(smem + L.dual);
    cpx* twp_s = reinterpret_cast<cpx*>(smem + L.twp);
    cpx* twn_s = reinterpret_cast<cpx*>(smem + L.twn);
    float* ola_s = reinterpret_cast<float*>(smem + L.ola);
    cpx* ex_s = reinterpret_cast<cpx*>(smem + L.ex);
    const int FR = NW * FPW;
    const int n_tiles = p.B * p.C * p.n_tiles_t;

#if defined(KB_HOST_EMU)
    std::vector<KbThreadRegs> kb_regs(kb_nt);
#else
    KbThreadRegs kb_regs;
#endif

    KB_PHASE_BEGIN
        (void)R;
        for (int i = tid; i < win + 2; i += kb_nt) dual_s[i] = p.dual[i];
        for (int i = tid; i < Q * 33; i += kb_nt) { float2 t = p.twp[i]; twp_s[i] = cmake(t.x, t.y); }
        for (int i = tid; i < P / 2; i += kb_nt) { float2 t = p.twn[i]; twn_s[i] = cmake(t.x, t.y); }
    KB_PHASE_END
    KB_SYNC_CTA;

    for (int tile = cta; tile < n_tiles; tile += n_cta) {
        const int sig = tile / p.n_tiles_t;
        const int tt = tile - sig * p.n_tiles_t;
        const int b = sig / p.C, c = sig - b * p.C;
        const int h0 = tt * p.hops_out;          // first output hop of the tile
        const int tf0 = h0 - (R_ - 1);           // first (possibly negative) frame computed
        const float2* Xsig = p.X + (long long)b * p.x_sb + (long long)c * p.x_sc;

        KB_PHASE_BEGIN
            (void)R;
            for (int i = tid; i < L.ola_len; i += kb_nt) ola_s[i] = 0.0f;
        KB_PHASE_END
        KB_SYNC_CTA;

        for (int cl = 0; cl < R_; ++cl) {
            const int n_in_class = (TFc - cl + R_ - 1) / R_;   // frames fi = cl + R*idx < TFc
            const int n_rounds = (n_in_class + FR - 1) / FR;
            for (int round = 0; round < n_rounds; ++round) {
                // ---- phase 1: X[k], X[P-k] -> conj(2 Z[k]), conj(2 Z[P-k]) in natural order ----
                KB_PHASE_BEGIN
                    (void)R;
                    const int warp = tid >> 5, lane = tid & 31;
                    // all of the round's global loads are issued before the first use (FPW frames x Q/2 bin pairs)
                    float2 xa[FPW][Q / 2], xb[FPW][Q / 2], xm[FPW];
                    bool live[FPW];
#pragma unroll
                    for (int gg = 0; gg < FPW; ++gg) {
                        const int idx = round * FR + warp * FPW + gg;
                        const int fi = cl + R_ * idx;
                        const int t = tf0 + fi;
                        live[gg] = !(fi >= TFc || t < 0 || t >= p.T);
                        xm[gg] = make_float2(0.0f, 0.0f);
                        if (live[gg]) {
                            const float2* Xf = Xsig + (long long)t * p.x_st;
#pragma unroll
                            for (int i = 0; i < Q / 2; ++i) {
                                const int k = lane + 32 * i;
                                xa[gg][i] = Xf[(long long)k * p.x_sk];
                                xb[gg][i] = Xf[(long long)(P - k) * p.x_sk];
                            }
                            if (lane == 0) xm[gg] = Xf[(long long)(P / 2) * p.x_sk];
                        }
                    }
#pragma unroll
                    for (int gg = 0; gg < FPW; ++gg) {
                        if (!live[gg]) continue;
                        cpx* zf = ex_s + warp * EXW + gg * ZSTR;
#pragma unroll
                        for (int i = 0; i < Q / 2; ++i) {
                            const int k = lane + 32 * i;
                            float2 a = xa[gg][i];
                            float2 bq = xb[gg][i];
                            if (k == 0) { a.y = 0.0f; bq.y = 0.0f; }  // C2R ignores Im of DC / Nyquist
                            const cpx W = twn_s[k];
                            const float Er = a.x + bq.x, Ei = a.y - bq.y;
                            const float dr = a.x - bq.x, di = a.y + bq.y;
                            const float Or = W.re * dr + W.im * di;
                            const float Oi = W.re * di - W.im * dr;
                            zf[k] = cmake(Er - Oi, -(Ei + Or));
                            if (k != 0) zf[P - k] = cmake(Er + Oi, Ei - Or);
                        }
                        if (lane == 0) zf[P / 2] = cmake(2.0f * xm[gg].x, 2.0f * xm[gg].y);
                    }
                KB_PHASE_END
                KB_SYNC_WARP;
                // ---- phase 2a: strided gather of the packed sequence into registers ----------
                KB_PHASE_BEGIN
                    const int warp = tid >> 5, lane = tid & 31;
                    const int g = lane / Q, q = lane % Q;
                    const int idx = round * FR + warp * FPW + g;
                    const int fi = cl + R_ * idx;
                    const int t = tf0 + fi;
                    if (fi < TFc && t >= 0 && t < p.T) {
                        const cpx* zf = ex_s + warp * EXW + g * ZSTR;
#pragma unroll
                        for (int j = 0; j < 32; ++j) R.v[j] = zf[q + Q * j];
                    }
                KB_PHASE_END
                KB_SYNC_WARP;
                // ---- phase 2b: 32-point DFTs, twiddle, transpose-store ------------------------
                KB_PHASE_BEGIN
                    const int warp = tid >> 5, lane = tid & 31;
                    const int g = lane / Q, q = lane % Q;
                    const int idx = round * FR + warp * FPW + g;
                    const int fi = cl + R_ * idx;
                    const int t = tf0 + fi;
                    if (fi < TFc && t >= 0 && t < p.T) {
                        kb_fft_dif<32>(R.v);
                        cpx* ex = ex_s + warp * EXW + (g * Q + q) * 33;
                        const cpx* tw = twp_s + q * 33;
                        ex[0] = R.v[0];
#pragma unroll
                        for (int k1 = 1; k1 < 32; ++k1) ex[k1] = cmul(R.v[kb_brev<32>(k1)], tw[k1]);
                    }
                KB_PHASE_END
                KB_SYNC_WARP;
                // ---- phase 3: column gather, Q-point DFTs, window, overlap-add ----------------
                KB_PHASE_BEGIN
                    const int warp = tid >> 5, lane = tid & 31;
                    const int g = lane / Q, q = lane % Q;
                    const int idx = round * FR + warp * FPW + g;
                    const int fi = cl + R_ * idx;
                    const int t = tf0 + fi;
                    if (fi < TFc && t >= 0 && t < p.T) {
                        const cpx* ex = ex_s + warp * EXW + (g * Q) * 33;
#pragma unroll
                        for (int i = 0; i < FPW; ++i) {
                            const int k1 = q + Q * i;
#pragma unroll
                            for (int q2 = 0; q2 < Q; ++q2) R.v[i * Q + q2] = ex[q2 * 33 + k1];
                        }
                        float* of = ola_s + fi * H;
                        const bool pair_ok = ((fi * H) & 1) == 0;
#pragma unroll
                        for (int i = 0; i < FPW; ++i) {
                            kb_fft_dif<Q>(R.v + i * Q);
                            const int k1 = q + Q * i;
#pragma unroll
                            for (int k2 = 0; k2 < Q; ++k2) {
                                const int m2 = 2 * (k1 + 32 * k2);
                                const cpx r = R.v[i * Q + kb_brev<Q>(k2)];
                                if (pair_ok && m2 + 1 < win) {      // aligned sample pair: one 8-byte read-modify-write
                                    cpx* o2 = reinterpret_cast<cpx*>(of + m2);
                                    const cpx d2 = *reinterpret_cast<const cpx*>(dual_s + m2);
                                    *o2 = cadd(*o2, cmul_elem(r, d2));
                                } else {
                                    if (m2 < win) of[m2] += r.re * dual_s[m2];
                                    if (m2 + 1 < win) of[m2 + 1] += r.im * dual_s[m2 + 1];
                                }
                            }
                        }
                    }
                KB_PHASE_END
                KB_SYNC_WARP;
            }
            KB_SYNC_CTA;  // the next class overlaps this one's samples
        }

        // ---- copy the tile's completed hops out ------------------------------------------------
        KB_PHASE_BEGIN
            (void)R;
            const long long s_begin = (long long)h0 * H;
            long long s_end = s_begin + (long long)p.hops_out * H;
            if (s_end > p.out_len) s_end = p.out_len;
            float* ysig = p.y + (long long)b * p.y_sb + (long long)c * p.y_sc;
            const int shift = (R_ - 1) * H;
            for (long long s = s_begin + tid; s < s_end; s += kb_nt)
                ysig[s * p.y_sl] = ola_s[(int)(s - s_begin) + shift];
        KB_PHASE_END
        KB_SYNC_CTA;
    }
}
