// Fused forward kernel body: frame + window + real FFT (+ |.| + filterbank + dB).
//
// Replaces, in one pass over the waveform, the reference op chain
//   kapre/time_frequency.py:164-187  (transpose, pad_begin, tf.signal.stft, transpose)
//   kapre/time_frequency.py:351-359  (tf.abs)
//   kapre/time_frequency.py:535-548  (tf.tensordot with the filterbank, transpose)
//   kapre/backend.py:186-188         (10*log10(max(x, amin)) - 10*log10(max(amin, ref)))
// The per-item clamp of kapre/backend.py:190-192 needs the item-wide maximum, so this
// kernel only reduces that maximum (atomicMax) and a second tiny kernel applies the clamp.
//
// Real FFT of length N as one complex FFT of length P = N/2 = 32*Q on the packed signal
// z[n] = x[2n] + i x[2n+1]; Q lanes of a warp cooperate on one frame (32/Q frames per warp):
//   pass 1  lane q: 32-point DFT over j of z[q + Q j]        (registers)
//           twiddle exp(-2 pi i q k1 / P), transpose through the warp's shared buffer
//   pass 2  lane q: Q-point DFTs over the columns k1 = q + Q i (registers)
//           Z[k1 + 32 k2] written back in natural order
//   pair    X[k], X[P-k] from Z[k], Z[P-k] and exp(-2 pi i k / N)  -> output epilogue
#pragma once
#include "fft_regs.cuh"

struct KbStftSmem {
    int wh, twp, twn, samples, ex, mag, total;  // byte offsets
    int span;   // samples staged per tile
    int TFp;    // padded column stride of mag_s
    int Mp;     // padded band stride of out_s
};

// Shared-memory carve-up; used by the host launcher (size) and by the kernel (offsets).
KB_HD KbStftSmem kb_stft_smem_layout(int Q, int n_fft, int hop, int TF, int n_warps, int mode, int n_bands) {
    KbStftSmem s;
    const int P = 32 * Q;
    int off = 0;
    s.wh = off;  off += kb_align16(n_fft * 4);
    s.twp = off; off += kb_align16(Q * 33 * 8);
    s.twn = off; off += kb_align16((P / 2) * 8);
    s.span = (TF - 1) * hop + n_fft;
    s.TFp = TF | 1;
    s.Mp = n_bands | 1;
    int samp_floats = s.span + 2;
    const bool fb = (mode == KB_OUT_FB || mode == KB_OUT_FB_DB);
    if (fb && TF * s.Mp > samp_floats) samp_floats = TF * s.Mp;  // out_s aliases the sample buffer
    s.samples = off; off += kb_align16(samp_floats * 4);
    s.ex = off;  off += n_warps * (32 * 33 * 8);
    s.mag = off; if (fb) off += kb_align16((P + 1) * s.TFp * 4);
    s.total = off;
    return s;
}

struct KbThreadRegs {
    cpx v[32];
    float runmax;
};

#if defined(KB_HOST_EMU)
#include <vector>
#include <algorithm>
#define KB_PHASE_BEGIN for (int tid = 0; tid < kb_nt; ++tid) { KbThreadRegs& R = kb_regs[tid];
#define KB_PHASE_END }
#define KB_SYNC_WARP
#define KB_SYNC_CTA
static inline float kb_sqrt(float v) { return std::sqrt(v); }
static inline float kb_log2(float v) { return std::log2(v); }
static inline float kb_ldg(const float* p) { return *p; }
static inline void kb_atomic_max_u32(unsigned int* p, unsigned int v) { if (v > *p) *p = v; }
static inline unsigned int kb_f2u(float f) { unsigned int u; std::memcpy(&u, &f, 4); return u; }
#else
#define KB_PHASE_BEGIN { const int tid = threadIdx.x; KbThreadRegs& R = kb_regs;
#define KB_PHASE_END }
#define KB_SYNC_WARP __syncwarp()
#define KB_SYNC_CTA __syncthreads()
KB_D float kb_sqrt(float v) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
KB_D float kb_log2(float v) { return __log2f(v); }
KB_D float kb_ldg(const float* p) { return __ldg(p); }
KB_D void kb_atomic_max_u32(unsigned int* p, unsigned int v) { atomicMax(p, v); }
KB_D unsigned int kb_f2u(float f) { return __float_as_uint(f); }
#endif

// One CTA's share of the work: tiles cta, cta + n_cta, ... ; a tile is TF consecutive frames
// of one (batch, channel) signal.
template <int Q>
#if defined(KB_HOST_EMU)
inline void kb_stft_cta(const KbStftParams& p, char* smem, int cta, int n_cta)
#else
__device__ __forceinline__ void kb_stft_cta(const KbStftParams& p, char* smem, int cta, int n_cta)
#endif
{
    constexpr int P = 32 * Q;        // complex FFT length
    constexpr int FPW = 32 / Q;      // frames per warp per round
    constexpr int ZSTR = P + Q;      // frame stride (complex) of the natural-order buffer
    constexpr int EXW = 32 * 33;     // complex elements per warp in the exchange buffer
    const int NW = p.n_warps;
    const int kb_nt = NW * 32;
    (void)kb_nt;
    const int H = p.hop, N = p.n_fft, TF = p.TF;
    const KbStftSmem L = kb_stft_smem_layout(Q, N, H, TF, NW, p.mode, p.n_bands);
    float* wh_s = reinterpret_cast<float*>(smem + L.wh);
    cpx* twp_s = reinterpret_cast<cpx*>(smem + L.twp);
    cpx* twn_s = reinterpret_cast<cpx*>(smem + L.twn);
    float* smp_s = reinterpret_cast<float*>(smem + L.samples);
    cpx* ex_s = reinterpret_cast<cpx*>(smem + L.ex);
    float* mag_s = reinterpret_cast<float*>(smem + L.mag);
    float* out_s = smp_s;  // aliases the sample buffer (dead once the FFT rounds are done)
    const bool fbmode = (p.mode == KB_OUT_FB || p.mode == KB_OUT_FB_DB);
    const bool dbmode = (p.mode == KB_OUT_MAG_DB || p.mode == KB_OUT_FB_DB);
    const int FR = NW * FPW;                       // frames per round
    const int n_rounds = (TF + FR - 1) / FR;
    const int n_tiles = p.B * p.C * p.n_tiles_t;

#if defined(KB_HOST_EMU)
    std::vector<KbThreadRegs> kb_regs(kb_nt);
#else
    KbThreadRegs kb_regs;
#endif

    // ---- one-time: tables to shared memory -------------------------------------------------
    KB_PHASE_BEGIN
        (void)R;
        for (int i = tid; i < N; i += kb_nt) wh_s[i] = p.wh[i];
        for (int i = tid; i < Q * 33; i += kb_nt) { float2 t = p.twp[i]; twp_s[i] = cmake(t.x, t.y); }
        for (int i = tid; i < P / 2; i += kb_nt) { float2 t = p.twn[i]; twn_s[i] = cmake(t.x, t.y); }
    KB_PHASE_END
    KB_SYNC_CTA;

    for (int tile = cta; tile < n_tiles; tile += n_cta) {
        const int sig = tile / p.n_tiles_t;
        const int tt = tile - sig * p.n_tiles_t;
        const int b = sig / p.C, c = sig - b * p.C;
        const int t0 = tt * TF;
        const float* xsig = p.x + (long long)b * p.x_sb + (long long)c * p.x_sc;
        const long long obase = (long long)b * p.o_sb + (long long)c * p.o_sc;

        // ---- phase 0: stage the tile's samples (zero outside [0, L): pad_begin / pad_end) ----
        KB_PHASE_BEGIN
            R.runmax = 0.0f;
            const long long s_first = (long long)t0 * H - p.pad_left;
            for (int i = tid; i < L.span; i += kb_nt) {
                const long long s = s_first + i;
                float val = 0.0f;
                if (s >= 0 && s < p.L) val = xsig[s * p.x_sl];
                smp_s[i] = val;
            }
        KB_PHASE_END
        KB_SYNC_CTA;

        for (int round = 0; round < n_rounds; ++round) {
            // ---- phase 1: window, 32-point DFTs, twiddle, transpose-store ------------------
            KB_PHASE_BEGIN
                const int warp = tid >> 5, lane = tid & 31;
                const int g = lane / Q, q = lane % Q;
                const int col = round * FR + warp * FPW + g;
                if (col < TF) {
                    const float* fr = smp_s + col * H;
                    if ((H & 1) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n2 = 2 * (q + Q * j);
                            const float2 xv = *reinterpret_cast<const float2*>(fr + n2);
                            const float2 wv = *reinterpret_cast<const float2*>(wh_s + n2);
                            R.v[j] = cmake(xv.x * wv.x, xv.y * wv.y);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n2 = 2 * (q + Q * j);
                            R.v[j] = cmake(fr[n2] * wh_s[n2], fr[n2 + 1] * wh_s[n2 + 1]);
                        }
                    }
                    kb_fft_dif<32>(R.v);
                    cpx* ex = ex_s + warp * EXW + (g * Q + q) * 33;
                    const cpx* tw = twp_s + q * 33;
                    ex[0] = R.v[0];
#pragma unroll
                    for (int k1 = 1; k1 < 32; ++k1) ex[k1] = cmul(R.v[kb_brev<32>(k1)], tw[k1]);
                }
            KB_PHASE_END
            KB_SYNC_WARP;
            // ---- phase 2: gather this lane's columns --------------------------------------
            KB_PHASE_BEGIN
                const int warp = tid >> 5, lane = tid & 31;
                const int g = lane / Q, q = lane % Q;
                const int col = round * FR + warp * FPW + g;
                if (col < TF) {
                    const cpx* ex = ex_s + warp * EXW + (g * Q) * 33;
#pragma unroll
                    for (int i = 0; i < FPW; ++i) {
                        const int k1 = q + Q * i;
#pragma unroll
                        for (int q2 = 0; q2 < Q; ++q2) R.v[i * Q + q2] = ex[q2 * 33 + k1];
                    }
                }
            KB_PHASE_END
            KB_SYNC_WARP;
            // ---- phase 3: Q-point DFTs, natural-order store (aliases the exchange buffer) ---
            KB_PHASE_BEGIN
                const int warp = tid >> 5, lane = tid & 31;
                const int g = lane / Q, q = lane % Q;
                const int col = round * FR + warp * FPW + g;
                if (col < TF) {
                    cpx* zs = ex_s + warp * EXW + g * ZSTR;
#pragma unroll
                    for (int i = 0; i < FPW; ++i) {
                        kb_fft_dif<Q>(R.v + i * Q);
                        const int k1 = q + Q * i;
#pragma unroll
                        for (int k2 = 0; k2 < Q; ++k2) zs[k1 + 32 * k2] = R.v[i * Q + kb_brev<Q>(k2)];
                    }
                }
            KB_PHASE_END
            KB_SYNC_WARP;
            // ---- phase 4: real-FFT pair post-processing + epilogue -------------------------
            KB_PHASE_BEGIN
                const int warp = tid >> 5, lane = tid & 31;
#pragma unroll 1
                for (int gg = 0; gg < FPW; ++gg) {
                    const int col = round * FR + warp * FPW + gg;
                    const int t = t0 + col;
                    if (col >= TF) break;
                    const bool valid = t < p.T;
                    const cpx* zf = ex_s + warp * EXW + gg * ZSTR;
                    const long long ofr = obase + (long long)t * p.o_st;
#pragma unroll
                    for (int i = 0; i <= Q / 2; ++i) {
                        cpx X1, X2;
                        int k, kk;
                        if (i < Q / 2) {
                            k = lane + 32 * i;
                            kk = P - k;
                            const cpx A = zf[k];
                            const cpx Bv = zf[kk & (P - 1)];
                            const cpx W = twn_s[k];
                            const float Er = A.re + Bv.re, Ei = A.im - Bv.im;
                            const float Dr = A.re - Bv.re, Di = A.im + Bv.im;
                            const float Tr = W.re * Di + W.im * Dr;
                            const float Ti = W.im * Di - W.re * Dr;
                            X1 = cmake(Er + Tr, Ei + Ti);
                            X2 = cmake(Er - Tr, Ti - Ei);
                        } else {  // bin P/2 pairs with itself; lane 0 only
                            if (lane != 0) continue;
                            k = P / 2;
                            kk = -1;
                            const cpx A = zf[P / 2];
                            X1 = cmake(2.0f * A.re, -2.0f * A.im);
                            X2 = X1;
                        }
                        if (p.mode == KB_OUT_COMPLEX) {
                            if (valid) {
                                float2* o = reinterpret_cast<float2*>(p.out);
                                o[ofr + (long long)k * p.o_sk] = make_float2(X1.re, X1.im);
                                if (kk >= 0) o[ofr + (long long)kk * p.o_sk] = make_float2(X2.re, X2.im);
                            }
                        } else {
                            const float m1 = kb_sqrt(X1.re * X1.re + X1.im * X1.im);
                            const float m2 = kb_sqrt(X2.re * X2.re + X2.im * X2.im);
                            if (fbmode) {
                                mag_s[k * L.TFp + col] = m1;
                                if (kk >= 0) mag_s[kk * L.TFp + col] = m2;
                            } else if (valid) {
                                float* o = reinterpret_cast<float*>(p.out);
                                if (dbmode) {
                                    const float a1 = fmaxf(m1, p.amin), a2 = fmaxf(m2, p.amin);
                                    R.runmax = fmaxf(R.runmax, a1);
                                    o[ofr + (long long)k * p.o_sk] = p.db_mul * kb_log2(a1) - p.db_sub;
                                    if (kk >= 0) {
                                        R.runmax = fmaxf(R.runmax, a2);
                                        o[ofr + (long long)kk * p.o_sk] = p.db_mul * kb_log2(a2) - p.db_sub;
                                    }
                                } else {
                                    o[ofr + (long long)k * p.o_sk] = m1;
                                    if (kk >= 0) o[ofr + (long long)kk * p.o_sk] = m2;
                                }
                            }
                        }
                    }
                }
            KB_PHASE_END
            KB_SYNC_WARP;
        }  // rounds

        if (fbmode) {
            KB_SYNC_CTA;
            // ---- phase 5: filterbank, one lane per frame column (uniform weights) ----------
            KB_PHASE_BEGIN
                const int warp = tid >> 5, lane = tid & 31;
                const int cpw = TF < 32 ? TF : 32;     // columns per warp
                const int subs = 32 / cpw;             // band sub-groups per warp
                const int colm = lane % cpw, sub = lane / cpw;
                const bool valid = (t0 + colm) < p.T;
                const float* mcol = mag_s + colm;
                for (int m = warp * subs + sub; m < p.n_bands; m += NW * subs) {
                    const KbBand bd = p.bands[m];
                    const float* w = p.fbw + bd.off - bd.lo;
                    float a0 = 0.0f, a1 = 0.0f;
                    int k = bd.lo;
                    for (; k + 1 < bd.hi; k += 2) {
                        a0 += kb_ldg(w + k) * mcol[k * L.TFp];
                        a1 += kb_ldg(w + k + 1) * mcol[(k + 1) * L.TFp];
                    }
                    if (k < bd.hi) a0 += kb_ldg(w + k) * mcol[k * L.TFp];
                    float acc = a0 + a1;
                    if (dbmode) {
                        acc = fmaxf(acc, p.amin);
                        if (valid) R.runmax = fmaxf(R.runmax, acc);
                        acc = p.db_mul * kb_log2(acc) - p.db_sub;
                    }
                    out_s[colm * L.Mp + m] = acc;
                }
            KB_PHASE_END
            KB_SYNC_CTA;
            // ---- phase 6: coalesced copy-out of the (TF x n_bands) block --------------------
            KB_PHASE_BEGIN
                (void)R;
                float* o = reinterpret_cast<float*>(p.out);
                const int M = p.n_bands;
                const int tot = TF * M;
                for (int idx = tid; idx < tot; idx += kb_nt) {
                    const int colm = idx / M, m = idx - colm * M;
                    const int t = t0 + colm;
                    if (t < p.T) o[obase + (long long)t * p.o_st + (long long)m * p.o_sk] = out_s[colm * L.Mp + m];
                }
            KB_PHASE_END
        }

        // ---- per-item maximum for the decibel clamp (kapre/backend.py:190-192) --------------
        if (dbmode) {
#if defined(KB_HOST_EMU)
            for (int tid = 0; tid < kb_nt; ++tid)
                kb_atomic_max_u32(p.item_max + b, kb_f2u(kb_regs[tid].runmax));
#else
            const unsigned int wm = __reduce_max_sync(0xffffffffu, kb_f2u(kb_regs.runmax));
            if ((threadIdx.x & 31) == 0 && wm != 0u) kb_atomic_max_u32(p.item_max + b, wm);
#endif
        }
        KB_SYNC_CTA;  // sample / mag buffers are reused by the next tile
    }
}
