// Fused log-mel kernel with the real FFT on the 5th-generation tensor cores (tcgen05 + TMEM), n_fft = 1024.
//
// Replaces the same reference chain as stft_core.cuh (kapre/time_frequency.py:164-187 STFT, :351-359 |.|,
// :535-548 filterbank, kapre/backend.py:186-188 decibel) for the headline configuration: n_fft = win_length = 1024,
// a cosine-sum window (tf.signal.hann_window -- kapre's default -- or hamming_window), hop 128 or 256.
//
// 1024 = 32 x 32 Cooley-Tukey, n = 32 n1 + n2, k = k1 + 32 k2, both stages as fp32-grade GEMMs (3xTF32 split,
// fp32 accumulators in TMEM):
//   stage 1 (tc_dft.cuh)  S[(f, n2), k1] = sum_n1 x[hop f + 32 n1 + n2] e^{-2 pi i n1 k1 / 32}: A = the RAW hop-overlapped
//            sample buffer read in place as an MN-major SWIZZLE_128B_BASE32B operand, M = 4 frames x 32, K = 32, N = 32
//            (k1 = 0..15 as (Re, Im); the zero Im(k1 = 0) column carries the real k1 = 16).
//   between  one thread per row (f, n2):  U[k1] = a w^k1 S[k1]  (w = e^{-2 pi i n2 / 1024}, table);  the WINDOW
//            a - b cos(2 pi n / N) is a 3-tap filter along k1 of the twiddled values: T[k1] = U[k1] - b/(2a) (U[k1-1] + U[k1+1])
//            (U[-1] = conj U[1], U[17] = a w^17 conj S[15]) -- so the samples are never windowed in the time domain and
//            every sample is split into TF32 hi/lo once, not once per overlapping frame.  T (k1 = 1..16) is split
//            and stored as the MN-major A operand of stage 2 (the transpose between the stages is free: a row of
//            stage 1 is a K row of stage 2).  k1 = 0 is a REAL sequence: its 32-point DFT over n2 (bins 0, 32, .., 512)
//            is done with warp shuffles.
//   stage 2  X[k1 + 32 k2] = sum_n2 T[n2, k1] e^{-2 pi i n2 k2 / 32}: rows (f, k1) M = 8 frames x 16, K = 32, as four
//            real products D_re = A_re C + A_im S, D_im = A_im C - A_re S (N = 32 each; the minus through the
//            instruction descriptor's negate-A bit).  Thread (f, k1) then holds bins k1 + 32 k2 (k2 < 16) and, through
//            conjugate symmetry, 1024 - (k1 + 32 k2) (k2 >= 16): magnitudes straight from registers.
//   mel + dB  banded filterbank from the chunk lists of kb_make_fb_chunks on the CUDA cores, decibel and the per-item
//            maximum as in stft_core.cuh.
//
// One CTA of 17 warps per SM, three roles: warps 0-7 (X) stage samples (cp.async into a raw buffer, then the hi/lo
// split) and run the between-stages step; warps 8-15 (Y) turn stage-2 accumulators into magnitudes, mel bands,
// decibels and stores; one lane of warp 16 (M) issues every tcgen05.mma.  The roles meet only through mbarriers
// (tcgen05.commit for MMA completion, 256-arrival barriers for "operand written" / "accumulator read"), so the
// tensor pipe, the X warps and the Y warps work on different 8-frame units at the same time and no X thread ever
// waits for another X thread.  TMEM: 2 x 64 columns stage 1, 2 x 64 stage 2.
#pragma once
#include "tc_dft.cuh"
#include "stft_core.cuh"

#define TCM_UNIT 8                   // frames per unit (one stage-2 GEMM tile: 8 frames x 16 k1 = 128 rows)
#define TCM_TF 16                    // frames per tile (two units share one sample staging)
#define TCM_XT 256                   // threads per role (X, Y); warp 16 issues the MMAs
#define TCM_THREADS 544
#define TCM_MAXROWS 152              // 128-byte sample rows per tile at hop 256: 15 * 8 + 32
#define TCM_MS 516                   // floats per frame in the magnitude buffer (513 bins + 3 pad)
#define TCM_MP 129                   // floats per frame in the band buffer

struct KbTcMelParams {
    const float* x;                  // waveform, strides (batch, channel), sample stride 1
    long long x_sb, x_sc;
    int B, C, L, T, hop, pad_left;
    float wc;                        // b / (2 a) of the window a - b cos(2 pi n / 1024)
    const float* f1;                 // stage-1 matrix, (hi|lo, 8 kchunk, 32 col, 4)
    const float* cs;                 // stage-2 matrices C_hi, C_lo, S_hi, S_lo, each (8 kchunk, 32 col, 4)
    const unsigned short* csb;       // C, S as bf16 (4 kchunk, 32 col, 8) for the lo-part products (kind::f16)
    const float2* tw;                // [18][32]: a e^{-2 pi i n2 k1 / 1024}, k1 = 0..17
    const float2* w32;               // [16]: e^{-2 pi i j / 32}
    const kb_f4* cw;                 // filterbank chunk lists (kb_make_fb_chunks, 32 groups)
    const kb_i2* cm;
    const int* cg;
    int n_chunks, n_bands;
    float* out;                      // (batch, channel, frame, band) element strides
    long long o_sb, o_sc, o_st, o_sk;
    int db;                          // 1: decibel output + per-item maximum
    float amin, db_mul, db_sub;
    unsigned int* item_max;
    float2* dbg;                     // optional: complex spectrum (B*C, T, 513), debugging only
    int n_tiles_t;                   // ceil(T / TCM_TF)
    int ablate;                      // timing experiments only: 1 Y skips its work, 2 X skips the between-stages math, 4 no MMAs, 8 no filterbank
};

struct KbTcMelSmem { int hi, lo, a2, raw, f1, cs, csb, tw, w32, mag, outs, mag0, cw, cm, cg, bar, total; };
#define TCM_A2_BYTES 49152           // one stage-2 A operand set: re_hi, im_hi (TF32, 16 KB each), re_lo, im_lo (bf16, 8 KB each)
KB_HD KbTcMelSmem kb_tcm_smem_layout(int n_chunks) {
    KbTcMelSmem s;
    int off = 0;
    s.hi = off; off += TCM_MAXROWS * 128;                  // 19456 = 19 * 1024
    s.lo = off; off += TCM_MAXROWS * 128;
    s.a2 = off; off += 2 * TCM_A2_BYTES;                   // two A operand sets (double buffer: unit 0 / unit 1 of a tile)
    s.raw = off; off += TCM_MAXROWS * 128;                 // next tile's samples as they arrive (cp.async), linear
    s.f1 = off; off += 8192;
    s.cs = off; off += 16384;
    s.csb = off; off += 4096;
    s.tw = off; off += 18 * 32 * 8;
    s.w32 = off; off += 16 * 8;
    s.mag = off; off += TCM_UNIT * TCM_MS * 4;
    s.outs = off; off += ((TCM_UNIT * TCM_MP * 4 + 15) & ~15);
    s.mag0 = off; off += 2 * TCM_UNIT * 20 * 4;            // [slot][frame][17 -> 20]
    s.cw = off; off += n_chunks * 16;
    s.cm = off; off += ((n_chunks * 8 + 15) & ~15);
    s.cg = off; off += 36 * 4;
    s.bar = off; off += 128;
    s.total = off + 1024;                                  // + alignment slack
    return s;
}

#if defined(__CUDACC__)
namespace kbtc {

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void role_sync(int id) {          // named barrier of one 256-thread role
    asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory");
}
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
// nearest TF32 value (cvt.rna): the remainder v - tf32_rn(v) is exact in fp32 and half the size of the truncation remainder
__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ int brev5(int v) { return (int)(__brev((unsigned)v) >> 27); }
// two floats -> packed bf16 pair (lo at the lower address)
__device__ __forceinline__ uint32_t bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// D (+)= A . B with bf16 operands (K = 16), fp32 accumulate; always accumulating
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}

}  // namespace kbtc

__global__ void __launch_bounds__(TCM_THREADS, 1) kb_tc_mel_kernel(const __grid_constant__ KbTcMelParams p) {
    using namespace kbtc;
    extern __shared__ char kb_tcm_raw[];
    const uint32_t raw = smem_u32(kb_tcm_raw);
    char* sm = kb_tcm_raw + (((raw + 1023u) & ~1023u) - raw);
    const KbTcMelSmem L = kb_tcm_smem_layout(p.n_chunks);
    char* hi_s = sm + L.hi;
    char* lo_s = sm + L.lo;
    char* a2_s = sm + L.a2;
    float* raw_s = reinterpret_cast<float*>(sm + L.raw);
    float* f1_s = reinterpret_cast<float*>(sm + L.f1);
    float* cs_s = reinterpret_cast<float*>(sm + L.cs);
    unsigned short* csb_s = reinterpret_cast<unsigned short*>(sm + L.csb);
    cpx* tw_s = reinterpret_cast<cpx*>(sm + L.tw);
    cpx* w32_s = reinterpret_cast<cpx*>(sm + L.w32);
    float* mag_s = reinterpret_cast<float*>(sm + L.mag);
    float* out_s = reinterpret_cast<float*>(sm + L.outs);
    float* mag0_s = reinterpret_cast<float*>(sm + L.mag0);
    kb_f4* cw_s = reinterpret_cast<kb_f4*>(sm + L.cw);
    kb_i2* cm_s = reinterpret_cast<kb_i2*>(sm + L.cm);
    int* cg_s = reinterpret_cast<int*>(sm + L.cg);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.bar + 96);
    // MMA completion (tcgen05.commit, count 1):          s1_done[2] (stage 1 of unit 0 / 1), s2_done[2]
    // operand written / accumulator read (count 256):    d1_free[2] (X read D1), d2_free[2] (Y read D2 + mag0),
    //                                                    a2_ready[2] (X wrote stage-2 A operand set 0 / 1), smp_ready (X staged a tile)
    uint64_t* s1_done = bars;
    uint64_t* s2_done = bars + 2;
    uint64_t* d1_free = bars + 4;
    uint64_t* d2_free = bars + 6;
    uint64_t* a2_ready = bars + 8;                          // [2]: one per A operand set
    uint64_t* smp_ready = bars + 10;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time: constant tables, barriers, TMEM (256 columns) -------------------------------------------------
    for (int i = tid; i < 2048; i += TCM_THREADS) f1_s[i] = p.f1[i];
    for (int i = tid; i < 4096; i += TCM_THREADS) cs_s[i] = p.cs[i];
    for (int i = tid; i < 2048; i += TCM_THREADS) csb_s[i] = p.csb[i];
    for (int i = tid; i < 18 * 32; i += TCM_THREADS) { const float2 t = p.tw[i]; tw_s[i] = cmake(t.x, t.y); }
    for (int i = tid; i < 16; i += TCM_THREADS) { const float2 t = p.w32[i]; w32_s[i] = cmake(t.x, t.y); }
    for (int i = tid; i < p.n_chunks; i += TCM_THREADS) { cw_s[i] = p.cw[i]; cm_s[i] = p.cm[i]; }
    for (int i = tid; i <= 32; i += TCM_THREADS) cg_s[i] = p.cg[i];
    for (int i = tid; i < TCM_UNIT * TCM_MS; i += TCM_THREADS) mag_s[i] = 0.0f;     // incl. the pad bins, never written again
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) bar_init(bars + i, 1);
        for (int i = 4; i < 11; ++i) bar_init(bars + i, TCM_XT);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the constant operands (generic stores) -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");               // PDL: everything above touched constants only
    asm volatile("griddepcontrol.launch_dependents;");
    const uint32_t tmem = *tmem_slot;
    const int n_tiles = p.B * p.C * p.n_tiles_t;
    const int hop_rows = p.hop >> 5;
    const int tile_rows = (TCM_TF - 1) * hop_rows + 32;

    if (warp == 16) {
        // =========================================== role M: MMA issue ================================================
        if (lane == 0) {
            const uint32_t idesc1 = make_idesc_tf32(128, 32, 1, 0);
            const uint32_t idesc2 = idesc1;                      // stage 2 has the same shape / majors
            const uint32_t idesc2n = idesc1 | (1u << 13);        // negate A
            const uint32_t a_hi = smem_u32(hi_s), a_lo = smem_u32(lo_s);
            const uint32_t f_hi = smem_u32(f1_s), f_lo = f_hi + 4096;
            const uint32_t a2 = smem_u32(a2_s);
            const uint32_t c_hi = smem_u32(cs_s), c_lo = c_hi + 4096, s_hi = c_hi + 8192, s_lo = c_hi + 12288;
            const uint32_t lbo1 = (uint32_t)p.hop * 4u;
            // Descriptors are built once; inside the unrolled loops only their 14-bit start-address field (16-byte units)
            // moves, by compile-time constants.
            const uint64_t d_ah = make_desc(a_hi, lbo1, 512, 1), d_al = make_desc(a_lo, lbo1, 512, 1);
            const uint64_t d_fh = make_desc(f_hi, 512, 128, 0), d_fl = make_desc(f_lo, 512, 128, 0);
            // stage-2 A operand set b at a2 + b * TCM_A2_BYTES: re_hi, im_hi TF32 (SW128_32B, M-group = 2 frames, LBO 4096),
            // re_lo, im_lo bf16 (SWIZZLE_128B, MN atom = 64 elements = 4 frames, LBO 4096, 8-row K groups SBO 1024)
            const uint64_t d_rh = make_desc(a2 + 0, 4096, 512, 1), d_ih = make_desc(a2 + 16384, 4096, 512, 1);
            const uint64_t d_rl = make_desc(a2 + 32768, 4096, 1024, 2), d_il = make_desc(a2 + 40960, 4096, 1024, 2);
            const uint64_t d_ch = make_desc(c_hi, 512, 128, 0), d_cl = make_desc(c_lo, 512, 128, 0);
            const uint64_t d_sh = make_desc(s_hi, 512, 128, 0), d_sl = make_desc(s_lo, 512, 128, 0);
            const uint32_t cb = smem_u32(csb_s);
            const uint64_t d_cb = make_desc(cb, 512, 128, 0), d_sb = make_desc(cb + 2048, 512, 128, 0);
            // D = F32, A = B = BF16, A MN-major, B K-major, N = 32, M = 128 (kind::f16, K = 16 per instruction)
            const uint32_t idescb = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idescbn = idescb | (1u << 13);
            // stage 1 of unit `uu` (0 / 1) of the staged tile into D1 slot uu
            auto issue_s1 = [&](int uu) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (p.ablate & 4) { mma_commit(s1_done + uu); return; }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const uint32_t d = tmem + (uint32_t)(uu * 64 + mt * 32);
                    const uint64_t rowoff = (uint64_t)((uint32_t)((uu * TCM_UNIT + mt * 4) * hop_rows) * 8u);   // rows * 128 B / 16
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint64_t dah = d_ah + rowoff + (uint64_t)(j * 64), dal = d_al + rowoff + (uint64_t)(j * 64);
                        const uint64_t dbh = d_fh + (uint64_t)(j * 64), dbl = d_fl + (uint64_t)(j * 64);
                        if (j == 0) mma_tf32_c<false>(d, dal, dbh, idesc1); else mma_tf32_c<true>(d, dal, dbh, idesc1);
                        mma_tf32_c<true>(d, dah, dbl, idesc1);
                        mma_tf32_c<true>(d, dah, dbh, idesc1);
                    }
                }
                mma_commit(s1_done + uu);
            };
            // stage 2 of unit uu (A operand set uu) into D2 slot uu.  hi.hi and hi.lo products in TF32 (K = 8 per MMA), the
            // lo.hi products in bf16 (K = 16 per MMA): the lo parts are 2^-12 of the values, 8 bits of them are plenty.
            auto issue_s2 = [&](int uu) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (p.ablate & 4) { mma_commit(s2_done + uu); return; }
                const uint32_t dre = tmem + (uint32_t)(128 + uu * 64), dim = dre + 32;
                const uint64_t bo = (uint64_t)((uint32_t)uu * (TCM_A2_BYTES / 16));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t ko = (uint64_t)(j * 64);
                    const uint64_t rh = d_rh + bo + ko, ih = d_ih + bo + ko;
                    const uint64_t ch = d_ch + ko, cl = d_cl + ko, sh = d_sh + ko, sl = d_sl + ko;
                    // D_re += A_re C + A_im S
                    if (j == 0) mma_tf32_c<false>(dre, rh, cl, idesc2); else mma_tf32_c<true>(dre, rh, cl, idesc2);
                    mma_tf32_c<true>(dre, rh, ch, idesc2);
                    mma_tf32_c<true>(dre, ih, sl, idesc2);
                    mma_tf32_c<true>(dre, ih, sh, idesc2);
                    // D_im += A_im C - A_re S
                    if (j == 0) mma_tf32_c<false>(dim, ih, cl, idesc2); else mma_tf32_c<true>(dim, ih, cl, idesc2);
                    mma_tf32_c<true>(dim, ih, ch, idesc2);
                    mma_tf32_c<true>(dim, rh, sl, idesc2n);
                    mma_tf32_c<true>(dim, rh, sh, idesc2n);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint64_t ka = (uint64_t)(j * 128), kb = (uint64_t)(j * 64);   // 16 K rows = 2048 B of A, 2 K chunks = 1024 B of B
                    const uint64_t rl = d_rl + bo + ka, il = d_il + bo + ka;
                    mma_bf16(dre, rl, d_cb + kb, idescb);
                    mma_bf16(dre, il, d_sb + kb, idescb);
                    mma_bf16(dim, il, d_cb + kb, idescb);
                    mma_bf16(dim, rl, d_sb + kb, idescbn);
                }
                mma_commit(s2_done + uu);
            };
            int k = 0;
            if ((int)blockIdx.x < n_tiles) {
                bar_wait(smp_ready, 0u);
                issue_s1(0);
                issue_s1(1);
            }
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
                const bool has_next = tile + (int)gridDim.x < n_tiles;
                const uint32_t pk = (uint32_t)(k & 1), pk1 = (uint32_t)((k - 1) & 1);
                bar_wait(a2_ready + 0, pk);
                if (k > 0) bar_wait(d2_free + 0, pk1);
                issue_s2(0);
                if (has_next) {
                    bar_wait(smp_ready, pk ^ 1u);                 // tile k+1 staged (its (k+1)-th completion)
                    bar_wait(d1_free + 0, pk);                    // X has read D1 slot 0 of tile k
                    issue_s1(0);
                }
                bar_wait(a2_ready + 1, pk);
                if (k > 0) bar_wait(d2_free + 1, pk1);
                issue_s2(1);
                if (has_next) {
                    bar_wait(d1_free + 1, pk);
                    issue_s1(1);
                }
            }
        }
    } else if (warp < 8) {
        // =========================================== role X ===========================================================
        // ---- sample staging: global -> raw_s (cp.async, one tile ahead) -> (hi, lo) swizzled rows ---------------------
        // Element i of a tile is sample s0 + i of the signal (zero outside [0, L): pad_begin / pad_end / tails).  The copies
        // are issued a whole unit before their values are split, so their latency hides behind the between-stages step;
        // V = 4 / 2 / 1 floats per copy by the alignment of the tile's first sample.  A thread splits exactly the
        // elements it copied, so no barrier is needed in between.
        int pmode = 1;
        const int n_el = tile_rows * 32;
        auto tile_src = [&](int tile, const float*& xsig, long long& s0, int& t0, int& sig) {
            sig = tile / p.n_tiles_t;
            const int tt = tile - sig * p.n_tiles_t;
            const int b = sig / p.C, c = sig - b * p.C;
            t0 = tt * TCM_TF;
            xsig = p.x + (long long)b * p.x_sb + (long long)c * p.x_sc;
            s0 = (long long)t0 * p.hop - p.pad_left;
        };
        auto cp_elem = [&](int i, const float* xsig, long long s) {
            if (s >= 0 && s < p.L)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(raw_s + i)), "l"(xsig + s) : "memory");
            else raw_s[i] = 0.0f;
        };
        // Every thread owns the same 4-element groups i = 4 (tid + 256 it) in every alignment mode, so a thread that is
        // already copying the next tile never touches raw_s elements another thread has not split yet.
        auto prefetch = [&](const float* xsig, long long s0) {
            if (p.ablate & 16) return;
            const unsigned al = (unsigned)((reinterpret_cast<uintptr_t>(xsig) + (uintptr_t)(s0 * 4)) >> 2) & 3u;
            pmode = (al == 0) ? 4 : ((al & 1u) == 0 ? 2 : 1);
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const int i = 4 * (tid + TCM_XT * it);
                const long long s = s0 + i;
                if (i < n_el) {
                    const bool inside = s >= 0 && s + 3 < p.L;
                    if (inside && pmode == 4) {
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(raw_s + i)), "l"(xsig + s) : "memory");
                    } else if (inside && pmode == 2) {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(raw_s + i)), "l"(xsig + s) : "memory");
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(raw_s + i + 2)), "l"(xsig + s + 2) : "memory");
                    } else {
                        cp_elem(i, xsig, s); cp_elem(i + 1, xsig, s + 1); cp_elem(i + 2, xsig, s + 2); cp_elem(i + 3, xsig, s + 3);
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto commit_stage = [&]() {
            if (p.ablate & 16) { bar_arrive(smp_ready); return; }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const int i = 4 * (tid + TCM_XT * it);
                if (i < n_el) {
                    const float4 v = *reinterpret_cast<const float4*>(raw_s + i);
                    const uint32_t o = swz((uint32_t)i >> 5, (uint32_t)i & 31u);
                    float4 h, l;
                    h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                    *reinterpret_cast<float4*>(hi_s + o) = h;
                    *reinterpret_cast<float4*>(lo_s + o) = l;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bar_arrive(smp_ready);
        };

        // ---- between the stages, part 1 (registers only): one thread per row (frame = warp, n2 = lane) --------------------
        float tre[16], tim[16], mag0v = 0.0f;
        cpx y0v = cmake(0.0f, 0.0f);
        auto between_compute = [&](int uu) {
            const int n2 = lane;
            uint32_t sv[32];
            if (p.ablate & 64) { for (int i = 0; i < 32; ++i) sv[i] = (uint32_t)(lane + i); }
            else tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(uu * 64 + (warp >> 2) * 32), sv);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(d1_free + uu);                          // D1 slot uu may be overwritten by the next tile's stage 1
            if (p.ablate & 2) { for (int i = 0; i < 16; ++i) { tre[i] = __uint_as_float(sv[i]); tim[i] = __uint_as_float(sv[16 + i]); } return; }
            const float S0 = __uint_as_float(sv[0]), S16 = __uint_as_float(sv[1]);
            const float wc = p.wc;
            const cpx U0 = cmake(tw_s[n2].re * S0, 0.0f);
            cpx Uc = cmul(cmake(__uint_as_float(sv[2]), __uint_as_float(sv[3])), tw_s[1 * 32 + n2]);
            const float T0 = U0.re - 2.0f * wc * Uc.re;       // k1 = 0: real (U[-1] = conj U[1])
            cpx Um = U0;
#pragma unroll
            for (int k1 = 1; k1 <= 16; ++k1) {
                cpx Sn;
                if (k1 < 15) Sn = cmake(__uint_as_float(sv[2 * (k1 + 1)]), __uint_as_float(sv[2 * (k1 + 1) + 1]));
                else if (k1 == 15) Sn = cmake(S16, 0.0f);
                else Sn = cmake(__uint_as_float(sv[30]), -__uint_as_float(sv[31]));   // S[17] = conj S[15]
                const cpx Up = cmul(Sn, tw_s[(k1 + 1) * 32 + n2]);
                const cpx T = cfma_s(cadd(Um, Up), -wc, Uc);
                tre[k1 - 1] = T.re; tim[k1 - 1] = T.im;
                Um = Uc; Uc = Up;
            }
            // k1 = 0: 32-point DFT of the real sequence T0[n2] across the warp (radix-2 DIF, bit-reversed result)
            cpx y = cmake(T0, 0.0f);
#pragma unroll
            for (int s5 = 0; s5 < 5; ++s5) {
                const int half = 16 >> s5;
                cpx q;
                q.re = __shfl_xor_sync(0xffffffffu, y.re, half);
                q.im = __shfl_xor_sync(0xffffffffu, y.im, half);
                const cpx w = w32_s[(lane & (half - 1)) << s5];
                y = (lane & half) ? cmul(csub(q, y), w) : cadd(y, q);
            }
            y0v = y;
            mag0v = kb_sqrt(cnorm(y));
        };
        // part 2 (after A operand set uu and the mag0 slot are free): split, store, signal the issuer
        auto between_store = [&](int uu, int t0, int sig) {
            const int n2 = lane, fl = warp;
            if (p.ablate & 128) { bar_arrive(a2_ready + uu); return; }
            char* set = a2_s + uu * TCM_A2_BYTES;
            const uint32_t rowh = (uint32_t)((fl >> 1) * 4096 + n2 * 128);     // TF32 hi parts: M-group = frame pair
            const uint32_t rowl = (uint32_t)((fl >> 2) * 4096 + n2 * 128);     // bf16 lo parts: MN atom = 4 frames x 16
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t unit = (uint32_t)((fl & 1) * 2 + (i >> 1));
                const uint32_t o = rowh + (((unit ^ ((uint32_t)n2 & 3u)) << 5) | (uint32_t)((i & 1) << 4));
                float4 rh, ih;
                rh.x = tf32_rn(tre[4 * i]); rh.y = tf32_rn(tre[4 * i + 1]); rh.z = tf32_rn(tre[4 * i + 2]); rh.w = tf32_rn(tre[4 * i + 3]);
                ih.x = tf32_rn(tim[4 * i]); ih.y = tf32_rn(tim[4 * i + 1]); ih.z = tf32_rn(tim[4 * i + 2]); ih.w = tf32_rn(tim[4 * i + 3]);
                *reinterpret_cast<float4*>(set + o) = rh;
                *reinterpret_cast<float4*>(set + 16384 + o) = ih;
                tre[4 * i] -= rh.x; tre[4 * i + 1] -= rh.y; tre[4 * i + 2] -= rh.z; tre[4 * i + 3] -= rh.w;   // exact remainders
                tim[4 * i] -= ih.x; tim[4 * i + 1] -= ih.y; tim[4 * i + 2] -= ih.z; tim[4 * i + 3] -= ih.w;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {                                      // 8 bf16 = one 16-byte chunk per store
                const uint32_t chunk = (uint32_t)((fl & 3) * 2 + i);
                const uint32_t o = rowl + ((chunk ^ ((uint32_t)n2 & 7u)) << 4);
                uint4 rl, il;
                rl.x = bf16x2(tre[8 * i], tre[8 * i + 1]); rl.y = bf16x2(tre[8 * i + 2], tre[8 * i + 3]);
                rl.z = bf16x2(tre[8 * i + 4], tre[8 * i + 5]); rl.w = bf16x2(tre[8 * i + 6], tre[8 * i + 7]);
                il.x = bf16x2(tim[8 * i], tim[8 * i + 1]); il.y = bf16x2(tim[8 * i + 2], tim[8 * i + 3]);
                il.z = bf16x2(tim[8 * i + 4], tim[8 * i + 5]); il.w = bf16x2(tim[8 * i + 6], tim[8 * i + 7]);
                *reinterpret_cast<uint4*>(set + 32768 + o) = rl;
                *reinterpret_cast<uint4*>(set + 40960 + o) = il;
            }
            const int k2 = brev5(lane);
            if (k2 <= 16) {
                mag0_s[(uu * TCM_UNIT + fl) * 20 + k2] = mag0v;
                if (p.dbg) {
                    const int t = t0 + uu * TCM_UNIT + fl;
                    if (t < p.T) p.dbg[((long long)sig * p.T + t) * 513 + 32 * k2] = make_float2(y0v.re, y0v.im);
                }
            }
            if (!(p.ablate & 32)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // A operand: generic stores -> async proxy
            bar_arrive(a2_ready + uu);
        };

        // ---- tile loop (per thread, no barrier among the X warps) -----------------------------------------------------------
        int k = 0;                                              // this CTA's tile counter
        const float* xsig; long long s0; int t0, sig;
        if ((int)blockIdx.x < n_tiles) {
            tile_src((int)blockIdx.x, xsig, s0, t0, sig);
            prefetch(xsig, s0);
            commit_stage();
        }
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
            const bool has_next = tile + (int)gridDim.x < n_tiles;
            const int t0c = t0, sigc = sig;                      // this tile (for the debug dump)
            const uint32_t pk = (uint32_t)(k & 1), pk1 = (uint32_t)((k - 1) & 1);
            if (has_next) {
                tile_src(tile + (int)gridDim.x, xsig, s0, t0, sig);
                prefetch(xsig, s0);                             // raw_s was consumed by this thread's previous commit_stage
            }
            // ---- unit 0 ----
            bar_wait(s1_done + 0, pk);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            between_compute(0);
            if (k > 0) {
                bar_wait(s2_done + 0, pk1);                     // stage 2 of unit 0 of the previous tile has read operand set 0
                bar_wait(d2_free + 0, pk1);                     // Y has read mag0 slot 0 of the previous tile
            }
            between_store(0, t0c, sigc);
            // ---- unit 1 ----
            bar_wait(s1_done + 1, pk);                          // both stage-1 GEMMs of the tile are done: hi_s / lo_s are free
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (has_next) commit_stage();                       // early: the issuer can start the next tile's stage 1
            between_compute(1);
            if (k > 0) {
                bar_wait(s2_done + 1, pk1);
                bar_wait(d2_free + 1, pk1);
            }
            between_store(1, t0c, sigc);
        }
    } else {
        // =========================================== role Y ===========================================================
        const int ty = tid - TCM_XT, wy = warp - 8;
        const int q = wy & 3, h = wy >> 2;
        int k = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
            const int sig = tile / p.n_tiles_t;
            const int tt = tile - sig * p.n_tiles_t;
            const int b = sig / p.C, c = sig - b * p.C;
            const int t0 = tt * TCM_TF;
            const long long obase = (long long)b * p.o_sb + (long long)c * p.o_sc;
            float rmax = 0.0f;
            for (int uu = 0; uu < 2; ++uu) {
                bar_wait(s2_done + uu, (uint32_t)(k & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (p.ablate & 1) { bar_arrive(d2_free + uu); continue; }
                // ---- stage-2 accumulators -> magnitudes (row = (frame, k1), 16 of the 32 k2 per thread) ---------------
                {
                    const int fl = 2 * q + (lane >> 4), k1 = 1 + (lane & 15);
                    uint32_t re[16], im[16];
                    const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(128 + uu * 64 + 16 * h);
                    tmem_ld16(ta, re);
                    tmem_ld16(ta + 32, im);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    float m0 = 0.0f;
                    int f0 = 0, k2c = 0;
                    const bool copier = ty < TCM_UNIT * 17;     // bins 0, 32, .., 512 come from the X role's shuffle DFT
                    if (copier) { f0 = ty / 17; k2c = ty - f0 * 17; m0 = mag0_s[(uu * TCM_UNIT + f0) * 20 + k2c]; }
                    bar_arrive(d2_free + uu);                   // D2 slot uu and mag0 slot uu may be overwritten
                    float* mrow = mag_s + fl * TCM_MS;
                    const int t = t0 + uu * TCM_UNIT + fl;
                    if (!(h == 1 && k1 == 16)) {          // k1 = 16, k2 >= 16 mirrors onto the k1 = 16, k2 < 16 bins
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float xr = __uint_as_float(re[j]), xi = __uint_as_float(im[j]);
                            const int kk = k1 + 32 * (16 * h + j);
                            const int bin = h ? 1024 - kk : kk;
                            mrow[bin] = kb_sqrt(xr * xr + xi * xi);
                            if (p.dbg && t < p.T) p.dbg[((long long)sig * p.T + t) * 513 + bin] = make_float2(xr, h ? -xi : xi);
                        }
                    }
                    if (copier) mag_s[f0 * TCM_MS + 32 * k2c] = m0;
                }
                role_sync(2);
                // ---- mel filterbank: 32 lane groups x 8 frames walk the 4-bin chunk lists ----------------------------
                if (!(p.ablate & 8)) {
                    const int grp = ty >> 3, f = ty & 7;
                    const float* mrow = mag_s + f * TCM_MS;
                    float a0 = 0.0f, a1 = 0.0f;
                    const int ce = cg_s[grp + 1];
                    for (int i = cg_s[grp]; i < ce; ++i) {
                        const kb_f4 wv = cw_s[i];
                        const kb_i2 mt = cm_s[i];
                        const float2 m01 = *reinterpret_cast<const float2*>(mrow + mt.x);
                        const float2 m23 = *reinterpret_cast<const float2*>(mrow + mt.x + 2);
                        a0 += wv.x * m01.x; a1 += wv.y * m01.y;
                        a0 += wv.z * m23.x; a1 += wv.w * m23.y;
                        if (mt.y >= 0) { out_s[f * TCM_MP + mt.y] = a0 + a1; a0 = 0.0f; a1 = 0.0f; }
                    }
                }
                role_sync(2);
                // ---- decibel + coalesced store: warp = frame ----------------------------------------------------------
                {
                    const int t = t0 + uu * TCM_UNIT + wy;
                    if (t < p.T) {
                        float* orow = p.out + obase + (long long)t * p.o_st;
                        const float* srow = out_s + wy * TCM_MP;
                        for (int m = lane; m < p.n_bands; m += 32) {
                            float v = srow[m];
                            if (p.db) {
                                v = kb_floor_keepnan(v, p.amin);
                                rmax = kb_max_keepnan(rmax, v);
                                v = p.db_mul * kb_log2(v) - p.db_sub;
                            }
                            orow[(long long)m * p.o_sk] = v;
                        }
                    }
                }
                // the next unit's magnitudes may be written now (every Y thread passed the barrier after the filterbank);
                // out_s is rewritten only after the next role_sync
            }
            if (p.db) {
                const unsigned int wm = __reduce_max_sync(0xffffffffu, __float_as_uint(rmax));
                if (lane == 0 && wm != 0u) atomicMax(p.item_max + b, wm);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
}
#endif  // __CUDACC__
