// Tensor-core DFT stages (tcgen05 / TMEM, sm_100a) -- measured prototype for the real FFT of the fused
// forward kernel, n_fft = 1024 = 32 x 32 (SURVEY section 7 "hard part 1").
//
// Cooley-Tukey with n = 32 n1 + n2, k = k1 + 32 k2:
//   stage 1   S[f, n2, k1] = sum_n1 x[hop f + 32 n1 + n2] exp(-2 pi i n1 k1 / 32)          (this file)
//   twiddle   T = S * exp(-2 pi i n2 k1 / 1024);   stage 2   X[k1 + 32 k2] = sum_n2 T exp(-2 pi i n2 k2 / 32)
// Stage 1 as ONE GEMM per tile of 4 frames:  D[(f, n2), c] = A[(f, n2), n1] . F[n1, c],  M = 128, K = 32, N = 32.
//   * A is the RAW hop-overlapped sample buffer: element ((f, n2), n1) = x[32 (8 f + n1) + n2], i.e. the
//     natural [row of 32 samples][32] view is exactly an MN-major operand with 128-byte K rows -- the
//     SWIZZLE_128B_BASE32B MN-major canonical layout of the UMMA shared-memory descriptor (the one TF32 MN-major
//     operands must use; cute/atom/mma_traits_sm100.hpp Layout_MN_SW128_32B_Atom: atoms of 4 K rows x 128 bytes).
//     Frames overlap: the M-group stride (LBO) is hop * 4 bytes, a K step of 8 rows advances the start address by
//     1024 bytes.  No per-frame copy of the samples is made.
//   * x is real, so only k1 = 0..16 are needed: columns c = (Re, Im) of k1 = 0..15; the identically-zero
//     Im(k1 = 0) column carries k1 = 16 (also real), so N = 32 with no padding.
//   * fp32-grade accuracy from TF32 operands by the 3-product split (hi = top 19 bits, lo = exact remainder):
//     D = A_lo.F_hi + A_hi.F_lo + A_hi.F_hi, accumulated in TMEM (fp32).
// One elected thread issues the tcgen05.mma instructions; completion is signalled through tcgen05.commit
// on an mbarrier; the accumulators are read back with tcgen05.ld (32 lanes x 32 columns per warp).
#pragma once
#include <stdint.h>

#define KB_TC_TILE_F 32                      // frames per CTA tile (8 MMA tiles of 4 frames)
#define KB_TC_HOP 256
#define KB_TC_NFFT 1024
#define KB_TC_ROWS ((KB_TC_TILE_F - 1) * (KB_TC_HOP / 32) + KB_TC_NFFT / 32)   // 128-byte sample rows per tile: 280
#define KB_TC_THREADS 256

struct KbTcParams {
    const float* x;            // (items, item_stride) waveforms
    long long item_stride;
    int n_items, length, T;    // T frames per item (hop 256, n_fft 1024, no padding)
    float* out;                // store != 0: (items, T, 32 n2, 32 cols) stage-1 output; else one float per thread
    int store;
    const float* fmat;         // (2, 8 kchunks, 32 cols, 4) hi / lo stage-1 matrix in K-major interleaved order
    int n_tiles_t;             // ceil(T / KB_TC_TILE_F)
};

#if defined(__CUDACC__)
namespace kbtc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout type << 61
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, A MN-major (bit 15), B K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// same with a compile-time accumulate flag (no predicate set-up per instruction in unrolled issue loops)
template <bool ACC>
__device__ __forceinline__ void mma_tf32_c(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    if (ACC)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// bounded wait: a descriptor / barrier mistake must not hang the GPU box -- trap instead
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t b = smem_u32(bar);
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(b), "r"(parity) : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// byte offset of sample (row u, column c) inside a buffer of 128-byte rows in the SWIZZLE_128B_BASE32B pattern --
// "for mn-major tf32 operands, SW128_32B is the only available smem layout" (cutlass sm100_common.inl:92):
// cute Swizzle<2,5,2>: the 32-byte unit index (address bits 5-6) is XORed with the row index mod 4 (bits 7-8).
__device__ __forceinline__ uint32_t swz(uint32_t u, uint32_t c) {
    return u * 128u + ((((c >> 3) ^ (u & 3u)) << 5) | ((c & 7u) << 2));
}

}  // namespace kbtc

// Shared-memory carve-up (bytes, relative to the 1024 B-aligned base)
#define KB_TC_OFF_HI 0
#define KB_TC_OFF_LO (KB_TC_ROWS * 128)                     // 35840, a multiple of 1024
#define KB_TC_OFF_F (2 * KB_TC_ROWS * 128)                  // F hi (4 KB) then F lo (4 KB)
#define KB_TC_OFF_BAR (KB_TC_OFF_F + 8192)
#define KB_TC_SMEM (KB_TC_OFF_BAR + 64 + 1024)              // + alignment slack

__global__ void __launch_bounds__(KB_TC_THREADS, 2) kb_tc_dft_stage1_kernel(const __grid_constant__ KbTcParams p) {
    using namespace kbtc;
    extern __shared__ char kb_tc_raw[];
    const uint32_t raw = smem_u32(kb_tc_raw);
    char* sm = kb_tc_raw + (((raw + 1023u) & ~1023u) - raw);
    char* hi_s = sm + KB_TC_OFF_HI;
    char* lo_s = sm + KB_TC_OFF_LO;
    float* f_s = reinterpret_cast<float*>(sm + KB_TC_OFF_F);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + KB_TC_OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + KB_TC_OFF_BAR + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time: matrix to shared memory, barrier, TMEM allocation (256 columns = 8 tiles x 32) ----
    for (int i = tid; i < 2048; i += KB_TC_THREADS) f_s[i] = p.fmat[i];
    if (tid == 0) bar_init(bar, 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = make_idesc_tf32(128, 32, 1, 0);
    const uint32_t f_hi_addr = smem_u32(f_s), f_lo_addr = f_hi_addr + 4096;
    const uint32_t a_hi_addr = smem_u32(hi_s), a_lo_addr = smem_u32(lo_s);

    const int n_tiles = p.n_items * p.n_tiles_t;
    uint32_t parity = 0;
    float sink = 0.0f;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int item = tile / p.n_tiles_t;
        const int t0 = (tile - item * p.n_tiles_t) * KB_TC_TILE_F;
        const float* xi = p.x + (long long)item * p.item_stride + (long long)t0 * KB_TC_HOP;
        const int avail = p.length - t0 * KB_TC_HOP;               // samples of this item from the tile start
        // ---- stage the tile's samples: hi = top 19 bits (what a TF32 operand keeps), lo = exact remainder ----
        for (int i = tid; i < KB_TC_ROWS * 32; i += KB_TC_THREADS) {
            const float v = i < avail ? __ldg(xi + i) : 0.0f;
            const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            const uint32_t o = swz((uint32_t)i >> 5, (uint32_t)i & 31u);
            *reinterpret_cast<float*>(hi_s + o) = h;
            *reinterpret_cast<float*>(lo_s + o) = v - h;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (tensor core)
        __syncthreads();
        // ---- one thread issues 8 tiles x 4 K steps x 3 split products ----------------------------------------
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int mt = 0; mt < KB_TC_TILE_F / 4; ++mt) {
                const uint32_t d = tmem + (uint32_t)(mt * 32);
                uint32_t acc = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // A: rows 8 (4 mt + f) + 8 j ..; M groups of 32 (one frame) are hop * 4 = 1024 bytes apart
                    const uint32_t aoff = (uint32_t)(mt * 4 * 1024 + j * 1024);
                    // layout type 1 = SWIZZLE_128B_BASE32B: atoms of 4 K rows (SBO = 512 bytes between them)
                    const uint64_t a_hi = make_desc(a_hi_addr + aoff, 1024, 512, 1);
                    const uint64_t a_lo = make_desc(a_lo_addr + aoff, 1024, 512, 1);
                    // B (K-major, no swizzle): [kchunk][col][4]: 8-column groups 128 B apart (SBO), K chunks 512 B (LBO)
                    const uint64_t b_hi = make_desc(f_hi_addr + (uint32_t)(j * 1024), 512, 128, 0);
                    const uint64_t b_lo = make_desc(f_lo_addr + (uint32_t)(j * 1024), 512, 128, 0);
                    mma_tf32(d, a_lo, b_hi, idesc, acc); acc = 1;
                    mma_tf32(d, a_hi, b_lo, idesc, 1);
                    mma_tf32(d, a_hi, b_hi, idesc, 1);
                }
            }
            mma_commit(bar);
        }
        // ---- epilogue: TMEM -> registers -> global ------------------------------------------------------------
        bar_wait(bar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            const int mt = (warp >> 2) * 4 + q;                       // warps 0-3: tiles 0-3, warps 4-7: tiles 4-7
            const int f = t0 + mt * 4 + (warp & 3);                   // TMEM lanes 32 (warp % 4) .. +31 = one frame
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * 32), v);
            if (p.store) {
                if (f < p.T) {
                    float4* o = reinterpret_cast<float4*>(p.out + (((long long)item * p.T + f) * 32 + lane) * 32);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        o[c] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]),
                                           __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3]));
                }
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) sink += __uint_as_float(v[c]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                              // TMEM and the sample buffers are free again
    }
    if (!p.store) p.out[(long long)blockIdx.x * KB_TC_THREADS + tid] = sink;
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
}
#endif  // __CUDACC__
