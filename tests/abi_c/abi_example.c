/*
 * Plain-C client of include/kapre_b200.h: no Python, no torch -- the CUDA runtime for buffers and
 * libkapre_b200.so for the transform.  Computes the log-mel spectrogram of a waveform file and writes
 * it out; tests/test_gpu_parity.py::test_c_client_matches_oracle compiles this with gcc on the GPU box,
 * runs it and compares the result with the oracle.  It is also the shortest complete usage example of
 * the C ABI (what a cgo / JNI / ctypes binding does, in C).
 *
 *   abi_example <wave.f32> <batch> <length> <window.f32> <n_fft> <hop> <fb.f32> <n_mels> <out.f32>
 *
 * wave: (batch, length) float32, window: n_fft float32, fb: (n_fft/2+1, n_mels) float32 row-major.
 * out: (batch, frames, n_mels) float32 in decibel (ref 1.0, amin 1e-5, dynamic range 80).
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kapre_b200.h"

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } \
    } while (0)
#define KP(call)                                                                         \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, kapre_last_error()); return 3; }   \
    } while (0)

static float* read_f32(const char* path, size_t n) {
    FILE* f = fopen(path, "rb");
    float* p = (float*)malloc(n * sizeof(float));
    if (!f || !p || fread(p, sizeof(float), n, f) != n) { fprintf(stderr, "cannot read %zu floats from %s\n", n, path); exit(1); }
    fclose(f);
    return p;
}

int main(int argc, char** argv) {
    if (argc != 10) { fprintf(stderr, "usage: see the header of abi_example.c\n"); return 1; }
    const int B = atoi(argv[2]), L = atoi(argv[3]), n_fft = atoi(argv[5]), hop = atoi(argv[6]), n_mels = atoi(argv[8]);
    const int n_freq = n_fft / 2 + 1;
    float* wave = read_f32(argv[1], (size_t)B * L);
    float* window = read_f32(argv[4], (size_t)n_fft);
    float* fbm = read_f32(argv[7], (size_t)n_freq * n_mels);

    kapre_stft_plan* plan = NULL;
    kapre_filterbank* fb = NULL;
    KP(kapre_stft_plan_create(n_fft, n_fft, hop, window, &plan));
    KP(kapre_filterbank_create(fbm, n_freq, n_mels, &fb));
    const int T = kapre_stft_num_frames(plan, L, /*pad_begin=*/0, /*pad_end=*/0);
    if (!kapre_stft_supports_mode(plan, KAPRE_OUT_FB_DB)) { fprintf(stderr, "no fused log-mel for n_fft=%d\n", n_fft); return 4; }

    float *x_dev = NULL, *y_dev = NULL;
    void* ws_dev = NULL;
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    CK(cudaMalloc((void**)&x_dev, (size_t)B * L * sizeof(float)));
    CK(cudaMalloc((void**)&y_dev, (size_t)B * T * n_mels * sizeof(float)));
    CK(cudaMalloc(&ws_dev, (size_t)B * 8));
    CK(cudaMemsetAsync(ws_dev, 0, (size_t)B * 8, st));                     /* once: the workspace is self-cleaning */
    CK(cudaMemcpyAsync(x_dev, wave, (size_t)B * L * sizeof(float), cudaMemcpyHostToDevice, st));

    /* mono (batch, length) waveform -> (batch, frames, mel) spectrogram, described by element strides */
    const kapre_wave_desc xd = {B, 1, L, (int64_t)L, (int64_t)L, 1};
    const kapre_spec_desc yd = {(int64_t)T * n_mels, (int64_t)T * n_mels, (int64_t)n_mels, 1};
    const kapre_db_cfg db = {1.0f, 1e-5f, 80.0f};
    for (int rep = 0; rep < 2; ++rep)          /* twice: the second call reuses the workspace without a memset */
        KP(kapre_stft_forward(plan, x_dev, &xd, 0, 0, KAPRE_OUT_FB_DB, y_dev, &yd, fb, &db, ws_dev, (void*)st));

    float* out = (float*)malloc((size_t)B * T * n_mels * sizeof(float));
    CK(cudaMemcpyAsync(out, y_dev, (size_t)B * T * n_mels * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    FILE* f = fopen(argv[9], "wb");
    if (!f || fwrite(out, sizeof(float), (size_t)B * T * n_mels, f) != (size_t)B * T * n_mels) return 5;
    fclose(f);
    printf("frames=%d launches=%llu %s\n", T, (unsigned long long)kapre_launch_count(), kapre_last_launch_info());

    kapre_filterbank_destroy(fb);
    kapre_stft_plan_destroy(plan);
    cudaFree(x_dev); cudaFree(y_dev); cudaFree(ws_dev);
    cudaStreamDestroy(st);
    free(wave); free(window); free(fbm); free(out);
    return 0;
}
