// lv_harness.cpp -- the Linux test/bench harness of BASELINE.md section 4, in C++ over the C ABI.
// Links the product (liblivevideo.so) and, as the checker and the CPU baseline only, the oracle (liblv_oracle.so).
//
//   ./lv_harness [cif|1080p|4k|720p32|all] [--iters N] [--threads N]
//
// Prints one row per (config, implementation):
//   config, impl(cpu-restated|gpu), threads|gpus, Mpx/s, GB/s(5.5 B/px), %8TB/s, %6.551TB/s, bit_exact
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "livevideo.h"
#include "lv_oracle.h"

struct Config { const char *name; int w, h, streams, frames; };
static const Config kConfigs[] = {
    {"cif", 352, 288, 1, 300},        // BASELINE.json configs[0]
    {"1080p", 1920, 1080, 1, 64},     // configs[1]
    {"4k", 3840, 2160, 1, 32},        // configs[2]
    {"720p32", 1280, 720, 32, 16},    // configs[3]: one GPU's share of 256 streams at 8 GPUs
};

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
#define LV(x) do { int r_ = (x); if (r_ != LV_OK) { fprintf(stderr, "%s: %s (%s)\n", #x, lv_strerror(r_), lv_last_error()); exit(1); } } while (0)

static void row(const char *cfg, const char *impl, int units, double px, double sec, const char *exact)
{
    const double mpx = px / sec / 1e6, gbs = px * 5.5 / sec / 1e9;
    printf("%-8s %-13s %4d %12.0f %9.1f %7.1f%% %7.1f%%  %s\n", cfg, impl, units, mpx, gbs, gbs / 8000 * 100, gbs / 6551 * 100, exact);
}

static void run(const Config &c, int iters, int threads)
{
    const size_t px = (size_t)c.w * c.h, fb = px * 3 / 2;
    const int n = c.streams * c.frames;
    const int mbs = ((c.w + 15) / 16) * ((c.h + 15) / 16);
    // ---- GPU ------------------------------------------------------------------------------------------------
    lv_ctx *ctx;
    LV(lv_create(0, c.w, c.h, c.streams, 20, 0, &ctx));
    uint8_t *d_in, *d_bgr, *d_map; uint16_t *d_cnt; uint32_t *d_sum;
    CK(cudaMalloc(&d_in, fb * n)); CK(cudaMalloc(&d_bgr, 3 * px * n)); CK(cudaMalloc(&d_cnt, 2 * (size_t)mbs * n));
    CK(cudaMalloc(&d_map, (size_t)mbs * n)); CK(cudaMalloc(&d_sum, 8 * (size_t)n));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    LV(lv_gen_synthetic(1, 1, 0, 0, c.streams, c.frames, c.w, c.h, d_in, st));
    // first step from the zero state: this is the one compared with the oracle
    LV(lv_convert_diff_batch(ctx, d_in, 0, c.streams, c.frames, d_bgr, nullptr, nullptr, d_cnt, d_map, d_sum, st));
    CK(cudaStreamSynchronize(st));
    // ---- oracle on a bounded sample (first stream, up to 16 frames) -------------------------------------------
    const int sf = std::min(c.frames, 16);
    std::vector<uint8_t> h_in(fb * sf), h_bgr(3 * px * sf), o_bgr(3 * px * sf), prev(px, 0), h_map((size_t)mbs * sf), o_map((size_t)mbs * sf);
    std::vector<uint16_t> h_cnt((size_t)mbs * sf), o_cnt((size_t)mbs * sf);
    std::vector<uint32_t> h_sum(2 * sf), o_sum(2 * sf);
    CK(cudaMemcpy(h_in.data(), d_in, fb * sf, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_bgr.data(), d_bgr, 3 * px * sf, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_cnt.data(), d_cnt, 2 * (size_t)mbs * sf, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_map.data(), d_map, (size_t)mbs * sf, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_sum.data(), d_sum, 8 * (size_t)sf, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> gen(fb);
    lvo_gen_camera(1, 0, 0, c.w, c.h, gen.data());
    bool exact = memcmp(gen.data(), h_in.data(), fb) == 0;            // device generator == C generator
    lvo_convert_diff_batch(h_in.data(), prev.data(), 1, sf, c.w, c.h, 20, 0, o_bgr.data(), nullptr, o_cnt.data(), o_map.data(), o_sum.data(), threads);
    exact = exact && h_bgr == o_bgr && h_cnt == o_cnt && h_map == o_map && h_sum == o_sum;
    // ---- GPU timing -----------------------------------------------------------------------------------------------
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) LV(lv_convert_diff_batch(ctx, d_in, 0, c.streams, c.frames, d_bgr, nullptr, nullptr, d_cnt, d_map, d_sum, st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; i++) LV(lv_convert_diff_batch(ctx, d_in, 0, c.streams, c.frames, d_bgr, nullptr, nullptr, d_cnt, d_map, d_sum, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    row(c.name, "gpu", 1, (double)px * n, ms / iters * 1e-3, exact ? "yes" : "NO");
    // ---- CPU timing (same sample) -----------------------------------------------------------------------------------
    for (int t : {1, threads}) {
        double best = 1e30;
        for (int rep = 0; rep < 3; rep++) {
            std::fill(prev.begin(), prev.end(), 0);
            auto t0 = std::chrono::steady_clock::now();
            lvo_convert_diff_batch(h_in.data(), prev.data(), 1, sf, c.w, c.h, 20, 0, o_bgr.data(), nullptr, o_cnt.data(), o_map.data(), o_sum.data(), t);
            best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        row(c.name, "cpu-restated", t, (double)px * sf, best, "oracle");
        if (threads == 1) break;
    }
    cudaFree(d_in); cudaFree(d_bgr); cudaFree(d_cnt); cudaFree(d_map); cudaFree(d_sum);
    lv_destroy(ctx);
}

int main(int argc, char **argv)
{
    std::string which = "all";
    int iters = 50, threads = (int)std::thread::hardware_concurrency();
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else which = argv[i];
    }
    if (lv_device_count() < 1) { fprintf(stderr, "no CUDA device: the product has no CPU fallback\n"); return 1; }
    printf("%-8s %-13s %4s %12s %9s %8s %8s  %s\n", "config", "impl", "n", "Mpx/s", "GB/s", "%8TB/s", "%6.55TB", "bit_exact");
    for (const Config &c : kConfigs)
        if (which == "all" || which == c.name) run(c, iters, threads < 1 ? 1 : threads);
    return 0;
}
