// mempat.cu -- memory-system probe for the fused step's access pattern (no pixel math): which thread->pixel
// mapping streams 1.5 B/px in + 1 B/px previous luma (L2) + 3 B/px out fastest on B200?
//   P1 tile   : block 32 units x 8 row pairs (512 px x 16 rows), lane = unit            (v1 kernel)
//   P2 strip  : warp 8 units x 4 row pairs (128 px x 8 rows)                              (strip / TMA kernels)
//   P3 linear : items (row pair, unit) flattened row-major, 256 consecutive items per block
//   P4 tile + stores transposed through shared memory (every warp store = 512 contiguous bytes)
//   P5 linear + transposed stores
//   P6 wide   : block 64 units x 4 row pairs (1024 px x 8 rows)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

struct G { int w, h, units, nf; size_t px, frame; const uint8_t *in; uint8_t *out; };

__device__ __forceinline__ uint4 ldg16(const uint8_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ uint4 ldcs16(const uint8_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }   // .cs loads cost 9 % (profiles/README.md)
__device__ __forceinline__ uint2 ldcs8(const uint8_t *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
__device__ __forceinline__ void stcs16(uint8_t *p, const uint4 &v) { __stcs(reinterpret_cast<uint4 *>(p), v); }

template <bool kSmemStore, bool kFwdOut = false, bool kNoUV = false, bool kNoPrev = false>
__device__ __forceinline__ void item(const G &g, int f, int unit, int rp, uint8_t *sm /* per-warp 2*1536 B */)
{
    const int row = 2 * rp;
    const uint8_t *yf = g.in + (size_t)f * g.frame;
    const uint8_t *rf = f > 0 ? yf - g.frame : yf;
    const size_t off = (size_t)row * g.w + (size_t)unit * 16;
    const uint4 y0 = ldg16(yf + off), y1 = ldg16(yf + off + g.w);
    const uint4 p0 = kNoPrev ? y0 : ldcs16(rf + off), p1 = kNoPrev ? y1 : ldcs16(rf + off + g.w);
    const size_t coff = (size_t)rp * (g.w >> 1) + (size_t)unit * 8;
    const uint2 u = kNoUV ? make_uint2(y0.x, y0.y) : ldcs8(yf + g.px + coff), v = kNoUV ? make_uint2(y1.x, y1.y) : ldcs8(yf + g.px + (g.px >> 2) + coff);
    const uint4 q = make_uint4(u.x ^ p0.x, u.y ^ p0.y ^ p0.z, v.x ^ p1.x ^ p0.w, v.y ^ p1.y ^ p1.z ^ p1.w);
    uint8_t *d = g.out + (size_t)f * 3 * g.px + (size_t)(kFwdOut ? row : g.h - 1 - row) * 3 * g.w + (size_t)unit * 48;
    if (!kSmemStore) {
        stcs16(d, y0); stcs16(d + 16, q); stcs16(d + 32, y1);
        if (kFwdOut) d += 3 * (size_t)g.w; else d -= 3 * (size_t)g.w;
        stcs16(d, y1); stcs16(d + 16, y0); stcs16(d + 32, q);
    } else {
        // lanes hold 48 contiguous bytes each; re-deal so that each warp store writes 512 contiguous bytes.
        // Only valid when the 32 lanes of the warp are 32 consecutive units of one row (P4) -- P5 handles
        // the row break by falling back to direct stores for straddling warps.
        const int lane = threadIdx.x & 31;
        uint4 *s = reinterpret_cast<uint4 *>(sm);
        s[lane * 3 + 0] = y0; s[lane * 3 + 1] = q; s[lane * 3 + 2] = y1;
        s[96 + lane * 3 + 0] = y1; s[96 + lane * 3 + 1] = y0; s[96 + lane * 3 + 2] = q;
        __syncwarp();
        uint8_t *base = d - (size_t)lane * 48;          // lane 0's address
#pragma unroll
        for (int k = 0; k < 3; k++) stcs16(base + (k * 32 + lane) * 16, s[k * 32 + lane]);
        base -= 3 * (size_t)g.w;
#pragma unroll
        for (int k = 0; k < 3; k++) stcs16(base + (k * 32 + lane) * 16, s[96 + k * 32 + lane]);
        __syncwarp();
    }
}

// one 16-pixel unit of ONE row per thread (chroma row loaded by both rows of a pair: the second load hits L1/L2)
__device__ __forceinline__ void item1(const G &g, int f, int unit, int row)
{
    const uint8_t *yf = g.in + (size_t)f * g.frame;
    const uint8_t *rf = f > 0 ? yf - g.frame : yf;
    const size_t off = (size_t)row * g.w + (size_t)unit * 16;
    const uint4 y0 = ldg16(yf + off);
    const uint4 p0 = ldcs16(rf + off);
    const size_t coff = (size_t)(row >> 1) * (g.w >> 1) + (size_t)unit * 8;
    const uint2 u = __ldg(reinterpret_cast<const uint2 *>(yf + g.px + coff)), v = __ldg(reinterpret_cast<const uint2 *>(yf + g.px + (g.px >> 2) + coff));
    const uint4 q = make_uint4(u.x ^ p0.x, u.y ^ p0.y ^ p0.z, v.x ^ p0.w, v.y);
    uint8_t *d = g.out + (size_t)f * 3 * g.px + (size_t)(g.h - 1 - row) * 3 * g.w + (size_t)unit * 48;
    stcs16(d, y0); stcs16(d + 16, q); stcs16(d + 32, y0);
}

template <int MODE>
__global__ void __launch_bounds__(256) probe(const G g)
{
    __shared__ __align__(16) uint8_t sm[8][3072];
    const int f = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (MODE == 1 || MODE == 4) {           // tile: blockIdx.x = 32-unit tile, blockIdx.y = 8 row pairs
        const int unit = blockIdx.x * 32 + lane, rp = blockIdx.y * 8 + warp;
        if (unit < g.units && 2 * rp < g.h) {
            if (MODE == 4 && blockIdx.x * 32 + 32 <= g.units) item<true>(g, f, unit, rp, sm[warp]);
            else item<false>(g, f, unit, rp, sm[warp]);
        }
    } else if (MODE == 2) {                 // strip: warp = 8 units x 4 rp; block = 4 strips x 2 bands
        const int unit = (blockIdx.x * 4 + (warp & 3)) * 8 + (lane & 7), rp = (blockIdx.y * 2 + (warp >> 2)) * 4 + (lane >> 3);
        if (unit < g.units && 2 * rp < g.h) item<false>(g, f, unit, rp, sm[warp]);
    } else if (MODE == 3 || MODE == 5) {    // linear
        const int i = blockIdx.x * 256 + threadIdx.x;
        const int rp = i / g.units, unit = i - rp * g.units;
        if (2 * rp < g.h) {
            const int i0 = i - lane;        // warp's first item: same row?
            const bool same_row = (i0 / g.units) == ((i0 + 31) / g.units) && 2 * ((i0 + 31) / g.units) < g.h;
            if (MODE == 5 && same_row) item<true>(g, f, unit, rp, sm[warp]);
            else item<false>(g, f, unit, rp, sm[warp]);
        }
    } else if (MODE >= 7 && MODE <= 10) {   // linear variants: 7 forward output rows, 8 no chroma loads, 9 no prev loads, 10 all three
        const int i = blockIdx.x * 256 + threadIdx.x;
        const int rp = i / g.units, unit = i - rp * g.units;
        if (2 * rp < g.h) {
            if (MODE == 7) item<false, true, false, false>(g, f, unit, rp, sm[warp]);
            if (MODE == 8) item<false, false, true, false>(g, f, unit, rp, sm[warp]);
            if (MODE == 9) item<false, false, false, true>(g, f, unit, rp, sm[warp]);
            if (MODE == 10) item<false, true, true, true>(g, f, unit, rp, sm[warp]);
        }
    } else if (MODE == 11) {                // tile 32 units x 8 rows, one row per thread
        const int unit = blockIdx.x * 32 + lane, row = blockIdx.y * 8 + warp;
        if (unit < g.units && row < g.h) item1(g, f, unit, row);
    } else if (MODE == 12) {                // linear over (row, unit), one row per thread
        const int i = blockIdx.x * 256 + threadIdx.x;
        const int row = i / g.units, unit = i - row * g.units;
        if (row < g.h) item1(g, f, unit, row);
    } else if (MODE == 6) {                 // wide: 64 units x 4 rp
        const int unit = blockIdx.x * 64 + (warp & 1) * 32 + lane, rp = blockIdx.y * 4 + (warp >> 1);
        if (unit < g.units && 2 * rp < g.h) item<false>(g, f, unit, rp, sm[warp]);
    }
}

template <int MODE>
float run(const G &g, dim3 grid)
{
    for (int i = 0; i < 3; i++) probe<MODE><<<grid, 256>>>(g);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        for (int i = 0; i < 10; i++) probe<MODE><<<grid, 256>>>(g);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms / 10 < best) best = ms / 10;
    }
    if (cudaGetLastError() != cudaSuccess) printf("CUDA error\n");
    return best * 1e3f;
}

int main(int argc, char **argv)
{
    int cfgs[2][3] = {{1920, 1080, 64}, {3840, 2160, 32}};
    for (auto &c : cfgs) {
        G g;
        g.w = c[0]; g.h = c[1]; g.nf = c[2]; g.units = g.w / 16; g.px = (size_t)g.w * g.h; g.frame = g.px * 3 / 2;
        uint8_t *in, *out;
        cudaMalloc(&in, g.frame * g.nf); cudaMalloc(&out, 3 * g.px * g.nf);
        cudaMemset(in, 1, g.frame * g.nf);
        g.in = in; g.out = out;
        const int rps = g.h / 2;
        const double px = (double)g.px * g.nf;
        float t;
        t = run<1>(g, dim3((g.units + 31) / 32, (rps + 7) / 8, g.nf));  printf("%dx%dx%d P1 tile        %7.1f us  %6.0f GB/s alg\n", g.w, g.h, g.nf, t, px * 5.5 / t / 1e3);
        t = run<2>(g, dim3((g.units + 31) / 32, (rps + 7) / 8, g.nf));  printf("%dx%dx%d P2 strip       %7.1f us  %6.0f GB/s alg\n", g.w, g.h, g.nf, t, px * 5.5 / t / 1e3);
        t = run<3>(g, dim3((g.units * rps + 255) / 256, 1, g.nf));      printf("%dx%dx%d P3 linear      %7.1f us  %6.0f GB/s alg\n", g.w, g.h, g.nf, t, px * 5.5 / t / 1e3);
        t = run<4>(g, dim3((g.units + 31) / 32, (rps + 7) / 8, g.nf));  printf("%dx%dx%d P4 tile+smemst %7.1f us  %6.0f GB/s alg\n", g.w, g.h, g.nf, t, px * 5.5 / t / 1e3);
        t = run<5>(g, dim3((g.units * rps + 255) / 256, 1, g.nf));      printf("%dx%dx%d P5 lin+smemst  %7.1f us  %6.0f GB/s alg\n", g.w, g.h, g.nf, t, px * 5.5 / t / 1e3);
        t = run<6>(g, dim3((g.units + 63) / 64, (rps + 3) / 4, g.nf));  printf("%dx%dx%d P6 wide        %7.1f us  %6.0f GB/s alg\n", g.w, g.h, g.nf, t, px * 5.5 / t / 1e3);
        t = run<7>(g, dim3((g.units * rps + 255) / 256, 1, g.nf));      printf("%dx%dx%d P7 lin fwd-out %7.1f us  %6.0f GB/s alg\n", g.w, g.h, g.nf, t, px * 5.5 / t / 1e3);
        t = run<8>(g, dim3((g.units * rps + 255) / 256, 1, g.nf));      printf("%dx%dx%d P8 lin no-UV   %7.1f us  (4 B/px DRAM)\n", g.w, g.h, g.nf, t);
        t = run<9>(g, dim3((g.units * rps + 255) / 256, 1, g.nf));      printf("%dx%dx%d P9 lin no-prev %7.1f us\n", g.w, g.h, g.nf, t);
        t = run<10>(g, dim3((g.units * rps + 255) / 256, 1, g.nf));     printf("%dx%dx%d P10 lin fwd,noUV,noprev %7.1f us (4 B/px DRAM)\n", g.w, g.h, g.nf, t);
        t = run<11>(g, dim3((g.units + 31) / 32, (g.h + 7) / 8, g.nf)); printf("%dx%dx%d P11 tile 1row/thr %7.1f us  %6.0f GB/s alg\n", g.w, g.h, g.nf, t, px * 5.5 / t / 1e3);
        t = run<12>(g, dim3((g.units * g.h + 255) / 256, 1, g.nf));     printf("%dx%dx%d P12 lin  1row/thr %7.1f us  %6.0f GB/s alg\n", g.w, g.h, g.nf, t, px * 5.5 / t / 1e3);
        cudaFree(in); cudaFree(out);
    }
    return 0;
}
