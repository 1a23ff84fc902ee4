// lv_formats.cu -- the other five converters of the reference's cscc.lib (Convert.h:25-30), SURVEY.md 8(f) rank 4.
// The application never calls them (grep pCon. -> only YV12_to_RGB24), so they are written for clarity, not for
// the roofline: one thread per output byte (or per 2x2 block for the RGB one), coalesced over the output.
// Semantics restated from the object code (Convert.obj): RGB24_to_YV12 .text+0x1d0 (packed 10-bit tables built by
// the constructor at .text+0x025..0x10a), YVU9_to_YV12 +0x740, YUY2_to_YV12 +0x8b0, YV12_to_YVU9 +0x960,
// YV12_to_YUY2 +0xa40; pinned bit-exact against that object code through the oracle (tests/test_oracle.py,
// tests/test_gpu_parity.py).
#include "lv_kernels.h"

namespace lv {

namespace {

// r2yuv / g2yuv / b2yuv: Y | Cb<<10 | Cr<<20, every contribution the truncation (toward zero) of coef * i in
// double precision, negative ones kept as their low 8 bits (the constructor's `and 0xff; shl`)
__device__ __forceinline__ void build_rgb_tables(uint32_t (&t)[3][256])
{
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const uint32_t half = (uint32_t)(int)(0.5 * i);
        t[0][i] = (half << 20) | (((uint32_t)(int)(-0.1687 * i) & 0xffu) << 10) | (uint32_t)(int)(0.299 * i);
        t[1][i] = (((uint32_t)(int)(-0.4187 * i) & 0xffu) << 20) | (((uint32_t)(int)(-0.3313 * i) & 0xffu) << 10) |
                  (uint32_t)(int)(0.587 * i);
        t[2][i] = (((uint32_t)(int)(-0.0813 * i) & 0xffu) << 20) | (half << 10) | (uint32_t)(int)(0.114 * i);
    }
    __syncthreads();
}

// one thread per 2x2 block; input rows are bottom-up, bytes R,G,B; output Y, Cb, Cr planes
__global__ void __launch_bounds__(256) k_rgb24_to_yv12(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h)
{
    __shared__ uint32_t t[3][256];
    build_rgb_tables(t);
    const size_t px = (size_t)w * h;
    in += (size_t)blockIdx.y * 3 * px;
    out += (size_t)blockIdx.y * (px + 2 * (px / 4));
    const int bw = (w + 1) >> 1, bh = h >> 1;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < bw * bh; b += gridDim.x * blockDim.x) {
        const int by = b / bw, bx = b - by * bw;
        const uint8_t *p = in + (size_t)(h - 1 - 2 * by) * 3 * w + 6 * bx;    // top image row of the block
        const uint8_t *q = p - 3 * w;                                         // the row below it
        const uint32_t s = t[0][p[0]] + t[1][p[1]] + t[2][p[2]];
        uint8_t *y = out + (size_t)(2 * by) * w + 2 * bx;
        y[0] = (uint8_t)s;
        y[1] = (uint8_t)(t[0][p[3]] + t[1][p[4]] + t[2][p[5]]);
        y[w] = (uint8_t)(t[0][q[0]] + t[1][q[1]] + t[2][q[2]]);
        y[w + 1] = (uint8_t)(t[0][q[3]] + t[1][q[4]] + t[2][q[5]]);
        out[px + (size_t)by * bw + bx] = (uint8_t)((s >> 10) ^ 0x80u);
        out[px + px / 4 + (size_t)by * bw + bx] = (uint8_t)((s >> 20) ^ 0x80u);
    }
}

// kind: 2 YVU9->YV12, 3 YUY2->YV12, 4 YV12->YVU9, 5 YV12->YUY2; one thread per output byte
template <int KIND>
__global__ void __launch_bounds__(256) k_shuffle(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h,
                                                 size_t in_frame, size_t out_frame)
{
    const size_t px = (size_t)w * h;
    const int cw2 = w / 2, cw4 = w / 4, ch4 = h / 4;
    in += (size_t)blockIdx.y * in_frame;
    out += (size_t)blockIdx.y * out_frame;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < out_frame; i += (size_t)gridDim.x * blockDim.x) {
        uint8_t v;
        if (KIND == 2) {            // planes read bottom-up, chroma replicated 2x2
            if (i < px) {
                const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
                v = in[(size_t)(h - 1 - y) * w + x];
            } else {
                const size_t c = i - px;
                const int pl = c >= px / 4, y = (int)((c - pl * (px / 4)) / cw2), x = (int)((c - pl * (px / 4)) - (size_t)y * cw2);
                v = in[px + (size_t)pl * ((size_t)h * cw4 / 4) + (size_t)(ch4 - 1 - (y >> 1)) * cw4 + (x >> 1)];
            }
        } else if (KIND == 3) {     // Y0 U Y1 V rows; byte 3 -> first chroma plane, byte 1 -> second; odd rows' chroma
            if (i < px) {
                v = in[2 * i];
            } else {
                const size_t c = i - px;
                const int pl = c >= px / 4, y = (int)((c - pl * (px / 4)) / cw2), x = (int)((c - pl * (px / 4)) - (size_t)y * cw2);
                v = in[(size_t)(2 * y + 1) * 2 * w + 4 * x + (pl ? 1 : 3)];
            }
        } else if (KIND == 4) {     // luma copied; chroma: even samples of even rows
            if (i < px) {
                v = in[i];
            } else {
                const size_t c = i - px, plane = (size_t)cw4 * ch4;
                const int pl = c >= plane, y = (int)((c - pl * plane) / cw4), x = (int)((c - pl * plane) - (size_t)y * cw4);
                v = in[px + (size_t)pl * (px / 4) + (size_t)(2 * y) * cw2 + 2 * x];
            }
        } else {                    // KIND 5: Y0 [2nd plane] Y1 [1st plane]
            const int y = (int)(i / (2 * (size_t)w)), k = (int)(i - (size_t)y * 2 * w);
            const int x2 = k >> 2, sel = k & 3;
            if ((sel & 1) == 0) v = in[(size_t)y * w + 2 * x2 + (sel >> 1)];
            else v = in[px + (sel == 1 ? px / 4 : 0) + (size_t)(y >> 1) * cw2 + x2];
        }
        out[i] = v;
    }
}

}  // namespace

size_t format_bytes_in(int kind, int w, int h)
{
    static const int num[6] = {12, 24, 9, 16, 12, 12};
    return (size_t)w * h / 8 * num[kind];
}
size_t format_bytes_out(int kind, int w, int h)
{
    static const int num[6] = {24, 12, 12, 12, 9, 16};
    return (size_t)w * h / 8 * num[kind];
}

// kind 1..5 as in lv_convert_format (include/livevideo.h)
int launch_format(int kind, int w, int h, const uint8_t *in, uint8_t *out, int n_frames, cudaStream_t st)
{
    if (n_frames <= 0) return 0;
    const size_t inb = format_bytes_in(kind, w, h), outb = format_bytes_out(kind, w, h);
    int launches = 0;
    for (int f0 = 0; f0 < n_frames; f0 += 65535) {
        const int nf = n_frames - f0 < 65535 ? n_frames - f0 : 65535;
        const uint8_t *i0 = in + (size_t)f0 * inb;
        uint8_t *o0 = out + (size_t)f0 * outb;
        unsigned bx = (unsigned)((outb + 255) / 256);
        if (bx > 4096) bx = 4096;
        if (bx < 1) bx = 1;
        const dim3 grid(bx, nf);
        switch (kind) {
        case 1: k_rgb24_to_yv12<<<grid, 256, 0, st>>>(i0, o0, w, h); break;
        case 2: k_shuffle<2><<<grid, 256, 0, st>>>(i0, o0, w, h, inb, outb); break;
        case 3: k_shuffle<3><<<grid, 256, 0, st>>>(i0, o0, w, h, inb, outb); break;
        case 4: k_shuffle<4><<<grid, 256, 0, st>>>(i0, o0, w, h, inb, outb); break;
        case 5: k_shuffle<5><<<grid, 256, 0, st>>>(i0, o0, w, h, inb, outb); break;
        default: return -(int)cudaErrorInvalidValue;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return -(int)e;
        launches++;
    }
    return launches;
}

}  // namespace lv
