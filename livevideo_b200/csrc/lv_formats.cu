// lv_formats.cu -- the other five converters of the reference's cscc.lib (Convert.h:25-30), SURVEY.md 8(f) rank 4.
// The application never calls them (grep pCon. -> only YV12_to_RGB24).  Two sets of kernels:
//   * width % 32 == 0: vector kernels -- every thread moves 16-byte words (one 16/32-pixel unit over 2 or 4 rows), bytes
//     are re-dealt in registers with PRMT, and RGB24_to_YV12 replaces the constructor's packed 10-bit tables by the
//     exact integer form of each entry ((int)(c*i) == (m*i) >> 16 for all i in 0..255, multipliers below);
//   * any other legal geometry: one thread per output byte (or per 2x2 block for the RGB one), the plain restatement.
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


// ---- vector kernels (w % 32 == 0, h % 4 == 0) --------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg16(const uint8_t *p) { return __ldcs(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ uint2 ldg8(const uint8_t *p) { return __ldcs(reinterpret_cast<const uint2 *>(p)); }
__device__ __forceinline__ void stg16(uint8_t *p, const uint4 &v) { __stcs(reinterpret_cast<uint4 *>(p), v); }
__device__ __forceinline__ void stg8(uint8_t *p, const uint2 &v) { __stcs(reinterpret_cast<uint2 *>(p), v); }
__device__ __forceinline__ uint32_t bp(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }

// Exact integer form of the constructor's tables (Convert.obj .text+0x025..0x10a): (int)(c * i) for i in 0..255 equals
// (m * i) >> 16 with these multipliers (checked for all 256 i; the GPU test runs all 2^24 triples against the oracle).
//   0.299 -> 19596   0.587 -> 38472   0.114 -> 7472   0.1687 -> 11056   0.3313 -> 21712   0.4187 -> 27440   0.0813 -> 5328
__device__ __forceinline__ uint32_t rgb_y(uint32_t r, uint32_t g, uint32_t b)
{
    return ((r * 19596u) >> 16) + ((g * 38472u) >> 16) + ((b * 7472u) >> 16);     // <= 254: no carry out of the byte
}
__device__ __forceinline__ uint32_t rgb_cb(uint32_t r, uint32_t g, uint32_t b)
{
    return ((b >> 1) - ((r * 11056u) >> 16) - ((g * 21712u) >> 16) + 128u) & 0xffu;   // fields add mod 256, then ^0x80
}
__device__ __forceinline__ uint32_t rgb_cr(uint32_t r, uint32_t g, uint32_t b)
{
    return ((r >> 1) - ((g * 27440u) >> 16) - ((b * 5328u) >> 16) + 128u) & 0xffu;
}
template <int K>
__device__ __forceinline__ uint32_t byte_of(const uint32_t (&wd)[12])      // byte K of 48 packed bytes
{
    return bp(wd[K >> 2], 0u, 0x4440u | (uint32_t)(K & 3));
}
template <int P>
__device__ __forceinline__ uint32_t y_of(const uint32_t (&wd)[12])
{
    return rgb_y(byte_of<3 * P>(wd), byte_of<3 * P + 1>(wd), byte_of<3 * P + 2>(wd));
}
__device__ __forceinline__ uint4 y16_of(const uint32_t (&wd)[12])
{
    uint4 o;
    o.x = y_of<0>(wd) | (y_of<1>(wd) << 8) | (y_of<2>(wd) << 16) | (y_of<3>(wd) << 24);
    o.y = y_of<4>(wd) | (y_of<5>(wd) << 8) | (y_of<6>(wd) << 16) | (y_of<7>(wd) << 24);
    o.z = y_of<8>(wd) | (y_of<9>(wd) << 8) | (y_of<10>(wd) << 16) | (y_of<11>(wd) << 24);
    o.w = y_of<12>(wd) | (y_of<13>(wd) << 8) | (y_of<14>(wd) << 16) | (y_of<15>(wd) << 24);
    return o;
}
template <int P>
__device__ __forceinline__ void c_of(const uint32_t (&wd)[12], uint32_t &cb, uint32_t &cr)   // pixel P (even) of the top row
{
    const uint32_t r = byte_of<3 * P>(wd), g = byte_of<3 * P + 1>(wd), b = byte_of<3 * P + 2>(wd);
    cb = rgb_cb(r, g, b);
    cr = rgb_cr(r, g, b);
}

// work item = 16 pixels x one row pair; grid.y = frame
__global__ void __launch_bounds__(256) k_rgb24_to_yv12_v(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h)
{
    const size_t px = (size_t)w * h;
    in += (size_t)blockIdx.y * 3 * px;
    out += (size_t)blockIdx.y * (px + 2 * (px / 4));
    const int upr = w >> 4, items = upr * (h >> 1);
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
        const int by = it / upr, u = it - by * upr;
        const uint8_t *p = in + (size_t)(h - 1 - 2 * by) * 3 * w + 48 * u;      // top image row of the pair (input is bottom-up)
        const uint8_t *q = p - 3 * (size_t)w;
        uint32_t a[12], b[12];
        const uint4 p0 = ldg16(p), p1 = ldg16(p + 16), p2 = ldg16(p + 32);
        const uint4 q0 = ldg16(q), q1 = ldg16(q + 16), q2 = ldg16(q + 32);
        a[0] = p0.x; a[1] = p0.y; a[2] = p0.z; a[3] = p0.w; a[4] = p1.x; a[5] = p1.y; a[6] = p1.z; a[7] = p1.w;
        a[8] = p2.x; a[9] = p2.y; a[10] = p2.z; a[11] = p2.w;
        b[0] = q0.x; b[1] = q0.y; b[2] = q0.z; b[3] = q0.w; b[4] = q1.x; b[5] = q1.y; b[6] = q1.z; b[7] = q1.w;
        b[8] = q2.x; b[9] = q2.y; b[10] = q2.z; b[11] = q2.w;
        uint8_t *y = out + (size_t)(2 * by) * w + 16 * u;
        stg16(y, y16_of(a));
        stg16(y + w, y16_of(b));
        uint32_t cb[8], cr[8];
        c_of<0>(a, cb[0], cr[0]); c_of<2>(a, cb[1], cr[1]); c_of<4>(a, cb[2], cr[2]); c_of<6>(a, cb[3], cr[3]);
        c_of<8>(a, cb[4], cr[4]); c_of<10>(a, cb[5], cr[5]); c_of<12>(a, cb[6], cr[6]); c_of<14>(a, cb[7], cr[7]);
        const size_t co = (size_t)by * (w >> 1) + 8 * u;
        stg8(out + px + co, make_uint2(cb[0] | (cb[1] << 8) | (cb[2] << 16) | (cb[3] << 24),
                                       cb[4] | (cb[5] << 8) | (cb[6] << 16) | (cb[7] << 24)));
        stg8(out + px + px / 4 + co, make_uint2(cr[0] | (cr[1] << 8) | (cr[2] << 16) | (cr[3] << 24),
                                                cr[4] | (cr[5] << 8) | (cr[6] << 16) | (cr[7] << 24)));
    }
}

// each byte of an 8-byte word doubled -> 16 bytes
__device__ __forceinline__ uint4 dup8(const uint2 &v)
{
    return make_uint4(bp(v.x, 0u, 0x1100u), bp(v.x, 0u, 0x3322u), bp(v.y, 0u, 0x1100u), bp(v.y, 0u, 0x3322u));
}
// even bytes of 16 -> 8
__device__ __forceinline__ uint2 even8(const uint4 &v)
{
    return make_uint2(bp(v.x, v.y, 0x6420u), bp(v.z, v.w, 0x6420u));
}
// odd bytes 1,5,9,13 (sel = 1) or 3,7,11,15 (sel = 3) of two 16-byte words -> 8
__device__ __forceinline__ uint2 pick8(const uint4 &a, const uint4 &b, int sel)
{
    const uint32_t s = sel == 1 ? 0x5151u : 0x7373u;     // bytes {sel of lo, sel of hi} twice; the low half is used
    const uint32_t w0 = bp(a.x, a.y, s), w1 = bp(a.z, a.w, s), w2 = bp(b.x, b.y, s), w3 = bp(b.z, b.w, s);
    return make_uint2(bp(w0, w1, 0x5410u), bp(w2, w3, 0x5410u));
}

// kind 2, YVU9 -> YV12: work item = 32 luma pixels x 4 rows (+ 8 chroma samples of each plane -> 16 x 2 rows)
__global__ void __launch_bounds__(256, 4) k_yvu9_to_yv12_v(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h,
                                                        size_t in_frame, size_t out_frame)
{
    const size_t px = (size_t)w * h;
    in += (size_t)blockIdx.y * in_frame;
    out += (size_t)blockIdx.y * out_frame;
    const int upr = w >> 5, items = upr * (h >> 2), cw2 = w >> 1, cw4 = w >> 2, ch4 = h >> 2;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
        const int q = it / upr, u = it - q * upr;
        // every load of the item is issued before the first store (10 in flight per thread)
        uint4 v0[4], v1[4];
        uint2 c[2];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint8_t *src = in + (size_t)(h - 1 - (4 * q + r)) * w + 32 * u;      // planes are read bottom-up
            v0[r] = ldg16(src);
            v1[r] = ldg16(src + 16);
        }
#pragma unroll
        for (int pl = 0; pl < 2; pl++)
            c[pl] = ldg8(in + px + (size_t)pl * ((size_t)cw4 * ch4) + (size_t)(ch4 - 1 - q) * cw4 + 8 * u);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            uint8_t *dst = out + (size_t)(4 * q + r) * w + 32 * u;
            stg16(dst, v0[r]);
            stg16(dst + 16, v1[r]);
        }
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            const uint4 d = dup8(c[pl]);
            uint8_t *dst = out + px + (size_t)pl * (px / 4) + (size_t)(2 * q) * cw2 + 16 * u;
            stg16(dst, d);
            stg16(dst + cw2, d);
        }
    }
}

// kind 3, YUY2 -> YV12: work item = 16 pixels x one row pair
__global__ void __launch_bounds__(256) k_yuy2_to_yv12_v(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h,
                                                        size_t in_frame, size_t out_frame)
{
    const size_t px = (size_t)w * h;
    in += (size_t)blockIdx.y * in_frame;
    out += (size_t)blockIdx.y * out_frame;
    const int upr = w >> 4, items = upr * (h >> 1), cw2 = w >> 1;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
        const int rp = it / upr, u = it - rp * upr;
        const uint8_t *r0 = in + (size_t)(2 * rp) * 2 * w + 32 * u, *r1 = r0 + 2 * (size_t)w;
        const uint4 a0 = ldg16(r0), a1 = ldg16(r0 + 16), b0 = ldg16(r1), b1 = ldg16(r1 + 16);
        const uint2 ya0 = even8(a0), ya1 = even8(a1), yb0 = even8(b0), yb1 = even8(b1);
        uint8_t *y = out + (size_t)(2 * rp) * w + 16 * u;
        stg16(y, make_uint4(ya0.x, ya0.y, ya1.x, ya1.y));
        stg16(y + w, make_uint4(yb0.x, yb0.y, yb1.x, yb1.y));
        const size_t co = (size_t)rp * cw2 + 8 * u;                    // the odd row of the pair supplies the chroma
        stg8(out + px + co, pick8(b0, b1, 3));
        stg8(out + px + px / 4 + co, pick8(b0, b1, 1));
    }
}

// kind 4, YV12 -> YVU9: work item = 32 luma pixels x 4 rows (+ 16 chroma samples of an even row -> 8)
__global__ void __launch_bounds__(256, 4) k_yv12_to_yvu9_v(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h,
                                                        size_t in_frame, size_t out_frame)
{
    const size_t px = (size_t)w * h;
    in += (size_t)blockIdx.y * in_frame;
    out += (size_t)blockIdx.y * out_frame;
    const int upr = w >> 5, items = upr * (h >> 2), cw2 = w >> 1, cw4 = w >> 2, ch4 = h >> 2;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
        const int q = it / upr, u = it - q * upr;
        uint4 v0[4], v1[4], c[2];                 // loads first, then stores
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const size_t o = (size_t)(4 * q + r) * w + 32 * u;
            v0[r] = ldg16(in + o);
            v1[r] = ldg16(in + o + 16);
        }
#pragma unroll
        for (int pl = 0; pl < 2; pl++)
            c[pl] = ldg16(in + px + (size_t)pl * (px / 4) + (size_t)(2 * q) * cw2 + 16 * u);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const size_t o = (size_t)(4 * q + r) * w + 32 * u;
            stg16(out + o, v0[r]);
            stg16(out + o + 16, v1[r]);
        }
#pragma unroll
        for (int pl = 0; pl < 2; pl++)
            stg8(out + px + (size_t)pl * ((size_t)cw4 * ch4) + (size_t)q * cw4 + 8 * u, even8(c[pl]));
    }
}

// kind 5, YV12 -> YUY2: work item = 16 pixels x one row pair; word k of a row = Y(2k), 2nd plane, Y(2k+1), 1st plane
__global__ void __launch_bounds__(256) k_yv12_to_yuy2_v(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int w, int h,
                                                        size_t in_frame, size_t out_frame)
{
    const size_t px = (size_t)w * h;
    in += (size_t)blockIdx.y * in_frame;
    out += (size_t)blockIdx.y * out_frame;
    const int upr = w >> 4, items = upr * (h >> 1), cw2 = w >> 1;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
        const int rp = it / upr, u = it - rp * upr;
        const size_t co = (size_t)rp * cw2 + 8 * u;
        const uint2 c3 = ldg8(in + px + co), c1 = ldg8(in + px + px / 4 + co);
        // chroma words: byte 1 <- c1[k], byte 3 <- c3[k]
        uint32_t cw[8];
        cw[0] = bp(c1.x, c3.x, 0x4000u) & 0xff00ff00u; cw[1] = bp(c1.x, c3.x, 0x5010u) & 0xff00ff00u;
        cw[2] = bp(c1.x, c3.x, 0x6020u) & 0xff00ff00u; cw[3] = bp(c1.x, c3.x, 0x7030u) & 0xff00ff00u;
        cw[4] = bp(c1.y, c3.y, 0x4000u) & 0xff00ff00u; cw[5] = bp(c1.y, c3.y, 0x5010u) & 0xff00ff00u;
        cw[6] = bp(c1.y, c3.y, 0x6020u) & 0xff00ff00u; cw[7] = bp(c1.y, c3.y, 0x7030u) & 0xff00ff00u;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const uint4 yv = ldg16(in + (size_t)(2 * rp + r) * w + 16 * u);
            const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
            uint32_t o[8];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                o[2 * k] = bp(yw[k], 0u, 0x4140u) | cw[2 * k];            // Y0 . Y1 .
                o[2 * k + 1] = bp(yw[k], 0u, 0x4342u) | cw[2 * k + 1];
            }
            uint8_t *dst = out + (size_t)(2 * rp + r) * 2 * w + 32 * u;
            stg16(dst, make_uint4(o[0], o[1], o[2], o[3]));
            stg16(dst + 16, make_uint4(o[4], o[5], o[6], o[7]));
        }
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
        // the vector kernels move 16-byte words: both buffers must be 16-byte aligned, otherwise the byte kernels run
        const bool al16 = ((reinterpret_cast<uintptr_t>(i0) | reinterpret_cast<uintptr_t>(o0)) & 15) == 0;
        if (w % 32 == 0 && h % 4 == 0 && al16 && inb % 16 == 0 && outb % 16 == 0) {
            // vector kernels: one work item = 16 px x 2 rows (kinds 1, 3, 5) or 32 px x 4 rows (kinds 2, 4)
            const size_t items = (kind == 2 || kind == 4) ? (size_t)(w / 32) * (h / 4) : (size_t)(w / 16) * (h / 2);
            unsigned vx = (unsigned)((items + 255) / 256);
            if (vx < 1) vx = 1;
            const dim3 vgrid(vx, nf);
            switch (kind) {
            case 1: k_rgb24_to_yv12_v<<<vgrid, 256, 0, st>>>(i0, o0, w, h); break;
            case 2: k_yvu9_to_yv12_v<<<vgrid, 256, 0, st>>>(i0, o0, w, h, inb, outb); break;
            case 3: k_yuy2_to_yv12_v<<<vgrid, 256, 0, st>>>(i0, o0, w, h, inb, outb); break;
            case 4: k_yv12_to_yvu9_v<<<vgrid, 256, 0, st>>>(i0, o0, w, h, inb, outb); break;
            case 5: k_yv12_to_yuy2_v<<<vgrid, 256, 0, st>>>(i0, o0, w, h, inb, outb); break;
            default: return -(int)cudaErrorInvalidValue;
            }
        } else {
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
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return -(int)e;
        launches++;
    }
    return launches;
}

}  // namespace lv
