// lv_aux.cu -- the two steps either side of the fused path that the application displays / consumes
// (SURVEY.md 8(f) ranks 2 and 3), written for sm_100a byte SIMD.
//
// k_write_bks  : the literal CMainDlg::WriteBks (MainDlg.cpp:771-835; same primitive in lib/JM/image.c:4791-4816):
//                per sample of an I420 frame, IsNearlyZero(cur - bkd, tol) ? fill : cur with fill = 16 for luma and
//                128 for chroma, against ONE static background frame; optionally also emits bkd_obj (image.c:4798,
//                4803) as a luma bit mask.  Pure stream: 1.5 B/px in + 1.5 B/px out; the background frame (3 MB at
//                1080p) stays in L2.
// k_object_box : min/max x,y of the set mask pixels inside a window (image.c:4818-4838), from the byte mask or
//                the bit mask; empty => {x1+1, x0-1, y1+1, y0-1} exactly as the reference initialises them.
#include "lv_kernels.h"
#include "lv_pixel.cuh"

namespace lv {

namespace {

// bytes where |a-b| <= tol keep `fill`, others keep a.  All four bytes of a word are luma or all are chroma.
__device__ __forceinline__ uint32_t bks_word(uint32_t a, uint32_t b, uint32_t fill4, uint32_t tol4, uint32_t ntol7,
                                             uint32_t &changed_bits)
{
    const uint32_t gt = diff_gt4(a, b, tol4, ntol7);          // bit 7 of each byte: |a-b| > tol
    changed_bits = gt & 0x80808080u;
    uint32_t m;                                               // replicate bit 7 over each byte: 0xff = changed
    asm("prmt.b32 %0, %1, %2, 0xba98;" : "=r"(m) : "r"(gt), "r"(0u));
    return (a & m) | (fill4 & ~m);
}

// 16 bytes of one frame: words i..i+3 (word index inside the frame)
__device__ __forceinline__ uint4 bks16(const uint4 &cv, const uint4 &bv, size_t i, size_t ywords, uint32_t tol4, uint32_t ntol7,
                                       uint32_t &g16)
{
    uint32_t bits[4];
    uint4 o;
    o.x = bks_word(cv.x, bv.x, i + 0 < ywords ? 0x10101010u : 0x80808080u, tol4, ntol7, bits[0]);
    o.y = bks_word(cv.y, bv.y, i + 1 < ywords ? 0x10101010u : 0x80808080u, tol4, ntol7, bits[1]);
    o.z = bks_word(cv.z, bv.z, i + 2 < ywords ? 0x10101010u : 0x80808080u, tol4, ntol7, bits[2]);
    o.w = bks_word(cv.w, bv.w, i + 3 < ywords ? 0x10101010u : 0x80808080u, tol4, ntol7, bits[3]);
    // 16 luma pixels = one mask unit: gather bit 7 of every byte (see diff_mask16)
    g16 = (bits[0] * 0x00204081u >> 28) | ((bits[1] * 0x00204081u >> 28) << 4) | ((bits[2] * 0x00204081u >> 28) << 8) |
          ((bits[3] * 0x00204081u >> 28) << 12);
    return o;
}

// frame % 16 == 0 (any width % 8 == 0): a block owns kUnroll * 256 consecutive 16-byte words of one frame and every
// thread keeps kUnroll loads of the current frame (HBM) and of the background (L2) in flight before it computes.
template <int kUnroll>
__global__ void __launch_bounds__(256) k_write_bks_v(const uint8_t *__restrict__ cur, const uint8_t *__restrict__ bkd,
                                                     uint8_t *__restrict__ out, uint16_t *__restrict__ obj_bits, size_t px,
                                                     size_t frame, int tol, int units_ok)
{
    const uint32_t tol4 = (uint32_t)tol * 0x01010101u, ntol7 = ~tol4 & 0x7f7f7f7fu;
    const size_t vecs = frame / 16, ywords = px / 4;
    const int f = blockIdx.y;
    const uint4 *c = reinterpret_cast<const uint4 *>(cur + (size_t)f * frame);
    const uint4 *b = reinterpret_cast<const uint4 *>(bkd);
    uint4 *o = reinterpret_cast<uint4 *>(out + (size_t)f * frame);
    const size_t base = (size_t)blockIdx.x * (256 * kUnroll) + threadIdx.x;
    uint4 cv[kUnroll], bv[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; k++) {
        const size_t v = base + 256 * k;
        if (v < vecs) {
            cv[k] = __ldcs(c + v);
            bv[k] = __ldg(b + v);
        }
    }
#pragma unroll
    for (int k = 0; k < kUnroll; k++) {
        const size_t v = base + 256 * k;
        if (v < vecs) {
            uint32_t g;
            __stcs(o + v, bks16(cv[k], bv[k], 4 * v, ywords, tol4, ntol7, g));
            if (obj_bits && units_ok && 4 * v + 4 <= ywords) obj_bits[(size_t)f * (px / 16) + v] = (uint16_t)g;
        }
    }
}

__global__ void __launch_bounds__(256) k_write_bks(const uint8_t *__restrict__ cur, const uint8_t *__restrict__ bkd,
                                                   uint8_t *__restrict__ out, uint16_t *__restrict__ obj_bits,
                                                   size_t px, size_t frame, int n_frames, int tol, int units_ok)
{
    const uint32_t tol4 = (uint32_t)tol * 0x01010101u, ntol7 = ~tol4 & 0x7f7f7f7fu;
    const size_t words = frame / 4, ywords = px / 4;
    const int f = blockIdx.y;
    const uint32_t *c = reinterpret_cast<const uint32_t *>(cur + (size_t)f * frame);
    const uint32_t *b = reinterpret_cast<const uint32_t *>(bkd);
    uint32_t *o = reinterpret_cast<uint32_t *>(out + (size_t)f * frame);
    // generic geometry (frame % 16 != 0): 4 words per thread per iteration, word by word
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < words; i += (size_t)gridDim.x * blockDim.x * 4) {
        uint32_t cw[4], bw[4], ow[4], bits[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            cw[k] = i + k < words ? c[i + k] : 0u;
            bw[k] = i + k < words ? b[i + k] : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            ow[k] = bks_word(cw[k], bw[k], i + k < ywords ? 0x10101010u : 0x80808080u, tol4, ntol7, bits[k]);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (i + k < words) o[i + k] = ow[k];
        if (obj_bits && units_ok && i + 4 <= ywords) {
            const uint32_t g = (bits[0] * 0x00204081u >> 28) | ((bits[1] * 0x00204081u >> 28) << 4) |
                               ((bits[2] * 0x00204081u >> 28) << 8) | ((bits[3] * 0x00204081u >> 28) << 12);
            obj_bits[(size_t)f * (px / 16) + i / 4] = (uint16_t)g;
        }
    }
}

// Block-wide min/max, then at most one atomic per bound per block -- and only when it would move the bound (the four
// ints of 8 frames share a 128-byte line: per-warp atomics from every block serialise in one L2 slice and cost more
// than the whole scan).  The plain pre-read can only be stale towards "less tight", so a skipped atomic is always safe.
__device__ __forceinline__ void box_commit(int *box, int mnx, int mxx, int mny, int mxy)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    __shared__ int4 part[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) part[warp] = make_int4(mnx, mxx, mny, mxy);
    __syncthreads();
    if (warp == 0) {
        const int4 p = part[lane & 7];
        mnx = p.x; mxx = p.y; mny = p.z; mxy = p.w;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
            mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        }
        if (lane == 0 && mxx >= mnx) {
            const int4 cur = __ldcg(reinterpret_cast<const int4 *>(box));
            if (mnx < cur.x) atomicMin(box + 0, mnx);
            if (mxx > cur.y) atomicMax(box + 1, mxx);
            if (mny < cur.z) atomicMin(box + 2, mny);
            if (mxy > cur.w) atomicMax(box + 3, mxy);
        }
    }
}

// box[0..3] = {min_x, max_x, min_y, max_y} per frame; must be pre-initialised to {x1+1, x0-1, y1+1, y0-1}.
// Generic geometry (byte mask, w % 16 != 0): one pixel per thread iteration.
__global__ void __launch_bounds__(256) k_object_box_scalar(const uint8_t *__restrict__ mask, int w, size_t frame_stride, int x0,
                                                           int x1, int y0, int y1, int *boxes)
{
    mask += (size_t)blockIdx.y * frame_stride;
    int *box = boxes + 4 * (size_t)blockIdx.y;
    int mnx = x1 + 1, mxx = x0 - 1, mny = y1 + 1, mxy = y0 - 1;
    const int ww = x1 - x0 + 1;
    const int total = ww * (y1 - y0 + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = y0 + i / ww, x = x0 + i % ww;
        if (mask[(size_t)y * w + x] == 1) {
            mnx = min(mnx, x); mxx = max(mxx, x); mny = min(mny, y); mxy = max(mxy, y);
        }
    }
    box_commit(box, mnx, mxx, mny, mxy);
}

// bit 7 of every byte of the result = (that byte of v == 1) & clip7; exact per byte (no borrow between bytes)
__device__ __forceinline__ uint32_t eq1_bits(uint32_t v, uint32_t clip7)
{
    const uint32_t y = v ^ 0x01010101u;                        // zero byte <=> mask byte == 1 (image.c tests == 1)
    const uint32_t t = (y & 0x7f7f7f7fu) + 0x7f7f7f7fu;        // bit 7 set <=> low 7 bits non-zero
    return ~(t | y) & clip7;                                   // one LOP3
}

// per-byte clip mask (0x80 = inside) for the 4 pixels x..x+3 against [x0, x1]
__device__ __forceinline__ uint32_t clip7_of(int x, int x0, int x1)
{
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (x + k >= x0 && x + k <= x1) m |= 0x80u << (8 * k);
    return m;
}

// w % 16 == 0.  Item = one 16-byte load: 16 pixels of the byte mask (kBits = false) or 8 units = 128 pixels of the bit
// mask (kBits = true).  A thread keeps ONE column of items for all its rows: the per-item work is an OR into a 16-byte
// accumulator plus "any pixel set" for the row bounds; bit positions are looked at once per column, at the end.
// blockIdx.y = frame, blockIdx.x = a band of `rows_per_block` window rows; kBoxUnroll loads in flight per thread.
constexpr int kBoxUnroll = 4;

template <bool kBits>
__device__ __forceinline__ uint4 box_load(const uint8_t *mask, size_t row_bytes, int w, int y, int c)
{
    if (!kBits) {
        // volatile asm: ptxas otherwise re-uses the destination registers and serialises load -> use -> load
        uint4 v;
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(mask + (size_t)y * row_bytes + (size_t)c * 16));
        return v;
    }
    // a row of the bit mask is w/16 uint16 units = w/8 bytes: rows need not be 16-byte aligned and the last item of a
    // row may be short
    const uint8_t *p = mask + (size_t)y * row_bytes + (size_t)c * 16;
    if (c * 128 + 128 <= w && (reinterpret_cast<uintptr_t>(p) & 15) == 0) return __ldg(reinterpret_cast<const uint4 *>(p));
    const uint16_t *pu = reinterpret_cast<const uint16_t *>(p);
    const int nu = min(8, (w - c * 128) >> 4);
    uint32_t wd[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (k < nu) wd[k >> 1] |= (uint32_t)__ldg(pu + k) << (16 * (k & 1));
    return make_uint4(wd[0], wd[1], wd[2], wd[3]);
}

template <bool kBits>
__global__ void __launch_bounds__(256, 2) k_object_box(const uint8_t *__restrict__ mask, int w, size_t frame_stride, int x0, int x1,
                                                    int y0, int y1, int rows_per_block, int *boxes)
{
    mask += (size_t)blockIdx.y * frame_stride;
    int *box = boxes + 4 * (size_t)blockIdx.y;
    int mnx = x1 + 1, mxx = x0 - 1, mny = y1 + 1, mxy = y0 - 1;
    constexpr int kPx = kBits ? 128 : 16;                 // pixels per item
    const int c0 = x0 / kPx, cols = x1 / kPx - c0 + 1;
    const int cpb = min(cols, 256), rpp = 256 / cpb;      // columns side by side in the block, rows per pass
    const int j = threadIdx.x % cpb, q = threadIdx.x / cpb;
    const size_t row_bytes = kBits ? (size_t)(w >> 3) : (size_t)w;
    const int rb0 = y0 + blockIdx.x * rows_per_block, rb1 = min(y1 + 1, rb0 + rows_per_block);
    if (q < rpp) {
        for (int c = c0 + j; c < c0 + cols; c += cpb) {
            const int xb = c * kPx;
            uint32_t clip[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (kBits) {
                    const int xw = xb + 32 * k;
                    uint32_t m = 0xffffffffu;
                    if (xw < x0) m = x0 - xw >= 32 ? 0u : m & (0xffffffffu << (x0 - xw));
                    if (xw + 31 > x1) m = x1 < xw ? 0u : m & (0xffffffffu >> (xw + 31 - x1));
                    clip[k] = m;
                } else {
                    clip[k] = clip7_of(xb + 4 * k, x0, x1);
                }
            }
            uint4 acc = make_uint4(0u, 0u, 0u, 0u);
            for (int y = rb0 + q; y < rb1; y += kBoxUnroll * rpp) {
                uint4 v[kBoxUnroll];
                int yc[kBoxUnroll];
#pragma unroll
                for (int k = 0; k < kBoxUnroll; k++)      // unconditional (row clamped) so that all loads are issued before any use
                    yc[k] = min(y + k * rpp, rb1 - 1);
#pragma unroll
                for (int k = 0; k < kBoxUnroll; k++)
                    v[k] = box_load<kBits>(mask, row_bytes, w, yc[k], c);
#pragma unroll
                for (int k = 0; k < kBoxUnroll; k++) {
                    uint4 m;
                    if (kBits) {
                        m = make_uint4(v[k].x & clip[0], v[k].y & clip[1], v[k].z & clip[2], v[k].w & clip[3]);
                    } else {
                        m = make_uint4(eq1_bits(v[k].x, clip[0]), eq1_bits(v[k].y, clip[1]), eq1_bits(v[k].z, clip[2]),
                                       eq1_bits(v[k].w, clip[3]));
                    }
                    if ((m.x | m.y | m.z | m.w) != 0u) {
                        mny = min(mny, yc[k]);                  // a clamped duplicate of the band's last row changes nothing
                        mxy = max(mxy, yc[k]);
                    }
                    acc.x |= m.x; acc.y |= m.y; acc.z |= m.z; acc.w |= m.w;
                }
            }
            // column bounds from the accumulated item
            const uint32_t aw[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (aw[k]) {
                    if (kBits) {
                        mnx = min(mnx, xb + 32 * k + __ffs(aw[k]) - 1);
                        mxx = max(mxx, xb + 32 * k + 31 - __clz(aw[k]));
                    } else {      // bit 7 of byte b = pixel xb + 4k + b
                        mnx = min(mnx, xb + 4 * k + ((__ffs(aw[k]) - 1) >> 3));
                        mxx = max(mxx, xb + 4 * k + ((31 - __clz(aw[k])) >> 3));
                    }
                }
            }
        }
    }
    box_commit(box, mnx, mxx, mny, mxy);
}

__global__ void k_box_init(int *boxes, int n, int x0, int x1, int y0, int y1)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n) reinterpret_cast<int4 *>(boxes)[f] = make_int4(x1 + 1, x0 - 1, y1 + 1, y0 - 1);
}

// ---- decoder picture (unsigned-short planes) -> tight 8-bit I420: img2buf + write_out_picture on the GPU ----------------
// lib/JM/output.c:87-178: with imgpel == unsigned short (global.h:61) and 1-byte output symbols the reference copies ONE
// byte of every little-endian sample -- its low byte, no clipping -- and cuts the crop rectangle away, sample by sample;
// write_out_picture (output.c:362-453) does that for Y, then imgUV[0], then imgUV[1], the chroma planes with half the
// luma crop.  3 B/px in (16-bit 4:2:0) + 1.5 B/px out.
struct NarrowArgs {
    const uint16_t *pic16;
    uint8_t *i420;
    int size_x, size_y, crop_left, crop_top, w, h;
    size_t pic_samples, frame;       // strides between pictures (samples) and between output frames (bytes)
};

// item = 16 output bytes of one plane row; kNarrowUnroll items (2 x 16-byte loads each) in flight per thread
constexpr int kNarrowUnroll = 4;

__device__ __forceinline__ void narrow_addr(const NarrowArgs &a, int i, int ny, int nc, const uint16_t *pic, uint8_t *out,
                                            const uint16_t *&src, uint8_t *&dst)
{
    if (i < ny) {
        const int per = a.w >> 4, row = i / per, col = i - row * per;
        src = pic + (size_t)(row + a.crop_top) * a.size_x + a.crop_left + 16 * col;
        dst = out + (size_t)row * a.w + 16 * col;
    } else {
        int j = i - ny;
        const int pl = j >= nc;
        j -= pl * nc;
        const int cw = a.w >> 1, cx = a.size_x >> 1, cy = a.size_y >> 1;
        const int per = cw >> 4, row = j / per, col = j - row * per;
        src = pic + (size_t)a.size_x * a.size_y + (size_t)pl * cx * cy + (size_t)(row + (a.crop_top >> 1)) * cx +
              (a.crop_left >> 1) + 16 * col;
        dst = out + (size_t)a.w * a.h + (size_t)pl * cw * (a.h >> 1) + (size_t)row * cw + 16 * col;
    }
}

__global__ void __launch_bounds__(256, 2) k_narrow_v(const NarrowArgs a)
{
    const uint16_t *pic = a.pic16 + (size_t)blockIdx.y * a.pic_samples;
    uint8_t *out = a.i420 + (size_t)blockIdx.y * a.frame;
    const int ny = (a.w >> 4) * a.h, nc = (a.w >> 5) * (a.h >> 1), total = ny + 2 * nc;
    const int base = blockIdx.x * (256 * kNarrowUnroll) + threadIdx.x;
    uint4 lo[kNarrowUnroll], hi[kNarrowUnroll];
    uint8_t *dst[kNarrowUnroll];
#pragma unroll
    for (int k = 0; k < kNarrowUnroll; k++) {
        const int i = min(base + 256 * k, total - 1);        // clamped: every load is issued, the store is guarded
        const uint16_t *src;
        narrow_addr(a, i, ny, nc, pic, out, src, dst[k]);
        lo[k] = __ldcs(reinterpret_cast<const uint4 *>(src));
        hi[k] = __ldcs(reinterpret_cast<const uint4 *>(src) + 1);
    }
#pragma unroll
    for (int k = 0; k < kNarrowUnroll; k++) {
        if (base + 256 * k < total) {
            // low byte of every little-endian short: bytes 0, 2 of each word
            const uint4 o = make_uint4(__byte_perm(lo[k].x, lo[k].y, 0x6420u), __byte_perm(lo[k].z, lo[k].w, 0x6420u),
                                       __byte_perm(hi[k].x, hi[k].y, 0x6420u), __byte_perm(hi[k].z, hi[k].w, 0x6420u));
            __stcs(reinterpret_cast<uint4 *>(dst[k]), o);
        }
    }
}

// any legal geometry: one thread per output byte
__global__ void __launch_bounds__(256) k_narrow_scalar(const NarrowArgs a)
{
    const uint16_t *pic = a.pic16 + (size_t)blockIdx.y * a.pic_samples;
    uint8_t *out = a.i420 + (size_t)blockIdx.y * a.frame;
    const size_t px = (size_t)a.w * a.h, cpx = (size_t)(a.w >> 1) * (a.h >> 1);
    const int cx = a.size_x >> 1, cy = a.size_y >> 1, cw = a.w >> 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.frame; i += (size_t)gridDim.x * blockDim.x) {
        uint16_t v;
        if (i < px) {
            const int row = (int)(i / a.w), col = (int)(i - (size_t)row * a.w);
            v = pic[(size_t)(row + a.crop_top) * a.size_x + a.crop_left + col];
        } else {
            size_t j = i - px;
            const int pl = j >= cpx;
            j -= pl * cpx;
            const int row = (int)(j / cw), col = (int)(j - (size_t)row * cw);
            v = pic[(size_t)a.size_x * a.size_y + (size_t)pl * cx * cy + (size_t)(row + (a.crop_top >> 1)) * cx +
                    (a.crop_left >> 1) + col];
        }
        out[i] = (uint8_t)(v & 0xffu);
    }
}

}  // namespace

int launch_write_bks(const Geom &g, const uint8_t *cur, const uint8_t *bkd, int n_frames, int tol, uint8_t *out,
                     uint16_t *obj_bits, cudaStream_t st)
{
    if (n_frames <= 0) return 0;
    int launches = 0;
    for (int f0 = 0; f0 < n_frames; f0 += 65535) {
        const int nf = n_frames - f0 < 65535 ? n_frames - f0 : 65535;
        const uint8_t *c0 = cur + (size_t)f0 * g.frame;
        uint8_t *o0 = out + (size_t)f0 * g.frame;
        uint16_t *b0 = obj_bits ? obj_bits + (size_t)f0 * (g.px / 16) : nullptr;
        if (g.frame % 16 == 0 && ((reinterpret_cast<uintptr_t>(c0) | reinterpret_cast<uintptr_t>(bkd) | reinterpret_cast<uintptr_t>(o0)) & 15) == 0) {
            constexpr int kUnroll = 4;
            const size_t vecs = g.frame / 16;
            const unsigned bx = (unsigned)((vecs + 256 * kUnroll - 1) / (256 * kUnroll));
            k_write_bks_v<kUnroll><<<dim3(bx, nf), 256, 0, st>>>(c0, bkd, o0, b0, g.px, g.frame, tol, g.w % 16 == 0);
        } else {
            const size_t words = g.frame / 4;
            unsigned bx = (unsigned)((words / 4 + 255) / 256);
            if (bx < 1) bx = 1;
            if (bx > 8192) bx = 8192;
            k_write_bks<<<dim3(bx, nf), 256, 0, st>>>(c0, bkd, o0, b0, g.px, g.frame, nf, tol, g.w % 16 == 0);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return -(int)e;
        launches++;
    }
    return launches;
}

int launch_object_box(int w, int h, const uint8_t *mask, bool bits, int n_frames, int x0, int x1, int y0, int y1, int *d_boxes,
                      cudaStream_t st)
{
    if (n_frames <= 0) return 0;
    int launches = 0;
    const size_t stride = bits ? (size_t)(w >> 3) * h : (size_t)w * h;
    for (int f0 = 0; f0 < n_frames; f0 += 65535) {
        const int nf = n_frames - f0 < 65535 ? n_frames - f0 : 65535;
        int *boxes = d_boxes + 4 * (size_t)f0;
        const uint8_t *m = mask + (size_t)f0 * stride;
        k_box_init<<<(nf + 255) / 256, 256, 0, st>>>(boxes, nf, x0, x1, y0, y1);
        launches++;
        const long long rows = (long long)y1 - y0 + 1;
        if (x1 >= x0 && rows > 0) {
            if (w % 16 != 0) {
                const long long total = (long long)(x1 - x0 + 1) * rows;
                long long bx = (total + 1023) / 1024;
                const long long cap = (8LL * 148 * 8 + nf - 1) / nf;
                if (bx > cap) bx = cap;
                if (bx < 1) bx = 1;
                k_object_box_scalar<<<dim3((unsigned)bx, nf), 256, 0, st>>>(m, w, stride, x0, x1, y0, y1, boxes);
            } else {
                const int per = bits ? 128 : 16;
                const int cols = x1 / per - x0 / per + 1, cpb = cols < 256 ? cols : 256, rpp = 256 / cpb;
                // One wave: at most 148 SMs x 8 resident blocks in the whole launch, so every block is loading from the
                // first microsecond and its fixed costs (set-up, reduction, the bound pre-read and atomics) are paid once
                // per many passes; a band is at least one pass of kBoxUnroll x rpp rows.
                long long bands = (rows + 4LL * rpp - 1) / (4LL * rpp);
                const long long cap = (148LL * 8) / nf > 0 ? (148LL * 8) / nf : 1;
                if (bands > cap) bands = cap;
                const long long rpb = (rows + bands - 1) / bands;
                const unsigned bx = (unsigned)((rows + rpb - 1) / rpb);
                if (bits) k_object_box<true><<<dim3(bx, nf), 256, 0, st>>>(m, w, stride, x0, x1, y0, y1, (int)rpb, boxes);
                else k_object_box<false><<<dim3(bx, nf), 256, 0, st>>>(m, w, stride, x0, x1, y0, y1, (int)rpb, boxes);
            }
            launches++;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return -(int)e;
    }
    return launches;
}

int launch_narrow(const Geom &g, const uint16_t *pic16, int n_frames, int size_x, int size_y, int crop_left, int crop_top,
                  uint8_t *i420, cudaStream_t st)
{
    if (n_frames <= 0) return 0;
    NarrowArgs a{};
    a.size_x = size_x; a.size_y = size_y; a.crop_left = crop_left; a.crop_top = crop_top; a.w = g.w; a.h = g.h;
    a.pic_samples = (size_t)size_x * size_y + 2 * ((size_t)(size_x >> 1) * (size_y >> 1));
    a.frame = g.frame;
    const bool vec = g.w % 32 == 0 && size_x % 16 == 0 && crop_left % 16 == 0 && ((size_t)size_x * size_y) % 8 == 0 &&
                     ((size_t)(size_x >> 1) * (size_y >> 1)) % 8 == 0 && g.frame % 16 == 0 &&
                     ((reinterpret_cast<uintptr_t>(pic16) | reinterpret_cast<uintptr_t>(i420)) & 15) == 0;
    int launches = 0;
    for (int f0 = 0; f0 < n_frames; f0 += 65535) {
        const int nf = n_frames - f0 < 65535 ? n_frames - f0 : 65535;
        a.pic16 = pic16 + (size_t)f0 * a.pic_samples;
        a.i420 = i420 + (size_t)f0 * g.frame;
        if (vec) {
            const int total = (g.w >> 4) * g.h + 2 * ((g.w >> 5) * (g.h >> 1));
            k_narrow_v<<<dim3((total + 256 * kNarrowUnroll - 1) / (256 * kNarrowUnroll), nf), 256, 0, st>>>(a);
        } else {
            unsigned bx = (unsigned)((g.frame + 255) / 256);
            if (bx > 4096) bx = 4096;
            k_narrow_scalar<<<dim3(bx, nf), 256, 0, st>>>(a);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return -(int)e;
        launches++;
    }
    return launches;
}

// ---- object bookkeeping either side of the box ---------------------------------------------------------------------
namespace {
// lib/JM/image.h:23-24
__device__ __forceinline__ int clip_x(int x, int w) { return x < 0 ? 0 : (x >= w ? w - 1 : x); }

// window of an object: its last position +- half its size +- the decoding margin, clipped to the picture
// (lib/JM/image.c:4786-4789; HOR/VER_DECODING_MARGIN = 8, lib/JM/global.h:119-120)
__global__ void k_object_windows(const int4 *obj, int n, int w, int h, int hm, int vm, int4 *win)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 o = obj[i];      // PosX, PosY, Width, Height
    win[i] = make_int4(clip_x(o.x - o.z / 2 - hm, w), clip_x(o.x + o.z / 2 + hm, w), clip_x(o.y - o.w / 2 - vm, h),
                       clip_x(o.y + o.w / 2 + vm, h));
}

// position / size of the object from its box, only when the box is not degenerate (lib/JM/image.c:4840-4848)
__global__ void k_object_update(const int4 *box, int n, int4 *obj)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 b = box[i];      // min_x, max_x, min_y, max_y
    if (b.x < b.y && b.z < b.w) obj[i] = make_int4(b.x + (b.y - b.x) / 2, b.z + (b.w - b.z) / 2, b.y - b.x, b.w - b.z);
}
}  // namespace

int launch_object_windows(int w, int h, const int *objects, int n, int hor_margin, int ver_margin, int *windows, cudaStream_t st)
{
    k_object_windows<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const int4 *>(objects), n, w, h, hor_margin, ver_margin,
                                                       reinterpret_cast<int4 *>(windows));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

int launch_object_update(const int *boxes, int n, int *objects, cudaStream_t st)
{
    k_object_update<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const int4 *>(boxes), n, reinterpret_cast<int4 *>(objects));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

}  // namespace lv
