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
    // 4 words (16 bytes) per thread per iteration; the tail of a frame that is not a multiple of 16 bytes is
    // handled word by word by the same loop (vector loads only where all four words exist)
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < words; i += (size_t)gridDim.x * blockDim.x * 4) {
        uint32_t cw[4], bw[4], ow[4], bits[4];
        const bool full = i + 4 <= words && (frame % 16 == 0);
        if (full) {
            const uint4 cv = __ldcs(reinterpret_cast<const uint4 *>(c + i));
            const uint4 bv = __ldg(reinterpret_cast<const uint4 *>(b + i));
            cw[0] = cv.x; cw[1] = cv.y; cw[2] = cv.z; cw[3] = cv.w;
            bw[0] = bv.x; bw[1] = bv.y; bw[2] = bv.z; bw[3] = bv.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                cw[k] = i + k < words ? c[i + k] : 0u;
                bw[k] = i + k < words ? b[i + k] : 0u;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            ow[k] = bks_word(cw[k], bw[k], i + k < ywords ? 0x10101010u : 0x80808080u, tol4, ntol7, bits[k]);
        if (full) {
            __stcs(reinterpret_cast<uint4 *>(o + i), make_uint4(ow[0], ow[1], ow[2], ow[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (i + k < words) o[i + k] = ow[k];
        }
        if (obj_bits && units_ok && i + 4 <= ywords) {
            // 16 luma pixels = one mask unit: gather bit 7 of every byte (see diff_mask16)
            const uint32_t g = (bits[0] * 0x00204081u >> 28) | ((bits[1] * 0x00204081u >> 28) << 4) |
                               ((bits[2] * 0x00204081u >> 28) << 8) | ((bits[3] * 0x00204081u >> 28) << 12);
            obj_bits[(size_t)f * (px / 16) + i / 4] = (uint16_t)g;
        }
    }
}

// box[0..3] = {min_x, max_x, min_y, max_y}; must be pre-initialised to {x1+1, x0-1, y1+1, y0-1}
template <bool kBits>
__global__ void __launch_bounds__(256) k_object_box(const uint8_t *__restrict__ mask, int w, int x0, int x1, int y0,
                                                    int y1, int *box)
{
    int mnx = x1 + 1, mxx = x0 - 1, mny = y1 + 1, mxy = y0 - 1;
    const int ww = x1 - x0 + 1;
    const long long total = (long long)ww * (y1 - y0 + 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = y0 + (int)(i / ww), x = x0 + (int)(i % ww);
        int on;
        if (kBits) {
            const uint16_t u = reinterpret_cast<const uint16_t *>(mask)[(size_t)y * (w / 16) + (x >> 4)];
            on = (u >> (x & 15)) & 1;
        } else {
            on = mask[(size_t)y * w + x] == 1;
        }
        if (on) {
            mnx = min(mnx, x); mxx = max(mxx, x); mny = min(mny, y); mxy = max(mxy, y);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(box + 0, mnx); atomicMax(box + 1, mxx); atomicMin(box + 2, mny); atomicMax(box + 3, mxy);
    }
}

__global__ void k_box_init(int *box, int x0, int x1, int y0, int y1)
{
    box[0] = x1 + 1; box[1] = x0 - 1; box[2] = y1 + 1; box[3] = y0 - 1;
}

}  // namespace

int launch_write_bks(const Geom &g, const uint8_t *cur, const uint8_t *bkd, int n_frames, int tol, uint8_t *out,
                     uint16_t *obj_bits, cudaStream_t st)
{
    if (n_frames <= 0) return 0;
    int launches = 0;
    for (int f0 = 0; f0 < n_frames; f0 += 65535) {
        const int nf = n_frames - f0 < 65535 ? n_frames - f0 : 65535;
        const size_t words = g.frame / 4;
        unsigned bx = (unsigned)((words / 4 + 255) / 256);
        if (bx < 1) bx = 1;
        if (bx > 8192) bx = 8192;
        k_write_bks<<<dim3(bx, nf), 256, 0, st>>>(cur + (size_t)f0 * g.frame, bkd, out + (size_t)f0 * g.frame,
                                                    obj_bits ? obj_bits + (size_t)f0 * (g.px / 16) : nullptr, g.px, g.frame,
                                                    nf, tol, g.w % 16 == 0);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return -(int)e;
        launches++;
    }
    return launches;
}

int launch_object_box(int w, const uint8_t *mask, bool bits, int x0, int x1, int y0, int y1, int *d_box, cudaStream_t st)
{
    k_box_init<<<1, 1, 0, st>>>(d_box, x0, x1, y0, y1);
    const long long total = (long long)(x1 - x0 + 1) * (y1 - y0 + 1);
    if (total > 0) {
        unsigned bx = (unsigned)((total + 255) / 256);
        if (bx > 1184) bx = 1184;
        if (bits) k_object_box<true><<<bx, 256, 0, st>>>(mask, w, x0, x1, y0, y1, d_box);
        else k_object_box<false><<<bx, 256, 0, st>>>(mask, w, x0, x1, y0, y1, d_box);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? (total > 0 ? 2 : 1) : -(int)e;
}

}  // namespace lv
