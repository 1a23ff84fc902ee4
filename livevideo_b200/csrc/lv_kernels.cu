// lv_kernels.cu -- sm_100a kernels of the fused YUV420 -> BGR24 + frame-difference hot path.
//
// K1 k_tile<true,true>   fused conversion + difference + per-macroblock counts/map (+ masks, summary)
// K2 k_tile<true,false>  conversion only  (ColorSpaceConversions::YV12_to_RGB24, Convert.h:26)
// K3 k_tile<false,true>  difference only  (IsNearlyZero loops, lib/JM/image.c:4791-4804)
//
// Work decomposition (all three): one thread owns a 16-pixel x 2-row patch, i.e. one 16-byte luma
// load per row, one 8-byte load of U and of V (the 2x2 chroma replication of the reference stays in
// registers), one 16-byte load per row of the reference luma, and 3 x 16-byte BGR stores per row.
// A block is 32 x 8 such threads = 512 pixels x 16 rows = 32 macroblocks of one macroblock row
// (k = mby*mb_w + mbx, lib/JM/image.c:4981); the 8 row-pair partial counts of a macroblock are
// summed through shared memory, so mb_count/mb_map need no atomics.  grid = (x tiles, mb rows,
// frames) with the frame index slowest, so frame t-1's luma (read as "current" a moment ago) is
// still in L2 when frame t reads it as the reference.
#include "lv_kernels.h"
#include "lv_pixel.cuh"

#include <cstdlib>

#ifndef LV_TILE_MINBLOCKS
#define LV_TILE_MINBLOCKS 5
#endif
#ifndef LV_TILE_FLAT
#define LV_TILE_FLAT 1
#endif
#ifndef LV_DEFAULT_TMA
#define LV_DEFAULT_TMA 0
#endif

namespace lv {

namespace {

struct KArgs {
    Geom g;
    const uint8_t *y0;          // luma of frame 0
    size_t y_stride;            // bytes between consecutive frames' luma (frame for I420, px for planes)
    const uint8_t *ref_planes;  // K3: explicit reference planes (stride px); else NULL
    const uint8_t *state_in;    // K1: carried previous luma, one plane per stream
    uint8_t *state_out;
    int n_frames;               // frames per stream
    uint32_t nf_magic;          // ceil(2^32 / n_frames): f / n_frames == umulhi(f, magic) for f < 2^16
    uint32_t mbw_magic;         // flat macroblock mapping: ceil(2^32 / mb_w), or 0 = one block row per macroblock row
    int f_base;                 // first frame index of this launch (gridDim.z is limited to 65535)
    int static_ref;             // 1: the stream's stored plane is the reference of EVERY frame (static background)
    int tol, mb_thresh;
    uint8_t *bgr;
    uint16_t *mask_bits;
    uint8_t *mask_bytes;
    uint16_t *mb_count;
    uint8_t *mb_map;
    uint32_t *summary;
    // kBox instantiation only: per-frame object windows {x0,x1,y0,y1} (inclusive) and boxes {min_x,max_x,min_y,max_y}
    const int4 *windows;
    int4 *boxes;
    int n_obj;
    // kPic instantiation only: the frames are the decoder's 16-bit pictures (imgpel = unsigned short, lib/JM/global.h:61;
    // planes Y, U, V of size_x x size_y / quarter size, memalloc.c:26-39) and the kernel does what img2buf +
    // write_out_picture do on the way (lib/JM/output.c:87-178, 362-453): low byte of every sample, crop cut away
    const uint16_t *pic0;       // first picture
    size_t pic_stride;          // samples between consecutive pictures
    int size_x;                 // luma pitch in samples (chroma: size_x / 2)
    uint32_t pic_y_off;         // crop_top * size_x + crop_left
    uint32_t pic_c_off;         // size_x * size_y + (crop_top / 2) * (size_x / 2) + crop_left / 2
    uint32_t pic_c_plane;       // (size_x / 2) * (size_y / 2): from U to V
};

// Cache policies (A/B measured on B200, profiles/README.md).  Loads: 0 = read-only path (ld.global.nc), 1 = .cs
// streaming, 2 = nc + L1::no_allocate, 3 = nc + L2 evict_first hint, 4 = nc + L1::no_allocate + L2 evict_first hint,
// 5 = nc + L2 evict_last hint.  Stores: 0 = .cs, 1 = default, 4 = L1::no_allocate + L2 evict_first hint.
#ifndef LV_LDY
#define LV_LDY 0
#endif
#ifndef LV_LDP
#define LV_LDP 0
#endif
#ifndef LV_LDC
#define LV_LDC 0
#endif
#ifndef LV_ST_POLICY
#define LV_ST_POLICY 0
#endif

template <int POL>
__device__ __forceinline__ uint4 ld16_pol(const uint8_t *p)
{
    uint4 v;
    if (POL == 0) return __ldg(reinterpret_cast<const uint4 *>(p));
    if (POL == 1) return __ldcs(reinterpret_cast<const uint4 *>(p));
    if (POL == 2) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        return v;
    }
    uint64_t pol;
    if (POL == 5) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (POL == 4)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    else
        asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
template <int POL>
__device__ __forceinline__ uint2 ld8_pol(const uint8_t *p)
{
    uint2 v;
    if (POL == 0) return __ldg(reinterpret_cast<const uint2 *>(p));
    if (POL == 1) return __ldcs(reinterpret_cast<const uint2 *>(p));
    if (POL == 2) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
        return v;
    }
    uint64_t pol;
    if (POL == 5) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (POL == 4)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    else
        asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint4 ld16(const uint8_t *p) { return ld16_pol<LV_LDY>(p); }
__device__ __forceinline__ uint4 ld16_stream(const uint8_t *p) { return ld16_pol<LV_LDP>(p); }
__device__ __forceinline__ uint2 ld8_stream(const uint8_t *p) { return ld8_pol<LV_LDC>(p); }
__device__ __forceinline__ void st16_stream(uint8_t *p, const uint4 &v)
{
#if LV_ST_POLICY == 0
    __stcs(reinterpret_cast<uint4 *>(p), v);
#elif LV_ST_POLICY == 1
    *reinterpret_cast<uint4 *>(p) = v;
#else
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
#endif
}

// 16 (8) consecutive 16-bit samples -> their low bytes (img2buf keeps the first byte of every little-endian sample)
__device__ __forceinline__ uint4 narrow16(const uint16_t *p)
{
    const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(p)), hi = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
    return make_uint4(prmt(lo.x, lo.y, 0x6420u), prmt(lo.z, lo.w, 0x6420u), prmt(hi.x, hi.y, 0x6420u), prmt(hi.z, hi.w, 0x6420u));
}
__device__ __forceinline__ uint2 narrow8(const uint16_t *p)
{
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    return make_uint2(prmt(v.x, v.y, 0x6420u), prmt(v.z, v.w, 0x6420u));
}

// 16 pixels of one row: luma words yw[4], chroma samples c[8] -> 48 bytes
__device__ __forceinline__ void convert_row16(const uint4 &y, const Chroma2 (&c)[8], uint8_t *dst)
{
    uint4 o0, o1, o2;
    convert4(y.x, c[0], c[1], o0.x, o0.y, o0.z);
    convert4(y.y, c[2], c[3], o0.w, o1.x, o1.y);
    convert4(y.z, c[4], c[5], o1.z, o1.w, o2.x);
    convert4(y.w, c[6], c[7], o2.y, o2.z, o2.w);
    st16_stream(dst, o0);
    st16_stream(dst + 16, o1);
    st16_stream(dst + 32, o2);
}

// All addressing is a block-uniform 64-bit base (frame / plane starts) plus one 32-bit per-thread offset;
// the frame -> (stream, t) split is a multiply-high by a host-computed reciprocal.  The optional mask
// outputs live in their own instantiation (kMask) so the common path carries none of their code; so does the object
// box of the step after the path (kBox: lib/JM/image.c:4818-4838 -- min / max x, y of the changed pixels inside each
// object's window, accumulated here so that no mask ever goes through HBM).
constexpr int kMaxObjects = 16;

//
// kEarly (plain fused instantiation only): the chroma loads are issued right behind the luma loads, in front of the
// reference-luma loads, instead of after the difference.  Same instructions, same bytes, different order of the six loads
// -- and this kernel's speed follows that order: see launch_fused_plain() for which pictures take which.
template <bool kConv, bool kDiff, bool kMask, bool kBox = false, bool kPic = false, bool kEarly = false>
__global__ void __launch_bounds__(256, kPic ? 4 : LV_TILE_MINBLOCKS) k_tile(const KArgs a)
{
    const Geom &g = a.g;
    const int tx = threadIdx.x, ty = threadIdx.y;
    // Flat mapping (mbw_magic != 0): a block owns 32 CONSECUTIVE macroblocks in raster order, so every lane has work
    // whatever the width (1080p: 120 macroblocks per row = 3.75 blocks of 32; 720p: 2.5; CIF: 0.69) -- a run simply
    // continues on the next macroblock row.  Otherwise blockIdx.y is the macroblock row and blockIdx.x tiles it.
    int unit, mby;
    if (a.mbw_magic) {
        const int m = blockIdx.x * 32 + tx;
        mby = (int)__umulhi((uint32_t)m, a.mbw_magic);
        unit = m - mby * g.mb_w;
    } else {
        unit = blockIdx.x * 32 + tx;
        mby = blockIdx.y;
    }
    const int row0 = mby * 16 + 2 * ty;
    const int f = a.f_base + blockIdx.z;
    const bool active = unit < g.units && row0 < g.h;
    const int w = g.w;

    uint32_t cnt = 0;
    [[maybe_unused]] uint32_t bm0 = 0, bm1 = 0;      // kBox: the thread's two 16-pixel mask rows
    // kBox: lane k of the block's second warp fetches object k's window NOW, with the block's first loads, so that the
    // box work behind the barrier starts from registers (a dependent global load there would keep the whole block's
    // registers and shared memory allocated for another memory latency with one warp alive)
    [[maybe_unused]] int4 wpre = make_int4(0, -1, 0, -1);
    if constexpr (kBox) {
        if (ty == 1 && tx < a.n_obj) wpre = __ldg(a.windows + (size_t)f * a.n_obj + tx);
    }
    if (active) {
        const uint8_t *yf = a.y0 + (size_t)f * a.y_stride;
        const uint32_t off = (uint32_t)row0 * (uint32_t)w + (uint32_t)unit * 16u;
        [[maybe_unused]] const uint16_t *pf = nullptr;
        [[maybe_unused]] uint32_t poff = 0;
        uint4 y0, y1;
        if constexpr (kPic) {
            pf = a.pic0 + (size_t)f * a.pic_stride;
            poff = a.pic_y_off + (uint32_t)row0 * (uint32_t)a.size_x + (uint32_t)unit * 16u;
            y0 = narrow16(pf + poff);
            y1 = narrow16(pf + poff + a.size_x);
        } else {
            y0 = ld16(yf + off);                  // re-read by the next frame as its reference: keep in L2
            y1 = ld16(yf + off + w);
        }
        [[maybe_unused]] uint2 u_early = make_uint2(0, 0), v_early = make_uint2(0, 0);
        if constexpr (kEarly && kConv) {
            if constexpr (kPic) {
                const uint32_t pcoff = a.pic_c_off + (uint32_t)(row0 >> 1) * (uint32_t)(a.size_x >> 1) + (uint32_t)unit * 8u;
                u_early = narrow8(pf + pcoff);
                v_early = narrow8(pf + pcoff + a.pic_c_plane);
            } else {
                const uint32_t coff = (uint32_t)(row0 >> 1) * (uint32_t)(w >> 1) + (uint32_t)unit * 8u;
                u_early = ld8_stream(yf + g.px + coff);
                v_early = ld8_stream(yf + g.px + (g.px >> 2) + coff);
            }
        }

        if constexpr (kDiff) {
            const int s = a.n_frames == 1 ? f : (int)__umulhi((uint32_t)f, a.nf_magic);
            const int t = f - s * a.n_frames;
            uint4 p0, p1;
            if (kPic && !(t == 0 || a.static_ref)) {
                // the previous PICTURE is the reference (block-uniform branch): narrowed again, from L2
                p0 = narrow16(pf - a.pic_stride + poff);
                p1 = narrow16(pf - a.pic_stride + poff + a.size_x);
            } else {
                const uint8_t *ref;
                if (a.ref_planes) ref = a.ref_planes + (size_t)f * g.px;
                else ref = (t == 0 || a.static_ref) ? a.state_in + (size_t)s * g.px : yf - a.y_stride;
                p0 = ld16_stream(ref + off);
                p1 = ld16_stream(ref + off + w);
            }
            const uint32_t tol4 = (uint32_t)a.tol * 0x01010101u;
            const uint32_t ntol7 = ~tol4 & 0x7f7f7f7fu;
            if constexpr (kMask || kBox) {
                const uint32_t m0 = diff_mask16(y0, p0, tol4, ntol7);
                const uint32_t m1 = diff_mask16(y1, p1, tol4, ntol7);
                if constexpr (kBox) { bm0 = m0; bm1 = m1; }
                cnt = (uint32_t)(__popc(m0) + __popc(m1)) << 7;
                if (a.mask_bits) {
                    uint16_t *mb = a.mask_bits + ((size_t)f * g.h + row0) * g.units + unit;
                    mb[0] = (uint16_t)m0;
                    mb[g.units] = (uint16_t)m1;
                }
                if (a.mask_bytes) {
                    uint8_t *mp = a.mask_bytes + (size_t)f * g.px + off;
                    st16_stream(mp, make_uint4(mask_bytes4(m0, 0), mask_bytes4(m0, 1), mask_bytes4(m0, 2), mask_bytes4(m0, 3)));
                    st16_stream(mp + w, make_uint4(mask_bytes4(m1, 0), mask_bytes4(m1, 1), mask_bytes4(m1, 2), mask_bytes4(m1, 3)));
                }
            } else {
                cnt = diff_count16(y1, p1, tol4, ntol7, diff_count16(y0, p0, tol4, ntol7, 0u));
            }
            if (a.state_out && t == a.n_frames - 1) {
                // prev_frm <- curr_frm (lib/JM/image.c:1553-1561): the last frame becomes the state
                uint8_t *so = a.state_out + (size_t)s * g.px + off;
                *reinterpret_cast<uint4 *>(so) = y0;
                *reinterpret_cast<uint4 *>(so + w) = y1;
            }
        }

        if constexpr (kConv) {
            uint2 u, v;
            if constexpr (kEarly) {
                u = u_early;
                v = v_early;
            } else if constexpr (kPic) {
                const uint32_t pcoff = a.pic_c_off + (uint32_t)(row0 >> 1) * (uint32_t)(a.size_x >> 1) + (uint32_t)unit * 8u;
                u = narrow8(pf + pcoff);
                v = narrow8(pf + pcoff + a.pic_c_plane);
            } else {
                const uint8_t *uf = yf + g.px;
                const uint32_t coff = (uint32_t)(row0 >> 1) * (uint32_t)(w >> 1) + (uint32_t)unit * 8u;
                u = ld8_stream(uf + coff);
                v = ld8_stream(uf + (g.px >> 2) + coff);
            }
            Chroma2 c[8];
            c[0] = chroma_terms(ubyte<0>(u.x), ubyte<0>(v.x));
            c[1] = chroma_terms(ubyte<1>(u.x), ubyte<1>(v.x));
            c[2] = chroma_terms(ubyte<2>(u.x), ubyte<2>(v.x));
            c[3] = chroma_terms(ubyte<3>(u.x), ubyte<3>(v.x));
            c[4] = chroma_terms(ubyte<0>(u.y), ubyte<0>(v.y));
            c[5] = chroma_terms(ubyte<1>(u.y), ubyte<1>(v.y));
            c[6] = chroma_terms(ubyte<2>(u.y), ubyte<2>(v.y));
            c[7] = chroma_terms(ubyte<3>(u.y), ubyte<3>(v.y));
            // bottom-up DIB: source row r lands in destination row h-1-r (Convert.obj .text+0x350..0x39f)
            uint8_t *d0 = a.bgr + (size_t)f * 3 * g.px + ((uint32_t)(g.h - 1 - row0) * 3u * (uint32_t)w + (uint32_t)unit * 48u);
            convert_row16(y0, c, d0);
            convert_row16(y1, c, d0 - 3 * w);
        }
    }

    // kBox: every thread parks its two 16-pixel mask rows (one word); behind the barrier of the macroblock reduction the
    // block's SECOND warp owns the boxes: lane tx holds the 16 x 16 mask of the block's macroblock tx in 8 registers, clips
    // it to each object's window with two bit masks (columns, rows), and the warp reduces min / max once per object
    // (REDUX) -- per object the other seven warps do nothing at all.
    __shared__ uint32_t smask[kBox ? 8 : 1][32];
    if constexpr (kBox) smask[ty][tx] = bm0 | (bm1 << 16);

    if constexpr (kDiff) {
        // 8 row-pair partial counts per macroblock -> shared memory -> the first warp finishes the block
        __shared__ uint32_t part[8][32];
        part[ty][tx] = cnt;
        __syncthreads();
        if constexpr (kBox) {
            if (ty == 1) {
                uint32_t mw[8];
#pragma unroll
                for (int i = 0; i < 8; i++) mw[i] = smask[i][tx];
                const int xs = unit * 16, yb = mby * 16;
                const bool mine = unit < g.units;            // lanes past the last macroblock of the picture hold zeros anyway
                for (int k = 0; k < a.n_obj; k++) {
                    // x0, x1, y0, y1 (inclusive) of object k, from lane k's registers
                    const int4 wd = make_int4(__shfl_sync(0xffffffffu, wpre.x, k), __shfl_sync(0xffffffffu, wpre.y, k),
                                              __shfl_sync(0xffffffffu, wpre.z, k), __shfl_sync(0xffffffffu, wpre.w, k));
                    int lo = wd.x - xs, hi = wd.y - xs, rlo = wd.z - yb, rhi = wd.w - yb;
                    lo = lo < 0 ? 0 : lo;
                    hi = hi > 15 ? 15 : hi;
                    rlo = rlo < 0 ? 0 : rlo;
                    rhi = rhi > 15 ? 15 : rhi;
                    const uint32_t cm = (mine && lo <= hi) ? ((2u << hi) - (1u << lo)) : 0u;
                    const uint32_t rm = rlo <= rhi ? ((2u << rhi) - (1u << rlo)) : 0u;
                    if (!__any_sync(0xffffffffu, cm != 0u && rm != 0u)) continue;      // the block misses the window
                    const uint32_t cm2 = cm * 0x00010001u;
                    uint32_t accx = 0u, rows = 0u;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        // rows 2i (low halfword) and 2i+1 (high halfword) of the macroblock
                        const uint32_t sel = ((rm >> (2 * i)) & 1u ? 0x0000ffffu : 0u) | ((rm >> (2 * i + 1)) & 1u ? 0xffff0000u : 0u);
                        const uint32_t v = mw[i] & cm2 & sel;
                        accx |= v;
                        rows |= (v & 0xffffu ? 1u << (2 * i) : 0u) | (v >> 16 ? 2u << (2 * i) : 0u);
                    }
                    const uint32_t anyx = (accx | (accx >> 16)) & 0xffffu;
                    if (!__any_sync(0xffffffffu, anyx != 0u)) continue;                 // nothing changed inside it
                    int mnx = 0x7fffffff, mxx = -0x7fffffff, mny = 0x7fffffff, mxy = -0x7fffffff;
                    if (anyx) {
                        mnx = xs + __ffs((int)anyx) - 1;
                        mxx = xs + 31 - __clz((int)anyx);
                        mny = yb + __ffs((int)rows) - 1;
                        mxy = yb + 31 - __clz((int)rows);
                    }
                    mnx = __reduce_min_sync(0xffffffffu, mnx);
                    mxx = __reduce_max_sync(0xffffffffu, mxx);
                    mny = __reduce_min_sync(0xffffffffu, mny);
                    mxy = __reduce_max_sync(0xffffffffu, mxy);
                    // lanes 0..3 take one bound each: fire-and-forget RED.MIN / RED.MAX (reading the bound first to skip
                    // the atomic would put a dependent global load at the very end of the block's life)
                    if (tx < 4) {
                        int *dst = reinterpret_cast<int *>(a.boxes + (size_t)f * a.n_obj + k) + tx;
                        const int v = tx == 0 ? mnx : tx == 1 ? mxx : tx == 2 ? mny : mxy;
                        if (tx & 1) atomicMax(dst, v);
                        else atomicMin(dst, v);
                    }
                }
            }
        }
        if (ty == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) tot += part[i][tx];
            tot >>= 7;                                     // diff_count16 counts in units of 128
            const bool valid = unit < g.units && mby < g.mb_h;
            const uint32_t on = valid && (int)tot > a.mb_thresh;
            if (valid) {
                const size_t k = (size_t)f * g.mb_w * g.mb_h + (size_t)mby * g.mb_w + unit;
                if (a.mb_count) a.mb_count[k] = (uint16_t)tot;
                if (a.mb_map) a.mb_map[k] = (uint8_t)on;
            }
            if (a.summary) {
                uint32_t spx = valid ? tot : 0, smb = on;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    spx += __shfl_xor_sync(0xffffffffu, spx, o);
                    smb += __shfl_xor_sync(0xffffffffu, smb, o);
                }
                if (tx == 0) {
                    if (spx) atomicAdd(a.summary + 2 * (size_t)f, spx);
                    if (smb) atomicAdd(a.summary + 2 * (size_t)f + 1, smb);
                }
            }
        }
    }
}

template <bool kConv, bool kDiff, bool kMask, bool kBox = false, bool kPic = false, bool kEarly = false>
int launch_tile_t(const KArgs &a, int total_frames, cudaStream_t st)
{
    int launches = 0;
    for (int f0 = 0; f0 < total_frames; f0 += 65535) {   // gridDim.z <= 65535
        const int nf = total_frames - f0 < 65535 ? total_frames - f0 : 65535;
        KArgs b = a;
        b.f_base = f0;
        dim3 grid((a.g.units + 31) / 32, a.g.mb_h, nf), block(32, 8);
        // flat macroblock mapping whenever the multiply-high division m / mb_w is exact for every m < mb_w * mb_h
        const unsigned long long mbw = (unsigned long long)a.g.mb_w, mbs = mbw * (unsigned long long)a.g.mb_h;
        // ... and the row tiling would leave more than a tenth of the lanes idle (A/B on B200: 720p 401 -> 384 us; 1080p
        // and 4K, 94 % busy either way, are within 1 % and keep the row tiling whose warps never straddle two rows)
        const unsigned tiled_lanes = 32u * ((a.g.units + 31) / 32);
        // (mb_w == 1 has no 32-bit reciprocal -- ceil(2^32 / 1) -- and needs none: its row tiling is the same mapping)
        if (LV_TILE_FLAT && a.g.units == a.g.mb_w && mbw >= 2 && mbs * mbw < (1ULL << 32) && 10u * a.g.units < 9u * tiled_lanes) {
            b.mbw_magic = (uint32_t)(((1ULL << 32) + mbw - 1) / mbw);
            grid = dim3((unsigned)((mbs + 31) / 32), 1, nf);
        }
        k_tile<kConv, kDiff, kMask, kBox, kPic, kEarly><<<grid, block, 0, st>>>(b);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return -(int)e;
        launches++;
    }
    return launches;
}

// One launch in front of a step with boxes: zero the summary (in place of the memset) and set every box to the reference's
// initial values {x1+1, x0-1, y1+1, y0-1} (lib/JM/image.c:4818-4821), which are also the result for an empty window.
__global__ void k_step_init(uint32_t *summary, int n_summary_words, const int4 *windows, int4 *boxes, int n_boxes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (summary && i < n_summary_words) summary[i] = 0u;
    if (i < n_boxes) {
        const int4 w = windows[i];
        boxes[i] = make_int4(w.y + 1, w.x - 1, w.w + 1, w.z - 1);
    }
}

// ---- the object boxes as a pass of their own over the WINDOWS only -----------------------------------------------------------
// lv_convert_diff_box_batch_ex(LV_BOX_WINDOWED): the plain fused step (early load order, 111 us) followed by this kernel, which
// re-reads current and reference luma inside each object's window (lib/JM/image.c:4818-4838) -- from L2 or HBM, a few KB per
// window of the reference's size -- instead of carrying the mask and the box machinery through every block of the picture
// (+8 us).  Measured (1080p x 64): 117.4 us against 119.4 us fused with one window per picture, 119.5 against 121.0 with four:
// a small kernel behind the step costs ~5 us however little it does (dependent launch + two memory latencies), so the gain is
// 2 us; a window as large as the picture costs a second, slow pass over the luma (340 us against 125): the fused form stays
// the default.  grid = (slabs of the window's rows, objects, frames).
constexpr int kBoxSlabs = 8, kBoxThreads = 64, kBoxInFlight = 8;     // small blocks: every block of a launch is resident at once
__global__ void __launch_bounds__(kBoxThreads) k_box_windows(const KArgs a)
{
    const Geom &g = a.g;
    const int f = a.f_base + blockIdx.z, k = blockIdx.y;
    const int4 wd = __ldg(a.windows + (size_t)f * a.n_obj + k);                    // x0, x1, y0, y1, inclusive
    const int x0 = wd.x < 0 ? 0 : wd.x, x1 = wd.y > g.w - 1 ? g.w - 1 : wd.y;
    const int y0 = wd.z < 0 ? 0 : wd.z, y1 = wd.w > g.h - 1 ? g.h - 1 : wd.w;
    if (x0 > x1 || y0 > y1) return;
    const int ny = y1 - y0 + 1;
    const int r0 = y0 + (int)(((long long)ny * blockIdx.x) / gridDim.x), r1 = y0 + (int)(((long long)ny * (blockIdx.x + 1)) / gridDim.x);
    if (r0 >= r1) return;
    const int s = a.n_frames == 1 ? f : (int)__umulhi((uint32_t)f, a.nf_magic);
    const int t = f - s * a.n_frames;
    const uint8_t *cur = a.y0 + (size_t)f * a.y_stride;
    const uint8_t *ref = (t == 0 || a.static_ref) ? a.state_in + (size_t)s * g.px : cur - a.y_stride;
    const uint32_t tol4 = (uint32_t)a.tol * 0x01010101u, ntol7 = ~tol4 & 0x7f7f7f7fu;
    const int wa = x0 >> 2, nw = (x1 >> 2) - wa + 1;                               // 4-pixel words of a row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int mnx = 0x7fffffff, mxx = -0x7fffffff, mny = 0x7fffffff, mxy = -0x7fffffff;
    // the slab's words as one flat index space, eight per thread in flight: a window of the reference's size is a single
    // memory latency per block.  i / nw by a multiply-high with ceil(2^32 / nw) is exact while total * nw < 2^32 (every window
    // of a 4K picture); beyond that -- windows of pictures tens of thousands of pixels wide -- the plain division
    const int total = nw * (r1 - r0);
    const bool exact = (unsigned long long)total * (unsigned)nw < (1ULL << 32);
    const uint32_t nw_magic = nw > 1 ? (uint32_t)((0x100000000ULL + (uint32_t)nw - 1) / (uint32_t)nw) : 0u;
    for (int base = threadIdx.x; base < total; base += kBoxInFlight * kBoxThreads) {
        uint32_t c[kBoxInFlight], p[kBoxInFlight];
        int rr[kBoxInFlight], xa[kBoxInFlight];
#pragma unroll
        for (int u = 0; u < kBoxInFlight; u++) {
            const int i = base + u * kBoxThreads;
            c[u] = p[u] = 0u;
            rr[u] = xa[u] = 0;
            if (i < total) {
                const int q = nw == 1 ? i : exact ? (int)__umulhi((uint32_t)i, nw_magic) : i / nw;
                rr[u] = r0 + q;
                xa[u] = (wa + (i - q * nw)) * 4;
                const size_t o = (size_t)rr[u] * g.w + (size_t)xa[u];
                c[u] = __ldg(reinterpret_cast<const uint32_t *>(cur + o));
                p[u] = __ldg(reinterpret_cast<const uint32_t *>(ref + o));
            }
        }
#pragma unroll
        for (int u = 0; u < kBoxInFlight; u++) {
            uint32_t m = diff_gt4(c[u], p[u], tol4, ntol7) & 0x80808080u;
            if (base + u * kBoxThreads >= total) m = 0u;
            if (xa[u] < x0) m &= 0xffffffffu << (8 * (x0 - xa[u]));                // pixels left of the window
            if (xa[u] + 3 > x1) m &= 0xffffffffu >> (8 * (xa[u] + 3 - x1));        // pixels right of it
            if (m) {
                const int first = (__ffs((int)m) - 1) >> 3, last = (31 - __clz((int)m)) >> 3;
                mnx = min(mnx, xa[u] + first);
                mxx = max(mxx, xa[u] + last);
                mny = min(mny, rr[u]);
                mxy = max(mxy, rr[u]);
            }
        }
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx);
    mxx = __reduce_max_sync(0xffffffffu, mxx);
    mny = __reduce_min_sync(0xffffffffu, mny);
    mxy = __reduce_max_sync(0xffffffffu, mxy);
    __shared__ int red[kBoxThreads / 32][4];
    if (lane == 0) { red[warp][0] = mnx; red[warp][1] = mxx; red[warp][2] = mny; red[warp][3] = mxy; }
    __syncthreads();
    if (threadIdx.x < 4) {
        int v = red[0][threadIdx.x], top = red[0][1];
#pragma unroll
        for (int i = 1; i < kBoxThreads / 32; i++) {
            v = (threadIdx.x & 1) ? max(v, red[i][threadIdx.x]) : min(v, red[i][threadIdx.x]);
            top = max(top, red[i][1]);
        }
        if (top > -0x7fffffff) {                                                  // something changed inside the slab
            int *dst = reinterpret_cast<int *>(a.boxes + (size_t)f * a.n_obj + k) + threadIdx.x;
            if (threadIdx.x & 1) atomicMax(dst, v);
            else atomicMin(dst, v);
        }
    }
}

int launch_box_windows(const KArgs &a, int total_frames, cudaStream_t st)
{
    int launches = 0;
    for (int f0 = 0; f0 < total_frames; f0 += 65535) {
        const int nf = total_frames - f0 < 65535 ? total_frames - f0 : 65535;
        KArgs b = a;
        b.f_base = f0;
        k_box_windows<<<dim3(kBoxSlabs, a.n_obj, nf), kBoxThreads, 0, st>>>(b);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return -(int)e;
        launches++;
    }
    return launches;
}

// ---- order of the six loads of the fused kernels --------------------------------------------------------------------------
// Measured on three B200 boards (profiles/r02_ab_kernel_variants.txt, "load order"): chroma loads issued early are 1.5 % faster
// at 1080p and 1.7 - 2 % at 720p on every board; at 4K the same order is bimodal (213 us or 235 - 245 us from one round of 30
// launches to the next, against a steady 218 - 220 us for the late order).  So the rule is static and by row length; a run-time
// tuner that timed both instantiations on the caller's first launches was built and dropped (the 4K bimodality defeats a
// handful of samples).  lv_set_kernel_variant(2 / 3) or LV_LOAD_ORDER=0 / 1 force one order (tests, A/B runs).
bool early_order(const KArgs &a)
{
    static const int forced = [] {
        const char *s = getenv("LV_LOAD_ORDER");
        return (s && (s[0] == '0' || s[0] == '1') && !s[1]) ? s[0] - '0' : -1;
    }();
    const int var = kernel_variant();
    return var >= 2 ? var == 3 : forced >= 0 ? forced == 1 : a.g.w <= 2048;
}

// Fused instantiations that come in both orders: the plain step and the step from decoder pictures (1080p x 64: 113.2 -> 111.2 us
// and 150.0 -> 142.1 us with the early order).  The mask and box instantiations keep the late order: early costs them
// registers they do not have (bit mask 117.7 -> 119.4 us, box 119.5 -> 127.6 us with spills).
template <bool kPic>
int launch_fused(const KArgs &a, int total_frames, cudaStream_t st)
{
    return early_order(a) ? launch_tile_t<true, true, false, false, kPic, true>(a, total_frames, st)
                          : launch_tile_t<true, true, false, false, kPic, false>(a, total_frames, st);
}

template <bool kConv, bool kDiff>
int launch_tile(KArgs a, int total_frames, cudaStream_t st)
{
    if (total_frames <= 0) return 0;
    // the multiply-high division is exact for f < 2^16 and n_frames < 2^16; longer batches are cut by the caller
    a.nf_magic = a.n_frames > 1 ? (uint32_t)((0x100000000ULL + (uint64_t)a.n_frames - 1) / (uint64_t)a.n_frames) : 0u;
    const bool mask = a.mask_bits || a.mask_bytes;
    if constexpr (kConv && kDiff) {
        if (a.pic0) return mask ? launch_tile_t<true, true, true, false, true>(a, total_frames, st) : launch_fused<true>(a, total_frames, st);
        if (!a.boxes && !mask) return launch_fused<false>(a, total_frames, st);
    }
    if (kDiff && a.boxes) return launch_tile_t<kConv, kDiff, kDiff, kDiff>(a, total_frames, st);   // masks too, when asked
    if (kDiff && mask) return launch_tile_t<kConv, kDiff, kDiff>(a, total_frames, st);
    return launch_tile_t<kConv, kDiff, false>(a, total_frames, st);
}

// ---- generic geometry (width % 4 == 0, not a multiple of 16): one thread per 2x2 quad ------------
__global__ void k_convert_generic(int w, int h, const uint8_t *__restrict__ y, const uint8_t *__restrict__ u,
                                  const uint8_t *__restrict__ v, uint8_t *__restrict__ bgr)
{
    const int cx = blockIdx.x * blockDim.x + threadIdx.x;
    const int cy = blockIdx.y * blockDim.y + threadIdx.y;
    if (cx >= (w >> 1) || cy >= (h >> 1)) return;
    const Chroma2 c = chroma_terms(u[(size_t)cy * (w >> 1) + cx], v[(size_t)cy * (w >> 1) + cx]);
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int row = 2 * cy + r;
        const uint8_t *yp = y + (size_t)row * w + 2 * cx;
        const uint32_t t2 = luma_pair(yp[0], yp[1]);
        const uint32_t B = chan_pair(t2, c.cb2), G = chan_pair(t2, c.ng2), R = chan_pair(t2, c.cr2);
        uint8_t *d = bgr + (size_t)(h - 1 - row) * 3 * w + 6 * cx;
        d[0] = chan_even(B); d[1] = chan_even(G); d[2] = chan_even(R);
        d[3] = chan_odd(B); d[4] = chan_odd(G); d[5] = chan_odd(R);
    }
}

// one block (16x16 threads) per macroblock
__global__ void k_diff_generic(int w, int h, const uint8_t *__restrict__ cur, const uint8_t *__restrict__ ref,
                               int tol, int mb_thresh, uint8_t *mask, uint16_t *mb_count, uint8_t *mb_map)
{
    const int x = blockIdx.x * 16 + threadIdx.x, yy = blockIdx.y * 16 + threadIdx.y;
    int m = 0;
    if (x < w && yy < h) {
        const int d = (int)cur[(size_t)yy * w + x] - (int)ref[(size_t)yy * w + x];
        m = (d > tol) || (d < -tol);   // !IsNearlyZero(d, tol), lib/JM/image.h:22
        if (mask) mask[(size_t)yy * w + x] = (uint8_t)m;
    }
    const int tot = __syncthreads_count(m);
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const size_t k = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        if (mb_count) mb_count[k] = (uint16_t)tot;
        if (mb_map) mb_map[k] = (uint8_t)(tot > mb_thresh);
    }
}

}  // namespace

// Kernel variant: 0 = tiled (default, fastest: profiles/), 1 = TMA pipeline (the one measured-and-rejected design kept
// selectable), 2 / 3 = tiled with the late / early order of the chroma loads forced (0 picks by row length, see
// launch_fused_plain).  All produce identical bytes (tests/test_gpu_parity.py runs the parity suite over each).  The
// environment variable LV_KERNEL=tile|tma sets the initial value; lv_set_kernel_variant() overrides it.
static int g_variant = -1;
int kernel_variant()
{
    if (g_variant < 0) {
        const char *e = getenv("LV_KERNEL");
        g_variant = (e && e[0] == 't' && e[1] == 'm') ? 1 : 0;
    }
    return g_variant;
}
void set_kernel_variant(int v) { g_variant = v >= 1 && v <= 3 ? v : 0; }
static bool use_tma() { return kernel_variant() == 1; }

// The fused picture kernel moves 16-byte words of 16-bit samples: luma rows, chroma rows, both chroma planes and every
// picture must start on a multiple of 8 samples, and all offsets must fit the kernel's 32-bit arithmetic.
bool pic_fusable(const Geom &g, const PicLayout &p)
{
    const size_t ypl = (size_t)p.size_x * p.size_y, cpl = (size_t)(p.size_x >> 1) * (p.size_y >> 1);
    return p.pic16 && (reinterpret_cast<uintptr_t>(p.pic16) & 15) == 0 && p.size_x % 16 == 0 && p.crop_left % 16 == 0 &&
           p.crop_top % 2 == 0 && ypl % 8 == 0 && cpl % 8 == 0 && p.samples % 8 == 0 && ypl + 2 * cpl < (1ull << 31) &&
           p.crop_left + g.w <= p.size_x && p.crop_top + g.h <= p.size_y;
}

int launch_step(const StepArgs &s, cudaStream_t st)
{
    if (use_tma() && !s.static_ref && !s.boxes && !s.pic.pic16)
        if (int r = launch_step_tma(s, st)) return r;   // 0 = geometry/alignment not eligible for the TMA pipeline
    KArgs a{};
    a.static_ref = s.static_ref;
    a.g = s.g; a.y0 = s.i420; a.y_stride = s.g.frame; a.ref_planes = nullptr;
    a.state_in = s.state_in; a.state_out = s.state_out; a.n_frames = s.n_frames;
    a.tol = s.tol; a.mb_thresh = s.mb_thresh;
    a.bgr = s.bgr; a.mask_bits = s.mask_bits; a.mask_bytes = s.mask_bytes;
    a.mb_count = s.mb_count; a.mb_map = s.mb_map; a.summary = s.summary;
    a.windows = reinterpret_cast<const int4 *>(s.windows); a.boxes = reinterpret_cast<int4 *>(s.boxes); a.n_obj = s.n_obj;
    if (s.boxes && (s.n_obj < 1 || s.n_obj > kMaxObjects || !s.windows)) return -(int)cudaErrorInvalidValue;
    const PicLayout &pl = s.pic;
    if (pl.pic16) {
        // decoder pictures as input: conversion is mandatory (the chroma planes are only reachable through the kernel),
        // boxes and the static-reference mode keep to the I420 form
        if (!s.bgr || s.boxes || !pic_fusable(s.g, pl)) return -(int)cudaErrorInvalidValue;
        a.pic0 = pl.pic16; a.pic_stride = pl.samples; a.size_x = pl.size_x;
        a.pic_y_off = (uint32_t)pl.crop_top * (uint32_t)pl.size_x + (uint32_t)pl.crop_left;
        a.pic_c_off = (uint32_t)pl.size_x * (uint32_t)pl.size_y + (uint32_t)(pl.crop_top >> 1) * (uint32_t)(pl.size_x >> 1) +
                      (uint32_t)(pl.crop_left >> 1);
        a.pic_c_plane = (uint32_t)(pl.size_x >> 1) * (uint32_t)(pl.size_y >> 1);
    }
    if (s.n_frames > 65535) return -(int)cudaErrorInvalidValue;
    // the kernel's frame -> (stream, t) split is exact below 2^16 frames per launch group: cut by streams
    const int group = 65535 / s.n_frames;
    const size_t mbs = (size_t)s.g.mb_w * s.g.mb_h;
    int launches = 0;
    for (int s0 = 0; s0 < s.n_streams; s0 += group) {
        const int ns = s.n_streams - s0 < group ? s.n_streams - s0 : group;
        const size_t f0 = (size_t)s0 * s.n_frames;
        KArgs b = a;
        b.y0 = a.y0 + f0 * s.g.frame;
        if (a.pic0) b.pic0 = a.pic0 + f0 * a.pic_stride;
        b.state_in = a.state_in + (size_t)s0 * s.g.px;
        if (a.state_out) b.state_out = a.state_out + (size_t)s0 * s.g.px;
        if (a.bgr) b.bgr = a.bgr + f0 * 3 * s.g.px;
        if (a.mask_bits) b.mask_bits = a.mask_bits + f0 * s.g.h * s.g.units;
        if (a.mask_bytes) b.mask_bytes = a.mask_bytes + f0 * s.g.px;
        if (a.mb_count) b.mb_count = a.mb_count + f0 * mbs;
        if (a.mb_map) b.mb_map = a.mb_map + f0 * mbs;
        if (a.summary) b.summary = a.summary + 2 * f0;
        if (a.boxes) { b.windows = a.windows + f0 * s.n_obj; b.boxes = a.boxes + f0 * s.n_obj; }
        KArgs step = b;
        if (s.box_windowed) step.boxes = nullptr;          // the plain step; the boxes come from the pass over the windows
        const int r = s.bgr ? launch_tile<true, true>(step, ns * s.n_frames, st) : launch_tile<false, true>(step, ns * s.n_frames, st);
        if (r < 0) return r;
        launches += r;
        if (a.boxes && s.box_windowed) {
            b.nf_magic = b.n_frames > 1 ? (uint32_t)((0x100000000ULL + (uint64_t)b.n_frames - 1) / (uint64_t)b.n_frames) : 0u;
            const int rb = launch_box_windows(b, ns * s.n_frames, st);
            if (rb < 0) return rb;
            launches += rb;
        }
    }
    return launches;
}

int launch_step_init(uint32_t *summary, size_t n_frames, const int *windows, int *boxes, int n_obj, cudaStream_t st)
{
    const size_t nb = n_frames * (size_t)n_obj, ns = summary ? 2 * n_frames : 0, n = nb > ns ? nb : ns;
    if (n == 0) return 0;
    if (n > 0x7fffffffu) return -(int)cudaErrorInvalidValue;
    k_step_init<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(summary, (int)ns, reinterpret_cast<const int4 *>(windows),
                                                            reinterpret_cast<int4 *>(boxes), (int)nb);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

int launch_convert(const Geom &g, const uint8_t *i420, int n_frames, uint8_t *bgr, cudaStream_t st)
{
    if (use_tma())
        if (int r = launch_convert_tma(g, i420, n_frames, bgr, st)) return r;
    KArgs a{};
    a.g = g; a.y0 = i420; a.y_stride = g.frame; a.n_frames = n_frames > 0 ? n_frames : 1; a.bgr = bgr;
    return launch_tile<true, false>(a, n_frames, st);
}

int launch_diff(const DiffArgs &d, cudaStream_t st)
{
    if (use_tma())
        if (int r = launch_diff_tma(d, st)) return r;
    KArgs a{};
    a.g = d.g; a.y0 = d.cur_y; a.y_stride = d.g.px; a.ref_planes = d.ref_y; a.n_frames = 1;
    a.tol = d.tol; a.mb_thresh = d.mb_thresh;
    a.mask_bits = d.mask_bits; a.mask_bytes = d.mask_bytes; a.mb_count = d.mb_count; a.mb_map = d.mb_map;
    a.summary = d.summary;
    return launch_tile<false, true>(a, d.n_frames, st);
}

int launch_convert_generic(int w, int h, const uint8_t *y, const uint8_t *u, const uint8_t *v, uint8_t *bgr,
                           cudaStream_t st)
{
    dim3 block(32, 8), grid(((w >> 1) + 31) / 32, ((h >> 1) + 7) / 8);
    k_convert_generic<<<grid, block, 0, st>>>(w, h, y, u, v, bgr);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

int launch_diff_generic(int w, int h, const uint8_t *cur, const uint8_t *ref, int tol, int mb_thresh,
                        uint8_t *mask_bytes, uint16_t *mb_count, uint8_t *mb_map, cudaStream_t st)
{
    dim3 block(16, 16), grid((w + 15) / 16, (h + 15) / 16);
    k_diff_generic<<<grid, block, 0, st>>>(w, h, cur, ref, tol, mb_thresh, mask_bytes, mb_count, mb_map);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

}  // namespace lv
