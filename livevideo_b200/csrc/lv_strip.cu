// lv_strip.cu -- the production kernel of the fused step (sm_100a): register-pipelined column strips.
//
// The kernel is bound by the integer pipes, not by HBM, once the memory system is kept busy (ncu:
// profiles/), so the design goal is the smallest possible instruction stream around the pixel math:
//   * a warp owns a strip 128 pixels wide and `rows_per_task` macroblock rows tall and walks down it in
//     8-row bands; lane (u = lane&7, rp = lane>>3) owns the 16-pixel x 2-row patch at column 16u, rows
//     2rp, 2rp+1 of the band.  Every warp-wide 16-byte load therefore covers four full 128-byte lines;
//   * all addresses are a per-warp base that advances by a constant per band plus a per-lane constant --
//     no divisions, no per-access 64-bit multiplies, no shared memory, no block-level barrier;
//   * the next band's luma / reference luma / chroma are loaded into a second register set before the
//     current band is computed (software pipelining in registers), so loads are always in flight;
//   * macroblock counts are thread-local over the two bands of a macroblock row, then two shuffles
//     across the four row-pair lanes; per-frame summaries are accumulated per strip and issued as one
//     pair of reductions per warp.
// grid = (ceil(x tiles / 4), strip segments, frames), block = 4 warps = 4 neighbouring strips, frame index
// slowest: frame t-1's luma, read as "current" one grid-slice earlier, is still in L2 when frame t reads
// it as the reference, so the 1 B/px of previous luma costs L2 bandwidth, not HBM bandwidth.
//
// Reference semantics (arithmetic in lv_pixel.cuh): ColorSpaceConversions::YV12_to_RGB24 (Convert.h:26,
// cscc.lib Convert.obj .text+0x350); IsNearlyZero difference (lib/JM/image.h:22, image.c:4791-4804);
// 16x16 macroblock scan (image.c:4981, 5143-5150); prev_frm <- curr_frm (image.c:1553-1561).
#include "lv_kernels.h"
#include "lv_pixel.cuh"

namespace lv {

namespace {

struct SArgs {
    int w, h, units, mb_w, mb_h;
    int rows_per_task;          // macroblock rows per warp task
    int n_frames;               // frames per stream
    int f_base;                 // first frame of this launch (gridDim.z <= 65535)
    uint32_t nf_magic;          // ceil(2^32 / n_frames): f / n_frames == umulhi(f, magic) for f < 2^16
    int tol, mb_thresh;
    const uint8_t *y0;          // luma of frame 0
    size_t y_stride;            // bytes between consecutive frames' luma
    size_t px;
    const uint8_t *ref_planes;  // K3: explicit reference planes (stride px), else NULL
    const uint8_t *state_in;
    uint8_t *state_out;
    uint8_t *bgr;
    uint16_t *mask_bits;
    uint8_t *mask_bytes;
    uint16_t *mb_count;
    uint8_t *mb_map;
    uint32_t *summary;
};

#ifndef LV_PREFETCH_BANDS
#define LV_PREFETCH_BANDS 5
#endif
constexpr int kPrefetch = LV_PREFETCH_BANDS;
#ifndef LV_STRIP_MINBLOCKS
#define LV_STRIP_MINBLOCKS 6
#endif
#ifndef LV_STRIP_RPT
#define LV_STRIP_RPT 4
#endif

struct Band {
    uint4 y0, y1;   // current luma, rows 2rp and 2rp+1
    uint4 p0, p1;   // reference luma
    uint2 u, v;     // chroma row rp
};

__device__ __forceinline__ uint4 ldg16(const uint8_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ uint4 ldcs16(const uint8_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }   // .cs loads cost 9 % (profiles/README.md)
__device__ __forceinline__ uint2 ldcs8(const uint8_t *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
__device__ __forceinline__ void stcs16(uint8_t *p, const uint4 &v) { __stcs(reinterpret_cast<uint4 *>(p), v); }
__device__ __forceinline__ void prefetch_l2(const uint8_t *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <bool kConv, bool kDiff>
__device__ __forceinline__ void load_band(Band &b, bool on, const uint8_t *py, const uint8_t *pr, const uint8_t *pu,
                                          const uint8_t *pv, int w)
{
    if (!on) {      // outside the frame (ragged right edge / below the last row): zeros difference to zero
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        b.y0 = z4; b.y1 = z4; b.p0 = z4; b.p1 = z4;
        b.u = make_uint2(0u, 0u); b.v = make_uint2(0u, 0u);
    } else {
        b.y0 = ldg16(py);                 // re-read as the reference by the next frame: keep it in L2
        b.y1 = ldg16(py + w);
        if (kDiff) {
            b.p0 = ldcs16(pr);            // last use: streaming
            b.p1 = ldcs16(pr + w);
        }
        if (kConv) {
            b.u = ldcs8(pu);
            b.v = ldcs8(pv);
        }
    }
}

__device__ __forceinline__ void convert_row16(const uint4 &y, const Chroma2 (&c)[8], uint8_t *dst)
{
    uint4 o0, o1, o2;
    convert4(y.x, c[0], c[1], o0.x, o0.y, o0.z);
    convert4(y.y, c[2], c[3], o0.w, o1.x, o1.y);
    convert4(y.z, c[4], c[5], o1.z, o1.w, o2.x);
    convert4(y.w, c[6], c[7], o2.y, o2.z, o2.w);
    stcs16(dst, o0);
    stcs16(dst + 16, o1);
    stcs16(dst + 32, o2);
}

template <bool kConv, bool kDiff, bool kMask>
__global__ void __launch_bounds__(128, LV_STRIP_MINBLOCKS) k_strip(const SArgs a)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int u = lane & 7, rp = lane >> 3;
    const int xt = blockIdx.x * 4 + warp;                    // 128-pixel strip index
    if (xt * 8 >= a.units) return;                           // whole warp outside (no barriers in this kernel)
    const int unit = xt * 8 + u;
    const bool col_on = unit < a.units;
    const int mby0 = blockIdx.y * a.rows_per_task;
    int mby1 = mby0 + a.rows_per_task;
    if (mby1 > a.mb_h) mby1 = a.mb_h;
    const int f = a.f_base + blockIdx.z;
    const int s = !kDiff ? 0 : (a.n_frames == 1 ? f : (int)__umulhi((uint32_t)f, a.nf_magic));
    const int t = f - s * a.n_frames;

    // per-warp bases (uniform) + per-lane offsets (32-bit)
    const int w = a.w;
    const uint8_t *yf = a.y0 + (size_t)f * a.y_stride;
    const uint8_t *rf = yf;
    if (kDiff) {
        if (a.ref_planes) rf = a.ref_planes + (size_t)f * a.px;
        else rf = (t == 0) ? a.state_in + (size_t)s * a.px : yf - a.y_stride;
    }
    int row = mby0 * 16 + 2 * rp;                            // this lane's first row in the current band
    const uint32_t yoff = (uint32_t)row * (uint32_t)w + (uint32_t)unit * 16u;
    const uint32_t coff = (uint32_t)(row >> 1) * (uint32_t)(w >> 1) + (uint32_t)unit * 8u;
    const uint8_t *py = yf + yoff, *pr = rf + yoff;
    const uint8_t *pu = yf + a.px + coff, *pv = pu + (a.px >> 2);
    uint8_t *pd = nullptr;
    if (kConv)   // bottom-up DIB: source row r lands in destination row h-1-r (Convert.obj .text+0x350..0x39f)
        pd = a.bgr + (size_t)f * 3 * a.px + (size_t)(a.h - 1 - row) * 3 * (size_t)w + (size_t)unit * 48;
    const int ystep = 8 * w, cstep = 4 * (w >> 1), dstep = 24 * w;

    const uint32_t tol4 = (uint32_t)a.tol * 0x01010101u;
    const uint32_t ntol7 = ~tol4 & 0x7f7f7f7fu;
    const bool write_state = kDiff && a.state_out != nullptr && t == a.n_frames - 1;

    Band bA, bB;
    load_band<kConv, kDiff>(bA, col_on && row < a.h, py, pr, pu, pv, w);
    uint32_t sum_px = 0, sum_mb = 0;
    uint32_t cnt = 0;                                        // 128 x changed pixels (see diff_count16)

    // one 8-row band: issue the next band's loads into `nxt`, then compute and store `cur`
    auto band = [&](Band &cur, Band &nxt, bool more) {
        const bool on = col_on && row < a.h;
        load_band<kConv, kDiff>(nxt, more && col_on && row + 8 < a.h, py + ystep, pr + ystep, pu + cstep, pv + cstep, w);
        if (kPrefetch > 0 && col_on && row + 8 * kPrefetch < a.h) {
            // pull the band kPrefetch steps down the strip from HBM into L2 (costs no registers); the
            // reference luma is not prefetched: it is the previous frame's luma and already sits in L2
            prefetch_l2(py + kPrefetch * ystep);
            prefetch_l2(py + kPrefetch * ystep + w);
            if (kConv) {
                prefetch_l2(pu + kPrefetch * cstep);
                prefetch_l2(pv + kPrefetch * cstep);
            }
        }
        if (kDiff) {
            if (kMask) {
                const uint32_t m0 = diff_mask16(cur.y0, cur.p0, tol4, ntol7);
                const uint32_t m1 = diff_mask16(cur.y1, cur.p1, tol4, ntol7);
                cnt += (uint32_t)(__popc(m0) + __popc(m1)) << 7;
                if (on) {
                    if (a.mask_bits) {
                        uint16_t *mb = a.mask_bits + ((size_t)f * a.h + row) * a.units + unit;
                        mb[0] = (uint16_t)m0;
                        mb[a.units] = (uint16_t)m1;
                    }
                    if (a.mask_bytes) {
                        uint8_t *mp = a.mask_bytes + ((size_t)f * a.h + row) * w + (size_t)unit * 16;
                        stcs16(mp, make_uint4(mask_bytes4(m0, 0), mask_bytes4(m0, 1), mask_bytes4(m0, 2), mask_bytes4(m0, 3)));
                        stcs16(mp + w, make_uint4(mask_bytes4(m1, 0), mask_bytes4(m1, 1), mask_bytes4(m1, 2), mask_bytes4(m1, 3)));
                    }
                }
            } else {
                cnt = diff_count16(cur.y0, cur.p0, tol4, ntol7, cnt);
                cnt = diff_count16(cur.y1, cur.p1, tol4, ntol7, cnt);
            }
            if (write_state && on) {
                // prev_frm <- curr_frm (lib/JM/image.c:1553-1561): the last frame becomes the state
                uint8_t *so = a.state_out + (size_t)s * a.px + (size_t)(py - yf);
                *reinterpret_cast<uint4 *>(so) = cur.y0;
                *reinterpret_cast<uint4 *>(so + w) = cur.y1;
            }
        }
#ifdef LV_STRIP_NOCOMPUTE
        if (kConv && on) {   // memory-system probe: same loads and stores, no pixel math
            uint4 q = make_uint4(cur.u.x ^ cur.p0.x, cur.u.y ^ cur.p0.y, cur.v.x ^ cur.p1.x, cur.v.y ^ cur.p1.y);
            stcs16(pd, cur.y0); stcs16(pd + 16, q); stcs16(pd + 32, cur.y1);
            stcs16(pd - (size_t)3 * w, cur.y1); stcs16(pd - (size_t)3 * w + 16, cur.y0); stcs16(pd - (size_t)3 * w + 32, q);
        }
        if (false) {
            Chroma2 c[8];
#else
        if (kConv && on) {
            Chroma2 c[8];
#endif
            c[0] = chroma_terms(ubyte<0>(cur.u.x), ubyte<0>(cur.v.x));
            c[1] = chroma_terms(ubyte<1>(cur.u.x), ubyte<1>(cur.v.x));
            c[2] = chroma_terms(ubyte<2>(cur.u.x), ubyte<2>(cur.v.x));
            c[3] = chroma_terms(ubyte<3>(cur.u.x), ubyte<3>(cur.v.x));
            c[4] = chroma_terms(ubyte<0>(cur.u.y), ubyte<0>(cur.v.y));
            c[5] = chroma_terms(ubyte<1>(cur.u.y), ubyte<1>(cur.v.y));
            c[6] = chroma_terms(ubyte<2>(cur.u.y), ubyte<2>(cur.v.y));
            c[7] = chroma_terms(ubyte<3>(cur.u.y), ubyte<3>(cur.v.y));
            convert_row16(cur.y0, c, pd);
            convert_row16(cur.y1, c, pd - (size_t)3 * w);
        }
        row += 8;
        py += ystep; pr += ystep; pu += cstep; pv += cstep;
        if (kConv) pd -= dstep;
    };

    for (int mby = mby0; mby < mby1; mby++) {
        cnt = 0;
        band(bA, bB, true);                  // rows 0..7 of the macroblock row; prefetch rows 8..15 into bB
        band(bB, bA, mby + 1 < mby1);        // rows 8..15; prefetch the next macroblock row into bA
        if (kDiff) {
            uint32_t tot = cnt;
            tot += __shfl_xor_sync(0xffffffffu, tot, 8);
            tot += __shfl_xor_sync(0xffffffffu, tot, 16);
            tot >>= 7;
            const bool valid = rp == 0 && col_on;
            const uint32_t on_mb = valid && (int)tot > a.mb_thresh;
            if (valid) {
                const size_t k = ((size_t)f * a.mb_h + mby) * a.mb_w + unit;
                if (a.mb_count) a.mb_count[k] = (uint16_t)tot;
                if (a.mb_map) a.mb_map[k] = (uint8_t)on_mb;
                sum_px += tot;
            }
            sum_mb += on_mb;
        }
    }
    if (kDiff && a.summary) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            sum_px += __shfl_xor_sync(0xffffffffu, sum_px, o);
            sum_mb += __shfl_xor_sync(0xffffffffu, sum_mb, o);
        }
        if (lane == 0) {
            if (sum_px) atomicAdd(a.summary + 2 * (size_t)f, sum_px);
            if (sum_mb) atomicAdd(a.summary + 2 * (size_t)f + 1, sum_mb);
        }
    }
}

int sm_count_cached()
{
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!cached[dev]) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
    return cached[dev] > 0 ? cached[dev] : 148;
}

template <bool kConv, bool kDiff, bool kMask>
int launch_strip(SArgs a, int total_frames, cudaStream_t st)
{
    if (total_frames <= 0) return 0;
    const int xt = (a.units + 7) / 8;
    // enough tasks to fill the machine several times over, as tall as possible otherwise
    int rpt = LV_STRIP_RPT;
    const long long want = (long long)sm_count_cached() * 24 * 4;
    while (rpt > 1 && (long long)xt * ((a.mb_h + rpt - 1) / rpt) * total_frames < want) rpt >>= 1;
    a.rows_per_task = rpt;
    int launches = 0;
    for (int f0 = 0; f0 < total_frames; f0 += 65535) {   // gridDim.z <= 65535
        const int nf = total_frames - f0 < 65535 ? total_frames - f0 : 65535;
        a.f_base = f0;
        dim3 grid((xt + 3) / 4, (a.mb_h + rpt - 1) / rpt, nf);
        k_strip<kConv, kDiff, kMask><<<grid, 128, 0, st>>>(a);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return -(int)e;
        launches++;
    }
    return launches;
}

SArgs base_args(const Geom &g)
{
    SArgs a{};
    a.w = g.w; a.h = g.h; a.units = g.units; a.mb_w = g.mb_w; a.mb_h = g.mb_h; a.px = g.px;
    a.n_frames = 1; a.nf_magic = 0;
    return a;
}

}  // namespace

bool strip_supported(const Geom &g) { return g.w % 16 == 0 && g.h % 2 == 0 && g.w >= 16; }

int launch_step_strip(const StepArgs &s, cudaStream_t st)
{
    const int total = s.n_streams * s.n_frames;
    if (total <= 0 || !strip_supported(s.g)) return 0;
    if (total >= 65536) return 0;                             // umulhi division bound; the tiled kernel takes over
    SArgs a = base_args(s.g);
    a.n_frames = s.n_frames;
    a.nf_magic = (uint32_t)((0x100000000ULL + (uint64_t)s.n_frames - 1) / (uint64_t)s.n_frames);
    a.tol = s.tol; a.mb_thresh = s.mb_thresh;
    a.y0 = s.i420; a.y_stride = s.g.frame;
    a.state_in = s.state_in; a.state_out = s.state_out;
    a.bgr = s.bgr; a.mask_bits = s.mask_bits; a.mask_bytes = s.mask_bytes;
    a.mb_count = s.mb_count; a.mb_map = s.mb_map; a.summary = s.summary;
    const bool mask = s.mask_bits || s.mask_bytes;
    if (s.bgr) return mask ? launch_strip<true, true, true>(a, total, st) : launch_strip<true, true, false>(a, total, st);
    return mask ? launch_strip<false, true, true>(a, total, st) : launch_strip<false, true, false>(a, total, st);
}

int launch_convert_strip(const Geom &g, const uint8_t *i420, int n_frames, uint8_t *bgr, cudaStream_t st)
{
    if (n_frames <= 0 || !strip_supported(g)) return 0;
    SArgs a = base_args(g);
    a.y0 = i420; a.y_stride = g.frame; a.bgr = bgr;
    return launch_strip<true, false, false>(a, n_frames, st);
}

int launch_diff_strip(const DiffArgs &d, cudaStream_t st)
{
    if (d.n_frames <= 0 || !strip_supported(d.g)) return 0;
    SArgs a = base_args(d.g);
    a.tol = d.tol; a.mb_thresh = d.mb_thresh;
    a.y0 = d.cur_y; a.y_stride = d.g.px; a.ref_planes = d.ref_y;
    a.n_frames = 1; a.nf_magic = 0;   // s = 0, t = f: unused with explicit reference planes
    a.mask_bits = d.mask_bits; a.mask_bytes = d.mask_bytes; a.mb_count = d.mb_count; a.mb_map = d.mb_map;
    a.summary = d.summary;
    const bool mask = d.mask_bits || d.mask_bytes;
    return mask ? launch_strip<false, true, true>(a, d.n_frames, st) : launch_strip<false, true, false>(a, d.n_frames, st);
}

}  // namespace lv
