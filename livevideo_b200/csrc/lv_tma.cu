// lv_tma.cu -- OPT-IN variant (lv_set_kernel_variant(1) / LV_KERNEL=tma), NOT the default: a warp-autonomous TMA pipeline for
// the fused step (sm_100a).  Bit-identical to the tiled kernel of lv_kernels.cu but measured at about half its speed
// (148-158 us at 1080p x 64, 532-576 us at 4K x 32: stall_mio on UTMALDG with 1 KB boxes); kept as the one
// measured-and-rejected design that stays selectable.
//
// Every warp is its own producer/consumer pipeline; there is no block-level synchronisation at all.
//   item   = one macroblock-row segment of 128 pixels x 16 rows of one frame (8 macroblocks)
//   step   = one 8-row band of an item (1024 pixels): lane (u = lane&7, r = lane>>3) owns the
//            16-pixel x 2-row patch at columns 16u.., rows 2r, 2r+1 of the band
//   loads  = cp.async.bulk.tensor (TMA) into a 3-stage ring in shared memory, completion on an mbarrier:
//              * luma of frame t-1 AND t in ONE 3-D box {128 B, 8 rows, 2 frames}  (t = 0: two boxes,
//                the previous luma comes from the carried per-stream state)
//              * U and V in ONE 4-D box {64 B, 4 rows, 2 planes, 1 frame}
//   stores = the 8 x 384-byte BGR band is staged in shared memory in destination order (the DIB is
//            bottom-up, so the band's rows are reversed in the staging buffer) and written by ONE TMA
//            store; partial tiles (ragged width, 1080 % 16 = 8) are clipped / zero-filled by TMA.
//   counts = thread-local over the two steps of an item, then 2 shuffles across the 4 row-pair lanes.
// Work is dealt to warps round-robin in frame-major order, so frame t-1's luma is still in L2 when
// frame t needs it as the reference: the previous-luma byte of the 5.5 B/px never reaches DRAM.
//
// Reference semantics implemented (see lv_pixel.cuh for the arithmetic):
//   ColorSpaceConversions::YV12_to_RGB24 (Convert.h:26, cscc.lib Convert.obj .text+0x350),
//   IsNearlyZero difference (lib/JM/image.h:22; image.c:4791-4804), 16x16 macroblock scan
//   (image.c:4981, 5143-5150), prev_frm <- curr_frm (image.c:1553-1561).
#include "lv_kernels.h"
#include "lv_pixel.cuh"

#include <cuda.h>
#include <cstdio>
#include <cstdlib>

namespace lv {

namespace {

constexpr int kStages = 4;                      // ring of 8-row bands = 2 items in flight per warp
constexpr int kWarps = 4;                       // warps per CTA
constexpr int kYSlab = 8 * 128;                 // one 8-row x 128-byte luma band
constexpr int kInBytes = 2 * kYSlab + 512;      // prev band | cur band | U (4x64) | V (4x64)
constexpr int kOutBytes = 8 * 384;              // 8 rows x 128 px x 3 B
constexpr int kWarpSmem = kStages * kInBytes + kOutBytes;       // 13312
constexpr int kSmemBytes = kWarps * kWarpSmem + kWarps * kStages * 8 + 1024;   // + barriers + alignment slack

struct Maps {
    CUtensorMap y2;         // luma of all frames, box {16, 8, 2}: frames t-1 and t
    CUtensorMap y1;         // same tensor, box {16, 8, 1}            (K3: the `cur` planes)
    CUtensorMap uv;         // chroma, 4-D {w/16, h/2, 2 planes, frames}, box {8, 4, 2, 1}
    CUtensorMap state_in;   // carried previous luma, {w/8, h, streams}, box {16, 8, 1}   (K3: the `ref` planes)
    CUtensorMap state_out;  // state planes written by the last frame of every stream
    CUtensorMap bgr;        // output, {3w/8, h, frames}, box {48, 8, 1}
    CUtensorMap bgr2;       // same tensor, box {48, 2, 1}: row pairs of a band that sticks out of the frame
};

struct TArgs {
    int w, h, units, mb_w, mb_h;
    int xt;                 // 128-pixel column tiles per row
    int n_frames;           // frames per stream
    int explicit_ref;       // K3: reference plane i belongs to frame i
    int has_state_out;
    int tol, mb_thresh;
    int n_items;            // total_frames * mb_h * xt
    int dx, dmby, df;       // (number of warps in the grid) written in (x, mby, frame) digits
    int n_sm;               // CTAs are renumbered so that the CTAs resident on one SM work on adjacent items
    uint16_t *mask_bits;
    uint8_t *mask_bytes;
    uint16_t *mb_count;
    uint8_t *mb_map;
    uint32_t *summary;
};

// ---- PTX wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of the (converged) warp; lets ptxas treat the guarded region as single-threaded, so the TMA
// instructions get their uniform-register operands without a per-lane waterfall loop
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint4 lds16(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds8(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t a, const uint4 &v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- the walk over this warp's items (warp-uniform) ---------------------------------------------------
// Items are numbered i = (f*mb_h + mby)*xt + x; warp g owns items g, g+G, g+2G, ...; the coordinates
// advance with precomputed carries instead of divisions.  Every item is two 8-row bands; when the
// frame height leaves only one band in the last macroblock row (1080 = 67*16 + 8) the second band
// is a phantom: nothing is loaded, computed or stored for it (its barrier phase is completed empty so
// the ring keeps its rhythm).  Bands that are only partly outside are zero-filled / clipped by TMA.
struct Walk {
    int item;
    int f, t, s;         // frame index, frame within its stream, stream
    int mby, x;          // macroblock row, 128-pixel column tile
};

__device__ __forceinline__ void walk_next(Walk &wk, const TArgs &a, int G)
{
    wk.item += G;
    wk.x += a.dx;
    int df = a.df;
    if (wk.x >= a.xt) { wk.x -= a.xt; wk.mby++; }
    wk.mby += a.dmby;
    if (wk.mby >= a.mb_h) { wk.mby -= a.mb_h; df++; }
    wk.f += df;
    wk.t += df;
    while (wk.t >= a.n_frames) { wk.t -= a.n_frames; wk.s++; }
}

// both bands of one item -> stages `stage`, `stage + kInBytes` (barriers bar, bar + 8); one elected lane
template <bool kConv, bool kDiff>
__device__ __forceinline__ void issue_item(const Maps &m, const TArgs &a, const Walk &wk, uint32_t stage, uint32_t bar)
{
    const int cx = wk.x * 16;                       // 128 px = 16 eight-byte elements
    constexpr uint32_t bytes = kYSlab + (kDiff ? kYSlab : 0) + (kConv ? 512 : 0);
#pragma unroll
    for (int b = 0; b < 2; b++) {
        const int y0 = wk.mby * 16 + b * 8;
        const uint32_t st = stage + b * kInBytes, br = bar + b * 8;
        if (y0 >= a.h) {            // phantom band (a box entirely outside the tensor traps): complete the phase empty
            mbar_arrive(br);
            continue;
        }
        mbar_expect_tx(br, bytes);
        if (kDiff) {
            if (a.explicit_ref) {
                tma_load_3d(st, &m.state_in, cx, y0, wk.f, br);
                tma_load_3d(st + kYSlab, &m.y1, cx, y0, wk.f, br);
            } else if (wk.t == 0) {
                tma_load_3d(st, &m.state_in, cx, y0, wk.s, br);
                tma_load_3d(st + kYSlab, &m.y1, cx, y0, wk.f, br);
            } else {
                tma_load_3d(st, &m.y2, cx, y0, wk.f - 1, br);   // frames f-1 and f in one box
            }
        } else {
            tma_load_3d(st + kYSlab, &m.y1, cx, y0, wk.f, br);
        }
        if (kConv) tma_load_4d(st + 2 * kYSlab, &m.uv, wk.x * 8, y0 >> 1, 0, wk.f, br);
    }
}

template <bool kConv, bool kDiff>
__global__ void __launch_bounds__(kWarps * 32, 4)
k_tma(const __grid_constant__ Maps m, const TArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = lane & 7, r = lane >> 3;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t in0 = base + warp * kWarpSmem;
    const uint32_t out = in0 + kStages * kInBytes;
    const uint32_t bar0 = base + kWarps * kWarpSmem + warp * kStages * 8;

    if (elect_one()) {
#pragma unroll
        for (int s = 0; s < kStages; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();     // make the initialised barriers visible to the async (TMA) proxy
    }
    __syncwarp();

    const int G = (int)gridDim.x * kWarps;
    Walk cons;
    {
        // CTA b lands on SM b % n_sm (first wave, round-robin): give co-resident CTAs neighbouring items so
        // that one SM's TMA traffic stays within a few pages
        int vb = (int)blockIdx.x;
        if (a.n_sm > 0) {
            const int per = (int)gridDim.x / a.n_sm;                 // CTAs per SM (grid is a multiple of n_sm)
            vb = ((int)blockIdx.x % a.n_sm) * per + (int)blockIdx.x / a.n_sm;
        }
        const int g = vb * kWarps + warp;
        const int per_frame = a.mb_h * a.xt;
        cons.item = g;
        cons.f = g / per_frame;
        const int rem = g - cons.f * per_frame;
        cons.mby = rem / a.xt;
        cons.x = rem - cons.mby * a.xt;
        cons.s = cons.f / a.n_frames;
        cons.t = cons.f - cons.s * a.n_frames;
    }
    Walk prod = cons;

    // prologue: two items in flight
#pragma unroll
    for (int i = 0; i < 2; i++) {
        if (prod.item < a.n_items) {
            if (elect_one()) issue_item<kConv, kDiff>(m, a, prod, in0 + i * 2 * kInBytes, bar0 + i * 16);
            walk_next(prod, a, G);
        }
    }

    const uint32_t tol4 = (uint32_t)a.tol * 0x01010101u;
    const uint32_t ntol7 = ~tol4 & 0x7f7f7f7fu;
    const uint32_t lane_off = (uint32_t)(2 * r) * 128u + (uint32_t)u * 16u;
    const uint32_t uv_off = 2 * kYSlab + (uint32_t)r * 64u + (uint32_t)u * 8u;
    const uint32_t o_row0 = out + (uint32_t)(7 - 2 * r) * 384u + (uint32_t)u * 48u;
    const bool want_mask = kDiff && (a.mask_bits != nullptr || a.mask_bytes != nullptr);
    uint32_t half = 0;           // which half of the ring (stages 0,1 / 2,3) this item uses
    uint32_t parity = 0;

    while (cons.item < a.n_items) {
        uint32_t cnt = 0;        // 128 x changed pixels of this lane's macroblock column
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const uint32_t in = in0 + (half * 2 + b) * kInBytes;
            mbar_wait(bar0 + (half * 2 + b) * 8, parity);
            const int y0 = cons.mby * 16 + b * 8;          // first source row of the band
            if (y0 >= a.h) continue;                       // phantom band
            const uint4 c0 = lds16(in + kYSlab + lane_off);
            const uint4 c1 = lds16(in + kYSlab + lane_off + 128u);

            if (kDiff) {
                const uint4 p0 = lds16(in + lane_off);
                const uint4 p1 = lds16(in + lane_off + 128u);
                if (want_mask) {
                    const uint32_t m0 = diff_mask16(c0, p0, tol4, ntol7);
                    const uint32_t m1 = diff_mask16(c1, p1, tol4, ntol7);
                    cnt += (uint32_t)(__popc(m0) + __popc(m1)) << 7;     // same unit as diff_count16
                    const int unit = cons.x * 8 + u, row = y0 + 2 * r;
                    if (unit < a.units && row < a.h) {
                        if (a.mask_bits) {
                            uint16_t *mb = a.mask_bits + ((size_t)cons.f * a.h + row) * a.units + unit;
                            mb[0] = (uint16_t)m0;
                            mb[a.units] = (uint16_t)m1;
                        }
                        if (a.mask_bytes) {
                            uint8_t *mp = a.mask_bytes + ((size_t)cons.f * a.h + row) * a.w + (size_t)unit * 16;
                            *reinterpret_cast<uint4 *>(mp) = make_uint4(mask_bytes4(m0, 0), mask_bytes4(m0, 1), mask_bytes4(m0, 2), mask_bytes4(m0, 3));
                            *reinterpret_cast<uint4 *>(mp + a.w) = make_uint4(mask_bytes4(m1, 0), mask_bytes4(m1, 1), mask_bytes4(m1, 2), mask_bytes4(m1, 3));
                        }
                    }
                } else {
                    cnt = diff_count16(c0, p0, tol4, ntol7, cnt);
                    cnt = diff_count16(c1, p1, tol4, ntol7, cnt);
                }
            }

            if (kConv) {
                const uint2 uu = lds8(in + uv_off);
                const uint2 vv = lds8(in + uv_off + 256u);
                Chroma2 c[8];
                c[0] = chroma_terms(ubyte<0>(uu.x), ubyte<0>(vv.x));
                c[1] = chroma_terms(ubyte<1>(uu.x), ubyte<1>(vv.x));
                c[2] = chroma_terms(ubyte<2>(uu.x), ubyte<2>(vv.x));
                c[3] = chroma_terms(ubyte<3>(uu.x), ubyte<3>(vv.x));
                c[4] = chroma_terms(ubyte<0>(uu.y), ubyte<0>(vv.y));
                c[5] = chroma_terms(ubyte<1>(uu.y), ubyte<1>(vv.y));
                c[6] = chroma_terms(ubyte<2>(uu.y), ubyte<2>(vv.y));
                c[7] = chroma_terms(ubyte<3>(uu.y), ubyte<3>(vv.y));
                uint4 q0, q1, q2, s0, s1, s2;
                convert4(c0.x, c[0], c[1], q0.x, q0.y, q0.z);
                convert4(c0.y, c[2], c[3], q0.w, q1.x, q1.y);
                convert4(c0.z, c[4], c[5], q1.z, q1.w, q2.x);
                convert4(c0.w, c[6], c[7], q2.y, q2.z, q2.w);
                convert4(c1.x, c[0], c[1], s0.x, s0.y, s0.z);
                convert4(c1.y, c[2], c[3], s0.w, s1.x, s1.y);
                convert4(c1.z, c[4], c[5], s1.z, s1.w, s2.x);
                convert4(c1.w, c[6], c[7], s2.y, s2.z, s2.w);
                // the store that last read the staging buffer (previous band) must have drained it
                if (elect_one()) tma_wait_read<0>();
                __syncwarp();
                // band rows are staged in destination order: source row y0+j sits in staging row 7-j
                sts16(o_row0, q0); sts16(o_row0 + 16, q1); sts16(o_row0 + 32, q2);
                sts16(o_row0 - 384u, s0); sts16(o_row0 - 384u + 16, s1); sts16(o_row0 - 384u + 32, s2);
                fence_proxy_async();          // generic-proxy writes -> visible to the TMA store
                __syncwarp();
                if (elect_one()) {
                    // destination rows h-1-(y0+7) .. h-1-y0 (bottom-up DIB, Convert.obj .text+0x350..0x39f);
                    // a band that is only partly inside the frame is handled below
                    const int dst0 = a.h - 8 - y0;
                    if (dst0 >= 0) {
                        tma_store_3d(&m.bgr, out, cons.x * 48, dst0, cons.f);
                    } else {
                        // h % 8 != 0: a store box may not start at a negative row (it traps), so the valid
                        // row pairs of the band go out one by one
                        for (int j = 0; 2 * j < a.h - y0; j++)
                            tma_store_3d(&m.bgr2, out + (uint32_t)(6 - 2 * j) * 384u, cons.x * 48, a.h - 2 - y0 - 2 * j, cons.f);
                    }
                    tma_commit();
                }
            }

            if (kDiff && a.has_state_out && cons.t == a.n_frames - 1) {
                // prev_frm <- curr_frm (lib/JM/image.c:1553-1561): the band just consumed is the new state
                __syncwarp();
                if (elect_one()) {
                    tma_store_3d(&m.state_out, in + kYSlab, cons.x * 16, y0, cons.s);
                    tma_commit();
                    tma_wait_read<0>();   // the stage is refilled after this item
                }
            }
        }

        if (kDiff) {
            uint32_t tot = cnt;
            tot += __shfl_xor_sync(0xffffffffu, tot, 8);
            tot += __shfl_xor_sync(0xffffffffu, tot, 16);
            tot >>= 7;
            const int mbx = cons.x * 8 + u;
            const bool valid = r == 0 && mbx < a.mb_w;
            const uint32_t on = valid && (int)tot > a.mb_thresh;
            if (valid) {
                const size_t kk = ((size_t)cons.f * a.mb_h + cons.mby) * a.mb_w + mbx;
                if (a.mb_count) a.mb_count[kk] = (uint16_t)tot;
                if (a.mb_map) a.mb_map[kk] = (uint8_t)on;
            }
            if (a.summary) {
                uint32_t spx = valid ? tot : 0u, smb = on;
                spx += __shfl_xor_sync(0xffffffffu, spx, 1); smb += __shfl_xor_sync(0xffffffffu, smb, 1);
                spx += __shfl_xor_sync(0xffffffffu, spx, 2); smb += __shfl_xor_sync(0xffffffffu, smb, 2);
                spx += __shfl_xor_sync(0xffffffffu, spx, 4); smb += __shfl_xor_sync(0xffffffffu, smb, 4);
                if (lane == 0) {
                    if (spx) atomicAdd(a.summary + 2 * (size_t)cons.f, spx);
                    if (smb) atomicAdd(a.summary + 2 * (size_t)cons.f + 1, smb);
                }
            }
        }

        // every lane has consumed its registers from both stages of this half: refill it
        __syncwarp();
        if (prod.item < a.n_items) {
            if (elect_one()) issue_item<kConv, kDiff>(m, a, prod, in0 + half * 2 * kInBytes, bar0 + half * 16);
            walk_next(prod, a, G);
        }
        walk_next(cons, a, G);
        half ^= 1u;
        if (half == 0) parity ^= 1u;
    }
    if (elect_one()) tma_wait_all<0>();
}

// ---- host side: tensor maps ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// all maps use 8-byte elements so that a 384-byte box row is 48 elements (box dims are limited to 256)
bool make_map(CUtensorMap *out, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
              const uint32_t *box)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gd[5];
    cuuint64_t gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i < rank - 1; i++) gs[i] = strides_bytes[i];
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <bool kConv, bool kDiff>
int launch_k(const Maps &m, const TArgs &a, int sm_count, cudaStream_t st)
{
    if (a.n_items <= 0) return 0;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_tma<kConv, kDiff>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return -(int)e;
        attr_done = true;
    }
    long long ctas = ((long long)a.n_items + kWarps - 1) / kWarps;
    const long long max_ctas = (long long)sm_count * 4;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    TArgs b = a;
    b.n_sm = ctas == max_ctas ? sm_count : 0;
    {   // the warp count of the grid in (x, mby, frame) digits: the kernel walks items without dividing
        const long long G = ctas * kWarps, per_frame = (long long)a.mb_h * a.xt;
        b.df = (int)(G / per_frame);
        const int rem = (int)(G % per_frame);
        b.dmby = rem / a.xt;
        b.dx = rem % a.xt;
    }
    k_tma<kConv, kDiff><<<(unsigned)ctas, kWarps * 32, kSmemBytes, st>>>(m, b);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

}  // namespace

bool tma_supported(const Geom &g)
{
    // 16-byte global strides for every map: w % 16 (luma rows, chroma rows w/2 % 8 -> needs % 16 for the stride),
    // plane stride w*h/4 % 16, frame stride 1.5*w*h % 16
    return g.w % 32 == 0 && g.h % 2 == 0 && (g.px / 4) % 16 == 0 && g.frame % 16 == 0 && encode_fn() != nullptr;
}

static int sm_count_of_current_device()
{
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!cached[dev]) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
    return cached[dev] > 0 ? cached[dev] : 148;
}

int launch_step_tma(const StepArgs &s, cudaStream_t st)
{
    const Geom &g = s.g;
    const int total = s.n_streams * s.n_frames;
    if (total <= 0) return 0;
    if (!tma_supported(g) || !aligned16(s.i420) || !aligned16(s.state_in) || (s.state_out && !aligned16(s.state_out)) ||
        (s.bgr && !aligned16(s.bgr)))
        return 0;   // caller falls back to the tiled kernel
    Maps m;
    const uint64_t W8 = (uint64_t)g.w / 8;
    {
        const uint64_t d[3] = {W8, (uint64_t)g.h, (uint64_t)total}, sb[2] = {(uint64_t)g.w, (uint64_t)g.frame};
        const uint32_t b2[3] = {16, 8, 2}, b1[3] = {16, 8, 1};
        if (!make_map(&m.y2, s.i420, 3, d, sb, total >= 2 ? b2 : b1)) return 0;
        if (!make_map(&m.y1, s.i420, 3, d, sb, b1)) return 0;
    }
    {
        const uint64_t d[4] = {(uint64_t)g.w / 16, (uint64_t)g.h / 2, 2, (uint64_t)total};
        const uint64_t sb[3] = {(uint64_t)g.w / 2, (uint64_t)g.px / 4, (uint64_t)g.frame};
        const uint32_t b[4] = {8, 4, 2, 1};
        if (!make_map(&m.uv, s.i420 + g.px, 4, d, sb, b)) return 0;
    }
    {
        const uint64_t d[3] = {W8, (uint64_t)g.h, (uint64_t)s.n_streams}, sb[2] = {(uint64_t)g.w, (uint64_t)g.px};
        const uint32_t b[3] = {16, 8, 1};
        if (!make_map(&m.state_in, s.state_in, 3, d, sb, b)) return 0;
        if (!make_map(&m.state_out, s.state_out ? s.state_out : s.state_in, 3, d, sb, b)) return 0;
    }
    if (s.bgr) {
        const uint64_t d[3] = {3 * W8, (uint64_t)g.h, (uint64_t)total}, sb[2] = {(uint64_t)3 * g.w, (uint64_t)3 * g.px};
        const uint32_t b[3] = {48, 8, 1}, b2[3] = {48, 2, 1};
        if (!make_map(&m.bgr, s.bgr, 3, d, sb, b)) return 0;
        if (!make_map(&m.bgr2, s.bgr, 3, d, sb, b2)) return 0;
    } else {
        m.bgr = m.y1; m.bgr2 = m.y1;
    }
    TArgs a{};
    a.w = g.w; a.h = g.h; a.units = g.units; a.mb_w = g.mb_w; a.mb_h = g.mb_h;
    a.xt = (g.w + 127) / 128;
    a.n_frames = s.n_frames; a.explicit_ref = 0; a.has_state_out = s.state_out != nullptr;
    a.tol = s.tol; a.mb_thresh = s.mb_thresh;
    if ((long long)total * g.mb_h * a.xt > 0x7fffffffLL - 65536) return 0;
    a.n_items = total * g.mb_h * a.xt;
    a.mask_bits = s.mask_bits; a.mask_bytes = s.mask_bytes; a.mb_count = s.mb_count; a.mb_map = s.mb_map;
    a.summary = s.summary;
    const int sms = sm_count_of_current_device();
    return s.bgr ? launch_k<true, true>(m, a, sms, st) : launch_k<false, true>(m, a, sms, st);
}

int launch_convert_tma(const Geom &g, const uint8_t *i420, int n_frames, uint8_t *bgr, cudaStream_t st)
{
    if (n_frames <= 0) return 0;
    if (!tma_supported(g) || !aligned16(i420) || !aligned16(bgr)) return 0;
    Maps m;
    const uint64_t W8 = (uint64_t)g.w / 8;
    {
        const uint64_t d[3] = {W8, (uint64_t)g.h, (uint64_t)n_frames}, sb[2] = {(uint64_t)g.w, (uint64_t)g.frame};
        const uint32_t b1[3] = {16, 8, 1};
        if (!make_map(&m.y1, i420, 3, d, sb, b1)) return 0;
        m.y2 = m.y1; m.state_in = m.y1; m.state_out = m.y1;
    }
    {
        const uint64_t d[4] = {(uint64_t)g.w / 16, (uint64_t)g.h / 2, 2, (uint64_t)n_frames};
        const uint64_t sb[3] = {(uint64_t)g.w / 2, (uint64_t)g.px / 4, (uint64_t)g.frame};
        const uint32_t b[4] = {8, 4, 2, 1};
        if (!make_map(&m.uv, i420 + g.px, 4, d, sb, b)) return 0;
    }
    {
        const uint64_t d[3] = {3 * W8, (uint64_t)g.h, (uint64_t)n_frames}, sb[2] = {(uint64_t)3 * g.w, (uint64_t)3 * g.px};
        const uint32_t b[3] = {48, 8, 1}, b2[3] = {48, 2, 1};
        if (!make_map(&m.bgr, bgr, 3, d, sb, b)) return 0;
        if (!make_map(&m.bgr2, bgr, 3, d, sb, b2)) return 0;
    }
    TArgs a{};
    a.w = g.w; a.h = g.h; a.units = g.units; a.mb_w = g.mb_w; a.mb_h = g.mb_h;
    a.xt = (g.w + 127) / 128;
    a.n_frames = n_frames;
    if ((long long)n_frames * g.mb_h * a.xt > 0x7fffffffLL - 65536) return 0;
    a.n_items = n_frames * g.mb_h * a.xt;
    return launch_k<true, false>(m, a, sm_count_of_current_device(), st);
}

int launch_diff_tma(const DiffArgs &d, cudaStream_t st)
{
    const Geom &g = d.g;
    if (d.n_frames <= 0) return 0;
    if (!(g.w % 16 == 0 && g.h % 2 == 0 && g.px % 16 == 0 && encode_fn() != nullptr) || !aligned16(d.cur_y) ||
        !aligned16(d.ref_y))
        return 0;
    Maps m;
    const uint64_t W8 = (uint64_t)g.w / 8;
    const uint64_t dd[3] = {W8, (uint64_t)g.h, (uint64_t)d.n_frames}, sb[2] = {(uint64_t)g.w, (uint64_t)g.px};
    const uint32_t b1[3] = {16, 8, 1};
    if (!make_map(&m.y1, d.cur_y, 3, dd, sb, b1)) return 0;
    if (!make_map(&m.state_in, d.ref_y, 3, dd, sb, b1)) return 0;
    m.y2 = m.y1; m.uv = m.y1; m.state_out = m.y1; m.bgr = m.y1; m.bgr2 = m.y1;
    TArgs a{};
    a.w = g.w; a.h = g.h; a.units = g.units; a.mb_w = g.mb_w; a.mb_h = g.mb_h;
    a.xt = (g.w + 127) / 128;
    a.n_frames = 1; a.explicit_ref = 1; a.has_state_out = 0;
    a.tol = d.tol; a.mb_thresh = d.mb_thresh;
    if ((long long)d.n_frames * g.mb_h * a.xt > 0x7fffffffLL - 65536) return 0;
    a.n_items = d.n_frames * g.mb_h * a.xt;
    a.mask_bits = d.mask_bits; a.mask_bytes = d.mask_bytes; a.mb_count = d.mb_count; a.mb_map = d.mb_map;
    a.summary = d.summary;
    return launch_k<false, true>(m, a, sm_count_of_current_device(), st);
}

}  // namespace lv
