// lv_zerocopy.cu -- the drop-in conversion call on buffers the application pinned in place (lv_host_register /
// lv_host_alloc): ONE kernel reads the three planes straight from host memory over PCIe and writes the bottom-up BGR
// picture straight back, so a call is one launch + one synchronise and the upload and the download of the SAME picture
// overlap (a staged call serialises 3 H2D copies -> kernel -> D2H copy: six API calls, 203 us per 1080p picture against
// ~115 us of link time for its 3.1 MB in / 6.2 MB out).
//
// Replaces, for pinned buffers, the staging inside ColorSpaceConversions::YV12_to_RGB24 (Convert.h:26) as the
// application calls it: planes are SEPARATE allocations (MainDlg.cpp:39-51: srcY[i], srcU[i], srcV[i], RGBbuf[i]), five
// pictures back to back in CMainDlg::DisplayPic (MainDlg.cpp:946-974) -- hence a table of up to 8 pictures per launch.
//
// The kernel is link-bound (PCIe Gen5 x16, ~55 GB/s per direction, two orders of magnitude under HBM), so its layout is
// chosen for the link, not for the SMs: same thread -> pixel mapping as the tile kernel (16 px x 2 rows per thread, 512
// contiguous bytes per warp request) for the reads; the 48 bytes per thread and row are re-dealt through shared memory so
// that every store instruction of a warp writes 512 CONTIGUOUS bytes (full 128-byte write transactions on the link; the
// tile kernel's 16-bytes-at-stride-48 stores rely on L2 to merge sectors, which host memory does not get).
#include "lv_kernels.h"
#include "lv_pixel.cuh"

namespace lv {

namespace {

template <bool kStage>
__global__ void __launch_bounds__(256) k_convert_planes(const Geom g, const PlaneTable t)
{
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int unit = blockIdx.x * 32 + tx, row0 = blockIdx.y * 16 + 2 * ty, pic = blockIdx.z;
    const bool active = unit < g.units && row0 < g.h;
    const int w = g.w;
    __shared__ uint4 stage[kStage ? 8 : 1][kStage ? 2 : 1][kStage ? 96 : 1];    // per warp: two rows of 32 x 48 bytes

    uint4 o[2][3] = {};
    if (active) {
        const uint32_t off = (uint32_t)row0 * (uint32_t)w + (uint32_t)unit * 16u;
        const uint32_t coff = (uint32_t)(row0 >> 1) * (uint32_t)(w >> 1) + (uint32_t)unit * 8u;
        const uint4 y0 = __ldg(reinterpret_cast<const uint4 *>(t.y[pic] + off));
        const uint4 y1 = __ldg(reinterpret_cast<const uint4 *>(t.y[pic] + off + w));
        const uint2 u = __ldg(reinterpret_cast<const uint2 *>(t.u[pic] + coff));
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(t.v[pic] + coff));
        Chroma2 c[8];
        c[0] = chroma_terms(ubyte<0>(u.x), ubyte<0>(v.x));
        c[1] = chroma_terms(ubyte<1>(u.x), ubyte<1>(v.x));
        c[2] = chroma_terms(ubyte<2>(u.x), ubyte<2>(v.x));
        c[3] = chroma_terms(ubyte<3>(u.x), ubyte<3>(v.x));
        c[4] = chroma_terms(ubyte<0>(u.y), ubyte<0>(v.y));
        c[5] = chroma_terms(ubyte<1>(u.y), ubyte<1>(v.y));
        c[6] = chroma_terms(ubyte<2>(u.y), ubyte<2>(v.y));
        c[7] = chroma_terms(ubyte<3>(u.y), ubyte<3>(v.y));
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const uint4 &y = r ? y1 : y0;
            convert4(y.x, c[0], c[1], o[r][0].x, o[r][0].y, o[r][0].z);
            convert4(y.y, c[2], c[3], o[r][0].w, o[r][1].x, o[r][1].y);
            convert4(y.z, c[4], c[5], o[r][1].z, o[r][1].w, o[r][2].x);
            convert4(y.w, c[6], c[7], o[r][2].y, o[r][2].z, o[r][2].w);
        }
    }
    if (row0 >= g.h) return;                     // warp-uniform: a warp is one row pair
    // bottom-up DIB: source row r lands in destination row h-1-r (Convert.obj .text+0x350..0x39f)
    uint8_t *d0 = t.dst[pic] + ((size_t)(g.h - 1 - row0) * 3u * (size_t)w + (size_t)blockIdx.x * 32u * 48u);
    if constexpr (kStage) {
        const int left = g.units - (int)blockIdx.x * 32;
        const int chunks = 3 * (left < 32 ? left : 32);           // 16-byte pieces of this warp's row
#pragma unroll
        for (int r = 0; r < 2; r++) {
            uint4 *s = stage[ty][r];
            s[3 * tx] = o[r][0];
            s[3 * tx + 1] = o[r][1];
            s[3 * tx + 2] = o[r][2];
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; r++) {
            uint4 *d = reinterpret_cast<uint4 *>(d0 - (size_t)r * 3u * (size_t)w);
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int cidx = tx + 32 * j;
                if (cidx < chunks) d[cidx] = stage[ty][r][cidx];
            }
        }
    } else {
        if (active) {
#pragma unroll
            for (int r = 0; r < 2; r++) {
                uint4 *d = reinterpret_cast<uint4 *>(d0 - (size_t)r * 3u * (size_t)w) + 3 * tx;
                d[0] = o[r][0];
                d[1] = o[r][1];
                d[2] = o[r][2];
            }
        }
    }
}

}  // namespace

int launch_convert_planes(const Geom &g, const PlaneTable &t, int n, bool staged_stores, cudaStream_t st)
{
    if (n <= 0) return 0;
    if (n > PlaneTable::kMax) return -(int)cudaErrorInvalidValue;
    dim3 grid((g.units + 31) / 32, g.mb_h, n), block(32, 8);
    if (staged_stores) k_convert_planes<true><<<grid, block, 0, st>>>(g, t);
    else k_convert_planes<false><<<grid, block, 0, st>>>(g, t);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

}  // namespace lv
