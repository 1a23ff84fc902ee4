// lv_synth.cu -- synthetic frame generators on the device (benchmark / test inputs only).
// Counter-based, so the bytes are identical to the C and numpy statements of the same generator
// (SURVEY.md 8(d)): lowbias32 hash, `uniform` planes and the `camera` luma model (static background,
// per-macroblock noise amplitude, moving box).  Not part of the reference's path; it stands in for
// the JM decoder's output (lib/JM/output.c:427-451) when no bitstream is available.
#include "../../include/livevideo.h"
#include <cuda_runtime.h>

namespace {

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

struct SynthArgs {
    int kind;            // 0 uniform, 1 camera
    uint32_t seed, stream0, frame0;
    int n_streams, n_frames, w, h;
    uint8_t *out;
};

// one thread per 4 output bytes; grid.y = frame index (stream-major)
__global__ void k_synth(const SynthArgs a)
{
    const size_t px = (size_t)a.w * a.h, fsz = px + px / 2;
    const int f = blockIdx.y;
    const uint32_t stream = a.stream0 + f / a.n_frames, frame = a.frame0 + f % a.n_frames;
    const uint32_t k0 = lowbias32(a.seed + 0x9E3779B9U * (stream + 1));
    uint8_t *dst = a.out + (size_t)f * fsz;
    const uint32_t kY = lowbias32(k0 ^ (3 * frame + 0)), kU = lowbias32(k0 ^ (3 * frame + 1)),
                   kV = lowbias32(k0 ^ (3 * frame + 2));
    const uint32_t kbg = lowbias32(k0 ^ 0xB5297A4DU);
    const int mb_w = (a.w + 15) / 16, bw = a.w / 8, bh = a.h / 8;
    const int x0 = (int)((8 * frame) % (uint32_t)(a.w - bw)), y0 = (int)((4 * frame) % (uint32_t)(a.h - bh));
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < fsz / 4; q += (size_t)gridDim.x * blockDim.x) {
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const size_t i = 4 * q + b;
            uint32_t val;
            if (i >= px) {
                const bool isU = i < px + px / 4;
                const uint32_t j = (uint32_t)(isU ? i - px : i - px - px / 4);
                val = lowbias32((isU ? kU : kV) + j) >> 24;
            } else if (a.kind == 0) {
                val = lowbias32(kY + (uint32_t)i) >> 24;
            } else {
                const int x = (int)(i % a.w), y = (int)(i / a.w);
                const int bg = 32 + (int)((lowbias32(kbg + (uint32_t)i) >> 24) * 3 / 4);
                if (x >= x0 && x < x0 + bw && y >= y0 && y < y0 + bh) {
                    val = (uint32_t)(255 - bg);
                } else {
                    const int sel = lowbias32(k0 ^ 0x68E31DA4U ^ (uint32_t)((y / 16) * mb_w + x / 16)) & 3;
                    const int amp = sel == 0 ? 0 : (3 << sel);   // {0,6,12,24}
                    const uint32_t n = lowbias32(kY + (uint32_t)i) >> 8;
                    int vv = bg + (int)(n % (uint32_t)(2 * amp + 1)) - amp;
                    vv = vv < 0 ? 0 : vv > 255 ? 255 : vv;
                    val = (uint32_t)vv;
                }
            }
            word |= val << (8 * b);
        }
        reinterpret_cast<uint32_t *>(dst)[q] = word;
    }
}

}  // namespace

extern "C" int lv_gen_synthetic(int kind, uint32_t seed, uint32_t first_stream, uint32_t first_frame, int n_streams,
                                int n_frames, int width, int height, uint8_t *d_i420, void *cuda_stream)
{
    if (!d_i420 || n_streams < 0 || n_frames < 0 || kind < 0 || kind > 1) return LV_E_ARG;
    if (width < 16 || height < 16 || width % 4 != 0 || height % 2 != 0) return LV_E_SIZE;
    if ((((size_t)width * height * 3 / 2) & 3) != 0) return LV_E_SIZE;
    const int total = n_streams * n_frames;
    if (total == 0) return LV_OK;
    if (total > 65535) return LV_E_ARG;
    SynthArgs a{kind, seed, first_stream, first_frame, n_streams, n_frames, width, height, d_i420};
    const size_t words = (size_t)width * height * 3 / 2 / 4;
    unsigned bx = (unsigned)((words + 255) / 256);
    if (bx > 4096) bx = 4096;
    k_synth<<<dim3(bx, total), 256, 0, (cudaStream_t)cuda_stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? LV_OK : LV_E_CUDA;
}
