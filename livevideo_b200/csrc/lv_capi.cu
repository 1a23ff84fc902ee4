// lv_capi.cu -- the C ABI of include/livevideo.h over the CUDA runtime.
//
// Host-side logic only: argument checking, the per-stream previous-luma state (ping-pong planes, so
// a launch never reads a plane that the same launch rewrites), the pinned double-buffered host step
// and the two legacy drop-in entries.  No CPU implementation of the pixel math exists here or
// anywhere in the product: if CUDA is unusable every entry fails loudly.
#include "../../include/livevideo.h"
#include "lv_kernels.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#ifndef LV_HOST_CHUNK_MB_DEFAULT
#define LV_HOST_CHUNK_MB_DEFAULT 16
#endif

namespace {

thread_local char g_err[256] = "";

int cuda_fail(cudaError_t e, const char *what)
{
    snprintf(g_err, sizeof g_err, "%s: %s (%d)", what, cudaGetErrorString(e), (int)e);
    return e == cudaErrorMemoryAllocation ? LV_E_NOMEM : LV_E_CUDA;
}

#define LV_CUDA(call)                                                  \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);            \
    } while (0)

int launch_result(int r, uint64_t &counter)
{
    if (r < 0) return cuda_fail((cudaError_t)(-r), "kernel launch");
    counter += (uint64_t)r;
    return LV_OK;
}

}  // namespace

struct lv_ctx {
    int device = 0;
    int sm_count = 0;
    lv::Geom g{};
    int max_streams = 0;
    int tol = 20, mb_thresh = 0;
    uint8_t *state[2] = {nullptr, nullptr};   // ping-pong previous-luma planes, max_streams x px each
    uint8_t *live = nullptr;                  // [max_streams] which of the two planes holds stream s's previous luma:
                                              // per STREAM, so a step over streams [a,b) touches nobody else's state
    int ref_mode = 0;                         // LV_REF_PREVIOUS / LV_REF_STATIC
    uint64_t launches = 0;
    // host-step pipeline (created on first lv_step_host)
    cudaStream_t s_copy_in = nullptr, s_compute = nullptr, s_copy_out = nullptr;
    static const int kSlots = 3;
    uint8_t *d_in[kSlots] = {}, *d_bgr[kSlots] = {};
    uint16_t *d_bits[kSlots] = {}, *d_cnt[kSlots] = {};
    uint8_t *d_map[kSlots] = {};
    uint32_t *d_sum[kSlots] = {};
    cudaEvent_t ev_in[kSlots] = {}, ev_k[kSlots] = {}, ev_out[kSlots] = {};
    int slot_frames = 0;                      // capacity of one slot in frames
    // the small outputs (macroblock counts / map, summaries) of a WHOLE host step, in the caller's layout: one download
    // per array behind the last chunk instead of three small copies between the picture downloads of every chunk
    uint16_t *d_step_cnt = nullptr;
    uint8_t *d_step_map = nullptr;
    uint32_t *d_step_sum = nullptr;
    size_t step_frames = 0;                   // their capacity in frames
    cudaEvent_t ev_small = nullptr;
    uint16_t *d_pic[kSlots] = {};             // lv_step_host_pictures: raw 16-bit pictures of a chunk
    uint8_t *d_narrow = nullptr;              // lv_convert_diff_pictures, layouts the fused kernel cannot take: I420 scratch
    size_t narrow_cap = 0;
    size_t pic_cap = 0;                       // capacity of one d_pic slot in samples
    // stand-alone upload slots (lv_upload_async)
    cudaStream_t s_upload = nullptr;
    uint8_t *d_up[2] = {nullptr, nullptr};
    size_t up_cap[2] = {0, 0};
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_up_free[2] = {nullptr, nullptr};
    bool up_pending_free[2] = {false, false};
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

bool geom_fast(int w, int h) { return w >= 16 && h >= 2 && w % 16 == 0 && h % 2 == 0; }

// The tile kernels move 16-byte words (luma, BGR, byte mask, state) and 8-byte words (chroma); frame and plane sizes
// are multiples of 16 bytes for every geometry they accept, so it is enough that each BASE pointer is 16-byte aligned.
// A misaligned vector access would raise a sticky cudaErrorMisalignedAddress that kills the context: refuse instead.
bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }
int misaligned(const char *what)
{
    snprintf(g_err, sizeof g_err, "%s is not aligned as include/livevideo.h requires", what);
    return LV_E_ARG;
}

// ---- host ranges pinned AND mapped through this library (lv_host_alloc / lv_host_register) -----------------------------
// The legacy entries look their pointers up here (no CUDA call on the way): a hit means the kernel may address the
// range directly.  dev - host is the constant offset of the mapping (0 on every UVA platform we have seen).
struct PinnedRange {
    uintptr_t base;
    size_t bytes;
    intptr_t dev_off;
};
std::mutex g_pinned_mu;
std::vector<PinnedRange> g_pinned;

void pinned_add(void *p, size_t bytes)
{
    void *d = nullptr;
    if (cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess || !d) {
        cudaGetLastError();        // not mappable here: the range is still pinned, the staged path serves it
        return;
    }
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    g_pinned.push_back({reinterpret_cast<uintptr_t>(p), bytes, (intptr_t)(reinterpret_cast<uintptr_t>(d) - reinterpret_cast<uintptr_t>(p))});
}

void pinned_remove(void *p)
{
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    for (size_t i = 0; i < g_pinned.size(); i++)
        if (g_pinned[i].base == reinterpret_cast<uintptr_t>(p)) {
            g_pinned.erase(g_pinned.begin() + (long)i);
            return;
        }
}

// device address of host bytes [p, p + n) when the whole span lies in one mapped range, else NULL
uint8_t *mapped(const void *p, size_t n)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    for (const PinnedRange &r : g_pinned)
        if (a >= r.base && a + n <= r.base + r.bytes) return reinterpret_cast<uint8_t *>(a + r.dev_off);
    return nullptr;
}

// LV_AUTO_PIN=1 (opt-in): the legacy entries pin + map, in place, every buffer they are handed for the SECOND time with the same
// address and size -- what a maintainer would do with one lv_host_register() per buffer after its malloc, for an application
// that is relinked without touching its source.  The reference allocates its planes and its RGB buffer once per dialog and
// keeps them (MainDlg.cpp:39-51), which is the only pattern this is safe for: a range stays registered until the process
// ends, so memory that is freed and handed out again by the allocator must not come through here.
struct SeenBuffer {
    uintptr_t p;
    size_t n;
    int count;
};
std::mutex g_seen_mu;
std::vector<SeenBuffer> g_seen;
void auto_pin(const void *p, size_t n)
{
    static const bool on = [] { const char *e = getenv("LV_AUTO_PIN"); return e && e[0] == '1'; }();
    if (!on || !p || !n || mapped(p, n)) return;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    {
        std::lock_guard<std::mutex> lk(g_seen_mu);
        SeenBuffer *hit = nullptr;
        for (SeenBuffer &b : g_seen)
            if (b.p == a) { hit = &b; break; }
        if (!hit) {
            if (g_seen.size() >= 256) g_seen.erase(g_seen.begin());
            g_seen.push_back({a, n, 1});
            return;
        }
        if (hit->n != n) { hit->n = n; hit->count = 1; return; }
        if (++hit->count != 2) return;                 // pin on the second sighting; one attempt per buffer
    }
    if (cudaHostRegister(const_cast<void *>(p), n, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess)
        pinned_add(const_cast<void *>(p), n);
    else
        cudaGetLastError();                             // e.g. a page shared with a range already registered: stays pageable
}

// 0 = never address host memory from a kernel (always stage through device buffers), 1 = zero-copy with the tile
// kernel's store pattern, 2 = zero-copy with link-friendly stores (default).  LV_ZEROCOPY sets the initial value.
int g_zero_copy = -1;
int zero_copy_mode()
{
    if (g_zero_copy < 0) {
        const char *e = getenv("LV_ZEROCOPY");
        g_zero_copy = (e && e[0] >= '0' && e[0] <= '2' && !e[1]) ? e[0] - '0' : 2;
    }
    return g_zero_copy;
}

void free_pipeline(lv_ctx *c)
{
    for (int i = 0; i < lv_ctx::kSlots; i++) {
        cudaFree(c->d_pic[i]); c->d_pic[i] = nullptr;
        cudaFree(c->d_in[i]); cudaFree(c->d_bgr[i]); cudaFree(c->d_bits[i]); cudaFree(c->d_cnt[i]);
        cudaFree(c->d_map[i]); cudaFree(c->d_sum[i]);
        c->d_in[i] = c->d_bgr[i] = c->d_map[i] = nullptr; c->d_bits[i] = c->d_cnt[i] = nullptr; c->d_sum[i] = nullptr;
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
        c->ev_in[i] = c->ev_k[i] = c->ev_out[i] = nullptr;
    }
    if (c->s_copy_in) cudaStreamDestroy(c->s_copy_in);
    if (c->s_compute) cudaStreamDestroy(c->s_compute);
    if (c->s_copy_out) cudaStreamDestroy(c->s_copy_out);
    c->s_copy_in = c->s_compute = c->s_copy_out = nullptr;
    c->slot_frames = 0;
    c->pic_cap = 0;
    cudaFree(c->d_step_cnt); cudaFree(c->d_step_map); cudaFree(c->d_step_sum);
    c->d_step_cnt = nullptr; c->d_step_map = nullptr; c->d_step_sum = nullptr;
    c->step_frames = 0;
    if (c->ev_small) cudaEventDestroy(c->ev_small);
    c->ev_small = nullptr;
}

// Whole-step staging of the small outputs: `frames` = every frame of one host step.
int ensure_step_small(lv_ctx *c, size_t frames)
{
    if (!c->ev_small) LV_CUDA(cudaEventCreateWithFlags(&c->ev_small, cudaEventDisableTiming));
    if (c->step_frames >= frames) return LV_OK;
    cudaFree(c->d_step_cnt); cudaFree(c->d_step_map); cudaFree(c->d_step_sum);
    c->d_step_cnt = nullptr; c->d_step_map = nullptr; c->d_step_sum = nullptr;
    c->step_frames = 0;
    const size_t mbs = (size_t)c->g.mb_w * c->g.mb_h;
    LV_CUDA(cudaMalloc(&c->d_step_cnt, sizeof(uint16_t) * mbs * frames));
    LV_CUDA(cudaMalloc(&c->d_step_map, mbs * frames));
    LV_CUDA(cudaMalloc(&c->d_step_sum, sizeof(uint32_t) * 2 * frames));
    c->step_frames = frames;
    return LV_OK;
}

// A chunk of `ns` streams x `nt` frames wrote its small outputs chunk-major into a slot; the caller's arrays are
// stream-major over the whole step (`n_frames` frames per stream).  One block per frame of the chunk; dst pointers are
// already offset to the chunk's first frame.
__global__ void k_place_small(const uint16_t *__restrict__ cnt_src, uint16_t *__restrict__ cnt_dst,
                              const uint8_t *__restrict__ map_src, uint8_t *__restrict__ map_dst,
                              const uint32_t *__restrict__ sum_src, uint32_t *__restrict__ sum_dst, int mbs, int nt, int n_frames)
{
    const int f = blockIdx.x, s = f / nt, t = f - s * nt;
    const size_t src = (size_t)f * mbs, dst = ((size_t)s * n_frames + t) * mbs;
    for (int i = threadIdx.x; i < mbs; i += blockDim.x) {
        if (cnt_src) cnt_dst[dst + i] = cnt_src[src + i];
        if (map_src) map_dst[dst + i] = map_src[src + i];
    }
    if (sum_src && threadIdx.x < 2) sum_dst[((size_t)s * n_frames + t) * 2 + threadIdx.x] = sum_src[(size_t)f * 2 + threadIdx.x];
}

// Device slots for the host step: `frames` frames each.
int ensure_pipeline(lv_ctx *c, int frames)
{
    if (c->slot_frames >= frames) return LV_OK;
    free_pipeline(c);
    const lv::Geom &g = c->g;
    LV_CUDA(cudaStreamCreateWithFlags(&c->s_copy_in, cudaStreamNonBlocking));
    LV_CUDA(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
    LV_CUDA(cudaStreamCreateWithFlags(&c->s_copy_out, cudaStreamNonBlocking));
    const size_t mbs = (size_t)g.mb_w * g.mb_h;
    for (int i = 0; i < lv_ctx::kSlots; i++) {
        LV_CUDA(cudaMalloc(&c->d_in[i], g.frame * frames));
        LV_CUDA(cudaMalloc(&c->d_bgr[i], 3 * g.px * frames));
        LV_CUDA(cudaMalloc(&c->d_bits[i], sizeof(uint16_t) * (size_t)g.units * g.h * frames));
        LV_CUDA(cudaMalloc(&c->d_cnt[i], sizeof(uint16_t) * mbs * frames));
        LV_CUDA(cudaMalloc(&c->d_map[i], mbs * frames));
        LV_CUDA(cudaMalloc(&c->d_sum[i], sizeof(uint32_t) * 2 * frames));
        LV_CUDA(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        LV_CUDA(cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
        LV_CUDA(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
    }
    c->slot_frames = frames;
    return LV_OK;
}

}  // namespace

extern "C" {

const char *lv_strerror(int code)
{
    switch (code) {
    case LV_OK: return "ok";
    case LV_E_ARG: return "invalid argument";
    case LV_E_SIZE: return "unsupported frame geometry";
    case LV_E_NODEVICE: return "no usable CUDA device";
    case LV_E_CUDA: return "CUDA runtime error";
    case LV_E_NOMEM: return "out of memory";
    case LV_E_STATE: return "invalid call sequence";
    default: return "unknown error";
    }
}

const char *lv_last_error(void) { return g_err; }
int lv_version(void) { return 100; }

int lv_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int lv_create(int device, int width, int height, int max_streams, int tol, int mb_thresh, lv_ctx **out)
{
    if (!out) return LV_E_ARG;
    *out = nullptr;
    if (max_streams < 1 || tol < 0 || tol > 255 || mb_thresh < 0 || mb_thresh > 256) return LV_E_ARG;
    if (width < 4 || height < 2 || width % 4 != 0 || height % 2 != 0) return LV_E_SIZE;
    if ((size_t)width * height > ((size_t)1 << 30)) return LV_E_SIZE;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cuda_fail(e, "cudaGetDeviceCount");
        return LV_E_NODEVICE;
    }
    if (device < 0 || device >= n) return LV_E_NODEVICE;
    DeviceGuard guard(device);
    if (!guard.ok) return LV_E_NODEVICE;
    lv_ctx *c = new (std::nothrow) lv_ctx;
    if (!c) return LV_E_NOMEM;
    c->device = device;
    c->g = lv::make_geom(width, height);
    c->max_streams = max_streams;
    c->tol = tol;
    c->mb_thresh = mb_thresh;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    const size_t bytes = c->g.px * (size_t)max_streams;
    c->live = new (std::nothrow) uint8_t[max_streams]();
    if (!c->live) {
        delete c;
        return LV_E_NOMEM;
    }
    for (int i = 0; i < 2; i++) {
        e = cudaMalloc(&c->state[i], bytes);
        if (e == cudaSuccess) e = cudaMemset(c->state[i], 0, bytes);   // calloc'ed prev_frm, image.c:564-570
        if (e != cudaSuccess) {
            int rc = cuda_fail(e, "state allocation");
            lv_destroy(c);
            return rc;
        }
    }
    *out = c;
    return LV_OK;
}

void lv_destroy(lv_ctx *c)
{
    if (!c) return;
    DeviceGuard guard(c->device);
    free_pipeline(c);
    cudaFree(c->d_narrow);
    for (int i = 0; i < 2; i++) {
        cudaFree(c->d_up[i]);
        if (c->ev_up[i]) cudaEventDestroy(c->ev_up[i]);
        if (c->ev_up_free[i]) cudaEventDestroy(c->ev_up_free[i]);
    }
    if (c->s_upload) cudaStreamDestroy(c->s_upload);
    cudaFree(c->state[0]);
    cudaFree(c->state[1]);
    delete[] c->live;
    delete c;
}

int lv_set_ref_mode(lv_ctx *c, int mode)
{
    if (!c || (mode != LV_REF_PREVIOUS && mode != LV_REF_STATIC)) return LV_E_ARG;
    c->ref_mode = mode;
    return LV_OK;
}

int lv_set_params(lv_ctx *c, int tol, int mb_thresh)
{
    if (!c || tol < 0 || tol > 255 || mb_thresh < 0 || mb_thresh > 256) return LV_E_ARG;
    c->tol = tol;
    c->mb_thresh = mb_thresh;
    return LV_OK;
}

size_t lv_frame_bytes_i420(const lv_ctx *c) { return c ? c->g.frame : 0; }
size_t lv_frame_bytes_bgr(const lv_ctx *c) { return c ? 3 * c->g.px : 0; }
size_t lv_frame_mask_units(const lv_ctx *c) { return c ? (size_t)c->g.units * c->g.h : 0; }
size_t lv_frame_mbs(const lv_ctx *c) { return c ? (size_t)c->g.mb_w * c->g.mb_h : 0; }
uint64_t lv_kernel_launches(const lv_ctx *c) { return c ? c->launches : 0; }
int lv_device_sm_count(const lv_ctx *c) { return c ? c->sm_count : 0; }

int lv_set_prev_luma(lv_ctx *c, int stream, const uint8_t *y, int is_device, void *cuda_stream)
{
    if (!c || !y || stream < 0 || stream >= c->max_streams) return LV_E_ARG;
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    LV_CUDA(cudaMemcpyAsync(c->state[c->live[stream]] + (size_t)stream * c->g.px, y, c->g.px,
                            is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (!is_device) LV_CUDA(cudaStreamSynchronize(st));
    return LV_OK;
}

int lv_get_prev_luma(lv_ctx *c, int stream, uint8_t *y, int is_device, void *cuda_stream)
{
    if (!c || !y || stream < 0 || stream >= c->max_streams) return LV_E_ARG;
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    LV_CUDA(cudaMemcpyAsync(y, c->state[c->live[stream]] + (size_t)stream * c->g.px, c->g.px,
                            is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (!is_device) LV_CUDA(cudaStreamSynchronize(st));
    return LV_OK;
}

int lv_reset_prev_luma(lv_ctx *c, void *cuda_stream)
{
    if (!c) return LV_E_ARG;
    DeviceGuard guard(c->device);
    LV_CUDA(cudaMemsetAsync(c->state[0], 0, c->g.px * (size_t)c->max_streams, (cudaStream_t)cuda_stream));
    memset(c->live, 0, (size_t)c->max_streams);
    return LV_OK;
}

int lv_convert_diff_batch(lv_ctx *c, const uint8_t *d_i420, int first_stream, int n_streams, int n_frames,
                          uint8_t *d_bgr, uint16_t *d_mask_bits, uint8_t *d_mask_bytes, uint16_t *d_mb_count,
                          uint8_t *d_mb_map, uint32_t *d_summary, void *cuda_stream)
{
    return lv_convert_diff_batch_ex(c, d_i420, first_stream, n_streams, n_frames, d_bgr, d_mask_bits, d_mask_bytes, d_mb_count,
                                    d_mb_map, d_summary, cuda_stream, 0);
}

}  // extern "C"

namespace {
int step_impl(lv_ctx *c, const uint8_t *d_i420, int first_stream, int n_streams, int n_frames, uint8_t *d_bgr,
              uint16_t *d_mask_bits, uint8_t *d_mask_bytes, uint16_t *d_mb_count, uint8_t *d_mb_map, uint32_t *d_summary,
              const int *d_windows, int n_objects, int *d_boxes, void *cuda_stream, int flags,
              const lv::PicLayout *pic = nullptr)
{
    if (!c || (!d_i420 && !pic)) return LV_E_ARG;
    if (n_streams < 0 || n_frames < 0 || first_stream < 0 || first_stream + n_streams > c->max_streams) return LV_E_ARG;
    if (!geom_fast(c->g.w, c->g.h)) return LV_E_SIZE;
    if (!aligned(d_i420, 16)) return misaligned("d_i420 (16 bytes)");
    if (!aligned(d_bgr, 16)) return misaligned("d_bgr (16 bytes)");
    if (!aligned(d_mask_bytes, 16)) return misaligned("d_mask_bytes (16 bytes)");
    if (!aligned(d_mask_bits, 2) || !aligned(d_mb_count, 2)) return misaligned("d_mask_bits / d_mb_count (2 bytes)");
    if (!aligned(d_summary, 4)) return misaligned("d_summary (4 bytes)");
    if (n_streams == 0 || n_frames == 0) return LV_OK;
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (d_boxes) {
        // one small kernel zeroes the summary (in place of the memset) and sets every box to its initial value
        int rc = launch_result(lv::launch_step_init((flags & LV_STEP_SUMMARY_PREZEROED) ? nullptr : d_summary,
                                                    (size_t)n_streams * n_frames, d_windows, d_boxes, n_objects, st), c->launches);
        if (rc != LV_OK) return rc;
    } else if (d_summary && !(flags & LV_STEP_SUMMARY_PREZEROED))
        LV_CUDA(cudaMemsetAsync(d_summary, 0, sizeof(uint32_t) * 2 * (size_t)n_streams * n_frames, st));
    const lv::Geom &g = c->g;
    const size_t mbs = (size_t)g.mb_w * g.mb_h;
    const bool is_static = c->ref_mode == LV_REF_STATIC;
    // One launch per run of streams whose previous luma lives in the same plane of the ping-pong pair (a launch reads
    // state[p] and writes state[p ^ 1], so it never reads a plane it rewrites).  Streams that always step together --
    // the whole context, or the same sub-range every time -- form ONE run; nothing is ever copied between the planes.
    for (int s0 = 0; s0 < n_streams;) {
        const int p = c->live[first_stream + s0];
        int s1 = s0 + 1;
        while (s1 < n_streams && c->live[first_stream + s1] == p) s1++;
        const size_t f0 = (size_t)s0 * n_frames;
        lv::StepArgs a{};
        a.g = g;
        if (pic) {
            a.pic = *pic;
            a.pic.pic16 = pic->pic16 + f0 * pic->samples;
        } else {
            a.i420 = d_i420 + f0 * g.frame;
        }
        a.state_in = c->state[p] + (size_t)(first_stream + s0) * g.px;
        // static background (MainDlg.cpp:238-246): the stored plane is the reference of every frame and stays as it is
        a.state_out = is_static ? nullptr : c->state[p ^ 1] + (size_t)(first_stream + s0) * g.px;
        a.static_ref = is_static ? 1 : 0;
        a.n_streams = s1 - s0; a.n_frames = n_frames;
        a.tol = c->tol; a.mb_thresh = c->mb_thresh;
        a.bgr = d_bgr ? d_bgr + f0 * 3 * g.px : nullptr;
        a.mask_bits = d_mask_bits ? d_mask_bits + f0 * g.h * g.units : nullptr;
        a.mask_bytes = d_mask_bytes ? d_mask_bytes + f0 * g.px : nullptr;
        a.mb_count = d_mb_count ? d_mb_count + f0 * mbs : nullptr;
        a.mb_map = d_mb_map ? d_mb_map + f0 * mbs : nullptr;
        a.summary = d_summary ? d_summary + 2 * f0 : nullptr;
        if (d_boxes) {
            a.windows = d_windows + 4 * f0 * n_objects;
            a.boxes = d_boxes + 4 * f0 * n_objects;
            a.n_obj = n_objects;
            a.box_windowed = (flags & LV_BOX_WINDOWED) ? 1 : 0;
        }
        const int rc = launch_result(lv::launch_step(a, st), c->launches);
        if (rc != LV_OK) return rc;
        if (!is_static)
            for (int s = s0; s < s1; s++) c->live[first_stream + s] ^= 1;
        s0 = s1;
    }
    return LV_OK;
}
}  // namespace

extern "C" {

int lv_convert_diff_batch_ex(lv_ctx *c, const uint8_t *d_i420, int first_stream, int n_streams, int n_frames,
                             uint8_t *d_bgr, uint16_t *d_mask_bits, uint8_t *d_mask_bytes, uint16_t *d_mb_count,
                             uint8_t *d_mb_map, uint32_t *d_summary, void *cuda_stream, int flags)
{
    return step_impl(c, d_i420, first_stream, n_streams, n_frames, d_bgr, d_mask_bits, d_mask_bytes, d_mb_count, d_mb_map,
                     d_summary, nullptr, 0, nullptr, cuda_stream, flags);
}

int lv_convert_diff_box_batch(lv_ctx *c, const uint8_t *d_i420, int first_stream, int n_streams, int n_frames, uint8_t *d_bgr,
                              uint16_t *d_mb_count, uint8_t *d_mb_map, uint32_t *d_summary, const int *d_windows,
                              int n_objects, int *d_boxes, void *cuda_stream)
{
    if (!d_windows || !d_boxes || n_objects < 1 || n_objects > LV_MAX_OBJECTS) return LV_E_ARG;
    if (!aligned(d_windows, 16) || !aligned(d_boxes, 16)) return misaligned("d_windows / d_boxes (16 bytes)");
    return step_impl(c, d_i420, first_stream, n_streams, n_frames, d_bgr, nullptr, nullptr, d_mb_count, d_mb_map, d_summary,
                     d_windows, n_objects, d_boxes, cuda_stream, 0);
}

int lv_convert_diff_box_batch_ex(lv_ctx *c, const uint8_t *d_i420, int first_stream, int n_streams, int n_frames, uint8_t *d_bgr,
                                 uint16_t *d_mb_count, uint8_t *d_mb_map, uint32_t *d_summary, const int *d_windows,
                                 int n_objects, int *d_boxes, void *cuda_stream, int flags)
{
    if (!d_windows || !d_boxes || n_objects < 1 || n_objects > LV_MAX_OBJECTS) return LV_E_ARG;
    if (flags & ~(LV_BOX_WINDOWED | LV_STEP_SUMMARY_PREZEROED)) return LV_E_ARG;
    if (!aligned(d_windows, 16) || !aligned(d_boxes, 16)) return misaligned("d_windows / d_boxes (16 bytes)");
    return step_impl(c, d_i420, first_stream, n_streams, n_frames, d_bgr, nullptr, nullptr, d_mb_count, d_mb_map, d_summary,
                     d_windows, n_objects, d_boxes, cuda_stream, flags);
}

int lv_object_windows(lv_ctx *c, const int *d_objects, int n, int hor_margin, int ver_margin, int *d_windows, void *cuda_stream)
{
    if (!c || !d_objects || !d_windows || n < 0) return LV_E_ARG;
    if (!aligned(d_objects, 16) || !aligned(d_windows, 16)) return misaligned("d_objects / d_windows (16 bytes)");
    if (n == 0) return LV_OK;
    DeviceGuard guard(c->device);
    return launch_result(lv::launch_object_windows(c->g.w, c->g.h, d_objects, n, hor_margin, ver_margin, d_windows,
                                                   (cudaStream_t)cuda_stream), c->launches);
}

int lv_object_update(lv_ctx *c, const int *d_boxes, int n, int *d_objects, void *cuda_stream)
{
    if (!c || !d_boxes || !d_objects || n < 0) return LV_E_ARG;
    if (!aligned(d_boxes, 16) || !aligned(d_objects, 16)) return misaligned("d_boxes / d_objects (16 bytes)");
    if (n == 0) return LV_OK;
    DeviceGuard guard(c->device);
    return launch_result(lv::launch_object_update(d_boxes, n, d_objects, (cudaStream_t)cuda_stream), c->launches);
}

int lv_convert_batch(lv_ctx *c, const uint8_t *d_i420, int n_frames, uint8_t *d_bgr, void *cuda_stream)
{
    if (!c || !d_i420 || !d_bgr || n_frames < 0) return LV_E_ARG;
    if (!geom_fast(c->g.w, c->g.h)) return LV_E_SIZE;
    if (!aligned(d_i420, 16) || !aligned(d_bgr, 16)) return misaligned("d_i420 / d_bgr (16 bytes)");
    if (n_frames == 0) return LV_OK;
    DeviceGuard guard(c->device);
    return launch_result(lv::launch_convert(c->g, d_i420, n_frames, d_bgr, (cudaStream_t)cuda_stream), c->launches);
}

int lv_diff_batch(lv_ctx *c, const uint8_t *d_cur_y, const uint8_t *d_ref_y, int n_frames, uint16_t *d_mask_bits,
                  uint8_t *d_mask_bytes, uint16_t *d_mb_count, uint8_t *d_mb_map, uint32_t *d_summary,
                  void *cuda_stream)
{
    if (!c || !d_cur_y || !d_ref_y || n_frames < 0) return LV_E_ARG;
    if (!geom_fast(c->g.w, c->g.h)) return LV_E_SIZE;
    if (!aligned(d_cur_y, 16) || !aligned(d_ref_y, 16) || !aligned(d_mask_bytes, 16))
        return misaligned("d_cur_y / d_ref_y / d_mask_bytes (16 bytes)");
    if (!aligned(d_mask_bits, 2) || !aligned(d_mb_count, 2) || !aligned(d_summary, 4))
        return misaligned("d_mask_bits / d_mb_count (2 bytes) / d_summary (4 bytes)");
    if (n_frames == 0) return LV_OK;
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (d_summary) LV_CUDA(cudaMemsetAsync(d_summary, 0, sizeof(uint32_t) * 2 * (size_t)n_frames, st));
    lv::DiffArgs a{};
    a.g = c->g; a.cur_y = d_cur_y; a.ref_y = d_ref_y; a.n_frames = n_frames;
    a.tol = c->tol; a.mb_thresh = c->mb_thresh;
    a.mask_bits = d_mask_bits; a.mask_bytes = d_mask_bytes; a.mb_count = d_mb_count; a.mb_map = d_mb_map;
    a.summary = d_summary;
    return launch_result(lv::launch_diff(a, st), c->launches);
}

int lv_write_bks(lv_ctx *c, const uint8_t *d_cur, const uint8_t *d_bkd, int n_frames, int tol, uint8_t *d_out,
                 uint16_t *d_obj_bits, void *cuda_stream)
{
    if (!c || !d_cur || !d_bkd || !d_out || n_frames < 0 || tol < 0 || tol > 255) return LV_E_ARG;
    if (d_obj_bits && c->g.w % 16 != 0) return LV_E_SIZE;
    // 16-byte aligned frames take the vector kernel, anything 4-byte aligned the word kernel
    if (!aligned(d_cur, 4) || !aligned(d_bkd, 4) || !aligned(d_out, 4)) return misaligned("d_cur_i420 / d_bkd_i420 / d_out_i420 (4 bytes)");
    if (!aligned(d_obj_bits, 2)) return misaligned("d_obj_bits (2 bytes)");
    if (n_frames == 0) return LV_OK;
    DeviceGuard guard(c->device);
    return launch_result(lv::launch_write_bks(c->g, d_cur, d_bkd, n_frames, tol, d_out, d_obj_bits, (cudaStream_t)cuda_stream),
                         c->launches);
}

int lv_object_box_batch(lv_ctx *c, const void *d_mask, int mask_is_bits, int n_frames, int x0, int x1, int y0, int y1,
                        int *d_boxes, void *cuda_stream)
{
    if (!c || !d_mask || !d_boxes || n_frames < 0) return LV_E_ARG;
    if (x0 < 0 || y0 < 0 || x1 >= c->g.w || y1 >= c->g.h) return LV_E_ARG;
    if (mask_is_bits && c->g.w % 16 != 0) return LV_E_SIZE;
    if ((reinterpret_cast<uintptr_t>(d_boxes) & 15) != 0) return LV_E_ARG;
    if (!mask_is_bits && c->g.w % 16 == 0 && (reinterpret_cast<uintptr_t>(d_mask) & 15) != 0) return LV_E_ARG;
    if (n_frames == 0) return LV_OK;
    DeviceGuard guard(c->device);
    return launch_result(lv::launch_object_box(c->g.w, c->g.h, static_cast<const uint8_t *>(d_mask), mask_is_bits != 0, n_frames,
                                               x0, x1, y0, y1, d_boxes, (cudaStream_t)cuda_stream), c->launches);
}

int lv_object_box(lv_ctx *c, const void *d_mask, int mask_is_bits, int x0, int x1, int y0, int y1, int *d_box4,
                  void *cuda_stream)
{
    return lv_object_box_batch(c, d_mask, mask_is_bits, 1, x0, x1, y0, y1, d_box4, cuda_stream);
}

int lv_narrow_pictures(lv_ctx *c, const uint16_t *d_pic16, int n_frames, int size_x, int size_y, int crop_left, int crop_top,
                       uint8_t *d_i420, void *cuda_stream)
{
    if (!c || !d_pic16 || !d_i420 || n_frames < 0) return LV_E_ARG;
    // 4:2:0 pictures: even sizes, even luma crops (twice the SPS offsets, lib/JM/output.c:375-380), crop inside the picture
    if (size_x <= 0 || size_y <= 0 || (size_x & 1) || (size_y & 1) || crop_left < 0 || crop_top < 0 || (crop_left & 1) ||
        (crop_top & 1) || crop_left + c->g.w > size_x || crop_top + c->g.h > size_y)
        return LV_E_SIZE;
    if ((reinterpret_cast<uintptr_t>(d_pic16) & 1) != 0) return LV_E_ARG;
    if (n_frames == 0) return LV_OK;
    DeviceGuard guard(c->device);
    return launch_result(lv::launch_narrow(c->g, d_pic16, n_frames, size_x, size_y, crop_left, crop_top, d_i420,
                                           (cudaStream_t)cuda_stream), c->launches);
}

int lv_convert_diff_pictures(lv_ctx *c, const uint16_t *d_pic16, int size_x, int size_y, int crop_left, int crop_top,
                             int first_stream, int n_streams, int n_frames, uint8_t *d_bgr, uint16_t *d_mask_bits,
                             uint8_t *d_mask_bytes, uint16_t *d_mb_count, uint8_t *d_mb_map, uint32_t *d_summary, void *cuda_stream)
{
    if (!c || !d_pic16 || !d_bgr || n_streams < 0 || n_frames < 0) return LV_E_ARG;
    if (size_x <= 0 || size_y <= 0 || (size_x & 1) || (size_y & 1) || crop_left < 0 || crop_top < 0 || (crop_left & 1) ||
        (crop_top & 1) || crop_left + c->g.w > size_x || crop_top + c->g.h > size_y)
        return LV_E_SIZE;
    if ((reinterpret_cast<uintptr_t>(d_pic16) & 1) != 0) return misaligned("d_pic16 (2 bytes)");
    lv::PicLayout p;
    p.pic16 = d_pic16; p.size_x = size_x; p.size_y = size_y; p.crop_left = crop_left; p.crop_top = crop_top;
    p.samples = (size_t)size_x * size_y + 2 * ((size_t)(size_x >> 1) * (size_y >> 1));
    if (geom_fast(c->g.w, c->g.h) && lv::pic_fusable(c->g, p) && lv::kernel_variant() != 1)
        return step_impl(c, nullptr, first_stream, n_streams, n_frames, d_bgr, d_mask_bits, d_mask_bytes, d_mb_count, d_mb_map,
                         d_summary, nullptr, 0, nullptr, cuda_stream, 0, &p);
    // a layout the fused kernel cannot take (odd pitches / crops, misaligned pictures): narrow into a scratch, then step
    const size_t n = (size_t)n_streams * n_frames;
    if (n == 0) return LV_OK;
    DeviceGuard guard(c->device);
    if (c->narrow_cap < n * c->g.frame) {
        LV_CUDA(cudaStreamSynchronize((cudaStream_t)cuda_stream));
        cudaFree(c->d_narrow);
        c->d_narrow = nullptr;
        c->narrow_cap = 0;
        LV_CUDA(cudaMalloc(&c->d_narrow, n * c->g.frame));
        c->narrow_cap = n * c->g.frame;
    }
    int rc = lv_narrow_pictures(c, d_pic16, (int)n, size_x, size_y, crop_left, crop_top, c->d_narrow, cuda_stream);
    if (rc != LV_OK) return rc;
    return step_impl(c, c->d_narrow, first_stream, n_streams, n_frames, d_bgr, d_mask_bits, d_mask_bytes, d_mb_count, d_mb_map,
                     d_summary, nullptr, 0, nullptr, cuda_stream, 0);
}

int lv_convert_format(lv_ctx *c, int kind, const uint8_t *d_in, uint8_t *d_out, int n_frames, void *cuda_stream)
{
    if (!c || !d_in || !d_out || kind < 1 || kind > 5 || n_frames < 0) return LV_E_ARG;
    if (c->g.w % 4 != 0 || c->g.h % 4 != 0) return LV_E_SIZE;
    if (n_frames == 0) return LV_OK;
    DeviceGuard guard(c->device);
    return launch_result(lv::launch_format(kind, c->g.w, c->g.h, d_in, d_out, n_frames, (cudaStream_t)cuda_stream), c->launches);
}

size_t lv_format_bytes(int kind, int width, int height, int output)
{
    if (kind < 1 || kind > 5 || width <= 0 || height <= 0) return 0;
    return output ? lv::format_bytes_out(kind, width, height) : lv::format_bytes_in(kind, width, height);
}

int lv_set_kernel_variant(int variant)
{
    if (variant < 0 || variant > 3) return LV_E_ARG;
    lv::set_kernel_variant(variant);
    return LV_OK;
}

int lv_get_kernel_variant(void) { return lv::kernel_variant(); }

int lv_upload_async(lv_ctx *c, const uint8_t *h_i420, int n_frames_total, int slot, const uint8_t **d_i420)
{
    if (!c || !h_i420 || !d_i420 || n_frames_total < 1 || slot < 0 || slot > 1) return LV_E_ARG;
    DeviceGuard guard(c->device);
    if (!c->s_upload) {
        LV_CUDA(cudaStreamCreateWithFlags(&c->s_upload, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            LV_CUDA(cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming));
            LV_CUDA(cudaEventCreateWithFlags(&c->ev_up_free[i], cudaEventDisableTiming));
        }
    }
    const size_t bytes = c->g.frame * (size_t)n_frames_total;
    // the consumer of this slot's previous contents must be done before the slot is overwritten (or re-allocated)
    if (c->up_pending_free[slot]) LV_CUDA(cudaStreamWaitEvent(c->s_upload, c->ev_up_free[slot], 0));
    if (bytes > c->up_cap[slot]) {
        if (c->up_pending_free[slot]) LV_CUDA(cudaEventSynchronize(c->ev_up_free[slot]));
        LV_CUDA(cudaStreamSynchronize(c->s_upload));
        cudaFree(c->d_up[slot]);
        c->d_up[slot] = nullptr;
        c->up_cap[slot] = 0;
        LV_CUDA(cudaMalloc(&c->d_up[slot], bytes));
        c->up_cap[slot] = bytes;
    }
    c->up_pending_free[slot] = false;
    LV_CUDA(cudaMemcpyAsync(c->d_up[slot], h_i420, bytes, cudaMemcpyHostToDevice, c->s_upload));
    LV_CUDA(cudaEventRecord(c->ev_up[slot], c->s_upload));
    *d_i420 = c->d_up[slot];
    return LV_OK;
}

int lv_upload_wait(lv_ctx *c, int slot, void *cuda_stream)
{
    if (!c || slot < 0 || slot > 1) return LV_E_ARG;
    if (!c->s_upload) return LV_E_STATE;
    DeviceGuard guard(c->device);
    LV_CUDA(cudaStreamWaitEvent((cudaStream_t)cuda_stream, c->ev_up[slot], 0));
    return LV_OK;
}

int lv_upload_release(lv_ctx *c, int slot, void *cuda_stream)
{
    if (!c || slot < 0 || slot > 1) return LV_E_ARG;
    if (!c->s_upload) return LV_E_STATE;
    DeviceGuard guard(c->device);
    LV_CUDA(cudaEventRecord(c->ev_up_free[slot], (cudaStream_t)cuda_stream));
    c->up_pending_free[slot] = true;
    return LV_OK;
}

int lv_host_alloc(void **p, size_t bytes)
{
    if (!p) return LV_E_ARG;
    *p = nullptr;
    LV_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped));
    pinned_add(*p, bytes ? bytes : 1);
    return LV_OK;
}

int lv_host_free(void *p)
{
    if (!p) return LV_OK;
    pinned_remove(p);
    LV_CUDA(cudaFreeHost(p));
    return LV_OK;
}

// Pin memory the application already owns (the reference's dialogs malloc their planes and the RGB buffer once,
// MainDlg.cpp:39-51, and reuse them for every picture): the legacy entries and lv_step_host then DMA straight from / to
// it instead of going through the driver's pageable staging, and -- the range is also MAPPED into the device's address
// space -- the conversion entries read and write it directly from the kernel (lv_zerocopy.cu).  The range must stay
// allocated until lv_host_unregister.
int lv_host_register(void *p, size_t bytes)
{
    if (!p || !bytes) return LV_E_ARG;
    LV_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    pinned_add(p, bytes);
    return LV_OK;
}

int lv_host_is_mapped(const void *p, size_t bytes) { return p && bytes && mapped(p, bytes) ? 1 : 0; }

int lv_host_unregister(void *p)
{
    if (!p) return LV_OK;
    pinned_remove(p);
    LV_CUDA(cudaHostUnregister(p));
    return LV_OK;
}

int lv_set_zero_copy(int mode)
{
    if (mode < 0 || mode > 2) return LV_E_ARG;
    g_zero_copy = mode;
    return LV_OK;
}

int lv_get_zero_copy(void) { return zero_copy_mode(); }

// Host step.  The batch is cut along the frame axis into chunks (all streams, a few frames each):
// chunk k's previous luma is chunk k-1's last frame, which is exactly what the carried state holds
// after chunk k-1's kernel, so chunking does not change any result.
//   copy-in stream : H2D chunk k          (waits until the kernel that last used the slot is done)
//   compute stream : kernel chunk k       (waits for its H2D)
//   copy-out stream: D2H chunk k          (waits for its kernel)
// With 3 slots the upload of k+1, the kernel of k and the download of k-1 are in flight together.
namespace {
// optional 16-bit source of the host step (lv_step_host_pictures): decoder pictures instead of I420 frames
struct PicSource {
    const uint16_t *h_pic16;
    int size_x, size_y, crop_left, crop_top;
    size_t samples;              // per picture
};

// ---- how a host step is cut into chunks (pure host logic: lv_host_step_plan exposes it, the CPU tests walk it) ----------
// chunk = `nt` consecutive frames of `ns` consecutive streams.  Sizes: ~LV_HOST_CHUNK_MB MiB of input per chunk, at least one
// frame; a device slot holds at most LV_HOST_SLOT_MB MiB of I420 input (default 256), so when one frame of all streams is
// larger the streams are cut too (streams keep their state independently: a cut by streams costs nothing).  The first chunks
// of a step ramp up (1, 2, 4 ... frames, at least 1 MiB of input): nothing can be downloaded before the first chunk has been
// uploaded and computed, so the pipeline is filled with a small one.
struct HostChunk {
    int s0, ns, t0, nt;
};
struct HostPlan {
    int cs = 0, cf = 0;              // largest chunk: streams x frames (what a slot must hold)
    std::vector<HostChunk> chunks;
};
HostPlan plan_host_step(size_t in_frame, size_t dev_frame, int n_streams, int n_frames)
{
    auto env_mb = [](const char *name, long dflt, long max) -> size_t {
        const char *e = getenv(name);
        const long mb = e ? atol(e) : 0;
        return (size_t)(mb > 0 && mb <= max ? mb : dflt) << 20;
    };
    HostPlan p;
    if (n_streams <= 0 || n_frames <= 0 || !in_frame || !dev_frame) return p;
    const size_t chunk_bytes = env_mb("LV_HOST_CHUNK_MB", LV_HOST_CHUNK_MB_DEFAULT, 1024);
    int cf = (int)(chunk_bytes / (in_frame * (size_t)n_streams));
    if (cf < 1) cf = 1;
    if (cf > n_frames) cf = n_frames;
    int cs = n_streams;
    const size_t slot_bytes = env_mb("LV_HOST_SLOT_MB", 256, 16384);
    const size_t cap_frames = slot_bytes / dev_frame > 0 ? slot_bytes / dev_frame : 1;
    if ((size_t)cs * cf > cap_frames) { cf = 1; if ((size_t)cs > cap_frames) cs = (int)cap_frames; }
    p.cs = cs; p.cf = cf;
    const char *ramp_env = getenv("LV_HOST_RAMP");
    const bool ramp_on = !(ramp_env && ramp_env[0] == '0');
    const size_t ramp_in = in_frame * (size_t)cs;                       // input bytes of one frame of every stream in a chunk
    const int ramp0 = ramp_in >= ((size_t)1 << 20) ? 1 : (int)((((size_t)1 << 20) + ramp_in - 1) / ramp_in);   // >= 1 MiB
    for (int s0 = 0; s0 < n_streams; s0 += cs) {
        const int ns = n_streams - s0 < cs ? n_streams - s0 : cs;
        int ramp = s0 == 0 && ramp_on ? ramp0 : cf;                     // later stream groups find the pipeline full
        for (int t0 = 0, nt = 0; t0 < n_frames; t0 += nt) {
            nt = ramp < cf ? ramp : cf;
            if (nt > n_frames - t0) nt = n_frames - t0;
            ramp = ramp < cf ? ramp * 2 : cf;
            p.chunks.push_back(HostChunk{s0, ns, t0, nt});
        }
    }
    return p;
}

int ensure_pic_slots(lv_ctx *c, size_t samples)
{
    if (c->pic_cap >= samples) return LV_OK;
    for (int i = 0; i < lv_ctx::kSlots; i++) {
        LV_CUDA(cudaStreamSynchronize(c->s_compute));
        cudaFree(c->d_pic[i]);
        c->d_pic[i] = nullptr;
        LV_CUDA(cudaMalloc(&c->d_pic[i], samples * sizeof(uint16_t)));
    }
    c->pic_cap = samples;
    return LV_OK;
}

int step_host_impl(lv_ctx *c, const uint8_t *h_i420, const PicSource *pic, int first_stream, int n_streams, int n_frames,
                   uint8_t *h_bgr, uint16_t *h_mask_bits, uint16_t *h_mb_count, uint8_t *h_mb_map, uint32_t *h_summary)
{
    if (!c || (!h_i420 && !pic)) return LV_E_ARG;
    if (n_streams < 0 || n_frames < 0 || first_stream < 0 || first_stream + n_streams > c->max_streams) return LV_E_ARG;
    if (!geom_fast(c->g.w, c->g.h)) return LV_E_SIZE;
    if (n_streams == 0 || n_frames == 0) return LV_OK;
    DeviceGuard guard(c->device);
    const lv::Geom &g = c->g;
    const size_t mbs = (size_t)g.mb_w * g.mb_h, units = (size_t)g.units * g.h;

    const size_t in_frame = pic ? pic->samples * sizeof(uint16_t) : g.frame;     // host bytes per frame
    const HostPlan plan = plan_host_step(in_frame, g.frame, n_streams, n_frames);
    const int cs = plan.cs, cf = plan.cf;
    int rc = ensure_pipeline(c, cs * cf);
    if (rc != LV_OK) return rc;
    if (pic) {
        rc = ensure_pic_slots(c, pic->samples * (size_t)cs * cf);
        if (rc != LV_OK) return rc;
    }
    const bool small = h_mb_count || h_mb_map || h_summary;
    if (small) {
        rc = ensure_step_small(c, (size_t)n_streams * n_frames);
        if (rc != LV_OK) return rc;
    }

    // A chunk is `ns` runs of `nt` consecutive frames, `n_frames` frames apart in the host batch.  One strided copy moves
    // it when both pitches fit the device's limit (cudaMemcpy2DAsync rejects larger ones: 2 GiB - 1 on B200, i.e. 87
    // frames of 4K BGR); a single run is one plain copy; otherwise one plain copy per stream.
    int max_pitch = 0;
    cudaDeviceGetAttribute(&max_pitch, cudaDevAttrMaxPitch, c->device);
    auto move = [&](void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, int rows, cudaMemcpyKind kind,
                    cudaStream_t stm) -> cudaError_t {
        if (rows == 1 || (dpitch == width && spitch == width)) return cudaMemcpyAsync(dst, src, width * (size_t)rows, kind, stm);
        if (dpitch <= (size_t)max_pitch && spitch <= (size_t)max_pitch) return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, (size_t)rows, kind, stm);
        for (int r = 0; r < rows; r++) {
            cudaError_t e = cudaMemcpyAsync(static_cast<uint8_t *>(dst) + (size_t)r * dpitch, static_cast<const uint8_t *>(src) + (size_t)r * spitch,
                                            width, kind, stm);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    // any failure inside the loop: the copies already enqueued still read / write the caller's host buffers -- drain the
    // three streams before handing the buffers back
    auto bail = [&](int code) -> int {
        cudaStreamSynchronize(c->s_copy_in);
        cudaStreamSynchronize(c->s_compute);
        cudaStreamSynchronize(c->s_copy_out);
        return code;
    };
#define LV_PIPE(call)                                                  \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) return bail(cuda_fail(e_, #call));      \
    } while (0)

    int k = 0;
    for (const HostChunk &ch : plan.chunks) {
        {
            const int s0 = ch.s0, ns = ch.ns, t0 = ch.t0, nt = ch.nt;
            const int slot = k % lv_ctx::kSlots;
            // the slot's previous download must have drained before its buffers are reused
            if (k >= lv_ctx::kSlots) {
                LV_PIPE(cudaStreamWaitEvent(c->s_copy_in, c->ev_k[slot], 0));
                LV_PIPE(cudaStreamWaitEvent(c->s_compute, c->ev_out[slot], 0));
            }
            const size_t hf = (size_t)s0 * n_frames + t0;
            if (pic)
                LV_PIPE(move(c->d_pic[slot], (size_t)nt * pic->samples * sizeof(uint16_t), pic->h_pic16 + hf * pic->samples,
                             (size_t)n_frames * pic->samples * sizeof(uint16_t), (size_t)nt * pic->samples * sizeof(uint16_t), ns,
                             cudaMemcpyHostToDevice, c->s_copy_in));
            else
                LV_PIPE(move(c->d_in[slot], (size_t)nt * g.frame, h_i420 + hf * g.frame, (size_t)n_frames * g.frame,
                             (size_t)nt * g.frame, ns, cudaMemcpyHostToDevice, c->s_copy_in));
            LV_PIPE(cudaEventRecord(c->ev_in[slot], c->s_copy_in));
            LV_PIPE(cudaStreamWaitEvent(c->s_compute, c->ev_in[slot], 0));
            // img2buf + write_out_picture on the device: inside the fused kernel when the picture layout allows it (7 B/px
            // of traffic instead of 4.5 + 5.5), else 16-bit planes -> the slot's tight I420 frames -> the fused kernel
            lv::PicLayout pl;
            bool fused = false;
            if (pic) {
                pl.pic16 = c->d_pic[slot]; pl.size_x = pic->size_x; pl.size_y = pic->size_y; pl.crop_left = pic->crop_left;
                pl.crop_top = pic->crop_top; pl.samples = pic->samples;
                fused = h_bgr && lv::pic_fusable(g, pl) && lv::kernel_variant() != 1;
                if (!fused) {
                    rc = launch_result(lv::launch_narrow(g, c->d_pic[slot], ns * nt, pic->size_x, pic->size_y, pic->crop_left,
                                                         pic->crop_top, c->d_in[slot], c->s_compute), c->launches);
                    if (rc != LV_OK) return bail(rc);
                }
            }
            // small outputs: straight into the whole-step arrays when the chunk is one run there (a single stream, or
            // every frame of its streams), else into the slot and placed by a tiny kernel behind the step
            const bool direct = ns == 1 || nt == n_frames;
            uint16_t *k_cnt = !h_mb_count ? nullptr : direct ? c->d_step_cnt + hf * mbs : c->d_cnt[slot];
            uint8_t *k_map = !h_mb_map ? nullptr : direct ? c->d_step_map + hf * mbs : c->d_map[slot];
            uint32_t *k_sum = !h_summary ? nullptr : direct ? c->d_step_sum + hf * 2 : c->d_sum[slot];
            rc = step_impl(c, fused ? nullptr : c->d_in[slot], first_stream + s0, ns, nt, h_bgr ? c->d_bgr[slot] : nullptr,
                           h_mask_bits ? c->d_bits[slot] : nullptr, nullptr, k_cnt, k_map, k_sum, nullptr, 0, nullptr,
                           c->s_compute, 0, fused ? &pl : nullptr);
            if (rc != LV_OK) return bail(rc);
            if (small && !direct) {
                k_place_small<<<ns * nt, 256, 0, c->s_compute>>>(k_cnt, c->d_step_cnt + hf * mbs, k_map, c->d_step_map + hf * mbs, k_sum,
                                                              c->d_step_sum + hf * 2, (int)mbs, nt, n_frames);
                LV_PIPE(cudaGetLastError());
                c->launches++;
            }
            LV_PIPE(cudaEventRecord(c->ev_k[slot], c->s_compute));
            LV_PIPE(cudaStreamWaitEvent(c->s_copy_out, c->ev_k[slot], 0));
            // D2H: the same strided shape per output (width = this chunk's nt frames of one stream, rows = streams)
            auto down = [&](void *h_base, const void *d_base, size_t bytes_per_frame) -> cudaError_t {
                return move(static_cast<uint8_t *>(h_base) + hf * bytes_per_frame, (size_t)n_frames * bytes_per_frame, d_base,
                            (size_t)nt * bytes_per_frame, (size_t)nt * bytes_per_frame, ns, cudaMemcpyDeviceToHost, c->s_copy_out);
            };
            if (h_bgr) LV_PIPE(down(h_bgr, c->d_bgr[slot], 3 * g.px));
            if (h_mask_bits) LV_PIPE(down(h_mask_bits, c->d_bits[slot], sizeof(uint16_t) * units));
            LV_PIPE(cudaEventRecord(c->ev_out[slot], c->s_copy_out));
            k++;
        }
    }
    // the small outputs of the whole step: one copy per array, on the upload stream (idle by now), beside the last
    // picture downloads
    if (small) {
        const size_t frames = (size_t)n_streams * n_frames;
        LV_PIPE(cudaEventRecord(c->ev_small, c->s_compute));
        LV_PIPE(cudaStreamWaitEvent(c->s_copy_in, c->ev_small, 0));
        if (h_mb_count) LV_PIPE(cudaMemcpyAsync(h_mb_count, c->d_step_cnt, sizeof(uint16_t) * mbs * frames, cudaMemcpyDeviceToHost, c->s_copy_in));
        if (h_mb_map) LV_PIPE(cudaMemcpyAsync(h_mb_map, c->d_step_map, mbs * frames, cudaMemcpyDeviceToHost, c->s_copy_in));
        if (h_summary) LV_PIPE(cudaMemcpyAsync(h_summary, c->d_step_sum, sizeof(uint32_t) * 2 * frames, cudaMemcpyDeviceToHost, c->s_copy_in));
    }
    LV_PIPE(cudaStreamSynchronize(c->s_copy_out));
    LV_PIPE(cudaStreamSynchronize(c->s_copy_in));
    LV_PIPE(cudaStreamSynchronize(c->s_compute));
#undef LV_PIPE
    return LV_OK;
}
}  // namespace

int lv_host_step_plan(int width, int height, size_t in_frame_bytes, int n_streams, int n_frames, int *chunks4, int max_chunks)
{
    if (width < 4 || height < 2 || width % 4 != 0 || height % 2 != 0 || n_streams < 0 || n_frames < 0 || max_chunks < 0 ||
        (max_chunks > 0 && !chunks4))
        return LV_E_ARG;
    const size_t frame = (size_t)width * height * 3 / 2;
    const HostPlan p = plan_host_step(in_frame_bytes ? in_frame_bytes : frame, frame, n_streams, n_frames);
    const int n = (int)p.chunks.size();
    for (int i = 0; i < n && i < max_chunks; i++) {
        chunks4[4 * i] = p.chunks[i].s0; chunks4[4 * i + 1] = p.chunks[i].ns;
        chunks4[4 * i + 2] = p.chunks[i].t0; chunks4[4 * i + 3] = p.chunks[i].nt;
    }
    return n;
}

int lv_step_host(lv_ctx *c, const uint8_t *h_i420, int first_stream, int n_streams, int n_frames, uint8_t *h_bgr,
                 uint16_t *h_mask_bits, uint16_t *h_mb_count, uint8_t *h_mb_map, uint32_t *h_summary)
{
    if (!h_i420) return LV_E_ARG;
    return step_host_impl(c, h_i420, nullptr, first_stream, n_streams, n_frames, h_bgr, h_mask_bits, h_mb_count, h_mb_map,
                          h_summary);
}

int lv_step_host_pictures(lv_ctx *c, const uint16_t *h_pic16, int size_x, int size_y, int crop_left, int crop_top,
                          int first_stream, int n_streams, int n_frames, uint8_t *h_bgr, uint16_t *h_mask_bits,
                          uint16_t *h_mb_count, uint8_t *h_mb_map, uint32_t *h_summary)
{
    if (!c || !h_pic16) return LV_E_ARG;
    if (size_x <= 0 || size_y <= 0 || (size_x & 1) || (size_y & 1) || crop_left < 0 || crop_top < 0 || (crop_left & 1) ||
        (crop_top & 1) || crop_left + c->g.w > size_x || crop_top + c->g.h > size_y)
        return LV_E_SIZE;
    PicSource p{h_pic16, size_x, size_y, crop_left, crop_top,
                (size_t)size_x * size_y + 2 * ((size_t)(size_x >> 1) * (size_y >> 1))};
    return step_host_impl(c, nullptr, &p, first_stream, n_streams, n_frames, h_bgr, h_mask_bits, h_mb_count, h_mb_map,
                          h_summary);
}

// ---- legacy drop-in entries ------------------------------------------------------------------------

namespace {

[[noreturn]] void legacy_die(const char *fn, const char *msg)
{
    fprintf(stderr, "livevideo-b200: %s: %s (%s)\n", fn, msg, g_err);
    abort();
}

struct LegacyBuf {
    uint8_t *d = nullptr;
    size_t cap = 0;
    uint8_t *get(size_t n)
    {
        if (n > cap) {
            cudaFree(d);
            d = nullptr;
            cap = 0;
            if (cudaMalloc(&d, n) != cudaSuccess) return nullptr;
            cap = n;
        }
        return d;
    }
};
thread_local LegacyBuf t_in, t_out;
thread_local cudaStream_t t_stream = nullptr;
thread_local int t_device = -1;       // the device this thread's stream and scratch buffers live on

// The per-thread stream and scratch buffers are created on the device that is current at the thread's first legacy
// call and stay there: later calls switch to that device for their duration (the application, or another library in
// the process, may have made a different device current in between).
struct LegacyScope {
    DeviceGuard *guard = nullptr;
    cudaStream_t st = nullptr;
    explicit LegacyScope(const char *fn)
    {
        if (t_device < 0) {
            int d = 0;
            if (cudaGetDevice(&d) != cudaSuccess) {
                cuda_fail(cudaGetLastError(), "cudaGetDevice");
                legacy_die(fn, "no usable CUDA device");
            }
            t_device = d;
        }
        guard = new DeviceGuard(t_device);
        if (!guard->ok) {
            cuda_fail(cudaGetLastError(), "cudaSetDevice");
            legacy_die(fn, "no usable CUDA device");
        }
        if (!t_stream && cudaStreamCreateWithFlags(&t_stream, cudaStreamNonBlocking) != cudaSuccess) {
            cuda_fail(cudaGetLastError(), "cudaStreamCreate");
            legacy_die(fn, "no usable CUDA device");
        }
        st = t_stream;
    }
    ~LegacyScope() { delete guard; }
};

}  // namespace

void YV12_to_RGB24(unsigned char *src0, unsigned char *src1, unsigned char *src2, unsigned char *dst_ori, int width,
                   int height)
{
    static const char *fn = "YV12_to_RGB24";
    if (!src0 || !src1 || !src2 || !dst_ori) legacy_die(fn, "NULL plane");
    if (width <= 0 || height <= 0) return;   // the original's loops run zero times
    if (width % 4 != 0 || height % 2 != 0) legacy_die(fn, "needs width % 4 == 0 and height % 2 == 0 (the original overruns)");
    LegacyScope scope(fn);
    cudaStream_t st = scope.st;
    const size_t px = (size_t)width * height;
    auto_pin(src0, px); auto_pin(src1, px / 4); auto_pin(src2, px / 4); auto_pin(dst_ori, 3 * px);
    if (zero_copy_mode() && geom_fast(width, height)) {
        // all four buffers pinned + mapped in place: one kernel reads the planes and writes the picture over the link
        lv::PlaneTable t{};
        t.y[0] = mapped(src0, px); t.u[0] = mapped(src1, px / 4); t.v[0] = mapped(src2, px / 4); t.dst[0] = mapped(dst_ori, 3 * px);
        if (t.y[0] && t.u[0] && t.v[0] && t.dst[0] && aligned(t.y[0], 16) && aligned(t.dst[0], 16) && aligned(t.u[0], 8) &&
            aligned(t.v[0], 8)) {
            const int r = lv::launch_convert_planes(lv::make_geom(width, height), t, 1, zero_copy_mode() == 2, st);
            if (r < 0) { cuda_fail((cudaError_t)(-r), "launch"); legacy_die(fn, "kernel launch failed"); }
            const cudaError_t ez = cudaStreamSynchronize(st);
            if (ez != cudaSuccess) { cuda_fail(ez, "zero-copy kernel"); legacy_die(fn, "kernel failed"); }
            return;
        }
    }
    uint8_t *din = t_in.get(px + px / 2), *dout = t_out.get(3 * px);
    if (!din || !dout) legacy_die(fn, "device allocation failed");
    cudaError_t e = cudaMemcpyAsync(din, src0, px, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(din + px, src1, px / 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(din + px + px / 4, src2, px / 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cuda_fail(e, "H2D"); legacy_die(fn, "copy failed"); }
    int r = geom_fast(width, height)
                ? lv::launch_convert(lv::make_geom(width, height), din, 1, dout, st)
                : lv::launch_convert_generic(width, height, din, din + px, din + px + px / 4, dout, st);
    if (r < 0) { cuda_fail((cudaError_t)(-r), "launch"); legacy_die(fn, "kernel launch failed"); }
    e = cudaMemcpyAsync(dst_ori, dout, 3 * px, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cuda_fail(e, "D2H"); legacy_die(fn, "copy failed"); }
}

namespace {

void legacy_format(const char *fn, int kind, unsigned char *in, unsigned char *out, int w, int h)
{
    if (!in || !out) legacy_die(fn, "NULL buffer");
    if (w <= 0 || h <= 0) return;
    if (w % 4 != 0 || h % 4 != 0) legacy_die(fn, "needs width % 4 == 0 and height % 4 == 0");
    LegacyScope scope(fn);
    cudaStream_t st = scope.st;
    const size_t inb = lv::format_bytes_in(kind, w, h), outb = lv::format_bytes_out(kind, w, h);
    uint8_t *din = t_in.get(inb), *dout = t_out.get(outb);
    if (!din || !dout) legacy_die(fn, "device allocation failed");
    cudaError_t e = cudaMemcpyAsync(din, in, inb, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cuda_fail(e, "H2D"); legacy_die(fn, "copy failed"); }
    const int r = lv::launch_format(kind, w, h, din, dout, 1, st);
    if (r < 0) { cuda_fail((cudaError_t)(-r), "launch"); legacy_die(fn, "kernel launch failed"); }
    e = cudaMemcpyAsync(out, dout, outb, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cuda_fail(e, "D2H"); legacy_die(fn, "copy failed"); }
}

}  // namespace

void RGB24_to_YV12(unsigned char *in, unsigned char *out, int w, int h) { legacy_format("RGB24_to_YV12", 1, in, out, w, h); }
void YVU9_to_YV12(unsigned char *in, unsigned char *out, int w, int h) { legacy_format("YVU9_to_YV12", 2, in, out, w, h); }
void YUY2_to_YV12(unsigned char *in, unsigned char *out, int w, int h) { legacy_format("YUY2_to_YV12", 3, in, out, w, h); }
void YV12_to_YVU9(unsigned char *in, unsigned char *out, int w, int h) { legacy_format("YV12_to_YVU9", 4, in, out, w, h); }
void YV12_to_YUY2(unsigned char *in, unsigned char *out, int w, int h) { legacy_format("YV12_to_YUY2", 5, in, out, w, h); }

void FrameDiff_MB(const unsigned char *cur_y, const unsigned char *ref_y, int width, int height, int tol, int mb_thresh,
                  unsigned char *mask, unsigned short *mb_count, unsigned char *mb_map)
{
    static const char *fn = "FrameDiff_MB";
    if (!cur_y || !ref_y) legacy_die(fn, "NULL plane");
    if (width <= 0 || height <= 0) return;
    if (tol < 0) tol = -1;          // IsNearlyZero(a, tol<0) is never true except ... nothing: every pixel changed
    if (tol > 255) tol = 255;
    LegacyScope scope(fn);
    cudaStream_t st = scope.st;
    const size_t px = (size_t)width * height;
    const int mb_w = (width + 15) / 16, mb_h = (height + 15) / 16;
    const size_t mbs = (size_t)mb_w * mb_h;
    auto_pin(cur_y, px); auto_pin(ref_y, px);
    if (mask) auto_pin(mask, px);
    if (mb_count) auto_pin(mb_count, 2 * mbs);
    if (mb_map) auto_pin(mb_map, mbs);
    if (zero_copy_mode() && geom_fast(width, height) && tol >= 0 && lv::kernel_variant() != 1) {
        // the planes (and the byte mask) pinned + mapped in place: the difference kernel addresses them directly; the two
        // small per-macroblock outputs go the same way when they are mapped too, else through the device scratch
        lv::DiffArgs a{};
        a.g = lv::make_geom(width, height); a.n_frames = 1; a.tol = tol; a.mb_thresh = mb_thresh;
        a.cur_y = mapped(cur_y, px); a.ref_y = mapped(ref_y, px);
        a.mask_bytes = mask ? mapped(mask, px) : nullptr;
        if (a.cur_y && a.ref_y && aligned(a.cur_y, 16) && aligned(a.ref_y, 16) && (!mask || (a.mask_bytes && aligned(a.mask_bytes, 16)))) {
            uint16_t *zc_cnt = mb_count ? reinterpret_cast<uint16_t *>(mapped(mb_count, 2 * mbs)) : nullptr;
            uint8_t *zc_map = mb_map ? mapped(mb_map, mbs) : nullptr;
            if (zc_cnt && !aligned(zc_cnt, 2)) zc_cnt = nullptr;
            const size_t cnt_bytes = (2 * mbs + 15) & ~(size_t)15;
            uint8_t *scratch = ((mb_count && !zc_cnt) || (mb_map && !zc_map)) ? t_out.get(cnt_bytes + mbs) : nullptr;
            if (((mb_count && !zc_cnt) || (mb_map && !zc_map)) && !scratch) legacy_die(fn, "device allocation failed");
            a.mb_count = mb_count ? (zc_cnt ? zc_cnt : reinterpret_cast<uint16_t *>(scratch)) : nullptr;
            a.mb_map = mb_map ? (zc_map ? zc_map : scratch + cnt_bytes) : nullptr;
            const int r = lv::launch_diff(a, st);
            if (r < 0) { cuda_fail((cudaError_t)(-r), "launch"); legacy_die(fn, "kernel launch failed"); }
            cudaError_t ez = cudaSuccess;
            if (mb_count && !zc_cnt) ez = cudaMemcpyAsync(mb_count, scratch, 2 * mbs, cudaMemcpyDeviceToHost, st);
            if (mb_map && !zc_map && ez == cudaSuccess) ez = cudaMemcpyAsync(mb_map, scratch + cnt_bytes, mbs, cudaMemcpyDeviceToHost, st);
            if (ez == cudaSuccess) ez = cudaStreamSynchronize(st);
            if (ez != cudaSuccess) { cuda_fail(ez, "zero-copy kernel"); legacy_die(fn, "kernel failed"); }
            return;
        }
    }
    // layout of the output scratch: mask bytes | counts (u16, 2-byte aligned) | map
    const size_t off_cnt = (px + 15) & ~(size_t)15, off_map = off_cnt + ((2 * mbs + 15) & ~(size_t)15);
    uint8_t *din = t_in.get(2 * ((px + 15) & ~(size_t)15)), *dout = t_out.get(off_map + mbs);
    if (!din || !dout) legacy_die(fn, "device allocation failed");
    uint8_t *dref = din + ((px + 15) & ~(size_t)15);
    cudaError_t e = cudaMemcpyAsync(din, cur_y, px, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dref, ref_y, px, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cuda_fail(e, "H2D"); legacy_die(fn, "copy failed"); }
    uint8_t *dmask = mask ? dout : nullptr;
    uint16_t *dcnt = reinterpret_cast<uint16_t *>(dout + off_cnt);
    uint8_t *dmap = dout + off_map;
    int r;
    if (geom_fast(width, height) && tol >= 0) {
        lv::DiffArgs a{};
        a.g = lv::make_geom(width, height); a.cur_y = din; a.ref_y = dref; a.n_frames = 1;
        a.tol = tol; a.mb_thresh = mb_thresh; a.mask_bytes = dmask; a.mb_count = dcnt; a.mb_map = dmap;
        r = lv::launch_diff(a, st);
    } else {
        r = lv::launch_diff_generic(width, height, din, dref, tol, mb_thresh, dmask, dcnt, dmap, st);
    }
    if (r < 0) { cuda_fail((cudaError_t)(-r), "launch"); legacy_die(fn, "kernel launch failed"); }
    if (mask && e == cudaSuccess) e = cudaMemcpyAsync(mask, dmask, px, cudaMemcpyDeviceToHost, st);
    if (mb_count && e == cudaSuccess) e = cudaMemcpyAsync(mb_count, dcnt, 2 * mbs, cudaMemcpyDeviceToHost, st);
    if (mb_map && e == cudaSuccess) e = cudaMemcpyAsync(mb_map, dmap, mbs, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cuda_fail(e, "D2H"); legacy_die(fn, "copy failed"); }
}

// ---- several pictures per call: the reference's real call pattern -------------------------------------------------
// CMainDlg::DisplayPic converts FIVE pictures back to back (MainDlg.cpp:946-974: srcY[i], srcU[i], srcV[i] -> RGBbuf[i]).
// One synchronous call per picture serialises H2D -> kernel -> D2H five times; here the upload of picture i+1, the
// kernel of picture i and the download of picture i-1 run on three streams and the host waits once, at the end.
}  // extern "C"

namespace {

struct MultiPipe {
    static const int kSlots = 3;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[kSlots] = {}, ev_k[kSlots] = {}, ev_out[kSlots] = {};
    LegacyBuf in[kSlots], out[kSlots];
    bool ready = false;
    int init()
    {
        if (ready) return LV_OK;
        LV_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        LV_CUDA(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
        LV_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        for (int i = 0; i < kSlots; i++) {
            LV_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
            LV_CUDA(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
            LV_CUDA(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
        }
        ready = true;
        return LV_OK;
    }
    int drain(int code)
    {
        if (s_in) cudaStreamSynchronize(s_in);
        if (s_k) cudaStreamSynchronize(s_k);
        if (s_out) cudaStreamSynchronize(s_out);
        return code;
    }
};
thread_local MultiPipe t_pipe;

// the per-thread device of the legacy entries, without the abort-on-failure of LegacyScope
int multi_scope(DeviceGuard *&guard)
{
    if (t_device < 0) {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess) {
            cuda_fail(cudaGetLastError(), "cudaGetDevice");
            return LV_E_NODEVICE;
        }
        t_device = d;
    }
    guard = new (std::nothrow) DeviceGuard(t_device);
    if (!guard) return LV_E_NOMEM;
    if (!guard->ok) {
        cuda_fail(cudaGetLastError(), "cudaSetDevice");
        return LV_E_NODEVICE;
    }
    return t_pipe.init();
}
struct GuardOwner {
    DeviceGuard *g = nullptr;
    ~GuardOwner() { delete g; }
};

#define LV_MP(call)                                                            \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) return P.drain(cuda_fail(e_, #call));           \
    } while (0)

}  // namespace

extern "C" {

int lv_convert_host_multi(const unsigned char *const *src0, const unsigned char *const *src1, const unsigned char *const *src2,
                          unsigned char *const *dst_ori, int n, int width, int height)
{
    if (n < 0 || (n > 0 && (!src0 || !src1 || !src2 || !dst_ori))) return LV_E_ARG;
    for (int i = 0; i < n; i++)
        if (!src0[i] || !src1[i] || !src2[i] || !dst_ori[i]) return LV_E_ARG;
    if (n == 0 || width <= 0 || height <= 0) return LV_OK;
    if (width % 4 != 0 || height % 2 != 0) return LV_E_SIZE;
    GuardOwner go;
    int rc = multi_scope(go.g);
    if (rc != LV_OK) return rc;
    MultiPipe &P = t_pipe;
    const size_t px = (size_t)width * height;
    const bool fast = geom_fast(width, height);
    const lv::Geom g = lv::make_geom(width, height);
    if (zero_copy_mode() && fast) {
        // every picture's four buffers pinned + mapped in place: up to 8 pictures per launch straight over the link
        bool all = true;
        for (int i = 0; i < n && all; i++) {
            const uint8_t *y = mapped(src0[i], px), *u = mapped(src1[i], px / 4), *v = mapped(src2[i], px / 4), *d = mapped(dst_ori[i], 3 * px);
            all = y && u && v && d && aligned(y, 16) && aligned(d, 16) && aligned(u, 8) && aligned(v, 8);
        }
        if (all) {
            for (int i0 = 0; i0 < n; i0 += lv::PlaneTable::kMax) {
                const int m = n - i0 < lv::PlaneTable::kMax ? n - i0 : lv::PlaneTable::kMax;
                lv::PlaneTable t{};
                for (int j = 0; j < m; j++) {
                    t.y[j] = mapped(src0[i0 + j], px); t.u[j] = mapped(src1[i0 + j], px / 4); t.v[j] = mapped(src2[i0 + j], px / 4);
                    t.dst[j] = mapped(dst_ori[i0 + j], 3 * px);
                }
                const int r = lv::launch_convert_planes(g, t, m, zero_copy_mode() == 2, P.s_k);
                if (r < 0) return P.drain(cuda_fail((cudaError_t)(-r), "kernel launch"));
            }
            LV_MP(cudaStreamSynchronize(P.s_k));
            return LV_OK;
        }
    }
    for (int i = 0; i < n; i++) {
        const int slot = i % MultiPipe::kSlots;
        uint8_t *din = P.in[slot].get(px + px / 2), *dout = P.out[slot].get(3 * px);
        if (!din || !dout) return P.drain(LV_E_NOMEM);
        if (i >= MultiPipe::kSlots) {
            LV_MP(cudaStreamWaitEvent(P.s_in, P.ev_k[slot], 0));     // the kernel that read this slot's input
            LV_MP(cudaStreamWaitEvent(P.s_k, P.ev_out[slot], 0));    // the download of this slot's output
        }
        LV_MP(cudaMemcpyAsync(din, src0[i], px, cudaMemcpyHostToDevice, P.s_in));
        LV_MP(cudaMemcpyAsync(din + px, src1[i], px / 4, cudaMemcpyHostToDevice, P.s_in));
        LV_MP(cudaMemcpyAsync(din + px + px / 4, src2[i], px / 4, cudaMemcpyHostToDevice, P.s_in));
        LV_MP(cudaEventRecord(P.ev_in[slot], P.s_in));
        LV_MP(cudaStreamWaitEvent(P.s_k, P.ev_in[slot], 0));
        const int r = fast ? lv::launch_convert(g, din, 1, dout, P.s_k)
                           : lv::launch_convert_generic(width, height, din, din + px, din + px + px / 4, dout, P.s_k);
        if (r < 0) return P.drain(cuda_fail((cudaError_t)(-r), "kernel launch"));
        LV_MP(cudaEventRecord(P.ev_k[slot], P.s_k));
        LV_MP(cudaStreamWaitEvent(P.s_out, P.ev_k[slot], 0));
        LV_MP(cudaMemcpyAsync(dst_ori[i], dout, 3 * px, cudaMemcpyDeviceToHost, P.s_out));
        LV_MP(cudaEventRecord(P.ev_out[slot], P.s_out));
    }
    LV_MP(cudaStreamSynchronize(P.s_out));
    return LV_OK;
}

int lv_diff_host_multi(const unsigned char *const *cur_y, const unsigned char *const *ref_y, int n, int width, int height, int tol,
                       int mb_thresh, unsigned char *const *mask, unsigned short *const *mb_count, unsigned char *const *mb_map)
{
    if (n < 0 || (n > 0 && (!cur_y || !ref_y))) return LV_E_ARG;
    for (int i = 0; i < n; i++)
        if (!cur_y[i] || !ref_y[i]) return LV_E_ARG;
    if (n == 0 || width <= 0 || height <= 0) return LV_OK;
    if (tol < 0) tol = -1;
    if (tol > 255) tol = 255;
    GuardOwner go;
    int rc = multi_scope(go.g);
    if (rc != LV_OK) return rc;
    MultiPipe &P = t_pipe;
    const size_t px = (size_t)width * height, pxa = (px + 15) & ~(size_t)15;
    const int mb_w = (width + 15) / 16, mb_h = (height + 15) / 16;
    const size_t mbs = (size_t)mb_w * mb_h;
    const size_t off_cnt = pxa, off_map = off_cnt + ((2 * mbs + 15) & ~(size_t)15);
    const bool fast = geom_fast(width, height) && tol >= 0;
    const lv::Geom g = lv::make_geom(width, height);
    if (zero_copy_mode() && fast && lv::kernel_variant() != 1) {
        // every buffer of every picture pinned + mapped in place: the difference kernel addresses them directly, one launch
        // per picture back to back on one stream, one synchronise
        bool all = true;
        for (int i = 0; i < n && all; i++) {
            const uint8_t *c0 = mapped(cur_y[i], px), *r0 = mapped(ref_y[i], px);
            all = c0 && r0 && aligned(c0, 16) && aligned(r0, 16);
            if (all && mask && mask[i]) { const uint8_t *m0 = mapped(mask[i], px); all = m0 && aligned(m0, 16); }
            if (all && mb_count && mb_count[i]) { const uint8_t *q = mapped(mb_count[i], 2 * mbs); all = q && aligned(q, 2); }
            if (all && mb_map && mb_map[i]) all = mapped(mb_map[i], mbs) != nullptr;
        }
        if (all) {
            for (int i = 0; i < n; i++) {
                lv::DiffArgs a{};
                a.g = g; a.n_frames = 1; a.tol = tol; a.mb_thresh = mb_thresh;
                a.cur_y = mapped(cur_y[i], px); a.ref_y = mapped(ref_y[i], px);
                a.mask_bytes = (mask && mask[i]) ? mapped(mask[i], px) : nullptr;
                a.mb_count = (mb_count && mb_count[i]) ? reinterpret_cast<uint16_t *>(mapped(mb_count[i], 2 * mbs)) : nullptr;
                a.mb_map = (mb_map && mb_map[i]) ? mapped(mb_map[i], mbs) : nullptr;
                const int r = lv::launch_diff(a, P.s_k);
                if (r < 0) return P.drain(cuda_fail((cudaError_t)(-r), "kernel launch"));
            }
            LV_MP(cudaStreamSynchronize(P.s_k));
            return LV_OK;
        }
    }
    for (int i = 0; i < n; i++) {
        const int slot = i % MultiPipe::kSlots, pslot = (i + MultiPipe::kSlots - 1) % MultiPipe::kSlots;
        uint8_t *din = P.in[slot].get(2 * pxa), *dout = P.out[slot].get(off_map + mbs);
        if (!din || !dout) return P.drain(LV_E_NOMEM);
        // a sequence differenced against its own previous picture (ref_y[i] == cur_y[i-1], the prev_frm <- curr_frm of
        // lib/JM/image.c:1553-1561) uploads every plane once: the previous slot's current plane is this one's reference
        const bool chained = i > 0 && ref_y[i] == cur_y[i - 1];
        if (i >= MultiPipe::kSlots) {
            LV_MP(cudaStreamWaitEvent(P.s_in, P.ev_k[slot], 0));
            LV_MP(cudaStreamWaitEvent(P.s_k, P.ev_out[slot], 0));
        }
        if (i >= MultiPipe::kSlots - 1)       // picture i-2's kernel may have read this slot's current plane as ITS reference
            LV_MP(cudaStreamWaitEvent(P.s_in, P.ev_k[(i + 1) % MultiPipe::kSlots], 0));
        LV_MP(cudaMemcpyAsync(din, cur_y[i], px, cudaMemcpyHostToDevice, P.s_in));
        const uint8_t *dref = chained ? P.in[pslot].d : din + pxa;
        if (!chained) LV_MP(cudaMemcpyAsync(din + pxa, ref_y[i], px, cudaMemcpyHostToDevice, P.s_in));
        LV_MP(cudaEventRecord(P.ev_in[slot], P.s_in));
        LV_MP(cudaStreamWaitEvent(P.s_k, P.ev_in[slot], 0));
        uint8_t *m = mask ? mask[i] : nullptr;
        unsigned short *hc = mb_count ? mb_count[i] : nullptr;
        uint8_t *hm = mb_map ? mb_map[i] : nullptr;
        uint8_t *dmask = m ? dout : nullptr;
        uint16_t *dcnt = reinterpret_cast<uint16_t *>(dout + off_cnt);
        uint8_t *dmap = dout + off_map;
        int r;
        if (fast) {
            lv::DiffArgs a{};
            a.g = g; a.cur_y = din; a.ref_y = dref; a.n_frames = 1;
            a.tol = tol; a.mb_thresh = mb_thresh; a.mask_bytes = dmask; a.mb_count = dcnt; a.mb_map = dmap;
            r = lv::launch_diff(a, P.s_k);
        } else {
            r = lv::launch_diff_generic(width, height, din, dref, tol, mb_thresh, dmask, dcnt, dmap, P.s_k);
        }
        if (r < 0) return P.drain(cuda_fail((cudaError_t)(-r), "kernel launch"));
        LV_MP(cudaEventRecord(P.ev_k[slot], P.s_k));
        LV_MP(cudaStreamWaitEvent(P.s_out, P.ev_k[slot], 0));
        if (m) LV_MP(cudaMemcpyAsync(m, dmask, px, cudaMemcpyDeviceToHost, P.s_out));
        if (hc) LV_MP(cudaMemcpyAsync(hc, dcnt, 2 * mbs, cudaMemcpyDeviceToHost, P.s_out));
        if (hm) LV_MP(cudaMemcpyAsync(hm, dmap, mbs, cudaMemcpyDeviceToHost, P.s_out));
        LV_MP(cudaEventRecord(P.ev_out[slot], P.s_out));
    }
    LV_MP(cudaStreamSynchronize(P.s_out));
    return LV_OK;
}

}  // extern "C"
