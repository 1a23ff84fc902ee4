// lv_shard.cu -- stream-ID sharding over the GPUs of one box from ONE host process (C ABI: lv_shard_*).
//
// Camera streams are independent units -- a stream's only state is its own previous luma
// (lib/JM/image.c:1553-1561) and no pixel crosses streams -- so device g simply owns the contiguous block of
// streams [g*S/G, (g+1)*S/G) and runs the same fused kernel on it.  lv_shard_step only
// enqueues work (one launch per device on that device's stream), so all GPUs run concurrently.  The single
// exchange the path has, the per-(stream, frame) change summaries {changed_px, changed_mbs}, is an NCCL
// all-gather (ncclGroupStart / ncclAllGather per device / ncclGroupEnd): every device ends up with the whole
// table.  NCCL is bound at run time (dlopen of libnccl.so.2) so that the library itself has no link-time
// dependency on it; with one device no collective is issued at all.
//
// The Python / torchrun deployment (one process per GPU, livevideo_b200/shard.py) is the same partition with
// torch.distributed doing the gather.
#include "../../include/livevideo.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

namespace {

// ---- the few NCCL entry points we need, resolved at run time -------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;                        // ncclSuccess == 0
typedef int ncclDataType_t;                      // ncclUint32 == 3 (nccl.h: ncclInt8 0, ncclUint8 1, ncclInt32 2, ncclUint32 3)
struct Nccl {
    void *h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (h) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
        GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd;
    }
};
Nccl g_nccl;

struct Rank {
    int device = 0;
    lv_ctx *ctx = nullptr;
    cudaStream_t stream = nullptr;
    int first = 0, count = 0;          // this rank's block of streams
    uint32_t *d_local = nullptr;       // summaries of the last step, padded to cap streams
    uint32_t *d_all = nullptr;         // gathered table: n_dev x cap x frames x 2
    ncclComm_t comm = nullptr;
};

}  // namespace

struct lv_shard {
    int n_dev = 0, total_streams = 0, cap = 0;   // cap = largest block
    int max_frames = 0, last_frames = 0;
    std::vector<Rank> r;
    bool have_nccl = false;
};

namespace {

void shard_block(int total, int world, int rank, int *first, int *count)
{
    const int a = (int)((long long)rank * total / world), b = (int)((long long)(rank + 1) * total / world);
    *first = a;
    *count = b - a;
}

int ensure_summary_buffers(lv_shard *s, int n_frames)
{
    if (n_frames <= s->max_frames) return LV_OK;
    for (Rank &k : s->r) {
        if (cudaSetDevice(k.device) != cudaSuccess) return LV_E_CUDA;
        cudaFree(k.d_local);
        cudaFree(k.d_all);
        k.d_local = k.d_all = nullptr;
        const size_t words = (size_t)s->cap * n_frames * 2;
        if (cudaMalloc(&k.d_local, words * sizeof(uint32_t)) != cudaSuccess) return LV_E_NOMEM;
        if (cudaMalloc(&k.d_all, words * s->n_dev * sizeof(uint32_t)) != cudaSuccess) return LV_E_NOMEM;
        if (cudaMemset(k.d_local, 0, words * sizeof(uint32_t)) != cudaSuccess) return LV_E_CUDA;
    }
    s->max_frames = n_frames;
    return LV_OK;
}

}  // namespace

extern "C" {

int lv_shard_create(int n_devices, const int *devices, int width, int height, int total_streams, int tol, int mb_thresh,
                    lv_shard **out)
{
    if (!out) return LV_E_ARG;
    *out = nullptr;
    if (n_devices < 1 || total_streams < 1) return LV_E_ARG;
    int avail = 0;
    if (cudaGetDeviceCount(&avail) != cudaSuccess || avail == 0) return LV_E_NODEVICE;
    lv_shard *s = new (std::nothrow) lv_shard;
    if (!s) return LV_E_NOMEM;
    s->n_dev = n_devices;
    s->total_streams = total_streams;
    s->r.resize(n_devices);
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = LV_OK;
    std::vector<int> devs(n_devices);
    for (int i = 0; i < n_devices && rc == LV_OK; i++) {
        Rank &k = s->r[i];
        k.device = devices ? devices[i] : i;
        devs[i] = k.device;
        if (k.device < 0 || k.device >= avail) { rc = LV_E_NODEVICE; break; }
        shard_block(total_streams, n_devices, i, &k.first, &k.count);
        if (k.count > s->cap) s->cap = k.count;
        rc = lv_create(k.device, width, height, k.count > 0 ? k.count : 1, tol, mb_thresh, &k.ctx);
        if (rc != LV_OK) break;
        if (cudaSetDevice(k.device) != cudaSuccess || cudaStreamCreateWithFlags(&k.stream, cudaStreamNonBlocking) != cudaSuccess)
            rc = LV_E_CUDA;
    }
    if (rc == LV_OK && n_devices > 1) {
        // one communicator over the devices of this process; duplicates cannot form a communicator
        if (!g_nccl.load()) rc = LV_E_STATE;
        else {
            std::vector<ncclComm_t> comms(n_devices);
            if (g_nccl.CommInitAll(comms.data(), n_devices, devs.data()) != 0) rc = LV_E_STATE;
            else {
                for (int i = 0; i < n_devices; i++) s->r[i].comm = comms[i];
                s->have_nccl = true;
            }
        }
    }
    cudaSetDevice(prev);
    if (rc != LV_OK) {
        lv_shard_destroy(s);
        return rc;
    }
    *out = s;
    return LV_OK;
}

void lv_shard_destroy(lv_shard *s)
{
    if (!s) return;
    int prev = 0;
    cudaGetDevice(&prev);
    for (Rank &k : s->r) {
        cudaSetDevice(k.device);
        if (k.stream) cudaStreamSynchronize(k.stream);
        if (k.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(k.comm);
        cudaFree(k.d_local);
        cudaFree(k.d_all);
        if (k.stream) cudaStreamDestroy(k.stream);
        lv_destroy(k.ctx);
    }
    cudaSetDevice(prev);
    delete s;
}

int lv_shard_device_count(const lv_shard *s) { return s ? s->n_dev : 0; }

int lv_shard_range(const lv_shard *s, int rank, int *device, int *first_stream, int *n_streams)
{
    if (!s || rank < 0 || rank >= s->n_dev) return LV_E_ARG;
    if (device) *device = s->r[rank].device;
    if (first_stream) *first_stream = s->r[rank].first;
    if (n_streams) *n_streams = s->r[rank].count;
    return LV_OK;
}

lv_ctx *lv_shard_ctx(lv_shard *s, int rank) { return (s && rank >= 0 && rank < s->n_dev) ? s->r[rank].ctx : nullptr; }

void *lv_shard_stream(lv_shard *s, int rank) { return (s && rank >= 0 && rank < s->n_dev) ? (void *)s->r[rank].stream : nullptr; }

int lv_shard_step(lv_shard *s, const uint8_t *const *d_i420, int n_frames, uint8_t *const *d_bgr,
                  uint16_t *const *d_mb_count, uint8_t *const *d_mb_map)
{
    if (!s || !d_i420 || n_frames < 1) return LV_E_ARG;
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = ensure_summary_buffers(s, n_frames);      // switches devices: restore the caller's below
    if (rc != LV_OK) {
        cudaSetDevice(prev);
        return rc;
    }
    for (int i = 0; i < s->n_dev && rc == LV_OK; i++) {
        Rank &k = s->r[i];
        if (k.count == 0) continue;
        if (!d_i420[i]) { rc = LV_E_ARG; break; }
        // asynchronous: returns as soon as the launch is enqueued, so the devices run concurrently
        rc = lv_convert_diff_batch(k.ctx, d_i420[i], 0, k.count, n_frames, d_bgr ? d_bgr[i] : nullptr, nullptr, nullptr,
                                   d_mb_count ? d_mb_count[i] : nullptr, d_mb_map ? d_mb_map[i] : nullptr, k.d_local, k.stream);
    }
    cudaSetDevice(prev);
    if (rc == LV_OK) s->last_frames = n_frames;
    return rc;
}

int lv_gather_summaries(lv_shard *s, uint32_t *h_summary)
{
    if (!s || s->last_frames < 1) return s ? LV_E_STATE : LV_E_ARG;
    const int T = s->last_frames;
    const size_t words = (size_t)s->cap * T * 2;            // per-rank (padded) block
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = LV_OK;
    if (s->n_dev > 1) {
        if (!s->have_nccl) { cudaSetDevice(prev); return LV_E_STATE; }
        // the one collective of the path: every device receives every block (<= 64 KiB per rank at 256 streams x 16)
        if (g_nccl.GroupStart() != 0) rc = LV_E_STATE;
        for (int i = 0; i < s->n_dev && rc == LV_OK; i++) {
            Rank &k = s->r[i];
            if (cudaSetDevice(k.device) != cudaSuccess) { rc = LV_E_CUDA; break; }
            if (g_nccl.AllGather(k.d_local, k.d_all, words, /*ncclUint32*/ 3, k.comm, k.stream) != 0) rc = LV_E_STATE;
        }
        if (g_nccl.GroupEnd() != 0 && rc == LV_OK) rc = LV_E_STATE;
    }
    if (rc == LV_OK && h_summary) {
        Rank &k0 = s->r[0];
        if (cudaSetDevice(k0.device) != cudaSuccess) rc = LV_E_CUDA;
        // unpad: rank i's block holds count_i streams
        const uint32_t *src = s->n_dev > 1 ? k0.d_all : k0.d_local;
        for (int i = 0; i < s->n_dev && rc == LV_OK; i++) {
            const Rank &k = s->r[i];
            if (k.count == 0) continue;
            if (cudaMemcpyAsync(h_summary + (size_t)k.first * T * 2, src + (size_t)i * words,
                                (size_t)k.count * T * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, k0.stream) != cudaSuccess)
                rc = LV_E_CUDA;
        }
        if (rc == LV_OK && cudaStreamSynchronize(k0.stream) != cudaSuccess) rc = LV_E_CUDA;
    }
    cudaSetDevice(prev);
    return rc;
}

const uint32_t *lv_shard_gathered(lv_shard *s, int rank)
{
    if (!s || rank < 0 || rank >= s->n_dev) return nullptr;
    return s->n_dev > 1 ? s->r[rank].d_all : s->r[rank].d_local;
}

int lv_shard_sync(lv_shard *s)
{
    if (!s) return LV_E_ARG;
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = LV_OK;
    for (Rank &k : s->r) {
        if (cudaSetDevice(k.device) != cudaSuccess || cudaStreamSynchronize(k.stream) != cudaSuccess) rc = LV_E_CUDA;
    }
    cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
