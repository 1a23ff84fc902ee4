// lv_exchange.cu -- peer-memory exchange of the per-(stream, frame) change summaries (C ABI: lv_exchange_*).
//
// The path shards by stream ID with no data-path collective; the one exchange it has is the few
// hundred bytes of {changed_px, changed_mbs} per step that every rank wants from every other rank.  A library
// all-gather of that size is pure latency (an 8-rank NCCL all-gather costs more than the whole 114 us kernel) and
// its kernels handshake across all GPUs every step.  Here each rank instead PUSHES its block straight into every
// peer's table with NVLink stores from one tiny kernel behind the fused kernel, followed by a sequence flag:
//
//   symmetric buffer (one per rank, same layout, peers map each other's through CUDA IPC or any shared mapping)
//     data [depth][world][slot_words]   u32   table of step k lives in ring slot k % depth; row r is rank r's block
//     flag [depth][world]               u32   flag[slot][r] == k+1  <=>  rank r's block of step k has landed here
//     ack  [world]                      u32   ack[r] == number of steps rank r has collected (retired)
//     err                               u32   set when a wait gave up (never spins forever: a hung GPU is worse)
//
//   publish(k): wait until every rank has retired step k-depth (ack >= k-depth+1; free in steady state), copy the
//               local block (the fused kernel wrote it directly into data[slot][rank] of the LOCAL buffer, so there is
//               no staging copy) to data[slot][rank] of every peer, fence, then store flag[slot][rank] = k+1 at every
//               peer (and locally).
//   collect(k): wait for flag[slot][r] == k+1 for all r (local memory, peers wrote it), copy data[slot] to the
//               caller's table, then store ack[rank] = k+1 at every peer (and locally).
//
// No rank ever waits for a peer unless it is a whole ring (depth steps) ahead: there is no per-step rendezvous.
#include "../../include/livevideo.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>

namespace {

constexpr int kMaxWorld = 32;               // one warp per peer in k_publish (<= 1024 threads)
constexpr unsigned long long kSpinLimitNs = 4000000000ull;   // a wait gives up after 4 s and raises err

thread_local char x_err[256] = "";

int x_fail(cudaError_t e, const char *what)
{
    snprintf(x_err, sizeof x_err, "%s: %s (%d)", what, cudaGetErrorString(e), (int)e);
    return e == cudaErrorMemoryAllocation ? LV_E_NOMEM : LV_E_CUDA;
}
#define X_CUDA(call)                                               \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return x_fail(e_, #call);           \
    } while (0)

struct Layout {
    int world, depth, slot_words;
    __host__ __device__ size_t data_words() const { return (size_t)depth * world * slot_words; }
    __host__ __device__ size_t flag_off() const { return data_words(); }                        // in u32 words
    __host__ __device__ size_t ack_off() const { return flag_off() + (size_t)depth * world; }
    __host__ __device__ size_t err_off() const { return ack_off() + (size_t)world; }
    __host__ __device__ size_t words() const { return (err_off() + 1 + 3) & ~(size_t)3; }
};

struct Peers {
    uint32_t *p[kMaxWorld];
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// spin until *p (local memory, written by a peer) is >= want (sequence numbers only grow); false on timeout
__device__ bool wait_ge(const uint32_t *p, uint32_t want)
{
    if ((int32_t)(ld_acquire_sys(p) - want) >= 0) return true;
    const unsigned long long t0 = now_ns();
    for (;;) {
        if ((int32_t)(ld_acquire_sys(p) - want) >= 0) return true;
        __nanosleep(200);
        if (now_ns() - t0 > kSpinLimitNs) return false;
    }
}

// one block of 32 * world threads (warp r serves peer r); n_words <= slot_words
__global__ void k_publish(Peers peers, Layout L, int rank, uint32_t step, int n_words, uint32_t *local_copy)
{
    uint32_t *local = peers.p[rank];
    const int slot = (int)(step % (uint32_t)L.depth);
    const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    // flow control: rank r must have retired the step that last used this ring slot
    if (step >= (uint32_t)L.depth && lane == 0) {
        if (!wait_ge(local + L.ack_off() + r, step - (uint32_t)L.depth + 1u)) ok = 0;
    }
    __syncthreads();
    if (!ok) {
        if (threadIdx.x == 0) atomicExch(local + L.err_off(), 1u);
        return;
    }
    const size_t row = ((size_t)slot * L.world + rank) * L.slot_words;
    if (r != rank) {
        const uint32_t *src = local + row;
        uint32_t *dst = peers.p[r] + row;
        if (((n_words | L.slot_words) & 3) == 0) {
            for (int i = lane; i < n_words / 4; i += 32)
                reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
        } else {
            for (int i = lane; i < n_words; i += 32) dst[i] = src[i];
        }
    } else if (local_copy) {      // the caller's own summary buffer keeps its usual content
        for (int i = lane; i < n_words; i += 32) local_copy[i] = local[row + i];
    }
    __threadfence_system();
    __syncwarp();
    if (lane == 0) st_release_sys(peers.p[r] + L.flag_off() + (size_t)slot * L.world + rank, step + 1u);
}

// one block of 256 threads; out = world rows of slot_words
__global__ void k_collect(Peers peers, Layout L, int rank, uint32_t step, uint32_t *out)
{
    uint32_t *local = peers.p[rank];
    const int slot = (int)(step % (uint32_t)L.depth);
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < L.world) {
        if (!wait_ge(local + L.flag_off() + (size_t)slot * L.world + threadIdx.x, step + 1u)) ok = 0;
    }
    __syncthreads();
    if (!ok) {
        if (threadIdx.x == 0) atomicExch(local + L.err_off(), 2u);
        return;
    }
    const size_t n = (size_t)L.world * L.slot_words;
    const uint32_t *src = local + (size_t)slot * n;
    if (out) {
        if ((L.slot_words & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            for (size_t i = threadIdx.x; i < n / 4; i += blockDim.x)
                reinterpret_cast<uint4 *>(out)[i] = __ldcv(reinterpret_cast<const uint4 *>(src) + i);
        } else {
            for (size_t i = threadIdx.x; i < n; i += blockDim.x) out[i] = __ldcv(src + i);
        }
    }
    __syncthreads();
    // retire: the slot may be overwritten once every rank has said so
    if (threadIdx.x < L.world) st_release_sys(peers.p[threadIdx.x] + L.ack_off() + rank, step + 1u);
}

}  // namespace

struct lv_exchange {
    int device = 0, rank = 0, world = 1;
    Layout L{};
    uint32_t *local = nullptr;       // cudaMalloc'ed, IPC-exportable
    Peers peers{};
    bool opened[kMaxWorld] = {};     // peers.p[r] came from cudaIpcOpenMemHandle
    bool connected = false;
    uint64_t published = 0, collected = 0;
    // lv_exchange_step: the side stream the pushes / collects run on, ring of events and of collected tables
    cudaStream_t side = nullptr;
    cudaEvent_t *ev_kernel = nullptr, *ev_collected = nullptr;   // depth each
    uint32_t *tables = nullptr;                                  // depth x world x slot_words
    uint64_t steps = 0;                                          // next step of lv_exchange_step
    int lag = 2;                                                 // step k's table is collected when step k+lag is issued
};

namespace {
struct DevGuard {
    int prev = -1;
    explicit DevGuard(int d)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        cudaSetDevice(d);
    }
    ~DevGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
}  // namespace

extern "C" {

const char *lv_exchange_last_error(void) { return x_err; }

size_t lv_exchange_bytes(int world, int depth, int slot_words)
{
    if (world < 1 || world > kMaxWorld || depth < 2 || slot_words < 1) return 0;
    Layout L{world, depth, slot_words};
    return L.words() * sizeof(uint32_t);
}

int lv_exchange_create(int device, int rank, int world, int depth, int slot_words, lv_exchange **out)
{
    if (!out) return LV_E_ARG;
    *out = nullptr;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || depth < 2 || depth > 1024 || slot_words < 1)
        return LV_E_ARG;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        cudaGetLastError();
        return LV_E_NODEVICE;
    }
    DevGuard g(device);
    lv_exchange *x = new (std::nothrow) lv_exchange;
    if (!x) return LV_E_NOMEM;
    x->device = device; x->rank = rank; x->world = world;
    x->L = Layout{world, depth, slot_words};
    const size_t bytes = x->L.words() * sizeof(uint32_t);
    cudaError_t e = cudaMalloc(&x->local, bytes);
    if (e == cudaSuccess) e = cudaMemset(x->local, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        int rc = x_fail(e, "exchange buffer");
        cudaFree(x->local);
        delete x;
        return rc;
    }
    x->peers.p[rank] = x->local;
    if (world == 1) x->connected = true;
    *out = x;
    return LV_OK;
}

void lv_exchange_destroy(lv_exchange *x)
{
    if (!x) return;
    DevGuard g(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; r++)
        if (x->opened[r]) cudaIpcCloseMemHandle(x->peers.p[r]);
    if (x->ev_kernel)
        for (int i = 0; i < x->L.depth; i++) {
            if (x->ev_kernel[i]) cudaEventDestroy(x->ev_kernel[i]);
            if (x->ev_collected[i]) cudaEventDestroy(x->ev_collected[i]);
        }
    delete[] x->ev_kernel;
    delete[] x->ev_collected;
    if (x->side) cudaStreamDestroy(x->side);
    cudaFree(x->tables);
    cudaFree(x->local);
    delete x;
}

void *lv_exchange_buffer(lv_exchange *x) { return x ? x->local : nullptr; }

/* 64-byte CUDA IPC handle of the local buffer, to be handed to the other ranks by any means */
int lv_exchange_handle(lv_exchange *x, unsigned char *handle64)
{
    if (!x || !handle64) return LV_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DevGuard g(x->device);
    cudaIpcMemHandle_t h;
    X_CUDA(cudaIpcGetMemHandle(&h, x->local));
    memcpy(handle64, &h, 64);
    return LV_OK;
}

/* handles: world * 64 bytes, entry r = rank r's lv_exchange_handle (entry `rank` is ignored) */
int lv_exchange_connect(lv_exchange *x, const unsigned char *handles)
{
    if (!x || !handles) return LV_E_ARG;
    if (x->connected) return LV_E_STATE;
    DevGuard g(x->device);
    for (int r = 0; r < x->world; r++) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)r, 64);
        void *p = nullptr;
        X_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->peers.p[r] = (uint32_t *)p;
        x->opened[r] = true;
    }
    x->connected = true;
    return LV_OK;
}

/* same, from raw device pointers that are already valid in this process (single-process multi-rank tests, a symmetric
 * allocation made elsewhere, VMM mappings) */
int lv_exchange_connect_ptrs(lv_exchange *x, void *const *peer_buffers)
{
    if (!x || !peer_buffers) return LV_E_ARG;
    if (x->connected) return LV_E_STATE;
    for (int r = 0; r < x->world; r++) {
        if (r == x->rank) continue;
        if (!peer_buffers[r]) return LV_E_ARG;
        x->peers.p[r] = (uint32_t *)peer_buffers[r];
    }
    x->connected = true;
    return LV_OK;
}

/* where the fused kernel must write this rank's block of step `step` (pass it as d_summary): row `rank` of ring slot
 * step % depth of the LOCAL buffer, slot_words u32 */
uint32_t *lv_exchange_local_block(lv_exchange *x, uint64_t step)
{
    if (!x) return nullptr;
    const int slot = (int)(step % (uint64_t)x->L.depth);
    return x->local + ((size_t)slot * x->world + x->rank) * x->L.slot_words;
}

int lv_exchange_publish(lv_exchange *x, uint64_t step, int n_words, uint32_t *d_local_copy, void *cuda_stream)
{
    if (!x || n_words < 0 || n_words > x->L.slot_words) return LV_E_ARG;
    if (!x->connected || step != x->published) return LV_E_STATE;      // steps are published in order, once
    DevGuard g(x->device);
    k_publish<<<1, 32 * x->world, 0, (cudaStream_t)cuda_stream>>>(x->peers, x->L, x->rank, (uint32_t)step, n_words,
                                                                      d_local_copy);
    X_CUDA(cudaGetLastError());
    x->published++;
    return LV_OK;
}

/* d_table: world * slot_words u32 (row r = rank r's block), or NULL to retire the step without reading it */
int lv_exchange_collect(lv_exchange *x, uint64_t step, uint32_t *d_table, void *cuda_stream)
{
    if (!x) return LV_E_ARG;
    // steps are collected in order, once, and only after this rank has published its own block of that step
    if (!x->connected || step != x->collected || step >= x->published) return LV_E_STATE;
    DevGuard g(x->device);
    k_collect<<<1, 256, 0, (cudaStream_t)cuda_stream>>>(x->peers, x->L, x->rank, (uint32_t)step, d_table);
    X_CUDA(cudaGetLastError());
    x->collected++;
    return LV_OK;
}

/* 0 = fine, 1 = a publish gave up waiting for a peer's ack, 2 = a collect gave up waiting for a peer's block */
int lv_exchange_error(lv_exchange *x)
{
    if (!x) return LV_E_ARG;
    DevGuard g(x->device);
    uint32_t v = 0;
    X_CUDA(cudaMemcpy(&v, x->local + x->L.err_off(), sizeof v, cudaMemcpyDeviceToHost));
    return (int)v;
}

// ---- the whole sharded step in one call ------------------------------------------------------------------------------
// lv_exchange_step = lv_convert_diff_batch writing its summary straight into this rank's block of the ring, then (on the
// exchange's own side stream, behind an event) the push to every peer and the collect of step k - lag.  The caller's
// stream only ever waits for the collect of step k - depth, which finished long ago unless a peer is a ring behind.
namespace {
int ensure_step_state(lv_exchange *x)
{
    if (x->side) return LV_OK;
    const int d = x->L.depth;
    X_CUDA(cudaStreamCreateWithFlags(&x->side, cudaStreamNonBlocking));
    x->ev_kernel = new (std::nothrow) cudaEvent_t[d]();
    x->ev_collected = new (std::nothrow) cudaEvent_t[d]();
    if (!x->ev_kernel || !x->ev_collected) return LV_E_NOMEM;
    for (int i = 0; i < d; i++) {
        X_CUDA(cudaEventCreateWithFlags(&x->ev_kernel[i], cudaEventDisableTiming));
        X_CUDA(cudaEventCreateWithFlags(&x->ev_collected[i], cudaEventDisableTiming));
    }
    X_CUDA(cudaMalloc(&x->tables, sizeof(uint32_t) * (size_t)d * x->world * x->L.slot_words));
    return LV_OK;
}

int collect_through(lv_exchange *x, uint64_t step)
{
    while (x->collected <= step) {
        const uint64_t j = x->collected;
        const int slot = (int)(j % (uint64_t)x->L.depth);
        int rc = lv_exchange_collect(x, j, x->tables + (size_t)slot * x->world * x->L.slot_words, x->side);
        if (rc != LV_OK) return rc;
        X_CUDA(cudaEventRecord(x->ev_collected[slot], x->side));
    }
    return LV_OK;
}
}  // namespace

int lv_exchange_step(lv_exchange *x, lv_ctx *ctx, const uint8_t *d_i420, int first_stream, int n_streams, int n_frames,
                     uint8_t *d_bgr, uint16_t *d_mask_bits, uint8_t *d_mask_bytes, uint16_t *d_mb_count, uint8_t *d_mb_map,
                     uint32_t *d_summary, void *cuda_stream, uint64_t *step_out)
{
    if (!x || !ctx || n_streams < 0 || n_frames < 0) return LV_E_ARG;
    const long long n_words = 2LL * n_streams * n_frames;
    if (n_words > x->L.slot_words) return LV_E_ARG;
    if (!x->connected || x->steps != x->published) return LV_E_STATE;    // not mixed with manual publish calls
    DevGuard g(x->device);
    int rc = ensure_step_state(x);
    if (rc != LV_OK) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const uint64_t k = x->steps;
    const int slot = (int)(k % (uint64_t)x->L.depth);
    if (k >= (uint64_t)x->L.depth) {
        // the fused kernel of step k overwrites this rank's block in ring slot `slot`: step k-depth must have left it
        rc = collect_through(x, k - (uint64_t)x->L.depth);
        if (rc != LV_OK) return rc;
        X_CUDA(cudaStreamWaitEvent(st, x->ev_collected[slot], 0));
    }
    uint32_t *block = lv_exchange_local_block(x, k);
    if (n_words > 0) {
        rc = lv_convert_diff_batch(ctx, d_i420, first_stream, n_streams, n_frames, d_bgr, d_mask_bits, d_mask_bytes, d_mb_count,
                                   d_mb_map, block, cuda_stream);
        if (rc != LV_OK) return rc;
    }
    X_CUDA(cudaEventRecord(x->ev_kernel[slot], st));
    X_CUDA(cudaStreamWaitEvent(x->side, x->ev_kernel[slot], 0));
    rc = lv_exchange_publish(x, k, (int)n_words, d_summary, x->side);
    if (rc != LV_OK) return rc;
    x->steps++;
    if (k >= (uint64_t)x->lag) {
        rc = collect_through(x, k - (uint64_t)x->lag);
        if (rc != LV_OK) return rc;
    }
    if (step_out) *step_out = k;
    return LV_OK;
}

/* Table of `step` (world rows of slot_words u32, device memory owned by the exchange, valid until `depth` further steps
 * have been issued).  Blocks the host until every rank's block of that step has arrived. */
int lv_exchange_result(lv_exchange *x, uint64_t step, const uint32_t **d_table)
{
    if (!x || !d_table) return LV_E_ARG;
    if (!x->side || step >= x->steps || x->steps - step > (uint64_t)x->L.depth) return LV_E_STATE;
    DevGuard g(x->device);
    int rc = collect_through(x, step);
    if (rc != LV_OK) return rc;
    X_CUDA(cudaStreamSynchronize(x->side));
    const int err = lv_exchange_error(x);
    if (err != 0) {
        snprintf(x_err, sizeof x_err, "peer exchange gave up waiting for a peer (%s)", err == 1 ? "publish" : "collect");
        return LV_E_STATE;
    }
    *d_table = x->tables + (size_t)(step % (uint64_t)x->L.depth) * x->world * x->L.slot_words;
    return LV_OK;
}

}  // extern "C"
