// lv_exchange.cu -- peer-memory exchange of the per-(stream, frame) change summaries (C ABI: lv_exchange_*).
//
// The path shards by stream ID with no data-path collective; the one exchange it has is the few
// hundred bytes of {changed_px, changed_mbs} per step that every rank wants from every other rank.  A library
// all-gather of that size is pure latency (an 8-rank NCCL all-gather costs more than the whole 114 us kernel) and
// its kernels rendezvous across all GPUs every step.  Here each rank PUSHES its block straight into every peer's table
// with NVLink stores, and nothing on the step ever waits for a round trip:
//
//   symmetric buffer (one per rank, same layout, peers map each other's through CUDA IPC or any shared mapping)
//     data [depth][world][slot_words]  u64   word i of rank r's block of step k lives in ring slot k % depth as
//                                            {value (low 32 bits), k+1 (high 32 bits)}: ONE 8-byte store carries the
//                                            word and its own sequence number, so the receiver needs no flag behind a
//                                            fence -- a word has arrived when its upper half says k+1
//     ack  [world]                      u32   ack[r] == number of steps rank r has collected (retired)
//     gate [world]                      u32   gate[r] == number of lv_exchange_gate calls rank r has reached
//     err                               u32   set when a wait gave up (a wait never spins forever: a hung GPU is worse);
//                                             mirrored into a host-mapped word so the host reads it without a copy
//
//   The sequence number of step k is (k mod (2^32 - 1)) + 1: never 0 (0 = "nothing has arrived"), for any 64-bit k.
//   publish(k): wait (local reads) until every rank has retired step k-depth, then store the block, fire and forget.
//   collect(k): spin on LOCAL memory until every word of ring slot k % depth carries k+1, copy the values out, then
//               store ack = k+1 into every peer (the loads have returned by then; nothing to fence).
//
//   lv_exchange_step: the fused kernel of step k accumulates into row k % depth of a ring owned by the exchange; behind
//     an event, on the exchange's own side stream, one small kernel pushes that row (and copies it to the caller's
//     d_summary) and another collects step k-lag.  The caller's stream gets the two launches of the single-GPU step
//     plus one event record; it waits for the side stream only if the push of step k-depth has not finished when step
//     k is issued (checked on the host: practically never).
//
// Measured alternatives (profiles/README.md, N = 2, us per step over the single-GPU 116.3): this design with flag +
// __threadfence_system() instead of sequence-numbered words +1.4; the same pushes from a kernel IN the caller's stream
// (in place of the memset) +5.2 with or without fences -- a kernel that stores to a peer ends only after the NVLink
// write acknowledgements; letting block 0 of the fused kernel carry the exchange slowed the fused kernel's own code
// by 10-14 % (even as a never-taken noinline call); NCCL all-gather started asynchronously +4.7.
#include "../../include/livevideo.h"

#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <new>

namespace {

constexpr int kMaxWorld = 32;
constexpr unsigned long long kSpinLimitNs = 4000000000ull;   // a KERNEL gives up 4 s after it started and raises err

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
    __host__ __device__ size_t data_words64() const { return (size_t)depth * world * slot_words; }
    __host__ __device__ size_t ack_off32() const { return 2 * data_words64(); }                  // in u32 words
    __host__ __device__ size_t gate_off32() const { return ack_off32() + (size_t)world; }
    __host__ __device__ size_t err_off32() const { return gate_off32() + (size_t)world; }
    __host__ __device__ size_t bytes() const { return ((err_off32() + 1) * 4 + 15) & ~(size_t)15; }
};

struct Peers {
    unsigned long long *p[kMaxWorld];
};

__device__ __forceinline__ unsigned long long ld_volatile64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_volatile32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_volatile32(uint32_t *p, uint32_t v)
{
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t *ack_of(unsigned long long *buf, const Layout &L)
{
    return reinterpret_cast<uint32_t *>(buf) + L.ack_off32();
}
__device__ __forceinline__ uint32_t *gate_of(unsigned long long *buf, const Layout &L)
{
    return reinterpret_cast<uint32_t *>(buf) + L.gate_off32();
}
__device__ __forceinline__ uint32_t *err_of(unsigned long long *buf, const Layout &L)
{
    return reinterpret_cast<uint32_t *>(buf) + L.err_off32();
}
__host__ __device__ __forceinline__ uint32_t seq_of(unsigned long long step)
{
    return (uint32_t)(step % 0xffffffffull) + 1u;      // 1 .. 2^32-1, never the "empty" value 0
}

// One deadline per kernel: `ok` (shared memory) drops to 0 when ANY thread of the block has waited past t0 + 4 s, and
// every spin of every thread reads it, so a dead peer costs the stream 4 s in all -- not 4 s per word.
struct Waiter {
    volatile int *ok;
    unsigned long long t0;
    __device__ __forceinline__ bool alive() const { return *ok != 0; }
    // one back-off step of a spin; false = give up
    __device__ __forceinline__ bool pause() const
    {
        __nanosleep(200);
        if (!*ok) return false;
        if (now_ns() - t0 > kSpinLimitNs) { *ok = 0; return false; }
        return true;
    }
};

// Whole block: shared `ok` = 1 unless this rank's err word is already set (a broken exchange stays broken: later kernels
// return at once instead of spinning again).
__device__ __forceinline__ Waiter begin_wait(unsigned long long *local, const Layout &L, int *ok)
{
    if (threadIdx.x == 0) *ok = ld_volatile32(err_of(local, L)) == 0u ? 1 : 0;
    __syncthreads();
    return Waiter{ok, now_ns()};
}
__device__ __forceinline__ void raise_err(unsigned long long *local, const Layout &L, uint32_t *err_host, uint32_t code)
{
    atomicCAS(err_of(local, L), 0u, code);
    if (err_host && *reinterpret_cast<volatile uint32_t *>(err_host) == 0u) *reinterpret_cast<volatile uint32_t *>(err_host) = code;
}

// Whole block.  Store this rank's block of `step` (n_words words at src, zero beyond, slot_words in all) into row `rank`
// of ring slot `slot` of EVERY rank's table, each word with its sequence number; warp v serves ranks v, v+nwarps...
__device__ void publish_block(const Peers &peers, const Layout &L, int rank, unsigned long long step, int slot, const uint32_t *src,
                              int n_words, uint32_t *local_copy, uint32_t *err_host, int *ok)
{
    unsigned long long *local = peers.p[rank];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const Waiter wt = begin_wait(local, L, ok);
    // flow control: every rank must have retired the step that last used this ring slot (local reads; free in steady state)
    if (step >= (unsigned long long)L.depth && lane == 0 && wt.alive()) {
        const uint32_t want = (uint32_t)(step - (unsigned long long)L.depth + 1ull);
        for (int r = warp; r < L.world; r += nwarps) {
            const uint32_t *a = ack_of(local, L) + r;
            while ((int32_t)(ld_volatile32(a) - want) < 0)
                if (!wt.pause()) break;
            if (!wt.alive()) break;
        }
    }
    __syncthreads();
    if (!*ok) {
        if (threadIdx.x == 0) raise_err(local, L, err_host, 1u);
        return;
    }
    const size_t row = ((size_t)slot * L.world + rank) * L.slot_words;
    const unsigned long long seq = (unsigned long long)seq_of(step) << 32;
    for (int r = warp; r < L.world; r += nwarps) {
        unsigned long long *dst = peers.p[r] + row;
        if (r == rank && local_copy && local_copy != src) {
            // the caller's own summary buffer gets its usual content BEFORE this rank's words appear in its own table:
            // a complete table of step k (lv_exchange_result) then implies that d_summary of step k is in place
            for (int i = lane; i < n_words; i += 32) local_copy[i] = src[i];
            __threadfence();
        }
        for (int i = lane; i < L.slot_words; i += 32) st_volatile64(dst + i, seq | (i < n_words ? src[i] : 0u));
    }
}

// Whole block.  For each of the n_steps steps from `step` on, in order: wait until every word of the step's ring slot
// carries its sequence number, copy the values out (world x slot_words u32; `out` = one table, or `ring` = the base of a
// ring of `depth` tables indexed by slot; both may be NULL), then retire the step at every rank.  Several lagged steps
// cost one launch, not one each.
__device__ void collect_block(const Peers &peers, const Layout &L, int rank, unsigned long long step, int n_steps, uint32_t *out,
                              uint32_t *ring, uint32_t *err_host, unsigned long long *done_host, int *ok)
{
    unsigned long long *local = peers.p[rank];
    const Waiter wt = begin_wait(local, L, ok);
    const size_t n = (size_t)L.world * L.slot_words;
    for (int s = 0; s < n_steps; s++, step++) {
        const int slot = (int)(step % (unsigned long long)L.depth);
        const unsigned long long *src = local + (size_t)slot * n;
        uint32_t *dst = ring ? ring + (size_t)slot * n : out;
        const uint32_t want = seq_of(step);
        for (size_t i = threadIdx.x; i < n && wt.alive(); i += blockDim.x) {
            unsigned long long v = ld_volatile64(src + i);
            while ((uint32_t)(v >> 32) != want) {
                if (!wt.pause()) break;
                v = ld_volatile64(src + i);
            }
            if (dst) dst[i] = (uint32_t)v;
        }
        __threadfence_system();          // the table is complete for every later reader, on the device or the host ...
        __syncthreads();
        if (!*ok) {
            if (threadIdx.x == 0) raise_err(local, L, err_host, 2u);
            return;
        }
        // ... before the host is told (it polls this host-mapped word instead of synchronising the stream)
        if (done_host && threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(done_host) = step + 1ull;
        // retire: the slot may be overwritten once every rank has said so (all loads above have returned)
        const uint32_t done = (uint32_t)(step + 1ull);
        for (int r = threadIdx.x; r < L.world; r += blockDim.x) st_volatile32(ack_of(peers.p[r], L) + rank, done);
    }
}

// `trace` (NULL unless lv_exchange_trace was asked for): globaltimer stamps of the newest step, host-mapped --
// [0] publish entry, [1] publish exit, [2] collect entry, [3] collect exit, [4] end of the fused kernel (k_stamp)
__global__ void __launch_bounds__(256) k_publish(Peers peers, Layout L, int rank, unsigned long long step, int slot,
                                                 const uint32_t *src, int n_words, uint32_t *local_copy, uint32_t *err_host,
                                                 unsigned long long *trace)
{
    __shared__ int ok;
    if (trace && threadIdx.x == 0) trace[0] = now_ns();
    publish_block(peers, L, rank, step, slot, src, n_words, local_copy, err_host, &ok);
    __syncthreads();
    if (trace && threadIdx.x == 0) trace[1] = now_ns();
}

__global__ void k_stamp(unsigned long long *dst) { *dst = now_ns(); }

__global__ void __launch_bounds__(256) k_collect(Peers peers, Layout L, int rank, unsigned long long step, int n_steps, uint32_t *out,
                                                 uint32_t *ring, uint32_t *err_host, unsigned long long *done_host,
                                                 unsigned long long *trace)
{
    __shared__ int ok;
    if (trace && threadIdx.x == 0) trace[2] = now_ns();
    collect_block(peers, L, rank, step, n_steps, out, ring, err_host, done_host, &ok);
    __syncthreads();
    if (trace && threadIdx.x == 0) trace[3] = now_ns();
}

// Device-side rendezvous of all ranks IN the caller's stream: lane r tells rank r "I have reached gate number `count`",
// then waits until rank r has said the same here.  Everything enqueued behind it starts within a few microseconds of
// the same instant on every GPU (the hosts may be milliseconds apart).
//   `hold` (host-mapped, optional): first wait until the host has written `count` there (lv_exchange_gate_open) -- a rank
//   tells its peers that it has arrived only when its own host has the work behind the gate enqueued, so when the gate
//   opens EVERY GPU has that work in its queue.
__global__ void __launch_bounds__(32) k_gate(Peers peers, Layout L, int rank, uint32_t count, const uint32_t *hold, uint32_t *err_host)
{
    __shared__ int ok;
    unsigned long long *local = peers.p[rank];
    const Waiter wt = begin_wait(local, L, &ok);
    if (hold && threadIdx.x == 0)
        while ((int32_t)(*reinterpret_cast<const volatile uint32_t *>(hold) - count) < 0)
            if (!wt.pause()) break;
    __syncthreads();
    for (int r = threadIdx.x; r < L.world; r += 32) st_volatile32(gate_of(peers.p[r], L) + rank, count);
    for (int r = threadIdx.x; r < L.world && wt.alive(); r += 32) {
        const uint32_t *g = gate_of(local, L) + r;
        while ((int32_t)(ld_volatile32(g) - count) < 0)
            if (!wt.pause()) break;
    }
    __syncthreads();
    if (!ok && threadIdx.x == 0) raise_err(local, L, err_host, 3u);
}

}  // namespace

struct lv_exchange {
    int device = 0, rank = 0, world = 1;
    Layout L{};
    unsigned long long *local = nullptr;   // cudaMalloc'ed, IPC-exportable
    Peers peers{};
    bool opened[kMaxWorld] = {};           // peers.p[r] came from cudaIpcOpenMemHandle
    bool connected = false;
    uint64_t published = 0, collected = 0; // steps whose publish / collect has been enqueued
    // lv_exchange_step
    cudaStream_t side = nullptr;           // the pushes run here, never in the caller's stream
    cudaStream_t coll = nullptr;           // the collects run here: a collect only reads local memory, so it depends on no
                                           // push of this rank and can already be spinning when the last block arrives
    unsigned long long *h_done = nullptr;  // host-mapped: number of steps whose table is complete (written by the collects)
    cudaEvent_t ev_collected = nullptr;    // behind the newest collect launch (lv_exchange_wait_table)
    cudaEvent_t *ev_kernel = nullptr;      // [depth] fused kernel of the step in this ring slot has finished
    cudaEvent_t *ev_pushed = nullptr;      // [depth] the push of the step in this ring slot has read its row
    uint32_t *tables = nullptr;            // ring of collected tables: depth x world x slot_words
    uint32_t *own = nullptr;               // ring of this rank's own blocks (what the fused kernels accumulate into)
    uint64_t steps = 0;                    // next step of lv_exchange_step
    bool stepping = false;
    int lag = 0;                           // step k's table is collected when step k+lag is issued (or on result)
    uint32_t *h_err = nullptr;             // host-mapped mirror of the err word (read by the host without a copy)
    uint32_t gates = 0;                    // lv_exchange_gate calls so far
    uint32_t *h_go = nullptr;              // host-mapped: number of gates the host has opened (lv_exchange_gate_open)
    unsigned long long *trace = nullptr;   // host-mapped time stamps of the newest step (lv_exchange_trace), else NULL
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
    return L.bytes();
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
    // a step's table is collected lag + 1 steps later: peers may run up to ~depth/2 steps apart before anyone waits
    x->lag = depth / 2 < depth - 2 ? depth / 2 : depth - 2;
    const size_t bytes = x->L.bytes();
    cudaError_t e = cudaMalloc(&x->local, bytes);
    if (e == cudaSuccess) e = cudaMemset(x->local, 0, bytes);       // sequence number 0 = "nothing has arrived"
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&x->h_err), sizeof(uint32_t), cudaHostAllocMapped);
    if (e == cudaSuccess) *x->h_err = 0u;
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&x->h_done), sizeof(unsigned long long), cudaHostAllocMapped);
    if (e == cudaSuccess) *x->h_done = 0ull;
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        int rc = x_fail(e, "exchange buffer");
        cudaFree(x->local);
        if (x->h_err) cudaFreeHost(x->h_err);
        if (x->h_done) cudaFreeHost(x->h_done);
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
            if (x->ev_pushed[i]) cudaEventDestroy(x->ev_pushed[i]);
        }
    delete[] x->ev_kernel;
    delete[] x->ev_pushed;
    if (x->side) cudaStreamDestroy(x->side);
    if (x->coll) cudaStreamDestroy(x->coll);
    if (x->ev_collected) cudaEventDestroy(x->ev_collected);
    cudaFree(x->tables);
    cudaFree(x->own);
    cudaFree(x->local);
    if (x->h_err) cudaFreeHost(x->h_err);
    if (x->h_done) cudaFreeHost(x->h_done);
    if (x->h_go) cudaFreeHost(x->h_go);
    if (x->trace) cudaFreeHost(x->trace);
    delete x;
}

/* Diagnostics: 8 host-mapped u64 GPU-globaltimer stamps (ns) of the newest step -- [0] push kernel entry, [1] exit,
 * [2] collect kernel entry, [3] exit, [4] end of the fused kernel (a one-thread kernel behind it, only while tracing).
 * Returns NULL on failure.  Costs one extra launch per step: never enable it in a measured run. */
const unsigned long long *lv_exchange_trace(lv_exchange *x)
{
    if (!x) return nullptr;
    if (!x->trace) {
        DevGuard g(x->device);
        if (cudaHostAlloc(reinterpret_cast<void **>(&x->trace), 8 * sizeof(unsigned long long), cudaHostAllocMapped) != cudaSuccess) {
            x->trace = nullptr;
            return nullptr;
        }
        memset(x->trace, 0, 8 * sizeof(unsigned long long));
    }
    return x->trace;
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
        x->peers.p[r] = (unsigned long long *)p;
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
        x->peers.p[r] = (unsigned long long *)peer_buffers[r];
    }
    x->connected = true;
    return LV_OK;
}

/* ---- the two halves on their own (any stream order the caller likes) -------------------------------------------- */
int lv_exchange_publish(lv_exchange *x, uint64_t step, const uint32_t *d_block, int n_words, void *cuda_stream)
{
    if (!x || n_words < 0 || n_words > x->L.slot_words || (n_words > 0 && !d_block)) return LV_E_ARG;
    if (!x->connected || x->stepping || step != x->published) return LV_E_STATE;      // steps are published in order, once
    DevGuard g(x->device);
    k_publish<<<1, 256, 0, (cudaStream_t)cuda_stream>>>(x->peers, x->L, x->rank, step, (int)(step % (uint64_t)x->L.depth), d_block,
                                                        n_words, nullptr, x->h_err, x->trace);
    X_CUDA(cudaGetLastError());
    x->published++;
    return LV_OK;
}

/* d_table: world * slot_words u32 (row r = rank r's block), or NULL to retire the step without reading it */
int lv_exchange_collect(lv_exchange *x, uint64_t step, uint32_t *d_table, void *cuda_stream)
{
    if (!x) return LV_E_ARG;
    // steps are collected in order, once, and only after this rank has published its own block of that step
    if (!x->connected || x->stepping || step != x->collected || step >= x->published) return LV_E_STATE;
    DevGuard g(x->device);
    k_collect<<<1, 256, 0, (cudaStream_t)cuda_stream>>>(x->peers, x->L, x->rank, step, 1, d_table, nullptr, x->h_err, nullptr, x->trace);
    X_CUDA(cudaGetLastError());
    x->collected++;
    return LV_OK;
}

/* 0 = fine, 1 = a publish gave up waiting for a peer's ack, 2 = a collect gave up waiting for a peer's block, 3 = a gate
 * gave up.  Reads the host-mapped mirror of the err word: no copy, no synchronisation (kernels that have finished have
 * made their write visible). */
int lv_exchange_error(lv_exchange *x)
{
    if (!x) return LV_E_ARG;
    return (int)*reinterpret_cast<volatile uint32_t *>(x->h_err);
}

/* Device-side rendezvous in `cuda_stream`: work enqueued behind it starts on every rank within microseconds of the same
 * instant, however far apart the host threads are.  Collective: every rank calls it the same number of times. */
int lv_exchange_gate(lv_exchange *x, void *cuda_stream, int hold)
{
    if (!x) return LV_E_ARG;
    if (!x->connected) return LV_E_STATE;
    DevGuard g(x->device);
    if (hold && !x->h_go) {
        X_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&x->h_go), sizeof(uint32_t), cudaHostAllocMapped));
        *x->h_go = x->gates;                 // every earlier gate counts as opened
    }
    x->gates++;
    if (!hold && x->h_go) *reinterpret_cast<volatile uint32_t *>(x->h_go) = x->gates;
    k_gate<<<1, 32, 0, (cudaStream_t)cuda_stream>>>(x->peers, x->L, x->rank, x->gates, hold ? x->h_go : nullptr, x->h_err);
    X_CUDA(cudaGetLastError());
    return LV_OK;
}

/* Opens the newest held gate of this rank (see lv_exchange_gate).  A held gate that is never opened gives up after 4 s
 * (lv_exchange_error 3). */
int lv_exchange_gate_open(lv_exchange *x)
{
    if (!x) return LV_E_ARG;
    if (!x->h_go) return LV_E_STATE;
    *reinterpret_cast<volatile uint32_t *>(x->h_go) = x->gates;
    return LV_OK;
}

/* Makes `cuda_stream` wait until the caller's d_summary of `step` (an lv_exchange_step) has been written: with the peer
 * exchange that buffer is filled by the push kernel on the exchange's own stream, not by the fused kernel. */
int lv_exchange_wait_summary(lv_exchange *x, uint64_t step, void *cuda_stream)
{
    if (!x) return LV_E_ARG;
    if (!x->stepping || step >= x->steps || x->steps - step > (uint64_t)x->L.depth) return LV_E_STATE;
    DevGuard g(x->device);
    X_CUDA(cudaStreamWaitEvent((cudaStream_t)cuda_stream, x->ev_pushed[step % (uint64_t)x->L.depth], 0));
    return LV_OK;
}

/* ---- the whole sharded step in one call -------------------------------------------------------------------------- */
}  // extern "C"

namespace {
int ensure_step_state(lv_exchange *x)
{
    if (x->side) return LV_OK;
    const int d = x->L.depth;
    // highest priority: their one-block kernels are dispatched ahead of the fused kernel's remaining blocks
    int prio_lo = 0, prio_hi = 0;
    X_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    X_CUDA(cudaStreamCreateWithPriority(&x->side, cudaStreamNonBlocking, prio_hi));
    X_CUDA(cudaStreamCreateWithPriority(&x->coll, cudaStreamNonBlocking, prio_hi));
    X_CUDA(cudaEventCreateWithFlags(&x->ev_collected, cudaEventDisableTiming));
    x->ev_kernel = new (std::nothrow) cudaEvent_t[d]();
    x->ev_pushed = new (std::nothrow) cudaEvent_t[d]();
    if (!x->ev_kernel || !x->ev_pushed) return LV_E_NOMEM;
    for (int i = 0; i < d; i++) {
        X_CUDA(cudaEventCreateWithFlags(&x->ev_kernel[i], cudaEventDisableTiming));
        X_CUDA(cudaEventCreateWithFlags(&x->ev_pushed[i], cudaEventDisableTiming));
    }
    X_CUDA(cudaMalloc(&x->tables, sizeof(uint32_t) * (size_t)d * x->world * x->L.slot_words));
    X_CUDA(cudaMalloc(&x->own, sizeof(uint32_t) * (size_t)d * x->L.slot_words));
    return LV_OK;
}

int collect_through(lv_exchange *x, uint64_t step)
{
    if (x->collected <= step) {          // every outstanding step up to `step` in ONE launch, on the collects' own stream
        k_collect<<<1, 256, 0, x->coll>>>(x->peers, x->L, x->rank, x->collected, (int)(step + 1 - x->collected), nullptr, x->tables,
                                          x->h_err, x->h_done, x->trace);
        X_CUDA(cudaGetLastError());
        X_CUDA(cudaEventRecord(x->ev_collected, x->coll));
        x->collected = step + 1;
    }
    return LV_OK;
}
}  // namespace

extern "C" {

int lv_exchange_step(lv_exchange *x, lv_ctx *ctx, const uint8_t *d_i420, int first_stream, int n_streams, int n_frames,
                     uint8_t *d_bgr, uint16_t *d_mask_bits, uint8_t *d_mask_bytes, uint16_t *d_mb_count, uint8_t *d_mb_map,
                     uint32_t *d_summary, void *cuda_stream, uint64_t *step_out)
{
    if (!x || !ctx || n_streams < 0 || n_frames < 0) return LV_E_ARG;
    const long long n_words = 2LL * n_streams * n_frames;
    if (n_words > x->L.slot_words) return LV_E_ARG;
    if (!x->connected) return LV_E_STATE;
    if (!x->stepping && (x->published != 0 || x->collected != 0)) return LV_E_STATE;    // not mixed with the manual halves
    DevGuard g(x->device);
    int rc = ensure_step_state(x);
    if (rc != LV_OK) return rc;
    x->stepping = true;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const uint64_t k = x->steps;
    const int slot = (int)(k % (uint64_t)x->L.depth);
    uint32_t *block = x->own + (size_t)slot * x->L.slot_words;
    if (k >= (uint64_t)x->L.depth) {
        // The fused kernel of step k re-uses the row of step k-depth: its push must have read it.  That happened depth-1
        // steps ago, so ask on the host and make the caller's stream wait only if it really has not.
        const cudaError_t q = cudaEventQuery(x->ev_pushed[slot]);
        if (q == cudaErrorNotReady) X_CUDA(cudaStreamWaitEvent(st, x->ev_pushed[slot], 0));
        else if (q != cudaSuccess) return x_fail(q, "cudaEventQuery");
    }
    if (n_words > 0) {
        rc = lv_convert_diff_batch(ctx, d_i420, first_stream, n_streams, n_frames, d_bgr, d_mask_bits, d_mask_bytes, d_mb_count,
                                   d_mb_map, block, cuda_stream);
        if (rc != LV_OK) return rc;
    }
    X_CUDA(cudaEventRecord(x->ev_kernel[slot], st));
    X_CUDA(cudaStreamWaitEvent(x->side, x->ev_kernel[slot], 0));
    if (x->trace) k_stamp<<<1, 1, 0, st>>>(x->trace + 4);
    k_publish<<<1, 256, 0, x->side>>>(x->peers, x->L, x->rank, k, slot, block, (int)n_words, d_summary, x->h_err, x->trace);
    X_CUDA(cudaGetLastError());
    X_CUDA(cudaEventRecord(x->ev_pushed[slot], x->side));
    x->published = k + 1;
    x->steps = k + 1;
    if (k >= (uint64_t)x->lag) {
        rc = collect_through(x, k - (uint64_t)x->lag);
        if (rc != LV_OK) return rc;
    }
    if (step_out) *step_out = k;
    return LV_OK;
}

/* Kept for callers written against the in-stream variant: every push is already on its way when lv_exchange_step
 * returns, so there is nothing to flush. */
int lv_exchange_flush(lv_exchange *x) { return x ? LV_OK : LV_E_ARG; }

/* Table of `step` (world rows of slot_words u32, device memory owned by the exchange, valid until `depth` further steps
 * have been issued).  Blocks the host until every rank's block of that step has arrived; afterwards the caller's
 * d_summary of every step up to `step` holds this rank's block too. */
int lv_exchange_result(lv_exchange *x, uint64_t step, const uint32_t **d_table)
{
    if (!x || !d_table) return LV_E_ARG;
    if (!x->stepping || step >= x->steps || x->steps - step > (uint64_t)x->L.depth) return LV_E_STATE;
    DevGuard g(x->device);
    int rc = collect_through(x, step);
    if (rc != LV_OK) return rc;
    // The collect kernel tells the host through a host-mapped word: no stream synchronisation on the way out (a collect
    // never spins for more than 4 s, so neither does this loop -- 6 s covers a kernel that could not even start).
    {
        const volatile unsigned long long *done = x->h_done;
        const volatile uint32_t *errw = x->h_err;
        const auto t0 = std::chrono::steady_clock::now();
        unsigned spins = 0;
        while (*done < step + 1 && *errw == 0u) {
            if ((++spins & 0x3ffu) == 0) {
                if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(6)) {
                    snprintf(x_err, sizeof x_err, "peer exchange: the table of step %llu did not complete within 6 s",
                             (unsigned long long)step);
                    return LV_E_STATE;
                }
                const cudaError_t q = cudaStreamQuery(x->coll);     // a failed launch or a dead context must not spin here
                if (q != cudaSuccess && q != cudaErrorNotReady) return x_fail(q, "collect stream");
            }
        }
    }
    const int err = lv_exchange_error(x);
    if (err != 0) {
        snprintf(x_err, sizeof x_err, "peer exchange gave up waiting for a peer (%s)",
                 err == 1 ? "publish" : err == 2 ? "collect" : "gate");
        return LV_E_STATE;
    }
    *d_table = x->tables + (size_t)(step % (uint64_t)x->L.depth) * x->world * x->L.slot_words;
    return LV_OK;
}

/* The device-side form of lv_exchange_result: makes `cuda_stream` wait until the table of `step` is complete, without
 * involving the host (a kernel that consumes the table is simply enqueued behind this call); *d_table as in
 * lv_exchange_result. */
int lv_exchange_wait_table(lv_exchange *x, uint64_t step, void *cuda_stream, const uint32_t **d_table)
{
    if (!x) return LV_E_ARG;
    if (!x->stepping || step >= x->steps || x->steps - step > (uint64_t)x->L.depth) return LV_E_STATE;
    DevGuard g(x->device);
    int rc = collect_through(x, step);
    if (rc != LV_OK) return rc;
    X_CUDA(cudaStreamWaitEvent((cudaStream_t)cuda_stream, x->ev_collected, 0));     // every collect launched so far
    if (d_table) *d_table = x->tables + (size_t)(step % (uint64_t)x->L.depth) * x->world * x->L.slot_words;
    return LV_OK;
}

}  // extern "C"
