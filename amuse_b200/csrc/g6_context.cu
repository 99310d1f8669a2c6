// g6_context.cu -- host side of the B200 g6 library: device state, staging,
// kernel selection and the C ABI declared in include/g6_b200.h.
//
// Mirrors the behaviour (not the code) of the reference host shells
//   lib/sapporo_light/sapporo.cpp:19-272, send_fetch_data.cpp:30-167,
//   host_evaluate_gravity.cu:31-159, sapporoG6lib.cpp:3-81
// with the limits lifted (dynamic j capacity instead of 131072; 16384 pipes
// instead of 256) and no per-call blocking cudaMemcpy chain: one pinned
// staging buffer per direction, one stream, everything asynchronous until
// g6calc_lasthalf*_ has to hand results to the caller.
//
// There is NO CPU fallback: without a CUDA device g6_open_ fails loudly.

#include "g6_kernels.cuh"
#include "../../include/g6_b200.h"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

using namespace g6b;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "g6_b200: FATAL CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_),   \
                    __FILE__, __LINE__, cudaGetErrorString(e_));                                   \
            exit(-1);                                                                              \
        }                                                                                          \
    } while (0)

// G6_B200_TRACE: add the time since the last mark to phase k (needs locals t0, t1)
#define G6_TR(k)                      \
    if (G.trace) {                    \
        t1 = wall();                  \
        G.tr[k] += t1 - t0;           \
        t0 = t1;                      \
    }

namespace {

// Force-kernel variants (g6x_set_variant ids).
enum Variant {
    V_AUTO = 0,
    V_S4 = 1,    // scalar, 4 i/thread, 256 i-slots  (IB 1024)
    V_S2 = 2,    // scalar, 2 i/thread, 256 i-slots  (IB 512)
    V_S1 = 3,    // scalar, 1 i/thread, 256 i-slots  (IB 256)
    V_W1 = 4,    // scalar, 1 i/thread, 32 i-slots x 8 j-slots (IB 32)
    V_T1 = 5,    // scalar, 1 i/thread, 4 i-slots x 64 j-slots (IB 4)
    V_P4 = 6,    // packed f32x2, 4 i/thread, 256 i-slots (IB 1024)
    V_P2 = 7,    // packed f32x2, 2 i/thread, 256 i-slots (IB 512)
    V_P2W = 8,   // packed f32x2, 2 i/thread, 32 i-slots x 8 j-slots (IB 64)
    V_F4 = 9,    // speculative (mask-free groups + verification), packed, 4 i/thread (IB 1024)
    V_F2 = 10,   // speculative, packed, 2 i/thread (IB 512)
    V_COUNT
};

struct VariantInfo {
    int ib;          // i-particles per CTA
    int ctas_per_sm; // resident CTAs targeted
};

struct Context {
    bool open = false;
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    int npipes = 16384;
    int ngb_cap = 1024;
    int variant = V_AUTO;
    int force_nsplit = 0;
    int refine = 1;        // Newton-refined rsqrt (accuracy first); 0 = raw MUFU.RSQ
    bool ext_stream = false;
    int j_offset = 0;
    long long launches = 0;

    // j state
    int capacity = 0;  // slots allocated (multiple of TILE)
    int nj_hi = 0;     // 1 + highest address ever set
    JState js{};
    // j-memory order (rebuild_order): Morton permutation address -> slot, id table, neighbour bounds
    int *addr_of = nullptr;            // [slot] -> address (device)
    JState js2{};                      // second set of state arrays the permutation gathers into
    int *addr_of2 = nullptr;
    std::vector<int> h_slot_of;        // host mirror of js.slot_of
    std::vector<int> h_id;             // id last staged for every address (detects id changes)
    unsigned *d_keys = nullptr, *d_keys_tmp = nullptr;   // [capacity] sorted Morton keys / sort input
    int *d_vals = nullptr, *d_vals_tmp = nullptr;
    void *d_sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    int sort_cap = 0;
    int *d_box = nullptr, *h_box = nullptr;   // 6 ordered ints + (at byte 32) 4 doubles: sum of positions, count
    u64 *d_hash = nullptr;
    unsigned hash_size = 0;
    OrderInfo ord{};
    bool order_valid = false;
    bool order_tiny = false;           // <= 64 j: all pairs go to FP64
    int order_nj = -1;
    bool ids_dirty = false;
    long long updates_since_order = 0, updates_since_force = 0;
    long long order_rebuilds = 0;
    float kclose = 32.f;               // G6_B200_KCLOSE
    float farc = 0.125f;               // G6_B200_FARC
    int near_w = 32;                   // Morton window (each side) of the neighbour-bound scans
    int gmax = 256;                    // G6_B200_GMAX: group boxes a particle's FP64 radius may touch
    unsigned long long *d_stats = nullptr;   // block-class counters (-DG6_STATS builds)
    // FP64 pairs: global list of the speculative kernel + per-particle results (see ForceArgs)
    int2 *d_wl = nullptr;
    unsigned int *d_wl_count = nullptr;
    size_t wl_cap = 0;
    double *d_corr = nullptr;
    size_t corr_cap = 0;
    int wl_per_i = 1024;               // G6_B200_WL_PER_I: list entries reserved per i-particle of a launch
    double ti = 0.0;
    double predicted_ti = 0.0;
    int predicted_nj = -1;  // prefix predicted at predicted_ti (-1: none)
    bool j_dirty = false;

    // staging of j-updates
    std::vector<int> pending_of_addr;  // address -> position in the pending batch, -1
    JUpdate *h_up = nullptr;        // pinned batch being filled (= h_up2[up_cur])
    JUpdate *h_up2[2] = {nullptr, nullptr};   // two pinned batches: one fills while the other uploads
    cudaEvent_t up_done[2] = {nullptr, nullptr};
    int up_cur = 0;
    JUpdate *d_up = nullptr;
    int up_cap = 0, up_n = 0;

    // i-block buffers (npipes)
    float4 *h_i = nullptr;  // pinned [4][npipes]
    float4 *d_i = nullptr;  // [4][npipes]
    int *d_conf = nullptr;  // [npipes] slot of the j-particle with the i-particle's id
    std::vector<int> h_perm;   // packed position k of the pending i-block holds caller particle h_perm[k] (empty: identity)
    std::vector<unsigned> h_ikey, h_ikey2;
    std::vector<int> h_perm2;
    double *d_sum = nullptr;  // [7n doubles][n nearest-neighbour ids] of the current i-block
    u64 *d_key = nullptr;
    double *h_sum = nullptr;  // pinned, same layout
    // device-resident entry point scratch
    float4 *d_i2 = nullptr;   // [4][i2_cap]
    int *d_conf2 = nullptr;   // [i2_cap]
    int i2_cap = 0;
    unsigned *d_ikey = nullptr, *d_ikey_tmp = nullptr;   // Morton sort of a device-resident i-set
    int *d_iperm = nullptr, *d_iperm_tmp = nullptr;
    int isort_cap = 0;
    // partial workspace
    double *part_sum = nullptr;
    u64 *part_key = nullptr;
    size_t part_records = 0;
    unsigned int *tickets = nullptr;
    // neighbour lists
    int *d_ngb_cnt = nullptr, *d_ngb_list = nullptr;
    int *h_ngb_cnt = nullptr, *h_ngb_list = nullptr;
    bool ngb_valid = false;    // the last lasthalf2 asked for lists (some h2 > 0): they can be built
    bool ngb_built = false;    // ... have been built on the device
    bool ngb_fetched = false;  // ... and fetched to the host
    double *d_sum2 = nullptr;  // scratch outputs of the list-building pass
    u64 *d_key2 = nullptr;
    int *d_nnid2 = nullptr;

    // latency path (small i-blocks): i-block in the kernel parameters, results written by the
    // kernel into mapped pinned host memory, completion raised as a flag there
    int direct_max = 2048;             // i-blocks up to this size get their results by direct host writes
    int inline_max = 384;              // ... and up to this size travel in the kernel parameters
    int fuse = 1;                      // scatter+predict fused into one launch for small update batches
    int eager_flush = 8192;            // staged updates that trigger an upload while the caller is still staging
    int zc_up_max = 512;               // j-update batches up to this size are read zero-copy by scatter_kernel
    double *dev_h_sum = nullptr;       // device alias of h_sum
    JUpdate *dev_h_up2[2] = {nullptr, nullptr};   // device aliases of h_up2[]
    unsigned long long *h_flag = nullptr, *dev_h_flag = nullptr;
    unsigned long long flag_seq = 0;
    unsigned int *d_done = nullptr;
    bool cur_direct = false;           // the pending i-block signals through h_flag
    bool i_on_device = false;          // d_i holds the pending/last i-block
    // G6_B200_TRACE=1: host-side phase timers (seconds), printed by g6_close_
    int trace = 0;
    double tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tr_calls = 0;

    // multi-GPU exchange over peer memory (g6x_peer_*): see g6_kernels.cuh "Multi-GPU exchange"
    struct Peer {
        bool allocated = false, attached = false, ipc = true;
        int world = 1, rank = 0, cap = 0;
        unsigned char *buf = nullptr;            // own exchange buffer: [2 halves][world slots] + flags[2][world]
        unsigned char *peer_buf[MAX_PEERS + 1] = {};   // all ranks' buffers as seen from here ([rank] = buf)
        size_t slot_bytes = 0, half_bytes = 0, flags_off = 0, buf_bytes = 0;
        unsigned long long seq = 0;
        unsigned int *h_err = nullptr, *dev_h_err = nullptr;   // mapped: a combine kernel gave up waiting
    } peer;
    // mirrors of the launch being issued (set by g6x_calc_device_allreduce around launch_force)
    int win_lo = 0;   // j-window offset of the launch being issued (sharded Hermite step), multiple of TILE
    int mir_n = 0;
    double *mir_sum[MAX_PEERS] = {};
    u64 *mir_key[MAX_PEERS] = {};
    int *mir_id[MAX_PEERS] = {};

    // device-resident Hermite step (g6x_hermite_*)
    struct Hermite {
        int cap = 0;                         // active particles per pass the buffers hold
        int *h_ilist = nullptr, *dev_h_ilist = nullptr;        // mapped pinned
        double *h_olddt = nullptr, *dev_h_olddt = nullptr;     // mapped pinned
        double *h_outdt = nullptr, *dev_h_outdt = nullptr;     // mapped pinned
        double *h_outpot = nullptr, *dev_h_outpot = nullptr;
        int *h_outnn = nullptr, *dev_h_outnn = nullptr;
        double *d_pred = nullptr;                              // [cap][6]
        int *d_ilist = nullptr;                                // device copies of the block's addresses / steps
        double *d_olddt = nullptr;
        float4 *d_i = nullptr;                                 // [4][cap]
        int *d_conf = nullptr;                                 // [cap]
        double *d_sum = nullptr;                               // [cap][7]
        u64 *d_key = nullptr;
        int *d_nnid = nullptr;
        // integrator state of g6x_hermite_evolve (host side: the scheduler's view)
        std::vector<double> time, dt;
        double system_time = 0.0;
        long long block_steps = 0, particle_steps = 0;
        bool initialised = false;
        // multi-GPU: the state is replicated on every rank, the FORCES are sharded -- this rank sums over
        // the j-window [shard_lo, shard_hi) and the partials are exchanged over peer memory; every rank then
        // applies the same corrector to its replica, so no state ever travels (g6x_hermite_set_shard)
        int shard_lo = 0, shard_hi = 0;   // hi == 0: no sharding
    } herm;

    // captured by firsthalf
    int cur_ni = 0, cur_nj = 0;
    float cur_eps2 = 0.f;
    bool cur_any_h2 = false;
    bool cur_spread = false;   // the pending i-block went to all devices (multi-device mode)
    int cd_win_lo = 0, cd_win_hi = 0;   // g6x_set_j_window: slots the device-resident entry points sum over (hi == 0: all)
    cudaEvent_t i_ready = nullptr;   // root: the packed i-block has arrived in d_i
    bool pending = false;
};

// One context per CUDA device the process has opened; the g6 calls work on the current one.  With
// G6_B200_DEVICES > 1 the ABI layer ("multi-device" below) drives several of them as one j-memory.
constexpr int MAX_DEVICES = MAX_PEERS + 1;
Context g_ctx[MAX_DEVICES];
Context *g_cur = &g_ctx[0];
#define G (*g_cur)

int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    if (!s || !*s) return dflt;
    return atoi(s);
}

template <typename T>
void dev_alloc(T *&p, size_t n)
{
    CK(cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T)));
}
template <typename T>
void dev_free(T *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}
template <typename T>
void host_alloc(T *&p, size_t n)
{
    CK(cudaHostAlloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T), cudaHostAllocMapped | cudaHostAllocPortable));
}
template <typename T>
T *dev_alias(T *host_ptr)
{
    T *d = nullptr;
    CK(cudaHostGetDevicePointer((void **)&d, (void *)host_ptr, 0));
    return d;
}
inline double wall()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
template <typename T>
void host_free(T *&p)
{
    if (p) cudaFreeHost(p);
    p = nullptr;
}

template <typename T>
void grow_dev(T *&p, size_t old_n, size_t new_n, cudaStream_t st)
{
    T *q = nullptr;
    dev_alloc(q, new_n);
    CK(cudaMemsetAsync(q, 0, new_n * sizeof(T), st));
    if (p && old_n) CK(cudaMemcpyAsync(q, p, old_n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
    if (p) cudaFree(p);
    p = q;
}

void ensure_capacity(int need)
{
    if (need <= G.capacity) return;
    size_t newcap = std::max<size_t>(G.capacity ? (size_t)G.capacity * 2 : 65536, (size_t)need);
    newcap = (newcap + TILE - 1) / TILE * TILE;
    size_t old = G.capacity;
    for (int k = 0; k < 7; k++) grow_dev(G.js.q[k], old, newcap, G.stream);
    grow_dev(G.js.ia, old, newcap, G.stream);
    grow_dev(G.js.A, old, newcap, G.stream);
    grow_dev(G.js.B, old, newcap, G.stream);
    grow_dev(G.js.C, old, newcap, G.stream);
    grow_dev(G.js.L, old, newcap, G.stream);
    grow_dev(G.js.gbb, old / TILE * TBOX, newcap / TILE * TBOX, G.stream);
    grow_dev(G.js.near2, old, newcap, G.stream);
    grow_dev(G.js.slot_of, old, newcap, G.stream);
    grow_dev(G.addr_of, old, newcap, G.stream);
    order_fill_kernel<<<(unsigned)((newcap - old + 255) / 256), 256, 0, G.stream>>>((int)old, (int)newcap, G.js.slot_of,
                                                                                  G.addr_of, G.js.near2);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(G.stream));
    // the second set of arrays (target of the Morton permutation) and the sort buffers: no contents to keep
    for (int k = 0; k < 7; k++) { dev_free(G.js2.q[k]); dev_alloc(G.js2.q[k], newcap); }
    dev_free(G.js2.ia); dev_alloc(G.js2.ia, newcap);
    dev_free(G.js2.near2); dev_alloc(G.js2.near2, newcap);
    dev_free(G.js.capr2); dev_alloc(G.js.capr2, newcap);   // filled by the re-ordering that follows every growth
    G.js2.capr2 = G.js.capr2;
    dev_free(G.addr_of2); dev_alloc(G.addr_of2, newcap);
    dev_free(G.d_keys); dev_alloc(G.d_keys, newcap);
    dev_free(G.d_keys_tmp); dev_alloc(G.d_keys_tmp, newcap);
    dev_free(G.d_vals); dev_alloc(G.d_vals, newcap);
    dev_free(G.d_vals_tmp); dev_alloc(G.d_vals_tmp, newcap);
    G.capacity = (int)newcap;
    G.pending_of_addr.resize(newcap, -1);
    G.h_slot_of.resize(newcap);
    for (size_t a = old; a < newcap; a++) G.h_slot_of[a] = (int)a;
    G.h_id.resize(newcap, (int)0x80000000);
    G.predicted_nj = -1;
    G.order_valid = false;   // the sorted keys lived in the old buffers
}

void ensure_up_cap(int need)
{
    if (need <= G.up_cap) return;
    int newcap = std::max(need, std::max(4096, G.up_cap * 2));
    CK(cudaStreamSynchronize(G.stream));   // no upload in flight while the batches move
    for (int b = 0; b < 2; b++) {
        JUpdate *nh = nullptr;
        host_alloc(nh, newcap);
        if (b == G.up_cur && G.h_up2[b] && G.up_n) memcpy(nh, G.h_up2[b], sizeof(JUpdate) * G.up_n);
        host_free(G.h_up2[b]);
        G.h_up2[b] = nh;
        G.dev_h_up2[b] = dev_alias(nh);
        if (!G.up_done[b]) CK(cudaEventCreateWithFlags(&G.up_done[b], cudaEventDisableTiming));
    }
    G.h_up = G.h_up2[G.up_cur];
    dev_free(G.d_up);
    dev_alloc(G.d_up, newcap);
    G.up_cap = newcap;
}

void require_open(const char *fn)
{
    if (!G.open) {
        fprintf(stderr, "g6_b200: FATAL %s called before g6_open_\n", fn);
        exit(-1);
    }
}

void run_predictor(int nj);

// Spin on the completion flag a kernel raises in mapped host memory (latency path, Hermite step).  Every
// ~1M polls the stream is queried, so that a failed kernel ends the wait with its CUDA error.
void wait_flag(unsigned long long want, const char *what)
{
    volatile unsigned long long *flag = G.h_flag;
    unsigned long long spins = 0;
    double t_start = 0.0;
    // no wall-clock limit unless the caller sets one (G6_B200_WAIT_SECONDS): a full sweep of a very large
    // system, or a shared GPU, may legitimately take minutes; real failures surface through cudaStreamQuery
    static const int limit_s = env_int("G6_B200_WAIT_SECONDS", 0);
    while (*flag != want) {
        if ((++spins & 0xfffff) == 0) {
            if (t_start == 0.0) t_start = wall();
            if (limit_s > 0 && wall() - t_start > (double)limit_s) {
                fprintf(stderr, "g6_b200: FATAL %s did not complete within %d s (G6_B200_WAIT_SECONDS)\n", what, limit_s);
                exit(-1);
            }
            cudaError_t q = cudaStreamQuery(G.stream);
            if (q == cudaSuccess) {
                if (*flag == want) break;
                fprintf(stderr, "g6_b200: FATAL %s finished without raising its completion flag\n", what);
                exit(-1);
            }
            if (q != cudaErrorNotReady) CK(q);
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
}

// The pending batch has been handed to a kernel on the stream: forget it on the host side and switch
// to the other pinned buffer (its consumer of two flushes ago is long complete; the event wait is a
// formality).
void retire_batch()
{
    CK(cudaEventRecord(G.up_done[G.up_cur], G.stream));
    for (int k = 0; k < G.up_n; k++) G.pending_of_addr[G.h_up[k].addr] = -1;
    G.up_n = 0;
    G.up_cur ^= 1;
    G.h_up = G.h_up2[G.up_cur];
    CK(cudaEventSynchronize(G.up_done[G.up_cur]));
}

// Upload the pending j-updates and scatter them into the state arrays.  Nothing here waits for the
// GPU: the next batch is staged in the other pinned buffer (its upload of two flushes ago is long
// complete; the event wait below is a formality).
void flush_updates()
{
    if (G.up_n == 0) return;
    if (G.up_n <= G.zc_up_max) {
        // small batch (block-timestep regime): the kernel reads the records straight from the mapped
        // pinned batch over PCIe -- no copy-engine round trip
        scatter_kernel<<<(G.up_n + 63) / 64, 64, 0, G.stream>>>(G.up_n, G.dev_h_up2[G.up_cur], G.js);
    } else {
        CK(cudaMemcpyAsync(G.d_up, G.h_up, sizeof(JUpdate) * G.up_n, cudaMemcpyHostToDevice, G.stream));
        scatter_kernel<<<(G.up_n + 255) / 256, 256, 0, G.stream>>>(G.up_n, G.d_up, G.js);
    }
    G.launches++;
    CK(cudaGetLastError());
    retire_batch();
    G.j_dirty = true;
}

// The batch was consumed by a kernel that scatters AND predicts (update_predict_kernel or a fused
// single-launch force kernel): same bookkeeping, and the prediction is current afterwards.
void fill_inline_updates(InlineU &iu)
{
    iu.n = G.up_n;
    for (int k = 0; k < G.up_n; k++) iu.slot[k] = G.h_up[k].slot;
}

// scatter (if any updates are pending) + predict, in as few launches as possible
void predict_with_updates(int nj)
{
    if (G.up_n == 0 || G.up_n > UPD_MAX || !G.fuse) {
        flush_updates();
        run_predictor(nj);
        return;
    }
    if (nj > G.capacity) nj = G.capacity;
    int n = std::max(nj, std::min(G.nj_hi, G.capacity));
    InlineU iu;
    fill_inline_updates(iu);
    update_predict_kernel<<<(n + 255) / 256, 256, 0, G.stream>>>(n, G.ti, G.js, G.dev_h_up2[G.up_cur], iu);
    G.launches++;
    CK(cudaGetLastError());
    retire_batch();
    G.predicted_nj = n;
    G.predicted_ti = G.ti;
    G.j_dirty = false;
}

void run_predictor(int nj)
{
    if (nj > G.capacity) nj = G.capacity;
    if (nj <= 0) return;
    if (!G.j_dirty && G.predicted_nj >= nj && G.predicted_ti == G.ti) return;
    int n = std::max(nj, std::min(G.nj_hi, G.capacity));
    predict_kernel<<<(n + 255) / 256, 256, 0, G.stream>>>(n, G.ti, G.js);
    G.launches++;
    CK(cudaGetLastError());
    G.predicted_nj = n;
    G.predicted_ti = G.ti;
    G.j_dirty = false;
}

// ---- j-memory order ------------------------------------------------------------------------------
// Sort the slots that hold addresses [0, nj) by the Morton key of their positions (addresses >= nj keep
// slot == address order behind them), rebuild the id -> slot table and the neighbour bounds.  Called with
// no update pending; everything runs on the library stream, two short host synchronisations (the box
// comes back to fix the origin and the grid, the permutation to update the host mirror).
void rebuild_order(int nj)
{
    nj = std::min(nj, G.capacity);
    const int n = std::min(G.capacity, std::max(nj, G.nj_hi));
    G.order_valid = true;
    G.order_nj = nj;
    G.ids_dirty = false;
    G.updates_since_order = 0;
    G.ord = OrderInfo{};
    G.ord.kclose = G.kclose;
    G.ord.farc2 = G.farc * G.farc;
    G.order_tiny = false;
    if (n <= 0 || nj <= 0) return;
    G.order_rebuilds++;
    cudaStream_t st = G.stream;
    const int blocks = (n + 255) / 256;
    if (!G.d_box) {
        dev_alloc(G.d_box, 16);
        host_alloc(G.h_box, 16);
    }
    double *h_cen = reinterpret_cast<double *>(G.h_box + 8), *d_cen = reinterpret_cast<double *>(G.d_box + 8);
    for (int d = 0; d < 4; d++) h_cen[d] = 0.0;
    union { int i; float f; } cv;
    auto ord_i = [&](float f) { cv.f = f; return cv.i >= 0 ? cv.i : cv.i ^ 0x7fffffff; };
    auto ord_f = [&](int i) { cv.i = i >= 0 ? i : i ^ 0x7fffffff; return cv.f; };
    for (int d = 0; d < 3; d++) {
        G.h_box[d] = 0x7f7fffff;
        G.h_box[3 + d] = ord_i(-3.0e38f);
    }
    CK(cudaMemcpyAsync(G.d_box, G.h_box, 16 * sizeof(int), cudaMemcpyHostToDevice, st));
    order_box_kernel<<<blocks, 256, 0, st>>>(n, nj, G.js, G.addr_of, G.d_box, d_cen);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(G.h_box, G.d_box, 16 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    double lo[3], hi[3], diag2 = 0.0;
    for (int d = 0; d < 3; d++) {
        lo[d] = ord_f(G.h_box[d]);
        hi[d] = ord_f(G.h_box[3 + d]);
        if (!(hi[d] >= lo[d])) lo[d] = hi[d] = 0.0;   // no massive particle
        // origin: the mean position (not the box centre, which a single escaper drags away from the cluster);
        // in the multi-device mode all devices share device 0's origin, the i-block is packed once
        G.js.x0[d] = (g_cur != &g_ctx[0] && g_ctx[0].order_valid) ? g_ctx[0].js.x0[d]
                                                                   : (h_cen[3] > 0.0 ? h_cen[d] / h_cen[3] : 0.0);
        const double ext = hi[d] - lo[d];
        diag2 += ext * ext;
        G.ord.blo[d] = (float)(lo[d] - G.js.x0[d]);
        G.ord.binv[d] = ext > 0.0 ? (float)(1023.999 / ext) : 0.f;
    }
    for (int d = 0; d < 3; d++) G.js2.x0[d] = G.js.x0[d];
    G.ord.cap2 = (float)(diag2 / 64.0);   // (diagonal / 8)^2
    // a handful of j (two-body and few-body tests, BHTree's shortest lists): every pair in FP64
    G.order_tiny = std::min(nj, n) <= 64;
    if (G.order_tiny && G.kclose > 0.f) {
        G.ord.kclose = 1.0e30f;
        G.ord.cap2 = std::numeric_limits<float>::infinity();
    }
    // keys, sort, permutation
    order_key_kernel<<<blocks, 256, 0, st>>>(n, nj, G.js, G.addr_of, G.ord, G.d_keys_tmp, G.d_vals_tmp);
    CK(cudaGetLastError());
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, G.d_keys_tmp, G.d_keys, G.d_vals_tmp, G.d_vals, n, 0, 32, st));
    if (need > G.sort_tmp_bytes) {
        CK(cudaStreamSynchronize(st));
        if (G.d_sort_tmp) cudaFree(G.d_sort_tmp);
        CK(cudaMalloc(&G.d_sort_tmp, need));
        G.sort_tmp_bytes = need;
    }
    CK(cub::DeviceRadixSort::SortPairs(G.d_sort_tmp, need, G.d_keys_tmp, G.d_keys, G.d_vals_tmp, G.d_vals, n, 0, 32, st));
    order_permute_kernel<<<(G.capacity + 255) / 256, 256, 0, st>>>(n, G.capacity, G.d_vals, G.js, G.addr_of, G.js2,
                                                                   G.addr_of2, G.js.slot_of);
    CK(cudaGetLastError());
    for (int k = 0; k < 7; k++) std::swap(G.js.q[k], G.js2.q[k]);
    std::swap(G.js.ia, G.js2.ia);
    std::swap(G.js.near2, G.js2.near2);
    std::swap(G.addr_of, G.addr_of2);
    // id -> slot table over the prefix
    const int nprefix = std::min(nj, n);
    unsigned size = 1024;
    while (size < 2u * (unsigned)nprefix) size <<= 1;
    if (size > G.hash_size) {
        CK(cudaStreamSynchronize(st));
        dev_free(G.d_hash);
        dev_alloc(G.d_hash, size);
        G.hash_size = size;
    }
    CK(cudaMemsetAsync(G.d_hash, 0, sizeof(u64) * size, st));
    hash_insert_kernel<<<(nprefix + 255) / 256, 256, 0, st>>>(nprefix, G.js, G.d_hash, size - 1);
    CK(cudaGetLastError());
    G.ord.hash = G.d_hash;
    G.ord.hash_mask = size - 1;
    G.ord.keys = G.d_keys;
    G.ord.nkeys = nprefix;
    order_near_kernel<<<(nprefix + 255) / 256, 256, 0, st>>>(nprefix, G.js, G.near_w);
    CK(cudaGetLastError());
    // cap of the FP64 radius (particles whose K d^2 would reach a large part of the system)
    {
        const int ntiles = (nprefix + TILE - 1) / TILE;
        order_boxes_kernel<<<ntiles, TILE, 0, st>>>(nprefix, G.js);
        CK(cudaGetLastError());
        order_cap_kernel<<<(nprefix + 255) / 256, 256, 0, st>>>(nprefix, G.js, ntiles,
                                                                G.order_tiny ? 0x7fffffff : G.gmax);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(G.h_slot_of.data(), G.js.slot_of, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    G.launches += 10;
    G.j_dirty = true;       // slots moved and the origin changed: predict again
    G.predicted_nj = -1;
}

// Does the next force call over [0, nj) need the order rebuilt first?  (called with the pending updates
// counted but not necessarily flushed)
bool order_stale(int nj)
{
    nj = std::min(nj, std::max(G.capacity, 0));
    const bool stale = !G.order_valid || G.order_nj != nj || G.ids_dirty ||
                       G.updates_since_force >= std::max<long long>(1, nj / 2) ||
                       G.updates_since_order >= 4LL * std::max(nj, 1);
    G.updates_since_force = 0;
    return stale;
}

// per-particle FP64 pair results for launches of up to ni particles (zero between launches: whoever writes a
// particle's outputs clears its entry); with_list: also the global pair list of the speculative kernel
void ensure_close_buffers(size_t ni, bool with_list)
{
    if (ni > G.corr_cap) {
        CK(cudaStreamSynchronize(G.stream));
        dev_free(G.d_corr);
        const size_t cap = std::max<size_t>(ni, std::max<size_t>(G.npipes, 2 * G.corr_cap));
        dev_alloc(G.d_corr, cap * 7);
        CK(cudaMemsetAsync(G.d_corr, 0, cap * 7 * sizeof(double), G.stream));
        G.corr_cap = cap;
    }
    if (!with_list) return;
    if (!G.d_wl_count) {
        dev_alloc(G.d_wl_count, 1);
        CK(cudaMemsetAsync(G.d_wl_count, 0, sizeof(unsigned int), G.stream));
    }
    const size_t want = std::min<size_t>((size_t)G.wl_per_i * ni, (size_t)1 << 28);   // <= 2 GiB of (i, slot) pairs
    if (want > G.wl_cap) {
        CK(cudaStreamSynchronize(G.stream));
        dev_free(G.d_wl);
        dev_alloc(G.d_wl, want);
        G.wl_cap = want;
    }
}

void ensure_partials(size_t records)
{
    if (records <= G.part_records) return;
    CK(cudaStreamSynchronize(G.stream));
    dev_free(G.part_sum);
    dev_free(G.part_key);
    dev_alloc(G.part_sum, records * 7);
    dev_alloc(G.part_key, records);
    G.part_records = records;
}

template <int IPT, int NI_SLOTS, bool PACKED, bool NR, int MINB>
void launch_variant_nr(const ForceArgs &a, dim3 grid, bool nn, bool list, cudaStream_t st)
{
    size_t smem = sizeof(ForceSmem);
    static const InlineI<0> none{};
#define G6_LAUNCH(NN_, LIST_)                                                                       \
    do {                                                                                            \
        auto kern = force_kernel<IPT, NI_SLOTS, NN_, LIST_, PACKED, NR, MINB, 0>;                   \
        static bool attr_set[64] = {}; /* per device */                                                               \
        if (!attr_set[G.device & 63]) {                                                                            \
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_set[G.device & 63] = true;                                                                        \
        }                                                                                           \
        kern<<<grid, THREADS, smem, st>>>(a, none);                                                 \
    } while (0)
    if (list)
        G6_LAUNCH(true, true);
    else if (nn)
        G6_LAUNCH(true, false);
    else
        G6_LAUNCH(false, false);
#undef G6_LAUNCH
    CK(cudaGetLastError());
}

template <int IPT, int NI_SLOTS, bool PACKED, int MINB>
void launch_variant(const ForceArgs &a, dim3 grid, bool nn, bool list, cudaStream_t st)
{
    if (G.refine)
        launch_variant_nr<IPT, NI_SLOTS, PACKED, true, MINB>(a, grid, nn, list, st);
    else
        launch_variant_nr<IPT, NI_SLOTS, PACKED, false, MINB>(a, grid, nn, list, st);
}

// Latency path: the i-block rides in the kernel parameters (always with the neighbour search, no lists).
template <int IPT, int NI_SLOTS, bool PACKED, int MINB, int INL>
void launch_inline(const ForceArgs &a, dim3 grid, const InlineI<INL> &ii, cudaStream_t st)
{
    size_t smem = sizeof(ForceSmem);
#define G6_LAUNCH(NR_)                                                                              \
    do {                                                                                            \
        auto kern = force_kernel<IPT, NI_SLOTS, true, false, PACKED, NR_, MINB, INL>;               \
        static bool attr_set[64] = {}; /* per device */                                                               \
        if (!attr_set[G.device & 63]) {                                                                            \
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_set[G.device & 63] = true;                                                                        \
        }                                                                                           \
        kern<<<grid, THREADS, smem, st>>>(a, ii);                                                   \
    } while (0)
    if (G.refine) G6_LAUNCH(true); else G6_LAUNCH(false);
#undef G6_LAUNCH
    CK(cudaGetLastError());
}

template <int IPT, int MINB>
void launch_fast(const ForceArgs &a, dim3 grid, bool nn, cudaStream_t st)
{
    size_t smem = sizeof(FastSmem);
    const bool eps0 = (a.eps2 == 0.f);   // unsoftened: the 2^-52 of the reference is handled by the verification
#define G6_LAUNCH(NN_, NR_)                                                                         \
    do {                                                                                            \
        if (eps0) G6_LAUNCH2(NN_, NR_, true); else G6_LAUNCH2(NN_, NR_, false);                     \
    } while (0)
#define G6_LAUNCH2(NN_, NR_, E0_)                                                                   \
    do {                                                                                            \
        auto kern = force_fast_kernel<IPT, NN_, NR_, MINB, E0_>;                                    \
        static bool attr_set[64] = {}; /* per device */                                                               \
        if (!attr_set[G.device & 63]) {                                                                            \
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_set[G.device & 63] = true;                                                                        \
        }                                                                                           \
        kern<<<grid, THREADS, smem, st>>>(a);                                                       \
    } while (0)
    if (nn) {
        if (G.refine) G6_LAUNCH(true, true); else G6_LAUNCH(true, false);
    } else {
        if (G.refine) G6_LAUNCH(false, true); else G6_LAUNCH(false, false);
    }
#undef G6_LAUNCH
#undef G6_LAUNCH2
    CK(cudaGetLastError());
}

// Device-resident Hermite step, small blocks: masked kernels whose final-output stage runs the corrector.
template <int IPT, int NI_SLOTS, bool PACKED, int MINB>
void launch_herm(const ForceArgs &a, dim3 grid, cudaStream_t st)
{
    size_t smem = sizeof(ForceSmem);
    static const InlineI<0> none{};
#define G6_LAUNCH(NR_)                                                                              \
    do {                                                                                            \
        auto kern = force_kernel<IPT, NI_SLOTS, true, false, PACKED, NR_, MINB, 0, true>;           \
        static bool attr_set[64] = {}; /* per device */                                                               \
        if (!attr_set[G.device & 63]) {                                                                            \
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_set[G.device & 63] = true;                                                                        \
        }                                                                                           \
        kern<<<grid, THREADS, smem, st>>>(a, none);                                                 \
    } while (0)
    if (G.refine) G6_LAUNCH(true); else G6_LAUNCH(false);
#undef G6_LAUNCH
    CK(cudaGetLastError());
}

const VariantInfo &variant_info(int v)
{
    static const VariantInfo info[V_COUNT] = {
        {0, 0}, {1024, 1}, {512, 2}, {256, 2}, {32, 2}, {4, 2}, {1024, 1}, {512, 2}, {64, 2}, {1024, 1}, {512, 2},
    };
    return info[v];
}

// Small i-blocks (the block-timestep regime): every thread of a masked kernel walks IB pairs per
// j-tile (IB = i-particles per CTA), and a launch has ntiles x ceil(ni/IB) CTA-tiles to hand out.  Pick
// the CTA shape whose modelled makespan -- waves x (tiles per CTA x IB x pair cost + fixed cost), in
// units of one scalar pair per thread -- is smallest: with few j (small N) the narrow shapes spread
// the work over more SMs, with many j the wide ones amortise the tile loads.
int choose_variant(int ni, int nj)
{
    if (G.variant != V_AUTO) return G.variant;
    // the speculative kernel (512 i per CTA, every thread walks every j) needs enough pairs to fill the
    // chip; below that the narrow masked shapes finish sooner (measured: tools/ + profiles/ latency tables)
    if (ni > 384 && (double)ni * (double)nj >= 6.0e7) return V_F2;
    if (ni > 384) return V_P2W;
    const int cand[3] = {V_T1, V_W1, V_P2W};
    const double pair_cost[3] = {1.0, 1.0, 0.95};  // measured: in this latency-bound regime packed pairs buy little
    const double fixed = 64.0;                     // prologue + TMA latency + reduction, in pair units
    const double per_tile = 6.0;                   // barrier wait, FP64 flush, next TMA issue: paid per tile by every CTA
    const int ntiles = std::max(1, (nj + TILE - 1) / TILE);
    int best = V_P2W;
    double best_cost = 1e300;
    for (int c = 0; c < 3; c++) {
        const VariantInfo &vi = variant_info(cand[c]);
        const int slots = G.sm_count * vi.ctas_per_sm;
        const int nib = (ni + vi.ib - 1) / vi.ib;
        const long long cta_tiles = (long long)ntiles * nib;
        const double waves_min = std::max(1.0, (double)cta_tiles / slots);   // tiles each slot must take
        const double tps = std::ceil(waves_min);
        const double cost = tps * (vi.ib * pair_cost[c] + per_tile) + fixed * std::max(1.0, (double)nib / slots);
        if (cost < best_cost) {
            best_cost = cost;
            best = cand[c];
        }
    }
    return best;
}

// A packed i-block on the device: four float4 streams, the id-table result per particle, and (device
// path) the index every packed particle's outputs go to.
struct IBlock {
    const float4 *A, *B, *C, *D;
    int *conf;
    const int *iperm;
};
IBlock iblock_of(float4 *base, size_t stride, int *conf, const int *iperm = nullptr)
{
    return IBlock{base, base + stride, base + 2 * stride, base + 3 * stride, conf, iperm};
}

template <int INL>
void fill_inline(InlineI<INL> &ii, const float4 *src, int ni)
{
    for (int k = 0; k < 4; k++) memcpy(ii.d + INL * k, src + (size_t)ni * k, sizeof(float4) * ni);
}

// Launch the force kernel for the packed i-block ib of ni particles against j in [0,nj).
// inline_src != nullptr: HOST pointer to the packed i-block ([4][ni] float4, stride ni), carried in the
// kernel parameters (small i-blocks; masked kernels).  flag_seq != 0: outputs are host-mapped and the
// launch raises G.h_flag = flag_seq when they are complete.
void launch_force(int nj, int ni, const IBlock &ib, float eps2, bool nn, bool list, double *out_sum, u64 *out_key,
                  int *out_nnid, const float4 *inline_src = nullptr, unsigned long long flag_seq = 0,
                  const HermiteArgs *herm = nullptr)
{
    if (nj > G.capacity - G.win_lo) nj = G.capacity - G.win_lo;
    int v = choose_variant(ni, nj);
    if (herm && !(v == V_W1 || v == V_T1 || v == V_P2W)) {
        fprintf(stderr, "g6_b200: FATAL fused Hermite corrector with variant %d\n", v);
        exit(-1);
    }
    if (inline_src && !(v == V_W1 || v == V_T1 || v == V_P2W)) {
        fprintf(stderr, "g6_b200: FATAL inline i-block with variant %d\n", v);
        exit(-1);
    }
    if (list && v == V_F4) v = V_P4;   // neighbour lists test every pair against h2: masked kernels
    if (list && v == V_F2) v = V_P2;
    const VariantInfo &vi = variant_info(v);
    int n_iblocks = (ni + vi.ib - 1) / vi.ib;
    int ntiles = (nj + TILE - 1) / TILE;
    if (ntiles < 1) ntiles = 1;
    // j-splits: pick the split count that minimises the modelled makespan
    //     (waves + w0) x (tiles per CTA + c0)
    // c0 ~ fixed per-CTA cost (prologue + reduction) in tile units; w0 = 0.06 is measured: SMs do not
    // all run at the same speed, so one wave of long CTAs ends ~6 % later than four waves of short
    // ones that the hardware scheduler balances (profiles/r01_grid_granularity.txt).
    // The search result only depends on (variant, i-blocks, tiles): the block-timestep regime asks for the
    // same shape call after call, so the last answer is kept.
    const int slots = G.sm_count * vi.ctas_per_sm;
    static int memo_key[4] = {-1, -1, -1, -1}, memo_ns = 1;
    int nsplit;
    if (memo_key[0] == v && memo_key[1] == n_iblocks && memo_key[2] == ntiles && memo_key[3] == slots) {
        nsplit = memo_ns;
    } else {
        const double c0 = 1.5, w0 = 0.06;
        int best_ns = 1;
        double best_cost = 1e300;
        int max_ns = std::min(ntiles, std::max(1, 4 * slots / n_iblocks));
        for (int ns = 1; ns <= max_ns; ns++) {
            int tps_ = (ntiles + ns - 1) / ns;
            int ns_ = (ntiles + tps_ - 1) / tps_;
            long long ctas = (long long)ns_ * n_iblocks;
            long long waves = (ctas + slots - 1) / slots;
            double cost = ((double)waves + w0) * (tps_ + c0);
            if (cost < best_cost * 0.999) {
                best_cost = cost;
                best_ns = ns_;
            }
        }
        nsplit = best_ns;
        memo_key[0] = v; memo_key[1] = n_iblocks; memo_key[2] = ntiles; memo_key[3] = slots;
        memo_ns = nsplit;
    }
    if (G.force_nsplit > 0) nsplit = std::min(G.force_nsplit, ntiles);   // G6_B200_NSPLIT: experiments only
    int tps = (ntiles + nsplit - 1) / nsplit;
    nsplit = (ntiles + tps - 1) / tps;

    ForceArgs a{};
    // G.win_lo > 0: the launch works on the window [win_lo, win_lo + nj) of the slots (sharded Hermite
    // step on a replicated state); win_lo is a multiple of TILE
    a.jA = G.js.A + G.win_lo; a.jB = G.js.B + G.win_lo; a.jC = G.js.C + G.win_lo; a.jL = G.js.L + G.win_lo;
    a.jG = G.js.gbb + (size_t)(G.win_lo / TILE) * TBOX;
    a.iA = ib.A; a.iB = ib.B; a.iC = ib.C; a.iD = ib.D;
    a.conf = ib.conf;
    a.iperm = ib.iperm;
    a.stats = G.d_stats;
    a.ni = ni; a.nj = nj;
    a.tiles_per_split = tps; a.nsplit = nsplit;
    a.ni_pad = ni;
    a.j_offset = G.j_offset;
    a.slot0 = G.win_lo;
    a.js = G.js;
    a.ord = G.ord;
    // many splits of few i-blocks: sum the partials with a kernel of its own (one warp per i, spread
    // over the SMs) instead of the last CTA -- except for the 4-particle shape, whose last CTA puts
    // 32 lanes on each i
    // the speculative kernel always stops at the partials when FP64 pairs are in use: they are evaluated by
    // close_pairs_kernel before reduce_partials_kernel adds everything up
    const bool fastv = (v == V_F2 || v == V_F4);
    const bool use_corr = (G.ord.kclose > 0.f);
    const bool defer = (fastv && use_corr) || ((nsplit > 32) && (n_iblocks * 8 <= slots) && (vi.ib > 4) && !herm);
    a.defer_reduce = defer ? 1 : 0;
    a.eps2 = eps2;
    if (nsplit > 1 || (fastv && use_corr)) ensure_partials((size_t)nsplit * ni);
    if (use_corr) {
        ensure_close_buffers(ni, fastv);
        a.corr = G.d_corr;
        if (fastv) {
            a.wl = G.d_wl;
            a.wl_count = G.d_wl_count;
            a.wl_cap = (unsigned int)G.wl_cap;
        }
    }
    a.part_sum = G.part_sum; a.part_key = G.part_key;
    a.tickets = G.tickets;
    a.out_sum = out_sum; a.out_key = out_key; a.out_nnid = out_nnid;
    a.ngb_cnt = G.d_ngb_cnt; a.ngb_list = G.d_ngb_list; a.ngb_cap = G.ngb_cap;
    a.n_mirror = G.mir_n;
    for (int m = 0; m < G.mir_n; m++) {
        a.m_sum[m] = G.mir_sum[m];
        a.m_key[m] = G.mir_key[m];
        a.m_id[m] = G.mir_id[m];
    }
    dim3 grid(nsplit, n_iblocks);
    const int reduce_ctas = (ni + 7) / 8;   // one warp per i
    if (flag_seq) {
        a.done_counter = G.d_done;
        a.host_flag = G.dev_h_flag;
        a.flag_seq = flag_seq;
        a.done_expected = defer ? (unsigned)reduce_ctas : (unsigned)n_iblocks;
    }
    if (v == V_F2 || v == V_F4) {
        // pre-pass of the speculative kernel: id-table lookup and nearest-neighbour bound of every i-particle
        near_kernel<<<(ni + 7) / 8, 256, 0, G.stream>>>(a, G.near_w);   // one warp per particle
        CK(cudaGetLastError());
        G.launches++;
    }
    if (herm) {
        a.herm = *herm;
        if (v == V_T1) launch_herm<1, 4, false, 2>(a, grid, G.stream);
        else if (v == V_W1) launch_herm<1, 32, false, 2>(a, grid, G.stream);
        else launch_herm<2, 32, true, 2>(a, grid, G.stream);
    } else if (inline_src) {
        if (ni <= 64) {
            InlineI<64> ii;
            fill_inline(ii, inline_src, ni);
            if (v == V_T1) launch_inline<1, 4, false, 2, 64>(a, grid, ii, G.stream);
            else if (v == V_W1) launch_inline<1, 32, false, 2, 64>(a, grid, ii, G.stream);
            else launch_inline<2, 32, true, 2, 64>(a, grid, ii, G.stream);
        } else {
            InlineI<384> ii;
            fill_inline(ii, inline_src, ni);
            if (v == V_T1) launch_inline<1, 4, false, 2, 384>(a, grid, ii, G.stream);
            else if (v == V_W1) launch_inline<1, 32, false, 2, 384>(a, grid, ii, G.stream);
            else launch_inline<2, 32, true, 2, 384>(a, grid, ii, G.stream);
        }
    } else
    switch (v) {
        case V_S4: launch_variant<4, 256, false, 1>(a, grid, nn, list, G.stream); break;
        case V_S2: launch_variant<2, 256, false, 2>(a, grid, nn, list, G.stream); break;
        case V_S1: launch_variant<1, 256, false, 2>(a, grid, nn, list, G.stream); break;
        case V_W1: launch_variant<1, 32, false, 2>(a, grid, nn, list, G.stream); break;
        case V_T1: launch_variant<1, 4, false, 2>(a, grid, nn, list, G.stream); break;
        case V_P4: launch_variant<4, 256, true, 1>(a, grid, nn, list, G.stream); break;
        case V_P2: launch_variant<2, 256, true, 2>(a, grid, nn, list, G.stream); break;
        case V_P2W: launch_variant<2, 32, true, 2>(a, grid, nn, list, G.stream); break;
        case V_F4: launch_fast<4, 1>(a, grid, nn, G.stream); break;
        case V_F2: launch_fast<2, 2>(a, grid, nn, G.stream); break;
        default:
            fprintf(stderr, "g6_b200: FATAL unknown force variant %d\n", v);
            exit(-1);
    }
    G.launches++;
    if (fastv && use_corr) {
        close_pairs_kernel<<<G.sm_count * 4, 256, 0, G.stream>>>(a);
        CK(cudaGetLastError());
        G.launches++;
    }
    if (defer) {
        reduce_partials_kernel<<<reduce_ctas, 256, 0, G.stream>>>(a, nn ? 1 : 0);
        CK(cudaGetLastError());
        G.launches++;
    }
}

bool is_fast_variant(int v) { return v == V_F2 || v == V_F4; }

void free_all()
{
    for (int k = 0; k < 7; k++) { dev_free(G.js.q[k]); dev_free(G.js2.q[k]); }
    dev_free(G.js.ia); dev_free(G.js2.ia); dev_free(G.js.near2); dev_free(G.js2.near2); dev_free(G.js.capr2); G.js2.capr2 = nullptr;
    dev_free(G.js.A); dev_free(G.js.B); dev_free(G.js.C); dev_free(G.js.L); dev_free(G.js.gbb);
    dev_free(G.js.slot_of); dev_free(G.addr_of); dev_free(G.addr_of2);
    dev_free(G.d_keys); dev_free(G.d_keys_tmp); dev_free(G.d_vals); dev_free(G.d_vals_tmp);
    if (G.d_sort_tmp) cudaFree(G.d_sort_tmp);
    G.d_sort_tmp = nullptr; G.sort_tmp_bytes = 0;
    dev_free(G.d_box); host_free(G.h_box);
    dev_free(G.d_hash); G.hash_size = 0;
    G.ord = OrderInfo{};
    G.order_valid = false; G.order_nj = -1; G.ids_dirty = false;
    G.updates_since_order = G.updates_since_force = 0;
    G.h_slot_of.clear(); G.h_id.clear(); G.h_perm.clear();
    dev_free(G.d_conf); dev_free(G.d_conf2); dev_free(G.d_stats);
    dev_free(G.d_wl); dev_free(G.d_wl_count); dev_free(G.d_corr);
    G.wl_cap = G.corr_cap = 0;
    dev_free(G.d_ikey); dev_free(G.d_ikey_tmp); dev_free(G.d_iperm); dev_free(G.d_iperm_tmp);
    G.isort_cap = 0;
    dev_free(G.d_up);
    for (int b = 0; b < 2; b++) {
        host_free(G.h_up2[b]);
        if (G.up_done[b]) cudaEventDestroy(G.up_done[b]);
        G.up_done[b] = nullptr;
    }
    G.h_up = nullptr; G.up_cur = 0;
    host_free(G.h_i); dev_free(G.d_i); dev_free(G.d_i2);
    dev_free(G.d_sum); dev_free(G.d_key);
    host_free(G.h_sum);
    host_free(G.h_flag);
    dev_free(G.d_done);
    {
        Context::Hermite &H = G.herm;
        host_free(H.h_ilist); host_free(H.h_olddt); host_free(H.h_outdt); host_free(H.h_outpot); host_free(H.h_outnn);
        dev_free(H.d_pred); dev_free(H.d_i); dev_free(H.d_sum); dev_free(H.d_key); dev_free(H.d_nnid);
        dev_free(H.d_ilist); dev_free(H.d_olddt); dev_free(H.d_conf);
        H.cap = 0;
        H.time.clear(); H.dt.clear();
        H.initialised = false;
        H.shard_lo = H.shard_hi = 0;
    }
    G.dev_h_sum = nullptr; G.dev_h_flag = nullptr; G.dev_h_up2[0] = G.dev_h_up2[1] = nullptr;
    G.cur_direct = false; G.i_on_device = false;
    dev_free(G.part_sum); dev_free(G.part_key); dev_free(G.tickets);
    dev_free(G.d_ngb_cnt); dev_free(G.d_ngb_list);
    dev_free(G.d_sum2); dev_free(G.d_key2); dev_free(G.d_nnid2);
    host_free(G.h_ngb_cnt); host_free(G.h_ngb_list);
    G.capacity = 0; G.nj_hi = 0; G.up_cap = 0; G.up_n = 0; G.part_records = 0; G.i2_cap = 0;
    G.pending_of_addr.clear();
    G.predicted_nj = -1; G.j_dirty = false; G.pending = false;
    G.ngb_valid = G.ngb_built = G.ngb_fetched = false;
}

void stage_j(int address, int index, double tj, double mass, const double *j6, const double *a2, const double *v,
             const double *x, int key_address = -1)
{
    if (address < 0) {
        fprintf(stderr, "g6_b200: FATAL g6_set_j_particle address %d < 0\n", address);
        exit(-1);
    }
    ensure_capacity(address + 1);
    int pos = G.pending_of_addr[address];
    if (pos < 0) {  // last write wins within a batch (sapporo.cpp:83-110)
        ensure_up_cap(G.up_n + 1);
        pos = G.up_n++;
        G.pending_of_addr[address] = pos;
    }
    JUpdate &u = G.h_up[pos];
    for (int k = 0; k < 3; k++) {
        u.x[k] = x[k];
        u.v[k] = v[k];
        u.a[k] = 2.0 * a2[k];   // a2 = acc/2   (sapporo.cpp:94)
        u.j[k] = 6.0 * j6[k];   // j6 = jerk/6  (sapporo.cpp:95)
    }
    u.t = tj;
    u.m = mass;
    u.id = index;
    u.slot = G.h_slot_of[address];
    u.kaddr = key_address >= 0 ? key_address : address;
    u.addr = address;
    if (G.h_id[address] != index) {   // a new particle or a new id at this address: the id table is out of date
        G.h_id[address] = index;
        G.ids_dirty = true;
    }
    G.updates_since_order++;
    G.updates_since_force++;
    if (address + 1 > G.nj_hi) G.nj_hi = address + 1;
    // big batches (the caller is sending back a large block, or loading the system) go out while the
    // caller is still staging the rest, so the next force call does not start with a multi-MB upload
    if (G.eager_flush > 0 && G.up_n >= G.eager_flush) flush_updates();
}

// Chunk size of the device path: unlike the ABI path it is not tied to g6_npipes().  A launch of the
// speculative kernel should hand out four waves of CTAs (148 SMs x 2 CTAs x 4 = 1184 = i-blocks x
// j-splits, see launch_force) and give every CTA ~128 j-tiles, so that its prologue, neighbour re-scan
// and split reduction stay amortised: with all 1M j on one GPU that is 37 i-blocks x 32 splits
// (18944 i per launch); a rank that holds 1/8 of the j takes 296 i-blocks x 4 splits (151552 i).
int device_chunk(int ni, int nj)
{
    int chunk = G.npipes;
    if (G.variant == V_AUTO && ni > G.npipes) {
        const VariantInfo &vi = variant_info(V_F2);
        const int slots = G.sm_count * vi.ctas_per_sm;
        const int ntiles = std::max(1, (nj + TILE - 1) / TILE);
        int want_split = std::max(1, ntiles / 128);
        int split = 1;   // largest divisor of 4*slots that does not exceed want_split
        for (int sdiv = 1; sdiv <= want_split; sdiv++)
            if ((4 * slots) % sdiv == 0) split = sdiv;
        chunk = vi.ib * (4 * slots / split);
    }
    return chunk;
}

}  // namespace

// ===========================================================================
// Part 1: the GRAPE-6 ABI
// ===========================================================================
extern "C" {

int get_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// ---- multi-device layer ----------------------------------------------------------------------------
// G6_B200_DEVICES = n > 1 (or "all"): g6_open_ opens n devices in this one process and drives them as ONE
// g6 device -- the device-side form of ph4's MPI mode (gpu.cc:40,56-59 + idata.cc:284-313) without MPI.
// Like ph4's ranks, every device holds ALL particles (g6_set_j_particle_ goes to every device; 180 GB of HBM
// hold tens of millions), so the Morton order of the j-memory is the same everywhere.  A big i-block is sent to
// all devices, device k sums over its window of the SLOTS (contiguous tiles, ~1/n of the j each), the force
// kernels store their partial sums into device 0's exchange buffer over NVLink (peer access enabled directly)
// and device 0 combines them (sum, min key, id of the winner) for g6calc_lasthalf_.  A small i-block (the
// block-timestep regime, where launch latency is everything) runs on device 0 alone, exactly like the one-device
// library; the other devices keep their pending j-updates staged until they are needed.
struct Multi {
    int n = 0;           // devices open (0: library closed)
    int dev_now = -1;    // device the CUDA runtime currently has selected
    double spread_min = 1.0e8;   // G6_B200_SPREAD_MIN: ni x nj per device from which a block goes to all devices
} M;

static void use(int k)
{
    g_cur = &g_ctx[k];
    if (M.dev_now != G.device) {
        CK(cudaSetDevice(G.device));
        M.dev_now = G.device;
    }
}
// window of slots [lo, hi) device k sums over when a block is spread over nd devices (whole tiles)
static inline void slot_window(int nj, int k, int nd, int &lo, int &hi)
{
    const int ntiles = (nj + TILE - 1) / TILE;
    const int per = (ntiles + nd - 1) / nd;
    lo = std::min(nj, k * per * TILE);
    hi = std::min(nj, (k + 1) * per * TILE);
}

static void open_context(int dev)   // g_cur selected by the caller
{
    G.device = dev;
    CK(cudaSetDevice(dev));
    M.dev_now = dev;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) {
        fprintf(stderr, "g6_b200: FATAL device %d is sm_%d%d; this library is built for sm_100a only\n", dev,
                prop.major, prop.minor);
        exit(-1);
    }
    G.sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&G.own_stream, cudaStreamNonBlocking));
    G.stream = G.own_stream;
    G.npipes = std::max(1, env_int("G6_B200_NPIPES", 16384));
    G.ngb_cap = std::max(1, env_int("G6_B200_NGB_CAP", 1024));
    G.variant = env_int("G6_B200_VARIANT", V_AUTO);
    G.refine = env_int("G6_B200_REFINE", 1);
    G.force_nsplit = env_int("G6_B200_NSPLIT", 0);
    {
        const char *e = getenv("G6_B200_KCLOSE");
        G.kclose = (e && *e) ? (float)atof(e) : 32.f;
        e = getenv("G6_B200_FARC");
        G.farc = (e && *e) ? (float)atof(e) : 0.125f;
        G.near_w = std::max(1, env_int("G6_B200_NEAR_WINDOW", 32));
        G.gmax = std::max(1, env_int("G6_B200_GMAX", 256));
        G.wl_per_i = std::max(1, env_int("G6_B200_WL_PER_I", 1024));
    }
    host_alloc(G.h_i, (size_t)4 * G.npipes);
    dev_alloc(G.d_i, (size_t)4 * G.npipes);
    dev_alloc(G.d_conf, (size_t)G.npipes);
    dev_alloc(G.d_i2, (size_t)4 * G.npipes);
    dev_alloc(G.d_conf2, (size_t)G.npipes);
    G.i2_cap = G.npipes;
    dev_alloc(G.d_sum, (size_t)8 * G.npipes);    // [7n doubles][n ints] per i-block
    dev_alloc(G.d_key, (size_t)G.npipes);
    host_alloc(G.h_sum, (size_t)8 * G.npipes);
    G.dev_h_sum = dev_alias(G.h_sum);
    host_alloc(G.h_flag, 8);
    G.h_flag[0] = 0;
    G.dev_h_flag = dev_alias(G.h_flag);
    G.flag_seq = 0;
    dev_alloc(G.d_done, 1);
    CK(cudaMemsetAsync(G.d_done, 0, sizeof(unsigned int), G.stream));
    G.direct_max = std::min(G.npipes, std::max(0, env_int("G6_B200_DIRECT_MAX", 2048)));
    G.inline_max = std::min(384, std::max(0, env_int("G6_B200_INLINE_MAX", 384)));
    G.zc_up_max = std::max(0, env_int("G6_B200_ZC_UPDATES", 512));
    G.fuse = env_int("G6_B200_FUSE", 1);
    G.eager_flush = env_int("G6_B200_EAGER_FLUSH", 8192);
    G.trace = env_int("G6_B200_TRACE", 0);
    for (double &t : G.tr) t = 0;
    G.tr_calls = 0;
    dev_alloc(G.tickets, 65536);
    CK(cudaMemsetAsync(G.tickets, 0, 65536 * sizeof(unsigned int), G.stream));
    G.ti = 0.0;
    G.predicted_nj = -1;
    G.j_offset = 0;
    G.open = true;
    if (env_int("G6_B200_VERBOSE", 0))
        fprintf(stderr, "g6_b200: open device %d (%s, %d SMs), npipes %d\n", dev, prop.name, G.sm_count, G.npipes);
}

static void close_context()
{
    if (!G.open) return;
    CK(cudaSetDevice(G.device));
    M.dev_now = G.device;
    CK(cudaStreamSynchronize(G.stream));
    g6x_peer_detach();
    if (G.trace && G.tr_calls)
        fprintf(stderr,
                "g6_b200 trace: %lld force calls; mean us per call: flush %.2f predict %.2f pack %.2f h2d %.2f "
                "launch %.2f d2h-issue %.2f wait %.2f unpack %.2f; %lld order rebuilds\n",
                G.tr_calls, 1e6 * G.tr[0] / G.tr_calls, 1e6 * G.tr[1] / G.tr_calls, 1e6 * G.tr[2] / G.tr_calls,
                1e6 * G.tr[3] / G.tr_calls, 1e6 * G.tr[4] / G.tr_calls, 1e6 * G.tr[5] / G.tr_calls,
                1e6 * G.tr[6] / G.tr_calls, 1e6 * G.tr[7] / G.tr_calls, G.order_rebuilds);
    free_all();
    if (G.i_ready) cudaEventDestroy(G.i_ready);
    G.i_ready = nullptr;
    if (G.own_stream) cudaStreamDestroy(G.own_stream);
    G.own_stream = G.stream = nullptr;
    G.open = false;
}

static void peer_group_setup();   // below, with the exchange helpers

int g6_open_(int *id)
{
    int ndev = get_device_count();
    if (ndev <= 0) {
        fprintf(stderr, "g6_b200: FATAL no CUDA device found (this library has no CPU fallback)\n");
        exit(-1);
    }
    int dev = id ? *id : 0;
    if (dev < 0 || dev >= ndev) {
        // ph4's standalone driver passes an uninitialised gpu_id (jdata.h:137-170) and Fortran callers ignore
        // the return value: fold the id onto the devices that exist instead of failing.  G6_B200_STRICT_DEVICE=1
        // restores the reference's answer (sapporo.cpp:31-34: -1 for a bad id).
        if (env_int("G6_B200_STRICT_DEVICE", 0)) {
            fprintf(stderr, "g6_b200: g6_open: no CUDA device with id %d (%d present)\n", dev, ndev);
            return -1;
        }
        const int folded = ((dev % ndev) + ndev) % ndev;
        fprintf(stderr, "g6_b200: g6_open: no CUDA device with id %d (%d present): using device %d\n", dev, ndev, folded);
        dev = folded;
    }
    int want = 1;
    {
        const char *e = getenv("G6_B200_DEVICES");
        if (e && *e) want = (strcmp(e, "all") == 0) ? ndev : atoi(e);
        want = std::max(1, std::min(want, std::min(ndev, MAX_DEVICES)));
    }
    if (M.n > 0) {
        if (M.n == want && g_ctx[0].device == dev) return 0;
        int d = g_ctx[0].device;
        g6_close_(&d);
    }
    for (int k = 0; k < want; k++) {
        g_cur = &g_ctx[k];
        open_context((dev + k) % ndev);
    }
    M.n = want;
    {
        const char *e = getenv("G6_B200_SPREAD_MIN");
        M.spread_min = (e && *e) ? atof(e) : 1.0e8;
    }
    if (want > 1) peer_group_setup();
    use(0);
    return 0;
}

int g6_close_(int *id)
{
    (void)id;
    for (int k = 0; k < M.n; k++) {
        g_cur = &g_ctx[k];
        close_context();
    }
    M.n = 0;
    g_cur = &g_ctx[0];
    return 0;
}

int g6_npipes_(void)
{
    if (g_ctx[0].open) return g_ctx[0].npipes;
    return std::max(1, env_int("G6_B200_NPIPES", 16384));
}

int g6_set_tunit_(void *unused) { (void)unused; return 0; }
int g6_set_xunit_(void *unused) { (void)unused; return 0; }

int g6_set_ti_(int *id, double *ti)
{
    (void)id;
    require_open("g6_set_ti_");
    for (int k = 0; k < std::max(1, M.n); k++) g_ctx[k].ti = *ti;
    return 0;
}

int g6_set_j_particle_(int *cluster_id, int *address, int *index, double *tj, double *dtj, double *mass,
                       double k18[3], double j6[3], double a2[3], double v[3], double x[3])
{
    (void)cluster_id; (void)dtj; (void)k18;
    require_open("g6_set_j_particle_");
    if (M.n > 1) {
        for (int k = 0; k < M.n; k++) {
            g_cur = &g_ctx[k];
            // (stage_j only touches the device when a buffer grows or a big batch is flushed early)
            const bool touches = (*address + 1 > G.capacity) || (G.up_n + 1 > G.up_cap) ||
                                 (G.eager_flush > 0 && G.up_n + 1 >= G.eager_flush);
            if (touches) use(k);
            stage_j(*address, *index, *tj, *mass, j6, a2, v, x);
        }
        g_cur = &g_ctx[0];
        return 0;
    }
    stage_j(*address, *index, *tj, *mass, j6, a2, v, x);
    return 0;
}

// Morton order of an i-block on the host (radix sort, 3 passes of 10 bits): G.h_perm[k] = caller index of
// the particle at packed position k.  xi are the caller's coordinates.
static void host_morton_perm(int n, double xi[][3])
{
    const OrderInfo &o = G.ord;
    G.h_perm.resize(n);
    G.h_perm2.resize(n);
    G.h_ikey.resize(n);
    G.h_ikey2.resize(n);
    auto spread = [](unsigned v) {
        v = (v | (v << 16)) & 0x030000ffu;
        v = (v | (v << 8)) & 0x0300f00fu;
        v = (v | (v << 4)) & 0x030c30c3u;
        v = (v | (v << 2)) & 0x09249249u;
        return v;
    };
    for (int i = 0; i < n; i++) {
        unsigned c[3];
        for (int d = 0; d < 3; d++) {
            float f = ((float)(xi[i][d] - G.js.x0[d]) - o.blo[d]) * o.binv[d];
            f = std::min(std::max(f, 0.f), 1023.f);
            c[d] = (unsigned)f;
        }
        G.h_ikey[i] = spread(c[0]) | (spread(c[1]) << 1) | (spread(c[2]) << 2);
        G.h_perm[i] = i;
    }
    unsigned *ka = G.h_ikey.data(), *kb = G.h_ikey2.data();
    int *va = G.h_perm.data(), *vb = G.h_perm2.data();
    for (int pass = 0; pass < 3; pass++) {
        unsigned cnt[1025] = {0};
        const int sh = 10 * pass;
        for (int i = 0; i < n; i++) cnt[((ka[i] >> sh) & 1023u) + 1]++;
        for (int b = 0; b < 1024; b++) cnt[b + 1] += cnt[b];
        for (int i = 0; i < n; i++) {
            const unsigned pos = cnt[(ka[i] >> sh) & 1023u]++;
            kb[pos] = ka[i];
            vb[pos] = va[i];
        }
        std::swap(ka, kb);
        std::swap(va, vb);
    }
    if (va != G.h_perm.data()) G.h_perm.swap(G.h_perm2);   // three passes: the result sits in the second pair
}

struct ExchangeSlots;
static void gather_begin_all(int ni);
static void gather_launch(int k, int nj, int ni, const IBlock &ib, float eps2, double *out_sum, u64 *out_key,
                          int *out_nnid, const float4 *inline_src, unsigned long long flag_seq);
static void gather_finish_all(int ni, bool direct);

void g6calc_firsthalf_(int *cluster_id, int *nj, int *ni, int index[], double xi[][3], double vi[][3],
                       double aold[][3], double j6old[][3], double phiold[], double *eps2, double h2[])
{
    (void)cluster_id; (void)aold; (void)j6old; (void)phiold;
    require_open("g6calc_firsthalf_");
    const int nd = std::max(1, M.n);
    use(0);
    Context &R = g_ctx[0];   // root: holds the packed i-block and receives the results
    int n = *ni;
    if (n > R.npipes || n < 0) {
        fprintf(stderr, "g6_b200: FATAL g6calc_firsthalf ni = %d exceeds g6_npipes() = %d\n", n, R.npipes);
        exit(-1);
    }
    if (R.pending)
        for (int k = 0; k < nd; k++) {   // a firsthalf without its lasthalf
            use(k);
            CK(cudaStreamSynchronize(G.stream));
        }
    use(0);
    double t0 = G.trace ? wall() : 0.0, t1 = 0.0;
    // a big block goes to all devices (each sums over its window of the slots), a small one to device 0 alone
    const int njc = std::min(*nj, std::max(R.capacity, 0));
    const bool spread = (nd > 1) && ((double)n * (double)njc >= M.spread_min * nd);
    R.cur_spread = spread;
    // scatter (small batches: fused with the predictor, records read from mapped pinned memory) + predict;
    // the Morton order of the j-memory is rebuilt first when a bulk load or new ids made it stale -- on ALL
    // devices at once (same state in, same permutation out: the slot windows must mean the same particles
    // everywhere), with device 0's origin
    const bool stale = order_stale(*nj);
    for (int k = 0; k < nd; k++) {
        if (!stale && !spread && k > 0) break;
        use(k);
        if (stale) {
            flush_updates();
            rebuild_order(*nj);
        }
        if (k == 0 || spread) predict_with_updates(*nj);
    }
    use(0);
    G6_TR(0)
    G6_TR(1)
    const bool inl = (n > 0) && (n <= G.inline_max) && (G.variant == V_AUTO);
    // Morton-sort the block when it goes to the speculative kernel (whose warps want 64 neighbouring particles)
    int wlo = 0, whi = njc;
    if (spread) slot_window(njc, 0, nd, wlo, whi);
    const bool sorted = (n > 1) && is_fast_variant(choose_variant(n, whi - wlo)) && G.ord.nkeys > 0;
    if (sorted) host_morton_perm(n, xi); else G.h_perm.clear();
    // pack the i-block: double -> double-single relative to the origin (sapporo.cpp:125-134); the four
    // float4 streams are laid out back to back with stride n, so that they cross PCIe in ONE copy
    float4 *A = G.h_i, *B = G.h_i + n, *C = G.h_i + 2 * (size_t)n, *D = G.h_i + 3 * (size_t)n;
    bool any_h2 = false;
    const double ox = G.js.x0[0], oy = G.js.x0[1], oz = G.js.x0[2];
    for (int k = 0; k < n; k++) {
        const int i = sorted ? G.h_perm[k] : k;
        double x = xi[i][0] - ox, y = xi[i][1] - oy, z = xi[i][2] - oz;
        float xh = (float)x, yh = (float)y, zh = (float)z;
        float vxh = (float)vi[i][0], vyh = (float)vi[i][1], vzh = (float)vi[i][2];
        float hh = h2 ? (float)h2[i] : 0.f;
        any_h2 |= (hh > 0.f);
        A[k] = make_float4(xh, yh, zh, hh);
        union { int i; float f; } cv;
        cv.i = index[i];
        B[k] = make_float4((float)(x - (double)xh), (float)(y - (double)yh), (float)(z - (double)zh), cv.f);
        C[k] = make_float4(vxh, vyh, vzh, 0.f);
        D[k] = make_float4((float)(vi[i][0] - (double)vxh), (float)(vi[i][1] - (double)vyh),
                           (float)(vi[i][2] - (double)vzh), 0.f);
    }
    G6_TR(2)
    // small i-blocks travel in the kernel parameters (no H2D copy); their results are written by the
    // kernel into mapped host memory and announced by a flag (no D2H copy, no stream synchronisation)
    const bool direct = (n > 0) && (n <= G.direct_max);
    R.i_on_device = !inl;
    R.cur_direct = direct;
    R.cur_ni = n;
    R.cur_nj = *nj;
    R.cur_eps2 = (float)*eps2;
    R.cur_any_h2 = any_h2;
    R.pending = true;
    R.ngb_valid = R.ngb_fetched = false;
    // The force kernel starts here, asynchronously (GRAPE's firsthalf/lasthalf split exists for this
    // overlap): always with the nearest-neighbour search, which costs ~1 % and is simply not copied
    // back by g6calc_lasthalf_.  Neighbour-sphere lists are NOT built here: ph4 and phiGRAPE pass
    // h2 = eps2 on every call (gpu.cc:266,324; gravity.F:72) and never read the lists, so they are
    // built on demand by g6_read_neighbour_list_.
    // outputs: [7n doubles][n ints] back to back, so that they come back in ONE copy
    if (n > 0 && !spread) {
        if (!inl) CK(cudaMemcpyAsync(G.d_i, G.h_i, sizeof(float4) * 4 * (size_t)n, cudaMemcpyHostToDevice, G.stream));
        G6_TR(3)
        double *out = direct ? G.dev_h_sum : G.d_sum;
        launch_force(R.cur_nj, n, iblock_of(G.d_i, n, G.d_conf), R.cur_eps2, true, false, out, G.d_key,
                     reinterpret_cast<int *>(out + 7 * (size_t)n), inl ? G.h_i : nullptr,
                     direct ? ++G.flag_seq : 0ull);
    } else if (n > 0) {
        // every device gets the block (from the root's pinned copy) and sums over its window of the slots; the
        // partials meet in the root's exchange buffer, where the root's combine kernel writes the totals (and
        // raises the flag)
        gather_begin_all(n);
        if (direct) ++R.flag_seq;
        if (!inl) {   // the block crosses PCIe once (to the root) and reaches the other devices over NVLink
            use(0);
            CK(cudaMemcpyAsync(R.d_i, R.h_i, sizeof(float4) * 4 * (size_t)n, cudaMemcpyHostToDevice, R.stream));
            if (!R.i_ready) CK(cudaEventCreateWithFlags(&R.i_ready, cudaEventDisableTiming));
            CK(cudaEventRecord(R.i_ready, R.stream));
        }
        for (int k = 0; k < nd; k++) {
            use(k);
            if (!inl && k > 0) {
                CK(cudaStreamWaitEvent(G.stream, R.i_ready, 0));
                CK(cudaMemcpyPeerAsync(G.d_i, G.device, R.d_i, R.device, sizeof(float4) * 4 * (size_t)n, G.stream));
            }
            int lo, hi;
            slot_window(njc, k, nd, lo, hi);
            G.win_lo = lo;
            gather_launch(k, hi - lo, n, iblock_of(G.d_i, n, G.d_conf), R.cur_eps2, nullptr, nullptr, nullptr,
                          inl ? R.h_i : nullptr, 0ull);
            G.win_lo = 0;
            G.i_on_device = !inl;
        }
        gather_finish_all(n, direct);
        use(0);
        G6_TR(3)
    }
    G6_TR(4)
    if (G.trace) G.tr_calls++;
}

static int lasthalf_common(int nj, int ni, double acc[][3], double jerk[][3], double pot[], int *inn)
{
    require_open("g6calc_lasthalf_");
    use(0);
    if (!G.pending || ni != G.cur_ni) {
        fprintf(stderr, "g6_b200: FATAL g6calc_lasthalf without matching g6calc_firsthalf (ni %d vs %d)\n", ni,
                G.cur_ni);
        exit(-1);
    }
    (void)nj;
    bool nn = (inn != nullptr);
    if (ni > 0) {
        double t0 = G.trace ? wall() : 0.0, t1 = 0.0;
        if (G.cur_direct) {
            wait_flag(G.flag_seq, "force kernel");
            G6_TR(6)
        } else {
            const size_t bytes = sizeof(double) * 7 * (size_t)ni + (nn ? sizeof(int) * (size_t)ni : 0);
            CK(cudaMemcpyAsync(G.h_sum, G.d_sum, bytes, cudaMemcpyDeviceToHost, G.stream));
            G6_TR(5)
            CK(cudaStreamSynchronize(G.stream));
            G6_TR(6)
        }
        const int *h_nnid = reinterpret_cast<const int *>(G.h_sum + 7 * (size_t)ni);
        const bool sorted = !G.h_perm.empty();
        for (int k = 0; k < ni; k++) {
            const int i = sorted ? G.h_perm[k] : k;
            const double *s = G.h_sum + (size_t)7 * k;
            acc[i][0] = s[0]; acc[i][1] = s[1]; acc[i][2] = s[2];
            jerk[i][0] = s[3]; jerk[i][1] = s[4]; jerk[i][2] = s[5];
            pot[i] = -s[6];
            if (nn) inn[i] = h_nnid[k];
        }
        G6_TR(7)
    }
    G.pending = false;
    G.ngb_valid = nn && G.cur_any_h2 && ni > 0;   // lists of this i-block can be built on demand
    G.ngb_built = false;
    G.ngb_fetched = false;
    return 0;
}

int g6calc_lasthalf_(int *cluster_id, int *nj, int *ni, int index[], double xi[][3], double vi[][3], double *eps2,
                     double h2[], double acc[][3], double jerk[][3], double pot[])
{
    (void)cluster_id; (void)index; (void)xi; (void)vi; (void)eps2; (void)h2;
    return lasthalf_common(*nj, *ni, acc, jerk, pot, nullptr);
}

int g6calc_lasthalf2_(int *cluster_id, int *nj, int *ni, int index[], double xi[][3], double vi[][3], double *eps2,
                      double h2[], double acc[][3], double jerk[][3], double pot[], int inn[])
{
    (void)cluster_id; (void)index; (void)xi; (void)vi; (void)eps2; (void)h2;
    return lasthalf_common(*nj, *ni, acc, jerk, pot, inn);
}

int g6_initialize_jp_buffer_(int *cluster_id, int *buf_size) { (void)cluster_id; (void)buf_size; return 0; }
int g6_flush_jp_buffer_(int *cluster_id) { (void)cluster_id; return 0; }
int g6_reset_(int *cluster_id) { (void)cluster_id; return 0; }
int g6_reset_fofpga_(int *cluster_id) { (void)cluster_id; return 0; }

// second pass over the captured i-block on the current device (still in d_i; j state unchanged since its
// lasthalf2) with the list-building variant of the masked kernel; forces go to scratch
static void build_lists_here(int nj_local, int ni, float eps2)
{
    if (!G.d_ngb_cnt) {
        dev_alloc(G.d_ngb_cnt, (size_t)G.npipes);
        dev_alloc(G.d_ngb_list, (size_t)G.npipes * G.ngb_cap);
        host_alloc(G.h_ngb_cnt, (size_t)G.npipes);
        host_alloc(G.h_ngb_list, (size_t)G.npipes * G.ngb_cap);
        dev_alloc(G.d_sum2, (size_t)7 * G.npipes);
        dev_alloc(G.d_key2, (size_t)G.npipes);
        dev_alloc(G.d_nnid2, (size_t)G.npipes);
    }
    if (!G.i_on_device) {   // the block went out in the kernel parameters: d_i was never written
        CK(cudaMemcpyAsync(G.d_i, g_ctx[0].h_i, sizeof(float4) * 4 * (size_t)ni, cudaMemcpyHostToDevice, G.stream));
        G.i_on_device = true;
    }
    CK(cudaMemsetAsync(G.d_ngb_cnt, 0, sizeof(int) * ni, G.stream));
    launch_force(nj_local, ni, iblock_of(G.d_i, ni, G.d_conf), eps2, true, true, G.d_sum2, G.d_key2, G.d_nnid2);
}

int g6_read_neighbour_list_(int *cluster_id)
{
    (void)cluster_id;
    require_open("g6_read_neighbour_list_");
    Context &R = g_ctx[0];
    if (!R.ngb_valid) {
        R.ngb_fetched = false;
        return 0;  // no lists were requested (all h2 <= 0 or lasthalf without nn)
    }
    const int nd = R.cur_spread ? M.n : 1;   // devices that hold the i-block
    const int ni = R.cur_ni;
    if (!R.ngb_built) {
        const int njc = std::min(R.cur_nj, std::max(R.capacity, 0));
        for (int k = 0; k < nd; k++) {
            use(k);
            int lo = 0, hi = njc;
            if (nd > 1) slot_window(njc, k, nd, lo, hi);
            G.win_lo = lo;
            build_lists_here(hi - lo, ni, R.cur_eps2);
            G.win_lo = 0;
        }
        R.ngb_built = true;
    }
    for (int k = 0; k < nd; k++) {
        use(k);
        CK(cudaMemcpyAsync(G.h_ngb_cnt, G.d_ngb_cnt, sizeof(int) * ni, cudaMemcpyDeviceToHost, G.stream));
        CK(cudaMemcpyAsync(G.h_ngb_list, G.d_ngb_list, sizeof(int) * (size_t)ni * G.ngb_cap, cudaMemcpyDeviceToHost,
                           G.stream));
    }
    for (int k = 0; k < nd; k++) {
        use(k);
        CK(cudaStreamSynchronize(G.stream));
    }
    use(0);
    R.ngb_fetched = true;
    // caller index -> packed position (the block may have been Morton-sorted)
    R.h_perm2.assign(ni, 0);
    for (int k = 0; k < ni; k++) R.h_perm2[R.h_perm.empty() ? k : R.h_perm[k]] = k;
    int overflow = 0;
    for (int i = 0; i < ni; i++) {
        long long tot = 0;
        for (int k = 0; k < nd; k++) {
            tot += g_ctx[k].h_ngb_cnt[i];
            if (g_ctx[k].h_ngb_cnt[i] > g_ctx[k].ngb_cap) overflow = 1;
        }
        if (tot > R.ngb_cap) overflow = 1;
    }
    return overflow;
}

int g6_get_neighbour_list_(int *cluster_id, int *ipipe, int *maxlength, int *n_neighbours, int neighbour_list[])
{
    (void)cluster_id;
    require_open("g6_get_neighbour_list_");
    Context &R = g_ctx[0];
    int ip = *ipipe;
    if (ip < 0 || ip >= R.cur_ni) {
        fprintf(stderr, "g6_b200: FATAL g6_get_neighbour_list ipipe = %d >= ni = %d\n", ip, R.cur_ni);
        exit(-1);  // as sapporo.cpp:254-258
    }
    if (!R.ngb_valid || !R.ngb_fetched) {
        *n_neighbours = 0;
        return 0;
    }
    const int nd = R.cur_spread ? M.n : 1;
    const int pos = R.h_perm2[ip];
    static std::vector<int> merged;
    merged.clear();
    long long cnt = 0;
    bool truncated = false;
    for (int k = 0; k < nd; k++) {
        const Context &c = g_ctx[k];
        const int ck = c.h_ngb_cnt[pos];
        const int have = std::min(ck, c.ngb_cap);
        truncated |= (ck > c.ngb_cap);
        cnt += ck;
        merged.insert(merged.end(), c.h_ngb_list + (size_t)pos * c.ngb_cap, c.h_ngb_list + (size_t)pos * c.ngb_cap + have);
    }
    std::sort(merged.begin(), merged.end());
    const int ncopy = std::min((int)merged.size(), *maxlength);
    if (ncopy > 0) memcpy(neighbour_list, merged.data(), sizeof(int) * ncopy);
    *n_neighbours = (int)cnt;
    // overflow when the list does not fit with room to spare: nblen >= maxlength (sapporo.cpp:262-265; ph4 shrinks
    // h2 on it, gpu.cc:668-751)
    return (cnt >= *maxlength || truncated) ? 1 : 0;
}

// ---- by-value variants (lib/g6lib/g6lib.h:58-128) ---------------------------
int g6_open(int clusterid) { return g6_open_(&clusterid); }
int g6_close(int clusterid) { return g6_close_(&clusterid); }
int g6_npipes(void) { return g6_npipes_(); }
int g6_set_tunit(int newtunit) { (void)newtunit; return 0; }
int g6_set_xunit(int newxunit) { (void)newxunit; return 0; }
int g6_set_ti(int clusterid, double ti) { return g6_set_ti_(&clusterid, &ti); }
int g6_set_j_particle(int clusterid, int address, int index, double tj, double dtj, double mass, double a2by18[3],
                      double a1by6[3], double aby2[3], double v[3], double x[3])
{
    return g6_set_j_particle_(&clusterid, &address, &index, &tj, &dtj, &mass, a2by18, a1by6, aby2, v, x);
}
void g6calc_firsthalf(int clusterid, int nj, int ni, int index[], double xi[][3], double vi[][3], double fold[][3],
                      double jold[][3], double phiold[], double eps2, double h2[])
{
    g6calc_firsthalf_(&clusterid, &nj, &ni, index, xi, vi, fold, jold, phiold, &eps2, h2);
}
int g6calc_lasthalf(int clusterid, int nj, int ni, int index[], double xi[][3], double vi[][3], double eps2,
                    double h2[], double acc[][3], double jerk[][3], double pot[])
{
    return g6calc_lasthalf_(&clusterid, &nj, &ni, index, xi, vi, &eps2, h2, acc, jerk, pot);
}
int g6calc_lasthalf2(int clusterid, int nj, int ni, int index[], double xi[][3], double vi[][3], double eps2,
                     double h2[], double acc[][3], double jerk[][3], double pot[], int nnbindex[])
{
    return g6calc_lasthalf2_(&clusterid, &nj, &ni, index, xi, vi, &eps2, h2, acc, jerk, pot, nnbindex);
}
int g6_initialize_jp_buffer(int clusterid, int size) { (void)clusterid; (void)size; return 0; }
int g6_flush_jp_buffer(int clusterid) { (void)clusterid; return 0; }
void g6_reset(int devid) { (void)devid; }
int g6_reset_fofpga(int devid) { (void)devid; return 0; }
void g6_reinitialize(int clusterid) { (void)clusterid; }
int g6_get_number_of_pipelines(void) { return g6_npipes_(); }
int g6_read_neighbour_list(int clusterid) { return g6_read_neighbour_list_(&clusterid); }
int g6_get_neighbour_list(int clusterid, int ipipe, int maxlength, int *nblen, int nbl[])
{
    return g6_get_neighbour_list_(&clusterid, &ipipe, &maxlength, nblen, nbl);
}
static int g_sort_mode = 1;
void g6_set_neighbour_list_sort_mode(int mode) { g_sort_mode = mode; }
int g6_get_neighbour_list_sort_mode(void) { return g_sort_mode; }
int g6_set_overflow_flag_test_mode(int aflag, int jflag, int pflag) { (void)aflag; (void)jflag; (void)pflag; return 0; }
void force_j_particle_send(void)
{
    for (int k = 0; k < M.n; k++) {
        use(k);
        flush_updates();
    }
    if (M.n > 0) use(0);
}
int get_j_part_data(int addr, int nj, double *pos, double *vel, double *acc, double *jrk, double *ppos, double *pvel)
{
    require_open("get_j_part_data");
    if (addr < 0 || addr >= nj) return -1;
    use(0);
    const int la = addr;
    flush_updates();
    if (la >= G.capacity) {
        use(0);
        return -1;
    }
    const int slot = G.h_slot_of[la];
    double2 q[7];
    float4 abc[4];
    CK(cudaStreamSynchronize(G.stream));
    for (int k = 0; k < 7; k++) CK(cudaMemcpy(&q[k], G.js.q[k] + slot, sizeof(double2), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&abc[0], G.js.A + slot, sizeof(float4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&abc[1], G.js.B + slot, sizeof(float4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&abc[2], G.js.C + slot, sizeof(float4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&abc[3], G.js.L + slot, sizeof(float4), cudaMemcpyDeviceToHost));
    if (pos) { pos[0] = q[0].x; pos[1] = q[0].y; pos[2] = q[1].x; }
    if (vel) { vel[0] = q[2].x; vel[1] = q[2].y; vel[2] = q[3].x; }
    if (acc) { acc[0] = q[3].y; acc[1] = q[4].x; acc[2] = q[4].y; }
    if (jrk) { jrk[0] = q[5].x; jrk[1] = q[5].y; jrk[2] = q[6].x; }
    if (ppos)
        for (int d = 0; d < 3; d++) ppos[d] = (double)(&abc[0].x)[d] + (double)(&abc[1].x)[d] + G.js.x0[d];
    if (pvel)
        for (int d = 0; d < 3; d++) pvel[d] = (double)(&abc[2].x)[d] + (double)(&abc[3].x)[d];
    use(0);
    return 0;
}

// ===========================================================================
// Part 2: g6x_ extensions
// ===========================================================================
int g6x_version(void) { return 100; }

int g6x_set_stream(void *cuda_stream, int external)
{
    require_open("g6x_set_stream");
    CK(cudaStreamSynchronize(G.stream));
    // a NULL handle with external != 0 is the legacy default stream (torch's default stream)
    G.stream = external ? (cudaStream_t)cuda_stream : G.own_stream;
    G.ext_stream = external != 0;
    return 0;
}

int g6x_set_refine(int on)
{
    for (int k = 0; k < std::max(1, M.n); k++) g_ctx[k].refine = on ? 1 : 0;
    return 0;
}

int g6x_set_close_factor(double k_close, double far_factor)
{
    for (int k = 0; k < std::max(1, M.n); k++) {
        Context &c = g_ctx[k];
        if (k_close >= 0.0) c.kclose = (float)k_close;
        if (far_factor >= 0.0) c.farc = (float)far_factor;
        c.ord.kclose = (c.order_tiny && c.kclose > 0.f) ? 1.0e30f : c.kclose;
        c.ord.farc2 = c.farc * c.farc;
    }
    return 0;
}

long long g6x_order_rebuilds(void) { return g_ctx[0].order_rebuilds; }

// (warp x group) blocks the speculative kernel took FAR / NEAR / CLOSE and NEAR blocks it redid, since the last
// call (device 0).  Counted only by builds with -DG6_STATS (returns 0 there, -1 otherwise).
int g6x_block_stats(unsigned long long out[8])
{
    require_open("g6x_block_stats");
    for (int m = 0; m < 8; m++) out[m] = 0;
#ifdef G6_STATS
    use(0);
    CK(cudaStreamSynchronize(G.stream));
    if (!G.d_stats) {
        dev_alloc(G.d_stats, 8);
        CK(cudaMemset(G.d_stats, 0, 8 * sizeof(unsigned long long)));
        return 0;
    }
    CK(cudaMemcpy(out, G.d_stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CK(cudaMemset(G.d_stats, 0, 8 * sizeof(unsigned long long)));
    return 0;
#else
    return -1;
#endif
}
int g6x_device_count_open(void) { return M.n; }

int g6x_set_j_offset(int offset)
{
    G.j_offset = offset;
    return 0;
}

int g6x_set_j_particles(int n, const int *address, int address0, const int *index, const double *tj,
                        const double *mass, const double (*j6)[3], const double (*a2)[3], const double (*v)[3],
                        const double (*x)[3])
{
    require_open("g6x_set_j_particles");
    static const double zero3[3] = {0, 0, 0};
    if (n <= 0) return 0;
    if (M.n > 1) {   // every device holds all particles
        for (int d = 0; d < M.n; d++) {
            use(d);
            int maxa = address0 + n - 1;
            if (address)
                for (int k = 0; k < n; k++) maxa = std::max(maxa, address[k]);
            ensure_capacity(maxa + 1);
            ensure_up_cap(G.up_n + n);
            for (int k = 0; k < n; k++)
                stage_j(address ? address[k] : address0 + k, index[k], tj ? tj[k] : 0.0, mass[k], j6 ? j6[k] : zero3,
                        a2 ? a2[k] : zero3, v[k], x[k]);
        }
        use(0);
        return 0;
    }
    int maxaddr = address0 + n - 1;
    if (address)
        for (int k = 0; k < n; k++) maxaddr = std::max(maxaddr, address[k]);
    ensure_capacity(maxaddr + 1);
    ensure_up_cap(G.up_n + n);
    for (int k = 0; k < n; k++)
        stage_j(address ? address[k] : address0 + k, index[k], tj ? tj[k] : 0.0, mass[k], j6 ? j6[k] : zero3,
                a2 ? a2[k] : zero3, v[k], x[k]);
    return 0;
}

int g6x_predict(int nj, double ti)
{
    require_open("g6x_predict");
    const int nd = std::max(1, M.n);
    for (int k = 0; k < nd; k++) {
        use(k);
        G.ti = ti;
        flush_updates();
        if (order_stale(nj)) rebuild_order(nj);
        run_predictor(nj);
    }
    use(0);
    return 0;
}

// ---- peer exchange helpers (shared by g6x_calc_device_allreduce and the sharded Hermite step) ----
struct ExchangeSlots {
    unsigned char *half[MAX_PEERS + 1];   // slot of THIS rank in every rank's buffer, current half
};
static ExchangeSlots exchange_begin(int ni, const char *who)
{
    Context::Peer &P = G.peer;
    if (!P.attached || ni > P.cap) {
        fprintf(stderr, "g6_b200: FATAL %s: %s (ni %d, capacity %d)\n", who,
                P.attached ? "i-set exceeds the exchange capacity" : "no peers attached", ni, P.cap);
        exit(-1);
    }
    ExchangeSlots e{};
    if (ni <= 0) return e;   // nothing is exchanged (and no flag/combine kernel runs): the sequence must not advance
    P.seq++;
    for (int r = 0; r < P.world; r++) e.half[r] = P.peer_buf[r] + (P.seq & 1) * P.half_bytes + P.rank * P.slot_bytes;
    return e;
}
static double *slot_sum(unsigned char *b, int i0) { return reinterpret_cast<double *>(b) + 7 * (size_t)i0; }
static u64 *slot_key(unsigned char *b, int i0)
{
    return reinterpret_cast<u64 *>(b + sizeof(double) * 7 * (size_t)G.peer.cap) + i0;
}
static int *slot_id(unsigned char *b, int i0)
{
    return reinterpret_cast<int *>(b + sizeof(double) * 8 * (size_t)G.peer.cap) + i0;
}
static void exchange_set_mirrors(const ExchangeSlots &e, int i0)
{
    Context::Peer &P = G.peer;
    G.mir_n = 0;
    for (int r = 0; r < P.world; r++) {
        if (r == P.rank) continue;
        G.mir_sum[G.mir_n] = slot_sum(e.half[r], i0);
        G.mir_key[G.mir_n] = slot_key(e.half[r], i0);
        G.mir_id[G.mir_n] = slot_id(e.half[r], i0);
        G.mir_n++;
    }
}
static void exchange_finish(int ni, double *d_sum, unsigned long long *d_key, int *d_nnid)
{
    Context::Peer &P = G.peer;
    PeerSlots ps{};
    ps.world = P.world;
    ps.rank = P.rank;
    unsigned char *mine = P.buf + (P.seq & 1) * P.half_bytes;
    for (int r = 0; r < P.world; r++) {
        unsigned char *b = mine + r * P.slot_bytes;
        ps.sum[r] = reinterpret_cast<const double *>(b);
        ps.key[r] = reinterpret_cast<const u64 *>(b + sizeof(double) * 7 * (size_t)P.cap);
        ps.id[r] = reinterpret_cast<const int *>(b + sizeof(double) * 8 * (size_t)P.cap);
    }
    const size_t foff = P.flags_off + (P.seq & 1) * sizeof(unsigned long long) * P.world;
    ps.flag = reinterpret_cast<volatile unsigned long long *>(P.buf + foff);
    ps.n_remote = 0;
    for (int r = 0; r < P.world; r++)
        if (r != P.rank)
            ps.remote_flag[ps.n_remote++] = reinterpret_cast<unsigned long long *>(P.peer_buf[r] + foff) + P.rank;
    peer_flag_kernel<<<1, 32, 0, G.stream>>>(ps, P.seq);
    CK(cudaGetLastError());
    const int ctas = std::max(1, std::min(2 * G.sm_count, (ni + 255) / 256));
    peer_combine_kernel<<<ctas, 256, 0, G.stream>>>(ps, P.seq, ni, d_sum, reinterpret_cast<u64 *>(d_key), d_nnid,
                                                    P.dev_h_err, nullptr, 0ull);
    CK(cudaGetLastError());
    G.launches += 2;
}

// ---- in-process device group (multi-device ABI): peer access enabled directly, exchange buffers shared by
// pointer, partials gathered at device 0 ------------------------------------------------------------
static void peer_group_setup()
{
    const int n = M.n;
    for (int k = 0; k < n; k++) {
        use(k);
        for (int r = 0; r < n; r++) {
            if (r == k) continue;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, g_ctx[k].device, g_ctx[r].device));
            if (!can) {
                fprintf(stderr, "g6_b200: FATAL G6_B200_DEVICES=%d: device %d cannot access device %d's memory\n", n,
                        g_ctx[k].device, g_ctx[r].device);
                exit(-1);
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(g_ctx[r].device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
            else CK(e);
        }
        Context::Peer &P = G.peer;
        P.world = n;
        P.rank = k;
        P.cap = (G.npipes + 15) / 16 * 16;
        P.slot_bytes = (size_t)P.cap * (sizeof(double) * 7 + sizeof(u64) + sizeof(int));
        P.half_bytes = P.slot_bytes * n;
        P.flags_off = 2 * P.half_bytes;
        P.buf_bytes = P.flags_off + 2 * sizeof(unsigned long long) * n;
        CK(cudaMalloc((void **)&P.buf, P.buf_bytes));
        CK(cudaMemset(P.buf, 0, P.buf_bytes));
        host_alloc(P.h_err, 1);
        P.h_err[0] = 0;
        P.dev_h_err = dev_alias(P.h_err);
        P.seq = 0;
        P.allocated = true;
        P.ipc = false;
    }
    for (int k = 0; k < n; k++) {
        Context::Peer &P = g_ctx[k].peer;
        for (int r = 0; r < n; r++) P.peer_buf[r] = g_ctx[r].peer.buf;
        P.attached = true;
    }
    use(0);
}

// all devices: next exchange sequence number
static void gather_begin_all(int ni)
{
    for (int k = 0; k < M.n; k++) {
        Context::Peer &P = g_ctx[k].peer;
        if (ni > P.cap) {
            fprintf(stderr, "g6_b200: FATAL i-block of %d exceeds the exchange capacity %d\n", ni, P.cap);
            exit(-1);
        }
        P.seq++;
    }
}
// device k's force launch: its outputs go to slot k of the ROOT's exchange buffer (a local store for the root,
// stores over NVLink for the others)
static void gather_launch(int k, int nj, int ni, const IBlock &ib, float eps2, double *, u64 *, int *,
                          const float4 *inline_src, unsigned long long)
{
    Context::Peer &P = G.peer;   // == g_ctx[k].peer (the caller selected device k)
    unsigned char *slot = P.peer_buf[0] + (P.seq & 1) * P.half_bytes + (size_t)k * P.slot_bytes;
    double *ssum = reinterpret_cast<double *>(slot);
    u64 *skey = reinterpret_cast<u64 *>(slot + sizeof(double) * 7 * (size_t)P.cap);
    int *sid = reinterpret_cast<int *>(slot + sizeof(double) * 8 * (size_t)P.cap);
    G.mir_n = 0;
    launch_force(nj, ni, ib, eps2, true, false, ssum, skey, sid, inline_src);
    // tell the root this device's slot is complete
    PeerSlots ps{};
    ps.world = P.world;
    ps.rank = k;
    const size_t foff = P.flags_off + (P.seq & 1) * sizeof(unsigned long long) * P.world;
    ps.flag = reinterpret_cast<volatile unsigned long long *>(P.peer_buf[0] + foff);   // the root's flags
    ps.n_remote = 0;
    peer_flag_kernel<<<1, 32, 0, G.stream>>>(ps, P.seq);
    CK(cudaGetLastError());
    G.launches++;
}
// root: wait for all flags, combine, write the totals to d_sum (or straight to mapped host memory + flag)
static void gather_finish_all(int ni, bool direct)
{
    use(0);
    Context::Peer &P = G.peer;
    PeerSlots ps{};
    ps.world = P.world;
    ps.rank = 0;
    unsigned char *mine = P.buf + (P.seq & 1) * P.half_bytes;
    for (int r = 0; r < P.world; r++) {
        unsigned char *b = mine + r * P.slot_bytes;
        ps.sum[r] = reinterpret_cast<const double *>(b);
        ps.key[r] = reinterpret_cast<const u64 *>(b + sizeof(double) * 7 * (size_t)P.cap);
        ps.id[r] = reinterpret_cast<const int *>(b + sizeof(double) * 8 * (size_t)P.cap);
    }
    const size_t foff = P.flags_off + (P.seq & 1) * sizeof(unsigned long long) * P.world;
    ps.flag = reinterpret_cast<volatile unsigned long long *>(P.buf + foff);
    ps.n_remote = 0;
    double *out = direct ? G.dev_h_sum : G.d_sum;
    const int ctas = direct ? 1 : std::max(1, std::min(2 * G.sm_count, (ni + 255) / 256));
    peer_combine_kernel<<<ctas, 256, 0, G.stream>>>(ps, P.seq, ni, out, reinterpret_cast<u64 *>(G.d_key),
                                                    reinterpret_cast<int *>(out + 7 * (size_t)ni), P.dev_h_err,
                                                    direct ? G.dev_h_flag : nullptr, G.flag_seq);
    CK(cudaGetLastError());
    G.launches++;
}

// Shared body of the device-resident entry points.  With peers attached and `exchange` set, every launch
// writes this rank's partials into its slot of all exchange buffers (own + peers, over NVLink) and the
// call ends with the flag + combine kernels, so d_sum/d_key/d_nnid hold the totals over all j-shards.
static int calc_device_impl(int nj, int ni, const int *d_index, const double *d_xi, const double *d_vi,
                            const double *d_h2, double eps2, int flags, double *d_sum, unsigned long long *d_key,
                            int *d_nnid, bool exchange)
{
    flush_updates();
    if (order_stale(nj)) rebuild_order(nj);
    run_predictor(nj);
    bool nn = (flags & 1) != 0, list = (flags & 2) != 0;
    if (list) {
        fprintf(stderr, "g6_b200: FATAL g6x_calc_device: neighbour lists are served by the g6 ABI path only\n");
        exit(-1);
    }
    Context::Peer &P = G.peer;
    ExchangeSlots ex{};
    if (exchange) ex = exchange_begin(ni, "g6x_calc_device_allreduce");
    // g6x_set_j_window: this process sums over a window of the slots only (every rank of a multi-process run holds
    // all particles, like ph4's MPI ranks, and owns a contiguous piece of the Morton-ordered j-memory)
    int wlo = 0, whi = std::min(nj, G.capacity);
    if (G.cd_win_hi > 0) {
        wlo = std::min(G.cd_win_lo, whi);
        whi = std::min(G.cd_win_hi, whi);
    }
    const int njc = std::max(0, whi - wlo);
    const int chunk = device_chunk(ni, njc);
    if (chunk > G.i2_cap) {
        CK(cudaStreamSynchronize(G.stream));
        dev_free(G.d_i2);
        dev_free(G.d_conf2);
        dev_alloc(G.d_i2, (size_t)4 * chunk);
        dev_alloc(G.d_conf2, (size_t)chunk);
        G.i2_cap = chunk;
    }
    // the speculative kernel wants Morton-sorted i-particles (64 neighbours per warp): sort the whole set once,
    // pack every chunk through the permutation, and let the kernels write their outputs to the caller's index
    const bool sorted = (ni > 1) && G.ord.nkeys > 0 && is_fast_variant(choose_variant(std::min(ni, chunk), njc));
    if (sorted) {
        if (ni > G.isort_cap) {
            CK(cudaStreamSynchronize(G.stream));
            dev_free(G.d_ikey); dev_free(G.d_ikey_tmp); dev_free(G.d_iperm); dev_free(G.d_iperm_tmp);
            dev_alloc(G.d_ikey, (size_t)ni); dev_alloc(G.d_ikey_tmp, (size_t)ni);
            dev_alloc(G.d_iperm, (size_t)ni); dev_alloc(G.d_iperm_tmp, (size_t)ni);
            G.isort_cap = ni;
        }
        i_key_kernel<<<(ni + 255) / 256, 256, 0, G.stream>>>(ni, d_xi, G.js, G.ord, G.d_ikey_tmp, G.d_iperm_tmp);
        CK(cudaGetLastError());
        size_t need = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, need, G.d_ikey_tmp, G.d_ikey, G.d_iperm_tmp, G.d_iperm, ni, 0, 30,
                                           G.stream));
        if (need > G.sort_tmp_bytes) {
            CK(cudaStreamSynchronize(G.stream));
            if (G.d_sort_tmp) cudaFree(G.d_sort_tmp);
            CK(cudaMalloc(&G.d_sort_tmp, need));
            G.sort_tmp_bytes = need;
        }
        CK(cub::DeviceRadixSort::SortPairs(G.d_sort_tmp, need, G.d_ikey_tmp, G.d_ikey, G.d_iperm_tmp, G.d_iperm, ni, 0, 30,
                                           G.stream));
        G.launches += 2;
    }
    for (int i0 = 0; i0 < ni; i0 += chunk) {
        int n = std::min(chunk, ni - i0);
        const IBlock ib = iblock_of(G.d_i2, G.i2_cap, G.d_conf2, sorted ? G.d_iperm + i0 : nullptr);
        // sorted: packed particle k of this chunk is caller particle d_iperm[i0 + k] and its outputs go there;
        // otherwise the chunk is the caller's [i0, i0 + n) and so are its outputs
        const int ob = sorted ? 0 : i0;
        pack_i_kernel<<<(n + 255) / 256, 256, 0, G.stream>>>(
            n, ib.iperm, d_index + ob, d_xi + 3 * (size_t)ob, d_vi + 3 * (size_t)ob, d_h2 ? d_h2 + ob : nullptr,
            G.js.x0[0], G.js.x0[1], G.js.x0[2], const_cast<float4 *>(ib.A), const_cast<float4 *>(ib.B),
            const_cast<float4 *>(ib.C), const_cast<float4 *>(ib.D));
        G.launches++;
        CK(cudaGetLastError());
        G.win_lo = wlo;
        if (!exchange) {
            launch_force(njc, n, ib, (float)eps2, nn, false, d_sum + 7 * (size_t)ob, d_key + ob, d_nnid + ob);
            G.win_lo = 0;
            continue;
        }
        exchange_set_mirrors(ex, ob);
        unsigned char *own = ex.half[P.rank];
        launch_force(njc, n, ib, (float)eps2, nn, false, slot_sum(own, ob), slot_key(own, ob), slot_id(own, ob));
        G.win_lo = 0;
        G.mir_n = 0;
    }
    if (exchange && ni > 0) exchange_finish(ni, d_sum, d_key, d_nnid);
    return 0;
}

int g6x_calc_device(int nj, int ni, const int *d_index, const double *d_xi, const double *d_vi, const double *d_h2,
                    double eps2, int flags, double *d_sum, unsigned long long *d_key, int *d_nnid)
{
    require_open("g6x_calc_device");
    return calc_device_impl(nj, ni, d_index, d_xi, d_vi, d_h2, eps2, flags, d_sum, d_key, d_nnid, false);
}

int g6x_calc_device_allreduce(int nj, int ni, const int *d_index, const double *d_xi, const double *d_vi,
                              const double *d_h2, double eps2, int flags, double *d_sum, unsigned long long *d_key,
                              int *d_nnid)
{
    require_open("g6x_calc_device_allreduce");
    return calc_device_impl(nj, ni, d_index, d_xi, d_vi, d_h2, eps2, flags, d_sum, d_key, d_nnid, true);
}

int g6x_peer_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int g6x_peer_alloc(int world, int rank, int capacity, void *handle_out)
{
    require_open("g6x_peer_alloc");
    Context::Peer &P = G.peer;
    if (world < 1 || world > MAX_PEERS + 1 || rank < 0 || rank >= world || capacity <= 0) return -1;
    if (P.allocated) g6x_peer_detach();
    P.world = world;
    P.rank = rank;
    P.cap = (capacity + 15) / 16 * 16;
    P.slot_bytes = (size_t)P.cap * (sizeof(double) * 7 + sizeof(u64) + sizeof(int));
    P.half_bytes = P.slot_bytes * world;
    P.flags_off = 2 * P.half_bytes;
    P.buf_bytes = P.flags_off + 2 * sizeof(unsigned long long) * world;
    CK(cudaMalloc((void **)&P.buf, P.buf_bytes));
    CK(cudaMemset(P.buf, 0, P.buf_bytes));
    host_alloc(P.h_err, 1);
    P.h_err[0] = 0;
    P.dev_h_err = dev_alias(P.h_err);
    P.seq = 0;
    P.allocated = true;
    P.attached = false;
    P.ipc = true;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, P.buf));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

int g6x_peer_attach(const void *handles)
{
    require_open("g6x_peer_attach");
    Context::Peer &P = G.peer;
    if (!P.allocated) return -1;
    const cudaIpcMemHandle_t *h = static_cast<const cudaIpcMemHandle_t *>(handles);
    for (int r = 0; r < P.world; r++) {
        if (r == P.rank) {
            P.peer_buf[r] = P.buf;
            continue;
        }
        void *q = nullptr;
        CK(cudaIpcOpenMemHandle(&q, h[r], cudaIpcMemLazyEnablePeerAccess));
        P.peer_buf[r] = static_cast<unsigned char *>(q);
    }
    P.attached = true;
    return 0;
}

int g6x_peer_detach(void)
{
    Context::Peer &P = G.peer;
    if (!P.allocated) return 0;
    if (G.open) CK(cudaStreamSynchronize(G.stream));
    for (int r = 0; r < P.world; r++)
        if (P.attached && P.ipc && r != P.rank && P.peer_buf[r]) cudaIpcCloseMemHandle(P.peer_buf[r]);
    if (P.buf) cudaFree(P.buf);
    host_free(P.h_err);
    P = Context::Peer{};
    return 0;
}

int g6x_peer_error(void)
{
    return (G.peer.allocated && G.peer.h_err) ? (int)G.peer.h_err[0] : 0;
}

int g6x_device_chunk(int ni)
{
    require_open("g6x_device_chunk");
    int nj = std::min(G.nj_hi, G.capacity);
    if (G.cd_win_hi > 0) nj = std::max(0, std::min(G.cd_win_hi, nj) - std::min(G.cd_win_lo, nj));
    return device_chunk(ni, nj);
}

int g6x_set_j_window(int slot_lo, int slot_hi)
{
    require_open("g6x_set_j_window");
    if (slot_hi > 0 && (slot_lo < 0 || slot_lo % TILE != 0 || slot_hi < slot_lo)) {
        fprintf(stderr, "g6_b200: g6x_set_j_window: slot_lo must be a multiple of %d and <= slot_hi\n", TILE);
        return -1;
    }
    G.cd_win_lo = slot_hi > 0 ? slot_lo : 0;
    G.cd_win_hi = slot_hi > 0 ? slot_hi : 0;
    return 0;
}

int g6x_resolve_nn(int ni, const unsigned long long *d_key, int rank, int *d_nnid)
{
    require_open("g6x_resolve_nn");
    if (ni <= 0) return 0;
    resolve_nn_kernel<<<(ni + 255) / 256, 256, 0, G.stream>>>(ni, d_key, rank, G.j_offset,
                                                               std::min(G.nj_hi, G.capacity), G.js.slot_of, G.js.B, d_nnid,
                                                               G.cd_win_lo, G.cd_win_hi > 0 ? G.cd_win_hi : 0x7fffffff);
    G.launches++;
    CK(cudaGetLastError());
    return 0;
}

// ---- device-resident Hermite block step -----------------------------------------------------------
static void hermite_reserve(int n)
{
    Context::Hermite &H = G.herm;
    if (n <= H.cap) return;
    CK(cudaStreamSynchronize(G.stream));
    int cap = std::max(n, std::max(4096, 2 * H.cap));
    host_free(H.h_ilist); host_free(H.h_olddt); host_free(H.h_outdt); host_free(H.h_outpot); host_free(H.h_outnn);
    dev_free(H.d_pred); dev_free(H.d_i); dev_free(H.d_sum); dev_free(H.d_key); dev_free(H.d_nnid);
    dev_free(H.d_ilist); dev_free(H.d_olddt); dev_free(H.d_conf);
    dev_alloc(H.d_conf, (size_t)cap);
    dev_alloc(H.d_ilist, (size_t)cap);
    dev_alloc(H.d_olddt, (size_t)cap);
    host_alloc(H.h_ilist, cap);   H.dev_h_ilist = dev_alias(H.h_ilist);
    host_alloc(H.h_olddt, cap);   H.dev_h_olddt = dev_alias(H.h_olddt);
    host_alloc(H.h_outdt, cap);   H.dev_h_outdt = dev_alias(H.h_outdt);
    host_alloc(H.h_outpot, cap);  H.dev_h_outpot = dev_alias(H.h_outpot);
    host_alloc(H.h_outnn, cap);   H.dev_h_outnn = dev_alias(H.h_outnn);
    dev_alloc(H.d_pred, (size_t)6 * cap);
    dev_alloc(H.d_i, (size_t)4 * cap);
    dev_alloc(H.d_sum, (size_t)7 * cap);
    dev_alloc(H.d_key, (size_t)cap);
    dev_alloc(H.d_nnid, (size_t)cap);
    H.cap = cap;
}

// One pass over n <= cap active particles whose addresses / old steps are already in H.h_ilist / H.h_olddt.
// Blocks until the results are in H.h_outdt / h_outpot / h_outnn.
static void hermite_pass(int nj, int n, double tnext, double eta, double eps2, int mode)
{
    Context::Hermite &H = G.herm;
    HermiteArgs h{};
    h.ni = n;
    h.ilist = H.dev_h_ilist;
    h.old_dt = H.dev_h_olddt;
    h.ilist_d = H.d_ilist;
    h.olddt_d = H.d_olddt;
    h.tnext = tnext;
    h.eta = eta;
    h.js = G.js;
    h.slot_of = G.js.slot_of;
    h.iA = H.d_i; h.iB = H.d_i + n; h.iC = H.d_i + 2 * (size_t)n; h.iD = H.d_i + 3 * (size_t)n;
    const IBlock ib = iblock_of(H.d_i, n, H.d_conf);
    h.pred = H.d_pred;
    h.sum = H.d_sum;
    h.nnid = H.d_nnid;
    h.out_dt = H.dev_h_outdt; h.out_pot = H.dev_h_outpot; h.out_nn = H.dev_h_outnn;
    h.mode = mode;
    h.done_counter = G.d_done;
    h.host_flag = G.dev_h_flag;
    h.flag_seq = ++G.flag_seq;
    const int ctas = (n + 255) / 256;
    G.ti = tnext;
    if (H.shard_hi > 0) {
        // replicated state, sharded forces: gather the block (every rank, identical), predict (every rank, all j:
        // the neighbour bounds of the speculative kernel look at the whole j-memory), sum over this rank's window
        // of the SLOTS with the partials mirrored into the peers' exchange buffers, combine, and let every rank
        // correct its own replica with the identical totals.  The Morton permutation is a function of the state,
        // which is identical on all ranks, so a window of slots holds the same particles everywhere.
        int lo = H.shard_lo, hi = std::min(H.shard_hi, std::min(nj, G.capacity));
        if (hi <= lo) lo = hi = 0;   // empty window (more ranks than j-tiles): this rank contributes zeros
        hermite_gather_kernel<<<ctas, 256, 0, G.stream>>>(h);
        CK(cudaGetLastError());
        G.j_dirty = true;
        run_predictor(nj);
        ExchangeSlots ex = exchange_begin(n, "g6x_hermite_step (sharded)");
        exchange_set_mirrors(ex, 0);
        unsigned char *own = ex.half[G.peer.rank];
        G.win_lo = lo;
        launch_force(std::max(0, hi - lo), n, ib, (float)eps2, true, false, slot_sum(own, 0), slot_key(own, 0),
                     slot_id(own, 0));
        G.win_lo = 0;
        G.mir_n = 0;
        exchange_finish(n, H.d_sum, reinterpret_cast<unsigned long long *>(H.d_key), H.d_nnid);
        hermite_correct_kernel<<<ctas, 256, 0, G.stream>>>(h);
        CK(cudaGetLastError());
        G.launches += 3;
    } else if (n <= 384 && G.variant == V_AUTO) {
        // small block: two launches -- (predict all j + gather/predict the block), then the force kernel
        // whose final-output stage is the corrector
        const int njc = std::min(nj, G.capacity);
        const int npred = std::max(njc, std::min(G.nj_hi, G.capacity));
        const int ntiles = (npred + TILE - 1) / TILE;
        hermite_predict_gather_kernel<<<ntiles + ctas, TILE, 0, G.stream>>>(ntiles, npred, tnext, h);
        CK(cudaGetLastError());
        G.predicted_nj = npred;
        G.predicted_ti = tnext;
        G.j_dirty = false;
        launch_force(nj, n, ib, (float)eps2, true, false, H.d_sum, H.d_key, H.d_nnid, nullptr, h.flag_seq, &h);
        G.launches += 1;
    } else {
        hermite_gather_kernel<<<ctas, 256, 0, G.stream>>>(h);
        CK(cudaGetLastError());
        run_predictor(nj);
        launch_force(nj, n, ib, (float)eps2, true, false, H.d_sum, H.d_key, H.d_nnid);
        hermite_correct_kernel<<<ctas, 256, 0, G.stream>>>(h);
        CK(cudaGetLastError());
        G.launches += 2;
    }
    G.j_dirty = true;   // the active particles' state changed: predict again before the next force
    wait_flag(h.flag_seq, "Hermite step");
}

int g6x_hermite_step(int nj, int ni, const int *ilist, double tnext, double eta, double eps2, const double *old_dt,
                     double *new_dt, double *pot, int *nn)
{
    require_open("g6x_hermite_step");
    Context::Hermite &H = G.herm;
    if (G.pending) CK(cudaStreamSynchronize(G.stream));
    flush_updates();
    if (ni <= 0) return 0;
    G.updates_since_order += ni;   // the corrector moves particles without passing through stage_j
    if (order_stale(nj)) rebuild_order(nj);
    // the whole block in one pass, whatever its size: every force is computed against the PREDICTED state
    // of all j before any particle is corrected, as in idata::advance
    hermite_reserve(ni);
    // big blocks go to the speculative kernel, which wants Morton neighbours side by side: the slots ARE in
    // Morton order, so the block is walked in slot order
    const bool sorted = ni > 384;
    static std::vector<int> order;
    if (sorted) {
        order.resize(ni);
        for (int k = 0; k < ni; k++) order[k] = k;
        const int *so = G.h_slot_of.data();
        std::sort(order.begin(), order.end(), [&](int a, int b) { return so[ilist[a]] < so[ilist[b]]; });
        for (int k = 0; k < ni; k++) {
            H.h_ilist[k] = ilist[order[k]];
            H.h_olddt[k] = old_dt[order[k]];
        }
    } else {
        memcpy(H.h_ilist, ilist, sizeof(int) * ni);
        memcpy(H.h_olddt, old_dt, sizeof(double) * ni);
    }
    hermite_pass(nj, ni, tnext, eta, eps2, 0);
    for (int k = 0; k < ni; k++) {
        const int i = sorted ? order[k] : k;
        new_dt[i] = H.h_outdt[k];
        if (pot) pot[i] = H.h_outpot[k];
        if (nn) nn[i] = H.h_outnn[k];
    }
    return 0;
}

int g6x_hermite_set_shard(int j_lo, int j_hi)
{
    require_open("g6x_hermite_set_shard");
    if (j_hi > 0 && (j_lo < 0 || j_lo % TILE != 0 || j_hi < j_lo)) {
        fprintf(stderr, "g6_b200: g6x_hermite_set_shard: j_lo must be a multiple of %d and <= j_hi\n", TILE);
        return -1;
    }
    G.herm.shard_lo = j_hi > 0 ? j_lo : 0;
    G.herm.shard_hi = j_hi > 0 ? j_hi : 0;
    return 0;
}

int g6x_hermite_init(int nj, double t0, double eta, double eps2, double *timestep_out)
{
    require_open("g6x_hermite_init");
    Context::Hermite &H = G.herm;
    if (G.pending) CK(cudaStreamSynchronize(G.stream));
    flush_updates();
    nj = std::min(nj, G.capacity);
    if (nj <= 0) return -1;
    if (order_stale(nj)) rebuild_order(nj);
    hermite_reserve(nj);
    H.time.assign(nj, t0);
    H.dt.assign(nj, 0.0);
    {   // all particles, walked in slot (= Morton) order
        std::vector<int> addr_of_slot(nj, 0);
        for (int a = 0; a < nj; a++) addr_of_slot[G.h_slot_of[a]] = a;
        for (int k = 0; k < nj; k++) H.h_ilist[k] = addr_of_slot[k];
    }
    hermite_pass(nj, nj, t0, eta, eps2, 1);   // forces of all particles at t0, first steps (jdata.cc:503-548)
    for (int k = 0; k < nj; k++) H.dt[H.h_ilist[k]] = H.h_outdt[k];
    if (timestep_out) memcpy(timestep_out, H.dt.data(), sizeof(double) * nj);
    H.system_time = t0;
    H.block_steps = H.particle_steps = 0;
    H.initialised = true;
    return 0;
}

// ph4's jdata::advance loop (jdata.cc:752-795) with the scheduler on the host (a heap of next times;
// the reference keeps a sorted list, scheduler.cc) and everything else on the device.
// stats[0] = system time reached, [1] = block steps, [2] = particle steps, [3] = wall seconds.
long long g6x_hermite_evolve(int nj, double t_end, double eta, double eps2, long long max_block_steps, double *stats)
{
    require_open("g6x_hermite_evolve");
    Context::Hermite &H = G.herm;
    if (!H.initialised || (int)H.time.size() != std::min(nj, G.capacity)) {
        fprintf(stderr, "g6_b200: FATAL g6x_hermite_evolve before g6x_hermite_init\n");
        exit(-1);
    }
    nj = (int)H.time.size();
    typedef std::pair<double, int> Ev;
    std::vector<Ev> heap;
    heap.reserve(nj);
    for (int j = 0; j < nj; j++) heap.push_back(Ev(H.time[j] + H.dt[j], j));
    auto later = [](const Ev &a, const Ev &b) { return a.first > b.first || (a.first == b.first && a.second > b.second); };
    std::make_heap(heap.begin(), heap.end(), later);
    std::vector<int> ilist;
    std::vector<double> olddt, newdt;
    const double w0 = wall();
    long long steps = 0;
    while (H.system_time < t_end && (max_block_steps <= 0 || steps < max_block_steps)) {
        const double tnext = heap.front().first;
        ilist.clear();
        while (!heap.empty() && heap.front().first == tnext) {
            std::pop_heap(heap.begin(), heap.end(), later);
            ilist.push_back(heap.back().second);
            heap.pop_back();
        }
        const int ni = (int)ilist.size();
        olddt.resize(ni);
        newdt.resize(ni);
        for (int k = 0; k < ni; k++) olddt[k] = H.dt[ilist[k]];
        g6x_hermite_step(nj, ni, ilist.data(), tnext, eta, eps2, olddt.data(), newdt.data(), nullptr, nullptr);
        for (int k = 0; k < ni; k++) {
            const int j = ilist[k];
            if (!(newdt[k] > 0.0) || !std::isfinite(newdt[k])) {
                fprintf(stderr, "g6_b200: FATAL g6x_hermite_evolve: particle %d got time step %g at t = %.17g\n", j,
                        newdt[k], tnext);
                exit(-1);
            }
            H.time[j] = tnext;
            H.dt[j] = newdt[k];
            heap.push_back(Ev(tnext + newdt[k], j));
            std::push_heap(heap.begin(), heap.end(), later);
        }
        H.system_time = tnext;
        H.block_steps++;
        H.particle_steps += ni;
        steps++;
    }
    if (stats) {
        stats[0] = H.system_time;
        stats[1] = (double)H.block_steps;
        stats[2] = (double)H.particle_steps;
        stats[3] = wall() - w0;
    }
    return steps;
}

int g6x_hermite_get_state(int nj, double *t, double (*x)[3], double (*v)[3], double (*a)[3], double (*j)[3])
{
    require_open("g6x_hermite_get_state");
    flush_updates();
    nj = std::min(nj, G.capacity);
    std::vector<double2> q[7];
    CK(cudaStreamSynchronize(G.stream));
    for (int k = 0; k < 7; k++) {
        q[k].resize(nj);
        CK(cudaMemcpy(q[k].data(), G.js.q[k], sizeof(double2) * nj, cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i < nj; i++) {   // caller's address i lives in slot s (addresses [0, nj) occupy slots [0, nj))
        const int s = G.h_slot_of[i] < nj ? G.h_slot_of[i] : i;
        if (x) { x[i][0] = q[0][s].x; x[i][1] = q[0][s].y; x[i][2] = q[1][s].x; }
        if (t) t[i] = q[1][s].y;
        if (v) { v[i][0] = q[2][s].x; v[i][1] = q[2][s].y; v[i][2] = q[3][s].x; }
        if (a) { a[i][0] = q[3][s].y; a[i][1] = q[4][s].x; a[i][2] = q[4][s].y; }
        if (j) { j[i][0] = q[5][s].x; j[i][1] = q[5][s].y; j[i][2] = q[6][s].x; }
    }
    return 0;
}

int g6x_synchronize(void)
{
    require_open("g6x_synchronize");
    CK(cudaStreamSynchronize(G.stream));
    return 0;
}

long long g6x_launch_count(void) { return G.launches; }

int g6x_get_predicted(void **A, void **B, void **C, int *capacity)
{
    require_open("g6x_get_predicted");
    *A = G.js.A; *B = G.js.B; *C = G.js.C;
    *capacity = G.capacity;
    return 0;
}

int g6x_read_predicted(int nj, double (*pos)[3], double (*vel)[3])
{
    require_open("g6x_read_predicted");
    nj = std::min(nj, G.capacity);
    const int n = std::min(G.capacity, std::max(nj, G.nj_hi));   // slots that may hold addresses < nj
    std::vector<float4> A(n), B(n), C(n), L(n);
    CK(cudaStreamSynchronize(G.stream));
    CK(cudaMemcpy(A.data(), G.js.A, sizeof(float4) * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(B.data(), G.js.B, sizeof(float4) * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(C.data(), G.js.C, sizeof(float4) * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(L.data(), G.js.L, sizeof(float4) * n, cudaMemcpyDeviceToHost));
    for (int j = 0; j < nj; j++) {   // by the caller's address; positions back in the caller's frame
        const int s = G.h_slot_of[j];
        pos[j][0] = (double)A[s].x + (double)B[s].x + G.js.x0[0];
        pos[j][1] = (double)A[s].y + (double)B[s].y + G.js.x0[1];
        pos[j][2] = (double)A[s].z + (double)B[s].z + G.js.x0[2];
        vel[j][0] = (double)C[s].x + (double)L[s].x;
        vel[j][1] = (double)C[s].y + (double)L[s].y;
        vel[j][2] = (double)C[s].z + (double)L[s].z;
    }
    return 0;
}

double g6x_time_predictor(int nj, int reps)
{
    require_open("g6x_time_predictor");
    flush_updates();
    nj = std::min(nj, G.capacity);
    if (nj <= 0 || reps <= 0) return 0.0;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int r = 0; r < 3; r++) predict_kernel<<<(nj + 255) / 256, 256, 0, G.stream>>>(nj, G.ti, G.js);
    CK(cudaEventRecord(e0, G.stream));
    for (int r = 0; r < reps; r++) predict_kernel<<<(nj + 255) / 256, 256, 0, G.stream>>>(nj, G.ti, G.js);
    CK(cudaEventRecord(e1, G.stream));
    CK(cudaEventSynchronize(e1));
    G.launches += reps + 3;
    G.predicted_nj = -1;  // slots >= nj of the last tile were parked: predict again before the next force
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return (double)ms / reps;
}

double g6x_latency_probe(int kernels, int reps)
{
    require_open("g6x_latency_probe");
    CK(cudaStreamSynchronize(G.stream));
    volatile unsigned long long *flag = G.h_flag;
    double t0 = 0;
    for (int r = 0; r < reps + 20; r++) {
        if (r == 20) t0 = wall();
        const unsigned long long seq = ++G.flag_seq;
        for (int k = 0; k < kernels - 1; k++) latency_probe_kernel<<<1, 32, 0, G.stream>>>(nullptr, 0, G.d_done + 0);
        latency_probe_kernel<<<1, 32, 0, G.stream>>>(G.dev_h_flag, seq, nullptr);
        while (*flag != seq) {
        }
    }
    return 1e6 * (wall() - t0) / reps;
}

int g6x_set_variant(int variant)
{
    if (variant < 0 || variant >= V_COUNT) return -1;
    for (int k = 0; k < std::max(1, M.n); k++) g_ctx[k].variant = variant;
    return 0;
}

double g6x_fp32_peak(int mode)
{
    require_open("g6x_fp32_peak");
    float *d = nullptr;
    dev_alloc(d, 256);
    const int iters = 4096;
    int blocks = G.sm_count * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 6; r++) {
        CK(cudaEventRecord(e0, G.stream));
        switch (mode) {
            case 0: fp32_peak_kernel<0><<<blocks, 256, 0, G.stream>>>(d, iters, 1.0f); break;
            case 1: fp32_peak_kernel<1><<<blocks, 256, 0, G.stream>>>(d, iters, 1.0f); break;
            case 2: fp32_peak_kernel<2><<<blocks, 256, 0, G.stream>>>(d, iters, 1.0f); break;
            case 3: fp32_peak_kernel<3><<<blocks, 256, 0, G.stream>>>(d, iters, 1.0f); break;
            case 4: fp32_peak_kernel<4><<<blocks, 256, 0, G.stream>>>(d, iters, 1.0f); break;
            case 5: fp32_peak_kernel<5><<<blocks, 256, 0, G.stream>>>(d, iters, 1.0f); break;
            default: fp32_peak_kernel<6><<<blocks, 256, 0, G.stream>>>(d, iters, 1.0f); break;
        }
        CK(cudaEventRecord(e1, G.stream));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0) best = std::min(best, ms);
        G.launches++;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    dev_free(d);
    double fmas = (double)blocks * 256 * iters * 8 * 8 * ((mode == 0 || mode == 2) ? 1 : 2);
    return 2.0 * fmas / (best * 1e-3) / 1e12;
}

}  // extern "C"
