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

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
    double ti = 0.0;
    double predicted_ti = 0.0;
    int predicted_nj = -1;  // prefix predicted at predicted_ti (-1: none)
    bool j_dirty = false;

    // staging of j-updates
    std::vector<int> slot_of_addr;  // address -> slot in the pending batch, -1
    JUpdate *h_up = nullptr;        // pinned batch being filled (= h_up2[up_cur])
    JUpdate *h_up2[2] = {nullptr, nullptr};   // two pinned batches: one fills while the other uploads
    cudaEvent_t up_done[2] = {nullptr, nullptr};
    int up_cur = 0;
    JUpdate *d_up = nullptr;
    int up_cap = 0, up_n = 0;

    // i-block buffers (npipes)
    float4 *h_i = nullptr;  // pinned [3][npipes]
    float4 *d_i = nullptr;  // [3][npipes]
    double *d_sum = nullptr;  // [7n doubles][n nearest-neighbour ids] of the current i-block
    u64 *d_key = nullptr;
    double *h_sum = nullptr;  // pinned, same layout
    // device-resident entry point scratch
    float4 *d_i2 = nullptr;
    int i2_cap = 0;
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
        bool allocated = false, attached = false;
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
        float4 *d_i = nullptr;                                 // [3][cap]
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
    bool pending = false;
};

Context G;

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
    CK(cudaHostAlloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T), cudaHostAllocMapped));
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
    grow_dev(G.js.A, old, newcap, G.stream);
    grow_dev(G.js.B, old, newcap, G.stream);
    grow_dev(G.js.C, old, newcap, G.stream);
    G.capacity = (int)newcap;
    G.slot_of_addr.resize(newcap, -1);
    G.predicted_nj = -1;
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
    while (*flag != want) {
        if ((++spins & 0xfffff) == 0) {
            if (t_start == 0.0) t_start = wall();
            if (wall() - t_start > 300.0) {   // nothing on this path runs for minutes: do not hang the caller forever
                fprintf(stderr, "g6_b200: FATAL %s did not complete within 300 s\n", what);
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
    for (int k = 0; k < G.up_n; k++) G.slot_of_addr[G.h_up[k].addr] = -1;
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
    for (int k = 0; k < G.up_n; k++) iu.addr[k] = G.h_up[k].addr;
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
        static bool attr_set = false;                                                               \
        if (!attr_set) {                                                                            \
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_set = true;                                                                        \
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
        static bool attr_set = false;                                                               \
        if (!attr_set) {                                                                            \
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_set = true;                                                                        \
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
    size_t smem = sizeof(ForceSmem);
    const bool eps0 = (a.eps2 == 0.f);   // unsoftened: the 2^-52 of the reference is handled by the verification
#define G6_LAUNCH(NN_, NR_)                                                                         \
    do {                                                                                            \
        if (eps0) G6_LAUNCH2(NN_, NR_, true); else G6_LAUNCH2(NN_, NR_, false);                     \
    } while (0)
#define G6_LAUNCH2(NN_, NR_, E0_)                                                                   \
    do {                                                                                            \
        auto kern = force_fast_kernel<IPT, NN_, NR_, MINB, E0_>;                                    \
        static bool attr_set = false;                                                               \
        if (!attr_set) {                                                                            \
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_set = true;                                                                        \
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
        static bool attr_set = false;                                                               \
        if (!attr_set) {                                                                            \
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_set = true;                                                                        \
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

// Launch the force kernel for i-block (iA,iB,iC) of ni particles against j in [0,nj).
// inline_src != nullptr: HOST pointer to the packed i-block ([3][ni] float4, stride ni), carried in the
// kernel parameters (small i-blocks; masked kernels).  flag_seq != 0: outputs are host-mapped and the
// launch raises G.h_flag = flag_seq when they are complete.
void launch_force(int nj, int ni, const float4 *iA, const float4 *iB, const float4 *iC, float eps2, bool nn,
                  bool list, double *out_sum, u64 *out_key, int *out_nnid, const float4 *inline_src = nullptr,
                  unsigned long long flag_seq = 0, const HermiteArgs *herm = nullptr)
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
    const int slots = G.sm_count * vi.ctas_per_sm;
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
    int nsplit = best_ns;
    if (G.force_nsplit > 0) nsplit = std::min(G.force_nsplit, ntiles);   // G6_B200_NSPLIT: experiments only
    int tps = (ntiles + nsplit - 1) / nsplit;
    nsplit = (ntiles + tps - 1) / tps;

    ForceArgs a{};
    // G.win_lo > 0: the launch works on the j-window [win_lo, win_lo + nj) of the local arrays (sharded
    // Hermite step on a replicated state); win_lo is a multiple of TILE, so tiles keep their id ranges
    a.jA = G.js.A + G.win_lo; a.jB = G.js.B + G.win_lo; a.jC = G.js.C + G.win_lo;
    a.iA = iA; a.iB = iB; a.iC = iC;
    a.ni = ni; a.nj = nj;
    a.tiles_per_split = tps; a.nsplit = nsplit;
    a.ni_pad = ni;
    a.j_offset = G.j_offset + G.win_lo;
    // many splits of few i-blocks: sum the partials with a kernel of its own (one warp per i, spread
    // over the SMs) instead of the last CTA -- except for the 4-particle shape, whose last CTA puts
    // 32 lanes on each i
    const bool defer = (nsplit > 32) && (n_iblocks * 8 <= slots) && (vi.ib > 4) && !herm;
    a.defer_reduce = defer ? 1 : 0;
    a.eps2 = eps2;
    if (nsplit > 1) ensure_partials((size_t)nsplit * ni);
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
    if (herm) {
        a.herm = *herm;
        if (v == V_T1) launch_herm<1, 4, false, 2>(a, grid, G.stream);
        else if (v == V_W1) launch_herm<1, 32, false, 2>(a, grid, G.stream);
        else launch_herm<2, 32, true, 2>(a, grid, G.stream);
    } else if (inline_src) {
        if (ni <= 64) {
            InlineI<64> ii;
            for (int k = 0; k < 3; k++) memcpy(ii.d + 64 * k, inline_src + (size_t)ni * k, sizeof(float4) * ni);
            if (v == V_T1) launch_inline<1, 4, false, 2, 64>(a, grid, ii, G.stream);
            else if (v == V_W1) launch_inline<1, 32, false, 2, 64>(a, grid, ii, G.stream);
            else launch_inline<2, 32, true, 2, 64>(a, grid, ii, G.stream);
        } else {
            InlineI<384> ii;
            for (int k = 0; k < 3; k++) memcpy(ii.d + 384 * k, inline_src + (size_t)ni * k, sizeof(float4) * ni);
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
    if (defer) {
        reduce_partials_kernel<<<reduce_ctas, 256, 0, G.stream>>>(a, nn ? 1 : 0);
        CK(cudaGetLastError());
        G.launches++;
    }
}

void free_all()
{
    for (int k = 0; k < 7; k++) dev_free(G.js.q[k]);
    dev_free(G.js.A); dev_free(G.js.B); dev_free(G.js.C);
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
        dev_free(H.d_ilist); dev_free(H.d_olddt);
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
    G.slot_of_addr.clear();
    G.predicted_nj = -1; G.j_dirty = false; G.pending = false;
    G.ngb_valid = G.ngb_built = G.ngb_fetched = false;
}

void stage_j(int address, int index, double tj, double mass, const double *j6, const double *a2, const double *v,
             const double *x)
{
    if (address < 0) {
        fprintf(stderr, "g6_b200: FATAL g6_set_j_particle address %d < 0\n", address);
        exit(-1);
    }
    ensure_capacity(address + 1);
    int slot = G.slot_of_addr[address];
    if (slot < 0) {  // last write wins within a batch (sapporo.cpp:83-110)
        ensure_up_cap(G.up_n + 1);
        slot = G.up_n++;
        G.slot_of_addr[address] = slot;
    }
    JUpdate &u = G.h_up[slot];
    for (int k = 0; k < 3; k++) {
        u.x[k] = x[k];
        u.v[k] = v[k];
        u.a[k] = 2.0 * a2[k];   // a2 = acc/2   (sapporo.cpp:94)
        u.j[k] = 6.0 * j6[k];   // j6 = jerk/6  (sapporo.cpp:95)
    }
    u.t = tj;
    u.m = (float)mass;
    u.id = index;
    u.addr = address;
    u.pad[0] = u.pad[1] = u.pad[2] = 0;
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

int g6_open_(int *id)
{
    int ndev = get_device_count();
    if (ndev <= 0) {
        fprintf(stderr, "g6_b200: FATAL no CUDA device found (this library has no CPU fallback)\n");
        exit(-1);
    }
    int dev = id ? *id : 0;
    if (env_int("G6_B200_DEVICE_MODULO", 0)) dev = ((dev % ndev) + ndev) % ndev;
    if (dev < 0 || dev >= ndev) {
        fprintf(stderr, "g6_b200: g6_open: no CUDA device with id %d (%d present)\n", dev, ndev);
        return -1;
    }
    if (G.open) {
        if (dev == G.device) return 0;
        int d = G.device;
        g6_close_(&d);
    }
    G.device = dev;
    CK(cudaSetDevice(dev));
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
    host_alloc(G.h_i, (size_t)3 * G.npipes);
    dev_alloc(G.d_i, (size_t)3 * G.npipes);
    dev_alloc(G.d_i2, (size_t)3 * G.npipes);
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
    return 0;
}

int g6_close_(int *id)
{
    (void)id;
    if (!G.open) return 0;
    CK(cudaSetDevice(G.device));
    CK(cudaStreamSynchronize(G.stream));
    g6x_peer_detach();
    if (G.trace && G.tr_calls)
        fprintf(stderr,
                "g6_b200 trace: %lld force calls; mean us per call: flush %.2f predict %.2f pack %.2f h2d %.2f "
                "launch %.2f d2h-issue %.2f wait %.2f unpack %.2f\n",
                G.tr_calls, 1e6 * G.tr[0] / G.tr_calls, 1e6 * G.tr[1] / G.tr_calls, 1e6 * G.tr[2] / G.tr_calls,
                1e6 * G.tr[3] / G.tr_calls, 1e6 * G.tr[4] / G.tr_calls, 1e6 * G.tr[5] / G.tr_calls,
                1e6 * G.tr[6] / G.tr_calls, 1e6 * G.tr[7] / G.tr_calls);
    free_all();
    if (G.own_stream) cudaStreamDestroy(G.own_stream);
    G.own_stream = G.stream = nullptr;
    G.open = false;
    return 0;
}

int g6_npipes_(void)
{
    if (G.open) return G.npipes;
    return std::max(1, env_int("G6_B200_NPIPES", 16384));
}

int g6_set_tunit_(void *unused) { (void)unused; return 0; }
int g6_set_xunit_(void *unused) { (void)unused; return 0; }

int g6_set_ti_(int *id, double *ti)
{
    (void)id;
    require_open("g6_set_ti_");
    G.ti = *ti;
    return 0;
}

int g6_set_j_particle_(int *cluster_id, int *address, int *index, double *tj, double *dtj, double *mass,
                       double k18[3], double j6[3], double a2[3], double v[3], double x[3])
{
    (void)cluster_id; (void)dtj; (void)k18;
    require_open("g6_set_j_particle_");
    stage_j(*address, *index, *tj, *mass, j6, a2, v, x);
    return 0;
}

void g6calc_firsthalf_(int *cluster_id, int *nj, int *ni, int index[], double xi[][3], double vi[][3],
                       double aold[][3], double j6old[][3], double phiold[], double *eps2, double h2[])
{
    (void)cluster_id; (void)aold; (void)j6old; (void)phiold;
    require_open("g6calc_firsthalf_");
    int n = *ni;
    if (n > G.npipes || n < 0) {
        fprintf(stderr, "g6_b200: FATAL g6calc_firsthalf ni = %d exceeds g6_npipes() = %d\n", n, G.npipes);
        exit(-1);
    }
    if (G.pending) CK(cudaStreamSynchronize(G.stream));  // a firsthalf without its lasthalf
    double t0 = G.trace ? wall() : 0.0, t1 = 0.0;
    // scatter (small batches: fused with the predictor, records read from mapped pinned memory) + predict
    const bool inl = (n > 0) && (n <= G.inline_max) && (G.variant == V_AUTO);
    predict_with_updates(*nj);
    G6_TR(0)
    G6_TR(1)
    // pack the i-block: double -> double-single (sapporo.cpp:125-134); the three float4 streams are
    // laid out back to back with stride n, so that they cross PCIe in ONE copy
    float4 *A = G.h_i, *B = G.h_i + n, *C = G.h_i + 2 * (size_t)n;
    bool any_h2 = false;
    for (int i = 0; i < n; i++) {
        double x = xi[i][0], y = xi[i][1], z = xi[i][2];
        float xh = (float)x, yh = (float)y, zh = (float)z;
        float hh = h2 ? (float)h2[i] : 0.f;
        any_h2 |= (hh > 0.f);
        A[i] = make_float4(xh, yh, zh, hh);
        union { int i; float f; } cv;
        cv.i = index[i];
        B[i] = make_float4((float)(x - (double)xh), (float)(y - (double)yh), (float)(z - (double)zh), cv.f);
        C[i] = make_float4((float)vi[i][0], (float)vi[i][1], (float)vi[i][2], 0.f);
    }
    G6_TR(2)
    // small i-blocks travel in the kernel parameters (no H2D copy); their results are written by the
    // kernel into mapped host memory and announced by a flag (no D2H copy, no stream synchronisation)
    const bool direct = (n > 0) && (n <= G.direct_max);
    if (n > 0 && !inl) CK(cudaMemcpyAsync(G.d_i, G.h_i, sizeof(float4) * 3 * (size_t)n, cudaMemcpyHostToDevice, G.stream));
    G.i_on_device = !inl;
    G.cur_direct = direct;
    G6_TR(3)
    G.cur_ni = n;
    G.cur_nj = *nj;
    G.cur_eps2 = (float)*eps2;
    G.cur_any_h2 = any_h2;
    G.pending = true;
    G.ngb_valid = G.ngb_fetched = false;
    // The force kernel starts here, asynchronously (GRAPE's firsthalf/lasthalf split exists for this
    // overlap): always with the nearest-neighbour search, which costs ~1 % and is simply not copied
    // back by g6calc_lasthalf_.  Neighbour-sphere lists are NOT built here: ph4 and phiGRAPE pass
    // h2 = eps2 on every call (gpu.cc:266,324; gravity.F:72) and never read the lists, so they are
    // built on demand by g6_read_neighbour_list_.
    // outputs: [7n doubles][n ints] back to back, so that they come back in ONE copy
    if (n > 0) {
        double *out = direct ? G.dev_h_sum : G.d_sum;
        launch_force(G.cur_nj, n, G.d_i, G.d_i + n, G.d_i + 2 * (size_t)n, G.cur_eps2, true, false, out, G.d_key,
                     reinterpret_cast<int *>(out + 7 * (size_t)n), inl ? G.h_i : nullptr,
                     direct ? ++G.flag_seq : 0ull);
    }
    G6_TR(4)
    if (G.trace) G.tr_calls++;
}

static int lasthalf_common(int nj, int ni, double acc[][3], double jerk[][3], double pot[], int *inn)
{
    require_open("g6calc_lasthalf_");
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
        for (int i = 0; i < ni; i++) {
            const double *s = G.h_sum + (size_t)7 * i;
            acc[i][0] = s[0]; acc[i][1] = s[1]; acc[i][2] = s[2];
            jerk[i][0] = s[3]; jerk[i][1] = s[4]; jerk[i][2] = s[5];
            pot[i] = -s[6];
            if (nn) inn[i] = h_nnid[i];
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

int g6_read_neighbour_list_(int *cluster_id)
{
    (void)cluster_id;
    require_open("g6_read_neighbour_list_");
    if (!G.ngb_valid) {
        G.ngb_fetched = false;
        return 0;  // no lists were requested (all h2 <= 0 or lasthalf without nn)
    }
    int ni = G.cur_ni;
    if (!G.ngb_built) {
        // second pass over the captured i-block (still in d_i; j state unchanged since its lasthalf2)
        // with the list-building variant of the masked kernel; forces go to scratch
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
            CK(cudaMemcpyAsync(G.d_i, G.h_i, sizeof(float4) * 3 * (size_t)ni, cudaMemcpyHostToDevice, G.stream));
            G.i_on_device = true;
        }
        CK(cudaMemsetAsync(G.d_ngb_cnt, 0, sizeof(int) * ni, G.stream));
        launch_force(G.cur_nj, ni, G.d_i, G.d_i + ni, G.d_i + 2 * (size_t)ni, G.cur_eps2, true, true, G.d_sum2,
                     G.d_key2, G.d_nnid2);
        G.ngb_built = true;
    }
    CK(cudaMemcpyAsync(G.h_ngb_cnt, G.d_ngb_cnt, sizeof(int) * ni, cudaMemcpyDeviceToHost, G.stream));
    CK(cudaMemcpyAsync(G.h_ngb_list, G.d_ngb_list, sizeof(int) * (size_t)ni * G.ngb_cap, cudaMemcpyDeviceToHost,
                       G.stream));
    CK(cudaStreamSynchronize(G.stream));
    G.ngb_fetched = true;
    int overflow = 0;
    for (int i = 0; i < ni; i++)
        if (G.h_ngb_cnt[i] > G.ngb_cap) overflow = 1;
    return overflow;
}

int g6_get_neighbour_list_(int *cluster_id, int *ipipe, int *maxlength, int *n_neighbours, int neighbour_list[])
{
    (void)cluster_id;
    require_open("g6_get_neighbour_list_");
    int ip = *ipipe;
    if (ip < 0 || ip >= G.cur_ni) {
        fprintf(stderr, "g6_b200: FATAL g6_get_neighbour_list ipipe = %d >= ni = %d\n", ip, G.cur_ni);
        exit(-1);  // as sapporo.cpp:254-258
    }
    if (!G.ngb_valid || !G.ngb_fetched) {
        *n_neighbours = 0;
        return 0;
    }
    int cnt = G.h_ngb_cnt[ip];
    int have = std::min(cnt, G.ngb_cap);
    int *src = G.h_ngb_list + (size_t)ip * G.ngb_cap;
    std::sort(src, src + have);
    int ncopy = std::min(have, *maxlength);
    memcpy(neighbour_list, src, sizeof(int) * ncopy);
    *n_neighbours = cnt;
    return (cnt > *maxlength || cnt > G.ngb_cap) ? 1 : 0;
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
    if (G.open) flush_updates();
}
int get_j_part_data(int addr, int nj, double *pos, double *vel, double *acc, double *jrk, double *ppos, double *pvel)
{
    require_open("get_j_part_data");
    flush_updates();
    if (addr < 0 || addr >= nj || addr >= G.capacity) return -1;
    double2 q[7];
    float4 abc[3];
    CK(cudaStreamSynchronize(G.stream));
    for (int k = 0; k < 7; k++) CK(cudaMemcpy(&q[k], G.js.q[k] + addr, sizeof(double2), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&abc[0], G.js.A + addr, sizeof(float4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&abc[1], G.js.B + addr, sizeof(float4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&abc[2], G.js.C + addr, sizeof(float4), cudaMemcpyDeviceToHost));
    if (pos) { pos[0] = q[0].x; pos[1] = q[0].y; pos[2] = q[1].x; }
    if (vel) { vel[0] = q[2].x; vel[1] = q[2].y; vel[2] = q[3].x; }
    if (acc) { acc[0] = q[3].y; acc[1] = q[4].x; acc[2] = q[4].y; }
    if (jrk) { jrk[0] = q[5].x; jrk[1] = q[5].y; jrk[2] = q[6].x; }
    if (ppos) { ppos[0] = (double)abc[0].x + abc[1].x; ppos[1] = (double)abc[0].y + abc[1].y; ppos[2] = (double)abc[0].z + abc[1].z; }
    if (pvel) { pvel[0] = abc[2].x; pvel[1] = abc[2].y; pvel[2] = abc[2].z; }
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
    G.refine = on ? 1 : 0;
    return 0;
}

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
    G.ti = ti;
    flush_updates();
    run_predictor(nj);
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
    P.seq++;
    ExchangeSlots e{};
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
                                                    P.dev_h_err);
    CK(cudaGetLastError());
    G.launches += 2;
}

// Shared body of the device-resident entry points.  With peers attached and `exchange` set, every launch
// writes this rank's partials into its slot of all exchange buffers (own + peers, over NVLink) and the
// call ends with the flag + combine kernels, so d_sum/d_key/d_nnid hold the totals over all j-shards.
static int calc_device_impl(int nj, int ni, const int *d_index, const double *d_xi, const double *d_vi,
                            const double *d_h2, double eps2, int flags, double *d_sum, unsigned long long *d_key,
                            int *d_nnid, bool exchange)
{
    flush_updates();
    run_predictor(nj);
    bool nn = (flags & 1) != 0, list = (flags & 2) != 0;
    if (list) {
        fprintf(stderr, "g6_b200: FATAL g6x_calc_device: neighbour lists are served by the g6 ABI path only\n");
        exit(-1);
    }
    Context::Peer &P = G.peer;
    ExchangeSlots ex{};
    if (exchange) ex = exchange_begin(ni, "g6x_calc_device_allreduce");
    const int chunk = device_chunk(ni, std::min(nj, G.capacity));
    if (chunk > G.i2_cap) {
        CK(cudaStreamSynchronize(G.stream));
        dev_free(G.d_i2);
        dev_alloc(G.d_i2, (size_t)3 * chunk);
        G.i2_cap = chunk;
    }
    for (int i0 = 0; i0 < ni; i0 += chunk) {
        int n = std::min(chunk, ni - i0);
        float4 *A = G.d_i2, *B = G.d_i2 + G.i2_cap, *C = G.d_i2 + 2 * (size_t)G.i2_cap;
        pack_i_kernel<<<(n + 255) / 256, 256, 0, G.stream>>>(n, d_index + i0, d_xi + 3 * (size_t)i0,
                                                              d_vi + 3 * (size_t)i0, d_h2 ? d_h2 + i0 : nullptr, A, B,
                                                              C);
        G.launches++;
        CK(cudaGetLastError());
        if (!exchange) {
            launch_force(nj, n, A, B, C, (float)eps2, nn, false, d_sum + 7 * (size_t)i0, d_key + i0, d_nnid + i0);
            continue;
        }
        exchange_set_mirrors(ex, i0);
        unsigned char *own = ex.half[P.rank];
        launch_force(nj, n, A, B, C, (float)eps2, nn, false, slot_sum(own, i0), slot_key(own, i0), slot_id(own, i0));
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
        if (P.attached && r != P.rank && P.peer_buf[r]) cudaIpcCloseMemHandle(P.peer_buf[r]);
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
    return device_chunk(ni, std::min(G.nj_hi, G.capacity));
}

int g6x_resolve_nn(int ni, const unsigned long long *d_key, int rank, int *d_nnid)
{
    require_open("g6x_resolve_nn");
    if (ni <= 0) return 0;
    resolve_nn_kernel<<<(ni + 255) / 256, 256, 0, G.stream>>>(ni, d_key, rank, G.j_offset,
                                                               std::min(G.nj_hi, G.capacity), G.js.B, d_nnid);
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
    dev_free(H.d_ilist); dev_free(H.d_olddt);
    dev_alloc(H.d_ilist, (size_t)cap);
    dev_alloc(H.d_olddt, (size_t)cap);
    host_alloc(H.h_ilist, cap);   H.dev_h_ilist = dev_alias(H.h_ilist);
    host_alloc(H.h_olddt, cap);   H.dev_h_olddt = dev_alias(H.h_olddt);
    host_alloc(H.h_outdt, cap);   H.dev_h_outdt = dev_alias(H.h_outdt);
    host_alloc(H.h_outpot, cap);  H.dev_h_outpot = dev_alias(H.h_outpot);
    host_alloc(H.h_outnn, cap);   H.dev_h_outnn = dev_alias(H.h_outnn);
    dev_alloc(H.d_pred, (size_t)6 * cap);
    dev_alloc(H.d_i, (size_t)3 * cap);
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
    h.iA = H.d_i; h.iB = H.d_i + n; h.iC = H.d_i + 2 * (size_t)n;
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
        // replicated state, sharded forces: gather the block (every rank, identical), predict only this
        // rank's j-window, sum over it with the partials mirrored into the peers' exchange buffers, combine,
        // and let every rank correct its own replica with the identical totals
        int lo = H.shard_lo, hi = std::min(H.shard_hi, std::min(nj, G.capacity));
        if (hi <= lo) lo = hi = 0;   // empty window (more ranks than j-tiles): this rank contributes zeros
        hermite_gather_kernel<<<ctas, 256, 0, G.stream>>>(h);
        CK(cudaGetLastError());
        if (hi > lo) {
            const int tile0 = lo / TILE, ntiles = (hi - lo + TILE - 1) / TILE;
            const int npred = std::max(std::min(nj, G.capacity), std::min(G.nj_hi, G.capacity));
            predict_kernel<<<ntiles, TILE, 0, G.stream>>>(std::min(npred, hi), tnext, G.js, tile0);
            CK(cudaGetLastError());
        }
        ExchangeSlots ex = exchange_begin(n, "g6x_hermite_step (sharded)");
        exchange_set_mirrors(ex, 0);
        unsigned char *own = ex.half[G.peer.rank];
        G.win_lo = lo;
        launch_force(std::max(0, hi - lo), n, h.iA, h.iB, h.iC, (float)eps2, true, false, slot_sum(own, 0),
                     slot_key(own, 0), slot_id(own, 0));
        G.win_lo = 0;
        G.mir_n = 0;
        exchange_finish(n, H.d_sum, reinterpret_cast<unsigned long long *>(H.d_key), H.d_nnid);
        hermite_correct_kernel<<<ctas, 256, 0, G.stream>>>(h);
        CK(cudaGetLastError());
        G.launches += 3;
        G.predicted_nj = -1;   // only a window was predicted
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
        launch_force(nj, n, h.iA, h.iB, h.iC, (float)eps2, true, false, H.d_sum, H.d_key, H.d_nnid, nullptr,
                     h.flag_seq, &h);
        G.launches += 1;
    } else {
        hermite_gather_kernel<<<ctas, 256, 0, G.stream>>>(h);
        CK(cudaGetLastError());
        run_predictor(nj);
        launch_force(nj, n, h.iA, h.iB, h.iC, (float)eps2, true, false, H.d_sum, H.d_key, H.d_nnid);
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
    // the whole block in one pass, whatever its size: every force is computed against the PREDICTED state
    // of all j before any particle is corrected, as in idata::advance
    hermite_reserve(ni);
    memcpy(H.h_ilist, ilist, sizeof(int) * ni);
    memcpy(H.h_olddt, old_dt, sizeof(double) * ni);
    hermite_pass(nj, ni, tnext, eta, eps2, 0);
    memcpy(new_dt, H.h_outdt, sizeof(double) * ni);
    if (pot) memcpy(pot, H.h_outpot, sizeof(double) * ni);
    if (nn) memcpy(nn, H.h_outnn, sizeof(int) * ni);
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
    hermite_reserve(nj);
    H.time.assign(nj, t0);
    H.dt.assign(nj, 0.0);
    for (int k = 0; k < nj; k++) H.h_ilist[k] = k;
    hermite_pass(nj, nj, t0, eta, eps2, 1);   // forces of all particles at t0, first steps (jdata.cc:503-548)
    memcpy(H.dt.data(), H.h_outdt, sizeof(double) * nj);
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
    for (int i = 0; i < nj; i++) {
        if (x) { x[i][0] = q[0][i].x; x[i][1] = q[0][i].y; x[i][2] = q[1][i].x; }
        if (t) t[i] = q[1][i].y;
        if (v) { v[i][0] = q[2][i].x; v[i][1] = q[2][i].y; v[i][2] = q[3][i].x; }
        if (a) { a[i][0] = q[3][i].y; a[i][1] = q[4][i].x; a[i][2] = q[4][i].y; }
        if (j) { j[i][0] = q[5][i].x; j[i][1] = q[5][i].y; j[i][2] = q[6][i].x; }
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
    std::vector<float4> A(nj), B(nj), C(nj);
    CK(cudaStreamSynchronize(G.stream));
    CK(cudaMemcpy(A.data(), G.js.A, sizeof(float4) * nj, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(B.data(), G.js.B, sizeof(float4) * nj, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(C.data(), G.js.C, sizeof(float4) * nj, cudaMemcpyDeviceToHost));
    for (int j = 0; j < nj; j++) {
        pos[j][0] = (double)A[j].x + (double)B[j].x;
        pos[j][1] = (double)A[j].y + (double)B[j].y;
        pos[j][2] = (double)A[j].z + (double)B[j].z;
        vel[j][0] = C[j].x; vel[j][1] = C[j].y; vel[j][2] = C[j].z;
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
    G.variant = variant;
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
