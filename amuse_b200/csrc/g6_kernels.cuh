// g6_kernels.cuh -- hand-written sm_100a kernels of the B200 g6 force library.
//
//   predict_kernel          j-particle Hermite predictor                  (HBM-bound)
//   scatter_kernel          j-update scatter                              (HBM/latency-bound)
//   update_predict_kernel   scatter of a small update batch + predictor in one launch
//   pack_i_kernel           double -> double-single i-block packing (device callers)
//   force_fast_kernel<>     Hermite force, big i-blocks: acc, jerk, pot, nearest neighbour; (warp of i) x (group of j)
//                           blocks classified by bounding boxes into FAR (hi-part differences), NEAR (mask-free
//                           double-single) and CLOSE (FP64, the reference's rule)       (FP32-pipe-bound)
//   near_kernel             its pre-pass: id-table lookup + nearest-neighbour bound per i-particle
//   order_*_kernel, hash_insert_kernel, i_key_kernel   Morton order of the j-memory, id -> slot table
//   force_kernel<>          Hermite force, small i-blocks and neighbour-sphere lists: j split over the warps of a
//                           CTA, masks per pair; optional i-block in the kernel parameters, results and completion
//                           flag in mapped host memory, corrector in the output stage   (latency-bound)
//   reduce_partials_kernel  sum of j-split partials, one warp per i
//   resolve_nn_kernel       id lookup after a cross-rank min-reduction (NCCL path)
//   peer_flag_kernel, peer_combine_kernel   multi-GPU exchange over peer memory (the force kernels store their
//                           partials into the peers' buffers themselves, see store_outputs)
//   hermite_*_kernel        device-resident Hermite block step: i-predictor, corrector + Aarseth step, write-back
//   fp32_peak_kernel<>, latency_probe_kernel   roofline / latency-floor probes
//
// Reference behaviour being reproduced (not translated):
//   force loop   src/amuse_ph4/src/idata.cc:198-236 (oracle, FP64)
//   predictor    src/amuse_ph4/src/jdata.cc:726-747, i-predictor idata.cc:347-365 (oracle, FP64)
//   corrector    src/amuse_ph4/src/idata.cc:443-511, first step jdata.cc:503-548 (oracle, FP64)
//   reduction    src/amuse_ph4/src/idata.cc:284-313 (sum / min / owner's nn over the j-domains)
//   API/semantics lib/sapporo_light/dev_evaluate_gravity.cu:46-106 (DS positions,
//                self-exclusion by id :76-79, neighbour rule :60-72)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace g6b {

constexpr int THREADS = 256;   // threads per force CTA (8 warps)
constexpr int TILE = 256;      // j-particles per shared-memory stage
constexpr int STAGES = 3;      // TMA bulk-copy pipeline depth
constexpr int GROUPS_PER_TILE = TILE / 32;
constexpr int TBOX = 2 * GROUPS_PER_TILE + 2;   // float4 per tile in gbb: 8 group boxes + the tile's own box
#ifndef G6_DENSE_ROUNDS
#define G6_DENSE_ROUNDS 4   // masked kernels, dense FP64 pass: particles per warp whose lanes are summed with shuffles first
#endif
constexpr int CTA_WL = 768;    // FP64 pairs a CTA of a masked kernel can queue (more are evaluated in place)
#ifndef G6_FLUSH
#define G6_FLUSH 16
#endif
#ifndef G6_UNROLL
#define G6_UNROLL 2
#endif
constexpr int UNROLL = G6_UNROLL;  // j-iterations unrolled in the hot loop
constexpr int FLUSH = G6_FLUSH;  // pairs summed in FP32 before a flush to the FP64 totals
constexpr float TINYF = 2.220446049250313e-16f;  // 2^-52, stdinc.h:33
constexpr float FAR_AWAY = 1.0e18f;  // where massless / unused j are parked
constexpr unsigned long long KEY_NONE = 0x7f800000ffffffffULL;

// ---------------------------------------------------------------------------
// j state in HBM (capacity C, padded to a multiple of TILE), indexed by SLOT.  The library keeps
// the j-memory in Morton order of the positions (rebuilt after bulk loads, see "j-memory order"
// below); the caller's addresses are translated through slot_of[] when updates arrive.  All FP64
// like the reference's jdata arrays, packed as seven double2 streams (7 x LDG.128 per j):
//   q0 = (x, y)    q1 = (z, t_j)   q2 = (vx, vy)   q3 = (vz, ax)
//   q4 = (ay, az)  q5 = (jx, jy)   q6 = (jz, mass)
//   ia = (id, address reported in nearest-neighbour keys)
// predicted j (what the force kernels stream), float4 each, positions relative to the origin x0:
//   A = (x.hi, y.hi, z.hi, mass)  B = (x.lo, y.lo, z.lo, id bits)  C = (vx, vy, vz, address bits)
//   L = (vx.lo, vy.lo, vz.lo, mass.lo)   -- read by the FP64 pairs only
// and per tile TBOX float4 of bounding boxes of the hi parts of the massive members, (min x, min y,
// min z, largest |coordinate|), (max x, max y, max z, 0): eight group boxes (32 slots each), then the tile's.
// ---------------------------------------------------------------------------
struct JState {
    double2 *q[7];
    int2 *ia;
    float4 *A, *B, *C, *L;
    float4 *gbb;
    int *slot_of;        // [address] -> slot
    float *near2;        // [slot] squared distance to the nearest neighbour when the particle was last an
                         // i-particle (or an upper bound from the Morton window at the last re-ordering)
    float *capr2;        // [slot] largest FP64 radius^2 the particle may use: the radius within which it touches no
                         // more than gmax group boxes (order_cap_kernel; one buffer shared by both array sets)
    double x0[3];        // origin subtracted from every position before the hi/lo split
};

// One staged j-update (host pinned -> device staging -> scatter_kernel).
struct __align__(16) JUpdate {
    double x[3];
    double v[3];
    double a[3];
    double j[3];
    double t;
    double m;
    int id;
    int slot;      // where it goes (the host keeps a mirror of slot_of[])
    int kaddr;     // address reported in nearest-neighbour keys
    int addr;      // the caller's (local) address
};
static_assert(sizeof(JUpdate) == 128, "JUpdate layout");

__device__ __forceinline__ void store_update(const JUpdate &u, const JState &s)
{
    const int a = u.slot;
    s.q[0][a] = make_double2(u.x[0], u.x[1]);
    s.q[1][a] = make_double2(u.x[2], u.t);
    s.q[2][a] = make_double2(u.v[0], u.v[1]);
    s.q[3][a] = make_double2(u.v[2], u.a[0]);
    s.q[4][a] = make_double2(u.a[1], u.a[2]);
    s.q[5][a] = make_double2(u.j[0], u.j[1]);
    s.q[6][a] = make_double2(u.j[2], u.m);
    s.ia[a] = make_int2(u.id, u.kaddr);
}
__global__ void scatter_kernel(int n, const JUpdate *__restrict__ up, JState s)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    store_update(up[k], s);
}

// identity order / no neighbour bound for freshly allocated slots [lo, hi)
__global__ void order_fill_kernel(const int lo, const int hi, int *slot_of, int *addr_of, float *near2)
{
    const int j = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= hi) return;
    slot_of[j] = j;
    addr_of[j] = j;
    near2[j] = __int_as_float(0x7f800000);
}

// floats <-> integers with the same ordering (for min/max reductions with integer instructions)
__device__ __forceinline__ int f2ord(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// FP64 Hermite prediction of slot j to time ti (jdata.cc:726-747), same expression tree.
struct PredJ {
    double x, y, z, vx, vy, vz, m;
};
__device__ __forceinline__ PredJ predict_slot(const JState &s, const int j, const double ti)
{
    const double2 q0 = s.q[0][j], q1 = s.q[1][j], q2 = s.q[2][j], q3 = s.q[3][j], q4 = s.q[4][j], q5 = s.q[5][j],
                  q6 = s.q[6][j];
    const double x = q0.x, y = q0.y, z = q1.x, tj = q1.y, vx = q2.x, vy = q2.y, vz = q3.x;
    const double ax = q3.y, ay = q4.x, az = q4.y, jx = q5.x, jy = q5.y, jz = q6.x;
    const double dt = ti - tj;
    PredJ p{x, y, z, vx, vy, vz, q6.y};
    if (dt != 0.0) {  // same expression tree as jdata.cc:739-746
        p.x = x + dt * (vx + 0.5 * dt * (ax + dt * jx / 3));
        p.y = y + dt * (vy + 0.5 * dt * (ay + dt * jy / 3));
        p.z = z + dt * (vz + 0.5 * dt * (az + dt * jz / 3));
        p.vx = vx + dt * (ax + 0.5 * dt * jx);
        p.vy = vy + dt * (ay + 0.5 * dt * jy);
        p.vz = vz + dt * (az + 0.5 * dt * jz);
    }
    return p;
}

// Hermite predictor (jdata.cc:726-747) in FP64, output split to double-single.
// Algorithmic traffic: 120 B read (seven double2 + id/address) + 65 B written per j.
// One CTA = one j-tile (TILE = 256 slots; the arrays are padded to whole tiles), one warp = one
// group of 32 slots, whose bounding box it records for the force kernel's far/close decisions.
// sh: 7 x 8 ints of shared memory.
__device__ __forceinline__ void predict_tile(const int tile, const int n, const double ti, const JState &s)
{
    __shared__ int sh[7][TILE / 32];
    const int j = tile * TILE + threadIdx.x;   // always < capacity
    PredJ p = predict_slot(s, j, ti);
    const int2 ia = s.ia[j];
    const bool massive = (j < n) && (p.m > (double)TINYF);
    double px = p.x - s.x0[0], py = p.y - s.x0[1], pz = p.z - s.x0[2];
    if (!massive) {  // massless or never-set slot: park it (idata.cc:208)
        px = py = pz = (double)FAR_AWAY;
        p.vx = p.vy = p.vz = 0.0;
        p.m = 0.0;
    }
    // slots in [n, end of tile) are written too (parked): the force kernel bounds its j loop by nj anyway
    const float xh = (float)px, yh = (float)py, zh = (float)pz;
    const float xl = (float)(px - (double)xh), yl = (float)(py - (double)yh), zl = (float)(pz - (double)zh);
    const float vxh = (float)p.vx, vyh = (float)p.vy, vzh = (float)p.vz, mh = (float)p.m;
    s.A[j] = make_float4(xh, yh, zh, mh);
    s.B[j] = make_float4(xl, yl, zl, __int_as_float(ia.x));
    s.C[j] = make_float4(vxh, vyh, vzh, __int_as_float(ia.y));
    s.L[j] = make_float4((float)(p.vx - (double)vxh), (float)(p.vy - (double)vyh), (float)(p.vz - (double)vzh),
                         (float)(p.m - (double)mh));
    // bounding box of the group's massive members (empty group: an inverted box nothing is close to)
    const int big = 0x7f7fffff, small = f2ord(-3.0e38f);   // +-FLT_MAX-ish
    const int lx = __reduce_min_sync(0xffffffffu, massive ? f2ord(xh) : big);
    const int ly = __reduce_min_sync(0xffffffffu, massive ? f2ord(yh) : big);
    const int lz = __reduce_min_sync(0xffffffffu, massive ? f2ord(zh) : big);
    const int hx = __reduce_max_sync(0xffffffffu, massive ? f2ord(xh) : small);
    const int hy = __reduce_max_sync(0xffffffffu, massive ? f2ord(yh) : small);
    const int hz = __reduce_max_sync(0xffffffffu, massive ? f2ord(zh) : small);
    const float amax = massive ? fmaxf(fmaxf(fabsf(xh), fabsf(yh)), fabsf(zh)) : 0.f;
    const int sc = __reduce_max_sync(0xffffffffu, __float_as_int(amax));   // non-negative floats order like ints
    const int w = threadIdx.x >> 5;
    float4 *tb = s.gbb + (size_t)tile * TBOX;
    if ((threadIdx.x & 31) == 0) {
        tb[2 * w] = make_float4(ord2f(lx), ord2f(ly), ord2f(lz), __int_as_float(sc));
        tb[2 * w + 1] = make_float4(ord2f(hx), ord2f(hy), ord2f(hz), 0.f);
        sh[0][w] = lx; sh[1][w] = ly; sh[2][w] = lz; sh[3][w] = hx; sh[4][w] = hy; sh[5][w] = hz; sh[6][w] = sc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int r[7];
#pragma unroll
        for (int q = 0; q < 7; q++) r[q] = sh[q][0];
#pragma unroll
        for (int g = 1; g < TILE / 32; g++) {
#pragma unroll
            for (int q = 0; q < 3; q++) r[q] = min(r[q], sh[q][g]);
#pragma unroll
            for (int q = 3; q < 7; q++) r[q] = max(r[q], sh[q][g]);
        }
        tb[2 * GROUPS_PER_TILE] = make_float4(ord2f(r[0]), ord2f(r[1]), ord2f(r[2]), __int_as_float(r[6]));
        tb[2 * GROUPS_PER_TILE + 1] = make_float4(ord2f(r[3]), ord2f(r[4]), ord2f(r[5]), 0.f);
    }
}

// Pending j-updates applied by the kernel that predicts (small batches, block-timestep regime): the
// slots ride in the kernel parameters, the 128-byte records are read from mapped pinned host
// memory by the threads whose slot falls into the CTA's j range [j_lo, j_hi).
constexpr int UPD_MAX = 256;
struct InlineU {
    int n;
    int slot[UPD_MAX];
};
__device__ __forceinline__ void apply_updates(const InlineU &iu, const JUpdate *__restrict__ rec, const JState &s,
                                              const int j_lo, const int j_hi)
{
    for (int k = threadIdx.x; k < iu.n; k += blockDim.x) {
        const int a = iu.slot[k];
        if (a >= j_lo && a < j_hi) store_update(rec[k], s);
    }
    __syncthreads();   // the block's own global writes are visible to its threads after the barrier
}

#ifndef G6_PRED_MINB
#define G6_PRED_MINB 5   // CTAs per SM the predictor is compiled for (48 registers): loads in flight, not arithmetic, bound it
#endif
__global__ void __launch_bounds__(TILE, G6_PRED_MINB) predict_kernel(int n, double ti, JState s, int tile0 = 0)
{
    predict_tile(tile0 + blockIdx.x, n, ti, s);
}

// scatter + predict in one launch (small update batches).
__global__ void __launch_bounds__(TILE) update_predict_kernel(int n, double ti, JState s, const JUpdate *rec,
                                                              const __grid_constant__ InlineU iu)
{
    apply_updates(iu, rec, s, blockIdx.x * TILE, (blockIdx.x + 1) * TILE);
    predict_tile(blockIdx.x, n, ti, s);
}

// i-block packing for device-resident callers: double -> double-single, relative to the origin.
//   iA = (x.hi,y.hi,z.hi,h2)  iB = (x.lo,y.lo,z.lo,id bits)  iC = (vx,vy,vz,0)  iD = (vx.lo,vy.lo,vz.lo,d2)
// d2 (squared distance to a neighbour: an upper bound of the nearest-neighbour distance) is filled
// in by near_kernel.  perm != NULL: packed slot k holds caller particle perm[k] (Morton order).
__global__ void pack_i_kernel(int ni, const int *__restrict__ perm, const int *__restrict__ index,
                              const double *__restrict__ xi, const double *__restrict__ vi,
                              const double *__restrict__ h2, const double x0, const double y0, const double z0,
                              float4 *iA, float4 *iB, float4 *iC, float4 *iD)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ni) return;
    const int i = perm ? perm[k] : k;
    const double x = xi[3 * (size_t)i] - x0, y = xi[3 * (size_t)i + 1] - y0, z = xi[3 * (size_t)i + 2] - z0;
    const double vx = vi[3 * (size_t)i], vy = vi[3 * (size_t)i + 1], vz = vi[3 * (size_t)i + 2];
    const float xh = (float)x, yh = (float)y, zh = (float)z;
    const float vxh = (float)vx, vyh = (float)vy, vzh = (float)vz;
    iA[k] = make_float4(xh, yh, zh, h2 ? (float)h2[i] : 0.f);
    iB[k] = make_float4((float)(x - (double)xh), (float)(y - (double)yh), (float)(z - (double)zh),
                        __int_as_float(index[i]));
    iC[k] = make_float4(vxh, vyh, vzh, 0.f);
    iD[k] = make_float4((float)(vx - (double)vxh), (float)(vy - (double)vyh), (float)(vz - (double)vzh), 0.f);
}

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA 1-D bulk copy, packed f32x2 math, rsqrt.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// One Newton-Raphson step on MUFU.RSQ (max rel. error 2^-22.9 -> ~1 ulp):
//   e = 1 - x*y0^2 ; y = y0 + (y0/2)*e        (+4 FP32 ops per pair, see DESIGN.md "accuracy")
__device__ __forceinline__ float rsqrt_refined(float x)
{
    float y0 = rsqrt_approx(x);
    float e = fmaf(-x, y0 * y0, 1.0f);
    return fmaf(0.5f * y0, e, y0);
}
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi)
{
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// ---------------------------------------------------------------------------
// j-memory order: where an i-particle sits among the j, and how close its neighbours are.
//
// The slots are kept in Morton order of the positions (host: rebuild_order).  Three things follow:
//  * the force kernel for big i-blocks, whose i-particles are Morton-sorted too, can classify whole
//    (warp of i) x (group of 32 j) blocks by their bounding boxes: FAR blocks take position differences
//    from the hi parts alone, CLOSE blocks are evaluated in FP64 exactly as the reference does;
//  * an open-addressing table id -> slot finds the j-particle an i-particle IS (same id), which is the
//    only j its self-exclusion rule can apply to;
//  * near2[slot] remembers the squared nearest-neighbour distance of every particle, which sets the
//    radius inside which its pairs are evaluated in FP64 (the few close pairs that dominate acc and
//    jerk and would otherwise carry the 2^-24 rounding of dx into a cancelling sum).
// ---------------------------------------------------------------------------
struct OrderInfo {
    const u64 *hash;       // entries (id << 32 | slot + 2); 0 = empty; low word 1 = several slots carry this id
    unsigned hash_mask;    // table size - 1 (a power of two); 0 = no table
    const unsigned *keys;  // [nkeys] Morton keys of slots [0, nkeys), ascending
    int nkeys;
    float blo[3], binv[3]; // Morton grid: cell = (x - blo) * binv, 1024 cells per axis (x relative to x0)
    float cap2;            // upper bound of the close radius squared
    float kclose;          // K: pairs closer than sqrt(K) x (nearest-neighbour distance) go to FP64; 0 = off
    float farc2;           // FAR blocks: box gap^2 > farc2 * (largest |coordinate|)^2
};

__device__ __forceinline__ unsigned hash_id(int id)
{
    unsigned h = (unsigned)id * 2654435761u;
    return h ^ (h >> 15);
}
// slot of the j-particle with this id: >= 0, -1 = none, -2 = several
__device__ __forceinline__ int hash_lookup(const OrderInfo &o, const int id)
{
    if (o.hash_mask == 0u) return -1;
    unsigned h = hash_id(id) & o.hash_mask;
    for (unsigned probe = 0; probe <= o.hash_mask; probe++) {
        const u64 e = o.hash[h];
        if (e == 0ull) return -1;
        if ((int)(unsigned)(e >> 32) == id) {
            const unsigned lowv = (unsigned)(e & 0xffffffffu);
            return lowv == 1u ? -2 : (int)lowv - 2;
        }
        h = (h + 1u) & o.hash_mask;
    }
    return -1;
}
__global__ void hash_insert_kernel(const int n, const JState s, u64 *table, const unsigned mask)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    if (!(s.q[6][slot].y > (double)TINYF)) return;   // massless / unset slots exert no force: nothing to exclude
    const int id = s.ia[slot].x;
    const u64 mine = ((u64)(unsigned)id << 32) | (u64)(unsigned)(slot + 2);
    unsigned h = hash_id(id) & mask;
    for (unsigned probe = 0; probe <= mask; probe++) {
        const u64 old = atomicCAS(&table[h], 0ull, mine);
        if (old == 0ull) return;
        if ((int)(unsigned)(old >> 32) == id) {   // a second slot with this id: the i-particles that carry it take
            atomicExch(&table[h], ((u64)(unsigned)id << 32) | 1ull);   // the exact path for every j
            return;
        }
        h = (h + 1u) & mask;
    }
}

__device__ __forceinline__ unsigned spread3(unsigned v)   // 10 bits -> every third bit
{
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ unsigned morton30(const float x, const float y, const float z, const float *blo,
                                             const float *binv)
{
    const float fx = fminf(fmaxf((x - blo[0]) * binv[0], 0.f), 1023.f);
    const float fy = fminf(fmaxf((y - blo[1]) * binv[1], 0.f), 1023.f);
    const float fz = fminf(fmaxf((z - blo[2]) * binv[2], 0.f), 1023.f);
    return spread3((unsigned)fx) | (spread3((unsigned)fy) << 1) | (spread3((unsigned)fz) << 2);
}
// first slot whose key is >= key (binary search over the sorted keys), clamped to a valid slot
__device__ __forceinline__ int key_search(const OrderInfo &o, const unsigned key)
{
    int lo = 0, hi = o.nkeys;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (o.keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo < o.nkeys ? lo : o.nkeys - 1;
}

// ---------------------------------------------------------------------------
// Force kernel.
// ---------------------------------------------------------------------------
// Arguments of the device-resident Hermite step (see "Device-resident Hermite block step" below).
struct HermiteArgs {
    int ni;
    const int *ilist;          // j-addresses of the active particles (mapped pinned host memory or device)
    const double *old_dt;      // their current time steps (same memory); unused by the init pass
    const int *slot_of;        // address -> slot
    int *ilist_d;              // device copies made by the gather pass (SLOTS), read by the corrector
    double *olddt_d;
    double tnext, eta;
    JState js;
    float4 *iA, *iB, *iC, *iD; // packed i-block for the force kernels
    double *pred;              // [ni][6] predicted pos, vel (FP64) kept for the corrector
    const double *sum;         // [ni][7] force-kernel output: acc, jerk, +sum m/r
    const int *nnid;           // [ni]
    double *out_dt;            // [ni] new time step          (mapped pinned host memory)
    double *out_pot;           // [ni] potential (negative)   (mapped pinned host memory)
    int *out_nn;               // [ni] id of the nearest neighbour
    int mode;                  // 0: corrector;  1: initialisation (a, j <- forces; first time step, jdata.cc:503-548)
    unsigned int *done_counter;
    unsigned long long *host_flag;
    unsigned long long flag_seq;
};

constexpr int MAX_PEERS = 7;   // other ranks of one NVSwitch domain (8 GPUs)
struct ForceArgs {
    const float4 *jA, *jB, *jC;   // predicted j (device), offset to the first slot of this launch's j-window
    const float4 *jG;             // group boxes, same offset (in groups)
    const float4 *iA, *iB, *iC, *iD;   // packed i-block (device)
    int ni, nj;                   // i count, j prefix [0, nj) of the window
    int tiles_per_split, nsplit;  // j decomposition over blockIdx.x
    int ni_pad;                   // stride of the partial workspace
    int j_offset;                 // added to the addresses reported in nearest-neighbour keys
    int slot0;                    // first slot of the j-window (a multiple of TILE)
    int defer_reduce;             // 1: stop after the per-split partials (reduce_partials_kernel follows)
    float eps2;
    const float4 *jL;             // lo parts of the predicted velocities and masses (FP64 pairs), same offset
    JState js;                    // the whole j-memory (not offset): near2, predicted arrays by absolute slot
    OrderInfo ord;
    int *conf;                    // [ni] slot of the j-particle with particle i's id (-1 none, -2 several): read by the
                                  // speculative kernel (near_kernel wrote it), written by the masked kernels
    const int *iperm;             // outputs of packed particle i go to index iperm[i] (NULL: i)
    unsigned long long *stats;    // -DG6_STATS builds: (warp x group) blocks taken FAR / NEAR / CLOSE, NEAR redone
    // FP64 pairs are not evaluated where they are found (one lane of a warp would hold up the other 31 for a
    // thousand cycles: FP64 issues at 1/64 of the FP32 rate on this chip) but collected and evaluated densely,
    // one pair per thread: the speculative kernel appends (packed i, slot) to a global list that
    // close_pairs_kernel works off, the masked kernels keep a list per CTA in shared memory.  Either way the
    // results are added into corr[] with FP64 atomics, and whoever writes particle i's outputs adds corr[i]
    // (and clears it for the next launch).
    int2 *wl;                     // global list (speculative kernel)
    unsigned int *wl_count;       // entries appended (zero between launches)
    unsigned int wl_cap;
    double *corr;                 // [ni][7], zero between launches (NULL: FP64 pairs are evaluated in place)
    double *part_sum;             // [nsplit][ni_pad][7]
    u64 *part_key;                // [nsplit][ni_pad]   (min r2 bits << 32 | slot)
    unsigned int *tickets;        // [gridDim.y], zero between launches
    double *out_sum;              // [ni][7]: acc xyz, jerk xyz, +sum m/r
    u64 *out_key;                 // [ni]   (min r2 bits << 32 | reported address)
    int *out_nnid;                // [ni]
    int *ngb_cnt;                 // [ni]   (LIST only; zeroed by the host)
    int *ngb_list;                // [ni][ngb_cap]
    int ngb_cap;
    // latency path: out_sum/out_nnid point into mapped pinned HOST memory and the CTA that writes the
    // last outputs raises a flag there, so the host neither issues a D2H copy nor synchronises the stream
    unsigned int *done_counter;          // device, zero between launches
    unsigned long long *host_flag;       // mapped pinned host memory (NULL: no signal)
    unsigned long long flag_seq;         // value to write
    unsigned int done_expected;          // CTAs that write final outputs in this launch
    // multi-GPU exchange fused into the force kernel: whoever writes final outputs of this device's
    // j-shard also stores them into its slot of every peer's exchange buffer over NVLink (peer pointers),
    // so the partials travel while the other i-blocks are still being computed
    // device-resident Hermite step, small blocks: whoever writes particle i's final force also runs its
    // corrector (HERM kernels), so a block step is two launches (predict+gather, force+correct)
    HermiteArgs herm;
    int n_mirror;
    double *m_sum[MAX_PEERS];            // [ni][7] at each peer, already offset to this launch's first i
    u64 *m_key[MAX_PEERS];
    int *m_id[MAX_PEERS];
};

// i-block carried in the kernel parameters (constant bank) for small i-blocks: no H2D copy at all.
// Layout [4][N]: iA, iB, iC, iD.  N == 0 is a 64-byte dummy.
template <int N>
struct InlineI {
    float4 d[4 * (N > 0 ? N : 1)];
};

__device__ __forceinline__ void hermite_correct_one(const HermiteArgs &h, const int i, const double *f, const int nnid);

// Final outputs of particle i (local arrays + the peers' exchange slots).  kk carries a slot of this
// launch's j-window; what leaves the kernel carries the address the caller gave that particle.
template <bool NN, bool HERM = false>
__device__ __forceinline__ void store_outputs(const ForceArgs &p, const int i, const double *tot_in, const u64 kk)
{
    double tot[7];
#pragma unroll
    for (int q = 0; q < 7; q++) tot[q] = tot_in[q];
    if (p.corr) {   // the FP64 pairs of this particle (all of them: their atomics precede this point, see ForceArgs)
#pragma unroll
        for (int q = 0; q < 7; q++) {
            tot[q] += __ldcg(p.corr + (size_t)i * 7 + q);
            p.corr[(size_t)i * 7 + q] = 0.0;
        }
    }
    int id = -1;
    u64 ko = KEY_NONE;
    if (NN && kk != KEY_NONE) {
        const int sl = (int)(unsigned)(kk & 0xffffffffu);
        id = __float_as_int(p.jB[sl].w);
        ko = (kk & 0xffffffff00000000ull) | (u64)(unsigned)(__float_as_int(p.jC[sl].w) + p.j_offset);
        // remember how close this particle's nearest neighbour is (sets its FP64 radius next time)
        const int me = p.conf ? p.conf[i] : -1;
        if (me >= 0) p.js.near2[me] = __int_as_float((int)(unsigned)(kk >> 32));
    }
    if (HERM) {   // the particle's force is complete: correct it right here
        hermite_correct_one(p.herm, i, tot, id);
    } else {
        const int io = p.iperm ? p.iperm[i] : i;
#pragma unroll
        for (int q = 0; q < 7; q++) p.out_sum[(size_t)io * 7 + q] = tot[q];
        p.out_key[io] = ko;
        if (NN) p.out_nnid[io] = id;
        for (int m = 0; m < p.n_mirror; m++) {
#pragma unroll
            for (int q = 0; q < 7; q++) p.m_sum[m][(size_t)io * 7 + q] = tot[q];
            p.m_key[m][io] = ko;
            p.m_id[m][io] = id;
        }
    }
}

// Called by ALL threads of a CTA after it has written final outputs.
__device__ __forceinline__ void signal_done(const ForceArgs &p)
{
    if (!p.host_flag) return;
    __threadfence_system();
    __syncthreads();
    if (p.done_expected == 1u) {   // this CTA wrote all outputs: no counting
        if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(p.host_flag) = p.flag_seq;
        return;
    }
    if (threadIdx.x == 0) {
        const unsigned int k = atomicAdd(p.done_counter, 1u);
        if (k == p.done_expected - 1u) {
            *p.done_counter = 0u;   // ready for the next launch (stream order)
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(p.host_flag) = p.flag_seq;
        }
    }
}

// ---- FP64 pairs -------------------------------------------------------------------------------
// One (i, j) interaction with the reference's FP64 expression tree (idata.cc:206-233) on the double-single
// predicted j (positions A+B, velocities C+L, mass A.w+L.w: 48 bits each) and the double-single i-particle.
// Returns acc, jerk, +m/r.
struct F64Out {
    double v[7];
};
__device__ __forceinline__ void fp64_accumulate(double *v, const double dx, const double dy, const double dz,
                                                const double dvx, const double dvy, const double dvz, const double m,
                                                const double eps2t)
{
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double xv = dx * dvx + dy * dvy + dz * dvz;
    const double ri = rsqrt(r2 + eps2t);      // eps2t = eps2 + 2^-52 (idata.cc:216)
    const double r2i = ri * ri;
    const double mri = m * ri;
    const double mr3i = mri * r2i;
    const double a3 = -3.0 * xv * r2i;
    v[0] += mr3i * dx;
    v[1] += mr3i * dy;
    v[2] += mr3i * dz;
    v[3] += mr3i * (dvx + a3 * dx);
    v[4] += mr3i * (dvy + a3 * dy);
    v[5] += mr3i * (dvz + a3 * dz);
    if (r2 > (double)TINYF) v[6] += mri;
}
__device__ __noinline__ void fp64_pair(F64Out *o, const float4 a, const float4 b, const float4 c, const float4 l,
                                       const double xi, const double yi, const double zi, const double vxi,
                                       const double vyi, const double vzi, const double eps2t)
{
#pragma unroll
    for (int q = 0; q < 7; q++) o->v[q] = 0.0;
    fp64_accumulate(o->v, ((double)a.x + (double)b.x) - xi, ((double)a.y + (double)b.y) - yi,
                    ((double)a.z + (double)b.z) - zi, ((double)c.x + (double)l.x) - vxi, ((double)c.y + (double)l.y) - vyi,
                    ((double)c.z + (double)l.z) - vzi, (double)a.w + (double)l.w, eps2t);
}

// FP64 pair (packed particle i, slot j of the launch's window) added into corr[i] with atomics.
__device__ __forceinline__ void close_pair_to_corr(const ForceArgs &p, const float4 ia, const float4 ib, const float4 ic,
                                                   const float4 id, const int i, const int j)
{
    const float4 a = p.jA[j], b = p.jB[j], c = p.jC[j], l = p.jL[j];
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    fp64_accumulate(v, ((double)a.x + (double)b.x) - ((double)ia.x + (double)ib.x),
                    ((double)a.y + (double)b.y) - ((double)ia.y + (double)ib.y),
                    ((double)a.z + (double)b.z) - ((double)ia.z + (double)ib.z),
                    ((double)c.x + (double)l.x) - ((double)ic.x + (double)id.x),
                    ((double)c.y + (double)l.y) - ((double)ic.y + (double)id.y),
                    ((double)c.z + (double)l.z) - ((double)ic.z + (double)id.z), (double)a.w + (double)l.w,
                    (double)p.eps2 + (double)TINYF);
#pragma unroll
    for (int q = 0; q < 7; q++) atomicAdd(p.corr + (size_t)i * 7 + q, v[q]);
}

struct Acc7 {
    float ax, ay, az, jx, jy, jz, pot;
};

// One (i, j) interaction in scalar FP32 with double-single positions.
// 9 FADD (DS dx) + 3 FADD (dv) + 6 FMUL/FFMA (r2, xv) + 1 FADD (eps2) + MUFU.RSQ
// + 5 FMUL + 9 FFMA + 1 FADD, plus guards.  Counted as 60 flop by convention
// (src/amuse_ph4/src/jdata.cc:1038).  Returns true if the pair is closer than the i-particle's FP64
// radius: it is then left out of the FP32 sums and the caller evaluates it with fp64_pair.
template <bool NN, bool LIST, bool NR>
__device__ __forceinline__ bool interact(const float4 a, const float4 b, const float4 c, int jaddr, float xh,
                                         float yh, float zh, float xl, float yl, float zl, float vx, float vy,
                                         float vz, int iid, float h2, float eps2, float thr, Acc7 &s, float &r2min,
                                         int &jmin, int i_global, const ForceArgs &p)
{
    float dx = (a.x - xh) + (b.x - xl);
    float dy = (a.y - yh) + (b.y - yl);
    float dz = (a.z - zh) + (b.z - zl);
    float dvx = c.x - vx, dvy = c.y - vy, dvz = c.z - vz;
    float r2 = dx * dx + dy * dy + dz * dz;
    float xv = dx * dvx + dy * dvy + dz * dvz;
    // idata.cc:216-233: r2i = 1/(r2 + eps2 + TINY) feeds acc and jerk unconditionally; pot and the
    // neighbour search are guarded by r2 > TINY.  Equal ids are skipped altogether (g6 rule).
    const bool idok = (__float_as_int(b.w) != iid);
    const bool ok = idok && (r2 > TINYF);
    const bool hp = ok && (r2 < thr);
    const bool use = idok && !hp;
    float rinv = NR ? rsqrt_refined(r2 + eps2) : rsqrt_approx(r2 + eps2);   // eps2 holds eps2 + TINY
    rinv = use ? rinv : 0.f;
    float rinv2 = rinv * rinv;
    float mrinv = a.w * rinv;
    float mr3 = mrinv * rinv2;
    float a3 = -3.f * (xv * rinv2);
    s.ax = fmaf(mr3, dx, s.ax);
    s.ay = fmaf(mr3, dy, s.ay);
    s.az = fmaf(mr3, dz, s.az);
    s.jx = fmaf(mr3, fmaf(a3, dx, dvx), s.jx);
    s.jy = fmaf(mr3, fmaf(a3, dy, dvy), s.jy);
    s.jz = fmaf(mr3, fmaf(a3, dz, dvz), s.jz);
    s.pot += (ok && !hp) ? mrinv : 0.f;
    if (NN) {
        float r2n = ok ? r2 : __int_as_float(0x7f800000);
        if (r2n < r2min) {
            r2min = r2n;
            jmin = jaddr;
        }
    }
    if (LIST) {   // sapporo's rule (dev_evaluate_gravity.cu:60-67): r2 <= h2 and ids differ -- no 2^-52 guard
        if (idok && r2 <= h2) {
            int pos = atomicAdd(&p.ngb_cnt[i_global], 1);
            if (pos < p.ngb_cap) p.ngb_list[(size_t)i_global * p.ngb_cap + pos] = __float_as_int(b.w);
        }
    }
    return hp;
}

// Two i-particles at once with Blackwell's packed FP32 pipe (FADD2/FMUL2/FFMA2):
// halves of every 64-bit register hold i0 and i1, the j operand is broadcast.
// nX* hold NEGATED i coordinates so that differences are single FADD2s.
struct IPair {
    u64 nxh, nyh, nzh, nxl, nyl, nzl, nvx, nvy, nvz;
    int id0, id1;
    float h20, h21;
};
struct Acc7P {
    u64 ax, ay, az, jx, jy, jz, pot;
};

// returns bit 0 / bit 1: pair (i0, j) / (i1, j) is inside the FP64 radius and was left out of the sums
template <bool NN, bool LIST, bool NR, bool TRACKJ = true>
__device__ __forceinline__ int interact2(const float4 a, const float4 b, const float4 c, int jaddr, const IPair &I,
                                         u64 eps2p, const float thr0, const float thr1, Acc7P &s, float &r2min0,
                                         int &jmin0, float &r2min1, int &jmin1, int i_global0, const ForceArgs &p)
{
    u64 dx = add2(add2(pk(a.x, a.x), I.nxh), add2(pk(b.x, b.x), I.nxl));
    u64 dy = add2(add2(pk(a.y, a.y), I.nyh), add2(pk(b.y, b.y), I.nyl));
    u64 dz = add2(add2(pk(a.z, a.z), I.nzh), add2(pk(b.z, b.z), I.nzl));
    u64 dvx = add2(pk(c.x, c.x), I.nvx);
    u64 dvy = add2(pk(c.y, c.y), I.nvy);
    u64 dvz = add2(pk(c.z, c.z), I.nvz);
    u64 r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
    u64 xv = fma2(dz, dvz, fma2(dy, dvy, mul2(dx, dvx)));
    u64 r2e = add2(r2, eps2p);
    float r20, r21, e0, e1;
    upk(r2, r20, r21);
    upk(r2e, e0, e1);
    int jid = __float_as_int(b.w);
    const bool id0 = (jid != I.id0), id1 = (jid != I.id1);
    const bool ok0 = id0 && (r20 > TINYF), ok1 = id1 && (r21 > TINYF);
    const bool hp0 = ok0 && (r20 < thr0), hp1 = ok1 && (r21 < thr1);
    float ri0 = rsqrt_approx(e0);
    float ri1 = rsqrt_approx(e1);
    u64 rinv = pk(ri0, ri1);
    if (NR) {  // packed Newton step: e = r2e*y0^2 - 1 ; y = y0 - (y0/2)*e   (4 packed ops)
        u64 e = fma2(r2e, mul2(rinv, rinv), pk(-1.f, -1.f));
        rinv = fma2(mul2(rinv, pk(-0.5f, -0.5f)), e, rinv);
        upk(rinv, ri0, ri1);
    }
    rinv = pk((id0 && !hp0) ? ri0 : 0.f, (id1 && !hp1) ? ri1 : 0.f);
    u64 rinv2 = mul2(rinv, rinv);
    u64 mrinv = mul2(pk(a.w, a.w), rinv);
    u64 mr3 = mul2(mrinv, rinv2);
    u64 a3 = mul2(mul2(xv, rinv2), pk(-3.f, -3.f));
    s.ax = fma2(mr3, dx, s.ax);
    s.ay = fma2(mr3, dy, s.ay);
    s.az = fma2(mr3, dz, s.az);
    s.jx = fma2(mr3, fma2(a3, dx, dvx), s.jx);
    s.jy = fma2(mr3, fma2(a3, dy, dvy), s.jy);
    s.jz = fma2(mr3, fma2(a3, dz, dvz), s.jz);
    s.pot = fma2(pk(a.w, a.w), pk((ok0 && !hp0) ? ri0 : 0.f, (ok1 && !hp1) ? ri1 : 0.f), s.pot);
    if (NN) {
        float n0 = ok0 ? r20 : __int_as_float(0x7f800000);
        float n1 = ok1 ? r21 : __int_as_float(0x7f800000);
        if (!TRACKJ) {
            r2min0 = fminf(r2min0, n0);
            r2min1 = fminf(r2min1, n1);
        } else {
            if (n0 < r2min0) {
                r2min0 = n0;
                jmin0 = jaddr;
            }
            if (n1 < r2min1) {
                r2min1 = n1;
                jmin1 = jaddr;
            }
        }
    }
    if (LIST) {   // sapporo's rule (dev_evaluate_gravity.cu:60-67): r2 <= h2 and ids differ -- no 2^-52 guard
        if (id0 && r20 <= I.h20) {
            int pos = atomicAdd(&p.ngb_cnt[i_global0], 1);
            if (pos < p.ngb_cap) p.ngb_list[(size_t)i_global0 * p.ngb_cap + pos] = jid;
        }
        if (id1 && r21 <= I.h21) {
            int pos = atomicAdd(&p.ngb_cnt[i_global0 + 1], 1);
            if (pos < p.ngb_cap) p.ngb_list[(size_t)(i_global0 + 1) * p.ngb_cap + pos] = jid;
        }
    }
    return (hp0 ? 1 : 0) | (hp1 ? 2 : 0);
}

// Where particle i sits among the j and how large its FP64 radius is (masked kernels, per i-particle):
// the j-particle with its id if there is exactly one (the usual case: the active particles of a block
// step ARE j-particles), else the slot its Morton key falls on.  d = distance to that slot + that
// particle's own nearest-neighbour distance bounds i's nearest-neighbour distance from above.
__device__ __forceinline__ float close_radius2(const ForceArgs &p, const int iid, const float xh, const float yh,
                                               const float zh, const float xl, const float yl, const float zl,
                                               int &self_slot)
{
    self_slot = hash_lookup(p.ord, iid);
    if (!(p.ord.kclose > 0.f)) return 0.f;
    int s = self_slot;
    if (s < 0) {
        if (p.ord.nkeys <= 0) return 0.f;
        s = key_search(p.ord, morton30(xh, yh, zh, p.ord.blo, p.ord.binv));
    }
    const float4 a = p.js.A[s], b = p.js.B[s];
    const float dx = (a.x - xh) + (b.x - xl), dy = (a.y - yh) + (b.y - yl), dz = (a.z - zh) + (b.z - zl);
    const float d = sqrtf(dx * dx + dy * dy + dz * dz) + sqrtf(p.js.near2[s]);
    return fminf(fminf(p.ord.kclose * d * d, p.ord.cap2), p.js.capr2[s]);
}

// Split reduction shared by the force kernels: the last CTA of an i-block (ticket) sums the
// per-split partials and writes the outputs.  Big i-blocks (IB >= THREADS): one thread per i, splits
// added in order.  Small i-blocks: TPI = min(32, THREADS/IB) lanes share an i, each sums a strided
// subset of the splits, and a fixed butterfly of shuffles combines them -- the block-timestep regime
// has hundreds of splits of a handful of i, which one thread per i would walk serially.  Both orders
// are fixed, so results are deterministic.
template <bool NN, bool HERM = false>
__device__ __forceinline__ void reduce_splits(const ForceArgs &p, unsigned int *is_last, const int IB)
{
    const int tid = threadIdx.x;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int tk = atomicAdd(&p.tickets[blockIdx.y], 1u);
        *is_last = (tk == (unsigned)p.nsplit - 1u) ? 1u : 0u;
        if (*is_last) p.tickets[blockIdx.y] = 0u;  // ready for the next launch
    }
    __syncthreads();
    if (!*is_last) return;
    __threadfence();
    auto write_out = [&](int i, const double *tot, u64 kk) { store_outputs<NN, HERM>(p, i, tot, kk); };
    if (IB >= THREADS) {
        for (int il = tid; il < IB; il += THREADS) {
            int i = blockIdx.y * IB + il;
            if (i >= p.ni) continue;
            double tot[7] = {0, 0, 0, 0, 0, 0, 0};
            u64 kk = KEY_NONE;
            // the loads of eight splits are issued together: this loop is bound by L2 latency
            int sp = 0;
            for (; sp + 8 <= p.nsplit; sp += 8) {
                double v[8][7];
                u64 kv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    size_t o = (size_t)(sp + u) * p.ni_pad + i;
#pragma unroll
                    for (int q = 0; q < 7; q++) v[u][q] = __ldcg(p.part_sum + o * 7 + q);
                    kv[u] = __ldcg(p.part_key + o);
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
#pragma unroll
                    for (int q = 0; q < 7; q++) tot[q] += v[u][q];
                    kk = kv[u] < kk ? kv[u] : kk;
                }
            }
            for (; sp < p.nsplit; sp++) {
                size_t o = (size_t)sp * p.ni_pad + i;
                const double *r = p.part_sum + o * 7;
#pragma unroll
                for (int q = 0; q < 7; q++) tot[q] += __ldcg(r + q);
                u64 ok = __ldcg(p.part_key + o);
                kk = ok < kk ? ok : kk;
            }
            write_out(i, tot, kk);
        }
    } else {
        const int TPI = (THREADS / IB) < 32 ? (THREADS / IB) : 32;   // lanes per i (power of two)
        const int per_pass = THREADS / TPI;
        const int sub = tid % TPI;
        for (int base = 0; base < IB; base += per_pass) {
            const int il = base + tid / TPI;
            const int i = blockIdx.y * IB + il;
            const bool valid = (il < IB) && (i < p.ni);
            double tot[7] = {0, 0, 0, 0, 0, 0, 0};
            u64 kk = KEY_NONE;
            if (valid) {
                int sp = sub;
                for (; sp + 3 * TPI < p.nsplit; sp += 4 * TPI) {
                    double v[4][7];
                    u64 kv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        size_t o = (size_t)(sp + u * TPI) * p.ni_pad + i;
#pragma unroll
                        for (int q = 0; q < 7; q++) v[u][q] = __ldcg(p.part_sum + o * 7 + q);
                        kv[u] = __ldcg(p.part_key + o);
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
#pragma unroll
                        for (int q = 0; q < 7; q++) tot[q] += v[u][q];
                        kk = kv[u] < kk ? kv[u] : kk;
                    }
                }
                for (; sp < p.nsplit; sp += TPI) {
                    size_t o = (size_t)sp * p.ni_pad + i;
#pragma unroll
                    for (int q = 0; q < 7; q++) tot[q] += __ldcg(p.part_sum + o * 7 + q);
                    u64 ok = __ldcg(p.part_key + o);
                    kk = ok < kk ? ok : kk;
                }
            }
            for (int off = 1; off < TPI; off <<= 1) {
#pragma unroll
                for (int q = 0; q < 7; q++) tot[q] += __shfl_xor_sync(0xffffffffu, tot[q], off);
                u64 o = __shfl_xor_sync(0xffffffffu, kk, off);
                kk = o < kk ? o : kk;
            }
            if (valid && sub == 0) write_out(i, tot, kk);
        }
    }
    signal_done(p);
}

struct __align__(16) ForceSmem {
    float4 A[STAGES][TILE];
    float4 B[STAGES][TILE];
    float4 C[STAGES][TILE];
    float4 G[STAGES][TBOX];   // group boxes + tile box (speculative kernel only)
    uint64_t full[STAGES];
    unsigned int is_last;
    unsigned int wl_n;        // masked kernels: FP64 pairs found by this CTA
    int2 wl[CTA_WL];          // (packed i, slot)
};

// force_fast_kernel: a deeper pipeline and no CTA-wide barrier per tile.  The warps of a CTA take different
// times over a tile (a warp whose particles have FP64 pairs in it runs the masked loop, the others the
// mask-free one), so each warp releases a stage on its own (done[s]) and the last one to do so refills it;
// a warp may run up to FSTAGES - 1 tiles ahead of the slowest.
constexpr int FSTAGES = 8;
struct __align__(16) FastSmem {
    float4 A[FSTAGES][TILE];
    float4 B[FSTAGES][TILE];
    float4 C[FSTAGES][TILE];
    float4 G[FSTAGES][TBOX];
    uint64_t full[FSTAGES];
    unsigned int done[FSTAGES];   // warps that have finished with the stage's current tile
    unsigned int is_last;
};

__device__ __forceinline__ u64 make_key(float r2min, int jmin)
{
    return ((u64)(unsigned)__float_as_int(r2min) << 32) | (u64)(unsigned)jmin;
}

// Thread layout: tid = jslot * NI_SLOTS + islot.  A CTA owns IB = NI_SLOTS*IPT
// i-particles (i = blockIdx.y*IB + islot + k*NI_SLOTS, coalesced) and the j-tiles
// of split blockIdx.x; inside a tile the NJ_SLOTS = THREADS/NI_SLOTS j-slots take
// interleaved j.  i-particles stay in registers for the whole kernel; j tiles
// arrive by TMA bulk copies (3 per stage) signalled on an mbarrier; per-tile FP32
// partial sums are flushed to FP64 so that long sums keep ~1e-7 accuracy; pairs inside the
// i-particle's FP64 radius bypass the FP32 sums altogether (fp64_pair).
// Partials of the j-slots are reduced with warp shuffles + shared memory, and the
// partials of the j-splits by the last CTA to arrive (ticket), in fixed order.
template <int IPT, int NI_SLOTS, bool NN, bool LIST, bool PACKED, bool NR, int MINB, int INL, bool HERM = false>
__global__ void __launch_bounds__(THREADS, MINB) force_kernel(const __grid_constant__ ForceArgs p,
                                                              const __grid_constant__ InlineI<INL> ii)
{
    constexpr int NJ_SLOTS = THREADS / NI_SLOTS;
    constexpr int IB = NI_SLOTS * IPT;
    constexpr int LANES_PER_I = (NI_SLOTS >= 32) ? 1 : 32 / NI_SLOTS;   // lanes of a warp sharing an i
    constexpr int JG = (NI_SLOTS >= 32) ? NJ_SLOTS : THREADS / 32;      // j-groups left after the shuffle stage
    static_assert(!PACKED || (IPT % 2 == 0), "packed path handles i in pairs");
    static_assert(JG * IB * 64 <= (int)sizeof(float4) * 3 * STAGES * TILE || JG == 1, "reduction scratch fits tile smem");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    ForceSmem &sm = *reinterpret_cast<ForceSmem *>(smem_raw);

    const int tid = threadIdx.x;
    const int islot = tid % NI_SLOTS;
    const int jslot = tid / NI_SLOTS;
    const int i_base = blockIdx.y * IB + islot;

    // ---- j tiles of this split -------------------------------------------
    const int ntiles_total = (p.nj + TILE - 1) / TILE;
    // split s takes tiles s, s + nsplit, s + 2 nsplit, ...: the tiles that hold an i-block's own neighbourhood
    // (FP64 pairs, masked blocks) are dealt out over all splits instead of landing on one CTA
    const int tstride = p.nsplit;
    int ntiles = ((int)blockIdx.x < ntiles_total) ? (ntiles_total - (int)blockIdx.x + tstride - 1) / tstride : 0;
    auto tile_of = [&](const int t) -> int { return (int)blockIdx.x + t * tstride; };

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.wl_n = 0u;
    }
    __syncthreads();
    constexpr uint32_t STAGE_BYTES = 3u * TILE * sizeof(float4);
    if (tid == 0) {
        for (int s = 0; s < STAGES && s < ntiles; s++) {
            size_t off = (size_t)tile_of(s) * TILE;
            mbar_expect_tx(&sm.full[s], STAGE_BYTES);
            bulk_g2s(sm.A[s], p.jA + off, TILE * sizeof(float4), &sm.full[s]);
            bulk_g2s(sm.B[s], p.jB + off, TILE * sizeof(float4), &sm.full[s]);
            bulk_g2s(sm.C[s], p.jC + off, TILE * sizeof(float4), &sm.full[s]);
        }
    }

    // ---- register-resident i-particles -----------------------------------
    auto i_of = [&](int k) -> int {
        return PACKED ? (int)(blockIdx.y * IB + (islot * 2 + (k & 1)) + (k >> 1) * (2 * NI_SLOTS)) : i_base + k * NI_SLOTS;
    };
    float xh[IPT], yh[IPT], zh[IPT], xl[IPT], yl[IPT], zl[IPT], vx[IPT], vy[IPT], vz[IPT], h2[IPT], thr[IPT];
    int iid[IPT];
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        const int i = i_of(k);
        float4 a = make_float4(0.f, 0.f, 0.f, -1.f), b = make_float4(0.f, 0.f, 0.f, __int_as_float(0x80000000)),
               c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < p.ni) {
            if (INL > 0) {
                a = ii.d[i];
                b = ii.d[INL + i];
                c = ii.d[2 * INL + i];
            } else {
                a = p.iA[i];
                b = p.iB[i];
                c = p.iC[i];
            }
        }
        xh[k] = a.x; yh[k] = a.y; zh[k] = a.z; h2[k] = a.w;
        xl[k] = b.x; yl[k] = b.y; zl[k] = b.z; iid[k] = __float_as_int(b.w);
        vx[k] = c.x; vy[k] = c.y; vz[k] = c.z;
        thr[k] = 0.f;
        if (i < p.ni) {
            int self_slot;
            thr[k] = close_radius2(p, iid[k], xh[k], yh[k], zh[k], xl[k], yl[k], zl[k], self_slot);
            if (p.conf && blockIdx.x == 0 && jslot == 0) p.conf[i] = self_slot;
        }
    }
    const double eps2t = (double)p.eps2 + (double)TINYF;
    // Pairs (k, slot j of this launch's window) to be evaluated in FP64 are collected as bits per particle over a
    // flush group (no branch in the pair loop: with ~1.5 % of the pairs inside an FP64 radius nearly every
    // warp-iteration would diverge), then queued with one shared-memory atomic per particle and group for the
    // dense pass after the tile loop; without a corr buffer (or with the queue full) evaluated here into D[k].
    // bit u of bits: the pair with j = jbase + u * NJ_SLOTS.
    auto flush_close = [&](const int k, unsigned int bits, const int jbase, double *Dk) {
        const int i = i_of(k);
        unsigned int e = p.corr ? atomicAdd(&sm.wl_n, (unsigned)__popc(bits)) : (unsigned)CTA_WL;
        while (bits) {
            const int u = __ffs(bits) - 1;
            bits &= bits - 1u;
            const int jlocal = jbase + u * NJ_SLOTS;
            if (e < (unsigned)CTA_WL) {
                sm.wl[e++] = make_int2(i, jlocal);
                continue;
            }
            const float4 d = (INL > 0) ? ii.d[3 * INL + i] : p.iD[i];
            F64Out o;
            fp64_pair(&o, p.jA[jlocal], p.jB[jlocal], p.jC[jlocal], p.jL[jlocal], (double)xh[k] + (double)xl[k],
                      (double)yh[k] + (double)yl[k], (double)zh[k] + (double)zl[k], (double)vx[k] + (double)d.x,
                      (double)vy[k] + (double)d.y, (double)vz[k] + (double)d.z, eps2t);
#pragma unroll
            for (int q = 0; q < 7; q++) Dk[q] += o.v[q];
        }
    };

    double D[IPT][7];
    float r2min[IPT];
    int jmin[IPT];
#pragma unroll
    for (int k = 0; k < IPT; k++) {
#pragma unroll
        for (int q = 0; q < 7; q++) D[k][q] = 0.0;
        r2min[k] = __int_as_float(0x7f800000);
        jmin[k] = -1;
    }

    constexpr int NP = PACKED ? IPT / 2 : 1;
    IPair IP[NP];
    if (PACKED) {
#pragma unroll
        for (int q = 0; q < NP; q++) {
            IP[q].nxh = pk(-xh[2 * q], -xh[2 * q + 1]);
            IP[q].nyh = pk(-yh[2 * q], -yh[2 * q + 1]);
            IP[q].nzh = pk(-zh[2 * q], -zh[2 * q + 1]);
            IP[q].nxl = pk(-xl[2 * q], -xl[2 * q + 1]);
            IP[q].nyl = pk(-yl[2 * q], -yl[2 * q + 1]);
            IP[q].nzl = pk(-zl[2 * q], -zl[2 * q + 1]);
            IP[q].nvx = pk(-vx[2 * q], -vx[2 * q + 1]);
            IP[q].nvy = pk(-vy[2 * q], -vy[2 * q + 1]);
            IP[q].nvz = pk(-vz[2 * q], -vz[2 * q + 1]);
            IP[q].id0 = iid[2 * q];
            IP[q].id1 = iid[2 * q + 1];
            IP[q].h20 = h2[2 * q];
            IP[q].h21 = h2[2 * q + 1];
        }
    }
    const float eps2 = p.eps2 + TINYF;   // the reference softens by eps2 + 2^-52 (idata.cc:216)
    const u64 eps2p = pk(eps2, eps2);

    // ---- main loop over tiles --------------------------------------------
    for (int t = 0; t < ntiles; t++) {
        const int s = t % STAGES;
        const uint32_t phase = (uint32_t)(t / STAGES) & 1u;
        while (!mbar_try_wait(&sm.full[s], phase)) {
        }
        const int jtile = tile_of(t) * TILE;
        int cnt = p.nj - jtile;
        if (cnt > TILE) cnt = TILE;
        const float4 *tA = sm.A[s], *tB = sm.B[s], *tC = sm.C[s];

        // The thread's j of this tile are jj = jslot + u*NJ_SLOTS, u = 0..ITERS-1.  They are processed
        // in groups of FL: FP32 partial sums over at most FL pairs, then flushed to the FP64 totals
        // (F2F + DADD run on the XU / FP64 pipes, off the FP32 pipe that bounds the kernel), so the
        // summation error stays below the per-pair rounding error for any N.
        constexpr int ITERS = TILE / NJ_SLOTS;
        constexpr int FL = ITERS < FLUSH ? ITERS : FLUSH;
        for (int u0 = 0; u0 < ITERS; u0 += FL) {
            const int jj0 = jslot + u0 * NJ_SLOTS;
            if (jj0 >= cnt) break;
            const bool whole = (jj0 + (FL - 1) * NJ_SLOTS) < cnt;
            if (!PACKED) {
                Acc7 S[IPT];
#pragma unroll
                for (int k = 0; k < IPT; k++) S[k] = Acc7{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                unsigned int hb[IPT];
#pragma unroll
                for (int k = 0; k < IPT; k++) hb[k] = 0u;
                auto do_j = [&](const int jj, const int u) {
                    const float4 a = tA[jj], b = tB[jj], c = tC[jj];
                    const int jaddr = jtile + jj;
#pragma unroll
                    for (int k = 0; k < IPT; k++) {
                        const bool hp = interact<NN, LIST, NR>(a, b, c, jaddr, xh[k], yh[k], zh[k], xl[k], yl[k], zl[k],
                                                               vx[k], vy[k], vz[k], iid[k], h2[k], eps2, thr[k], S[k],
                                                               r2min[k], jmin[k], i_of(k), p);
                        hb[k] |= (hp ? 1u : 0u) << u;
                    }
                };
                if (whole) {
#pragma unroll UNROLL
                    for (int u = 0; u < FL; u++) do_j(jj0 + u * NJ_SLOTS, u);
                } else {
                    int u = 0;
                    for (int jj = jj0; jj < cnt; jj += NJ_SLOTS) do_j(jj, u++);
                }
#pragma unroll
                for (int k = 0; k < IPT; k++)
                    if (hb[k]) flush_close(k, hb[k], jtile + jj0, D[k]);
#pragma unroll
                for (int k = 0; k < IPT; k++) {
                    D[k][0] += (double)S[k].ax; D[k][1] += (double)S[k].ay; D[k][2] += (double)S[k].az;
                    D[k][3] += (double)S[k].jx; D[k][4] += (double)S[k].jy; D[k][5] += (double)S[k].jz;
                    D[k][6] += (double)S[k].pot;
                }
            } else {
                Acc7P S[NP];
#pragma unroll
                for (int q = 0; q < NP; q++) S[q] = Acc7P{0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
                unsigned int hb[IPT];
#pragma unroll
                for (int k = 0; k < IPT; k++) hb[k] = 0u;
                auto do_j = [&](const int jj, const int u) {
                    const float4 a = tA[jj], b = tB[jj], c = tC[jj];
                    const int jaddr = jtile + jj;
#pragma unroll
                    for (int q = 0; q < NP; q++) {
                        const int hpm = interact2<NN, LIST, NR>(a, b, c, jaddr, IP[q], eps2p, thr[2 * q], thr[2 * q + 1],
                                                                S[q], r2min[2 * q], jmin[2 * q], r2min[2 * q + 1],
                                                                jmin[2 * q + 1], i_of(2 * q), p);
                        hb[2 * q] |= (unsigned)(hpm & 1) << u;
                        hb[2 * q + 1] |= (unsigned)((hpm >> 1) & 1) << u;
                    }
                };
                if (whole) {
#pragma unroll UNROLL
                    for (int u = 0; u < FL; u++) do_j(jj0 + u * NJ_SLOTS, u);
                } else {
                    int u = 0;
                    for (int jj = jj0; jj < cnt; jj += NJ_SLOTS) do_j(jj, u++);
                }
#pragma unroll
                for (int k = 0; k < IPT; k++)
                    if (hb[k]) flush_close(k, hb[k], jtile + jj0, D[k]);
#pragma unroll
                for (int q = 0; q < NP; q++) {
                    float lo, hi;
                    upk(S[q].ax, lo, hi); D[2 * q][0] += (double)lo; D[2 * q + 1][0] += (double)hi;
                    upk(S[q].ay, lo, hi); D[2 * q][1] += (double)lo; D[2 * q + 1][1] += (double)hi;
                    upk(S[q].az, lo, hi); D[2 * q][2] += (double)lo; D[2 * q + 1][2] += (double)hi;
                    upk(S[q].jx, lo, hi); D[2 * q][3] += (double)lo; D[2 * q + 1][3] += (double)hi;
                    upk(S[q].jy, lo, hi); D[2 * q][4] += (double)lo; D[2 * q + 1][4] += (double)hi;
                    upk(S[q].jz, lo, hi); D[2 * q][5] += (double)lo; D[2 * q + 1][5] += (double)hi;
                    upk(S[q].pot, lo, hi); D[2 * q][6] += (double)lo; D[2 * q + 1][6] += (double)hi;
                }
            }
        }

        __syncthreads();  // everyone is done with stage s
        if (tid == 0 && t + STAGES < ntiles) {
            size_t off = (size_t)tile_of(t + STAGES) * TILE;
            mbar_expect_tx(&sm.full[s], STAGE_BYTES);
            bulk_g2s(sm.A[s], p.jA + off, TILE * sizeof(float4), &sm.full[s]);
            bulk_g2s(sm.B[s], p.jB + off, TILE * sizeof(float4), &sm.full[s]);
            bulk_g2s(sm.C[s], p.jC + off, TILE * sizeof(float4), &sm.full[s]);
        }
    }

    // ---- dense pass over the FP64 pairs this CTA queued: one pair per thread ----------------------
    if (p.corr) {
        __syncthreads();
        const unsigned int nwl = sm.wl_n < (unsigned)CTA_WL ? sm.wl_n : (unsigned)CTA_WL;
        const unsigned int nround = (nwl + 31u) & ~31u;   // whole warps stay in the loop (shuffles below)
        for (unsigned int e = tid; e < nround; e += THREADS) {
            int wi = -1;
            double v[7] = {0, 0, 0, 0, 0, 0, 0};
            if (e < nwl) {
                const int2 w = sm.wl[e];
                wi = w.x;
                const float4 ia = (INL > 0) ? ii.d[w.x] : p.iA[w.x], ib = (INL > 0) ? ii.d[INL + w.x] : p.iB[w.x],
                             ic = (INL > 0) ? ii.d[2 * INL + w.x] : p.iC[w.x], id = (INL > 0) ? ii.d[3 * INL + w.x] : p.iD[w.x];
                const float4 a = p.jA[w.y], b = p.jB[w.y], c = p.jC[w.y], l = p.jL[w.y];
                fp64_accumulate(v, ((double)a.x + (double)b.x) - ((double)ia.x + (double)ib.x),
                                ((double)a.y + (double)b.y) - ((double)ia.y + (double)ib.y),
                                ((double)a.z + (double)b.z) - ((double)ia.z + (double)ib.z),
                                ((double)c.x + (double)l.x) - ((double)ic.x + (double)id.x),
                                ((double)c.y + (double)l.y) - ((double)ic.y + (double)id.y),
                                ((double)c.z + (double)l.z) - ((double)ic.z + (double)id.z), (double)a.w + (double)l.w, eps2t);
            }
            // a CTA of a small block holds a handful of particles: sum the lanes of each particle with shuffles and
            // let one lane add the seven totals (hundreds of pairs per particle would otherwise queue up on seven
            // addresses); with many particles per warp the lanes add their own
            unsigned int todo = __ballot_sync(0xffffffffu, wi >= 0);
            for (int round = 0; round < G6_DENSE_ROUNDS && todo; round++) {
                const int cur = __shfl_sync(0xffffffffu, wi, __ffs(todo) - 1);
                const bool mine = (wi == cur);
                const unsigned int grp = __ballot_sync(0xffffffffu, mine);
#pragma unroll
                for (int q = 0; q < 7; q++) {
                    double x = mine ? v[q] : 0.0;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
                    if ((int)(tid & 31) == __ffs(grp) - 1) atomicAdd(p.corr + (size_t)cur * 7 + q, x);
                }
                if (mine) wi = -1;
                todo &= ~grp;
            }
            if (wi >= 0) {
#pragma unroll
                for (int q = 0; q < 7; q++) atomicAdd(p.corr + (size_t)wi * 7 + q, v[q]);
            }
        }
        __threadfence();   // the atomics precede this CTA's ticket / final stores
    }

    // ---- keys (slot of this launch's j-window) --------------------------------
    u64 key[IPT];
#pragma unroll
    for (int k = 0; k < IPT; k++)
        key[k] = (jmin[k] >= 0 && r2min[k] < 1.0e30f) ? make_key(r2min[k], jmin[k]) : KEY_NONE;   // >= 1e30: parked slots

    // ---- reduce over the j-slots of this CTA -------------------------------
    double *red = reinterpret_cast<double *>(smem_raw);  // tile buffers are dead now
    if (NJ_SLOTS > 1) {
        if (LANES_PER_I > 1) {
#pragma unroll
            for (int off = NI_SLOTS; off < 32; off <<= 1) {
#pragma unroll
                for (int k = 0; k < IPT; k++) {
#pragma unroll
                    for (int q = 0; q < 7; q++) D[k][q] += __shfl_xor_sync(0xffffffffu, D[k][q], off);
                    u64 o = __shfl_xor_sync(0xffffffffu, key[k], off);
                    key[k] = o < key[k] ? o : key[k];
                }
            }
        }
        __syncthreads();  // all warps past their last tile reads (also covers ntiles == 0)
        const int g = (NI_SLOTS >= 32) ? jslot : (tid >> 5);
        const bool writer = (NI_SLOTS >= 32) ? true : ((tid & 31) < NI_SLOTS);
        if (writer) {
#pragma unroll
            for (int k = 0; k < IPT; k++) {
                int il = PACKED ? (islot * 2 + (k & 1)) + (k >> 1) * (2 * NI_SLOTS) : islot + k * NI_SLOTS;
                double *r = red + ((size_t)g * IB + il) * 8;
#pragma unroll
                for (int q = 0; q < 7; q++) r[q] = D[k][q];
                reinterpret_cast<u64 *>(r)[7] = key[k];
            }
        }
        __syncthreads();
    }

    // ---- CTA totals -> global (final or per-split partial) ------------------
    const bool single = (p.nsplit == 1);
    auto emit = [&](int il, const double *tot, u64 kk) {
        int i = blockIdx.y * IB + il;
        if (i >= p.ni) return;
        if (single) {
            store_outputs<NN, HERM>(p, i, tot, kk);
        } else {
            size_t o = (size_t)blockIdx.x * p.ni_pad + i;
#pragma unroll
            for (int q = 0; q < 7; q++) p.part_sum[o * 7 + q] = tot[q];
            p.part_key[o] = kk;
        }
    };
    if (NJ_SLOTS > 1) {
        for (int il = tid; il < IB; il += THREADS) {
            double tot[7] = {0, 0, 0, 0, 0, 0, 0};
            u64 kk = KEY_NONE;
            for (int g = 0; g < JG; g++) {
                const double *r = red + ((size_t)g * IB + il) * 8;
#pragma unroll
                for (int q = 0; q < 7; q++) tot[q] += r[q];
                u64 o = reinterpret_cast<const u64 *>(r)[7];
                kk = o < kk ? o : kk;
            }
            emit(il, tot, kk);
        }
    } else {
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            int il = PACKED ? (islot * 2 + (k & 1)) + (k >> 1) * (2 * NI_SLOTS) : islot + k * NI_SLOTS;
            emit(il, D[k], key[k]);
        }
    }
    if (single) signal_done(p);
    if (single || p.defer_reduce) return;
    reduce_splits<NN, HERM>(p, &sm.is_last, IB);
}

// ---------------------------------------------------------------------------
// Speculative force kernel (large i-blocks, packed FP32), round 2.
//
// i-particles arrive Morton-sorted (64 consecutive ones per warp, two per thread), the j-memory is
// Morton-ordered, and predict_kernel has left a bounding box per group of 32 j and per tile.  Every
// (warp, group) block is classified by box distance, warp-uniformly:
//   FAR    gap^2 > max(farc2 * (largest |coordinate|)^2, every i's nearest-neighbour bound, every i's FP64
//          radius): position differences from the hi parts alone (31 packed FP32 operations per pair
//          instead of 37; the lo parts are below the rounding of such a difference), no neighbour search;
//   NEAR   mask-free pair function on the double-single differences, running minimum of r2 for the
//          neighbour search; a group whose minimum is not above R2_EXACT is redone exactly;
//   CLOSE  some particle of the warp has the group's box inside its FP64 radius, or the j-particle that
//          carries the id of one of the warp's particles sits in the group (conf[], from the id table), or
//          the group is the ragged tail: the block is evaluated pair by pair with the reference's rule
//          (idata.cc:206-233: equal ids skipped, pot and neighbour search only for r2 > 2^-52), and every pair
//          closer than its i-particle's FP64 radius is left out of the FP32 sums and queued for
//          close_pairs_kernel, which evaluates the queue densely in FP64 (one pair per thread).
// Results are therefore those of the masked rule for every input, and the pairs that dominate a
// particle's acc and jerk never see FP32 rounding.  The nearest neighbour is kept as (min r2, first group
// that lowered it); the exact j is found at the end by re-scanning that one group with the same r2
// instruction sequence.
// ---------------------------------------------------------------------------
#ifndef G6_GRP
#define G6_GRP 32
#endif
#ifndef G6_FUNROLL
#define G6_FUNROLL 16   // the whole group of 32 j unrolled (measured: 67.5 / 70.7 / 69.6 / 71.8 % for 2 / 4 / 8 / 16)
#endif
#ifndef G6_FAR_RAW
#define G6_FAR_RAW 1    // FAR blocks take MUFU.RSQ without the Newton step (28 instead of 31 operations per pair)
#endif
constexpr int GRP = G6_GRP;   // pairs per group (FP32 partial sums span one group); one lane per j in the FP64 path
constexpr int FUNROLL = G6_FUNROLL;  // j-pairs unrolled in the mask-free loop
static_assert(GRP == 32, "the FP64 path predicts one j per lane");

__device__ __forceinline__ float min3f(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// Geometry shared by the fast path, the FP64 path's neighbour search and the neighbour re-scan: identical
// instruction sequence, so r2 is bit-identical in all three.
__device__ __forceinline__ void pair_geometry(const float4 a, const float4 b, const IPair &I, u64 &dx, u64 &dy, u64 &dz,
                                              u64 &r2)
{
    dx = add2(add2(pk(a.x, a.x), I.nxh), add2(pk(b.x, b.x), I.nxl));
    dy = add2(add2(pk(a.y, a.y), I.nyh), add2(pk(b.y, b.y), I.nyl));
    dz = add2(add2(pk(a.z, a.z), I.nzh), add2(pk(b.z, b.z), I.nzl));
    r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
}

// One j against an i-pair, no masks: 37 (FAR: 31) packed FP32 operations + 2 MUFU.  Accumulates -acc, -jerk, +pot.
// EPS0: the caller runs unsoftened (eps2 = 0, ph4's AMUSE default).  The reference still adds 2^-52 to r2
// (idata.cc:216); in FP32 that add is the identity for r2 > 2^-26, so it is skipped here and the group
// verification (running minimum of r2) redoes any group that holds a closer pair with the exact path.
template <bool NR, bool EPS0, bool FAR>
__device__ __forceinline__ u64 interact2_fast(const float4 a, const float4 b, const float4 c, const IPair &I,
                                              const u64 eps2p, Acc7P &s)
{
    u64 dx, dy, dz, r2;
    if (FAR) {
        dx = add2(pk(a.x, a.x), I.nxh);
        dy = add2(pk(a.y, a.y), I.nyh);
        dz = add2(pk(a.z, a.z), I.nzh);
        r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
    } else {
        pair_geometry(a, b, I, dx, dy, dz, r2);
    }
    const u64 dvx = add2(pk(c.x, c.x), I.nvx);
    const u64 dvy = add2(pk(c.y, c.y), I.nvy);
    const u64 dvz = add2(pk(c.z, c.z), I.nvz);
    const u64 xv = fma2(dz, dvz, fma2(dy, dvy, mul2(dx, dvx)));
    const u64 r2e = EPS0 ? r2 : add2(r2, eps2p);
    float e0, e1;
    upk(r2e, e0, e1);
    const u64 y0 = pk(rsqrt_approx(e0), rsqrt_approx(e1));
    // rinv2 and mr3 carry a MINUS sign here (the caller subtracts the acc/jerk partial sums): with
    // x = r2 + eps2 and one Newton step from y0 = MUFU.RSQ(x),
    //   e2 = x*y0^2 - 2 ;  -1/r^2 = y0^2 * e2  (Newton for 1/x from y0^2) ;  1/r = y0 * (0.5 - e2/2)
    // -- no instruction of the step reads three distinct register pairs.
    u64 rinv2, mrinv;
    if (NR) {
        const u64 yy = mul2(y0, y0);
        const u64 e2 = fma2(r2e, yy, pk(-2.f, -2.f));
        rinv2 = mul2(yy, e2);
        mrinv = mul2(mul2(pk(a.w, a.w), y0), fma2(e2, pk(-0.5f, -0.5f), pk(0.5f, 0.5f)));
    } else {
        rinv2 = mul2(y0, mul2(y0, pk(-1.f, -1.f)));
        mrinv = mul2(pk(a.w, a.w), y0);
    }
    const u64 mr3 = mul2(mrinv, rinv2);                          // = -m/r^3
    const u64 a3 = mul2(mul2(xv, rinv2), pk(3.f, 3.f));          // = -3 x.v/r^2
    const u64 tx = fma2(a3, dx, dvx);
    const u64 ty = fma2(a3, dy, dvy);
    const u64 tz = fma2(a3, dz, dvz);
    s.ax = fma2(mr3, dx, s.ax);
    s.ay = fma2(mr3, dy, s.ay);
    s.az = fma2(mr3, dz, s.az);
    s.jx = fma2(mr3, tx, s.jx);
    s.jy = fma2(mr3, ty, s.jy);
    s.jz = fma2(mr3, tz, s.jz);
    s.pot = add2(s.pot, mrinv);
    return r2;
}

// FP64 pairs the speculative kernel queued, one per thread (grid-stride over the list).  A lane queued all pairs
// of one particle back to back, so a warp mostly holds runs of equal i: each run is summed with shuffles and
// its head lane issues the seven atomics.
__global__ void __launch_bounds__(256) close_pairs_kernel(const ForceArgs p)
{
    const unsigned int n = min(*p.wl_count, p.wl_cap);
    const unsigned int nround = (n + 31u) & ~31u;   // whole warps stay in the loop
    const int lane = threadIdx.x & 31;
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < nround; e += gridDim.x * blockDim.x) {
        int i = -1 - lane;   // lanes past the end: distinct negative ids, no contribution
        double v[7] = {0, 0, 0, 0, 0, 0, 0};
        if (e < n) {
            const int2 w = p.wl[e];
            i = w.x;
            const float4 ia = p.iA[i], ib = p.iB[i], ic = p.iC[i], id = p.iD[i];
            const float4 a = p.jA[w.y], b = p.jB[w.y], c = p.jC[w.y], l = p.jL[w.y];
            fp64_accumulate(v, ((double)a.x + (double)b.x) - ((double)ia.x + (double)ib.x),
                            ((double)a.y + (double)b.y) - ((double)ia.y + (double)ib.y),
                            ((double)a.z + (double)b.z) - ((double)ia.z + (double)ib.z),
                            ((double)c.x + (double)l.x) - ((double)ic.x + (double)id.x),
                            ((double)c.y + (double)l.y) - ((double)ic.y + (double)id.y),
                            ((double)c.z + (double)l.z) - ((double)ic.z + (double)id.z), (double)a.w + (double)l.w,
                            (double)p.eps2 + (double)TINYF);
        }
        // segmented sum over the runs of equal i (a run = consecutive lanes; the same i may come back in a later
        // run, so lanes are matched by run number, not by i): after the last step the first lane of a run holds
        // its sum
        const int iprev = __shfl_up_sync(0xffffffffu, i, 1);
        const bool head = (lane == 0) || (iprev != i);
        const unsigned int heads = __ballot_sync(0xffffffffu, head);
        const int run = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int ro = __shfl_down_sync(0xffffffffu, run, off);
            const bool take = (lane + off < 32) && (ro == run);
#pragma unroll
            for (int q = 0; q < 7; q++) {
                const double o = __shfl_down_sync(0xffffffffu, v[q], off);
                if (take) v[q] += o;
            }
        }
        if (i >= 0 && head) {
#pragma unroll
            for (int q = 0; q < 7; q++) atomicAdd(p.corr + (size_t)i * 7 + q, v[q]);
        }
    }
}

// Pre-pass of the speculative kernel, one WARP per packed i-particle (the lanes share the window, so the pass costs
// two dependent loads whatever the window): the slot of the j-particle that carries its id (conf[]) and a
// strict upper bound of its nearest-neighbour distance (iD.w) -- the smallest distance to the 2W predicted j
// around its place in the Morton order.
__global__ void __launch_bounds__(256) near_kernel(const ForceArgs p, const int W)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= p.ni) return;
    const float4 a = p.iA[i], b = p.iB[i];
    const int iid = __float_as_int(b.w);
    const int self = hash_lookup(p.ord, iid);
    float d2 = __int_as_float(0x7f800000), cap = __int_as_float(0x7f800000);
    if (p.ord.nkeys > 0) {
        int s = self;
        if (s < 0 || s >= p.ord.nkeys) s = key_search(p.ord, morton30(a.x, a.y, a.z, p.ord.blo, p.ord.binv));
        cap = p.js.capr2[s];
        const int lo = max(0, s - W), hi = min(p.ord.nkeys, s + W + 1);
        for (int j = lo + lane; j < hi; j += 32) {
            const float4 ja = p.js.A[j], jb = p.js.B[j];
            const float dx = (ja.x - a.x) + (jb.x - b.x), dy = (ja.y - a.y) + (jb.y - b.y), dz = (ja.z - a.z) + (jb.z - b.z);
            const float r2 = dx * dx + dy * dy + dz * dz;
            if (ja.w > 0.f && __float_as_int(jb.w) != iid && r2 > TINYF) d2 = fminf(d2, r2);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) d2 = fminf(d2, __shfl_xor_sync(0xffffffffu, d2, off));
    if (lane == 0) {
        p.conf[i] = self;
        float4 d = p.iD[i];
        d.w = d2 * 1.0001f;   // the kernel's own r2 of that pair may round differently
        const_cast<float4 *>(p.iD)[i] = d;
        float4 c = p.iC[i];
        c.w = cap;            // cap of the FP64 radius^2 (order_cap_kernel) of the slot the particle sits on
        const_cast<float4 *>(p.iC)[i] = c;
    }
}

template <int IPT, bool NN, bool NR, int MINB, bool EPS0>
__global__ void __launch_bounds__(THREADS, MINB) force_fast_kernel(const __grid_constant__ ForceArgs p)
{
    // a pair this close sends its group down the exact path (coincident pairs; with EPS0 also pairs for
    // which r2 + 2^-52 is not r2 in FP32)
    constexpr float R2_EXACT = EPS0 ? 1.4901161193847656e-08f /* 2^-26 */ : TINYF;
    static_assert(IPT % 2 == 0, "i-particles are processed in packed pairs");
    static_assert(TILE % GRP == 0 && GRP % (2 * G6_FUNROLL) == 0, "group shape");
    constexpr int NP = IPT / 2;
    constexpr int IB = THREADS * IPT;
    const float INF = __int_as_float(0x7f800000);

    extern __shared__ __align__(128) unsigned char smem_raw[];
    FastSmem &sm = *reinterpret_cast<FastSmem *>(smem_raw);
    const int tid = threadIdx.x;

    const int ntiles_total = (p.nj + TILE - 1) / TILE;
    // split s takes tiles s, s + nsplit, s + 2 nsplit, ...: the tiles that hold an i-block's own neighbourhood
    // (FP64 pairs, masked blocks) are dealt out over all splits instead of landing on one CTA
    const int tstride = p.nsplit;
    int ntiles = ((int)blockIdx.x < ntiles_total) ? (ntiles_total - (int)blockIdx.x + tstride - 1) / tstride : 0;
    auto tile_of = [&](const int t) -> int { return (int)blockIdx.x + t * tstride; };

    if (tid == 0) {
        for (int s = 0; s < FSTAGES; s++) {
            mbar_init(&sm.full[s], 1);
            sm.done[s] = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    constexpr uint32_t STAGE_BYTES = 3u * TILE * sizeof(float4) + TBOX * sizeof(float4);
    auto issue_stage = [&](const int s, const int tile) {
        const size_t off = (size_t)tile * TILE;
        mbar_expect_tx(&sm.full[s], STAGE_BYTES);
        bulk_g2s(sm.A[s], p.jA + off, TILE * sizeof(float4), &sm.full[s]);
        bulk_g2s(sm.B[s], p.jB + off, TILE * sizeof(float4), &sm.full[s]);
        bulk_g2s(sm.C[s], p.jC + off, TILE * sizeof(float4), &sm.full[s]);
        bulk_g2s(sm.G[s], p.jG + (size_t)tile * TBOX, TBOX * sizeof(float4), &sm.full[s]);
    };
    if (tid == 0)
        for (int s = 0; s < FSTAGES && s < ntiles; s++) issue_stage(s, tile_of(s));

    // ---- register-resident i-pairs: i = block base + IPT*tid + k (consecutive = Morton neighbours) --------
    IPair IP[NP];
    int iid[IPT], cgrp[IPT];
    float closek[IPT], d2k[IPT];   // FP64 radius^2 and nearest-neighbour bound^2 of each particle (-1: none)
    auto i_of = [&](int k) -> int { return blockIdx.y * IB + tid * IPT + k; };
    float wlo[3] = {INF, INF, INF}, whi[3] = {-INF, -INF, -INF};
    float closemax = 0.f, nnmax = 0.f;
    bool several = false;
#pragma unroll
    for (int q = 0; q < NP; q++) {
        float4 a[2], b[2], c[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int k = 2 * q + h;
            const int i = i_of(k);
            a[h] = make_float4(0.f, 0.f, 0.f, -1.f);
            b[h] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x80000000));
            c[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            closek[k] = -1.f;
            d2k[k] = -1.f;
            cgrp[k] = -1;
            if (i < p.ni) {
                a[h] = p.iA[i];
                b[h] = p.iB[i];
                c[h] = p.iC[i];
                const float d2 = p.iD[i].w;
                const int cf = p.conf[i];
                d2k[k] = d2;
                closek[k] = (p.ord.kclose > 0.f) ? fminf(fminf(p.ord.kclose * d2, p.ord.cap2), c[h].w) : -1.f;
                closemax = fmaxf(closemax, closek[k]);
                nnmax = fmaxf(nnmax, d2);
                several |= (cf == -2);
                if (cf >= p.slot0) cgrp[k] = (cf - p.slot0) >> 5;
                wlo[0] = fminf(wlo[0], a[h].x); wlo[1] = fminf(wlo[1], a[h].y); wlo[2] = fminf(wlo[2], a[h].z);
                whi[0] = fmaxf(whi[0], a[h].x); whi[1] = fmaxf(whi[1], a[h].y); whi[2] = fmaxf(whi[2], a[h].z);
            }
            iid[k] = __float_as_int(b[h].w);
        }
        IP[q].nxh = pk(-a[0].x, -a[1].x); IP[q].nyh = pk(-a[0].y, -a[1].y); IP[q].nzh = pk(-a[0].z, -a[1].z);
        IP[q].nxl = pk(-b[0].x, -b[1].x); IP[q].nyl = pk(-b[0].y, -b[1].y); IP[q].nzl = pk(-b[0].z, -b[1].z);
        IP[q].nvx = pk(-c[0].x, -c[1].x); IP[q].nvy = pk(-c[0].y, -c[1].y); IP[q].nvz = pk(-c[0].z, -c[1].z);
        IP[q].id0 = iid[2 * q]; IP[q].id1 = iid[2 * q + 1];
        IP[q].h20 = a[0].w; IP[q].h21 = a[1].w;
    }
    // the warp's box, largest FP64 radius and largest neighbour bound (warp-uniform from here on)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            wlo[d] = fminf(wlo[d], __shfl_xor_sync(0xffffffffu, wlo[d], off));
            whi[d] = fmaxf(whi[d], __shfl_xor_sync(0xffffffffu, whi[d], off));
        }
        closemax = fmaxf(closemax, __shfl_xor_sync(0xffffffffu, closemax, off));
        nnmax = fmaxf(nnmax, __shfl_xor_sync(0xffffffffu, nnmax, off));
    }
    const bool always_exact = __any_sync(0xffffffffu, several);
    const float wsc = fmaxf(fmaxf(fmaxf(fabsf(wlo[0]), fabsf(whi[0])), fmaxf(fabsf(wlo[1]), fabsf(whi[1]))),
                            fmaxf(fabsf(wlo[2]), fabsf(whi[2])));
    // a FAR group holds no FP64 pair, no nearest neighbour and no pair the unsoftened shortcut cannot take
    const float farlim = fmaxf(fmaxf(closemax, NN ? nnmax : 0.f), 1.0e-7f);
    auto box_gap2 = [&](const float4 lo, const float4 hi) -> float {
        const float gx = fmaxf(0.f, fmaxf(lo.x - whi[0], wlo[0] - hi.x));
        const float gy = fmaxf(0.f, fmaxf(lo.y - whi[1], wlo[1] - hi.y));
        const float gz = fmaxf(0.f, fmaxf(lo.z - whi[2], wlo[2] - hi.z));
        return gx * gx + gy * gy + gz * gz;
    };

    double D[IPT][7];
    float rmin[IPT], rprev[IPT];   // minimum of r2 over the current group; running minimum (masked rule) before it
    int jgrp[IPT];                 // first j of the group that last lowered it (slot of the window), -1: none
#pragma unroll
    for (int k = 0; k < IPT; k++) {
#pragma unroll
        for (int q = 0; q < 7; q++) D[k][q] = 0.0;
        rmin[k] = rprev[k] = INF;
        jgrp[k] = -1;
    }
    const float eps2 = p.eps2 + TINYF;   // the reference softens by eps2 + 2^-52 (idata.cc:216)
    const u64 eps2p = pk(eps2, eps2);
#ifdef G6_STATS
    unsigned int nmode[4] = {0u, 0u, 0u, 0u};
#endif

    for (int t = 0; t < ntiles; t++) {
        const int s = t % FSTAGES;
        const uint32_t phase = (uint32_t)(t / FSTAGES) & 1u;
        while (!mbar_try_wait(&sm.full[s], phase)) {
        }
        const int jtile = tile_of(t) * TILE;
        int cnt = p.nj - jtile;
        if (cnt > TILE) cnt = TILE;
        const float4 *tA = sm.A[s], *tB = sm.B[s], *tC = sm.C[s], *tG = sm.G[s];

        // whole tile FAR?  (its box is the union of the group boxes)
        bool tile_far = false;
        if (!always_exact && cnt == TILE) {
            const float4 lo = tG[2 * GROUPS_PER_TILE], hi = tG[2 * GROUPS_PER_TILE + 1];
            const float sc = fmaxf(wsc, lo.w);
            bool hit = false;
#pragma unroll
            for (int k = 0; k < IPT; k++) hit |= ((cgrp[k] >> 3) == tile_of(t)) & (cgrp[k] >= 0);
            tile_far = (box_gap2(lo, hi) > fmaxf(farlim, p.ord.farc2 * sc * sc)) && !__any_sync(0xffffffffu, hit);
        }

        for (int jj0 = 0; jj0 < cnt; jj0 += GRP) {
            int mode = 0;   // 0 FAR, 1 NEAR, 2 CLOSE (FP64)
            if (!tile_far) {
                const int gidx = (jtile + jj0) >> 5;
                const float4 lo = tG[2 * (jj0 >> 5)], hi = tG[2 * (jj0 >> 5) + 1];
                const float sc = fmaxf(wsc, lo.w);
                const float gap2 = box_gap2(lo, hi);
                bool hit = false;
#pragma unroll
                for (int k = 0; k < IPT; k++) hit |= (cgrp[k] == gidx);
                if (always_exact || (jj0 + GRP > cnt) || __any_sync(0xffffffffu, hit)) {
                    mode = 2;
                } else if (gap2 > fmaxf(farlim, p.ord.farc2 * sc * sc)) {
                    mode = 0;   // the warp's whole box is far from the group's
                } else {
                    // particle by particle: distance of the lane's own two particles to the group's box against
                    // their own FP64 radius (-> CLOSE), their own neighbour bound and the distance below which the
                    // lo parts of the coordinates matter (-> NEAR); a warp of particles that are NOT neighbours of
                    // each other (a chunk of a caller's i-list) keeps most groups FAR this way
                    bool cl = false, nr = false;
#pragma unroll
                    for (int q = 0; q < NP; q++) {
                        float x0, x1, y0, y1, z0, z1;
                        upk(IP[q].nxh, x0, x1); upk(IP[q].nyh, y0, y1); upk(IP[q].nzh, z0, z1);   // negated positions
                        const float ax = fmaxf(0.f, fmaxf(lo.x + x0, -x0 - hi.x)), ay = fmaxf(0.f, fmaxf(lo.y + y0, -y0 - hi.y)),
                                    az = fmaxf(0.f, fmaxf(lo.z + z0, -z0 - hi.z));
                        const float bx = fmaxf(0.f, fmaxf(lo.x + x1, -x1 - hi.x)), by = fmaxf(0.f, fmaxf(lo.y + y1, -y1 - hi.y)),
                                    bz = fmaxf(0.f, fmaxf(lo.z + z1, -z1 - hi.z));
                        const float pa = ax * ax + ay * ay + az * az, pb = bx * bx + by * by + bz * bz;
                        const float sa = fmaxf(lo.w, fmaxf(fmaxf(fabsf(x0), fabsf(y0)), fabsf(z0)));
                        const float sb = fmaxf(lo.w, fmaxf(fmaxf(fabsf(x1), fabsf(y1)), fabsf(z1)));
                        cl |= (pa <= closek[2 * q]) | (pb <= closek[2 * q + 1]);
                        nr |= (pa <= fmaxf(d2k[2 * q], p.ord.farc2 * sa * sa)) & (d2k[2 * q] >= 0.f);
                        nr |= (pb <= fmaxf(d2k[2 * q + 1], p.ord.farc2 * sb * sb)) & (d2k[2 * q + 1] >= 0.f);
                    }
                    const unsigned int clm = __ballot_sync(0xffffffffu, cl), nrm = __ballot_sync(0xffffffffu, nr);
                    mode = clm ? 2 : (nrm ? 1 : 0);
                }
            }
            Acc7P S[NP];
#pragma unroll
            for (int q = 0; q < NP; q++) S[q] = Acc7P{0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int k = 0; k < IPT; k++) rmin[k] = INF;
            if (mode == 0) {
#pragma unroll FUNROLL
                for (int u = 0; u < GRP; u += 2) {
                    const int jj = jj0 + u;
                    const float4 a0 = tA[jj], c0 = tC[jj];
                    const float4 a1 = tA[jj + 1], c1 = tC[jj + 1];
#pragma unroll
                    for (int q = 0; q < NP; q++) {
                        // far pairs are many and individually small: the 2^-22.9 error of the raw reciprocal square
                        // root averages out over them (the pairs that dominate a sum are NEAR or CLOSE)
                        interact2_fast<NR && !G6_FAR_RAW, EPS0, true>(a0, a0, c0, IP[q], eps2p, S[q]);
                        interact2_fast<NR && !G6_FAR_RAW, EPS0, true>(a1, a1, c1, IP[q], eps2p, S[q]);
                    }
                }
            } else if (mode == 1) {
#pragma unroll FUNROLL
                for (int u = 0; u < GRP; u += 2) {
                    const int jj = jj0 + u;
                    const float4 a0 = tA[jj], b0 = tB[jj], c0 = tC[jj];
                    const float4 a1 = tA[jj + 1], b1 = tB[jj + 1], c1 = tC[jj + 1];
#pragma unroll
                    for (int q = 0; q < NP; q++) {
                        const u64 ra = interact2_fast<NR, EPS0, false>(a0, b0, c0, IP[q], eps2p, S[q]);
                        const u64 rb = interact2_fast<NR, EPS0, false>(a1, b1, c1, IP[q], eps2p, S[q]);
                        float ra0, ra1, rb0, rb1;
                        upk(ra, ra0, ra1);
                        upk(rb, rb0, rb1);
                        rmin[2 * q] = min3f(rmin[2 * q], ra0, rb0);
                        rmin[2 * q + 1] = min3f(rmin[2 * q + 1], ra1, rb1);
                    }
                }
                bool bad = false;
#pragma unroll
                for (int k = 0; k < IPT; k++) bad |= !(rmin[k] > R2_EXACT);
                if (__any_sync(0xffffffffu, bad)) {   // a pair too close for the mask-free path
                    mode = 2;
#ifdef G6_STATS
                    nmode[3]++;
#endif
                }
            }
#ifdef G6_STATS
            nmode[mode]++;
#endif
            if (mode == 2) {
                // the reference's rule pair by pair (equal ids skipped, pot and neighbour search only for
                // r2 > 2^-52) in FP32; pairs inside a particle's FP64 radius are left out of the sums and queued
#pragma unroll
                for (int q = 0; q < NP; q++) S[q] = Acc7P{0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
#pragma unroll
                for (int k = 0; k < IPT; k++) rmin[k] = INF;
                const int jend = (jj0 + GRP < cnt) ? jj0 + GRP : cnt;
                unsigned int hpm[IPT];   // bit u: pair (particle k, j = jj0 + u) is inside k's FP64 radius
#pragma unroll
                for (int k = 0; k < IPT; k++) hpm[k] = 0u;
                for (int jj = jj0; jj < jend; jj++) {
                    const float4 a = tA[jj], b = tB[jj], c = tC[jj];
                    int unused0 = 0, unused1 = 0;
#pragma unroll
                    for (int q = 0; q < NP; q++) {
                        const int hp = interact2<true, false, NR, false>(a, b, c, 0, IP[q], eps2p, closek[2 * q],
                                                                         closek[2 * q + 1], S[q], rmin[2 * q], unused0,
                                                                         rmin[2 * q + 1], unused1, 0, p);
                        hpm[2 * q] |= (unsigned)(hp & 1) << (jj - jj0);
                        hpm[2 * q + 1] |= (unsigned)((hp >> 1) & 1) << (jj - jj0);
                    }
                }
                // queue the block's FP64 pairs with ONE atomic: warp prefix sum of the lanes' counts
                int mine = 0;
#pragma unroll
                for (int k = 0; k < IPT; k++) mine += __popc(hpm[k]);
                if (__any_sync(0xffffffffu, mine != 0)) {
                    const int lane = tid & 31;
                    int incl = mine;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, incl, off);
                        if (lane >= off) incl += o;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    unsigned int base = 0u;
                    if (lane == 0) base = atomicAdd(p.wl_count, (unsigned)total);
                    base = __shfl_sync(0xffffffffu, base, 0);
#ifdef G6_STATS
                    if (lane == 0 && p.stats) {
                        atomicAdd(&p.stats[4], (unsigned long long)total);
                        if (base + (unsigned)total > p.wl_cap) atomicAdd(&p.stats[5], 1ull);
                    }
#endif
                    unsigned int pos = base + (unsigned)(incl - mine);
#pragma unroll
                    for (int k = 0; k < IPT; k++) {
                        const int i = i_of(k);
                        unsigned int m = hpm[k];
                        while (m) {
                            const int u = __ffs(m) - 1;
                            m &= m - 1u;
                            if (pos < p.wl_cap) {
                                p.wl[pos] = make_int2(i, jtile + jj0 + u);
                            } else {   // list full: evaluate here
                                close_pair_to_corr(p, p.iA[i], p.iB[i], p.iC[i], p.iD[i], i, jtile + jj0 + u);
                            }
                            pos++;
                        }
                    }
                }
            }
            {
                const double sg = (mode == 2) ? 1.0 : -1.0;   // the mask-free pair function accumulates -acc, -jerk
#pragma unroll
                for (int q = 0; q < NP; q++) {
                    float lo, hi;
                    upk(S[q].ax, lo, hi); D[2 * q][0] += sg * (double)lo; D[2 * q + 1][0] += sg * (double)hi;
                    upk(S[q].ay, lo, hi); D[2 * q][1] += sg * (double)lo; D[2 * q + 1][1] += sg * (double)hi;
                    upk(S[q].az, lo, hi); D[2 * q][2] += sg * (double)lo; D[2 * q + 1][2] += sg * (double)hi;
                    upk(S[q].jx, lo, hi); D[2 * q][3] += sg * (double)lo; D[2 * q + 1][3] += sg * (double)hi;
                    upk(S[q].jy, lo, hi); D[2 * q][4] += sg * (double)lo; D[2 * q + 1][4] += sg * (double)hi;
                    upk(S[q].jz, lo, hi); D[2 * q][5] += sg * (double)lo; D[2 * q + 1][5] += sg * (double)hi;
                    upk(S[q].pot, lo, hi); D[2 * q][6] += (double)lo; D[2 * q + 1][6] += (double)hi;
                }
            }
            if (NN && mode != 0) {
#pragma unroll
                for (int k = 0; k < IPT; k++) {
                    if (rmin[k] < rprev[k]) {   // strict: the first group that reaches the minimum keeps it
                        jgrp[k] = jtile + jj0;
                        rprev[k] = rmin[k];
                    }
                }
            }
        }

        // this warp is done with stage s; the last warp of the CTA to say so refills it
        __syncwarp();
        if ((tid & 31) == 0) {
            unsigned int old;
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;"
                         : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(&sm.done[s])) : "memory");
            if (old == (unsigned)(THREADS / 32 - 1)) {
                sm.done[s] = 0u;
                if (t + FSTAGES < ntiles) issue_stage(s, tile_of(t + FSTAGES));
            }
        }
    }

#ifdef G6_STATS
    if (p.stats && (tid & 31) == 0)
        for (int m = 0; m < 4; m++) atomicAdd(&p.stats[m], (unsigned long long)nmode[m]);
#endif
    // ---- nearest neighbour: re-scan the one group that holds the minimum ---------------------
    u64 key[IPT];
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        int jm = -1;
        if (NN && jgrp[k] >= 0 && rprev[k] < 1.0e30f) {   // >= 1e30: only parked (massless) slots were seen
            const int jend = (jgrp[k] + GRP < p.nj) ? jgrp[k] + GRP : p.nj;
            for (int jj = jgrp[k]; jj < jend; jj++) {
                const float4 a = p.jA[jj], b = p.jB[jj];
                u64 dx, dy, dz, r2;
                pair_geometry(a, b, IP[k >> 1], dx, dy, dz, r2);
                float r0, r1;
                upk(r2, r0, r1);
                const float r = (k & 1) ? r1 : r0;
                if (__float_as_int(b.w) != iid[k] && r > TINYF && r == rprev[k]) {
                    jm = jj;
                    break;
                }
            }
        }
        key[k] = (jm >= 0) ? make_key(rprev[k], jm) : KEY_NONE;
    }

    // ---- totals -> global (final or per-split partial) -----------------------------------
    // (with the FP64 pair list in use the launch always ends at the partials: close_pairs_kernel and
    // reduce_partials_kernel follow)
    const bool single = (p.nsplit == 1) && !p.corr;
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        const int i = i_of(k);
        if (i >= p.ni) continue;
        if (single) {
            store_outputs<NN>(p, i, D[k], key[k]);
        } else {
            size_t o = (size_t)blockIdx.x * p.ni_pad + i;
#pragma unroll
            for (int q = 0; q < 7; q++) p.part_sum[o * 7 + q] = D[k][q];
            p.part_key[o] = key[k];
        }
    }
    if (single) signal_done(p);
    if (single || p.defer_reduce) return;
    reduce_splits<NN>(p, &sm.is_last, IB);
}

// Sum of the per-split partials as a kernel of its own, for launches of the speculative kernel with
// few i-blocks and many j-splits, where the last CTA of an i-block would sum hundreds of splits
// alone.  One warp per i: the lanes take strided subsets of the splits (all loads in flight at once)
// and a fixed butterfly of shuffles combines them (deterministic).
__global__ void __launch_bounds__(256) reduce_partials_kernel(const ForceArgs p, const int want_nn)
{
    if (p.wl_count && blockIdx.x == 0 && threadIdx.x == 0) *p.wl_count = 0u;   // close_pairs_kernel is done with it
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i < p.ni) {
        double tot[7] = {0, 0, 0, 0, 0, 0, 0};
        u64 kk = KEY_NONE;
        int sp = lane;
        for (; sp + 96 < p.nsplit; sp += 128) {
            double v[4][7];
            u64 kv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                size_t o = (size_t)(sp + 32 * u) * p.ni_pad + i;
#pragma unroll
                for (int q = 0; q < 7; q++) v[u][q] = __ldcg(p.part_sum + o * 7 + q);
                kv[u] = __ldcg(p.part_key + o);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
#pragma unroll
                for (int q = 0; q < 7; q++) tot[q] += v[u][q];
                kk = kv[u] < kk ? kv[u] : kk;
            }
        }
        for (; sp < p.nsplit; sp += 32) {
            size_t o = (size_t)sp * p.ni_pad + i;
#pragma unroll
            for (int q = 0; q < 7; q++) tot[q] += __ldcg(p.part_sum + o * 7 + q);
            u64 ok = __ldcg(p.part_key + o);
            kk = ok < kk ? ok : kk;
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
            for (int q = 0; q < 7; q++) tot[q] += __shfl_xor_sync(0xffffffffu, tot[q], off);
            u64 o = __shfl_xor_sync(0xffffffffu, kk, off);
            kk = o < kk ? o : kk;
        }
        if (lane == 0) {
            if (want_nn) store_outputs<true>(p, i, tot, kk);
            else store_outputs<false>(p, i, tot, kk);
        }
    }
    signal_done(p);
}

// After a min-reduction of keys over ranks (each rank holds a j-shard): the id of the winner if this rank
// owns it (keys carry the reported address = local address + j_offset).
__global__ void resolve_nn_kernel(int ni, const u64 *__restrict__ key, int rank, int j_offset, int nj_local,
                                  const int *__restrict__ slot_of, const float4 *__restrict__ jB, int *nnid,
                                  const int win_lo, const int win_hi)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ni) return;
    u64 k = key[i];
    int id = 0;
    if (k == KEY_NONE) {
        id = (rank == 0) ? -1 : 0;
    } else {
        int a = (int)(unsigned)(k & 0xffffffffu) - j_offset;
        if (a >= 0 && a < nj_local) {   // this rank owns it if it holds the address and the slot is in its window
            const int sl = slot_of[a];
            if (sl >= win_lo && sl < win_hi) id = __float_as_int(jB[sl].w);
        }
    }
    nnid[i] = id;
}

// ---------------------------------------------------------------------------
// j-memory order (host: rebuild_order): bounding box, Morton keys, permutation, id table, neighbour bounds.
// ---------------------------------------------------------------------------
// box[0..2] = min, box[3..5] = max of the positions of the massive particles with address < nj (ordered ints);
// cen[0..2] = sum of their positions, cen[3] = their number (the origin is their mean: the lo parts of the
// double-single coordinates are smallest where the particles are)
__global__ void __launch_bounds__(256) order_box_kernel(const int n, const int nj, const JState s, const int *addr_of,
                                                        int *box, double *cen)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool use = (j < n) && (addr_of[j] < nj) && (s.q[6][j].y > (double)TINYF);
    double xd = 0.0, yd = 0.0, zd = 0.0, cnt = use ? 1.0 : 0.0;
    if (use) {
        xd = s.q[0][j].x;
        yd = s.q[0][j].y;
        zd = s.q[1][j].x;
    }
    const float x = (float)xd, y = (float)yd, z = (float)zd;
    const int big = 0x7f7fffff, small = f2ord(-3.0e38f);
    int v[6] = {use ? f2ord(x) : big, use ? f2ord(y) : big, use ? f2ord(z) : big,
                use ? f2ord(x) : small, use ? f2ord(y) : small, use ? f2ord(z) : small};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        v[d] = __reduce_min_sync(0xffffffffu, v[d]);
        v[3 + d] = __reduce_max_sync(0xffffffffu, v[3 + d]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        xd += __shfl_xor_sync(0xffffffffu, xd, off);
        yd += __shfl_xor_sync(0xffffffffu, yd, off);
        zd += __shfl_xor_sync(0xffffffffu, zd, off);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0.0) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            atomicMin(&box[d], v[d]);
            atomicMax(&box[3 + d], v[3 + d]);
        }
        atomicAdd(&cen[0], xd);
        atomicAdd(&cen[1], yd);
        atomicAdd(&cen[2], zd);
        atomicAdd(&cen[3], cnt);
    }
}
// sort keys: Morton key of the position (relative to x0) for massive particles of the prefix [0, nj);
// massless ones go to the end of the prefix, addresses >= nj behind it in address order
__global__ void __launch_bounds__(256) order_key_kernel(const int n, const int nj, const JState s, const int *addr_of,
                                                        const OrderInfo o, unsigned *key, int *val)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int a = addr_of[j];
    unsigned k;
    if (a >= nj) {
        k = 0x80000000u | (unsigned)a;
    } else if (!(s.q[6][j].y > (double)TINYF)) {
        k = 0x7fffffffu;
    } else {
        k = morton30((float)(s.q[0][j].x - s.x0[0]), (float)(s.q[0][j].y - s.x0[1]), (float)(s.q[1][j].x - s.x0[2]), o.blo,
                     o.binv);
    }
    key[j] = k;
    val[j] = j;
}
// new slot j < n takes what old slot perm[j] held; slots [n, ncap) are copied as they are (the two sets of
// arrays swap roles afterwards, so every slot of the capacity must arrive in the target set)
__global__ void __launch_bounds__(256) order_permute_kernel(const int n, const int ncap, const int *__restrict__ perm,
                                                            const JState from, const int *__restrict__ addr_from,
                                                            JState to, int *addr_to, int *slot_of)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncap) return;
    const int o = j < n ? perm[j] : j;
#pragma unroll
    for (int k = 0; k < 7; k++) to.q[k][j] = from.q[k][o];
    to.ia[j] = from.ia[o];
    to.near2[j] = from.near2[o];
    const int a = addr_from[o];
    addr_to[j] = a;
    slot_of[a] = j;
}
// upper bound of every particle's nearest-neighbour distance from its 2W Morton neighbours (state positions)
__global__ void __launch_bounds__(256) order_near_kernel(const int n, const JState s, const int W)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float d2 = __int_as_float(0x7f800000);
    if (s.q[6][j].y > (double)TINYF) {
        const double x = s.q[0][j].x, y = s.q[0][j].y, z = s.q[1][j].x;
        const int id = s.ia[j].x;
        const int lo = max(0, j - W), hi = min(n, j + W + 1);
        for (int k = lo; k < hi; k++) {
            if (k == j || s.ia[k].x == id || !(s.q[6][k].y > (double)TINYF)) continue;
            const double dx = s.q[0][k].x - x, dy = s.q[0][k].y - y, dz = s.q[1][k].x - z;
            const float r2 = (float)(dx * dx + dy * dy + dz * dz);
            if (r2 > TINYF) d2 = fminf(d2, r2);
        }
    }
    s.near2[j] = d2;
}
// Group and tile boxes of the STATE positions (same layout as predict_tile's, which overwrites them at the next
// prediction): what order_cap_kernel measures against.  One CTA per tile, one warp per group.
__global__ void __launch_bounds__(TILE) order_boxes_kernel(const int n, const JState s)
{
    __shared__ int sh[6][TILE / 32];
    const int j = blockIdx.x * TILE + threadIdx.x;   // < capacity (whole tiles)
    const bool massive = (j < n) && (s.q[6][j].y > (double)TINYF);
    float x = 0.f, y = 0.f, z = 0.f;
    if (massive) {
        x = (float)(s.q[0][j].x - s.x0[0]);
        y = (float)(s.q[0][j].y - s.x0[1]);
        z = (float)(s.q[1][j].x - s.x0[2]);
    }
    const int big = 0x7f7fffff, small = f2ord(-3.0e38f);
    const int lx = __reduce_min_sync(0xffffffffu, massive ? f2ord(x) : big);
    const int ly = __reduce_min_sync(0xffffffffu, massive ? f2ord(y) : big);
    const int lz = __reduce_min_sync(0xffffffffu, massive ? f2ord(z) : big);
    const int hx = __reduce_max_sync(0xffffffffu, massive ? f2ord(x) : small);
    const int hy = __reduce_max_sync(0xffffffffu, massive ? f2ord(y) : small);
    const int hz = __reduce_max_sync(0xffffffffu, massive ? f2ord(z) : small);
    const int w = threadIdx.x >> 5;
    float4 *tb = s.gbb + (size_t)blockIdx.x * TBOX;
    if ((threadIdx.x & 31) == 0) {
        tb[2 * w] = make_float4(ord2f(lx), ord2f(ly), ord2f(lz), 0.f);
        tb[2 * w + 1] = make_float4(ord2f(hx), ord2f(hy), ord2f(hz), 0.f);
        sh[0][w] = lx; sh[1][w] = ly; sh[2][w] = lz; sh[3][w] = hx; sh[4][w] = hy; sh[5][w] = hz;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int r[6];
#pragma unroll
        for (int q = 0; q < 6; q++) r[q] = sh[q][0];
#pragma unroll
        for (int g = 1; g < TILE / 32; g++) {
#pragma unroll
            for (int q = 0; q < 3; q++) r[q] = min(r[q], sh[q][g]);
#pragma unroll
            for (int q = 3; q < 6; q++) r[q] = max(r[q], sh[q][g]);
        }
        tb[2 * GROUPS_PER_TILE] = make_float4(ord2f(r[0]), ord2f(r[1]), ord2f(r[2]), 0.f);
        tb[2 * GROUPS_PER_TILE + 1] = make_float4(ord2f(r[3]), ord2f(r[4]), ord2f(r[5]), 0.f);
    }
}
// Cap of every particle's FP64 radius.  The radius K d^2 is meant to hold the few hundred strongest pairs of a
// particle; for a particle of the halo whose nearest neighbour is far away while the dense core is not, it would
// hold a large part of the system (a few particles in a thousand, which then dominate the FP64 work and make
// their whole warp take the masked path against every group).  Such pairs are many and comparable, their FP32
// errors average out like those of the far field, so the radius is cut back to the largest of
// 64 d^2 / 2^b (b = 0..5) within which the particle touches at most gmax group boxes (never below 2 d^2: the
// nearest neighbour stays inside).  d^2 = near2 (the bound from the Morton window).
constexpr int CAP_BINS = 6;
__global__ void __launch_bounds__(256) order_cap_kernel(const int n, const JState s, const int ntiles, const int gmax)
{
    __shared__ float4 tb[2 * 256];
    const int j = blockIdx.x * 256 + threadIdx.x;
    const float INF = __int_as_float(0x7f800000);
    float near2 = (j < n) ? s.near2[j] : INF;
    const bool live = (j < n) && (s.q[6][j].y > (double)TINYF) && (near2 < 1.0e30f) && (near2 > 0.f);
    float x = 0.f, y = 0.f, z = 0.f;
    if (live) {
        x = (float)(s.q[0][j].x - s.x0[0]);
        y = (float)(s.q[0][j].y - s.x0[1]);
        z = (float)(s.q[1][j].x - s.x0[2]);
    }
    const float c0 = 64.f * near2;
    int cnt[CAP_BINS];
#pragma unroll
    for (int b = 0; b < CAP_BINS; b++) cnt[b] = 0;
    auto gap2 = [&](const float4 lo, const float4 hi) -> float {
        const float gx = fmaxf(0.f, fmaxf(lo.x - x, x - hi.x)), gy = fmaxf(0.f, fmaxf(lo.y - y, y - hi.y)),
                    gz = fmaxf(0.f, fmaxf(lo.z - z, z - hi.z));
        return gx * gx + gy * gy + gz * gz;
    };
    for (int t0 = 0; t0 < ntiles; t0 += 256) {
        __syncthreads();
        if (t0 + (int)threadIdx.x < ntiles) {
            const float4 *g = s.gbb + (size_t)(t0 + threadIdx.x) * TBOX;
            tb[2 * threadIdx.x] = g[2 * GROUPS_PER_TILE];
            tb[2 * threadIdx.x + 1] = g[2 * GROUPS_PER_TILE + 1];
        }
        __syncthreads();
        const int m = min(256, ntiles - t0);
        if (live) {
            for (int u = 0; u < m; u++) {
                if (gap2(tb[2 * u], tb[2 * u + 1]) <= c0) {
                    const float4 *g = s.gbb + (size_t)(t0 + u) * TBOX;
                    for (int k = 0; k < GROUPS_PER_TILE; k++) {
                        const float d2 = gap2(g[2 * k], g[2 * k + 1]);
                        float c = c0;
#pragma unroll
                        for (int b = 0; b < CAP_BINS; b++) {
                            cnt[b] += (d2 <= c) ? 1 : 0;
                            c *= 0.5f;
                        }
                    }
                }
            }
        }
    }
    if (j < n) {
        float cap = INF;
        if (live && cnt[0] > gmax) {
            cap = c0 * (1.f / (float)(1 << (CAP_BINS - 1)));
            float c = c0;
#pragma unroll
            for (int b = 1; b < CAP_BINS; b++) {
                c *= 0.5f;
                if (cnt[b] <= gmax) {
                    cap = c;
                    break;
                }
            }
        }
        s.capr2[j] = cap;
    }
}
// Morton keys of an i-set (device-resident callers), for the sort that precedes pack_i_kernel
__global__ void __launch_bounds__(256) i_key_kernel(const int ni, const double *__restrict__ xi, const JState s,
                                                    const OrderInfo o, unsigned *key, int *val)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ni) return;
    key[i] = morton30((float)(xi[3 * (size_t)i] - s.x0[0]), (float)(xi[3 * (size_t)i + 1] - s.x0[1]),
                      (float)(xi[3 * (size_t)i + 2] - s.x0[2]), o.blo, o.binv);
    val[i] = i;
}

// ---------------------------------------------------------------------------
// Device-resident Hermite block step (the steps either side of the force call in ph4's
// idata::advance, src/amuse_ph4/src/idata.cc:832-870): the active particles ARE j-particles, so their
// state is gathered from the j-memory, predicted (idata.cc:347-365), pushed through the force kernels,
// corrected with the Aarseth step and its block quantisation (idata.cc:443-511), and written back
// into the j-memory -- no host round trip of positions, no per-particle g6_set_j_particle.
// All of this is FP64 with the reference's expression trees.
// ---------------------------------------------------------------------------

__device__ __forceinline__ void hermite_gather_one(const HermiteArgs &h, const int i)
{
    const int a = h.slot_of[h.ilist[i]];
    const JState &s = h.js;
    const PredJ pj = predict_slot(s, a, h.tnext);   // idata.cc:353-361 is the same polynomial as jdata.cc:739-746
    const int id = s.ia[a].x;
    const double px = pj.x, py = pj.y, pz = pj.z, qx = pj.vx, qy = pj.vy, qz = pj.vz;
    double *pr = h.pred + (size_t)i * 6;
    pr[0] = px; pr[1] = py; pr[2] = pz; pr[3] = qx; pr[4] = qy; pr[5] = qz;
    h.ilist_d[i] = a;
    h.olddt_d[i] = (h.mode == 0) ? h.old_dt[i] : 0.0;
    const double rx = px - s.x0[0], ry = py - s.x0[1], rz = pz - s.x0[2];
    const float xh = (float)rx, yh = (float)ry, zh = (float)rz;
    const float vxh = (float)qx, vyh = (float)qy, vzh = (float)qz;
    h.iA[i] = make_float4(xh, yh, zh, 0.f);
    h.iB[i] = make_float4((float)(rx - (double)xh), (float)(ry - (double)yh), (float)(rz - (double)zh),
                          __int_as_float(id));
    h.iC[i] = make_float4(vxh, vyh, vzh, 0.f);
    h.iD[i] = make_float4((float)(qx - (double)vxh), (float)(qy - (double)vyh), (float)(qz - (double)vzh), 0.f);
}

// Corrector (or initialisation) of active particle i given its new force f[7] and neighbour id.
__device__ __forceinline__ void hermite_correct_one(const HermiteArgs &h, const int i, const double *f, const int nnid)
{
    const int a = h.ilist_d[i];
    const JState &s = h.js;
    const double2 q1 = s.q[1][a], q3 = s.q[3][a], q4 = s.q[4][a], q5 = s.q[5][a], q6 = s.q[6][a];
    const double told = q1.y;
    const double oa[3] = {q3.y, q4.x, q4.y}, oj[3] = {q5.x, q5.y, q6.x};
    const double ia[3] = {f[0], f[1], f[2]}, ij[3] = {f[3], f[4], f[5]};
    const double *pr = h.pred + (size_t)i * 6;
    double pos[3] = {pr[0], pr[1], pr[2]}, vel[3] = {pr[3], pr[4], pr[5]};
    double newstep;
    if (h.mode == 0) {   // idata.cc:443-511
        const double dt = h.tnext - told;
        const double dt2 = dt * dt;
        double a2 = 0, j2 = 0, k2 = 0, l2 = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double alpha = -3 * (oa[k] - ia[k]) - dt * (2 * oj[k] + ij[k]);
            const double beta = 2 * (oa[k] - ia[k]) + dt * (oj[k] + ij[k]);
            pos[k] += (alpha / 12 + beta / 20) * dt2;
            vel[k] += (alpha / 3 + beta / 4) * dt;
            a2 += ia[k] * ia[k];
            j2 += (dt * ij[k]) * (dt * ij[k]);
            k2 += (2 * alpha) * (2 * alpha);
            l2 += (6 * beta) * (6 * beta);
        }
        newstep = h.eta * dt * sqrt((sqrt(a2 * k2) + j2) / (sqrt(j2 * l2) + k2));
        int exponent;
        const double olddt = h.olddt_d[i];
        double oldstep2 = olddt / (2 * frexp(olddt, &exponent));
        // (bounded: a step that is not a positive power of two -- a caller error or a NaN force -- must not
        // hang the device; the host sees the NaN step and stops)
        for (int it = 0; it < 1200 && fmod(h.tnext, oldstep2) != 0; it++) oldstep2 /= 2;
        if (!(olddt > 0.0) || !(oldstep2 > 0.0)) oldstep2 = __longlong_as_double(0x7ff8000000000000ll);
        if (newstep < oldstep2) {
            newstep = oldstep2 / 2;
        } else {
            const double t2 = 2 * oldstep2;
            newstep = (newstep >= t2 && fmod(h.tnext, t2) == 0) ? t2 : oldstep2;
        }
    } else {             // jdata.cc:503-548 (fac 0.0625, limit 0.03125)
        double a2 = 0, j2 = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            a2 += ia[k] * ia[k];
            j2 += ij[k] * ij[k];
        }
        const double fac = 0.0625, limit = 0.03125;
        double first = (h.eta == 0.0) ? limit : ((a2 == 0.0 || j2 == 0.0) ? fac * h.eta : fac * h.eta * sqrt(a2 / j2));
        if (first != first) first = fac * h.eta;
        int exponent;
        first /= 2 * frexp(first, &exponent);
        for (int it = 0; it < 1200 && fmod(h.tnext, first) != 0; it++) first /= 2;
        for (int it = 0; it < 1200 && first > limit; it++) first /= 2;
        newstep = first;
    }
    s.q[0][a] = make_double2(pos[0], pos[1]);
    s.q[1][a] = make_double2(pos[2], h.tnext);
    s.q[2][a] = make_double2(vel[0], vel[1]);
    s.q[3][a] = make_double2(vel[2], ia[0]);
    s.q[4][a] = make_double2(ia[1], ia[2]);
    s.q[5][a] = make_double2(ij[0], ij[1]);
    s.q[6][a] = make_double2(ij[2], q6.y);
    h.out_dt[i] = newstep;
    h.out_pot[i] = -f[6];
    h.out_nn[i] = nnid;
}

__global__ void __launch_bounds__(256) hermite_gather_kernel(const HermiteArgs h)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < h.ni) hermite_gather_one(h, i);
}

__global__ void __launch_bounds__(256) hermite_correct_kernel(const HermiteArgs h)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < h.ni) hermite_correct_one(h, i, h.sum + (size_t)i * 7, h.nnid[i]);
    // completion flag in mapped host memory (same protocol as signal_done)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && h.host_flag) {
        bool last = true;
        if (gridDim.x > 1) {
            const unsigned int k = atomicAdd(h.done_counter, 1u);
            last = (k == gridDim.x - 1u);
            if (last) *h.done_counter = 0u;
        }
        if (last) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(h.host_flag) = h.flag_seq;
        }
    }
}

// predict all j tiles AND gather/predict the active particles in one launch:
// CTAs [0, ntiles) predict a tile each, the CTAs after them take 256 active particles each.
__global__ void __launch_bounds__(TILE) hermite_predict_gather_kernel(const int ntiles, const int n, const double ti,
                                                                      const HermiteArgs h)
{
    if ((int)blockIdx.x < ntiles) {
        predict_tile(blockIdx.x, n, ti, h.js);
        return;
    }
    const int i = (blockIdx.x - ntiles) * TILE + threadIdx.x;
    if (i < h.ni) hermite_gather_one(h, i);
}

// ---------------------------------------------------------------------------
// Multi-GPU exchange (one process per GPU, peer memory over NVLink/NVSwitch).
//
// Every rank owns an exchange buffer with one slot per rank: slot[r] = { sum[cap][7], key[cap],
// id[cap] } holds rank r's partial results (written by r's force kernels, see store_outputs) and
// flag[r] the sequence number of the last exchange r has completed into this buffer.  After its last
// force launch of an exchange, rank r runs peer_flag_kernel (remote stores of seq into flag[r] at every
// peer); peer_combine_kernel then waits for all flags of the LOCAL buffer and combines the slots in
// rank order -- the device-side form of idata.cc:284-313 (sum pot/acc/jerk, min dnn, nn of the winner),
// identical on every rank.
// ---------------------------------------------------------------------------
struct PeerSlots {
    int world, rank;
    const double *sum[MAX_PEERS + 1];   // local buffer, slot r: [cap][7]
    const u64 *key[MAX_PEERS + 1];
    const int *id[MAX_PEERS + 1];
    volatile unsigned long long *flag;  // local buffer: [world]
    unsigned long long *remote_flag[MAX_PEERS];   // &flag[rank] at every peer
    int n_remote;
};

__global__ void peer_flag_kernel(const PeerSlots ps, const unsigned long long seq)
{
    // the force kernels that wrote into the peers' slots precede this kernel on the stream
    __threadfence_system();
    if (threadIdx.x < ps.n_remote) {
        *reinterpret_cast<volatile unsigned long long *>(ps.remote_flag[threadIdx.x]) = seq;
    }
    if (threadIdx.x == 0) ps.flag[ps.rank] = seq;
}

// grid <= resident CTAs (the host sizes it): every CTA waits, then takes a grid-stride share of i
// host_flag != NULL (single CTA): the totals go to mapped host memory and the flag announces them.
__global__ void __launch_bounds__(256) peer_combine_kernel(const PeerSlots ps, const unsigned long long seq, const int ni,
                                                          double *out_sum, u64 *out_key, int *out_nnid,
                                                          unsigned int *error_word, unsigned long long *host_flag,
                                                          const unsigned long long flag_seq)
{
    __shared__ int ok;
    if (threadIdx.x == 0) {
        ok = 1;
        for (int r = 0; r < ps.world; r++) {
            unsigned long long spins = 0;
            while (ps.flag[r] < seq) {
                __nanosleep(200);
                if (++spins > (1ull << 26)) {   // ~15 s: a peer died; report instead of hanging the GPU
                    ok = 0;
                    atomicExch(error_word, 1u);
                    break;
                }
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (!ok) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ni; i += gridDim.x * blockDim.x) {
        double tot[7] = {0, 0, 0, 0, 0, 0, 0};
        u64 kk = KEY_NONE;
        int id = -1;
        for (int r = 0; r < ps.world; r++) {
            const double *s = ps.sum[r] + (size_t)i * 7;
#pragma unroll
            for (int q = 0; q < 7; q++) tot[q] += __ldcv(s + q);
            const u64 k = __ldcv(ps.key[r] + i);
            if (k < kk) {   // strict: the lowest rank wins exact ties, like the ascending-j CPU scan
                kk = k;
                id = __ldcv(ps.id[r] + i);
            }
        }
#pragma unroll
        for (int q = 0; q < 7; q++) out_sum[(size_t)i * 7 + q] = tot[q];
        out_key[i] = kk;
        out_nnid[i] = id;
    }
    if (host_flag) {   // gridDim.x == 1
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(host_flag) = flag_seq;
    }
}

// Launch-latency probe: the floor of "launch k dependent kernels, the last one raises a flag in mapped host
// memory, the host spins on it" -- what one block step of the latency path can cost at best.
__global__ void latency_probe_kernel(unsigned long long *host_flag, unsigned long long seq, unsigned int *sink)
{
    if (sink && threadIdx.x == 1234567) *sink = 1;
    if (host_flag && threadIdx.x == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(host_flag) = seq;
    }
}

// ---------------------------------------------------------------------------
// FP32 pipe microbenchmark (roofline denominator measured on the device).
// ---------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, int iters, float seed)
{
    // MODE 0: scalar FFMA, operands mostly from the reuse cache      (x = x*m + a)
    // MODE 1: packed FFMA2, same pattern
    // MODE 2: scalar FFMA, three distinct rotating register operands (x[c] = x[c+1]*x[c+2] + x[c])
    // MODE 3: packed FFMA2, three distinct rotating 64-bit operands
    // MODE 4: packed FFMA2 with a broadcast 32-bit operand           (x[c] = x[c+1]*s + x[c])
    // MODE 5: MODE 3 with one ALU op (FMNMX) per two FFMA2, as in the force loop
    // MODE 6: packed FADD2, two distinct operands
    constexpr int CH = 8;
    if (MODE == 0 || MODE == 2) {
        float x[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) x[c] = seed + threadIdx.x * 1e-6f + c * 0.001f;
        float m = 0.999999f, a = 1e-7f + seed;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int c = 0; c < CH; c++)
                    x[c] = (MODE == 0) ? fmaf(x[c], m, a) : fmaf(x[(c + 1) % CH], x[(c + 2) % CH], x[c]);
        }
        float s = 0;
#pragma unroll
        for (int c = 0; c < CH; c++) s += x[c];
        if (s == 12345.678f) out[threadIdx.x] = s;
    } else {
        u64 x[CH];
        float mn = seed;
#pragma unroll
        for (int c = 0; c < CH; c++) x[c] = pk(seed + threadIdx.x * 1e-6f + c * 0.001f, seed + c * 0.002f + 0.5f);
        u64 m = pk(0.999999f, 0.999998f), a = pk(1e-7f + seed, 2e-7f + seed);
        float sc = 0.99999f + seed * 1e-9f;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    if (MODE == 1) x[c] = fma2(x[c], m, a);
                    if (MODE == 3 || MODE == 5) x[c] = fma2(x[(c + 1) % CH], x[(c + 2) % CH], x[c]);
                    if (MODE == 4) x[c] = fma2(x[(c + 1) % CH], pk(sc, sc), x[c]);
                    if (MODE == 6) x[c] = add2(x[(c + 1) % CH], x[c]);
                    if (MODE == 5 && (c & 1)) {
                        float lo, hi;
                        upk(x[c], lo, hi);
                        mn = fminf(mn, lo);
                    }
                }
        }
        float s = mn;
#pragma unroll
        for (int c = 0; c < CH; c++) {
            float lo, hi;
            upk(x[c], lo, hi);
            s += lo + hi;
        }
        if (s == 12345.678f) out[threadIdx.x] = s;
    }
}

}  // namespace g6b
