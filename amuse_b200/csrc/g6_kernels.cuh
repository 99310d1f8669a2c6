// g6_kernels.cuh -- hand-written sm_100a kernels of the B200 g6 force library.
//
//   predict_kernel          j-particle Hermite predictor                  (HBM-bound)
//   scatter_kernel          j-update scatter                              (HBM/latency-bound)
//   update_predict_kernel   scatter of a small update batch + predictor in one launch
//   pack_i_kernel           double -> double-single i-block packing (device callers)
//   force_fast_kernel<>     Hermite force, big i-blocks: acc, jerk, pot, nearest neighbour; mask-free groups of
//                           pairs verified afterwards (speculative)       (FP32-pipe-bound)
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
// j state in HBM (capacity C, padded to a multiple of TILE), all FP64 like the
// reference's jdata arrays, packed as seven double2 streams (7 x LDG.128 per j):
//   q0 = (x, y)    q1 = (z, t_j)   q2 = (vx, vy)   q3 = (vz, ax)
//   q4 = (ay, az)  q5 = (jx, jy)   q6 = (jz, {float mass, int id})
// predicted j (what the force kernel streams), float4 each:
//   A = (x.hi, y.hi, z.hi, mass)  B = (x.lo, y.lo, z.lo, id bits)  C = (vx, vy, vz, 0)
// ---------------------------------------------------------------------------
struct JState {
    double2 *q[7];
    float4 *A, *B, *C;
};

// One staged j-update (host pinned -> device staging -> scatter_kernel).
struct __align__(16) JUpdate {
    double x[3];
    double v[3];
    double a[3];
    double j[3];
    double t;
    float m;
    int id;
    int addr;
    int pad[3];
};
static_assert(sizeof(JUpdate) == 128, "JUpdate layout");

__device__ __forceinline__ double pack_mass_id(float m, int id)
{
    return __hiloint2double(id, __float_as_int(m));
}

__device__ __forceinline__ void store_update(const JUpdate &u, const JState &s)
{
    const int a = u.addr;
    s.q[0][a] = make_double2(u.x[0], u.x[1]);
    s.q[1][a] = make_double2(u.x[2], u.t);
    s.q[2][a] = make_double2(u.v[0], u.v[1]);
    s.q[3][a] = make_double2(u.v[2], u.a[0]);
    s.q[4][a] = make_double2(u.a[1], u.a[2]);
    s.q[5][a] = make_double2(u.j[0], u.j[1]);
    s.q[6][a] = make_double2(u.j[2], pack_mass_id(u.m, u.id));
}
__global__ void scatter_kernel(int n, const JUpdate *__restrict__ up, JState s)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    store_update(up[k], s);
}

// Hermite predictor (jdata.cc:726-747) in FP64, output split to double-single.
// Algorithmic traffic: 112 B read + 48 B written per j.
// One CTA = one j-tile (TILE = 256 slots; the arrays are padded to whole tiles).  The CTA also
// records the id range of its massive particles in the spare .w lanes of the tile's first two C
// entries (C[tile*256].w = min id, C[tile*256+1].w = max id, as int bits): the force kernel uses
// it to decide per tile whether the self-exclusion masks can be skipped.
// predict_tile: the body, for one tile and the 256 threads of a CTA.  sh_lo/sh_hi: TILE/32 ints of
// shared memory each.
__device__ __forceinline__ void predict_tile(const int tile, const int n, const double ti, const JState &s, int *sh_lo,
                                             int *sh_hi)
{
    const int j = tile * TILE + threadIdx.x;   // always < capacity
    const double2 q0 = s.q[0][j], q1 = s.q[1][j], q2 = s.q[2][j], q3 = s.q[3][j], q4 = s.q[4][j], q5 = s.q[5][j],
                  q6 = s.q[6][j];
    const double x = q0.x, y = q0.y, z = q1.x, tj = q1.y, vx = q2.x, vy = q2.y, vz = q3.x;
    const double ax = q3.y, ay = q4.x, az = q4.y, jx = q5.x, jy = q5.y, jz = q6.x;
    const float m = __int_as_float(__double2loint(q6.y));
    const int id = __double2hiint(q6.y);
    const double dt = ti - tj;
    double px = x, py = y, pz = z, qx = vx, qy = vy, qz = vz;
    if (dt != 0.0) {  // same expression tree as jdata.cc:739-746
        px = x + dt * (vx + 0.5 * dt * (ax + dt * jx / 3));
        py = y + dt * (vy + 0.5 * dt * (ay + dt * jy / 3));
        pz = z + dt * (vz + 0.5 * dt * (az + dt * jz / 3));
        qx = vx + dt * (ax + 0.5 * dt * jx);
        qy = vy + dt * (ay + 0.5 * dt * jy);
        qz = vz + dt * (az + 0.5 * dt * jz);
    }
    const bool massive = (j < n) && (m > TINYF);
    if (!massive) {  // massless or never-set slot: park it (idata.cc:208)
        px = py = pz = (double)FAR_AWAY;
        qx = qy = qz = 0.0;
    }
    int lo = __reduce_min_sync(0xffffffffu, massive ? id : 0x7fffffff);
    int hi = __reduce_max_sync(0xffffffffu, massive ? id : (int)0x80000000);
    if ((threadIdx.x & 31) == 0) {
        sh_lo[threadIdx.x >> 5] = lo;
        sh_hi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    float cw = 0.f;
    if (threadIdx.x < 2) {
#pragma unroll
        for (int w = 0; w < TILE / 32; w++) {
            lo = min(lo, sh_lo[w]);
            hi = max(hi, sh_hi[w]);
        }
        cw = __int_as_float(threadIdx.x == 0 ? lo : hi);
    }
    // slots in [n, end of tile) are written too (parked): the tile's id range lives in its first two
    // entries and the force kernel bounds its j loop by nj anyway
    float xh = (float)px, yh = (float)py, zh = (float)pz;
    float xl = (float)(px - (double)xh), yl = (float)(py - (double)yh), zl = (float)(pz - (double)zh);
    s.A[j] = make_float4(xh, yh, zh, massive ? m : 0.f);
    s.B[j] = make_float4(xl, yl, zl, __int_as_float(id));
    s.C[j] = make_float4((float)qx, (float)qy, (float)qz, cw);
    __syncthreads();   // sh_lo/sh_hi may be reused by the next tile
}

// Pending j-updates applied by the kernel that predicts (small batches, block-timestep regime): the
// addresses ride in the kernel parameters, the 128-byte records are read from mapped pinned host
// memory by the threads whose address falls into the CTA's j range [j_lo, j_hi).
constexpr int UPD_MAX = 256;
struct InlineU {
    int n;
    int addr[UPD_MAX];
};
__device__ __forceinline__ void apply_updates(const InlineU &iu, const JUpdate *__restrict__ rec, const JState &s,
                                              const int j_lo, const int j_hi)
{
    for (int k = threadIdx.x; k < iu.n; k += blockDim.x) {
        const int a = iu.addr[k];
        if (a >= j_lo && a < j_hi) store_update(rec[k], s);
    }
    __syncthreads();   // the block's own global writes are visible to its threads after the barrier
}

__global__ void __launch_bounds__(TILE) predict_kernel(int n, double ti, JState s, int tile0 = 0)
{
    __shared__ int sh_lo[TILE / 32], sh_hi[TILE / 32];
    predict_tile(tile0 + blockIdx.x, n, ti, s, sh_lo, sh_hi);
}

// scatter + predict in one launch (small update batches).
__global__ void __launch_bounds__(TILE) update_predict_kernel(int n, double ti, JState s, const JUpdate *rec,
                                                              const __grid_constant__ InlineU iu)
{
    __shared__ int sh_lo[TILE / 32], sh_hi[TILE / 32];
    apply_updates(iu, rec, s, blockIdx.x * TILE, (blockIdx.x + 1) * TILE);
    predict_tile(blockIdx.x, n, ti, s, sh_lo, sh_hi);
}

// i-block packing for device-resident callers: double -> double-single.
//   iA = (x.hi,y.hi,z.hi,h2)  iB = (x.lo,y.lo,z.lo,id bits)  iC = (vx,vy,vz,0)
__global__ void pack_i_kernel(int ni, const int *__restrict__ index, const double *__restrict__ xi,
                              const double *__restrict__ vi, const double *__restrict__ h2,
                              float4 *iA, float4 *iB, float4 *iC)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ni) return;
    double x = xi[3 * i], y = xi[3 * i + 1], z = xi[3 * i + 2];
    float xh = (float)x, yh = (float)y, zh = (float)z;
    iA[i] = make_float4(xh, yh, zh, h2 ? (float)h2[i] : 0.f);
    iB[i] = make_float4((float)(x - (double)xh), (float)(y - (double)yh), (float)(z - (double)zh),
                        __int_as_float(index[i]));
    iC[i] = make_float4((float)vi[3 * i], (float)vi[3 * i + 1], (float)vi[3 * i + 2], 0.f);
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
// Force kernel.
// ---------------------------------------------------------------------------
// Arguments of the device-resident Hermite step (see "Device-resident Hermite block step" below).
struct HermiteArgs {
    int ni;
    const int *ilist;          // j-addresses of the active particles (mapped pinned host memory or device)
    const double *old_dt;      // their current time steps (same memory); unused by the init pass
    int *ilist_d;              // device copies made by the gather pass, read by the corrector
    double *olddt_d;
    double tnext, eta;
    JState js;
    float4 *iA, *iB, *iC;      // packed i-block for the force kernels
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
    const float4 *jA, *jB, *jC;   // predicted j (device)
    const float4 *iA, *iB, *iC;   // packed i-block (device)
    int ni, nj;                   // i count, j prefix [0, nj)
    int tiles_per_split, nsplit;  // j decomposition over blockIdx.x
    int ni_pad;                   // stride of the partial workspace
    int j_offset;                 // global address of local address 0
    int defer_reduce;             // 1: stop after the per-split partials (reduce_partials_kernel follows)
    float eps2;
    double *part_sum;             // [nsplit][ni_pad][7]
    u64 *part_key;                // [nsplit][ni_pad]
    unsigned int *tickets;        // [gridDim.y], zero between launches
    double *out_sum;              // [ni][7]: acc xyz, jerk xyz, +sum m/r
    u64 *out_key;                 // [ni]
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
    // multi-GPU exchange fused into the force kernel: whoever writes final outputs of this rank's
    // j-shard also stores them into its slot of every peer's exchange buffer over NVLink (peer pointers
    // from CUDA IPC), so the partials travel while the other i-blocks are still being computed
    // device-resident Hermite step, small blocks: whoever writes particle i's final force also runs its
    // corrector (HERM kernels), so a block step is two launches (predict+gather, force+correct)
    HermiteArgs herm;
    int n_mirror;
    double *m_sum[MAX_PEERS];            // [ni][7] at each peer, already offset to this launch's first i
    u64 *m_key[MAX_PEERS];
    int *m_id[MAX_PEERS];
};

// i-block carried in the kernel parameters (constant bank) for small i-blocks: no H2D copy at all.
// Layout [3][N]: iA, iB, iC.  N == 0 is a 48-byte dummy.
template <int N>
struct InlineI {
    float4 d[3 * (N > 0 ? N : 1)];
};

__device__ __forceinline__ void hermite_correct_one(const HermiteArgs &h, const int i, const double *f, const int nnid);

// Final outputs of particle i (local arrays + the peers' exchange slots).
template <bool NN, bool HERM = false>
__device__ __forceinline__ void store_outputs(const ForceArgs &p, const int i, const double *tot, const u64 kk)
{
    int id = -1;
    if (NN && kk != KEY_NONE) id = __float_as_int(p.jB[(int)(unsigned)(kk & 0xffffffffu) - p.j_offset].w);
    if (HERM) {   // the particle's force is complete: correct it right here
        hermite_correct_one(p.herm, i, tot, id);
    } else {
#pragma unroll
        for (int q = 0; q < 7; q++) p.out_sum[(size_t)i * 7 + q] = tot[q];
        p.out_key[i] = kk;
        if (NN) p.out_nnid[i] = id;
        for (int m = 0; m < p.n_mirror; m++) {
#pragma unroll
            for (int q = 0; q < 7; q++) p.m_sum[m][(size_t)i * 7 + q] = tot[q];
            p.m_key[m][i] = kk;
            p.m_id[m][i] = id;
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

struct Acc7 {
    float ax, ay, az, jx, jy, jz, pot;
};

// One (i, j) interaction in scalar FP32 with double-single positions.
// 9 FADD (DS dx) + 3 FADD (dv) + 6 FMUL/FFMA (r2, xv) + 1 FADD (eps2) + MUFU.RSQ
// + 5 FMUL + 9 FFMA + 1 FADD, plus guards.  Counted as 60 flop by convention
// (src/amuse_ph4/src/jdata.cc:1038).
template <bool NN, bool LIST, bool NR>
__device__ __forceinline__ void interact(const float4 a, const float4 b, const float4 c, int jaddr, float xh,
                                         float yh, float zh, float xl, float yl, float zl, float vx, float vy,
                                         float vz, int iid, float h2, float eps2, Acc7 &s, float &r2min,
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
    float rinv = NR ? rsqrt_refined(r2 + eps2) : rsqrt_approx(r2 + eps2);   // eps2 holds eps2 + TINY
    rinv = idok ? rinv : 0.f;
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
    s.pot += ok ? mrinv : 0.f;
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

template <bool NN, bool LIST, bool NR, bool TRACKJ = true>
__device__ __forceinline__ void interact2(const float4 a, const float4 b, const float4 c, int jaddr, const IPair &I,
                                          u64 eps2p, Acc7P &s, float &r2min0, int &jmin0, float &r2min1,
                                          int &jmin1, int i_global0, const ForceArgs &p)
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
    float ri0 = rsqrt_approx(e0);
    float ri1 = rsqrt_approx(e1);
    u64 rinv = pk(ri0, ri1);
    if (NR) {  // packed Newton step: e = r2e*y0^2 - 1 ; y = y0 - (y0/2)*e   (4 packed ops)
        u64 e = fma2(r2e, mul2(rinv, rinv), pk(-1.f, -1.f));
        rinv = fma2(mul2(rinv, pk(-0.5f, -0.5f)), e, rinv);
        upk(rinv, ri0, ri1);
    }
    rinv = pk(id0 ? ri0 : 0.f, id1 ? ri1 : 0.f);
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
    s.pot = fma2(pk(a.w, a.w), pk(ok0 ? ri0 : 0.f, ok1 ? ri1 : 0.f), s.pot);
    if (NN) {
        float n0 = ok0 ? r20 : __int_as_float(0x7f800000);
        float n1 = ok1 ? r21 : __int_as_float(0x7f800000);
        if (!TRACKJ) {   // the speculative kernel keeps only the minimum and finds j afterwards
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
    uint64_t full[STAGES];
    unsigned int is_last;
};

__device__ __forceinline__ u64 make_key(float r2min, int jmin_global)
{
    return ((u64)(unsigned)__float_as_int(r2min) << 32) | (u64)(unsigned)jmin_global;
}

// Thread layout: tid = jslot * NI_SLOTS + islot.  A CTA owns IB = NI_SLOTS*IPT
// i-particles (i = blockIdx.y*IB + islot + k*NI_SLOTS, coalesced) and the j-tiles
// of split blockIdx.x; inside a tile the NJ_SLOTS = THREADS/NI_SLOTS j-slots take
// interleaved j.  i-particles stay in registers for the whole kernel; j tiles
// arrive by TMA bulk copies (3 per stage) signalled on an mbarrier; per-tile FP32
// partial sums are flushed to FP64 so that long sums keep ~1e-7 accuracy.
// Partials of the j-slots are reduced with warp shuffles + shared memory, and the
// partials of the j-splits by the last CTA to arrive (ticket), in fixed order.
template <int IPT, int NI_SLOTS, bool NN, bool LIST, bool PACKED, bool NR, int MINB, int INL, bool HERM = false>
__global__ void __launch_bounds__(THREADS, MINB) force_kernel(const ForceArgs p,
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
    const int tile0 = blockIdx.x * p.tiles_per_split;
    int ntiles = ntiles_total - tile0;
    if (ntiles > p.tiles_per_split) ntiles = p.tiles_per_split;
    if (ntiles < 0) ntiles = 0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    constexpr uint32_t STAGE_BYTES = 3u * TILE * sizeof(float4);
    if (tid == 0) {
        for (int s = 0; s < STAGES && s < ntiles; s++) {
            size_t off = (size_t)(tile0 + s) * TILE;
            mbar_expect_tx(&sm.full[s], STAGE_BYTES);
            bulk_g2s(sm.A[s], p.jA + off, TILE * sizeof(float4), &sm.full[s]);
            bulk_g2s(sm.B[s], p.jB + off, TILE * sizeof(float4), &sm.full[s]);
            bulk_g2s(sm.C[s], p.jC + off, TILE * sizeof(float4), &sm.full[s]);
        }
    }

    // ---- register-resident i-particles -----------------------------------
    float xh[IPT], yh[IPT], zh[IPT], xl[IPT], yl[IPT], zl[IPT], vx[IPT], vy[IPT], vz[IPT], h2[IPT];
    int iid[IPT];
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        int i = i_base + k * NI_SLOTS;
        if (PACKED) i = blockIdx.y * IB + (islot * 2 + (k & 1)) + (k >> 1) * (2 * NI_SLOTS);
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
    }
    auto i_of = [&](int k) -> int {
        return PACKED ? (int)(blockIdx.y * IB + (islot * 2 + (k & 1)) + (k >> 1) * (2 * NI_SLOTS)) : i_base + k * NI_SLOTS;
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
        const int jtile = (tile0 + t) * TILE;
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
                auto do_j = [&](int jj) {
                    const float4 a = tA[jj], b = tB[jj], c = tC[jj];
                    const int jaddr = jtile + jj;
#pragma unroll
                    for (int k = 0; k < IPT; k++)
                        interact<NN, LIST, NR>(a, b, c, jaddr, xh[k], yh[k], zh[k], xl[k], yl[k], zl[k], vx[k], vy[k],
                                               vz[k], iid[k], h2[k], eps2, S[k], r2min[k], jmin[k], i_of(k), p);
                };
                if (whole) {
#pragma unroll UNROLL
                    for (int u = 0; u < FL; u++) do_j(jj0 + u * NJ_SLOTS);
                } else {
                    for (int jj = jj0; jj < cnt; jj += NJ_SLOTS) do_j(jj);
                }
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
                auto do_j = [&](int jj) {
                    const float4 a = tA[jj], b = tB[jj], c = tC[jj];
                    const int jaddr = jtile + jj;
#pragma unroll
                    for (int q = 0; q < NP; q++)
                        interact2<NN, LIST, NR>(a, b, c, jaddr, IP[q], eps2p, S[q], r2min[2 * q], jmin[2 * q],
                                                r2min[2 * q + 1], jmin[2 * q + 1], i_of(2 * q), p);
                };
                if (whole) {
#pragma unroll UNROLL
                    for (int u = 0; u < FL; u++) do_j(jj0 + u * NJ_SLOTS);
                } else {
                    for (int jj = jj0; jj < cnt; jj += NJ_SLOTS) do_j(jj);
                }
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
            size_t off = (size_t)(tile0 + t + STAGES) * TILE;
            mbar_expect_tx(&sm.full[s], STAGE_BYTES);
            bulk_g2s(sm.A[s], p.jA + off, TILE * sizeof(float4), &sm.full[s]);
            bulk_g2s(sm.B[s], p.jB + off, TILE * sizeof(float4), &sm.full[s]);
            bulk_g2s(sm.C[s], p.jC + off, TILE * sizeof(float4), &sm.full[s]);
        }
    }

    // ---- keys ---------------------------------------------------------------
    u64 key[IPT];
#pragma unroll
    for (int k = 0; k < IPT; k++)
        key[k] = (jmin[k] >= 0) ? make_key(r2min[k], jmin[k] + p.j_offset) : KEY_NONE;

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
// Speculative force kernel (large i-blocks, packed FP32).
//
// The masks of the exact pair rule (skip equal ids; pot and the neighbour search only for
// r2 > 2^-52) cost ~18 ALU-pipe instructions per packed pair and, on this chip, compete with the
// FMA pipe for register-file read bandwidth (profiles/: FFMA2 with three distinct register
// operands runs at 2/3 rate).  Almost no pair needs them, so the j stream is processed in groups
// of GRP pairs WITHOUT masks, and each group is verified afterwards:
//   * an equal-id pair can only occur in a tile whose id range [C[0].w, C[1].w] (written by
//     predict_kernel) contains the id of one of the warp's i-particles -> such tiles take the
//     masked path from the start;
//   * a pair with r2 <= 2^-52 shows up in the running minimum of r2, which is tracked anyway for
//     the nearest neighbour -> the group's FP32 partial sums are discarded and the group is redone
//     with the masked pair function.
// Results are therefore those of the masked rule for every input.  The nearest neighbour is kept
// as (min r2, first group that lowered it) with one 3-input FMNMX per i per two j; the exact j is
// found at the end by re-scanning that one group with the same r2 instruction sequence.
// ---------------------------------------------------------------------------
#ifndef G6_GRP
#define G6_GRP 32
#endif
#ifndef G6_FUNROLL
#define G6_FUNROLL 2
#endif
#ifndef G6_DUAL
#define G6_DUAL 0   // 0: one FP32 partial-sum set per group; k: two sets (even/odd j) in kernels with IPT >= k
#endif
constexpr int GRP = G6_GRP;   // pairs per speculation/flush group (FP32 partial sums span one group)
constexpr int FUNROLL = G6_FUNROLL;  // j-pairs unrolled in the mask-free loop

__device__ __forceinline__ float min3f(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// Geometry shared by the fast path, the masked path and the neighbour re-scan: identical
// instruction sequence, so r2 is bit-identical in all three.
__device__ __forceinline__ void pair_geometry(const float4 a, const float4 b, const IPair &I, u64 &dx, u64 &dy, u64 &dz,
                                              u64 &r2)
{
    dx = add2(add2(pk(a.x, a.x), I.nxh), add2(pk(b.x, b.x), I.nxl));
    dy = add2(add2(pk(a.y, a.y), I.nyh), add2(pk(b.y, b.y), I.nyl));
    dz = add2(add2(pk(a.z, a.z), I.nzh), add2(pk(b.z, b.z), I.nzl));
    r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
}

// One j against an i-pair, no masks: 38 packed FP32 operations + 2 MUFU.  Accumulates -acc, -jerk, +pot.
// EPS0: the caller runs unsoftened (eps2 = 0, ph4's AMUSE default).  The reference still adds 2^-52 to r2
// (idata.cc:216); in FP32 that add is the identity for r2 > 2^-26, so it is skipped here and the group
// verification (running minimum of r2) redoes any group that holds a closer pair with the exact path.
template <bool NR, bool EPS0>
__device__ __forceinline__ u64 interact2_fast(const float4 a, const float4 b, const float4 c, const IPair &I,
                                              const u64 eps2p, Acc7P &s)
{
    u64 dx, dy, dz, r2;
    pair_geometry(a, b, I, dx, dy, dz, r2);
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

template <int IPT, bool NN, bool NR, int MINB, bool EPS0>
__global__ void __launch_bounds__(THREADS, MINB) force_fast_kernel(const ForceArgs p)
{
    // a pair this close sends its group down the exact path (coincident pairs; with EPS0 also pairs for
    // which r2 + 2^-52 is not r2 in FP32)
    constexpr float R2_EXACT = EPS0 ? 1.4901161193847656e-08f /* 2^-26 */ : TINYF;
    constexpr bool DUAL = (G6_DUAL != 0) && (IPT >= G6_DUAL);
    static_assert(IPT % 2 == 0, "i-particles are processed in packed pairs");
    static_assert(TILE % GRP == 0 && GRP % (2 * G6_FUNROLL) == 0, "group shape");
    constexpr int NP = IPT / 2;
    constexpr int IB = THREADS * IPT;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    ForceSmem &sm = *reinterpret_cast<ForceSmem *>(smem_raw);
    const int tid = threadIdx.x;

    const int ntiles_total = (p.nj + TILE - 1) / TILE;
    const int tile0 = blockIdx.x * p.tiles_per_split;
    int ntiles = ntiles_total - tile0;
    if (ntiles > p.tiles_per_split) ntiles = p.tiles_per_split;
    if (ntiles < 0) ntiles = 0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    constexpr uint32_t STAGE_BYTES = 3u * TILE * sizeof(float4);
    if (tid == 0) {
        for (int s = 0; s < STAGES && s < ntiles; s++) {
            size_t off = (size_t)(tile0 + s) * TILE;
            mbar_expect_tx(&sm.full[s], STAGE_BYTES);
            bulk_g2s(sm.A[s], p.jA + off, TILE * sizeof(float4), &sm.full[s]);
            bulk_g2s(sm.B[s], p.jB + off, TILE * sizeof(float4), &sm.full[s]);
            bulk_g2s(sm.C[s], p.jC + off, TILE * sizeof(float4), &sm.full[s]);
        }
    }

    // ---- register-resident i-pairs: i = block base + 2*tid + {0,1} + q*2*THREADS -------------
    IPair IP[NP];
    int iid[IPT];
    auto i_of = [&](int k) -> int { return blockIdx.y * IB + (tid * 2 + (k & 1)) + (k >> 1) * (2 * THREADS); };
#pragma unroll
    for (int q = 0; q < NP; q++) {
        float4 a[2], b[2], c[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int i = i_of(2 * q + h);
            a[h] = make_float4(0.f, 0.f, 0.f, -1.f);
            b[h] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x80000000));
            c[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < p.ni) {
                a[h] = p.iA[i];
                b[h] = p.iB[i];
                c[h] = p.iC[i];
            }
            iid[2 * q + h] = __float_as_int(b[h].w);
        }
        IP[q].nxh = pk(-a[0].x, -a[1].x); IP[q].nyh = pk(-a[0].y, -a[1].y); IP[q].nzh = pk(-a[0].z, -a[1].z);
        IP[q].nxl = pk(-b[0].x, -b[1].x); IP[q].nyl = pk(-b[0].y, -b[1].y); IP[q].nzl = pk(-b[0].z, -b[1].z);
        IP[q].nvx = pk(-c[0].x, -c[1].x); IP[q].nvy = pk(-c[0].y, -c[1].y); IP[q].nvz = pk(-c[0].z, -c[1].z);
        IP[q].id0 = iid[2 * q]; IP[q].id1 = iid[2 * q + 1];
        IP[q].h20 = a[0].w; IP[q].h21 = a[1].w;
    }

    double D[IPT][7];
    float rmin[IPT], rprev[IPT];   // minimum of r2 over the current group; running minimum (masked rule) before it
    int jgrp[IPT];                 // first j of the group that last lowered it (local address), -1: none
#pragma unroll
    for (int k = 0; k < IPT; k++) {
#pragma unroll
        for (int q = 0; q < 7; q++) D[k][q] = 0.0;
        rmin[k] = rprev[k] = __int_as_float(0x7f800000);
        jgrp[k] = -1;
    }
    const float eps2 = p.eps2 + TINYF;   // the reference softens by eps2 + 2^-52 (idata.cc:216)
    const u64 eps2p = pk(eps2, eps2);

    for (int t = 0; t < ntiles; t++) {
        const int s = t % STAGES;
        const uint32_t phase = (uint32_t)(t / STAGES) & 1u;
        while (!mbar_try_wait(&sm.full[s], phase)) {
        }
        const int jtile = (tile0 + t) * TILE;
        int cnt = p.nj - jtile;
        if (cnt > TILE) cnt = TILE;
        const float4 *tA = sm.A[s], *tB = sm.B[s], *tC = sm.C[s];

        // may this warp meet an equal-id pair in this tile?
        const int idlo = __float_as_int(tC[0].w), idhi = __float_as_int(tC[1].w);
        bool hit = false;
#pragma unroll
        for (int k = 0; k < IPT; k++) hit |= (iid[k] >= idlo) & (iid[k] <= idhi);
        const bool tile_masked = __any_sync(0xffffffffu, hit) || (cnt < TILE);

        for (int jj0 = 0; jj0 < cnt; jj0 += GRP) {
            Acc7P S[NP];
            bool fast = !tile_masked;
#pragma unroll
            for (int k = 0; k < IPT; k++) rmin[k] = __int_as_float(0x7f800000);
            if (fast) {
                // DUAL: even and odd j of the group go to separate FP32 partial sums (each spans GRP/2
                // pairs, which is what bounds the rounding error of a sum dominated by one close pair)
                Acc7P S1[DUAL ? NP : 1];
#pragma unroll
                for (int q = 0; q < NP; q++) {
                    S[q] = Acc7P{0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
                    if (DUAL) S1[q] = Acc7P{0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
                }
#pragma unroll FUNROLL
                for (int u = 0; u < GRP; u += 2) {
                    const int jj = jj0 + u;
                    const float4 a0 = tA[jj], b0 = tB[jj], c0 = tC[jj];
                    const float4 a1 = tA[jj + 1], b1 = tB[jj + 1], c1 = tC[jj + 1];
#pragma unroll
                    for (int q = 0; q < NP; q++) {
                        const u64 ra = interact2_fast<NR, EPS0>(a0, b0, c0, IP[q], eps2p, S[q]);
                        const u64 rb = interact2_fast<NR, EPS0>(a1, b1, c1, IP[q], eps2p, DUAL ? S1[q] : S[q]);
                        float ra0, ra1, rb0, rb1;
                        upk(ra, ra0, ra1);
                        upk(rb, rb0, rb1);
                        rmin[2 * q] = min3f(rmin[2 * q], ra0, rb0);
                        rmin[2 * q + 1] = min3f(rmin[2 * q + 1], ra1, rb1);
                    }
                }
                if (DUAL) {
#pragma unroll
                    for (int q = 0; q < NP; q++) {
                        S[q].ax = add2(S[q].ax, S1[q].ax); S[q].ay = add2(S[q].ay, S1[q].ay);
                        S[q].az = add2(S[q].az, S1[q].az); S[q].jx = add2(S[q].jx, S1[q].jx);
                        S[q].jy = add2(S[q].jy, S1[q].jy); S[q].jz = add2(S[q].jz, S1[q].jz);
                        S[q].pot = add2(S[q].pot, S1[q].pot);
                    }
                }
                bool bad = false;
#pragma unroll
                for (int k = 0; k < IPT; k++) bad |= !(rmin[k] > R2_EXACT);
                if (__any_sync(0xffffffffu, bad)) {   // a pair too close for the fast path: redo the group exactly
                    fast = false;
#pragma unroll
                    for (int k = 0; k < IPT; k++) rmin[k] = __int_as_float(0x7f800000);
                }
            }
            if (!fast) {
#pragma unroll
                for (int q = 0; q < NP; q++) S[q] = Acc7P{0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
                const int jend = (jj0 + GRP < cnt) ? jj0 + GRP : cnt;
                for (int jj = jj0; jj < jend; jj++) {
                    const float4 a = tA[jj], b = tB[jj], c = tC[jj];
                    int unused0 = 0, unused1 = 0;
#pragma unroll
                    for (int q = 0; q < NP; q++)
                        interact2<true, false, NR, false>(a, b, c, 0, IP[q], eps2p, S[q], rmin[2 * q], unused0,
                                                          rmin[2 * q + 1], unused1, 0, p);
                }
            }
            const double sg = fast ? -1.0 : 1.0;   // the mask-free pair function accumulates -acc, -jerk
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
            if (NN) {
#pragma unroll
                for (int k = 0; k < IPT; k++) {
                    if (rmin[k] < rprev[k]) {   // strict: the first group that reaches the minimum keeps it
                        jgrp[k] = jtile + jj0;
                        rprev[k] = rmin[k];
                    }
                }
            }
        }

        __syncthreads();  // everyone is done with stage s
        if (tid == 0 && t + STAGES < ntiles) {
            size_t off = (size_t)(tile0 + t + STAGES) * TILE;
            mbar_expect_tx(&sm.full[s], STAGE_BYTES);
            bulk_g2s(sm.A[s], p.jA + off, TILE * sizeof(float4), &sm.full[s]);
            bulk_g2s(sm.B[s], p.jB + off, TILE * sizeof(float4), &sm.full[s]);
            bulk_g2s(sm.C[s], p.jC + off, TILE * sizeof(float4), &sm.full[s]);
        }
    }

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
        key[k] = (jm >= 0) ? make_key(rprev[k], jm + p.j_offset) : KEY_NONE;
    }

    // ---- totals -> global (final or per-split partial) -----------------------------------
    const bool single = (p.nsplit == 1);
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

// After a min-reduction of keys over ranks (each rank holds a j-shard).
__global__ void resolve_nn_kernel(int ni, const u64 *__restrict__ key, int rank, int j_offset, int nj_local,
                                  const float4 *__restrict__ jB, int *nnid)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ni) return;
    u64 k = key[i];
    int id = 0;
    if (k == KEY_NONE) {
        id = (rank == 0) ? -1 : 0;
    } else {
        int a = (int)(unsigned)(k & 0xffffffffu) - j_offset;
        if (a >= 0 && a < nj_local) id = __float_as_int(jB[a].w);
    }
    nnid[i] = id;
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
    const int a = h.ilist[i];
    const JState &s = h.js;
    const double2 q0 = s.q[0][a], q1 = s.q[1][a], q2 = s.q[2][a], q3 = s.q[3][a], q4 = s.q[4][a], q5 = s.q[5][a],
                  q6 = s.q[6][a];
    const double x = q0.x, y = q0.y, z = q1.x, tj = q1.y, vx = q2.x, vy = q2.y, vz = q3.x;
    const double ax = q3.y, ay = q4.x, az = q4.y, jx = q5.x, jy = q5.y, jz = q6.x;
    const int id = __double2hiint(q6.y);
    const double dt = h.tnext - tj;
    double px = x, py = y, pz = z, qx = vx, qy = vy, qz = vz;
    if (dt != 0.0) {  // idata.cc:353-361
        px = x + dt * (vx + 0.5 * dt * (ax + dt * jx / 3));
        py = y + dt * (vy + 0.5 * dt * (ay + dt * jy / 3));
        pz = z + dt * (vz + 0.5 * dt * (az + dt * jz / 3));
        qx = vx + dt * (ax + 0.5 * dt * jx);
        qy = vy + dt * (ay + 0.5 * dt * jy);
        qz = vz + dt * (az + 0.5 * dt * jz);
    }
    double *pr = h.pred + (size_t)i * 6;
    pr[0] = px; pr[1] = py; pr[2] = pz; pr[3] = qx; pr[4] = qy; pr[5] = qz;
    h.ilist_d[i] = a;
    h.olddt_d[i] = (h.mode == 0) ? h.old_dt[i] : 0.0;
    const float xh = (float)px, yh = (float)py, zh = (float)pz;
    h.iA[i] = make_float4(xh, yh, zh, 0.f);
    h.iB[i] = make_float4((float)(px - (double)xh), (float)(py - (double)yh), (float)(pz - (double)zh),
                          __int_as_float(id));
    h.iC[i] = make_float4((float)qx, (float)qy, (float)qz, 0.f);
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
        while (fmod(h.tnext, oldstep2) != 0) oldstep2 /= 2;
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
        while (fmod(h.tnext, first) != 0) first /= 2;
        while (first > limit) first /= 2;
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
    __shared__ int sh_lo[TILE / 32], sh_hi[TILE / 32];
    if ((int)blockIdx.x < ntiles) {
        predict_tile(blockIdx.x, n, ti, h.js, sh_lo, sh_hi);
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
__global__ void __launch_bounds__(256) peer_combine_kernel(const PeerSlots ps, const unsigned long long seq, const int ni,
                                                          double *out_sum, u64 *out_key, int *out_nnid,
                                                          unsigned int *error_word)
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
