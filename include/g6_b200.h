/*
 * g6_b200.h -- C ABI of the B200-native GRAPE-6/Sapporo force library.
 *
 * The library (libsapporo.so, alias libg6.so) is a drop-in for the two g6
 * implementations AMUSE ships:
 *     lib/sapporo_light/sapporoG6lib.cpp:5-81   (CUDA, 2008-era)
 *     lib/g6lib/g6lib.h:8-128, g6lib.c:190-481  (CPU emulation)
 * Callers that bind these symbols: src/amuse_ph4/src/grape.h:6-120 (ph4),
 * src/amuse_phigrape/src/{initgrape,sendbodies2grape,update_grape,gravity}.F
 * (phiGRAPE, Fortran: every argument by reference, trailing underscore),
 * src/amuse_bhtree/src/BHtree.C:763-885.
 *
 * Part 1 declares exactly the symbols those callers link against.
 * Part 2 ("g6x_") is the batched / device-resident extension used by the
 * benchmark, the multi-GPU path and any caller that keeps data in HBM.
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 */
#ifndef G6_B200_H
#define G6_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ======================================================================
 * Part 1 -- the GRAPE-6 ABI (Fortran convention: trailing '_', all pointers)
 * ====================================================================== */

/* Open the device.  *id = CUDA device ordinal (ph4 passes its MPI rank or
 * gpu_id, src/amuse_ph4/src/gpu.cc:56-59).  Returns 0, or -1 if no such
 * device (lib/sapporo_light/sapporo.cpp:19-41).  Re-opening is allowed
 * (phiGRAPE opens/closes around every evolve, interface.F:552-594). */
int g6_open_(int *id);
/* Free all device state (lib/sapporo_light/sapporo.cpp:43-56). */
int g6_close_(int *id);
/* Max i-particles per firsthalf/lasthalf call (sapporo_light: 256,
 * sapporo.h:149; g6lib: 1).  Here 16384 (phiGRAPE's NGP bound,
 * src/amuse_phigrape/src/gravity.F) unless env G6_B200_NPIPES overrides. */
int g6_npipes_(void);
/* Unit setters of the GRAPE hardware; ignored (sapporoG6lib.cpp:9-10).  ph4
 * declares double*, phiGRAPE passes int*: the argument is never read. */
int g6_set_tunit_(void *unused);
int g6_set_xunit_(void *unused);
/* Set the time all j-particles are predicted to by the next force call
 * (sapporo.cpp:58-62). */
int g6_set_ti_(int *id, double *ti);
/* Store j-particle at slot *address (sapporo.cpp:68-112).  a2 = acc/2,
 * j6 = jerk/6, k18 = snap/18 (ignored, as in sapporo.cpp:91-96).  *index is the
 * particle id used for self-exclusion and returned as nearest neighbour.  The
 * update becomes visible to the next g6calc_firsthalf_. */
int g6_set_j_particle_(int *cluster_id, int *address, int *index, double *tj,
                       double *dtj, double *mass, double k18[3], double j6[3],
                       double a2[3], double v[3], double x[3]);
/* Capture the i-block (ni <= g6_npipes_()) and start the device work for it:
 * pending j-updates are uploaded, all j in [0,*nj) are predicted to ti, the
 * i-block is uploaded (sapporo.cpp:114-152).  aold/j6old/phiold are GRAPE
 * scaling hints and are ignored.  h2[i] = squared neighbour radius. */
void g6calc_firsthalf_(int *cluster_id, int *nj, int *ni, int index[],
                       double xi[][3], double vi[][3], double aold[][3],
                       double j6old[][3], double phiold[], double *eps2,
                       double h2[]);
/* Finish the force evaluation of the captured i-block: acc, jerk and pot
 * (negative: -sum m/sqrt(r2+eps2)) against j in [0,*nj)  (sapporo.cpp:154-181). */
int g6calc_lasthalf_(int *cluster_id, int *nj, int *ni, int index[],
                     double xi[][3], double vi[][3], double *eps2, double h2[],
                     double acc[][3], double jerk[][3], double pot[]);
/* Same, and also inn[i] = index (id) of the nearest j by unsoftened distance,
 * self excluded (sapporo.cpp:184-232); -1 if there is none.  Also builds the
 * neighbour-sphere lists (r2 <= h2[i]) when any h2[i] > 0. */
int g6calc_lasthalf2_(int *cluster_id, int *nj, int *ni, int index[],
                      double xi[][3], double vi[][3], double *eps2, double h2[],
                      double acc[][3], double jerk[][3], double pot[],
                      int inn[]);
/* GRAPE buffer management: no-ops (sapporoG6lib.cpp:52-55). */
int g6_initialize_jp_buffer_(int *cluster_id, int *buf_size);
int g6_flush_jp_buffer_(int *cluster_id);
int g6_reset_(int *cluster_id);
int g6_reset_fofpga_(int *cluster_id);
/* Fetch the neighbour lists of the last lasthalf2 to the host.  Returns
 * non-zero iff any list overflowed the per-particle capacity
 * (sapporo.cpp:234-246; capacity there 256, here G6_B200_NGB_CAP, default 1024). */
int g6_read_neighbour_list_(int *cluster_id);
/* Copy the list of i-particle *ipipe: ids with r2 <= h2, self excluded, sorted
 * ascending; *n_neighbours = full count; returns non-zero iff the count does
 * not fit in *maxlength (sapporo.cpp:248-272). */
int g6_get_neighbour_list_(int *cluster_id, int *ipipe, int *maxlength,
                           int *n_neighbours, int neighbour_list[]);
/* Number of CUDA devices; the symbol AMUSE's configure probes for
 * (support/shared/m4/amuse_lib.m4:135-137; lib/sapporo_light/send_fetch_data.cpp). */
int get_device_count(void);

/* By-value variants exported by lib/g6lib (g6lib.h:58-128) and probed by
 * AC_SEARCH_LIBS(g6_npipes, g6) (amuse_lib.m4:126-128).  ph4 defines its own
 * copies in grape.cc:7-146; the executable's definition wins. */
int g6_open(int clusterid);
int g6_close(int clusterid);
int g6_npipes(void);
int g6_set_tunit(int newtunit);
int g6_set_xunit(int newxunit);
int g6_set_ti(int clusterid, double ti);
int g6_set_j_particle(int clusterid, int address, int index, double tj,
                      double dtj, double mass, double a2by18[3],
                      double a1by6[3], double aby2[3], double v[3], double x[3]);
void g6calc_firsthalf(int clusterid, int nj, int ni, int index[],
                      double xi[][3], double vi[][3], double fold[][3],
                      double jold[][3], double phiold[], double eps2,
                      double h2[]);
int g6calc_lasthalf(int clusterid, int nj, int ni, int index[], double xi[][3],
                    double vi[][3], double eps2, double h2[], double acc[][3],
                    double jerk[][3], double pot[]);
int g6calc_lasthalf2(int clusterid, int nj, int ni, int index[],
                     double xi[][3], double vi[][3], double eps2, double h2[],
                     double acc[][3], double jerk[][3], double pot[],
                     int nnbindex[]);
int g6_initialize_jp_buffer(int clusterid, int size);
int g6_flush_jp_buffer(int clusterid);
void g6_reset(int devid);
int g6_reset_fofpga(int devid);
void g6_reinitialize(int clusterid);
int g6_get_number_of_pipelines(void);
int g6_read_neighbour_list(int clusterid);
int g6_get_neighbour_list(int clusterid, int ipipe, int maxlength, int *nblen,
                          int nbl[]);
void g6_set_neighbour_list_sort_mode(int mode);
int g6_get_neighbour_list_sort_mode(void);
/* Debug hooks ph4 declares (grape.h:110-119). */
int g6_set_overflow_flag_test_mode(int aflag, int jflag, int pflag);
void force_j_particle_send(void);
/* grape.h:118-119 (its only call site is commented out, gpu.cc:423): read back j-particle `addr` of a
 * j-memory of nj particles -- stored pos, vel, acc, jerk and the last predicted pos, vel.  Returns 0, or
 * -1 if addr is outside [0, nj). */
int get_j_part_data(int addr, int nj, double *pos, double *vel, double *acc,
                    double *jrk, double *ppos, double *pvel);

/* ======================================================================
 * Part 2 -- g6x_: batched and device-resident entry points
 * ====================================================================== */

/* Library/ABI version (major*100 + minor). */
int g6x_version(void);
/* external != 0: run all work on the caller's CUDA stream (a cudaStream_t passed
 * as void*; NULL is the legacy default stream).  external == 0: back to the
 * library's own stream. */
int g6x_set_stream(void *cuda_stream, int external);
/* 1 (default): one Newton step on the reciprocal square root (per-pair error at
 * the FP32 rounding floor); 0: raw MUFU.RSQ (2^-22.9), ~10 % faster. */
int g6x_set_refine(int on);
/* Accuracy / speed parameters of the pair classification (defaults: G6_B200_KCLOSE = 32, G6_B200_FARC = 0.125):
 * k_close: pairs closer than sqrt(k_close) x (the i-particle's nearest-neighbour distance) are evaluated
 * in FP64 with the reference's expression tree (0 switches that off: plain FP32 pair arithmetic);
 * far_factor: (warp of i) x (group of j) blocks whose bounding boxes are further apart than far_factor x
 * (largest |coordinate| relative to the library's origin) take position differences from the hi parts of the
 * double-single coordinates alone.  A negative value leaves that parameter unchanged. */
int g6x_set_close_factor(double k_close, double far_factor);
/* How often the j-memory has been re-ordered (Morton sort + id table) since g6_open_. */
long long g6x_order_rebuilds(void);
/* Diagnostics of the pair classification: (warp of i) x (group of j) blocks taken FAR / NEAR / CLOSE and NEAR
 * blocks redone exactly, since the previous call.  Only builds with -DG6_STATS count (returns 0; the first
 * call arms the counters); the production build returns -1. */
int g6x_block_stats(unsigned long long out[8]);   /* [4] = FP64 pairs queued, [5] = appends that met a full list */
/* Devices g6_open_ opened (G6_B200_DEVICES; 0 when closed). */
int g6x_device_count_open(void);
/* This process owns j-addresses whose GLOBAL address is local + offset (used
 * when j is sharded over ranks; the offset is packed in the nearest-neighbour
 * keys so a min-reduction over ranks is meaningful). */
int g6x_set_j_offset(int offset);
/* Array form of g6_set_j_particle_ for n particles (host arrays; k18 omitted).
 * address == NULL means address[k] = address0 + k. */
int g6x_set_j_particles(int n, const int *address, int address0,
                        const int *index, const double *tj, const double *mass,
                        const double (*j6)[3], const double (*a2)[3],
                        const double (*v)[3], const double (*x)[3]);
/* Upload pending j-updates and predict j in [0,nj) to ti now (asynchronous). */
int g6x_predict(int nj, double ti);
/* Device-resident force evaluation, asynchronous on the library stream.
 * Inputs are DEVICE pointers: d_index[ni], d_xi[ni][3], d_vi[ni][3], d_h2[ni]
 * (may be NULL = 0).  Outputs are DEVICE pointers: d_sum[ni][7] =
 * (acc xyz, jerk xyz, +sum m/r), d_key[ni] = (float bits of min r2) << 32 |
 * global j address (0x7f800000ffffffff if none), d_nnid[ni] = id of nearest j
 * or -1 (valid for a single shard; see g6x_resolve_nn).
 * flags: bit0 = nearest neighbour wanted, bit1 = neighbour lists wanted.
 * Runs the predictor first if ti or any j changed.  ni is unlimited (the call
 * loops over npipes-sized chunks).  Returns 0. */
int g6x_calc_device(int nj, int ni, const int *d_index, const double *d_xi,
                    const double *d_vi, const double *d_h2, double eps2,
                    int flags, double *d_sum, unsigned long long *d_key,
                    int *d_nnid);
/* ---- multi-GPU exchange over peer memory (one process per GPU, NVLink/NVSwitch) ----
 * Replaces the five host MPI_Allreduce of idata::get_acc_and_jerk (idata.cc:284-313) by stores into
 * the peers' memory issued from inside the force kernels plus one combine kernel.
 *   1. every rank: g6x_peer_alloc(world, rank, capacity, handle) -- allocates its exchange buffer for
 *      i-sets of up to `capacity` particles and returns an opaque handle (g6x_peer_handle_bytes() bytes,
 *      a CUDA IPC memory handle);
 *   2. the caller gathers the handles of all ranks (any transport: torch.distributed, MPI, a file);
 *   3. every rank: g6x_peer_attach(handles[world]) -- maps the peers' buffers;
 *   4. g6x_calc_device_allreduce(...) -- as g6x_calc_device, but every rank's d_sum/d_key/d_nnid
 *      receive the combination over all j-shards (sum, min key, id of the winner).  Collective: all
 *      ranks must call it in the same order.  g6x_set_j_offset() gives the keys their global address.
 * g6x_peer_error() != 0 after a synchronize means a combine kernel gave up waiting for a peer. */
int g6x_peer_handle_bytes(void);
int g6x_peer_alloc(int world, int rank, int capacity, void *handle_out);
int g6x_peer_attach(const void *handles);
int g6x_peer_detach(void);
int g6x_peer_error(void);
int g6x_calc_device_allreduce(int nj, int ni, const int *d_index,
                              const double *d_xi, const double *d_vi,
                              const double *d_h2, double eps2, int flags,
                              double *d_sum, unsigned long long *d_key,
                              int *d_nnid);
/* Multi-process runs that replicate the j-memory (every rank loads ALL particles, like ph4's MPI ranks; the Morton
 * order is then the same everywhere): this process' device-resident entry points sum over the slots
 * [slot_lo, slot_hi) only (slot_lo a multiple of 256; the windows of all ranks tile [0, nj)).  The neighbour bounds
 * that set the FP64 radius stay global, so no close pair is evaluated twice.  slot_hi <= 0: the whole j-memory. */
int g6x_set_j_window(int slot_lo, int slot_hi);
/* i-particles per kernel launch that g6x_calc_device uses for an i-set of ni against
 * the j currently loaded (it picks the chunk whose i-blocks x j-splits make four full
 * waves of CTAs with ~128 j-tiles each; g6_npipes_() only bounds the ABI path). */
int g6x_device_chunk(int ni);
/* After a min-reduction of d_key over ranks: d_nnid[i] = id of the winning j if
 * this rank owns it, else 0 (so a sum over ranks gives the id), -1 on rank 0
 * if there is no neighbour.  Mirrors idata.cc:308-313. */
int g6x_resolve_nn(int ni, const unsigned long long *d_key, int rank,
                   int *d_nnid);
/* ---- device-resident Hermite block step (the steps either side of the force call) ----
 * ph4's idata::advance (idata.cc:832-870) gathers the active particles from its j-arrays, predicts
 * them, calls the g6 library, corrects them, and sends them back one g6_set_j_particle_ at a time.
 * The active particles ARE j-particles, so the library can do all of that where the state lives:
 *   g6x_hermite_init : forces on all nj particles at t0 (acc, jerk stored in the j-memory) and first
 *                      time steps (jdata::set_initial_timestep, jdata.cc:503-548) -> timestep_out[nj]
 *   g6x_hermite_step : for the ni addresses in ilist: predict to tnext (idata.cc:347-365), force
 *                      against all j predicted to tnext, corrector + Aarseth step with block
 *                      quantisation (idata.cc:443-511), state written back.  Host traffic: ilist and
 *                      old_dt in, new_dt (and optionally pot, nn ids) out.
 *   g6x_hermite_evolve : jdata::advance loop (jdata.cc:752-795), scheduler on the host, to t_end or
 *                      max_block_steps (> 0); returns block steps done, stats[4] = time reached,
 *                      block steps, particle steps (both cumulative), wall seconds of this call.
 *   g6x_hermite_get_state : copy time/pos/vel/acc/jerk of the j-memory back (any pointer may be NULL).
 * Multi-GPU (one process per GPU): every rank loads ALL particles (the state is replicated), attaches
 * the peers (g6x_peer_*) and calls g6x_hermite_set_shard(j_lo, j_hi) with its window of the j-addresses
 * (j_lo a multiple of 256; windows of all ranks tile [0, nj)).  A step then sums the forces over the
 * window only, exchanges the partials over peer memory, and every rank corrects its replica with the
 * identical totals -- no state and no host data ever travel; all ranks must make the same calls.
 * j_hi <= 0 switches back to the single-device step. */
int g6x_hermite_set_shard(int j_lo, int j_hi);
int g6x_hermite_init(int nj, double t0, double eta, double eps2, double *timestep_out);
int g6x_hermite_step(int nj, int ni, const int *ilist, double tnext, double eta,
                     double eps2, const double *old_dt, double *new_dt, double *pot,
                     int *nn);
long long g6x_hermite_evolve(int nj, double t_end, double eta, double eps2,
                             long long max_block_steps, double *stats);
int g6x_hermite_get_state(int nj, double *t, double (*x)[3], double (*v)[3],
                          double (*a)[3], double (*j)[3]);
/* Block until the library stream is idle. */
int g6x_synchronize(void);
/* Cumulative number of CUDA kernels this library has launched. */
long long g6x_launch_count(void);
/* Device pointers of the predicted j arrays (float4 each): A = (x.hi,y.hi,z.hi,m),
 * B = (x.lo,y.lo,z.lo,id bits), C = (vx,vy,vz,0); for tests/inspection. */
int g6x_get_predicted(void **A, void **B, void **C, int *capacity);
/* Copy predicted j state back to the host as doubles (tests): pos = hi+lo. */
int g6x_read_predicted(int nj, double (*pos)[3], double (*vel)[3]);
/* Time only the j-predictor / j-update kernels (CUDA events, ms per launch,
 * averaged over reps) for the HBM roofline of those kernels. */
double g6x_time_predictor(int nj, int reps);
/* Floor of the latency path: microseconds per round trip of `kernels` dependent empty launches
 * whose last one raises the completion flag in mapped host memory (mean over reps). */
double g6x_latency_probe(int kernels, int reps);
/* Which force-kernel variant the next calls use: 0 = auto, else a fixed variant
 * id; for benchmarking and tests. */
int g6x_set_variant(int variant);
/* FP32 pipe microbenchmark: returns achieved TFLOP/s (2 flop per FMA lane) of a
 * dependent-chain FFMA (mode 0) or packed FFMA2 (mode 1) kernel. */
double g6x_fp32_peak(int mode);

#ifdef __cplusplus
}
#endif
#endif /* G6_B200_H */
