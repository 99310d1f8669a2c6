// g6_latency.cc -- TEST INFRASTRUCTURE ONLY (caller harness), not product code.
//
// Latency of one g6 force evaluation as a C caller sees it (no Python in the loop): the call pattern of a
// block time step of ph4 (src/amuse_ph4/src/gpu.cc:163-230,365-407) and phiGRAPE (gravity.F:54-112,
// update_grape.F:19-40): ni x g6_set_j_particle_ (the block advanced last step), g6_set_ti_, g6calc_firsthalf_,
// g6calc_lasthalf2_.  j = uniform random sphere of N equal masses (latency does not depend on the geometry).
// usage: g6_latency <g6lib.so> <N> [reps]
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef int (*open_t)(int *);
typedef int (*npipes_t)(void);
typedef int (*set_ti_t)(int *, double *);
typedef int (*set_j_t)(int *, int *, int *, double *, double *, double *, double *, double *, double *, double *,
                       double *);
typedef void (*first_t)(int *, int *, int *, int *, double (*)[3], double (*)[3], double (*)[3], double (*)[3],
                        double *, double *, double *);
typedef int (*last2_t)(int *, int *, int *, int *, double (*)[3], double (*)[3], double *, double *, double (*)[3],
                       double (*)[3], double *, int *);
static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static double rnd()
{
    static unsigned long long s = 88172645463325252ull;
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (double)(s >> 11) / 9007199254740992.0;
}
int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    void *h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    const int n = atoi(argv[2]);
    const int reps = argc > 3 ? atoi(argv[3]) : 300;
    open_t g_open = (open_t)dlsym(h, "g6_open_"), g_close = (open_t)dlsym(h, "g6_close_");
    npipes_t g_np = (npipes_t)dlsym(h, "g6_npipes_");
    set_ti_t g_ti = (set_ti_t)dlsym(h, "g6_set_ti_");
    set_j_t g_j = (set_j_t)dlsym(h, "g6_set_j_particle_");
    first_t g_first = (first_t)dlsym(h, "g6calc_firsthalf_");
    last2_t g_last2 = (last2_t)dlsym(h, "g6calc_lasthalf2_");
    std::vector<double> x(3 * (size_t)n), v(3 * (size_t)n);
    for (size_t q = 0; q < 3 * (size_t)n; q++) { x[q] = 2 * rnd() - 1; v[q] = 0.5 * (2 * rnd() - 1); }
    int cid = 0;
    g_open(&cid);
    const int npipes = g_np();
    double zero3[3] = {0, 0, 0}, tj = 0, dtj = 0.125, mass = 1.0 / n;
    for (int j = 0; j < n; j++) { int id = j + 1; g_j(&cid, &j, &id, &tj, &dtj, &mass, zero3, zero3, zero3, &v[3 * (size_t)j], &x[3 * (size_t)j]); }
    std::vector<int> idx(npipes), inn(npipes);
    std::vector<double> xi(3 * (size_t)npipes), vi(3 * (size_t)npipes), acc(3 * (size_t)npipes), jerk(3 * (size_t)npipes),
        pot(npipes), h2(npipes, 0.0), z3(3 * (size_t)npipes, 0.0), z1(npipes, 0.0);
    double eps2 = 1e-4;
    int nj = n;
    printf("N = %d, npipes = %d, %d reps (us per block step: force call | force call + ni j-updates)\n", n, npipes, reps);
    const int sizes[] = {1, 4, 16, 42, 128, 225, 512, 2048, 8192};
    for (int ni : sizes) {
        if (ni > n || ni > npipes) continue;
        double res[2];
        for (int mode = 0; mode < 2; mode++) {
            double tsum = 0;
            for (int r = 0; r < reps + 10; r++) {
                int i0 = (int)(rnd() * (n - ni));
                for (int k = 0; k < ni; k++) {
                    idx[k] = i0 + k + 1;
                    for (int c = 0; c < 3; c++) { xi[3 * k + c] = x[3 * (size_t)(i0 + k) + c]; vi[3 * k + c] = v[3 * (size_t)(i0 + k) + c]; }
                }
                double ti = 1e-7 * (r + 1) + 1e-3 * mode;
                double t0 = now();
                if (mode == 1)
                    for (int k = 0; k < ni; k++) {
                        int a = i0 + k;
                        g_j(&cid, &a, &idx[k], &ti, &dtj, &mass, zero3, zero3, zero3, &vi[3 * k], &xi[3 * k]);
                    }
                g_ti(&cid, &ti);
                g_first(&cid, &nj, &ni, idx.data(), (double(*)[3])xi.data(), (double(*)[3])vi.data(), (double(*)[3])z3.data(),
                        (double(*)[3])z3.data(), z1.data(), &eps2, h2.data());
                g_last2(&cid, &nj, &ni, idx.data(), (double(*)[3])xi.data(), (double(*)[3])vi.data(), &eps2, h2.data(),
                        (double(*)[3])acc.data(), (double(*)[3])jerk.data(), pot.data(), inn.data());
                double t1 = now();
                if (r >= 10) tsum += t1 - t0;
            }
            res[mode] = 1e6 * tsum / reps;
        }
        printf("ni %5d: %9.1f | %9.1f   (%.3g interactions/s)\n", ni, res[0], res[1], (double)ni * n / (res[1] * 1e-6));
    }
    g_close(&cid);
    return 0;
}
