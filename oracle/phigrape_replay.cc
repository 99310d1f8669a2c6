// phigrape_replay.cc -- TEST INFRASTRUCTURE ONLY (caller replay), not product code.
//
// phiGRAPE is Fortran and cannot be built in this image (no gfortran), so its block-timestep
// loop -- the third caller of the g6 ABI and BASELINE configs[2] -- is restated here in C++,
// subroutine by subroutine, and driven against ANY g6 library given on the command line
// (dlopen): the B200 library, the FP64 oracle ABI (oracle/oracle_g6abi.c) or the reference's
// own CPU emulation (oracle/_ref/libg6ref.so).  Everything is passed by reference to the
// trailing-underscore symbols, as a Fortran caller does.
//
// Reference restated (paths relative to the reference tree):
//   src/amuse_phigrape/interface.F:652-752   commit_particles (init sequence)
//   src/amuse_phigrape/interface.F:1426-1520 evolve_model main loop
//   src/amuse_phigrape/src/initgrape.F:21-30, sendbodies2grape.F:13-36, update_grape.F:19-63
//   src/amuse_phigrape/src/get_min_t.F:19-24, selectactive.F:19-40, predictor.F:22-38
//   src/amuse_phigrape/src/gravity.F:54-112  (g6_set_ti; chunks of npipe <= NGP = 16384;
//                                             firsthalf + lasthalf2; h2 = eps2)
//   src/amuse_phigrape/src/corrector.F:20-135 (Hermite corrector + Makino-Aarseth step)
//   src/amuse_phigrape/src/timestep.F:21-50  (initial step), energy.F:17-60
//
// usage: phigrape_replay <g6lib.so> <input.bin> <t_end> [eps2] [eta] [eta_s] [max_block_steps] [dump.bin]
//   input.bin: int32 n, then float64 m[n], x[n][3], v[n][3]
// prints one JSON line.
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef int (*open_t)(int *);
typedef int (*npipes_t)(void);
typedef int (*unit_t)(void *);
typedef int (*set_ti_t)(int *, double *);
typedef int (*set_j_t)(int *, int *, int *, double *, double *, double *, double *, double *, double *, double *,
                       double *);
typedef void (*first_t)(int *, int *, int *, int *, double (*)[3], double (*)[3], double (*)[3], double (*)[3],
                        double *, double *, double *);
typedef int (*last_t)(int *, int *, int *, int *, double (*)[3], double (*)[3], double *, double *, double (*)[3],
                      double (*)[3], double *);
typedef int (*last2_t)(int *, int *, int *, int *, double (*)[3], double (*)[3], double *, double *, double (*)[3],
                       double (*)[3], double *, int *);

static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct G6 {
    open_t open, close;
    npipes_t npipes;
    unit_t tunit, xunit;
    set_ti_t set_ti;
    set_j_t set_j;
    first_t first;
    last_t last;
    last2_t last2;
};

template <typename T>
static T sym(void *h, const char *name)
{
    void *p = dlsym(h, name);
    if (!p) {
        fprintf(stderr, "phigrape_replay: symbol %s missing\n", name);
        exit(2);
    }
    return reinterpret_cast<T>(p);
}

int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: %s <g6lib.so> <input.bin> <t_end> [eps2] [eta] [eta_s] [max_block_steps] [dump.bin]\n",
                argv[0]);
        return 2;
    }
    const char *libpath = argv[1];
    const double t_end = atof(argv[3]);
    const double eps2 = argc > 4 ? atof(argv[4]) : 0.0;          // interface.py:190
    const double eta = argc > 5 ? atof(argv[5]) : 0.02;           // interface.F:628
    const double eta_s = argc > 6 ? atof(argv[6]) : 0.01;         // interface.F:627
    const long max_steps = argc > 7 ? atol(argv[7]) : 0;
    const char *dump = argc > 8 ? argv[8] : nullptr;
    const double dt_max = 1.0;                                    // interface.F:668
    const double dt_min = std::ldexp(1.0, -30);                   // paras.inc:13
    const int NGP = 16384;                                        // gravity.F:23

    void *h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        fprintf(stderr, "phigrape_replay: %s\n", dlerror());
        return 2;
    }
    G6 g;
    g.open = sym<open_t>(h, "g6_open_");
    g.close = sym<open_t>(h, "g6_close_");
    g.npipes = sym<npipes_t>(h, "g6_npipes_");
    g.tunit = sym<unit_t>(h, "g6_set_tunit_");
    g.xunit = sym<unit_t>(h, "g6_set_xunit_");
    g.set_ti = sym<set_ti_t>(h, "g6_set_ti_");
    g.set_j = sym<set_j_t>(h, "g6_set_j_particle_");
    g.first = sym<first_t>(h, "g6calc_firsthalf_");
    g.last = sym<last_t>(h, "g6calc_lasthalf_");
    g.last2 = sym<last2_t>(h, "g6calc_lasthalf2_");

    FILE *f = fopen(argv[2], "rb");
    if (!f) {
        perror(argv[2]);
        return 2;
    }
    int n = 0;
    if (fread(&n, sizeof(int), 1, f) != 1 || n <= 0) return 2;
    std::vector<double> m(n), x(3 * (size_t)n), v(3 * (size_t)n);
    if (fread(m.data(), 8, n, f) != (size_t)n || fread(x.data(), 8, 3 * (size_t)n, f) != 3 * (size_t)n ||
        fread(v.data(), 8, 3 * (size_t)n, f) != 3 * (size_t)n)
        return 2;
    fclose(f);

    std::vector<double> a(3 * (size_t)n, 0.0), adot(3 * (size_t)n, 0.0), pot(n, 0.0), t(n, 0.0), dt(n, dt_min);
    std::vector<int> ind(n);
    for (int i = 0; i < n; i++) ind[i] = i + 1;   // Fortran global indices

    // ---- initgrape -------------------------------------------------------------------------
    int clusterid = 0, unit = 48;
    g.open(&clusterid);
    int npipe = g.npipes();
    if (npipe > NGP) npipe = NGP;
    g.tunit(&unit);
    g.xunit(&unit);

    double t_lib_force = 0, t_lib_update = 0;
    long force_calls = 0, update_calls = 0;
    // latency by i-block size bucket (powers of two)
    double bucket_t[20] = {0};
    long bucket_n[20] = {0};

    auto update_grape = [&](const std::vector<int> &who) {   // update_grape.F (imode 0) / sendbodies2grape.F
        double t0 = now();
        double a2by18[3] = {0, 0, 0}, a1by6[3], aby2[3];
        for (int i : who) {
            for (int k = 0; k < 3; k++) {
                a1by6[k] = adot[3 * (size_t)i + k] * (1.0 / 6.0);
                aby2[k] = a[3 * (size_t)i + k] * 0.5;
            }
            int addr = i;
            g.set_j(&clusterid, &addr, &ind[i], &t[i], &dt[i], &m[i], a2by18, a1by6, aby2, &v[3 * (size_t)i],
                    &x[3 * (size_t)i]);
        }
        t_lib_update += now() - t0;
        update_calls += (long)who.size();
    };

    std::vector<int> act, all(n);
    for (int i = 0; i < n; i++) all[i] = i;
    std::vector<double> xp, vp, anew, jnew, pnew;
    std::vector<int> nn_i(NGP), index_i(NGP);
    std::vector<double> h2_i(NGP), a_i(3 * (size_t)NGP), j_i(3 * (size_t)NGP), p_i(NGP);
    std::vector<double> a_guess(3 * (size_t)n, 1.0), j_guess(3 * (size_t)n, 10.0), p_guess(n, -1.0);

    // gravity.F: forces on the active set `act` with predicted xp/vp (6 per particle)
    auto gravity = [&](double tnow, int ifirst) {
        double tt = tnow;
        double e2 = eps2;
        int nj = n;
        const int ni = (int)act.size();
        anew.resize(3 * (size_t)ni);
        jnew.resize(3 * (size_t)ni);
        pnew.resize(ni);
        double t0 = now();
        g.set_ti(&clusterid, &tt);
        for (int i0 = 0; i0 < ni; i0 += npipe) {
            int nn = std::min(npipe, ni - i0);
            for (int ii = 0; ii < nn; ii++) {
                int ig = act[i0 + ii];
                index_i[ii] = ind[ig];
                h2_i[ii] = eps2;
                for (int k = 0; k < 3; k++) {
                    a_i[3 * (size_t)ii + k] = a_guess[3 * (size_t)ig + k];
                    j_i[3 * (size_t)ii + k] = j_guess[3 * (size_t)ig + k];
                }
                p_i[ii] = p_guess[ig];
            }
            double(*xi)[3] = reinterpret_cast<double(*)[3]>(&xp[3 * (size_t)i0]);
            double(*vi)[3] = reinterpret_cast<double(*)[3]>(&vp[3 * (size_t)i0]);
            double(*ai)[3] = reinterpret_cast<double(*)[3]>(a_i.data());
            double(*ji)[3] = reinterpret_cast<double(*)[3]>(j_i.data());
            double tc0 = now();
            if (ifirst) {   // gravity.F:86-101: first a call on a bad guess, jerk zero test
                g.first(&clusterid, &nj, &nn, index_i.data(), xi, vi, ai, ji, p_i.data(), &e2, h2_i.data());
                g.last(&clusterid, &nj, &nn, index_i.data(), xi, vi, &e2, h2_i.data(), ai, ji, p_i.data());
                for (int q = 0; q < 3 * nn; q++)
                    if (j_i[q] == 0.0) j_i[q] = 1e-5;
            }
            g.first(&clusterid, &nj, &nn, index_i.data(), xi, vi, ai, ji, p_i.data(), &e2, h2_i.data());
            g.last2(&clusterid, &nj, &nn, index_i.data(), xi, vi, &e2, h2_i.data(), ai, ji, p_i.data(), nn_i.data());
            double tc1 = now();
            if (!ifirst) {
                int b = 0;
                while ((1 << b) < nn) b++;
                bucket_t[b] += tc1 - tc0;
                bucket_n[b]++;
            }
            force_calls++;
            for (int ii = 0; ii < nn; ii++) {
                int ig = act[i0 + ii];
                for (int k = 0; k < 3; k++) {
                    anew[3 * (size_t)(i0 + ii) + k] = a_i[3 * (size_t)ii + k];
                    jnew[3 * (size_t)(i0 + ii) + k] = j_i[3 * (size_t)ii + k];
                    a_guess[3 * (size_t)ig + k] = a_i[3 * (size_t)ii + k];
                    j_guess[3 * (size_t)ig + k] = j_i[3 * (size_t)ii + k];
                }
                pnew[i0 + ii] = p_i[ii];
                p_guess[ig] = p_i[ii];
            }
        }
        t_lib_force += now() - t0;
    };

    auto energy = [&](double tcur, double &ekin, double &epot) {   // energy.F + predict_potential.F
        act = all;
        xp.resize(3 * (size_t)n);
        vp.resize(3 * (size_t)n);
        for (int i = 0; i < n; i++) {
            double d = tcur - t[i], d2 = 0.5 * d * d, d3 = d * d2 / 3.0;
            for (int k = 0; k < 3; k++) {
                size_t q = 3 * (size_t)i + k;
                xp[q] = x[q] + v[q] * d + a[q] * d2 + adot[q] * d3;
                vp[q] = v[q] + a[q] * d + adot[q] * d2;
            }
        }
        double sf = t_lib_force;
        long fc = force_calls;
        gravity(tcur, 0);
        t_lib_force = sf;   // diagnostics are not part of the timed run
        force_calls = fc;
        ekin = epot = 0;
        for (int i = 0; i < n; i++) {
            epot += m[i] * pnew[i];
            ekin += m[i] * (vp[3 * (size_t)i] * vp[3 * (size_t)i] + vp[3 * (size_t)i + 1] * vp[3 * (size_t)i + 1] +
                            vp[3 * (size_t)i + 2] * vp[3 * (size_t)i + 2]);
        }
        epot *= 0.5;
        ekin *= 0.5;
    };

    // ---- commit_particles (interface.F:652-752) ---------------------------------------------
    double time_cur = 0.0;
    update_grape(all);          // update_grape(1)
    update_grape(all);          // sendbodies2grape
    act = all;
    xp = x;                     // predictor(1): copy
    vp = v;
    gravity(time_cur, 1);
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) {
            a[3 * (size_t)i + k] = anew[3 * (size_t)i + k];
            adot[3 * (size_t)i + k] = jnew[3 * (size_t)i + k];
        }
        pot[i] = pnew[i];
    }
    for (int i = 0; i < n; i++) {   // timestep.F (imode 0)
        double a2 = 0, j2 = 0;
        for (int k = 0; k < 3; k++) {
            a2 += a[3 * (size_t)i + k] * a[3 * (size_t)i + k];
            j2 += adot[3 * (size_t)i + k] * adot[3 * (size_t)i + k];
        }
        double tmp = (j2 == 0.0) ? eta_s : eta_s * std::sqrt(a2 / j2);
        int power = (int)(std::log(tmp) / std::log(2.0)) - 1;
        tmp = std::ldexp(1.0, power);
        if (tmp > dt_max) tmp = dt_max;
        if (tmp < dt_min) tmp = dt_min;
        dt[i] = tmp;
    }
    update_grape(all);          // update_grape(1)
    double ek0, ep0;
    energy(time_cur, ek0, ep0);
    const double e0 = ek0 + ep0;

    // reset the counters: the timed region is evolve_model
    t_lib_force = t_lib_update = 0;
    force_calls = update_calls = 0;
    memset(bucket_t, 0, sizeof(bucket_t));
    memset(bucket_n, 0, sizeof(bucket_n));

    // ---- evolve_model main loop (interface.F:1426-1520) ---------------------------------------
    long block_steps = 0, particle_steps = 0;
    const double wall0 = now();
    while (time_cur < t_end) {
        if (max_steps && block_steps >= max_steps) break;
        // get_min_t
        double min_t = t[0] + dt[0];
        for (int i = 1; i < n; i++) {
            double tn = t[i] + dt[i];
            if (tn < min_t) min_t = tn;
        }
        // selectactive
        act.clear();
        for (int i = 0; i < n; i++)
            if (t[i] + dt[i] == min_t) act.push_back(i);
        const int na = (int)act.size();
        // predictor(0)
        xp.resize(3 * (size_t)na);
        vp.resize(3 * (size_t)na);
        for (int q = 0; q < na; q++) {
            int i = act[q];
            double d = dt[i], d2 = 0.5 * d * d, d3 = d * d2 / 3.0;
            for (int k = 0; k < 3; k++) {
                size_t s = 3 * (size_t)i + k;
                xp[3 * (size_t)q + k] = x[s] + v[s] * d + a[s] * d2 + adot[s] * d3;
                vp[3 * (size_t)q + k] = v[s] + a[s] * d + adot[s] * d2;
            }
        }
        gravity(min_t, 0);
        // corrector (corrector.F:20-135) + update_loc_p
        for (int q = 0; q < na; q++) {
            int i = act[q];
            double d = dt[i];
            double dt3over6 = d * d * d / 6.0, dt4over24 = dt3over6 * d / 4.0, dt5over120 = dt4over24 * d / 5.0;
            double dtinv = 1.0 / d, dt2inv = dtinv * dtinv, dt3inv = dt2inv * dtinv;
            double a2[3], a3[3];
            for (int k = 0; k < 3; k++) {
                size_t s = 3 * (size_t)i + k;
                double an = anew[3 * (size_t)q + k], jn = jnew[3 * (size_t)q + k];
                a2[k] = -6.0 * (a[s] - an) * dt2inv - (4.0 * adot[s] + 2.0 * jn) * dtinv;
                a3[k] = 12.0 * (a[s] - an) * dt3inv + 6.0 * (adot[s] + jn) * dt2inv;
            }
            double a1abs = 0, adot1abs = 0, a2dot1abs = 0, a3dot1abs = 0;
            for (int k = 0; k < 3; k++) {
                xp[3 * (size_t)q + k] += dt4over24 * a2[k] + dt5over120 * a3[k];
                vp[3 * (size_t)q + k] += dt3over6 * a2[k] + dt4over24 * a3[k];
                double an = anew[3 * (size_t)q + k], jn = jnew[3 * (size_t)q + k];
                a1abs += an * an;
                adot1abs += jn * jn;
                double a2d = a2[k] + d * a3[k];
                a2dot1abs += a2d * a2d;
                a3dot1abs += a3[k] * a3[k];
            }
            a1abs = std::sqrt(a1abs); adot1abs = std::sqrt(adot1abs);
            a2dot1abs = std::sqrt(a2dot1abs); a3dot1abs = std::sqrt(a3dot1abs);
            double dt_new = std::sqrt(eta * (a1abs * a2dot1abs + adot1abs * adot1abs) /
                                      (adot1abs * a3dot1abs + a2dot1abs * a2dot1abs));
            double dt_tmp = d;
            if (dt_new < dt_min) dt_new = dt_min;
            if (dt_new < dt_tmp && dt_new >= dt_min) {
                int power = (int)(std::log(dt_new) / std::log(2.0)) - 1;
                dt_tmp = std::ldexp(1.0, power);
            }
            {
                double y = 2.0 * dt_tmp, r = min_t / y;
                double dmod = (r - (double)(long long)r) * y;   // corrector.F:128-135
                if (dt_new >= 2.0 * dt_tmp && dmod == 0.0 && 2.0 * dt_tmp <= dt_max) dt_tmp = 2.0 * dt_tmp;
            }
            dt[i] = dt_tmp;
            t[i] = min_t;
            for (int k = 0; k < 3; k++) {
                size_t s = 3 * (size_t)i + k;
                x[s] = xp[3 * (size_t)q + k];
                v[s] = vp[3 * (size_t)q + k];
                a[s] = anew[3 * (size_t)q + k];
                adot[s] = jnew[3 * (size_t)q + k];
            }
            pot[i] = pnew[q];
        }
        update_grape(act);   // update_grape(0)
        time_cur = min_t;
        block_steps++;
        particle_steps += na;
    }
    const double wall = now() - wall0;

    double ek1, ep1;
    energy(time_cur, ek1, ep1);
    const double e1 = ek1 + ep1;
    g.close(&clusterid);

    if (dump) {
        FILE *o = fopen(dump, "wb");
        if (o) {
            fwrite(&n, sizeof(int), 1, o);
            fwrite(x.data(), 8, 3 * (size_t)n, o);
            fwrite(v.data(), 8, 3 * (size_t)n, o);
            fwrite(t.data(), 8, n, o);
            fwrite(pot.data(), 8, n, o);
            fclose(o);
        }
    }

    printf("{\"n\": %d, \"t\": %.17g, \"eps2\": %g, \"eta\": %g, \"npipe\": %d, \"seconds\": %.6f, "
           "\"block_steps\": %ld, \"particle_steps\": %ld, \"force_calls\": %ld, \"lib_force_s\": %.6f, "
           "\"lib_update_s\": %.6f, \"E0\": %.17g, \"E1\": %.17g, \"Ek0\": %.17g, \"Ep0\": %.17g, \"latency_us\": {",
           n, time_cur, eps2, eta, npipe, wall, block_steps, particle_steps, force_calls, t_lib_force, t_lib_update, e0,
           e1, ek0, ep0);
    bool firstb = true;
    for (int b = 0; b < 20; b++)
        if (bucket_n[b]) {
            printf("%s\"%d\": [%ld, %.2f]", firstb ? "" : ", ", 1 << b, bucket_n[b], 1e6 * bucket_t[b] / bucket_n[b]);
            firstb = false;
        }
    printf("}}\n");
    return 0;
}
