/*
 * oracle_g6abi.c -- TEST INFRASTRUCTURE ONLY (parity oracle), not product code.
 *
 * The GRAPE-6 ABI (include/g6_b200.h part 1) served by the double-precision
 * CPU restatement in oracle_g6.c, so that any g6 caller (oracle/phigrape_replay.cc,
 * the reference ph4 -DGPU objects) can be run against an FP64 oracle with the
 * same call sequence it uses on the B200 library.
 *
 * Why not the reference's own lib/g6lib/g6lib.c: its velocity predictor is
 * wrong for dt != 0 (g6lib.c:85-87 multiplies a1by6 by 3 dt^3 and a2by18 by
 * 6 dt^4), it caps N at 100000 (g6lib.c:7), its neighbour-list calls are empty
 * stubs (g6lib.c:469-481) and it has one pipe.  This file keeps g6lib's ABI
 * conventions (g6lib.c:190-381, sapporoG6lib.cpp:5-81):
 *   a2 = acc/2, j6 = jerk/6, pot returned negative, equal ids skipped,
 *   inn = id of the nearest j by unsoftened r2,
 * and takes the arithmetic from ph4's CPU loop (oracle_force / oracle_predict:
 * idata.cc:198-236, jdata.cc:726-747).
 *
 * Only tests/ and bench.py's reference legs may load this shared object.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void oracle_predict(int nj, double t, const double *time, const double *pos, const double *vel, const double *acc,
                    const double *jerk, double *pred_pos, double *pred_vel);
void oracle_force(int ni, const int *iid, const double *ipos, const double *ivel, int j_start, int j_end,
                  const int *jid, const double *mass, const double *pred_pos, const double *pred_vel, double eps2,
                  int use_ids, double *acc, double *jerk, double *pot, int *nn, double *dnn);
int oracle_neighbours(int iid, const double *ipos, double h2, int j_start, int j_end, const int *jid,
                      const double *mass, const double *pred_pos, int maxlen, int *list);

#define NPIPES 16384

static struct {
    int cap, nj_hi;
    double *t, *m, *x, *v, *a, *j, *px, *pv;
    int *id;
    double ti, pred_ti;
    int pred_valid;
    /* captured i-block */
    int ni, nj;
    int *iid;
    double *ix, *iv, *ih2;
    double eps2;
} S;

static void grow(int need)
{
    if (need <= S.cap) return;
    int nc = S.cap ? S.cap : 1024;
    while (nc < need) nc *= 2;
#define GROW(p, w, T)                                                   \
    do {                                                                \
        T *q = (T *)calloc((size_t)nc * (w), sizeof(T));                \
        if (S.p) memcpy(q, S.p, sizeof(T) * (size_t)S.cap * (w));       \
        free(S.p);                                                      \
        S.p = q;                                                        \
    } while (0)
    GROW(t, 1, double); GROW(m, 1, double); GROW(x, 3, double); GROW(v, 3, double); GROW(a, 3, double);
    GROW(j, 3, double); GROW(px, 3, double); GROW(pv, 3, double); GROW(id, 1, int);
#undef GROW
    S.cap = nc;
}

int g6_open_(int *id) { (void)id; return 0; }
int g6_close_(int *id)
{
    (void)id;
    free(S.t); free(S.m); free(S.x); free(S.v); free(S.a); free(S.j); free(S.px); free(S.pv); free(S.id);
    free(S.iid); free(S.ix); free(S.iv); free(S.ih2);
    memset(&S, 0, sizeof(S));
    return 0;
}
int g6_npipes_(void) { return NPIPES; }
int g6_set_tunit_(void *u) { (void)u; return 0; }
int g6_set_xunit_(void *u) { (void)u; return 0; }
int g6_set_ti_(int *id, double *ti) { (void)id; S.ti = *ti; return 0; }

int g6_set_j_particle_(int *cluster, int *address, int *index, double *tj, double *dtj, double *mass, double k18[3],
                       double j6[3], double a2[3], double v[3], double x[3])
{
    (void)cluster; (void)dtj; (void)k18;
    int a = *address;
    grow(a + 1);
    if (a + 1 > S.nj_hi) S.nj_hi = a + 1;
    S.t[a] = *tj;
    S.m[a] = *mass;
    S.id[a] = *index;
    for (int k = 0; k < 3; k++) {
        S.x[3 * a + k] = x[k];
        S.v[3 * a + k] = v[k];
        S.a[3 * a + k] = 2.0 * a2[k];
        S.j[3 * a + k] = 6.0 * j6[k];
    }
    S.pred_valid = 0;
    return 0;
}

void g6calc_firsthalf_(int *cluster, int *nj, int *ni, int index[], double xi[][3], double vi[][3], double aold[][3],
                       double j6old[][3], double phiold[], double *eps2, double h2[])
{
    (void)cluster; (void)aold; (void)j6old; (void)phiold;
    int n = *ni;
    if (n > NPIPES) {
        fprintf(stderr, "oracle_g6abi: ni %d > npipes\n", n);
        exit(-1);
    }
    if (!S.iid) {
        S.iid = (int *)malloc(sizeof(int) * NPIPES);
        S.ix = (double *)malloc(sizeof(double) * 3 * NPIPES);
        S.iv = (double *)malloc(sizeof(double) * 3 * NPIPES);
        S.ih2 = (double *)malloc(sizeof(double) * NPIPES);
    }
    S.ni = n;
    S.nj = *nj < S.cap ? *nj : S.cap;
    S.eps2 = *eps2;
    memcpy(S.iid, index, sizeof(int) * n);
    memcpy(S.ix, xi, sizeof(double) * 3 * n);
    memcpy(S.iv, vi, sizeof(double) * 3 * n);
    for (int i = 0; i < n; i++) S.ih2[i] = h2 ? h2[i] : 0.0;
    if (!S.pred_valid || S.pred_ti != S.ti) {
        oracle_predict(S.nj_hi, S.ti, S.t, S.x, S.v, S.a, S.j, S.px, S.pv);
        S.pred_valid = 1;
        S.pred_ti = S.ti;
    }
}

static int lasthalf(int ni, double acc[][3], double jerk[][3], double pot[], int *inn)
{
    if (ni != S.ni) {
        fprintf(stderr, "oracle_g6abi: lasthalf without matching firsthalf\n");
        exit(-1);
    }
    int *nn = (int *)malloc(sizeof(int) * (ni > 0 ? ni : 1));
    double *dnn = (double *)malloc(sizeof(double) * (ni > 0 ? ni : 1));
    oracle_force(ni, S.iid, S.ix, S.iv, 0, S.nj, S.id, S.m, S.px, S.pv, S.eps2, 1, &acc[0][0], &jerk[0][0], pot, nn,
                 dnn);
    if (inn)
        for (int i = 0; i < ni; i++) inn[i] = nn[i] >= 0 ? S.id[nn[i]] : -1;
    free(nn);
    free(dnn);
    return 0;
}

int g6calc_lasthalf_(int *cluster, int *nj, int *ni, int index[], double xi[][3], double vi[][3], double *eps2,
                     double h2[], double acc[][3], double jerk[][3], double pot[])
{
    (void)cluster; (void)nj; (void)index; (void)xi; (void)vi; (void)eps2; (void)h2;
    return lasthalf(*ni, acc, jerk, pot, NULL);
}
int g6calc_lasthalf2_(int *cluster, int *nj, int *ni, int index[], double xi[][3], double vi[][3], double *eps2,
                      double h2[], double acc[][3], double jerk[][3], double pot[], int inn[])
{
    (void)cluster; (void)nj; (void)index; (void)xi; (void)vi; (void)eps2; (void)h2;
    return lasthalf(*ni, acc, jerk, pot, inn);
}
int g6_initialize_jp_buffer_(int *c, int *n) { (void)c; (void)n; return 0; }
int g6_flush_jp_buffer_(int *c) { (void)c; return 0; }
int g6_reset_(int *c) { (void)c; return 0; }
int g6_reset_fofpga_(int *c) { (void)c; return 0; }
int g6_read_neighbour_list_(int *c) { (void)c; return 0; }
int g6_get_neighbour_list_(int *c, int *ipipe, int *maxlength, int *n_neighbours, int list[])
{
    (void)c;
    int ip = *ipipe;
    if (ip < 0 || ip >= S.ni) {
        fprintf(stderr, "oracle_g6abi: ipipe out of range\n");
        exit(-1);
    }
    int n = oracle_neighbours(S.iid[ip], S.ix + 3 * ip, S.ih2[ip], 0, S.nj, S.id, S.m, S.px, *maxlength, list);
    *n_neighbours = n;
    return n > *maxlength;
}
int get_device_count(void) { return 1; }
