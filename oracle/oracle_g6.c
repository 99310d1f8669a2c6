/*
 * oracle_g6.c -- TEST INFRASTRUCTURE ONLY (parity oracle), not product code.
 *
 * Plain-C, double-precision restatement of the reference CPU algorithms that
 * the B200 g6 library must reproduce.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object.  The product library (amuse_b200/csrc) never links it.
 *
 * Parity pinning: this restatement is checked in tests/test_oracle.py against
 *   (1) the golden vectors the survey recorded from the reference itself
 *       (plummer1k.in, eps=0.01, particle 0),
 *   (2) fixtures in tests/golden/ generated here by oracle/_ref (the
 *       UNMODIFIED reference ph4 sources compiled with -DNOMPI, see
 *       oracle/Makefile and oracle/ref_driver.cc), script
 *       oracle/make_golden.py,
 *   (3) the reference's own known-answer tests for the path
 *       (src/amuse_ph4/tests/test_ph4.py:34-52, :349-357, :804-822).
 *
 * Each function cites the reference file:line it follows (paths relative to
 * the reference tree root).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_TINY 2.220446049250313e-16 /* 2^-52, src/amuse_ph4/src/stdinc.h:33 */
#define ORACLE_INF 1.0e300                /* src/amuse_ph4/src/stdinc.h:32 */

/*
 * Hermite predictor, follows jdata::predict_all(),
 * src/amuse_ph4/src/jdata.cc:726-747:
 *   dt = t - t_j; if dt == 0 copy, else
 *   xp = x + dt*(v + 0.5*dt*(a + dt*j/3)),  vp = v + dt*(a + 0.5*dt*j)
 */
void oracle_predict(int nj, double t, const double *time, const double *pos,
                    const double *vel, const double *acc, const double *jerk,
                    double *pred_pos, double *pred_vel)
{
    for (int j = 0; j < nj; j++) {
        double dt = t - time[j];
        for (int k = 0; k < 3; k++) {
            int q = 3 * j + k;
            if (dt == 0) {
                pred_pos[q] = pos[q];
                pred_vel[q] = vel[q];
            } else {
                pred_pos[q] = pos[q] + dt * (vel[q] + 0.5 * dt * (acc[q] + dt * jerk[q] / 3));
                pred_vel[q] = vel[q] + dt * (acc[q] + 0.5 * dt * jerk[q]);
            }
        }
    }
}

/*
 * Partial force loop, follows idata::get_partial_acc_and_jerk(),
 * src/amuse_ph4/src/idata.cc:198-236, over the j-domain [j_start, j_end).
 *
 * Outputs: acc[ni][3], jerk[ni][3], pot[ni], nn[ni] (j index, -1 if none),
 * dnn[ni] (distance to nn, sqrt applied as in idata.cc:234; sqrt(1e300) if
 * none).  The loop skips massless j (idata.cc:208) and guards pot/nn with
 * r2 > 2^-52 (idata.cc:222).
 *
 * One extension that the reference CPU loop does not need but the g6 ABI
 * does (lib/sapporo_light/dev_evaluate_gravity.cu:76-79, lib/g6lib/g6lib.c:135):
 * when use_ids != 0, a pair with iid[i] == jid[j] is skipped.  With exact
 * self-coincidence (dx = 0) both rules give identical sums.
 */
void oracle_force(int ni, const int *iid, const double *ipos, const double *ivel,
                  int j_start, int j_end, const int *jid, const double *mass,
                  const double *pred_pos, const double *pred_vel, double eps2,
                  int use_ids, double *acc, double *jerk, double *pot, int *nn,
                  double *dnn)
{
    for (int i = 0; i < ni; i++) {
        double lpot = 0, ldnn = ORACLE_INF;
        double la[3] = {0, 0, 0}, lj[3] = {0, 0, 0};
        int lnn = -1;
        for (int j = j_start; j < j_end; j++) {
            if (!(mass[j] > ORACLE_TINY)) continue;
            if (use_ids && iid && jid && iid[i] == jid[j]) continue;
            double dx[3], dv[3], r2 = 0, xv = 0;
            for (int k = 0; k < 3; k++) {
                dx[k] = pred_pos[3 * j + k] - ipos[3 * i + k];
                dv[k] = pred_vel[3 * j + k] - ivel[3 * i + k];
                r2 += dx[k] * dx[k];
                xv += dx[k] * dv[k];
            }
            double r2i = 1 / (r2 + eps2 + ORACLE_TINY);
            double ri = sqrt(r2i);
            double mri = mass[j] * ri;
            double mr3i = mri * r2i;
            double a3 = -3 * xv * r2i;
            if (r2 > ORACLE_TINY) {
                lpot -= mri;
                if (r2 < ldnn) {
                    ldnn = r2;
                    lnn = j;
                }
            }
            for (int k = 0; k < 3; k++) {
                la[k] += mr3i * dx[k];
                lj[k] += mr3i * (dv[k] + a3 * dx[k]);
            }
        }
        for (int k = 0; k < 3; k++) {
            acc[3 * i + k] = la[k];
            jerk[3 * i + k] = lj[k];
        }
        pot[i] = lpot;
        nn[i] = lnn;
        dnn[i] = sqrt(ldnn);
    }
}

/*
 * Condition scales of the sums above (test metric only, no reference counterpart):
 * sacc[i] = sum_j |acc_ij|, sjerk[i] = sum_j |jerk_ij| over the same pairs and
 * with the same arithmetic as oracle_force.  A summation carried out in FP32
 * pair arithmetic is backward stable iff its error is a small multiple of
 * 2^-24 times these scales.
 */
void oracle_force_scales(int ni, const int *iid, const double *ipos, const double *ivel,
                         int j_start, int j_end, const int *jid, const double *mass,
                         const double *pred_pos, const double *pred_vel, double eps2,
                         int use_ids, double *sacc, double *sjerk)
{
    for (int i = 0; i < ni; i++) {
        double sa = 0, sj = 0;
        for (int j = j_start; j < j_end; j++) {
            if (!(mass[j] > ORACLE_TINY)) continue;
            if (use_ids && iid && jid && iid[i] == jid[j]) continue;
            double dx[3], dv[3], r2 = 0, xv = 0;
            for (int k = 0; k < 3; k++) {
                dx[k] = pred_pos[3 * j + k] - ipos[3 * i + k];
                dv[k] = pred_vel[3 * j + k] - ivel[3 * i + k];
                r2 += dx[k] * dx[k];
                xv += dx[k] * dv[k];
            }
            double r2i = 1 / (r2 + eps2 + ORACLE_TINY);
            double mr3i = mass[j] * sqrt(r2i) * r2i;
            double a3 = -3 * xv * r2i, a2 = 0, j2 = 0;
            for (int k = 0; k < 3; k++) {
                double a = mr3i * dx[k], jk = mr3i * (dv[k] + a3 * dx[k]);
                a2 += a * a;
                j2 += jk * jk;
            }
            sa += sqrt(a2);
            sj += sqrt(j2);
        }
        sacc[i] = sa;
        sjerk[i] = sj;
    }
}

/*
 * Combination of per-domain partial results, follows the reduction tail of
 * idata::get_acc_and_jerk(), src/amuse_ph4/src/idata.cc:284-313: sum pot, acc,
 * jerk; min over dnn; nn taken from the domain holding the minimum (first
 * domain on exact ties, as the "zero the losers" rule at :308-311 keeps rank
 * order).  Arrays are [ndom][ni][...].
 */
void oracle_combine(int ndom, int ni, const double *pacc, const double *pjerk,
                    const double *ppot, const int *pnn, const double *pdnn,
                    double *acc, double *jerk, double *pot, int *nn, double *dnn)
{
    for (int i = 0; i < ni; i++) {
        double a[3] = {0, 0, 0}, jk[3] = {0, 0, 0}, p = 0, d = ORACLE_INF;
        int n = -1;
        for (int r = 0; r < ndom; r++) {
            size_t o = (size_t)r * ni + i;
            for (int k = 0; k < 3; k++) {
                a[k] += pacc[3 * o + k];
                jk[k] += pjerk[3 * o + k];
            }
            p += ppot[o];
            if (pnn[o] >= 0 && pdnn[o] < d) {
                d = pdnn[o];
                n = pnn[o];
            }
        }
        for (int k = 0; k < 3; k++) {
            acc[3 * i + k] = a[k];
            jerk[3 * i + k] = jk[k];
        }
        pot[i] = p;
        nn[i] = n;
        dnn[i] = (n >= 0) ? d : sqrt(ORACLE_INF);
    }
}

/*
 * j-domain split, follows jdata::define_domain(),
 * src/amuse_ph4/src/jdata.cc:56-67.
 */
void oracle_define_domain(int nj, int size, int rank, int *j_start, int *j_end)
{
    int n = nj / size;
    if (n * size < nj) n++;
    *j_start = rank * n;
    *j_end = *j_start + n;
    if (rank == size - 1) *j_end = nj;
    if (*j_start >= nj) *j_end = *j_start;
}

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/*
 * Neighbour-sphere list of one i-particle.  Semantics follow
 * lib/sapporo_light/dev_evaluate_gravity.cu:60-67 (member iff r2 <= h2 and
 * ids differ) and lib/sapporo_light/sapporo.cpp:248-272 (ids returned, sorted
 * ascending).  Massless j are skipped like in the force loop.  Returns the
 * full count (may exceed maxlen; only the first maxlen sorted ids are stored).
 * "parity unpinned": the reference holds no test for neighbour lists
 * (SURVEY.md section 8c); this restatement is the only judge.
 */
int oracle_neighbours(int iid, const double *ipos, double h2, int j_start,
                      int j_end, const int *jid, const double *mass,
                      const double *pred_pos, int maxlen, int *list)
{
    int n = 0, cap = 64;
    int *tmp = (int *)malloc(sizeof(int) * cap);
    for (int j = j_start; j < j_end; j++) {
        if (!(mass[j] > ORACLE_TINY)) continue;
        if (jid[j] == iid) continue;
        double r2 = 0;
        for (int k = 0; k < 3; k++) {
            double d = pred_pos[3 * j + k] - ipos[k];
            r2 += d * d;
        }
        if (r2 <= h2) {
            if (n == cap) {
                cap *= 2;
                tmp = (int *)realloc(tmp, sizeof(int) * cap);
            }
            tmp[n++] = jid[j];
        }
    }
    qsort(tmp, n, sizeof(int), cmp_int);
    for (int k = 0; k < n && k < maxlen; k++) list[k] = tmp[k];
    free(tmp);
    return n;
}

/*
 * Hermite corrector and Aarseth time step with block quantisation, follows
 * idata::correct(), src/amuse_ph4/src/idata.cc:443-511 (the "#else" branch in use).
 * ipos/ivel: predicted on entry, corrected on exit; itime/itimestep updated in place.
 */
void oracle_correct(int ni, double tnext, double eta, double *itime, double *itimestep,
                    const double *old_acc, const double *old_jerk, const double *iacc,
                    const double *ijerk, double *ipos, double *ivel)
{
    for (int i = 0; i < ni; i++) {
        double dt = tnext - itime[i];
        double dt2 = dt * dt;
        double a2 = 0, j2 = 0, k2 = 0, l2 = 0;
        for (int k = 0; k < 3; k++) {
            int q = 3 * i + k;
            double alpha = -3 * (old_acc[q] - iacc[q]) - dt * (2 * old_jerk[q] + ijerk[q]);
            double beta = 2 * (old_acc[q] - iacc[q]) + dt * (old_jerk[q] + ijerk[q]);
            ipos[q] += (alpha / 12 + beta / 20) * dt2;
            ivel[q] += (alpha / 3 + beta / 4) * dt;
            a2 += iacc[q] * iacc[q];
            j2 += (dt * ijerk[q]) * (dt * ijerk[q]);
            k2 += (2 * alpha) * (2 * alpha);
            l2 += (6 * beta) * (6 * beta);
        }
        double newstep = eta * dt * sqrt((sqrt(a2 * k2) + j2) / (sqrt(j2 * l2) + k2));
        int exponent;
        double oldstep2 = itimestep[i] / (2 * frexp(itimestep[i], &exponent));
        while (fmod(tnext, oldstep2) != 0) oldstep2 /= 2;
        if (newstep < oldstep2) {
            newstep = oldstep2 / 2;
        } else {
            double t2 = 2 * oldstep2;
            if (newstep >= t2 && fmod(tnext, t2) == 0)
                newstep = t2;
            else
                newstep = oldstep2;
        }
        itime[i] = tnext;
        itimestep[i] = newstep;
    }
}

/*
 * First time step, follows jdata::set_initial_timestep(), src/amuse_ph4/src/jdata.cc:503-548
 * (fac = 0.0625, limit = 0.03125, no median limit).
 */
void oracle_initial_timestep(int nj, double system_time, double eta, const double *acc, const double *jerk,
                             double *timestep)
{
    const double fac = 0.0625, limit = 0.03125;
    for (int j = 0; j < nj; j++) {
        double a2 = 0, j2 = 0;
        for (int k = 0; k < 3; k++) {
            a2 += acc[3 * j + k] * acc[3 * j + k];
            j2 += jerk[3 * j + k] * jerk[3 * j + k];
        }
        double firststep;
        if (eta == 0.0) firststep = limit;
        else if (a2 == 0.0 || j2 == 0.0) firststep = fac * eta;
        else firststep = fac * eta * sqrt(a2 / j2);
        if (firststep != firststep) firststep = fac * eta;
        int exponent;
        firststep /= 2 * frexp(firststep, &exponent);
        while (fmod(system_time, firststep) != 0) firststep /= 2;
        while (firststep > limit) firststep /= 2;
        timestep[j] = firststep;
    }
}

/*
 * Close-encounter / collision detection over the current i-list, follows idata::check_encounters(),
 * src/amuse_ph4/src/idata.cc:634-733 (the nearest-neighbour branch): the i-particle with the largest
 * rmin/dnn (>= 1) defines the close pair, the one with the largest (r_i + r_nn)/dnn (>= 1) the colliding
 * pair; pairs of two massless particles are ignored; first maximum wins (strict >).
 * inn = j index of the nearest neighbour (-1: none), idnn its distance.  out[4] = close1, close2, coll1, coll2
 * (ids; -1 if none).
 */
void oracle_check_encounters(int ni, const int *iid, const int *inn, const double *idnn, const double *imass,
                             const double *iradius, const int *jid, const double *jmass, const double *jradius,
                             double rmin, int *out)
{
    double rmax_close = 0, rmax_coll = 0;
    int imax_close = -1, imax_coll = -1;
    out[0] = out[1] = out[2] = out[3] = -1;
    for (int i = 0; i < ni; i++) {
        int jnn = inn[i];
        if (jnn >= 0) {
            if ((jmass[jnn] > ORACLE_TINY) || (imass[i] > ORACLE_TINY)) {
                double r = rmin / idnn[i];
                if (r > rmax_close) {
                    rmax_close = r;
                    imax_close = i;
                }
                r = (iradius[i] + jradius[jnn]) / idnn[i];
                if (r > rmax_coll) {
                    rmax_coll = r;
                    imax_coll = i;
                }
            }
        }
    }
    if (rmax_close >= 1) {
        out[0] = iid[imax_close];
        out[1] = jid[inn[imax_close]];
    }
    if (rmax_coll >= 1) {
        out[2] = iid[imax_coll];
        out[3] = jid[inn[imax_coll]];
    }
}
