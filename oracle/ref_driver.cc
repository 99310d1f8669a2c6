// ref_driver.cc -- TEST INFRASTRUCTURE ONLY.
//
// Thin C-ABI driver (our code) around the UNMODIFIED reference ph4 sources
// (src/amuse_ph4/src/{jdata,idata,scheduler,...}.cc compiled with -DNOMPI where
// they lie under /root/reference; see oracle/Makefile).  It lets tests and the
// golden-vector generator call the reference's own double-precision force loop
// idata::get_partial_acc_and_jerk() (src/amuse_ph4/src/idata.cc:147-237) and
// predictor jdata::predict_all() (src/amuse_ph4/src/jdata.cc:710-750), and run
// the reference Hermite integrator (jdata::advance) in CPU mode, or -- when
// this file is compiled with -DGPU against a g6 library -- in GPU mode through
// the g6 ABI exactly as the ph4 worker does (src/amuse_ph4/src/gpu.cc).
//
// Output goes to oracle/_ref/ only; no reference source is copied.

#include "stdinc.h"
#include "jdata.h"
#include "idata.h"
#include "scheduler.h"
#include <sys/time.h>
#include <unistd.h>
#include <fcntl.h>

static double wall()
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

// The reference prints progress on stdout; silence it while we drive it.
struct quiet {
    int saved;
    quiet() {
        fflush(stdout);
        cout << flush;
        saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        dup2(nul, 1);
        close(nul);
    }
    ~quiet() {
        fflush(stdout);
        cout << flush;
        dup2(saved, 1);
        close(saved);
    }
};

static void load(jdata &jd, int n, const int *id, const double *mass,
                 const double *pos, const double *vel, double eps2, double eta,
                 bool use_gpu)
{
    jd.system_time = 0;
    jd.sync_time = 0;
    for (int j = 0; j < n; j++) {
        vec p(pos[3 * j], pos[3 * j + 1], pos[3 * j + 2]);
        vec v(vel[3 * j], vel[3 * j + 1], vel[3 * j + 2]);
        jd.add_particle(mass[j], 0.0, p, v, id ? id[j] : -1);
    }
    jd.eps2 = eps2;
    jd.eta = eta;
    jd.set_manage_encounters(0);
#ifdef GPU
    jd.have_gpu = true;
    jd.use_gpu = use_gpu;
    jd.gpu_id = 0;          // jdata() leaves it uninitialised (jdata.h:89,137-170)
#else
    jd.have_gpu = false;
    jd.use_gpu = false;
#endif
}

extern "C" {

// Full i = j sweep at t = 0 (the setup() sweep of idata.cc:66-83).
// nn is the reference's j-index, dnn the distance.
int ph4ref_full_sweep(int n, const int *id, const double *mass,
                      const double *pos, const double *vel, double eps2,
                      double *acc, double *jerk, double *pot, int *nn,
                      double *dnn, double *seconds)
{
    quiet q;
    jdata jd;
    load(jd, n, id, mass, pos, vel, eps2, 0.14, false);
    jd.initialize_arrays();
    idata id_(&jd);          // constructor runs setup(): full sweep
    double t0 = wall();
    id_.get_partial_acc_and_jerk();   // timed repeat of the hot loop alone
    double t1 = wall();
    if (seconds) *seconds = t1 - t0;
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) {
            acc[3 * i + k] = id_.iacc[i][k];
            jerk[3 * i + k] = id_.ijerk[i][k];
        }
        pot[i] = id_.ipot[i];
        nn[i] = id_.inn[i];
        dnn[i] = id_.idnn[i];
    }
    return 0;
}

// Forces on an arbitrary i-list (positions/velocities given) from a j-system
// whose state (time, pos, vel, acc, jerk) is given and predicted to time t by
// the reference predictor.  Returns predicted j too.
int ph4ref_predict_force(int nj, const double *mass, const double *tj,
                         const double *pos, const double *vel,
                         const double *acc, const double *jerk, double t,
                         double eps2, int ni, const double *ipos,
                         const double *ivel, double *pred_pos, double *pred_vel,
                         double *iacc, double *ijerk, double *ipot, int *inn,
                         double *idnn, double *seconds)
{
    quiet q;
    jdata jd;
    load(jd, nj, NULL, mass, pos, vel, eps2, 0.14, false);
    jd.initialize_arrays();
    for (int j = 0; j < nj; j++) {
        jd.time[j] = tj[j];
        for (int k = 0; k < 3; k++) {
            jd.acc[j][k] = acc[3 * j + k];
            jd.jerk[j][k] = jerk[3 * j + k];
        }
    }
    jd.predict_all(t, true);
    for (int j = 0; j < nj; j++)
        for (int k = 0; k < 3; k++) {
            pred_pos[3 * j + k] = jd.pred_pos[j][k];
            pred_vel[3 * j + k] = jd.pred_vel[j][k];
        }
    idata id_;               // no jdata yet: setup() is a no-op
    id_.jdat = &jd;
    id_.set_ni(ni);
    id_.ni = ni;
    for (int i = 0; i < ni; i++)
        for (int k = 0; k < 3; k++) {
            id_.ipos[i][k] = ipos[3 * i + k];
            id_.ivel[i][k] = ivel[3 * i + k];
        }
    double t0 = wall();
    id_.get_partial_acc_and_jerk();
    double t1 = wall();
    if (seconds) *seconds = t1 - t0;
    for (int i = 0; i < ni; i++) {
        for (int k = 0; k < 3; k++) {
            iacc[3 * i + k] = id_.iacc[i][k];
            ijerk[3 * i + k] = id_.ijerk[i][k];
        }
        ipot[i] = id_.ipot[i];
        inn[i] = id_.inn[i];
        idnn[i] = id_.idnn[i];
    }
    return 0;
}

// The reference corrector + Aarseth step with block quantisation, idata::correct()
// (src/amuse_ph4/src/idata.cc:399-514), on caller-supplied i-data.  ipos/ivel come in as the
// predicted values and go out corrected; itime/itimestep are updated in place.
int ph4ref_correct(int ni, double tnext, double eta, double *itime, double *itimestep,
                   const double *old_acc, const double *old_jerk, const double *iacc, const double *ijerk,
                   double *ipos, double *ivel)
{
    quiet q;
    jdata jd;
    jd.eta = eta;
    idata id_;
    id_.jdat = &jd;
    id_.set_ni(ni);
    id_.ni = ni;
    for (int i = 0; i < ni; i++) {
        id_.itime[i] = itime[i];
        id_.itimestep[i] = itimestep[i];
        for (int k = 0; k < 3; k++) {
            id_.old_acc[i][k] = old_acc[3 * i + k];
            id_.old_jerk[i][k] = old_jerk[3 * i + k];
            id_.iacc[i][k] = iacc[3 * i + k];
            id_.ijerk[i][k] = ijerk[3 * i + k];
            id_.ipos[i][k] = ipos[3 * i + k];
            id_.ivel[i][k] = ivel[3 * i + k];
        }
    }
    id_.correct(tnext);
    for (int i = 0; i < ni; i++) {
        itime[i] = id_.itime[i];
        itimestep[i] = id_.itimestep[i];
        for (int k = 0; k < 3; k++) {
            ipos[3 * i + k] = id_.ipos[i][k];
            ivel[3 * i + k] = id_.ivel[i][k];
        }
    }
    return 0;
}

// idata::check_encounters() (src/amuse_ph4/src/idata.cc:618-733) on caller-supplied nearest-neighbour data.
// jmass/jradius/jid describe the j system (nj particles), the i-arrays the current block.
// out[4] = close1, close2, coll1, coll2.
int ph4ref_check_encounters(int nj, const int *jid, const double *jmass, const double *jradius, const double *jvel,
                            int ni, const int *iid, const int *inn, const double *idnn, const double *imass,
                            const double *iradius, double rmin, int *out)
{
    quiet q;
    jdata jd;
    for (int j = 0; j < nj; j++) {
        vec p(0.0, 0.0, 0.0), v(jvel[3 * j], jvel[3 * j + 1], jvel[3 * j + 2]);
        jd.add_particle(jmass[j], jradius[j], p, v, jid[j]);
    }
    jd.rmin = rmin;
    jd.use_gpu = false;
    idata id_;
    id_.jdat = &jd;
    id_.set_ni(ni);
    id_.ni = ni;
    for (int i = 0; i < ni; i++) {
        id_.iid[i] = iid[i];
        id_.inn[i] = inn[i];
        id_.idnn[i] = idnn[i];
        id_.imass[i] = imass[i];
        id_.iradius[i] = iradius[i];
    }
    id_.check_encounters();
    out[0] = jd.close1; out[1] = jd.close2; out[2] = jd.coll1; out[3] = jd.coll2;
    return 0;
}

// Run the reference Hermite integrator to t_end (the loop of
// src/amuse_ph4/interface.cc:673-674 / parallel_hermite_4.cc run_hermite4).
// use_gpu selects the g6 ABI path when compiled -DGPU.
// out[0]=E0 out[1]=E(t_end) out[2]=block steps out[3]=particle steps
// out[4]=wall seconds of the advance loop  out[5]=final system_time
// max_block_steps > 0 bounds the advance loop (config 4: N = 1M with hard binaries takes
// ~1e8 block steps per time unit; a bounded number of them is timed and extrapolated).
int ph4ref_evolve_steps(int n, const int *id, const double *mass, const double *pos,
                        const double *vel, double eps2, double eta, double t_end,
                        int use_gpu, long max_block_steps, double *out, double *pos_out, double *vel_out);
int ph4ref_evolve(int n, const int *id, const double *mass, const double *pos,
                  const double *vel, double eps2, double eta, double t_end,
                  int use_gpu, double *out, double *pos_out, double *vel_out)
{
    return ph4ref_evolve_steps(n, id, mass, pos, vel, eps2, eta, t_end, use_gpu, 0, out, pos_out, vel_out);
}

int ph4ref_evolve_steps(int n, const int *id, const double *mass, const double *pos,
                        const double *vel, double eps2, double eta, double t_end,
                        int use_gpu, long max_block_steps, double *out, double *pos_out, double *vel_out)
{
    quiet q;
    jdata jd;
    load(jd, n, id, mass, pos, vel, eps2, eta, use_gpu != 0);
    jd.initialize_arrays();
    idata id_(&jd);
    jd.set_initial_timestep();
    scheduler sched(&jd);
    jd.E0 = jd.get_energy();
    double t0 = wall();
    while (jd.system_time < t_end && (max_block_steps <= 0 || jd.block_steps < max_block_steps)) jd.advance();
    jd.synchronize_all();
    double t1 = wall();
    out[0] = jd.E0;
    out[1] = jd.get_energy();
    out[2] = jd.block_steps;
    out[3] = jd.total_steps;
    out[4] = t1 - t0;
    out[5] = jd.system_time;
    if (pos_out && vel_out)
        for (int j = 0; j < jd.nj; j++) {
            int jj = jd.inverse_id.count(id ? id[j] : j) ? j : j;
            for (int k = 0; k < 3; k++) {
                pos_out[3 * jj + k] = jd.pos[j][k];
                vel_out[3 * jj + k] = jd.vel[j][k];
            }
        }
    return 0;
}

// The same loop with ph4's own close-encounter management switched on, as its standalone driver runs it
// (parallel_hermite_4.cc:185,215: set_manage_encounters(m) + advance_and_check_encounter(); m = 1 replaces a
// close pair by its centre of mass with unperturbed two-body motion, close_encounter.cc:1-9,66-73).  BASELINE
// configs[3]: without it the primordial binaries pin the time step.  out[6] = particles left in the j-memory.
int ph4ref_evolve_enc(int n, const int *id, const double *mass, const double *pos, const double *vel, double eps2,
                      double eta, double t_end, int use_gpu, int manage_encounters, long max_block_steps, double *out)
{
    quiet q;
    jdata jd;
    load(jd, n, id, mass, pos, vel, eps2, eta, use_gpu != 0);
    jd.set_manage_encounters(manage_encounters);
    jd.initialize_arrays();
    idata id_(&jd);
    jd.set_initial_timestep();
    scheduler sched(&jd);
    jd.E0 = jd.get_energy();
    double t0 = wall();
    long steps = 0;
    while (jd.system_time < t_end && (max_block_steps <= 0 || steps < max_block_steps)) {
        jd.advance_and_check_encounter();
        steps++;
    }
    double t1 = wall();
    jd.synchronize_all();
    out[0] = jd.E0;
    out[1] = jd.get_energy();
    out[2] = jd.block_steps;
    out[3] = jd.total_steps;
    out[4] = t1 - t0;
    out[5] = jd.system_time;
    out[6] = jd.nj;
    return 0;
}

}  // extern "C"
