/* qcxms_b200 -- C ABI of the B200-native QCxMS production-trajectory hot path.
 *
 * Every entry point takes plain pointers and sizes; all arrays are HOST memory owned by
 * the caller (the library keeps device buffers behind opaque handles and copies in/out).
 * Array layout: xyz/gradient/velo are [nat][3] C row-major == Fortran (3,nat) column-major,
 * so the reference's arrays can be passed unchanged through iso_c_binding.
 * Units as in the reference: bohr, Eh, Eh/bohr, electron masses, atomic time units, K.
 *
 * Return value of every function: 0 on success, non-zero on a library-level failure
 * (no CUDA device, bad argument, unsupported element).  Per-calculation status follows
 * the reference's convention and is returned through `stat`.
 */
#ifndef QCXMS_B200_H
#define QCXMS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* method selector ids, reference src/tblite.f90:34-40 */
#define QCXMS_B200_GFN1 1
#define QCXMS_B200_GFN2 2
#define QCXMS_B200_IPEA1 11
/* stat values, reference src/tblite.f90:55-58 */
#define QCXMS_B200_STAT_OK 0
#define QCXMS_B200_STAT_FATAL (-1)
#define QCXMS_B200_STAT_UNKNOWN_METHOD 5

/* library-level error codes */
#define QCXMS_B200_ERR_ARG 1
#define QCXMS_B200_ERR_CUDA 2
#define QCXMS_B200_ERR_UNSUPPORTED 3

/* ---------------------------------------------------------------------------------------
 * Replaces  subroutine get_xtb_egrad(num, xyz, charge, multiplicity, method, etemp,
 *                                    output_file, qat, energy, gradient, stat, spec_calc)
 * reference src/tblite.f90:65-66 (called from egrad src/iniqm.f90:625,644, iniqm :188,195,
 * eqm :393,409).  Cold-start SCC with accuracy 1.0 exactly like the reference; output_file /
 * spec_calc side effects are handled by the Fortran shim.
 */
int qcxms_b200_egrad(int nat, const int32_t *num, const double *xyz, int charge, int multiplicity,
                     int method_id, double etemp_kelvin, double *qat, double *energy, double *gradient,
                     int32_t *stat);

/* Batched form: nsys geometries of ONE composition evaluated as one ensemble (one CTA each).
 * xyz [nsys][nat][3]; qat [nsys][nat]; energy [nsys]; gradient [nsys][nat][3]; stat [nsys];
 * niter [nsys] (number of SCC cycles) may be NULL. */
int qcxms_b200_egrad_batch(int nsys, int nat, const int32_t *num, const double *xyz, int charge,
                           int multiplicity, int method_id, double etemp_kelvin, double *qat, double *energy,
                           double *gradient, int32_t *stat, int32_t *niter);

/* Number of basis functions of a composition (sizes the arrays of qcxms_b200_egrad_spec). */
int qcxms_b200_basis_size(int nat, const int32_t *num, int method_id, int32_t *nao);

/* get_xtb_egrad with spec_calc = .true. (reference src/tblite.f90:152-164): besides energy / gradient / charges it hands back
 * what write_qmo (reference src/mo_energ.f90:7-80) puts into tmp.mspec and qcxms.Mspec.tbxtb for getspec (src/mo_spec.f90):
 * nao, ihomo = max(HOMO of the alpha channel, 1) (1-based), orbital energies emo [nao] (Eh, ascending), occupations focc [nao],
 * and qmo [nao][nat] = Mulliken population of every orbital on every atom, + 1e-10, normalised per orbital.  The file writing
 * itself stays on the Fortran side (shim) so that the list-directed format is the compiler's own. */
int qcxms_b200_egrad_spec(int nat, const int32_t *num, const double *xyz, int charge, int multiplicity, int method_id,
                          double etemp_kelvin, double *qat, double *energy, double *gradient, int32_t *stat,
                          int32_t *nao, int32_t *ihomo, double *emo, double *focc, double *qmo);

/* ---------------------------------------------------------------------------------------
 * Small routines of the path exposed 1:1 for parity tests (device implementations).
 * fragment_structure(nat, oz, xyz, rcut, at1=1, at2=0, frag): reference src/fragments.f90:93-182
 * (integer result must be bit-identical to the reference for identical coordinates). */
int qcxms_b200_fragment_structure(int nsys, int nat, const int32_t *num, const double *xyz, double rcut,
                                  int32_t *frag /* [nsys][nat] */);

/* ---------------------------------------------------------------------------------------
 * Ensemble of trajectories: replaces "one qcxms --prod process per TMP.<n> directory"
 * (reference bin/pqcxms:88-98) and the md() loop (reference src/md.f90:34-708) for it > 0.
 */
typedef struct qcxms_b200_ensemble qcxms_b200_ensemble_t;

typedef struct {
    int32_t method_id;   /* QCXMS_B200_GFN2 ... */
    int32_t mchrg;       /* molecular charge of every trajectory (reference: mchrg) */
    int32_t nfragexit;   /* reference default 3 (2 for isec > 1), src/main.F90:2252 */
    int32_t exit_rules;  /* 1: reference EI exit rules (md.f90:623-679); 0: run exactly nmax steps */
    int32_t nmax;        /* maximum number of MD steps (reference nmax = tmax/tstep) */
    int32_t isec;        /* index of the (secondary) run, reference isec; affects the error exit */
    double tstep;        /* MD time step in atomic units (reference tstep after *fstoau) */
    double etemp_in;     /* reference etempin: >= 0 sets the electronic temperature of md()'s INITIAL single point only; the loop
                          * calls setetemp() on every step regardless (src/md.f90:167-172, 443-445); < 0: setetemp() throughout */
    double ieetemp;      /* reference common1 ieetemp (default 0) */
    double ax;           /* reference common1 ax (0 for xtb) */
} qcxms_b200_md_config_t;

typedef struct {
    int32_t mdok;        /* reference mdok */
    int32_t fragstate;   /* reference fragstate: 0 undefined, 1 normal, 2 nfrag=2 constant */
    int32_t nstep;       /* MD steps done */
    int32_t nfrag;       /* fragments at exit */
    int32_t status;      /* 0 running, 1 finished, 2 failed (egrad failure at start) */
    int32_t scc_iter_total; /* total SCC cycles spent (for the roofline accounting) */
    double Tav, Epav, Ekav, aTlast, dtime, ttime;
    double Epot, Ekin;   /* last values */
} qcxms_b200_md_result_t;

/* device: CUDA device ordinal.  num [nat] atomic numbers, mass [nat] in electron masses. */
int qcxms_b200_ensemble_create(const qcxms_b200_md_config_t *cfg, int ntraj, int nat, const int32_t *num,
                               const double *mass, int device, qcxms_b200_ensemble_t **out);
int qcxms_b200_ensemble_destroy(qcxms_b200_ensemble_t *h);

/* initial conditions of one trajectory = contents of start.xyz + qcxms.start
 * (reference src/utility.f90:379-422 rdstart): xyz, velo [nat][3], velof [nat], eimp (Eh), tadd (a.u.) */
int qcxms_b200_ensemble_set_trajectory(qcxms_b200_ensemble_t *h, int itrj, const double *xyz, const double *velo,
                                       const double *velof, double eimp, double tadd);
/* bulk variant: arrays carry a leading [ntraj] axis (reference run_settings layout, src/settings.f90:9-25) */
int qcxms_b200_ensemble_set_all(qcxms_b200_ensemble_t *h, const double *xyz, const double *velo,
                                const double *velof, const double *eimp, const double *tadd);

/* md(): initial egrad + MD loop until every trajectory has exited (or max_steps more steps were
 * taken per trajectory when max_steps > 0).  Returns the number of trajectory-MD-steps executed
 * in *steps_done (may be NULL). */
/* OPT-IN fast mode, NOT the reference protocol: start the SCC of every MD step from the converged shell charges /
 * atomic dipoles / quadrupoles of the previous step of the same trajectory instead of from zero (the reference creates a zeroed
 * wavefunction on every call, src/tblite.f90:133).  Fewer SCC cycles per step; energies move within the SCC thresholds
 * (~1e-7 Eh), so trajectories are no longer step-for-step comparable with the reference.  Off by default; call before run_md. */
int qcxms_b200_ensemble_set_warm_start(qcxms_b200_ensemble_t *h, int on);

/* Switches the ensemble to the mean-free-path MD between two collisions of a CID run: md() as the reference runs it with the
 * global method == 3 and icoll >= 1 (call site src/main.F90:1860-1866; branches src/md.f90:246-255, 325, 365, 438, 466-621, 672,
 * 694-699): no IEE heating, kinetic energy / temperature without the motion of the centre of mass, fragment structures averaged
 * over 50 steps after a fragmentation, end of the run moved to nstep + add_steps by every counted fragmentation.
 * new_velo [ntraj]: the reference's in/out argument new_velo (velocity of the ion's centre of mass, m/s) as cid() returned it.
 * Call after set_all / set_trajectory and before run_md. */
int qcxms_b200_ensemble_set_mfp(qcxms_b200_ensemble_t *h, int icoll, const double *new_velo);
/* Switches the ensemble to the heating MD that precedes the first collision of an ESI/CID run: md() with the global method == 3,
 * icoll = 0 and starting_md = .true. (call site src/main.F90:1357-1362): Berendsen scaling of the velocities towards tscale (K)
 * during the first nadd steps while the ion is intact (src/md.f90:428-434), no IEE heating, error threshold and fragment
 * averaging of the method-3 branch, kinetic energy with the centre-of-mass motion.  eimp = E_Scale and tadd = pretadd are
 * the per-trajectory values given to set_all / set_trajectory.  Call after them and before run_md. */
int qcxms_b200_ensemble_set_esi(qcxms_b200_ensemble_t *h, double tscale);
/* new_velo [ntraj] after run_md (mean-free-path mode only) */
int qcxms_b200_ensemble_get_new_velo(qcxms_b200_ensemble_t *h, double *new_velo);

/* The two ground-state runs that precede the production runs (reference src/main.F90:545-567): md() with it = -1 (equilibration:
 * velocities rescaled towards tsoll when the running mean temperature is more than 5 % off, src/md.f90:402-410; no energy test)
 * and it = 0 (NVE sampling: positions and velocities of EVERY step are recorded -- the lines of qcxms.gs, src/md.f90:380-385).  Both
 * keep the electronic temperature at etemp_in (>= 0 required), add no impact energy, look for no fragments and end after nmax
 * steps.  it = 1 returns to production mode.  Call before run_md. */
int qcxms_b200_ensemble_set_gs_mode(qcxms_b200_ensemble_t *h, int it, double tsoll);
/* it = 0: records first .. first + count - 1 of trajectory itrj, [count][nat][6] = (x, y, z, vx, vy, vz) per atom in a.u. */
int qcxms_b200_ensemble_get_gs(qcxms_b200_ensemble_t *h, int itrj, int first, int count, double *xyzvelo);

int qcxms_b200_ensemble_run_md(qcxms_b200_ensemble_t *h, int max_steps, int64_t *steps_done);

int qcxms_b200_ensemble_get_result(qcxms_b200_ensemble_t *h, int itrj, double *xyz, double *velo, double *grad,
                                   int32_t *list, double *achrg, double *axyz, qcxms_b200_md_result_t *res);

/* bulk variant of get_result: arrays carry a leading [ntraj] axis, res is an array of ntraj structs (any pointer may be NULL) */
int qcxms_b200_ensemble_get_all(qcxms_b200_ensemble_t *h, double *xyz, double *velo, double *grad, int32_t *list,
                                double *achrg, double *axyz, qcxms_b200_md_result_t *res);

/* device time (ms, CUDA events on the launching stream) and kernel launch count of the last run_md */
int qcxms_b200_ensemble_last_timing(qcxms_b200_ensemble_t *h, double *kernel_ms, int64_t *launches,
                                    int64_t *scc_iterations);

/* intenergy (reference src/md.f90:715-741): internal kinetic energy E_int [ntraj][10] (Eh) and temperature T [ntraj][10] (K) of
 * every fragment of the current state (list / velo of the last finished step), computed on the device in the reference's
 * summation order.  Slots beyond the number of fragments are 0. */
int qcxms_b200_ensemble_intenergy(qcxms_b200_ensemble_t *h, double *fragT, double *e_int);

/* Fragment-mass histogram of the finished trajectories of this ensemble: bin = nominal integer mass of a fragment (sum over its
 * atoms of the mass number of the most abundant isotope, or of the atom's own mass rounded when it carries an isotope label),
 * weight 1 per fragment.  A device-side preview of the first fragmentation generation; the spectrum proper -- statistical
 * charges from the fragment IPs, isotope patterns -- is assembled from the qcxms.res records (qcxms_b200/spectrum.py) and combined
 * with qcxms_b200_comm_allreduce_sum.  Returns the device pointer as well (may be NULL). */
int qcxms_b200_ensemble_histogram(qcxms_b200_ensemble_t *h, int nbins, double *bins_host, void **bins_device);

/* ---------------------------------------------------------------------------------------
 * The one collective of the path: trajectories are dealt over the GPUs of a box (reference bin/pqcxms:88-98 deals
 * TMP.<n> directories over processes) and only the spectrum is combined at the end, where the reference concatenates every
 * TMP.<n>/qcxms.res (bin/pqcxms:101-103).  One process per GPU; NCCL (libnccl.so.2, loaded on first use -- set
 * QCXMS_B200_NCCL_LIB to point at another copy) over NVLink / NVSwitch.  Bootstrap as NCCL prescribes: rank 0 calls
 * qcxms_b200_comm_unique_id and hands the 128 bytes to the other ranks by whatever the host program has (MPI_Bcast, a file, ...);
 * then every rank calls qcxms_b200_comm_create (collective: it ends with a one-element all-reduce, so that NCCL's lazy channel
 * set-up is paid here and not in the first spectrum reduction). */
typedef struct qcxms_b200_comm qcxms_b200_comm_t;
#define QCXMS_B200_UNIQUE_ID_BYTES 128
int qcxms_b200_comm_unique_id(void *id128);
int qcxms_b200_comm_create(const void *id128, int nranks, int rank, int device, qcxms_b200_comm_t **out);
int qcxms_b200_comm_destroy(qcxms_b200_comm_t *c);
/* in-place sum over the ranks of a HOST array (staged through device memory): the charge-weighted, isotope-expanded m/z
 * intensities a rank accumulated from its qcxms.res records */
int qcxms_b200_comm_allreduce_sum(qcxms_b200_comm_t *c, double *inout, int n);
/* fragment-mass histogram of this rank's ensemble (qcxms_b200_ensemble_histogram), summed over the ranks on the device and
 * returned in bins_host [nbins] */
int qcxms_b200_ensemble_allreduce_histogram(qcxms_b200_ensemble_t *h, qcxms_b200_comm_t *c, int nbins, double *bins_host);

/* ---------------------------------------------------------------------------------------
 * CID: one ion + collision-gas-atom collision MD per trajectory, replaces cid() (reference src/cid.f90:24-1111;
 * euler_rotation / rotation_velo src/rotation.f90, eigvec3x3 src/diag3x3.f90, vary_energies src/boxmuller.f90:46-76).
 * The reference draws 9 uniform random numbers per call with the Fortran intrinsic; here the caller passes them
 * (rnd[0..2] = a,b,c of euler_rotation; rnd[3..4] = dum,dum2 of vary_energies; rnd[5..8] = f,g,lmin,lpos of the
 * gas-atom placement), so any RNG can sit on the host side and runs are reproducible.
 * Supported gases: He / Ne / Ar (gas_z 2, 10, 18) and N2 (gas_z 7: two atoms of mass gas_mass each, the second one 1.09 A above
 * the first, src/cid.f90:660-667); options ConstVelo / MinPot / vScale are off (defaults). */
typedef struct {
    int32_t method_id, mchrg;
    int32_t gas_z;        /* reference gas%IndAtom */
    int32_t eexact;       /* reference eExact */
    int32_t manual_dist;  /* reference manual_dist (0: automatic) */
    int32_t ntot;         /* maximum number of steps, reference ntot = 15000 */
    double gas_mass;      /* reference gas%mIatom (electron masses) */
    double tstep;         /* a.u. */
    double etemp;         /* <= 0: 5000 K (src/cid.f90:293-295) */
    double elab, ecom;    /* eV; ecom > 0 wins (src/cid.f90:349-353) */
} qcxms_b200_cid_config_t;

typedef struct {
    int32_t stopcid, nstep, nfrag, collided, status, scc_iter_total;
    double velo_cm;       /* out: last centre-of-mass speed of the ion, m/s (reference velo_cm) */
    double aTlast, ttime, epot;
    double direc[3];
} qcxms_b200_cid_result_t;

/* One collision (index icoll >= 1) for ntraj ions of nuc atoms.  In/out arrays carry a leading [ntraj] axis:
 * xyz, velo [ntraj][nuc][3] (in: ion before the collision; out: after), rnd [ntraj][9], velo_cm [ntraj] (in for icoll > 1),
 * direc [ntraj][3] (in for icoll > 1, out for icoll == 1), collided [ntraj] (the reference's SAVEd flag, in/out),
 * grad [ntraj][nuc][3], achrg [ntraj][nuc], axyz [ntraj][nuc][3], list [ntraj][nuc], res [ntraj]. */
int qcxms_b200_cid_batch(const qcxms_b200_cid_config_t *cfg, int ntraj, int nuc, const int32_t *num, const double *mass, int icoll,
                         double *xyz, double *velo, const double *rnd, const double *velo_cm, double *direc, int32_t *collided,
                         double *grad, double *achrg, double *axyz, int32_t *list, qcxms_b200_cid_result_t *res, int device);

const char *qcxms_b200_last_error(void);
const char *qcxms_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
