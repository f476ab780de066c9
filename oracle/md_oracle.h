/* TEST INFRASTRUCTURE -- CPU restatement of the MD side of the QCxMS hot path (see xtb_oracle.h).
 * These functions follow the in-tree Fortran line by line (citations at each definition). */
#ifndef MD_ORACLE_H
#define MD_ORACLE_H
#include <stdint.h>

#include "../include/qcxms_b200.h" /* config / result PODs of the boundary (types only) */
#ifdef __cplusplus
extern "C" {
#endif

void md_oracle_leapfrog(int nat, const double *grad, const double *amass, double tstp, double *xyz, double *vel, double *ke);
void md_oracle_ekinet(int nat, const double *velo, const double *mass, double *e_kin, double *temp);
int md_oracle_impactscale(int nuc, double *velo, const double *mass, const double *velof, double eimp, double ff, double e0);
void md_oracle_fragment_structure(int nat, const int32_t *oz, const double *xyz, double rcut, int at1, int at2, int32_t *frag);
void md_oracle_fragmass(int nat, const int32_t *iat, const int32_t *list, const double *mass, const int32_t *imass, int32_t *nfrag,
                        double *fragx /*[10] amu*/, int32_t *fragat /*[10][200]*/);
void md_oracle_intenergy(int nuc, const int32_t *list, const double *mass, const double *velo, int nfrag, double *T, double *e_int);
int md_oracle_checkqc(int nuc, double *e, const double *grad, const double *qat, int mchrg);
double md_oracle_setetemp(int nfrag, double eimp, double ax, double ieetemp);
int md_oracle_getspin(int nat, const int32_t *ic, int chrg);
void md_oracle_center_of_mass(int nat, const double *mass, const double *xyz, double *cm);
/* egrad, xtb2 branch (src/iniqm.f90:641-655): returns gradfail; E = 0 on failure */
int md_oracle_egrad(int nuc, const double *xyz, const int32_t *iat, int mchrg, double etemp, int method_id, double *E, double *grad,
                    double *qat, int *niter);
/* md() for it > 0, EI (method 0), icoll = 0  (src/md.f90:34-708) */
int md_oracle_md(const qcxms_b200_md_config_t *cfg, int nuc, const int32_t *iat, const double *mass, double *xyz, double *velo,
                 const double *velof, double eimp, double tadd, int max_steps, double *grad, int32_t *list, double *achrg, double *axyz,
                 qcxms_b200_md_result_t *res);
/* md() for it = -1 / 0: the ground-state equilibration and sampling runs (src/md.f90 with it <= 0, called from src/main.F90:545-567) */
int md_oracle_md_gs(const qcxms_b200_md_config_t *cfg, int it, double Tsoll, int nuc, const int32_t *iat, const double *mass, double *xyz,
                    double *velo, double *grad, double *achrg, double *gs, qcxms_b200_md_result_t *res);
/* md() for it > 0 as the mean-free-path MD of a CID run (method 3, icoll >= 1; called from src/main.F90:1860-1866); new_velo in/out, m/s */
int md_oracle_md_mfp(const qcxms_b200_md_config_t *cfg, int nuc, const int32_t *iat, const double *mass, double *xyz, double *velo,
                     int icoll, double *new_velo, int step_limit, double *grad, int32_t *list, double *achrg, double *axyz,
                     qcxms_b200_md_result_t *res);
/* md() as the pre-collision heating MD of an ESI/CID run (method 3, icoll = 0, starting_md = .true.) */
int md_oracle_md_esi(const qcxms_b200_md_config_t *cfg, int nuc, const int32_t *iat, const double *mass, double *xyz, double *velo,
                     double tsoll, double eimp, double tadd, int step_limit, double *grad, int32_t *list, double *achrg, double *axyz,
                     qcxms_b200_md_result_t *res);
/* cid() pieces (src/rotation.f90, src/diag3x3.f90, src/boxmuller.f90, src/cid.f90) */
void md_oracle_eigvec3x3(double a[3][3], double w[3], double q[3][3]);
void md_oracle_euler_rotation(int nuc, double *xyz, double *velo, double a, double b, double c);
void md_oracle_rotation_velo(const double *xyz, int nuc, const double *mass, const double *velo, double *velo_rot, double *e_rot);
double md_oracle_vary_energies(double e_in, double e_distr, double dum, double dum2);
/* one collision for one ion; arrays as in qcxms_b200_cid_batch without the leading trajectory axis */
int md_oracle_cid(const qcxms_b200_cid_config_t *cfg, int nuc, const int32_t *iat, const double *mass, int icoll, double *xyz,
                  double *velo, const double *rnd, double velo_cm_in, double *direc, int32_t *collided, double *grad, double *achrg,
                  double *axyz, int32_t *list, qcxms_b200_cid_result_t *res);
/* fragment records (src/utility.f90:469-498, src/write_fragments.f90:402-441) */
void md_oracle_boltz(int nfrag, double temp, const double *ip, double *pop);
int md_oracle_res_line(char *buf, double charge, int mchrg, int itrj, int icoll, int isec, int j, int ntypes, const int32_t *types,
                       const int32_t *counts);
#ifdef __cplusplus
}
#endif
#endif
