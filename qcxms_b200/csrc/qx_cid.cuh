// CID collision MD on the device: one CTA per ion + gas-atom system.
//   cid()            reference src/cid.f90:24-1111 (set-up :300-670, loop :739-1052, hand-back :1058-1105)
//   euler_rotation   reference src/rotation.f90:11-88       rotation_velo  :92-182
//   eigvec3x3        reference src/diag3x3.f90:85-260        vary_energies  src/boxmuller.f90:46-76
// The scalar set-up runs on thread 0 in the reference's operation order; the loop reuses the ensemble's egrad.
#pragma once
#include "qx_md.cuh"

namespace qx {

struct CidConfig {
    int mchrg, gas_z, eexact, manual_dist, ntot;
    double gas_mass, tstep, etemp, elab, ecom;
};

// per-trajectory scalar state of the collision loop (reference src/cid.f90:739-1052), carried between chunk launches
struct CidScalars {
    int nstep, m, step_counter, distance_dump, xyzavg_dump, total_steps, collided, fragmented, count_average, check_fragmented, cnt,
        nfrag, status, stopcid, scc_total, pad_;
    int save_natf[10];
    double Tav, new_velo, lowestCOM, ttime, aTlast, old_cm[3], Tinit, summass, epot;
};

struct CidState {   // leading [ntraj] axis; nuc0 = nuc + 1 (N2: + 2) atoms for the ion + gas arrays
    double *xyz, *velo, *direc;                       // ion, in/out [ntraj][nuc][3]; direc [ntraj][3]
    const double *rnd, *velo_cm_in;                   // [ntraj][9], [ntraj]
    double *xyz0, *velo0, *grad0, *achrg0;            // [ntraj][nuc0][3] / [ntraj][nuc0]
    double *avxyz, *avxyz2, *store;                   // [ntraj][nuc][3]
    int *list;                                        // [ntraj][nuc]
    CidScalars *sc;                                   // [ntraj]
};

__device__ inline void cid_eigval3x3(double a[3][3], double w[3]) {
    const double twothirdpi = 8.0 * atan(1.0) / 3.0;
    double r = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    double q = (a[0][0] + a[1][1] + a[2][2]) / 3.0;
    w[0] = a[0][0] - q; w[1] = a[1][1] - q; w[2] = a[2][2] - q;
    double p = sqrt((w[0] * w[0] + w[1] * w[1] + w[2] * w[2] + 2 * r) / 6.0);
    r = (w[0] * (w[1] * w[2] - a[1][2] * a[1][2]) - a[0][1] * (a[0][1] * w[2] - a[1][2] * a[0][2]) +
         a[0][2] * (a[0][1] * a[1][2] - w[1] * a[0][2])) / (p * p * p) * 0.5;
    if (r <= -1.0) r = 0.5 * twothirdpi;
    else if (r >= 1.0) r = 0.0;
    else r = acos(r) / 3.0;
    w[2] = q + 2 * p * cos(r);
    w[0] = q + 2 * p * cos(r + twothirdpi);
    w[1] = 3 * q - w[0] - w[2];
}

__device__ inline void cid_eigvec3x3(double a[3][3], double w[3], double q[3][3]) {
    const double eps = 2.220446049250313e-16;
    double norm, n1, n2, n3, precon;
    int i;
    w[0] = fmax(fabs(a[0][0]), fabs(a[0][1]));
    w[1] = fmax(fabs(a[0][2]), fabs(a[1][1]));
    w[2] = fmax(fabs(a[1][2]), fabs(a[2][2]));
    precon = fmax(w[0], fmax(w[1], w[2]));
    if (precon < eps) {
        w[0] = w[1] = w[2] = 0.0;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) q[r][c] = r == c;
        return;
    }
    norm = 1.0 / precon;
    a[0][0] *= norm; a[0][1] *= norm; a[1][1] *= norm; a[0][2] *= norm; a[1][2] *= norm; a[2][2] *= norm;
    cid_eigval3x3(a, w);
    a[0][0] -= w[0]; a[1][1] -= w[0]; a[2][2] -= w[0];
    q[0][0] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
    q[1][0] = a[0][2] * a[0][1] - a[0][0] * a[1][2];
    q[2][0] = a[0][0] * a[1][1] - a[0][1] * a[0][1];
    q[0][1] = a[0][1] * a[2][2] - a[0][2] * a[1][2];
    q[1][1] = a[0][2] * a[0][2] - a[0][0] * a[2][2];
    q[2][1] = a[0][0] * a[1][2] - a[0][1] * a[0][2];
    q[0][2] = a[1][1] * a[2][2] - a[1][2] * a[1][2];
    q[1][2] = a[1][2] * a[0][2] - a[0][1] * a[2][2];
    q[2][2] = a[0][1] * a[1][2] - a[1][1] * a[0][2];
    n1 = q[0][0] * q[0][0] + q[1][0] * q[1][0] + q[2][0] * q[2][0];
    n2 = q[0][1] * q[0][1] + q[1][1] * q[1][1] + q[2][1] * q[2][1];
    n3 = q[0][2] * q[0][2] + q[1][2] * q[1][2] + q[2][2] * q[2][2];
    norm = n1; i = 1;
    if (n2 > norm) { i = 2; norm = n1; }
    if (n3 > norm) i = 3;
    if (i == 1) { norm = sqrt(1.0 / n1); q[0][0] *= norm; q[1][0] *= norm; q[2][0] *= norm; }
    else if (i == 2) { norm = sqrt(1.0 / n2); q[0][0] = q[0][1] * norm; q[1][0] = q[1][1] * norm; q[2][0] = q[2][1] * norm; }
    else { norm = sqrt(1.0 / n3); q[0][0] = q[0][2] * norm; q[1][0] = q[1][2] * norm; q[2][0] = q[2][2] * norm; }
    if (fabs(q[0][0]) > fabs(q[1][0])) {
        norm = sqrt(1.0 / (q[0][0] * q[0][0] + q[2][0] * q[2][0]));
        q[0][1] = -q[2][0] * norm; q[1][1] = 0.0; q[2][1] = +q[0][0] * norm;
    } else {
        norm = sqrt(1.0 / (q[1][0] * q[1][0] + q[2][0] * q[2][0]));
        q[0][1] = 0.0; q[1][1] = +q[2][0] * norm; q[2][1] = -q[1][0] * norm;
    }
    q[0][2] = q[1][0] * q[2][1] - q[2][0] * q[1][1];
    q[1][2] = q[2][0] * q[0][1] - q[0][0] * q[2][1];
    q[2][2] = q[0][0] * q[1][1] - q[1][0] * q[0][1];
    a[0][0] += w[0]; a[1][1] += w[0]; a[2][2] += w[0];
    n1 = a[0][0] * q[0][1] + a[0][1] * q[1][1] + a[0][2] * q[2][1];
    n2 = a[0][1] * q[0][1] + a[1][1] * q[1][1] + a[1][2] * q[2][1];
    n3 = a[0][2] * q[0][1] + a[1][2] * q[1][1] + a[2][2] * q[2][1];
    a[2][2] = a[0][2] * q[0][2] + a[1][2] * q[1][2] + a[2][2] * q[2][2];
    a[0][2] = a[0][0] * q[0][2] + a[0][1] * q[1][2] + a[0][2] * q[2][2];
    a[1][2] = a[0][1] * q[0][2] + a[1][1] * q[1][2] + a[1][2] * q[2][2];
    n1 = q[0][1] * n1 + q[1][1] * n2 + q[2][1] * n3 - w[1];
    n2 = q[0][1] * a[0][2] + q[1][1] * a[1][2] + q[2][1] * a[2][2];
    n3 = q[0][2] * a[0][2] + q[1][2] * a[1][2] + q[2][2] * a[2][2] - w[1];
    if (fabs(n1) >= fabs(n3)) {
        norm = fmax(fabs(n1), fabs(n2));
        if (norm > eps) {
            if (fabs(n1) >= fabs(n2)) { n2 = n2 / n1; n1 = sqrt(1.0 / (1.0 + n2 * n2)); n2 = n2 * n1; }
            else { n1 = n1 / n2; n2 = sqrt(1.0 / (1.0 + n1 * n1)); n1 = n1 * n2; }
            q[0][1] = n2 * q[0][1] - n1 * q[0][2];
            q[1][1] = n2 * q[1][1] - n1 * q[1][2];
            q[2][1] = n2 * q[2][1] - n1 * q[2][2];
        }
    } else {
        norm = fmax(fabs(n3), fabs(n2));
        if (norm > eps) {
            if (fabs(n3) >= fabs(n2)) { n2 = n2 / n3; n3 = sqrt(1.0 / (1.0 + n2 * n2)); n2 = n2 * n3; }
            else { n3 = n3 / n2; n2 = sqrt(1.0 / (1.0 + n3 * n3)); n3 = n3 * n2; }
            q[0][1] = n3 * q[0][1] - n2 * q[0][2];
            q[1][1] = n3 * q[1][1] - n2 * q[1][2];
            q[2][1] = n3 * q[2][1] - n2 * q[2][2];
        }
    }
    q[0][2] = q[1][0] * q[2][1] - q[2][0] * q[1][1];
    q[1][2] = q[2][0] * q[0][1] - q[0][0] * q[2][1];
    q[2][2] = q[0][0] * q[1][1] - q[1][0] * q[0][1];
    w[0] *= precon; w[1] *= precon; w[2] *= precon;
}

__device__ inline void cid_center_of_mass(int n, const double *mass, const double *xyz, double cm[3]) {
    double tm = 0.0;
    cm[0] = cm[1] = cm[2] = 0.0;
    for (int i = 0; i < n; ++i) {
        cm[0] += mass[i] * xyz[3 * i]; cm[1] += mass[i] * xyz[3 * i + 1]; cm[2] += mass[i] * xyz[3 * i + 2];
        tm += mass[i];
    }
    cm[0] /= tm; cm[1] /= tm; cm[2] /= tm;
}

__device__ inline double cid_ekinet(int n, const double *velo, const double *mass, double *temp) {
    double e = 0.0;
    for (int i = 0; i < n; ++i) e = e + mass[i] * (velo[3 * i] * velo[3 * i] + velo[3 * i + 1] * velo[3 * i + 1] + velo[3 * i + 2] * velo[3 * i + 2]);
    e = e * 0.5;
    *temp = e / (0.5 * 3.0 * n * QC_KB);
    return e;
}

// Collision set-up (thread 0): rotates / places the ion, positions the gas atom, fills xyz0 / velo0 (nuc0 = m.nat = nuc + 1 atoms, nuc + 2 for N2).
// xyz, velo: the ion (global, in/out); returns Tinit and summass through pointers.
__device__ inline void cid_setup_thread0(const DevModel &m, const CidConfig &c, int nuc, int icoll, double *xyz, double *velo, const double *rnd,
                                         double velo_cm_in, double *direc, double *xyz0, double *velo0, double *old_cm, double *tinit_out,
                                         double *summass_out) {
    const double pi = 3.14159265358979323846264338327950288;
    const double *mass = m.mass;
    double cm[3];
    if (icoll == 1) {
        cid_center_of_mass(nuc, mass, xyz, cm);
        for (int i = 0; i < nuc; ++i) for (int k = 0; k < 3; ++k) xyz[3 * i + k] -= cm[k];
        // euler_rotation
        const double al = rnd[0] * 2 * pi, be = rnd[1] * 2 * pi, ga = rnd[2] * pi;
        const double ra[3][3] = {{1, 0, 0}, {0, cos(al), -sin(al)}, {0, sin(al), cos(al)}};
        const double rb[3][3] = {{cos(be), 0, sin(be)}, {0, 1, 0}, {-sin(be), 0, cos(be)}};
        const double rg[3][3] = {{cos(ga), -sin(ga), 0}, {sin(ga), cos(ga), 0}, {0, 0, 1}};
        double d[3][3], R[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { d[i][j] = 0; for (int k = 0; k < 3; ++k) d[i][j] += ra[i][k] * rb[k][j]; }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { R[i][j] = 0; for (int k = 0; k < 3; ++k) R[i][j] += d[i][k] * rg[k][j]; }
        for (int i = 0; i < nuc; ++i) {
            const double r[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, v[3] = {velo[3 * i], velo[3 * i + 1], velo[3 * i + 2]};
            for (int k = 0; k < 3; ++k) {
                xyz[3 * i + k] = R[k][0] * r[0] + R[k][1] * r[1] + R[k][2] * r[2];
                velo[3 * i + k] = R[k][0] * v[0] + R[k][1] * v[1] + R[k][2] * v[2];
            }
        }
        // rotation_velo
        double mat[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, ev[3], evec[3][3], w_new[3], om[3][3], T0;
        for (int i = 0; i < nuc; ++i) {
            const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2], mm = mass[i];
            mat[0][0] += (y * y + z * z) * mm; mat[1][0] += (-x * y) * mm; mat[2][0] += (-x * z) * mm;
            mat[0][1] += (-y * x) * mm; mat[1][1] += (x * x + z * z) * mm; mat[2][1] += (-y * z) * mm;
            mat[0][2] += (-z * x) * mm; mat[1][2] += (-z * y) * mm; mat[2][2] += (x * x + y * y) * mm;
        }
        cid_eigvec3x3(mat, ev, evec);
        cid_ekinet(nuc, velo, mass, &T0);
        for (int k = 0; k < 3; ++k) w_new[k] = sqrt((QC_KB * T0) / ev[k]);
        for (int i = 0; i < 3; ++i) for (int r = 0; r < 3; ++r) om[r][i] = evec[r][i] * w_new[i];
        for (int i = 0; i < nuc; ++i) {
            const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
            double vx = 0, vy = 0, vz = 0;
            for (int j = 0; j < 3; ++j) {
                vx = vx + (om[1][j] * z - om[2][j] * y);
                vy = vy + (om[2][j] * x - om[0][j] * z);
                vz = vz + (om[0][j] * y - om[1][j] * x);
            }
            velo[3 * i] += vx; velo[3 * i + 1] += vy; velo[3 * i + 2] += vz;
        }
    }
    double summass = 0.0;
    for (int i = 0; i < nuc; ++i) summass = summass + mass[i];
    const double beta = c.gas_mass / (c.gas_mass + summass);
    double E_velo, fasti = 0.0;
    if (icoll == 1) {
        const double Eimpact = c.ecom > 0.0 ? c.ecom / beta : c.elab;
        if (!c.eexact) {
            const double sigma = Eimpact * 0.1, dum = rnd[3], dum2 = rnd[4];
            const double z0 = sqrt(-2.0 * log(dum)) * cos(2.0 * pi * dum2), z1 = sqrt(-2.0 * log(dum)) * sin(2.0 * pi * dum2);
            E_velo = (dum > 0.5 ? z0 * sigma + Eimpact : z1 * sigma + Eimpact) * QC_EVTOAU;
        } else
            E_velo = Eimpact * QC_EVTOAU;
        fasti = sqrt(2 * E_velo / summass);
    }
    double Tinit;
    cid_ekinet(nuc, velo, mass, &Tinit);
    cid_center_of_mass(nuc, mass, xyz, cm);
    const double f = rnd[5], g = rnd[6], lmin = rnd[7], lpos = rnd[8];
    double lowestx = 1.7976931348623157e308, lowesty = lowestx, highestx = -lowestx, highesty = -lowestx;
    for (int i = 0; i < nuc; ++i) {
        const double x = xyz[3 * i], y = xyz[3 * i + 1];
        if (x < lowestx) lowestx = x;
        if (x > highestx) highestx = x;
        if (y < lowesty) lowesty = y;
        if (y > highesty) highesty = y;
    }
    const double diff1 = lmin < 0.5 ? lowesty * f : highesty * f;
    const double diff2 = lpos < 0.5 ? lowestx * g : highestx * g;
    const int step_dist = c.manual_dist == 0 ? (2 * nuc * 10 > 800 ? 800 : 2 * nuc * 10) : c.manual_dist;
    double start_dist, xyzAr[3], scale_velo[3];
    if (icoll == 1) {
        start_dist = fasti * (2 * c.tstep);
        start_dist = step_dist * start_dist * QC_AUTOAA;
        if (start_dist < 10.0) start_dist = 10.0;
        const double xs[3] = {cm[0], cm[1], cm[2] + start_dist};
        for (int k = 0; k < 3; ++k) direc[k] = xs[k];
        for (int k = 0; k < 3; ++k) direc[k] = direc[k] / sqrt(direc[0] * direc[0] + direc[1] * direc[1] + direc[2] * direc[2]);
        xyzAr[0] = xs[0] + diff2 * 0.8; xyzAr[1] = xs[1] + diff1 * 0.8; xyzAr[2] = xs[2];
        for (int k = 0; k < 3; ++k) scale_velo[k] = direc[k] * fasti;
    } else {
        start_dist = (velo_cm_in * QC_MSTOAU) * (2 * c.tstep);
        start_dist = step_dist * start_dist * QC_AUTOAA;
        if (start_dist < 17.0) start_dist = 17.0;
        for (int i = 0; i < nuc; ++i) for (int k = 0; k < 3; ++k) xyz[3 * i + k] -= cm[k];
        xyzAr[0] = cm[0] + direc[0] * start_dist + diff2 * 0.7;
        xyzAr[1] = cm[1] + direc[1] * start_dist + diff1 * 0.7;
        xyzAr[2] = cm[2] + direc[2] * start_dist;
        scale_velo[0] = scale_velo[1] = scale_velo[2] = 0.0;
    }
    for (int k = 0; k < 3; ++k) old_cm[k] = cm[k];
    for (int i = 0; i < nuc; ++i)
        for (int k = 0; k < 3; ++k) { velo0[3 * i + k] = velo[3 * i + k] + scale_velo[k]; xyz0[3 * i + k] = xyz[3 * i + k]; }
    const int ig = m.nat - 1;   // last atom of the collision system: the gas atom distArCOM looks at; N2 puts its second atom before it
    for (int k = 0; k < 3; ++k) { xyz0[3 * ig + k] = xyzAr[k]; velo0[3 * ig + k] = 0.0; }
    if (m.nat - nuc == 2) {     // reference src/cid.f90:660-667
        xyz0[3 * nuc] = xyzAr[0]; xyz0[3 * nuc + 1] = xyzAr[1]; xyz0[3 * nuc + 2] = xyzAr[2] + 1.09 * QC_AATOAU;
        velo0[3 * nuc] = velo0[3 * nuc + 1] = velo0[3 * nuc + 2] = 0.0;
    }
    *tinit_out = Tinit;
    *summass_out = summass;
}

}  // namespace qx
