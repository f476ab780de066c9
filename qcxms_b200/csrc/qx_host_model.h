// Host-side construction of the per-composition model tables (basis set, H0 scaling, D4 reference
// C6, work lists) and their upload to the device.  This is the "calculator construction" that the
// reference repeats on every get_xtb_egrad call (src/tblite.f90:121-130); here it is
// done once per composition and shared by every trajectory of the ensemble.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "params/constants.h"
#include "params/d4_refdata.h"
#include "params/elem_tables.h"
#include "params/gfn2_params.h"
#include "params/gfn1_params.h"
#include "params/d3_refdata.h"
#include "params/stong_table.h"
#include "qx_model.h"

namespace qx {
__host__ __device__ inline int tc_padded_dim(int n);

struct HostModel {
    int nat = 0, nsh = 0, nao = 0, ntype = 0, ld = 0, ndim = 0, charge = 0, rows8 = 0;
    double nel[2] = {0, 0};
    std::vector<int> num, type, at_sh0, at_nsh, at_ao0, at_nao, at_nref, at_ngw;
    std::vector<double> at_rcov, at_rad, at_repa, at_repz, at_en, at_mprad, at_mpvcn, at_dk, at_qk, at_r4r2, at_zeff, at_gam,
        at_qcrad, mass, at_refcn, at_refq;
    std::vector<int> sh_at, sh_l, sh_ao0, sh_np;
    std::vector<double> sh_alpha, sh_coef, sh_level, sh_kcn, sh_poly, sh_refocc, sh_hub, sh_gam3, sh_zeta, hscale;
    std::vector<int> ao_at, ao_sh, ao_m;
    std::vector<double> c6ref;
    int method = 2;
    std::vector<double> at_gam3, at_xb, at_r2r4d3, d3ref;
    std::vector<int> at_mxc;
    std::vector<int2> task_int;
    std::vector<int> gr_ptr, gr_task;
    std::vector<double> scal_table;
    // device copy
    void *d_blob = nullptr;
    DevModel dev{};
};

inline double d4_zeta_h(double a, double c, double qref, double qmod) {
    return qmod < 0.0 ? std::exp(a) : std::exp(a * (1.0 - std::exp(c * (1.0 - qref / qmod))));
}

// returns empty string on success, otherwise the reason
inline std::string build_host_model(HostModel &h, int nat, const int32_t *num, const double *mass, int charge, int multiplicity, int method = 2) {
    h = HostModel();
    h.nat = nat;
    h.charge = charge;
    h.method = method;
    const bool gfn1 = method == 1;
    if (gfn1) gfn1_ensure_loaded();
    std::vector<int> types;
    for (int i = 0; i < nat; ++i) {
        int z = num[i];
        if (z < 1 || z > GFN2_MAXZ) return "element Z=" + std::to_string(z) + " is outside the parametrised range (H..Ar)";
        if (gfn1 && !GFN1_EXTRA[z].supported)
            return "no GFN1-xTB parameters for Z=" + std::to_string(z) + " (built-in: H, C, N, O, F, Cl; others through QCXMS_B200_GFN1_PARAM)";
        int t = (int)(std::find(types.begin(), types.end(), z) - types.begin());
        if (t == (int)types.size()) types.push_back(z);
        const gfn2_elem_t &e = gfn1 ? GFN1_ELEM[z] : GFN2_ELEM[z];
        h.num.push_back(z);
        h.type.push_back(t);
        h.at_gam3.push_back(gfn1 ? e.hubbard_deriv : 0.0);
        h.at_xb.push_back(gfn1 ? GFN1_EXTRA[z].xbond : 0.0);
        h.at_r2r4d3.push_back(D3_R2R4[z]);
        h.at_mxc.push_back(D3_MXC[z]);
        // the exponential CN of GFN1 / D3 uses the in-tree D3 radii (reference src/copyc6.f90 setrcov), like the oracle
        h.at_rcov.push_back(gfn1 ? D3_RCOV[z] : 4.0 / 3.0 * COVRAD2009_AA[z] * TB_AATOAU);
        h.at_rad.push_back(e.atomic_rad * TB_AATOAU);
        h.at_repa.push_back(e.rep_alpha);
        h.at_repz.push_back(e.rep_zeff);
        h.at_en.push_back(PAULING_EN[z]);
        h.at_mprad.push_back(e.mp_rad);
        h.at_mpvcn.push_back(e.mp_vcn);
        h.at_dk.push_back(e.dkernel);
        h.at_qk.push_back(e.qkernel);
        h.at_r4r2.push_back(D4_R4R2[z]);
        h.at_zeff.push_back(D4_ZEFF[z]);
        h.at_gam.push_back(D4_GAM[z]);
        h.at_qcrad.push_back(QC_AATOAU * QCXMS_RAD_AA[z]);
        h.mass.push_back(mass ? mass[i] : ATOMIC_MASS_AMU[z] * QC_AMUTOAU);
        h.at_sh0.push_back(h.nsh);
        h.at_nsh.push_back(e.nshell);
        h.at_ao0.push_back(h.nao);
        int nao_at = 0;
        for (int k = 0; k < e.nshell; ++k) {
            const int l = e.ang[k];
            const stong_entry_t *st = nullptr;
            for (int q = 0; q < STONG_NENTRY; ++q)
                if (STONG_TABLE[q].n == e.pqn[k] && STONG_TABLE[q].l == l && STONG_TABLE[q].ng == e.nprim[k]) st = &STONG_TABLE[q];
            if (!st) return "missing STO-nG expansion";
            h.sh_at.push_back(i);
            h.sh_l.push_back(l);
            h.sh_ao0.push_back(h.nao + nao_at);
            h.sh_np.push_back(e.nprim[k]);
            h.sh_level.push_back(e.selfenergy[k] * GFN2_EVTOAU);
            // GFN1: h = level (1 + kcn_l CN) written as level - kcn CN
            h.sh_kcn.push_back(gfn1 ? -e.selfenergy[k] * GFN2_EVTOAU * GFN1_KCN_L[l] : e.kcn[k] * GFN2_EVTOAU);
            h.sh_poly.push_back(e.shpoly[k]);
            h.sh_refocc.push_back(e.refocc[k]);
            h.sh_hub.push_back(e.hubbard * e.shell_hubbard[l]);
            h.sh_gam3.push_back(gfn1 ? 0.0 : e.hubbard_deriv * GFN2_KSHELL3[l]);
            h.sh_zeta.push_back(e.slater[k]);
            const double dfact = l == 2 ? 3.0 : 1.0;
            for (int p = 0; p < QX_MAXPRIM; ++p) {
                double a = 0.0, c = 0.0;
                if (p < e.nprim[k]) {
                    a = st->alpha[p] * e.slater[k] * e.slater[k];
                    c = st->coeff[p] * std::pow(2.0 * a / M_PI, 0.75) * std::pow(std::sqrt(4.0 * a), l) / std::sqrt(dfact);
                }
                h.sh_alpha.push_back(a);
                h.sh_coef.push_back(c);
            }
            // GFN1: a second s shell on the atom (H 2s) is Schmidt-orthogonalised to the first and renormalised (see the oracle's setup_basis)
            if (gfn1 && k > 0 && l == 0) {
                int first = -1;
                for (int kk = 0; kk < k; ++kk) if (e.ang[kk] == l) first = h.at_sh0[i] + kk;
                if (first >= 0) {
                    double *a1 = &h.sh_alpha[(size_t)first * QX_MAXPRIM], *c1 = &h.sh_coef[(size_t)first * QX_MAXPRIM];
                    double *a2 = &h.sh_alpha[(size_t)h.nsh * QX_MAXPRIM], *c2 = &h.sh_coef[(size_t)h.nsh * QX_MAXPRIM];
                    const int n1 = h.sh_np[first], n2 = h.sh_np[h.nsh];
                    if (n1 + n2 > QX_MAXPRIM) return "too many primitives in an orthogonalised shell";
                    double s11 = 0, s12 = 0, s22 = 0;
                    for (int p = 0; p < n1; ++p) for (int q = 0; q < n1; ++q) s11 += c1[p] * c1[q] * std::pow(M_PI / (a1[p] + a1[q]), 1.5);
                    for (int p = 0; p < n1; ++p) for (int q = 0; q < n2; ++q) s12 += c1[p] * c2[q] * std::pow(M_PI / (a1[p] + a2[q]), 1.5);
                    for (int p = 0; p < n2; ++p) for (int q = 0; q < n2; ++q) s22 += c2[p] * c2[q] * std::pow(M_PI / (a2[p] + a2[q]), 1.5);
                    const double f = s12 / s11, nrm = 1.0 / std::sqrt(s22 - s12 * s12 / s11);
                    for (int q = 0; q < n2; ++q) c2[q] *= nrm;
                    for (int p = 0; p < n1; ++p) { a2[n2 + p] = a1[p]; c2[n2 + p] = -f * nrm * c1[p]; }
                    h.sh_np[h.nsh] = n1 + n2;
                }
            }
            for (int m = 0; m < 2 * l + 1; ++m) {
                h.ao_at.push_back(i);
                h.ao_sh.push_back(h.nsh);
                h.ao_m.push_back(m);
            }
            nao_at += 2 * l + 1;
            h.nsh += 1;
        }
        h.at_nao.push_back(nao_at);
        h.nao += nao_at;
    }
    h.ntype = (int)types.size();
    // shared-memory matrices are zero padded to the strip-GEMM dimension; ld == 4 or 12 (mod 16) keeps DMMA fragment loads
    // and 128-bit row accesses free of bank conflicts
    h.rows8 = tc_padded_dim(h.nao) ? tc_padded_dim(h.nao) : h.nao;
    h.ld = h.nao;
    while (h.ld % 16 != 4 && h.ld % 16 != 12) h.ld += 1;
    h.ndim = gfn1 ? h.nsh : h.nsh + 9 * nat;   // GFN1 mixes the shell charges only (no multipoles)
    // occupation numbers (tblite get_occupation / get_alpha_beta_occupation; uhf = min(mult-1, 0), tblite.f90:111)
    double nocc = -(double)charge;
    for (double v : h.sh_refocc) nocc += v;
    if (nocc <= 0.0 || nocc > 2.0 * h.nao) return "no electrons (or too many) for this charge";
    int uhf = multiplicity - 1 < 0 ? multiplicity - 1 : 0;
    auto pmod2 = [](long v) { return (int)(((v % 2) + 2) % 2); };
    int nuhf = pmod2(uhf) == pmod2(std::lround(nocc)) ? uhf : pmod2(std::lround(nocc));
    double diff = nuhf < nocc ? nuhf : nocc, ntmp = nocc - diff;
    h.nel[0] = ntmp / 2 + diff;
    h.nel[1] = ntmp / 2;
    // H0 shell-pair scaling
    static const double kdiag[3] = {GFN2_KDIAG_S, GFN2_KDIAG_P, GFN2_KDIAG_D};
    h.hscale.assign((size_t)h.nsh * h.nsh, 0.0);
    for (int a = 0; a < h.nsh; ++a)
        for (int b = 0; b < h.nsh; ++b) {
            int la = h.sh_l[a], lb = h.sh_l[b];
            if (gfn1) {
                static const double k1[3] = {GFN1_KDIAG_S, GFN1_KDIAG_P, GFN1_KDIAG_D};
                const int ia = h.sh_at[a], ib = h.sh_at[b], za = h.num[ia], zb = h.num[ib];
                const bool va = GFN1_EXTRA[za].valence[a - h.at_sh0[ia]] != 0, vb = GFN1_EXTRA[zb].valence[b - h.at_sh0[ib]] != 0;
                double km;
                if (va && vb) {
                    const double kll = (la + lb == 1) ? GFN1_K_SP : 0.5 * (k1[la] + k1[lb]);
                    const double den = GFN1_ELEM[za].en - GFN1_ELEM[zb].en;
                    km = gfn1_kpair(za, zb) * kll * (1.0 + GFN1_ENSCALE * den * den);
                } else if (va) km = 0.5 * (k1[la] + GFN1_KDIFF);
                else if (vb) km = 0.5 * (k1[lb] + GFN1_KDIFF);
                else km = GFN1_KDIFF;
                h.hscale[(size_t)a * h.nsh + b] = km;
                continue;
            }
            double k;
            if (la == lb) k = kdiag[la];
            else if (la == 2 || lb == 2) k = (la + lb == 2) ? GFN2_K_SD : GFN2_K_PD;
            else k = 0.5 * (kdiag[la] + kdiag[lb]);
            double za = h.sh_zeta[a], zb = h.sh_zeta[b];
            double zij = std::pow(2.0 * std::sqrt(za * zb) / (za + zb), GFN2_WEXP);
            double den = GFN2_ELEM[h.num[h.sh_at[a]]].en - GFN2_ELEM[h.num[h.sh_at[b]]].en;
            h.hscale[(size_t)a * h.nsh + b] = zij * k * (1.0 + GFN2_ENSCALE * den * den);
        }
    // D4 reference systems (in-tree model: reference src/dftd4.f90:524-595) and reference C6 (src/dftd4.f90:451-497)
    static const double freq[D4_NFREQ] = {0.000001, 0.050000, 0.100000, 0.200000, 0.300000, 0.400000, 0.500000, 0.600000,
                                          0.700000, 0.800000, 0.900000, 1.000000, 1.200000, 1.400000, 1.600000, 1.800000,
                                          2.000000, 2.500000, 3.000000, 4.000000, 5.000000, 7.500000, 10.00000};
    std::vector<double> alpha((size_t)h.ntype * QX_MAXREF * D4_NFREQ, 0.0);
    std::vector<int> ngw_t((size_t)h.ntype * QX_MAXREF, 0);
    for (int t = 0; t < h.ntype; ++t) {
        int z = types[t], cnc[32] = {0};
        cnc[0] = 1;
        for (int r = 0; r < D4_REFN[z]; ++r) {
            int is = D4_REFSYS[z][r];
            double iz = D4_ZEFF[is];
            double zt = d4_zeta_h(GFN2_D4_GA, D4_GAM[is] * GFN2_D4_GC, D4_SECQ[is] + iz, D4_GFFH[z][r] + iz);
            for (int k = 0; k < D4_NFREQ; ++k) {
                double aiw = D4_SSCALE[is] * D4_SECAIW[is][k] * zt;
                double v = D4_ASCALE[z][r] * (D4_ALPHAIW[z][r][k] - D4_HCOUNT[z][r] * aiw);
                alpha[((size_t)t * QX_MAXREF + r) * D4_NFREQ + k] = v > 0.0 ? v : 0.0;
            }
            cnc[(int)std::lround(D4_REFCN[z][r])] += 1;
        }
        for (int r = 0; r < D4_REFN[z]; ++r) {
            int icn = cnc[(int)std::lround(D4_REFCN[z][r])];
            ngw_t[(size_t)t * QX_MAXREF + r] = icn * (icn + 1) / 2;
        }
    }
    h.c6ref.assign((size_t)h.ntype * h.ntype * QX_MAXREF * QX_MAXREF, 0.0);
    for (int ti = 0; ti < h.ntype; ++ti)
        for (int tj = 0; tj < h.ntype; ++tj)
            for (int ri = 0; ri < D4_REFN[types[ti]]; ++ri)
                for (int rj = 0; rj < D4_REFN[types[tj]]; ++rj) {
                    const double *ai = &alpha[((size_t)ti * QX_MAXREF + ri) * D4_NFREQ], *aj = &alpha[((size_t)tj * QX_MAXREF + rj) * D4_NFREQ];
                    double acc = 0.0;
                    for (int k = 0; k < D4_NFREQ - 1; ++k) acc += 0.5 * (freq[k + 1] - freq[k]) * (ai[k + 1] * aj[k + 1] + ai[k] * aj[k]);
                    h.c6ref[(((size_t)ti * h.ntype + tj) * QX_MAXREF + ri) * QX_MAXREF + rj] = 3.0 / M_PI * acc;
                }
    // GFN1: D3 reference table per type pair (in-tree data, reference include/pars.fh via src/copyc6.f90)
    h.d3ref.assign((size_t)h.ntype * h.ntype * D3_MAXC * D3_MAXC * 3, -1.0);
    for (int ti = 0; ti < h.ntype; ++ti)
        for (int tj = 0; tj < h.ntype; ++tj)
            for (int a = 0; a < D3_MAXC; ++a)
                for (int b = 0; b < D3_MAXC; ++b)
                    for (int c = 0; c < 3; ++c)
                        h.d3ref[((((size_t)ti * h.ntype + tj) * D3_MAXC + a) * D3_MAXC + b) * 3 + c] = D3_C6AB[types[ti]][types[tj]][a][b][c];
    for (int i = 0; i < nat; ++i) {
        int z = h.num[i];
        h.at_nref.push_back(D4_REFN[z]);
        for (int r = 0; r < QX_MAXREF; ++r) {
            h.at_refcn.push_back(D4_REFCOVCN[z][r]);
            h.at_refq.push_back(D4_GFFQ[z][r]);
            h.at_ngw.push_back(ngw_t[(size_t)h.type[i] * QX_MAXREF + r]);
        }
    }
    // AO-pair work list: bra AO a on atom J <= ket atom I of AO b; on-site blocks as ordered pairs.
    struct Key { int key; int2 p; };
    std::vector<Key> tasks;
    for (int b = 0; b < h.nao; ++b)
        for (int a = 0; a < h.nao; ++a) {
            int ja = h.ao_at[a], ib = h.ao_at[b];
            if (ja > ib) continue;
            int key = ((h.sh_l[h.ao_sh[a]] * 3 + h.sh_l[h.ao_sh[b]]) << 1) | (ja == ib);
            tasks.push_back({key, make_int2(a, b)});
        }
    std::stable_sort(tasks.begin(), tasks.end(), [](const Key &x, const Key &y) { return x.key < y.key; });
    for (auto &t : tasks) h.task_int.push_back(t.p);
    // per-atom reduction lists over the off-site tasks: code = task*2 + (1 if the atom is the bra atom)
    std::vector<std::vector<int>> per(nat);
    for (int t = 0; t < (int)h.task_int.size(); ++t) {
        int ja = h.ao_at[h.task_int[t].x], ib = h.ao_at[h.task_int[t].y];
        if (ja == ib) continue;
        per[ib].push_back(t * 2);
        per[ja].push_back(t * 2 + 1);
    }
    h.gr_ptr.push_back(0);
    for (int k = 0; k < nat; ++k) {
        for (int c : per[k]) h.gr_task.push_back(c);
        h.gr_ptr.push_back((int)h.gr_task.size());
    }
    return "";
}

template <class T>
inline size_t blob_add(std::vector<char> &blob, const std::vector<T> &v) {
    size_t off = (blob.size() + 15) & ~size_t(15);
    blob.resize(off + v.size() * sizeof(T));
    if (!v.empty()) std::memcpy(blob.data() + off, v.data(), v.size() * sizeof(T));
    return off;
}

inline cudaError_t upload_model(HostModel &h) {
    std::vector<char> blob;
#define ADD(name) size_t o_##name = blob_add(blob, h.name)
    ADD(num); ADD(type); ADD(at_sh0); ADD(at_nsh); ADD(at_ao0); ADD(at_nao); ADD(at_nref); ADD(at_ngw);
    ADD(at_rcov); ADD(at_rad); ADD(at_repa); ADD(at_repz); ADD(at_en); ADD(at_mprad); ADD(at_mpvcn); ADD(at_dk); ADD(at_qk);
    ADD(at_r4r2); ADD(at_zeff); ADD(at_gam); ADD(at_qcrad); ADD(mass); ADD(at_refcn); ADD(at_refq);
    ADD(sh_at); ADD(sh_l); ADD(sh_ao0); ADD(sh_np); ADD(sh_alpha); ADD(sh_coef); ADD(sh_level); ADD(sh_kcn); ADD(sh_poly);
    ADD(sh_refocc); ADD(sh_hub); ADD(sh_gam3); ADD(hscale); ADD(ao_at); ADD(ao_sh); ADD(ao_m); ADD(c6ref); ADD(task_int);
    ADD(gr_ptr); ADD(gr_task); ADD(at_gam3); ADD(at_xb); ADD(at_r2r4d3); ADD(at_mxc); ADD(d3ref);
    if (h.scal_table.empty()) {   // the reference's running sum scal = scal + 0.0002 (src/impact.f90:37)
        h.scal_table.resize(20001);
        volatile double scal = 0.0;
        for (int k = 0; k <= 20000; ++k) { h.scal_table[k] = scal; scal = scal + (double)0.0002f; }
    }
    ADD(scal_table);
#undef ADD
    cudaError_t err = cudaMalloc(&h.d_blob, blob.size());
    if (err != cudaSuccess) return err;
    err = cudaMemcpy(h.d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice);
    if (err != cudaSuccess) return err;
    char *base = (char *)h.d_blob;
    DevModel &d = h.dev;
    d.method = h.method;
    d.nat = h.nat; d.nsh = h.nsh; d.nao = h.nao; d.ntype = h.ntype; d.ld = h.ld; d.ndim = h.ndim; d.rows8 = h.rows8; d.mat_in_global = 0;
    d.ntask_int = (int)h.task_int.size(); d.ntask_grad = 0;
    d.nel[0] = h.nel[0]; d.nel[1] = h.nel[1];
#define PTR(name, T) d.name = (const T *)(base + o_##name)
    PTR(num, int); PTR(type, int); PTR(at_sh0, int); PTR(at_nsh, int); PTR(at_ao0, int); PTR(at_nao, int); PTR(at_nref, int); PTR(at_ngw, int);
    PTR(at_rcov, double); PTR(at_rad, double); PTR(at_repa, double); PTR(at_repz, double); PTR(at_en, double); PTR(at_mprad, double);
    PTR(at_mpvcn, double); PTR(at_dk, double); PTR(at_qk, double); PTR(at_r4r2, double); PTR(at_zeff, double); PTR(at_gam, double);
    PTR(at_qcrad, double); PTR(mass, double); PTR(at_refcn, double); PTR(at_refq, double);
    PTR(sh_at, int); PTR(sh_l, int); PTR(sh_ao0, int); PTR(sh_np, int); PTR(sh_alpha, double); PTR(sh_coef, double); PTR(sh_level, double);
    PTR(sh_kcn, double); PTR(sh_poly, double); PTR(sh_refocc, double); PTR(sh_hub, double); PTR(sh_gam3, double); PTR(hscale, double);
    PTR(ao_at, int); PTR(ao_sh, int); PTR(ao_m, int); PTR(c6ref, double); PTR(task_int, int2); PTR(gr_ptr, int); PTR(gr_task, int);
    PTR(scal_table, double);
    PTR(at_gam3, double); PTR(at_xb, double); PTR(at_r2r4d3, double); PTR(at_mxc, int); PTR(d3ref, double);
#undef PTR
    return cudaSuccess;
}

inline ScratchLayout make_layout(const HostModel &h) {
    ScratchLayout L{};
    size_t n2 = (size_t)h.nao * h.nao, nat = h.nat, off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 1) & ~size_t(1); return o; };
    L.S = take(n2); L.H0 = take(n2); L.Dt = take(3 * n2); L.Qt = take(6 * n2);
    L.T = take(std::max(n2, (size_t)(7 * nat + 11 * h.nao)));
    L.P = take(h.dev.polish ? 2 * (size_t)h.rows8 * h.ld : 0); L.W = take(0);   // P: Gram matrix / rotation and the parked C^T of jacobi_polish
    L.matA = take(h.dev.mat_in_global ? 2 * (size_t)h.rows8 * h.ld + 4 : 0); L.matC = L.matA;
    L.gamma = take((size_t)h.nsh * h.nsh);
    L.dcnp = take(nat * nat); L.dcnp4 = take(nat * nat); L.edisp = take(nat * nat); L.c6 = take(nat * nat); L.dc6 = take(nat * nat);
    L.taskout = take(std::max((size_t)5 * h.task_int.size(), 14 * nat + 4 * nat * nat) + nat * nat / 8 + 2 * nat + 16);
    L.br_df = take((size_t)QX_MAX_ITER * h.ndim);
    L.br_u = take((size_t)QX_MAX_ITER * h.ndim);
    L.br_a = take((size_t)2 * QX_MAX_ITER * QX_MAX_ITER);
    L.br_vec = take((size_t)4 * h.ndim + 2 * QX_MAX_ITER);
    L.total = off;
    return L;
}

}  // namespace qx
