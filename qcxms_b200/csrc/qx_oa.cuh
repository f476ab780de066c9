// GEMM-based refinement of the eigenpairs of the generalised problem H1 C = S C diag(e) (wide-CTA kernels, DevModel::oa).
//
// The SCC solves one eigenproblem per cycle whose matrix differs little from the one before (previous cycle, or the same cycle of
// the previous MD step).  Instead of rotating row pairs one after the other (one-sided Jacobi, a latency chain of ~65 rounds per
// sweep that leaves the FP64 pipe at ~10 %), the approximate eigenvectors are corrected all at once (Ogita & Aishima, Japan J.
// Indust. Appl. Math. 35, 1007 (2018), generalised to an overlap metric): with
//     Sm = C^T H1 C,   B = C^T S C,   d_i = Sm_ii / B_ii,
// the update C <- C (1 + E) with
//     E_ij = (Sm_ij - d_j B_ij) / (d_j - d_i)   (i != j),      E_ii = (1 - B_ii) / 2
// removes the first-order errors of both the diagonalisation and the S-orthonormality; convergence is quadratic.  Everything but
// the element-wise formula is a dense product and runs on the FP64 tensor pipe (tc_transform / tc_left_apply, DMMA).
// Pairs whose coupling is not small against their gap (|Sm_ij - d B_ij| > kappa |d_j - d_i|: near-degenerate levels, or a seed that
// is far off) cannot be treated to first order: they are collected into clusters (connected components), every cluster block is
// diagonalised exactly (small generalised Jacobi, one warp per cluster), the rotation is applied to C, Sm and B, and the
// first-order step then runs on the rotated quantities.  Anything unexpected -- a cluster larger than QX_OA_MAXC, no convergence in
// QX_OA_ITMAX passes, a diverging step -- returns false and the caller falls back to the Jacobi path, so the result never depends on
// this file being right about its assumptions, only on its convergence test: the pass that ends the iteration has no clusters
// and max |E_ij| < QX_OA_STOP, which leaves errors of O(STOP^2) (12 distorted caffeine cations against the oracle: dE 1e-11 Eh, dg 8e-13, dq 1e-11).
#pragma once
#include "qx_device.cuh"

namespace qx {

#define QX_OA_KAPPA 0.3
#define QX_OA_STOP 1e-6
#define QX_OA_DIVERGE 0.7
#define QX_OA_ITMAX 8
#define QX_OA_MAXC 10     // largest cluster diagonalised exactly
#define QX_OA_NSTORE 3    // eigenvectors of the first SCC cycles are kept per trajectory as seeds for the next MD step
#define QX_OA_SLOT (2 * QX_OA_MAXC * QX_OA_MAXC + QX_OA_MAXC + 2)   // doubles of warp scratch per concurrent cluster
#define QX_OA_NW 3        // 32-bit words per adjacency row (nao <= 96)
#define QX_OA_STAGE 8     // rotated elements staged per thread: nao^2 <= QX_OA_STAGE * QX_NT

// scratch carved out of the vectors that are idle during the eigen-solve (jw .. d4u, contiguous)
struct OaScratch {
    double *vstore;   // rotation blocks of all clusters, cluster c: m x m (row-major) at voff[c]
    double *wk;       // [nslot][QX_OA_SLOT] a | b | rotation parameters of the cluster a warp is working on
    int *lab, *oldlab, *slot_of, *pos_of, *nsize, *mem, *cl_root, *cl_size, *cl_moff, *cl_voff, *misc;   // misc[0] = ncl, [1] = nrot, [2] = flag
    unsigned *adj;    // [n][QX_OA_NW] adjacency bit rows: strong pairs and members of one cluster of the previous round
    int nslot;
};

__host__ __device__ inline size_t oa_scratch_doubles_available(int nat, int nao, int ntype) {
    return (size_t)(3 * nao + 8) + QX_BSOL + 11 * (size_t)nao + (nao & 1) + 7 * (size_t)nat * ntype;
}
__host__ __device__ inline size_t oa_scratch_doubles_fixed(int nao) {   // vstore + the integer arrays (10 + QX_OA_NW arrays of nao ints + misc)
    return (size_t)nao * QX_OA_MAXC + ((10 + QX_OA_NW) * (size_t)nao + 8 + 1) / 2 + 2;
}
// can the refinement run for this composition with this CTA width?
__host__ __device__ inline bool oa_supported(int nat, int nao, int ntype, int nthreads) {
    const size_t avail = oa_scratch_doubles_available(nat, nao, ntype), fixed = oa_scratch_doubles_fixed(nao);
    return nao <= 32 * QX_OA_NW && tc_padded_dim(nao) != 0 && tc_padded_dim(nao) / 8 <= nthreads / 32 && (size_t)nao * nao <= (size_t)QX_OA_STAGE * nthreads &&
           avail >= fixed + QX_OA_SLOT + 2;
}

__device__ inline void oa_carve(const DevModel &m, Sm &s, OaScratch &w) {
    const int n = m.nao;
    double *p = s.jw.ptr();
    w.vstore = p; p += (size_t)n * QX_OA_MAXC;
    int *ip = reinterpret_cast<int *>(p);
    w.lab = ip; w.oldlab = ip + n; w.slot_of = ip + 2 * n; w.pos_of = ip + 3 * n; w.nsize = ip + 4 * n; w.mem = ip + 5 * n;
    w.cl_root = ip + 6 * n; w.cl_size = ip + 7 * n; w.cl_moff = ip + 8 * n; w.cl_voff = ip + 9 * n;
    w.adj = reinterpret_cast<unsigned *>(ip + 10 * n); w.misc = ip + (10 + QX_OA_NW) * n;
    p += ((10 + QX_OA_NW) * (size_t)n + 8 + 1) / 2 + 1;
    if ((reinterpret_cast<size_t>(p) & 15) != 0) p += 1;
    w.wk = p;
    const size_t left = oa_scratch_doubles_available(m.nat, n, m.ntype) - (size_t)(p - s.jw.ptr());
    int ns = (int)(left / QX_OA_SLOT);
    w.nslot = ns < QX_NT / 32 ? ns : QX_NT / 32;
}

// |Sm_ij - d B_ij| > kappa |d_j - d_i| for either orientation (bid: B is the unit matrix)
__device__ __forceinline__ bool oa_strong(const double *Sm, const double *B, const double *d, int ld, bool bid, int i, int j, double kappa) {
    const double sm = Sm[(size_t)i * ld + j], b = bid ? 0.0 : B[(size_t)i * ld + j];
    const double gap = kappa * fabs(d[j] - d[i]);
    return fabs(sm - d[j] * b) > gap || fabs(sm - d[i] * b) > gap;
}

// Exact diagonalisation of one cluster block by one warp: a x = lambda b x with a, b (m x m, leading dimension QX_OA_MAXC) in shared
// memory, m <= QX_OA_MAXC <= 32.  On exit v (m x m, row-major, ld m) holds the b-orthonormal eigenvectors as columns.
// Cholesky b = L L^T, a' = L^-1 a L^-T, cyclic two-sided Jacobi on a' (lanes own rows / columns), v = L^-T y.
__device__ inline void oa_small_geneig(int m, double *a, double *b, double *v) {
    const int lane = threadIdx.x & 31;
    constexpr int LD = QX_OA_MAXC;
    for (int j = 0; j < m; ++j) {   // Cholesky, lower triangle of b in place
        if (lane == j) {
            double dj = b[j * LD + j];
            for (int k = 0; k < j; ++k) dj -= b[j * LD + k] * b[j * LD + k];
            b[j * LD + j] = sqrt(fmax(dj, 1e-300));
        }
        __syncwarp();
        if (lane > j && lane < m) {
            double x = b[lane * LD + j];
            for (int k = 0; k < j; ++k) x -= b[lane * LD + k] * b[j * LD + k];
            b[lane * LD + j] = x / b[j * LD + j];
        }
        __syncwarp();
    }
    if (lane < m) {   // X = L^-1 a: lane owns column `lane`
        for (int i = 0; i < m; ++i) {
            double x = a[i * LD + lane];
            for (int k = 0; k < i; ++k) x -= b[i * LD + k] * a[k * LD + lane];
            a[i * LD + lane] = x / b[i * LD + i];
        }
    }
    __syncwarp();
    if (lane < m) {   // a' = X L^-T: lane owns row `lane`
        for (int c = 0; c < m; ++c) {
            double x = a[lane * LD + c];
            for (int k = 0; k < c; ++k) x -= a[lane * LD + k] * b[c * LD + k];
            a[lane * LD + c] = x / b[c * LD + c];
        }
    }
    __syncwarp();
    if (lane < m)   // symmetrise (the two triangular solves leave rounding-level asymmetry) and start y = 1 in v
        for (int c = 0; c < m; ++c) v[lane * m + c] = lane == c ? 1.0 : 0.0;
    __syncwarp();
    if (lane < m)
        for (int c = lane + 1; c < m; ++c) { const double x = 0.5 * (a[lane * LD + c] + a[c * LD + lane]); a[lane * LD + c] = x; a[c * LD + lane] = x; }
    __syncwarp();
    // two-sided Jacobi in the parallel (round-robin) ordering: the m/2 disjoint rotations of a round are set up by different lanes
    // and applied together, first to the columns (lane = row index), then to the rows (lane = column index)
    double *cs = b + LD * LD;   // [2][LD / 2 + 1] cosines | sines of the round (behind the two matrices of the slot)
    const int M = (m + 1) & ~1, M1 = M - 1, npair = M >> 1;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, dia = 0.0;
        if (lane < m)
            for (int c = 0; c < m; ++c) { const double x = fabs(a[lane * LD + c]); if (c == lane) dia = x; else off = fmax(off, x); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { off = fmax(off, __shfl_xor_sync(0xffffffffu, off, o)); dia = fmax(dia, __shfl_xor_sync(0xffffffffu, dia, o)); }
        if (off <= 1e-14 * fmax(dia, 1e-300)) break;
        for (int r = 0; r < M1; ++r) {
            if (lane < npair) {
                int p = lane == 0 ? M1 : (r + lane) % M1, q = lane == 0 ? r : (r - lane + M1) % M1;
                double c = 1.0, sn = 0.0;
                if (p < m && q < m) {
                    const double apq = a[p * LD + q], app = a[p * LD + p], aqq = a[q * LD + q];
                    if (fabs(apq) > 1e-300) {
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0); sn = t * c;
                    }
                }
                cs[lane] = c; cs[LD / 2 + 1 + lane] = sn;
            }
            __syncwarp();
            if (lane < m) {   // A <- A R, y <- y R: row `lane`
                for (int k = 0; k < npair; ++k) {
                    const int p = k == 0 ? M1 : (r + k) % M1, q = k == 0 ? r : (r - k + M1) % M1;
                    if (p >= m || q >= m) continue;
                    const double c = cs[k], sn = cs[LD / 2 + 1 + k];
                    const double xp = a[lane * LD + p], xq = a[lane * LD + q], yp = v[lane * m + p], yq = v[lane * m + q];
                    a[lane * LD + p] = c * xp - sn * xq; a[lane * LD + q] = sn * xp + c * xq;
                    v[lane * m + p] = c * yp - sn * yq; v[lane * m + q] = sn * yp + c * yq;
                }
            }
            __syncwarp();
            if (lane < m) {   // A <- R^T A: column `lane`
                for (int k = 0; k < npair; ++k) {
                    const int p = k == 0 ? M1 : (r + k) % M1, q = k == 0 ? r : (r - k + M1) % M1;
                    if (p >= m || q >= m) continue;
                    const double c = cs[k], sn = cs[LD / 2 + 1 + k];
                    const double xp = a[p * LD + lane], xq = a[q * LD + lane];
                    a[p * LD + lane] = c * xp - sn * xq; a[q * LD + lane] = sn * xp + c * xq;
                }
            }
            __syncwarp();
        }
    }
    if (lane < m) {   // v = L^-T y: lane owns column `lane`
        for (int i = m - 1; i >= 0; --i) {
            double x = v[i * m + lane];
            for (int k = i + 1; k < m; ++k) x -= b[k * LD + i] * v[k * m + lane];
            v[i * m + lane] = x / b[i * LD + i];
        }
    }
    __syncwarp();
}

// the same for a pair (most clusters are pairs), in registers of one thread: v = {v00, v01, v10, v11}
__device__ inline void oa_geneig2(double a11, double a12, double a22, double b11, double b12, double b22, double *v) {
    const double l11 = sqrt(fmax(b11, 1e-300)), l21 = b12 / l11, l22 = sqrt(fmax(b22 - l21 * l21, 1e-300));
    const double x11 = a11 / l11, x12 = a12 / l11, x21 = (a12 - l21 * x11) / l22, x22 = (a22 - l21 * x12) / l22;
    const double p11 = x11 / l11, p12 = (x12 - p11 * l21) / l22, p21 = x21 / l11, p22 = (x22 - p21 * l21) / l22;
    const double apq = 0.5 * (p12 + p21);
    double c = 1.0, sn = 0.0;
    if (fabs(apq) > 1e-300) {
        const double theta = (p22 - p11) / (2.0 * apq);
        const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
        c = 1.0 / sqrt(t * t + 1.0); sn = t * c;
    }
    // y = (c, sn; -sn, c); v = L^-T y
    const double v10 = -sn / l22, v11 = c / l22;
    v[0] = (c - l21 * v10) / l11; v[1] = (sn - l21 * v11) / l11; v[2] = v10; v[3] = v11;
}

// M <- Q^T M for up to three matrices at once (rows of the clustered nodes mix, all `ncol` columns); staged through registers so
// that it runs in place with two barriers for all of them
__device__ inline void oa_rotate_rows(const OaScratch &w, double *M0, double *M1, double *M2, int ld, int ncol) {
    const int nrot = w.misc[1];
    double st[3][QX_OA_STAGE];
#pragma unroll
    for (int k = 0; k < QX_OA_STAGE; ++k) {
        const int e = threadIdx.x + k * QX_NT;
        st[0][k] = st[1][k] = st[2][k] = 0.0;
        if (e < nrot * ncol) {
            const int r = e / ncol, j = e - r * ncol, i = w.mem[r], c = w.slot_of[i], mm = w.cl_size[c];
            const double *v = w.vstore + w.cl_voff[c] + w.pos_of[i];
            const int *mem = w.mem + w.cl_moff[c];
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            for (int l = 0; l < mm; ++l) {
                const double vl = v[l * mm];
                const size_t o = (size_t)mem[l] * ld + j;
                a0 += vl * M0[o];
                if (M1) a1 += vl * M1[o];
                if (M2) a2 += vl * M2[o];
            }
            st[0][k] = a0; st[1][k] = a1; st[2][k] = a2;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < QX_OA_STAGE; ++k) {
        const int e = threadIdx.x + k * QX_NT;
        if (e < nrot * ncol) {
            const int r = e / ncol, j = e - r * ncol;
            const size_t o = (size_t)w.mem[r] * ld + j;
            M0[o] = st[0][k];
            if (M1) M1[o] = st[1][k];
            if (M2) M2[o] = st[2][k];
        }
    }
    __syncthreads();
}
// M <- M Q for up to two matrices (columns of the clustered nodes mix, first `nrow` rows)
__device__ inline void oa_rotate_cols(const OaScratch &w, double *M0, double *M1, int ld, int nrow) {
    const int nrot = w.misc[1];
    double st[2][QX_OA_STAGE];
#pragma unroll
    for (int k = 0; k < QX_OA_STAGE; ++k) {
        const int e = threadIdx.x + k * QX_NT;
        st[0][k] = st[1][k] = 0.0;
        if (e < nrot * nrow) {
            const int row = e / nrot, r = e - row * nrot, i = w.mem[r], c = w.slot_of[i], mm = w.cl_size[c];
            const double *v = w.vstore + w.cl_voff[c] + w.pos_of[i];
            const int *mem = w.mem + w.cl_moff[c];
            double a0 = 0.0, a1 = 0.0;
            for (int l = 0; l < mm; ++l) {
                const double vl = v[l * mm];
                const size_t o = (size_t)row * ld + mem[l];
                a0 += M0[o] * vl;
                if (M1) a1 += M1[o] * vl;
            }
            st[0][k] = a0; st[1][k] = a1;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < QX_OA_STAGE; ++k) {
        const int e = threadIdx.x + k * QX_NT;
        if (e < nrot * nrow) {
            const int row = e / nrot, r = e - row * nrot;
            const size_t o = (size_t)row * ld + w.mem[r];
            M0[o] = st[0][k];
            if (M1) M1[o] = st[1][k];
        }
    }
    __syncthreads();
}

// One refinement of the eigenpairs.  On entry s.A = H1 (kept), s.C = approximate eigenvectors (transposed: row k = vector k),
// s.S5 = S; bid_first: the vectors are S-orthonormal to working accuracy (previous cycle's result), so the first pass takes
// B = 1.  On success s.C holds the refined vectors and s.emo the eigenvalues; on failure s.C is undefined.
// Returns the number of passes (> 0) or 0 on failure.
template <int NT8>
static __device__ __noinline__ int oa_refine(const DevModel &m, Sm &s, bool bid_first) {
    const int n = m.nao, ld = m.ld, nfull = m.rows8 * ld;
    double *const H1 = s.A, *const Ct = s.C, *const Sm_ = s.X3, *const B = s.X4, *const S = s.S5;
    QX_ASSUME_SHARED(H1); QX_ASSUME_SHARED(Ct); QX_ASSUME_SHARED(Sm_); QX_ASSUME_SHARED(B); QX_ASSUME_SHARED(S);
    double *d = s.emo.ptr();
    OaScratch w;
    oa_carve(m, s, w);
    if (w.nslot < 1) return 0;
    QX_SUB_BEGIN();
    for (int pass = 1; pass <= QX_OA_ITMAX; ++pass) {
        const bool bid = bid_first && pass == 1;
        for (int t = threadIdx.x; t < nfull; t += QX_NT) { Sm_[t] = H1[t]; if (!bid) B[t] = S[t]; }
        __syncthreads();
        QX_SUB(8);
        tc_transform<NT8>(n, Ct, Sm_, ld);
        if (!bid) tc_transform<NT8>(n, Ct, B, ld);
        QX_SUB(9);
        for (int i = threadIdx.x; i < n; i += QX_NT) {
            d[i] = bid ? Sm_[(size_t)i * ld + i] : Sm_[(size_t)i * ld + i] / B[(size_t)i * ld + i];
            w.lab[i] = i; w.oldlab[i] = i;
        }
        __syncthreads();
        bool had_cluster = false;
        // ---- clusters of pairs that are not first-order: detect, diagonalise exactly, rotate, look again (the rotation moves d)
        for (int round = 0; round < 4; ++round) {
            // adjacency: strong pairs (all threads share the pairs) + the members of one cluster of the previous round
            for (int t = threadIdx.x; t < n * QX_OA_NW; t += QX_NT) w.adj[t] = 0u;
            __syncthreads();
            int any = 0;
            for (int t = threadIdx.x; t < n * n; t += QX_NT) {
                const int i = t / n, j = t - i * n;
                if (j <= i) continue;
                const bool same = w.oldlab[i] == w.oldlab[j];
                const bool strong = !same && oa_strong(Sm_, B, d, ld, bid, i, j, m.oa_kappa);
                if (same || strong) {
                    atomicOr(&w.adj[i * QX_OA_NW + (j >> 5)], 1u << (j & 31));
                    atomicOr(&w.adj[j * QX_OA_NW + (i >> 5)], 1u << (i & 31));
                }
                any |= strong;
            }
            const int any_all = __syncthreads_or(any);
            QX_SUB(10);
            if (!any_all) break;
            had_cluster = true;
            // connected components: the smallest node index of a component becomes the label of its members
            for (int it = 0; it < n; ++it) {
                int l = n, changed = 0;
                if (threadIdx.x < n) {
                    const int i = threadIdx.x;
                    l = w.lab[i];
#pragma unroll
                    for (int q = 0; q < QX_OA_NW; ++q) {
                        unsigned bits = w.adj[i * QX_OA_NW + q];
                        while (bits) {
                            const int j = 32 * q + __ffs(bits) - 1;
                            bits &= bits - 1;
                            const int lj = w.lab[j];
                            l = lj < l ? lj : l;
                        }
                    }
                    changed = l != w.lab[i];
                }
                __syncthreads();
                if (threadIdx.x < n) w.lab[threadIdx.x] = l;
                if (!__syncthreads_or(changed)) break;
            }
            // cluster tables: one slot per multi-member root (ascending root index), members in ascending order
            if (threadIdx.x == 0) w.misc[2] = 0;
            if (threadIdx.x < n) {
                const int i = threadIdx.x, li = w.lab[i];
                int size = 0, pos = 0;
                for (int j = 0; j < n; ++j) { const bool same = w.lab[j] == li; size += same; pos += same && j < i; }
                w.pos_of[i] = pos;
                w.nsize[i] = size;
            }
            __syncthreads();
            if (threadIdx.x < n) {
                const int i = threadIdx.x, li = w.lab[i], size = w.nsize[i];
                int slot = -1;
                if (size > 1) {
                    int moff = 0, voff = 0;
                    slot = 0;
                    for (int r = 0; r < li; ++r)
                        if (w.lab[r] == r && w.nsize[r] > 1) { slot += 1; moff += w.nsize[r]; voff += w.nsize[r] * w.nsize[r]; }
                    w.mem[moff + w.pos_of[i]] = i;
                    if (li == i) { w.cl_root[slot] = li; w.cl_size[slot] = size; w.cl_moff[slot] = moff; w.cl_voff[slot] = voff; }
                    if (size > QX_OA_MAXC) w.misc[2] = 1;
                }
                w.slot_of[i] = slot;
                w.oldlab[i] = li;
            }
            __syncthreads();
            if (w.misc[2]) return 0;   // a cluster too large for the exact small solver: the caller takes the Jacobi path
            if (threadIdx.x == 0) {
                int ncl = 0, nrot = 0;
                for (int i = 0; i < n; ++i)
                    if (w.lab[i] == i && w.nsize[i] > 1) { ncl += 1; nrot += w.nsize[i]; }
                w.misc[0] = ncl; w.misc[1] = nrot;
            }
            __syncthreads();
            QX_SUB(11);
            const int ncl = w.misc[0];
            // ---- exact small solves, one warp per cluster
            {
                // pairs: one thread each, from the last warps (the first ones take the larger clusters)
                const int tp = QX_NT - 1 - threadIdx.x;
                if (tp < ncl && w.cl_size[tp] == 2) {
                    const int gi = w.mem[w.cl_moff[tp]], gj = w.mem[w.cl_moff[tp] + 1];
                    oa_geneig2(Sm_[(size_t)gi * ld + gi], Sm_[(size_t)gi * ld + gj], Sm_[(size_t)gj * ld + gj], bid ? 1.0 : B[(size_t)gi * ld + gi],
                               bid ? 0.0 : B[(size_t)gi * ld + gj], bid ? 1.0 : B[(size_t)gj * ld + gj], w.vstore + w.cl_voff[tp]);
                }
                const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
                if (warp < w.nslot) {
                    double *a = w.wk + (size_t)warp * QX_OA_SLOT, *b = a + QX_OA_MAXC * QX_OA_MAXC;
                    for (int c = warp; c < ncl; c += w.nslot) {
                        const int mm = w.cl_size[c];
                        if (mm == 2) continue;
                        const int *mem = w.mem + w.cl_moff[c];
                        for (int e = lane; e < mm * mm; e += 32) {
                            const int r = e / mm, cc = e - r * mm, gi = mem[r], gj = mem[cc];
                            a[r * QX_OA_MAXC + cc] = Sm_[(size_t)gi * ld + gj];
                            b[r * QX_OA_MAXC + cc] = bid ? (r == cc ? 1.0 : 0.0) : B[(size_t)gi * ld + gj];
                        }
                        __syncwarp();
                        oa_small_geneig(mm, a, b, w.vstore + w.cl_voff[c]);
                    }
                }
            }
            __syncthreads();
            QX_SUB(12);
            // ---- rotate: C <- C Q (rows of C^T), Sm <- Q^T Sm Q, B <- Q^T B Q (stays 1 when it was 1)
            oa_rotate_rows(w, Ct, Sm_, bid ? nullptr : B, ld, n);
            oa_rotate_cols(w, Sm_, bid ? nullptr : B, ld, n);
            for (int i = threadIdx.x; i < n; i += QX_NT) d[i] = bid ? Sm_[(size_t)i * ld + i] : Sm_[(size_t)i * ld + i] / B[(size_t)i * ld + i];
            __syncthreads();
            QX_SUB(13);
        }
        // ---- first-order step: X3 <- (1 + E)^T, C^T <- (1 + E)^T C^T
        double emax = 0.0;
        for (int t = threadIdx.x; t < n * n; t += QX_NT) {
            const int k = t / n, l = t - k * n;   // element (k, l) of E^T = E_lk = (Sm_lk - d_k B_lk) / (d_k - d_l)
            const double sm = Sm_[(size_t)k * ld + l], b = bid ? (k == l ? 1.0 : 0.0) : B[(size_t)k * ld + l];
            double e;
            if (k == l) e = 1.0 + 0.5 * (1.0 - b);
            else if (w.oldlab[k] == w.oldlab[l]) e = -0.5 * b;
            else {
                const double num = sm - d[k] * b, gap = d[k] - d[l];
                e = num == 0.0 ? 0.0 : num / gap;
                emax = fmax(emax, fabs(e));
            }
            Sm_[(size_t)k * ld + l] = e;
        }
        emax = block_max(emax, s.red);   // (two barriers: the E matrix is complete afterwards)
        if (!(emax < QX_OA_DIVERGE)) return 0;
        // the eigenvalue estimates of this pass (second order in the remaining error) survive in d = s.emo
        QX_SUB(14);
        tc_left_apply<NT8>(n, Sm_, Ct, ld);
        QX_SUB(15);
        if (!had_cluster && emax < m.oa_stop) return pass;
    }
    return 0;
}

}  // namespace qx
