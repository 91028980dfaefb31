// block_qr_core.cuh -- Householder QR of one row-major block, written as barrier-separated PHASES (one CTA per block).
//
// Replaces, per charge block, the LAPACK geqrf/orgqr pair behind the reference's npc.qr (np_conserved.py:4139,
// `np.linalg.qr(block, mode)`).  A (m x n, row-major) is overwritten by R (k x n upper triangular in its first k rows,
// k = min(m, n)), Q (m x k) is formed explicitly; the diagonal of R is made non-negative (the unique factorisation the
// reference returns for pos_diag_R=True).  Unblocked algorithm (dgeqr2 + dorg2r): right for the small and medium blocks of
// charge-conserving tensors, one launch per Array with no host round trip; big dense blocks want the blocked (compact WY)
// multi-CTA version -- round-2 work together with the QR preconditioning of the Jacobi SVD.
//
// Threads own COLUMNS (row-major storage: for a fixed row, consecutive threads touch consecutive addresses); the only
// cross-thread quantity per Householder step is the squared norm of the pivot column below the diagonal, summed in a
// fixed order from per-thread partials (deterministic).  Every phase is a plain function of (tid, nthreads, pointers):
// the CUDA kernel runs it per thread between __syncthreads(), tests/csrc/block_qr_host.cpp runs it for tid = 0..T-1.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define BQ_HD __host__ __device__ __forceinline__
#else
#define BQ_HD inline
#endif

namespace b200 {
namespace bqr {

// PHASE 1: partial[tid] = sum over rows r > j (strided by threads) of A[r][j]^2
BQ_HD void col_partial(int tid, int T, const double *A, int m, int n, int j, double *partial) {
    double s = 0.0;
    for (int r = j + 1 + tid; r < m; r += T) {
        const double a = A[(int64_t)r * n + j];
        s = fma(a, a, s);
    }
    partial[tid] = s;
}

// PHASE 2 (one thread): reflector H = 1 - tau v v^T with v[j] = 1 that maps the pivot column to beta e_j.
// params[0] = tau, params[1] = scale (v[r] = A[r][j] * scale for r > j), params[2] = beta
BQ_HD void reflector(int T, const double *A, int n, int j, const double *partial, double *params) {
    double sigma = 0.0;
    for (int t = 0; t < T; ++t) sigma += partial[t];
    const double alpha = A[(int64_t)j * n + j];
    if (sigma == 0.0) {
        params[0] = 0.0;
        params[1] = 0.0;
        params[2] = alpha;
    } else {
        const double nrm = sqrt(alpha * alpha + sigma);
        const double beta = alpha >= 0.0 ? -nrm : nrm;
        params[0] = (beta - alpha) / beta;
        params[1] = 1.0 / (alpha - beta);
        params[2] = beta;
    }
}

// PHASE 3: store the reflector: V[r][j] = A[r][j] * scale (r > j), V[j][j] = 1, V[r][j] = 0 (r < j); A[r][j] = 0 below
// the diagonal, A[j][j] = beta.  V is m x k row-major.
BQ_HD void store_reflector(int tid, int T, double *A, double *V, int m, int n, int k, int j, const double *params) {
    const double scale = params[1];
    for (int r = tid; r < m; r += T) {
        double v = 0.0;
        if (r == j) {
            v = 1.0;
            A[(int64_t)r * n + j] = params[2];
        } else if (r > j) {
            v = A[(int64_t)r * n + j] * scale;
            A[(int64_t)r * n + j] = 0.0;
        }
        V[(int64_t)r * k + j] = v;
    }
}

// PHASE 4: apply H to the columns c > j of X (ld = ncol; X = A with c0 = j + 1, or X = Q with c0 = j): every thread owns
// columns c = c0 + tid, c0 + tid + T, ...:  w = sum_{r >= j} V[r][j] X[r][c];  X[r][c] -= tau V[r][j] w
BQ_HD void apply_reflector(int tid, int T, double *X, int ncol, const double *V, int m, int k, int j, int c0, double tau) {
    if (tau == 0.0) return;
    for (int c = c0 + tid; c < ncol; c += T) {
        double w = 0.0;
        for (int r = j; r < m; ++r) w = fma(V[(int64_t)r * k + j], X[(int64_t)r * ncol + c], w);
        w *= tau;
        for (int r = j; r < m; ++r) X[(int64_t)r * ncol + c] = fma(-V[(int64_t)r * k + j], w, X[(int64_t)r * ncol + c]);
    }
}

// PHASE Q0: Q = first k columns of the identity
BQ_HD void init_q(int tid, int T, double *Q, int m, int k) {
    for (int64_t e = tid; e < (int64_t)m * k; e += T) Q[e] = (e / k == e % k) ? 1.0 : 0.0;
}

// PHASE S: make diag(R) non-negative: rows i of R (= A) and columns i of Q with R[i][i] < 0 change sign.
// sign[i] must have been written by the previous phase (sign_of_diag).
BQ_HD void sign_of_diag(int tid, int T, const double *A, int n, int k, double *sign) {
    for (int i = tid; i < k; i += T) sign[i] = A[(int64_t)i * n + i] < 0.0 ? -1.0 : 1.0;
}
BQ_HD void flip_signs(int tid, int T, double *A, double *Q, int m, int n, int k, const double *sign) {
    for (int64_t e = tid; e < (int64_t)k * n; e += T) {
        const int i = (int)(e / n);
        if (sign[i] < 0.0) A[e] = -A[e];
    }
    for (int64_t e = tid; e < (int64_t)m * k; e += T) {
        const int i = (int)(e % k);
        if (sign[i] < 0.0) Q[e] = -Q[e];
    }
}

}  // namespace bqr
}  // namespace b200
