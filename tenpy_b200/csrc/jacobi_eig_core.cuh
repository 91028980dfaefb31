// jacobi_eig_core.cuh -- the 32x32 pivot eigenproblem of one Jacobi pair, written as barrier-separated PHASES.
//
// The pivot step of the block-Jacobi SVD/eigh (svd.cu) diagonalises G = P P^T (32x32, symmetric) by a parallel cyclic
// two-sided Jacobi in shared memory.  ncu (profiles/r01d_launch_shares.md) shows that this step is the latency floor of
// the whole decomposition: 100-116 us per round against 6 us (Gram) + 10 us (apply).  Version 1 (jacobi_eig_kernel)
// spends three barriers per rotation set: parameters | rows of G | columns of G and Q.  Version 2 applies the 16
// disjoint rotations of a set from both sides in ONE pass over a double-buffered G:
//
//     G'[i][j] = a_i a_j G[i][j] + a_i b_j G[i][pj] + b_i a_j G[pi][j] + b_i b_j G[pi][pj],   Q'[r][j] = a_j Q[r][j] + b_j Q[r][pj]
//
// (pi = partner of i in the current pairing; (a, b) = (c, -s) for the lower and (c, +s) for the upper index of a pair),
// i.e. two barriers per rotation set and no read-after-write hazard inside a phase.
//
// Every phase is a plain function of (thread index, shared arrays): the CUDA kernel calls it per thread between
// __syncthreads(), the host test (tests/csrc/eig_core_host.cpp, run by tests/test_eig_core_host.py) calls it for
// tid = 0..T-1 sequentially -- the same code is checked on a CPU before it ever runs on the GPU.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define JE_HD __host__ __device__ __forceinline__
#else
#define JE_HD inline
#endif

namespace b200 {
namespace jeig {

constexpr int N = 32;        // order of the pivot problem (= JP)
constexpr int LD = N + 1;    // shared-memory row stride (JLDG)
constexpr int NPAIR = N / 2;

JE_HD double je_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / std::sqrt(x);
#endif
}

// pair `t` (0..15) of rotation set `step` (0..N-2): round-robin tournament, p < q
JE_HD void pair_of(int step, int t, int &p, int &q) {
    int a, b;
    if (t == 0) {
        a = N - 1;
        b = step;
    } else {
        a = (step + t) % (N - 1);
        b = (step - t + (N - 1)) % (N - 1);
    }
    p = a < b ? a : b;
    q = a < b ? b : a;
}

// rotation (c, s) annihilating G[p][q]; identity if the element is negligible or a row is deflated.  Returns 1 if rotated.
JE_HD int rotation(double gpp, double gqq, double gpq, double defl2, double tol_in, double &c, double &s) {
    c = 1.0;
    s = 0.0;
    const double lim = tol_in * sqrt(fabs(gpp * gqq));
    if (fabs(gpq) > lim && gpp > defl2 && gqq > defl2) {
        const double aa = gqq - gpp, bb = 2.0 * gpq;
        const double hh = sqrt(aa * aa + bb * bb);
        const double tt = (aa >= 0.0) ? bb / (aa + hh) : bb / (aa - hh);
        c = je_rsqrt(1.0 + tt * tt);
        s = tt * c;
        return 1;
    }
    return 0;
}

// PHASE A (threads t < NPAIR): parameters of rotation set `step` from the current G
JE_HD int phase_params(int t, int step, const double *G, double defl2, double tol_in, int *partner, double *alpha,
                       double *beta) {
    int p, q;
    pair_of(step, t, p, q);
    double c, s;
    const int rot = rotation(G[p * LD + p], G[q * LD + q], G[p * LD + q], defl2, tol_in, c, s);
    partner[p] = q;
    partner[q] = p;
    alpha[p] = c;
    beta[p] = -s;   // new_p = c old_p - s old_q
    alpha[q] = c;
    beta[q] = s;    // new_q = s old_p + c old_q
    return rot;
}

// PHASE B (element e of N*N, any thread): two-sided update of G and one-sided update of Q, old -> new buffers
JE_HD void phase_apply_elem(int e, const double *Gold, double *Gnew, const double *Qold, double *Qnew,
                            const int *partner, const double *alpha, const double *beta) {
    const int i = e / N, j = e - i * N;
    const int pi = partner[i], pj = partner[j];
    const double ai = alpha[i], bi = beta[i], aj = alpha[j], bj = beta[j];
    const double top = aj * Gold[i * LD + j] + bj * Gold[i * LD + pj];
    const double bot = aj * Gold[pi * LD + j] + bj * Gold[pi * LD + pj];
    Gnew[i * LD + j] = ai * top + bi * bot;
    Qnew[i * LD + j] = aj * Qold[i * LD + j] + bj * Qold[i * LD + pj];
}

}  // namespace jeig
}  // namespace b200
