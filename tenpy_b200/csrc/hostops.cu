// hostops.cu -- host-side integer bookkeeping of the C ABI (no device code) and the device self test.
//
// Host twins of the integer helpers of the reference's native module tenpy/linalg/_npc_helper.pyx:
//   _find_row_differences pyx:635, the np.lexsort calls of _tensordot_pre_sort pyx:1357-1377,
//   ChargeInfo_make_valid pyx:478, _map_blocks pyx:732.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "common.cuh"

using namespace b200;

extern "C" int b200_find_row_differences(const int64_t *rows, int64_t n, int64_t width, int64_t *out, int64_t *n_out) {
    if (n < 0 || width < 0 || !out || !n_out) return set_error(B200_ERR_ARG, "bad arguments");
    int64_t cnt = 0;
    if (n == 0) {
        out[cnt++] = 0;
        *n_out = cnt;
        return B200_OK;
    }
    out[cnt++] = 0;
    for (int64_t i = 1; i < n; ++i) {
        const int64_t *x = rows + (i - 1) * width, *y = rows + i * width;
        bool diff = false;
        for (int64_t c = 0; c < width; ++c)
            if (x[c] != y[c]) {
                diff = true;
                break;
            }
        if (diff) out[cnt++] = i;
    }
    out[cnt++] = n;
    *n_out = cnt;
    return B200_OK;
}

extern "C" int b200_lexsort_rows(const int64_t *rows, int64_t n, int64_t width, int64_t *perm) {
    if (n < 0 || width < 0 || !perm) return set_error(B200_ERR_ARG, "bad arguments");
    std::iota(perm, perm + n, (int64_t)0);
    if (width == 0) return B200_OK;
    std::stable_sort(perm, perm + n, [=](int64_t i, int64_t j) {
        const int64_t *x = rows + i * width, *y = rows + j * width;
        for (int64_t c = width - 1; c >= 0; --c)
            if (x[c] != y[c]) return x[c] < y[c];
        return false;
    });
    return B200_OK;
}

extern "C" int b200_make_valid(int64_t *charges, int64_t n, int64_t qnumber, const int64_t *mod) {
    if (n < 0 || qnumber < 0) return set_error(B200_ERR_ARG, "bad arguments");
    for (int64_t c = 0; c < qnumber; ++c) {
        const int64_t md = mod[c];
        if (md <= 0) return set_error(B200_ERR_ARG, "mod must be > 0");
        if (md == 1) continue;
        for (int64_t i = 0; i < n; ++i) {
            int64_t v = charges[i * qnumber + c] % md;
            if (v < 0) v += md;
            charges[i * qnumber + c] = v;
        }
    }
    return B200_OK;
}

extern "C" int b200_map_blocks(const int64_t *blocksizes, int64_t n, int64_t *out) {
    int64_t at = 0;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < blocksizes[i]; ++j) out[at++] = i;
    return B200_OK;
}

// ---- device self test --------------------------------------------------------------------------------
namespace {
__global__ void dmma_test_kernel(const double *A, const double *B, double *C8, double *C4) {
    // one warp: C(16x8) = A(16x8) * B(8x8), row-major inputs
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double a[4] = {A[g * 8 + t], A[(g + 8) * 8 + t], A[g * 8 + t + 4], A[(g + 8) * 8 + t + 4]};
    double b[2] = {B[t * 8 + g], B[(t + 4) * 8 + g]};
    double c8[4] = {0, 0, 0, 0}, c4[4] = {0, 0, 0, 0};
    dmma_16x8x8_k8(c8, a, b);
    dmma_16x8x8_k4(c4, a, b);
    C8[g * 8 + 2 * t] = c8[0];
    C8[g * 8 + 2 * t + 1] = c8[1];
    C8[(g + 8) * 8 + 2 * t] = c8[2];
    C8[(g + 8) * 8 + 2 * t + 1] = c8[3];
    C4[g * 8 + 2 * t] = c4[0];
    C4[g * 8 + 2 * t + 1] = c4[1];
    C4[(g + 8) * 8 + 2 * t] = c4[2];
    C4[(g + 8) * 8 + 2 * t + 1] = c4[3];
}

double gemm_check(int m, int n, int k1, int k2) {
    // C = A1 B1 + A2 B2 through b200_grouped_gemm_f64, compared with a host triple loop
    std::vector<double> A((size_t)m * (k1 + k2)), B((size_t)(k1 + k2) * n), C((size_t)m * n), R((size_t)m * n, 0.0);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        return (double)(s % 2000001) / 1000000.0 - 1.0;
    };
    for (auto &x : A) x = rnd();
    for (auto &x : B) x = rnd();
    // layout: A1 (m x k1) at 0, A2 (m x k2) at m*k1 ; B1 (k1 x n) at 0, B2 at k1*n
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int l = 0; l < k1; ++l) acc += A[(size_t)i * k1 + l] * B[(size_t)l * n + j];
            for (int l = 0; l < k2; ++l) acc += A[(size_t)m * k1 + (size_t)i * k2 + l] * B[(size_t)k1 * n + (size_t)l * n + j];
            R[(size_t)i * n + j] = acc;
        }
    double *dA, *dB, *dC;
    if (cudaMalloc(&dA, A.size() * 8) != cudaSuccess) return -1.0;
    cudaMalloc(&dB, B.size() * 8);
    cudaMalloc(&dC, C.size() * 8);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
    int64_t mm = m, nn = n, coff = 0, pptr[2] = {0, 2}, kk[2] = {k1, k2}, ao[2] = {0, (int64_t)m * k1},
            bo[2] = {0, (int64_t)k1 * n};
    int rc = b200_grouped_gemm_f64(1, &mm, &nn, &coff, pptr, 2, kk, ao, bo, dA, dB, dC, nullptr);
    double err = -1.0;
    if (rc == B200_OK) {
        cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
        err = 0.0;
        for (size_t i = 0; i < C.size(); ++i) err = std::max(err, std::fabs(C[i] - R[i]));
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dC);
    return err;
}
}  // namespace

extern "C" int b200_selftest(double *out) {
    if (!out) return set_error(B200_ERR_ARG, "out is NULL");
    double hA[128], hB[64], hR[128], h8[128], h4[128];
    for (int i = 0; i < 128; ++i) hA[i] = std::sin(0.37 * i) + 0.01 * i;
    for (int i = 0; i < 64; ++i) hB[i] = std::cos(0.91 * i) - 0.02 * i;
    for (int i = 0; i < 16; ++i)
        for (int j = 0; j < 8; ++j) {
            double acc = 0.0;
            for (int l = 0; l < 8; ++l) acc += hA[i * 8 + l] * hB[l * 8 + j];
            hR[i * 8 + j] = acc;
        }
    double *dA, *dB, *d8, *d4;
    B200_CUDA_CHECK(cudaMalloc(&dA, sizeof(hA)));
    B200_CUDA_CHECK(cudaMalloc(&dB, sizeof(hB)));
    B200_CUDA_CHECK(cudaMalloc(&d8, sizeof(h8)));
    B200_CUDA_CHECK(cudaMalloc(&d4, sizeof(h4)));
    B200_CUDA_CHECK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice));
    B200_CUDA_CHECK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
    dmma_test_kernel<<<1, 32>>>(dA, dB, d8, d4);
    B200_CHECK_LAUNCH();
    B200_CUDA_CHECK(cudaMemcpy(h8, d8, sizeof(h8), cudaMemcpyDeviceToHost));
    B200_CUDA_CHECK(cudaMemcpy(h4, d4, sizeof(h4), cudaMemcpyDeviceToHost));
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(d8);
    cudaFree(d4);
    double e8 = 0.0, e4 = 0.0;
    for (int i = 0; i < 128; ++i) {
        e8 = std::max(e8, std::fabs(h8[i] - hR[i]));
        e4 = std::max(e4, std::fabs(h4[i] - hR[i]));
    }
    out[0] = e8;
    out[1] = e4;
    out[2] = gemm_check(200, 136, 48, 22);  // 128-tile config, vector path
    out[3] = gemm_check(61, 33, 7, 19);     // small tiles, scalar path
    return B200_OK;
}
