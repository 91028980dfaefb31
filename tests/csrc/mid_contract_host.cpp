// Host check of tenpy_b200/csrc/mid_contract_core.cuh (test infrastructure): the per-thread body of mid_contract_kernel is
// run for every (o, i) column and compared with the plain triple loop OUT[o,n,i] = sum_k M[n,k] T[o,k,i].
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "../../tenpy_b200/csrc/mid_contract_core.cuh"

template <int KMAX>
static double run(int K, int N, int64_t outer, int64_t inner, std::mt19937_64 &rng) {
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<double> M((size_t)N * K), T((size_t)outer * K * inner), OUT((size_t)outer * N * inner, -7.0);
    for (auto &x : M) x = nd(rng);
    for (auto &x : T) x = nd(rng);
    for (int64_t o = 0; o < outer; ++o)
        for (int64_t i = 0; i < inner; ++i) b200::midc::column<KMAX>(o, i, K, N, inner, M.data(), T.data(), OUT.data());
    double err = 0.0;
    for (int64_t o = 0; o < outer; ++o)
        for (int n = 0; n < N; ++n)
            for (int64_t i = 0; i < inner; ++i) {
                double s = 0.0;
                for (int k = 0; k < K; ++k) s += M[(size_t)n * K + k] * T[((size_t)o * K + k) * inner + i];
                err = std::fmax(err, std::fabs(s - OUT[((size_t)o * N + n) * inner + i]));
            }
    return err;
}

template <int KMAX>
static double run2(int K1, int K2, int N1, int N2, int64_t outer, int64_t inner, std::mt19937_64 &rng) {
    std::normal_distribution<double> nd(0.0, 1.0);
    const int K = K1 + K2, N = N1 + N2;
    std::vector<double> M((size_t)N * K), T1((size_t)outer * K1 * inner + 1), T2((size_t)outer * K2 * inner + 1),
        O1((size_t)outer * N1 * inner + 1, -7.0), O2((size_t)outer * N2 * inner + 1, -7.0);
    for (auto &x : M) x = nd(rng);
    for (auto &x : T1) x = nd(rng);
    for (auto &x : T2) x = nd(rng);
    for (int64_t o = 0; o < outer; ++o)
        for (int64_t i = 0; i < inner; ++i)
            b200::midc::column2<KMAX>(o, i, K1, K2, N1, N2, inner, M.data(), T1.data(), T2.data(), O1.data(), O2.data());
    double err = 0.0;
    for (int64_t o = 0; o < outer; ++o)
        for (int n = 0; n < N; ++n)
            for (int64_t i = 0; i < inner; ++i) {
                double s = 0.0;
                for (int k = 0; k < K; ++k)
                    s += M[(size_t)n * K + k] * (k < K1 ? T1[((size_t)o * K1 + k) * inner + i] : T2[((size_t)o * K2 + (k - K1)) * inner + i]);
                const double got = n < N1 ? O1[((size_t)o * N1 + n) * inner + i] : O2[((size_t)o * N2 + (n - N1)) * inner + i];
                err = std::fmax(err, std::fabs(s - got));
            }
    return err;
}

int main() {
    std::mt19937_64 rng(7);
    int bad = 0;
    const int cases[][4] = {{12, 12, 5, 7}, {1, 1, 1, 1}, {16, 3, 2, 33}, {7, 20, 4, 9}, {20, 20, 3, 17}, {32, 5, 2, 8}, {17, 40, 2, 5}};
    for (auto &c : cases) {
        double e = c[0] <= 16 ? run<16>(c[0], c[1], c[2], c[3], rng) : run<32>(c[0], c[1], c[2], c[3], rng);
        printf("K=%d N=%d outer=%d inner=%d  err %.2e\n", c[0], c[1], c[2], c[3], e);
        if (!(e < 1e-13)) ++bad;
    }
    const int cases2[][6] = {{8, 4, 8, 4, 5, 7}, {8, 4, 4, 8, 3, 9}, {12, 0, 12, 0, 2, 5}, {0, 5, 3, 2, 2, 4}, {16, 4, 16, 4, 2, 33},
                             {20, 12, 7, 0, 2, 6}};
    for (auto &c : cases2) {
        double e = (c[0] + c[1]) <= 16 ? run2<16>(c[0], c[1], c[2], c[3], c[4], c[5], rng)
                                       : run2<32>(c[0], c[1], c[2], c[3], c[4], c[5], rng);
        printf("K=%d+%d N=%d+%d outer=%d inner=%d  err %.2e\n", c[0], c[1], c[2], c[3], c[4], c[5], e);
        if (!(e < 1e-13)) ++bad;
    }
    printf("%s\n", bad ? "FAILED" : "ok");
    return bad ? 1 : 0;
}
