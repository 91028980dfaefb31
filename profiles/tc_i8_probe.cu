// tc_i8_probe.cu -- stand-alone bring-up probe for the tcgen05 int8 path (round 2, first tcgen05 GPU call).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tc_i8_probe tc_i8_probe.cu
// 1. correctness: C(128 x N, int32) = A(128 x K, int8) . B(N x K, int8)^T with operands pre-tiled in global memory as
//    the shared-memory image (128-byte-swizzled K-major tiles, or the un-swizzled "interleaved" core-matrix layout),
//    brought in by cp.async.bulk on a 2-stage mbarrier pipeline, accumulated in TMEM, read back with tcgen05.ld.
// 2. rate: MMAs issued back to back on resident shared-memory tiles (no loads): the int8 pipe ceiling per N.
// Every wait has a watchdog (tc05::mbar_wait), so a wrong descriptor cannot hang the device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../tenpy_b200/csrc/tc05.cuh"

using namespace b200::tc05;

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

struct ProbeArgs {
    const int8_t *A;  // [KT][128*128] bytes
    const int8_t *B;  // [KT][N*128] bytes
    int32_t *C;       // 128 x N row-major
    int N, KT;
    int layout;       // 2 = SW128, 0 = interleave
    uint32_t lbo, sbo, kstep;  // descriptor fields (bytes); kstep = start-address advance per K=32 instruction
    int *abort_flag;
};

constexpr int STAGES = 2;

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int N = p.N;
    const uint32_t a_bytes = 128 * 128, b_bytes = N * 128;
    uint8_t *sA[STAGES], *sB[STAGES];
    for (int s = 0; s < STAGES; ++s) {
        sA[s] = smem + s * (a_bytes + b_bytes);
        sB[s] = sA[s] + a_bytes;
    }
    __shared__ uint64_t full[STAGES], empty[STAGES], acc_full;
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t ncols = N < 32 ? 32 : N;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, ncols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_slot;

    if (warp == 0 && lane == 0) {
        // producer
        for (int kt = 0; kt < p.KT; ++kt) {
            int s = kt % STAGES;
            uint32_t ph = (kt / STAGES) & 1;
            if (!mbar_wait(&empty[s], ph ^ 1, p.abort_flag)) break;
            mbar_expect_tx(&full[s], a_bytes + b_bytes);
            bulk_g2s(sA[s], p.A + (size_t)kt * a_bytes, a_bytes, &full[s]);
            bulk_g2s(sB[s], p.B + (size_t)kt * b_bytes, b_bytes, &full[s]);
        }
    } else if (warp == 1 && lane == 0) {
        // MMA issuer
        const uint32_t idesc = idesc_s8(128, N);
        bool ok = true;
        for (int kt = 0; kt < p.KT && ok; ++kt) {
            int s = kt % STAGES;
            uint32_t ph = (kt / STAGES) & 1;
            ok = mbar_wait(&full[s], ph, p.abort_flag);
            if (!ok) break;
            fence_after_sync();
            uint32_t a0 = smem_addr(sA[s]), b0 = smem_addr(sB[s]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint64_t ad = smem_desc(a0 + k * p.kstep, p.lbo, p.sbo, p.layout);
                uint64_t bd = smem_desc(b0 + k * p.kstep, p.lbo, p.sbo, p.layout);
                mma_i8(tmem, ad, bd, idesc, (kt | k) ? 1u : 0u);
            }
            mma_commit(&empty[s]);
        }
        mma_commit(&acc_full);
    }
    __syncwarp();
    // epilogue: all four warps
    bool ok = mbar_wait(&acc_full, 0, p.abort_flag);
    fence_after_sync();
    if (ok) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
            tmem_ld_wait();
            int row = warp * 32 + lane;
#pragma unroll
            for (int j = 0; j < 32; ++j) p.C[(size_t)row * N + c0 + j] = (int32_t)r[j];
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

// rate probe: tiles resident, `iters` x 4 MMAs (K = 32 each) back to back into `nacc` accumulators round robin
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int nacc, int *abort_flag, int32_t *sink, int layout,
                                                      uint32_t lbo, uint32_t sbo, uint32_t kstep, int nk) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem, *sB = smem + 128 * 128;
    __shared__ uint64_t done;
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 * 128 + N * 128) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x01010101u * (i & 3);
    fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        mbar_init(&done, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_slot;
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = idesc_s8(128, N);
        uint32_t a0 = smem_addr(sA), b0 = smem_addr(sB);
        for (int it = 0; it < iters; ++it) {
            uint32_t d = tmem + (it % nacc) * N;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                mma_i8(d, smem_desc(a0 + (k % nk) * kstep, lbo, sbo, layout), smem_desc(b0 + (k % nk) * kstep, lbo, sbo, layout), idesc,
                       it >= nacc ? 1u : (k ? 1u : 0u));
        }
        mma_commit(&done);
    }
    __syncwarp();
    bool ok = mbar_wait(&done, 0, abort_flag, 1u << 28);
    fence_after_sync();
    if (ok && blockIdx.x == 0) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), r);
        tmem_ld_wait();
        sink[threadIdx.x] = (int32_t)r[0];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static uint32_t interleave_offset(uint32_t r, uint32_t kb) {  // core matrices [row group][k chunk], 128 B each
    return (r / 8) * 1024 + (kb / 16) * 128 + (r % 8) * 16 + (kb % 16);
}

static int run_case(const char *name, int N, int KT, int layout, uint32_t lbo, uint32_t sbo, uint32_t kstep) {
    const int M = 128, K = KT * 128;
    std::vector<int8_t> A((size_t)M * K), B((size_t)N * K);
    srand(1234 + N + KT);
    for (auto &x : A) x = (int8_t)(rand() % 129 - 64);
    for (auto &x : B) x = (int8_t)(rand() % 129 - 64);
    std::vector<int8_t> At((size_t)M * K), Bt((size_t)N * K);
    for (int kt = 0; kt < KT; ++kt) {
        for (int r = 0; r < M; ++r)
            for (int kb = 0; kb < 128; ++kb) {
                uint32_t o = layout == 2 ? sw128_offset(r, kb) : interleave_offset(r, kb);
                At[(size_t)kt * M * 128 + o] = A[(size_t)r * K + kt * 128 + kb];
            }
        for (int r = 0; r < N; ++r)
            for (int kb = 0; kb < 128; ++kb) {
                uint32_t o = layout == 2 ? sw128_offset(r, kb) : interleave_offset(r, kb);
                Bt[(size_t)kt * N * 128 + o] = B[(size_t)r * K + kt * 128 + kb];
            }
    }
    int8_t *dA, *dB;
    int32_t *dC;
    int *dflag;
    CK(cudaMalloc(&dA, At.size()));
    CK(cudaMalloc(&dB, Bt.size()));
    CK(cudaMalloc(&dC, (size_t)M * N * 4));
    CK(cudaMalloc(&dflag, 4));
    CK(cudaMemset(dflag, 0, 4));
    CK(cudaMemset(dC, 0xff, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, At.data(), At.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bt.data(), Bt.size(), cudaMemcpyHostToDevice));
    ProbeArgs p{dA, dB, dC, N, KT, layout, lbo, sbo, kstep, dflag};
    size_t smem = STAGES * (128 * 128 + N * 128) + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    int flag = -1;
    if (e == cudaSuccess) cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost);
    std::vector<int32_t> C((size_t)M * N);
    if (e == cudaSuccess) cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    int first_r = -1, first_c = -1;
    if (e == cudaSuccess && flag == 0) {
        for (int r = 0; r < M; ++r)
            for (int c = 0; c < N; ++c) {
                int32_t ref = 0;
                for (int k = 0; k < K; ++k) ref += (int32_t)A[(size_t)r * K + k] * (int32_t)B[(size_t)c * K + k];
                if (ref != C[(size_t)r * N + c]) {
                    if (!bad) first_r = r, first_c = c;
                    ++bad;
                }
            }
    }
    printf("case %-28s N=%3d KT=%2d layout=%d lbo=%4u sbo=%4u kstep=%3u : cuda=%s watchdog=%d mismatches=%ld/%d first=(%d,%d)\n",
           name, N, KT, layout, lbo, sbo, kstep, cudaGetErrorString(e), flag, bad, M * N, first_r, first_c);
    fflush(stdout);
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dC);
    cudaFree(dflag);
    return (e == cudaSuccess && flag == 0 && bad == 0) ? 0 : 1;
}

static int run_rate(int N, int nacc, int ctas, int layout = 2, uint32_t lbo = 16, uint32_t sbo = 1024, uint32_t kstep = 32, int nk = 4,
                    const char *tag = "sw128") {
    int *dflag;
    int32_t *sink;
    CK(cudaMalloc(&dflag, 4));
    CK(cudaMemset(dflag, 0, 4));
    CK(cudaMalloc(&sink, 512));
    size_t smem = 128 * 128 + N * 128 + 1024;
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rate_kernel<<<ctas, 128, smem>>>(N, 200, nacc, dflag, sink, layout, lbo, sbo, kstep, nk);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    rate_kernel<<<ctas, 128, smem>>>(N, iters, nacc, dflag, sink, layout, lbo, sbo, kstep, nk);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int flag;
    cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost);
    double macs = (double)ctas * iters * 4 * 128.0 * N * 32.0;
    printf("rate %-22s N=%3d nacc=%d ctas=%3d : %.3f ms  %.1f Tops/s (2*MAC)  watchdog=%d\n", tag, N, nacc, ctas, ms,
           2 * macs / ms * 1e-9, flag);
    fflush(stdout);
    cudaFree(dflag);
    cudaFree(sink);
    return 0;
}


// ---- fragment layout of tcgen05.ld.16x256b: write lane*1000 + column with 32x32b stores, read back with 16x256b.x2 ----
__global__ void __launch_bounds__(128, 1) ldshape_kernel(int32_t *out) {
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&tmem_base_slot, 64);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_slot;
    {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (uint32_t)((warp * 32 + lane) * 1000 + j);
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
            "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(tmem + ((uint32_t)(warp * 32) << 16)),
            "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
            "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
            "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
            "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    // every warp reads its own lane quarter, lower (h = 0) and upper (h = 1) 16 lanes: 16x256b.x2 = 16 columns, 8 registers
    for (int h = 0; h < 2; ++h) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t)(warp * 32 + h * 16) << 16))
                     : "memory");
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) out[((warp * 2 + h) * 32 + lane) * 8 + j] = (int32_t)r[j];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

static int run_ldshape() {
    int32_t *d;
    CK(cudaMalloc(&d, 4 * 2 * 32 * 8 * 4));
    ldshape_kernel<<<1, 128>>>(d);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> h(4 * 2 * 32 * 8);
    CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
    // expected (mma m16n8 accumulator layout per 8-column group g): r[4g+0], r[4g+1] = row lane/4, cols 8g + 2(lane%4) + {0,1};
    // r[4g+2], r[4g+3] = row lane/4 + 8, same columns
    long bad = 0;
    for (int w = 0; w < 4; ++w)
        for (int hh = 0; hh < 2; ++hh)
            for (int l = 0; l < 32; ++l)
                for (int j = 0; j < 8; ++j) {
                    int g = j / 4, i = j % 4;
                    int row = w * 32 + hh * 16 + l / 4 + (i >= 2 ? 8 : 0), col = 8 * g + 2 * (l % 4) + (i & 1);
                    if (h[(((w * 2 + hh) * 32 + l) * 8) + j] != row * 1000 + col) ++bad;
                }
    printf("ldshape 16x256b.x2: %ld of %d registers differ from the m16n8 accumulator layout\n", bad, 4 * 2 * 32 * 8);
    for (int l = 0; l < 8; ++l) {
        printf("  warp 1, lower half, lane %d:", l);
        for (int j = 0; j < 8; ++j) printf(" %d", h[(((1 * 2 + 0) * 32 + l) * 8) + j]);
        printf("\n");
    }
    cudaFree(d);
    return bad ? 1 : 0;
}

int main(int argc, char **argv) {
    int fails = 0;
    const char *mode = argc > 1 ? argv[1] : "all";
    if (!strcmp(mode, "all") || !strcmp(mode, "sw128")) {
        fails += run_case("sw128", 128, 1, 2, 16, 1024, 32);
        fails += run_case("sw128", 128, 4, 2, 16, 1024, 32);
        fails += run_case("sw128", 64, 3, 2, 16, 1024, 32);
        fails += run_case("sw128", 256, 5, 2, 16, 1024, 32);
        fails += run_case("sw128", 32, 2, 2, 16, 1024, 32);
    }
    if (!strcmp(mode, "all") || !strcmp(mode, "interleave")) {
        run_case("interleave lbo=K sbo=MN", 128, 2, 0, 128, 1024, 256);
        run_case("interleave lbo=MN sbo=K", 128, 2, 0, 1024, 128, 256);
    }
    if (!strcmp(mode, "all") || !strcmp(mode, "rate")) {
        int ns[] = {64, 128, 256};
        for (int N : ns) {
            run_rate(N, 1, 148);
            run_rate(N, 2, 148);
        }
        run_rate(128, 4, 148);
        run_rate(256, 1, 296);
        // un-swizzled core-matrix layouts and the 64-byte swizzle at N = 128 (the Ozaki kernel's tile shape)
        run_rate(128, 4, 148, 0, 128, 1024, 256, 4, "interleave 128/1024");
        run_rate(128, 4, 148, 0, 2048, 128, 4096, 2, "interleave 2048/128");
        run_rate(128, 4, 148, 4, 16, 512, 32, 2, "sw64");
        run_rate(128, 4, 148, 6, 16, 256, 32, 1, "sw32");
    }
    if (!strcmp(mode, "all") || !strcmp(mode, "ldshape")) run_ldshape();
    printf("probe %s (%d failing sw128 cases)\n", fails ? "FAILED" : "ok", fails);
    return fails ? 1 : 0;
}
