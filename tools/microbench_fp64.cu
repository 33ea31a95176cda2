// Micro-measurements that the multi-lane R.rho.R design rests on (B200, sm_100a).  Stand-alone binary:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench_fp64 tools/microbench_fp64.cu
// Prints cycles per dependent operation for a lone warp (latencies), cycles per instruction with several
// independent chains (issue cadence), the cost of exchanging doubles between lanes through shuffles and through
// shared memory, and whether a DMMA (mma.m8n8k4.f64) accumulates like a chain of FMAs in k order (bitwise).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e = (x);                                                          \
        if (e != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

constexpr int kReps = 2048;

// mode 0: one dependent DFMA chain; 1: 2 chains; 2: 4 chains; 3: 8 chains; 4: dependent DADD; 5: dependent DMUL
// 6: dependent shuffle of a double (2 SHFL); 7: smem round trip (STS.64, syncwarp, LDS.64 from the neighbour lane)
// 8: dependent DMMA chain; 9: 4 independent DMMA accumulators; 10: rcp.approx.ftz.f64 dependent; 11: 16 chains
// 12: smem all-gather of 4 doubles in a quad (1 STS.64 + syncwarp + 2 LDS.128); 13: same by 3 xor shuffles
__global__ void k_lat(int mode, double* out, long long* cyc, double seed) {
    __shared__ __align__(16) double sm[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = seed + 0.001 * (threadIdx.x + i);
    const double m = 0.999999, c = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
#pragma unroll 16
        for (int i = 0; i < kReps; ++i) a[0] = fma(a[0], m, c);
    } else if (mode == 1) {
#pragma unroll 8
        for (int i = 0; i < kReps; ++i) { a[0] = fma(a[0], m, c); a[1] = fma(a[1], m, c); }
    } else if (mode == 2) {
#pragma unroll 4
        for (int i = 0; i < kReps; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = fma(a[j], m, c);
    } else if (mode == 3) {
#pragma unroll 2
        for (int i = 0; i < kReps; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma(a[j], m, c);
    } else if (mode == 11) {
#pragma unroll 1
        for (int i = 0; i < kReps; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = fma(a[j], m, c);
    } else if (mode == 4) {
#pragma unroll 16
        for (int i = 0; i < kReps; ++i) a[0] = __dadd_rn(a[0], c);
    } else if (mode == 5) {
#pragma unroll 16
        for (int i = 0; i < kReps; ++i) a[0] = __dmul_rn(a[0], m);
    } else if (mode == 6) {
#pragma unroll 16
        for (int i = 0; i < kReps; ++i) a[0] = __shfl_xor_sync(0xffffffffu, a[0], 1);
    } else if (mode == 7) {
#pragma unroll 4
        for (int i = 0; i < kReps; ++i) {
            sm[warp][lane] = a[0];
            __syncwarp();
            a[0] = sm[warp][lane ^ 1];
            __syncwarp();
        }
    } else if (mode == 8) {
#pragma unroll 16
        for (int i = 0; i < kReps; ++i) dmma(a[0], a[1], m, m);
    } else if (mode == 9) {
#pragma unroll 4
        for (int i = 0; i < kReps; ++i) {
            dmma(a[0], a[1], m, m); dmma(a[2], a[3], m, m); dmma(a[4], a[5], m, m); dmma(a[6], a[7], m, m);
        }
    } else if (mode == 10) {
#pragma unroll 16
        for (int i = 0; i < kReps; ++i) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(a[0]));
    } else if (mode == 12) {
#pragma unroll 4
        for (int i = 0; i < kReps; ++i) {
            sm[warp][lane] = a[0];
            __syncwarp();
            const double2 u = *reinterpret_cast<const double2*>(&sm[warp][(lane & ~3)]);
            const double2 v = *reinterpret_cast<const double2*>(&sm[warp][(lane & ~3) + 2]);
            __syncwarp();
            a[0] = (u.x + u.y) + (v.x + v.y);
        }
    } else if (mode == 13) {
#pragma unroll 4
        for (int i = 0; i < kReps; ++i) {
            const double x1 = __shfl_xor_sync(0xffffffffu, a[0], 1);
            const double x2 = __shfl_xor_sync(0xffffffffu, a[0], 2);
            const double x3 = __shfl_xor_sync(0xffffffffu, a[0], 3);
            a[0] = (a[0] + x1) + (x2 + x3);
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 1234.5678) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// DMMA bitwise check: D = A(8x4) B(4x8) + C against three scalar evaluation orders.
__global__ void k_dmma_check(const double* A, const double* B, const double* C, double* D, int n) {
    const int lane = threadIdx.x;
    for (int t = 0; t < n; ++t) {
        const double a = A[t * 32 + (lane >> 2) * 4 + (lane & 3)];
        const double b = B[t * 32 + (lane & 3) * 8 + (lane >> 2)];
        double d0 = C[t * 64 + (lane >> 2) * 8 + 2 * (lane & 3)], d1 = C[t * 64 + (lane >> 2) * 8 + 2 * (lane & 3) + 1];
        dmma(d0, d1, a, b);
        D[t * 64 + (lane >> 2) * 8 + 2 * (lane & 3)] = d0;
        D[t * 64 + (lane >> 2) * 8 + 2 * (lane & 3) + 1] = d1;
    }
}

// the full single-thread iteration cost model: N dependent "phases" of W independent FMAs (ILP = W)
__global__ void k_ilp(int ilp, double* out, long long* cyc, double seed) {
    double a[32];
    for (int i = 0; i < 32; ++i) a[i] = seed + 0.001 * (threadIdx.x + i);
    const double m = 0.999999, c = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    if (ilp == 32) {
#pragma unroll 1
        for (int i = 0; i < kReps; ++i)
#pragma unroll
            for (int j = 0; j < 32; ++j) a[j] = fma(a[j], m, c);
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 32; ++i) s += a[i];
    if (s == 1234.5678) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    CK(cudaMalloc(&out, 1024));
    CK(cudaMalloc(&cyc, 1024 * sizeof(long long)));
    const char* names[] = {"DFMA dependent (1 chain)", "DFMA 2 chains", "DFMA 4 chains", "DFMA 8 chains", "DADD dependent",
                           "DMUL dependent", "shuffle double dependent (2 SHFL)", "smem round trip STS/syncwarp/LDS",
                           "DMMA m8n8k4 dependent", "DMMA 4 independent accumulators", "MUFU.RCP64H dependent",
                           "DFMA 16 chains", "quad all-gather of 4 doubles via smem (+3 adds)",
                           "quad all-gather of 4 doubles via 3 shuffles (+3 adds)"};
    const int per_iter[] = {1, 2, 4, 8, 1, 1, 1, 1, 1, 4, 1, 16, 1, 1};
    for (int threads : {32, 64, 128, 256}) {
        printf("--- %d threads per CTA (1 CTA): warps per scheduler = %g\n", threads, threads / 128.0);
        for (int mode = 0; mode < 14; ++mode) {
            k_lat<<<1, threads>>>(mode, out, cyc, 1.0);
            CK(cudaDeviceSynchronize());
            k_lat<<<1, threads>>>(mode, out, cyc, 1.0);
            CK(cudaDeviceSynchronize());
            long long c;
            CK(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
            printf("%-60s %8.2f cycles per loop trip, %6.2f per instruction\n", names[mode], (double)c / kReps,
                   (double)c / kReps / per_iter[mode]);
        }
        k_ilp<<<1, threads>>>(32, out, cyc, 1.0);
        CK(cudaDeviceSynchronize());
        long long c;
        CK(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
        printf("%-60s %8.2f cycles per loop trip, %6.2f per instruction\n", "DFMA 32 chains", (double)c / kReps,
               (double)c / kReps / 32);
    }

    // DMMA accumulation order
    const int n = 4096;
    std::vector<double> A(n * 32), B(n * 32), C(n * 64), D(n * 64);
    srand(7);
    auto rnd = [] { return (rand() / (double)RAND_MAX - 0.5) * exp2((double)(rand() % 40 - 20)); };
    for (auto& v : A) v = rnd();
    for (auto& v : B) v = rnd();
    for (auto& v : C) v = rnd();
    double *dA, *dB, *dC, *dD;
    CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dB, B.size() * 8));
    CK(cudaMalloc(&dC, C.size() * 8)); CK(cudaMalloc(&dD, D.size() * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice));
    k_dmma_check<<<1, 32>>>(dA, dB, dC, dD, n);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost));
    long same_fwd = 0, same_rev = 0, same_pair = 0, same_prodfirst = 0, total = 0;
    for (int t = 0; t < n; ++t)
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) {
                const double* a = &A[t * 32 + i * 4];
                auto b = [&](int k) { return B[t * 32 + k * 8 + j]; };
                const double c = C[t * 64 + i * 8 + j], d = D[t * 64 + i * 8 + j];
                double f = c;
                for (int k = 0; k < 4; ++k) f = __builtin_fma(a[k], b(k), f);
                double r = c;
                for (int k = 3; k >= 0; --k) r = __builtin_fma(a[k], b(k), r);
                const double p = c + (__builtin_fma(a[0], b(0), a[1] * b(1)) + __builtin_fma(a[2], b(2), a[3] * b(3)));
                double q = a[0] * b(0);
                for (int k = 1; k < 4; ++k) q = __builtin_fma(a[k], b(k), q);
                q += c;
                same_fwd += (memcmp(&f, &d, 8) == 0);
                same_rev += (memcmp(&r, &d, 8) == 0);
                same_pair += (memcmp(&p, &d, 8) == 0);
                same_prodfirst += (memcmp(&q, &d, 8) == 0);
                ++total;
            }
    printf("DMMA vs scalar orders over %ld outputs: fma chain k=0..3 onto C: %ld | k=3..0: %ld | pairwise: %ld | products first, C last: %ld\n",
           total, same_fwd, same_rev, same_pair, same_prodfirst);
    return 0;
}
