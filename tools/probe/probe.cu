// Hardware probe for the B200 Hessenberg build (not part of the product):
//  (1) cuBLAS DGEMM rates at the shapes of the rank-nb updates  -> FP64 roofline denominator
//  (2) FP64 mma.sync (DMMA) issue-rate microbenchmarks per shape, and DFMA
//  (3) read-only HBM bandwidth
//  (4) GEMV prototype bandwidth
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

// ---------------------------------------------------------------- DMMA microbench
template <int SHAPE>
__global__ void __launch_bounds__(256) dmma_bench(double *out, int iters)
{
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 4; i++) b[i] = 0.5 + threadIdx.x * 1e-9 + i;
    constexpr int NACC = 8;
    double c[NACC][4];
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                    : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
            } else if (SHAPE == 1) {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                    : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            } else if (SHAPE == 2) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                    : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            } else if (SHAPE == 3) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                    : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                    : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                      "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            } else {
                // DFMA: 4 independent chains per accumulator slot
#pragma unroll
                for (int j = 0; j < 4; j++) c[i][j] = fma(a[j], b[j], c[i][j]);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
static void run_dmma(const char *name, double flops_per_warp_inst, int warps_per_cta, int ctas_per_sm)
{
    int iters = 20000;
    int grid = 148 * ctas_per_sm, block = warps_per_cta * 32;
    double *out; CK(cudaMalloc(&out, sizeof(double) * grid * block));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dmma_bench<SHAPE><<<grid, block>>>(out, 100);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    dmma_bench<SHAPE><<<grid, block>>>(out, iters);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = time_ms(e0, e1);
    double flops = (double)grid * warps_per_cta * iters * 8.0 * flops_per_warp_inst;
    printf("dmma %-10s warps/cta=%d ctas/sm=%d : %8.3f ms  %8.2f TFLOP/s\n", name, warps_per_cta, ctas_per_sm, ms, flops / ms / 1e9);
    cudaFree(out);
}

// ---------------------------------------------------------------- read bandwidth
__global__ void __launch_bounds__(256) read_bw(const double2 *__restrict__ p, size_t n2, double *out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 a = p[i], b = p[i + stride], c = p[i + 2 * stride], d = p[i + 3 * stride];
        s0 += a.x + a.y; s1 += b.x + b.y; s2 += c.x + c.y; s3 += d.x + d.y;
    }
    for (; i < n2; i += stride) { double2 a = p[i]; s0 += a.x + a.y; }
    double s = s0 + s1 + s2 + s3;
    if (s == 1.2345) out[0] = s;
}

// ---------------------------------------------------------------- GEMV prototype
// y_part[s][r] = sum_{c in chunk s} A[r, c] * v[c];  A column-major, 16B-aligned base, ld even
template <int Q>   // double2 per thread; CTA of 128 threads covers 256*Q rows
__global__ void __launch_bounds__(128) gemv_proto(const double *__restrict__ A, int ld, int m, int k, int kc,
                                                  const double *__restrict__ v, double *__restrict__ part, int ldp)
{
    extern __shared__ double vs[];
    const int tid = threadIdx.x;
    const int c0 = blockIdx.y * kc;
    const int c1 = min(k, c0 + kc);
    const int nc = c1 - c0;
    for (int c = tid; c < nc; c += 128) vs[c] = v[c0 + c];
    __syncthreads();
    const int rbase = blockIdx.x * (256 * Q) + tid * 2;
    bool ok[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) ok[q] = (rbase + q * 256) < m;
    double2 acc[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) acc[q] = make_double2(0.0, 0.0);
    const double *Ap = A + (size_t)c0 * ld + rbase;
    int c = 0;
    constexpr int U = 4;
    for (; c + U <= nc; c += U) {
        double2 a[U][Q];
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int q = 0; q < Q; q++)
                a[u][q] = ok[q] ? __ldcs((const double2 *)(Ap + (size_t)(c + u) * ld + q * 256)) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < U; u++) {
            double x = vs[c + u];
#pragma unroll
            for (int q = 0; q < Q; q++) { acc[q].x = fma(a[u][q].x, x, acc[q].x); acc[q].y = fma(a[u][q].y, x, acc[q].y); }
        }
    }
    for (; c < nc; c++) {
        double x = vs[c];
#pragma unroll
        for (int q = 0; q < Q; q++) if (ok[q]) {
            double2 a = __ldcs((const double2 *)(Ap + (size_t)c * ld + q * 256));
            acc[q].x = fma(a.x, x, acc[q].x); acc[q].y = fma(a.y, x, acc[q].y);
        }
    }
    double *pp = part + (size_t)blockIdx.y * ldp;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        int r = rbase + q * 256;
        if (r < m) pp[r] = acc[q].x;
        if (r + 1 < m) pp[r + 1] = acc[q].y;
    }
}

template <int Q>
static void run_gemv(const double *A, int ld, int m, int k, const double *v, double *part, int ctas_per_sm, int kcmin)
{
    int rb = (m + 256 * Q - 1) / (256 * Q);
    int slots = 148 * ctas_per_sm;
    int S = slots / rb; if (S < 1) S = 1;
    int kc = (k + S - 1) / S; if (kc < kcmin) kc = kcmin;
    kc = (kc + 3) / 4 * 4;
    S = (k + kc - 1) / kc;
    dim3 grid(rb, S);
    size_t smem = kc * sizeof(double);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) gemv_proto<Q><<<grid, 128, smem>>>(A, ld, m, k, kc, v, part, ld);
    CK(cudaDeviceSynchronize());
    int reps = 10;
    cudaEventRecord(e0);
    for (int w = 0; w < reps; w++) gemv_proto<Q><<<grid, 128, smem>>>(A, ld, m, k, kc, v, part, ld);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = time_ms(e0, e1) / reps;
    printf("gemv Q=%d m=%d k=%d occ=%d grid=(%d,%d) kc=%d : %8.1f us  %8.1f GB/s\n", Q, m, k, ctas_per_sm, rb, S, kc, ms * 1e3,
           (double)m * k * 8 / ms / 1e6);
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs=%d cc=%d.%d L2=%d MB\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor, prop.l2CacheSize >> 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);

    // (2) DMMA microbench
    for (int occ = 1; occ <= 2; occ++) {
        run_dmma<0>("m8n8k4", 2.0 * 8 * 8 * 4, 8, occ);
        run_dmma<1>("m16n8k4", 2.0 * 16 * 8 * 4, 8, occ);
        run_dmma<2>("m16n8k8", 2.0 * 16 * 8 * 8, 8, occ);
        run_dmma<3>("m16n8k16", 2.0 * 16 * 8 * 16, 8, occ);
        run_dmma<4>("dfma", 2.0 * 4 * 32, 8, occ);
    }
    run_dmma<0>("m8n8k4", 2.0 * 8 * 8 * 4, 4, 1);
    run_dmma<3>("m16n8k16", 2.0 * 16 * 8 * 16, 4, 1);
    run_dmma<4>("dfma", 2.0 * 4 * 32, 16, 2);

    // (1) cuBLAS DGEMM
    cublasHandle_t h; cublasCreate(&h);
    {
        struct { int m, n, k; cublasOperation_t ta, tb; const char *name; } shapes[] = {
            {8192, 8192, 8192, CUBLAS_OP_N, CUBLAS_OP_N, "NN 8192^3"},
            {16384, 16384, 312, CUBLAS_OP_N, CUBLAS_OP_T, "NT rank-312 update 16384^2"},
            {16384, 312, 16384, CUBLAS_OP_T, CUBLAS_OP_N, "TN A^T V  16384x312 k=16384"},
            {16384, 312, 16384, CUBLAS_OP_N, CUBLAS_OP_N, "NN Q V    16384x312 k=16384"},
            {4096, 4096, 312, CUBLAS_OP_N, CUBLAS_OP_T, "NT rank-312 update 4096^2"},
            {4096, 312, 4096, CUBLAS_OP_T, CUBLAS_OP_N, "TN 4096x312 k=4096"},
        };
        size_t N = 16384;
        double *A, *B, *C;
        CK(cudaMalloc(&A, N * N * 8)); CK(cudaMalloc(&B, N * N * 8)); CK(cudaMalloc(&C, N * N * 8));
        CK(cudaMemset(A, 0, N * N * 8)); CK(cudaMemset(B, 0, N * N * 8)); CK(cudaMemset(C, 0, N * N * 8));
        for (auto &s : shapes) {
            double alpha = -1.0, beta = 1.0;
            int lda = (int)N, ldb = (int)N, ldc = (int)N;
            for (int w = 0; w < 2; w++) cublasDgemm(h, s.ta, s.tb, s.m, s.n, s.k, &alpha, A, lda, B, ldb, &beta, C, ldc);
            CK(cudaDeviceSynchronize());
            int reps = 5;
            cudaEventRecord(e0);
            for (int w = 0; w < reps; w++) cublasDgemm(h, s.ta, s.tb, s.m, s.n, s.k, &alpha, A, lda, B, ldb, &beta, C, ldc);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms = time_ms(e0, e1) / reps;
            printf("cublasDgemm %-32s : %8.3f ms  %8.2f TFLOP/s\n", s.name, ms, 2.0 * s.m * s.n * s.k / ms / 1e9);
        }
        // sustained: 3 s of 8192^3
        {
            double alpha = 1.0, beta = 0.0; int n = 8192;
            cudaEventRecord(e0);
            int reps = 100;
            for (int w = 0; w < reps; w++) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &alpha, A, (int)N, B, (int)N, &beta, C, (int)N);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms = time_ms(e0, e1) / reps;
            printf("cublasDgemm sustained 100x 8192^3 : %8.3f ms  %8.2f TFLOP/s\n", ms, 2.0 * n * n * (double)n / ms / 1e9);
        }
        // (3) read bandwidth over 2 GiB
        {
            size_t n2 = (size_t)N * N / 2;
            double *out; CK(cudaMalloc(&out, 8));
            for (int g = 2; g <= 8; g *= 2) {
                read_bw<<<148 * g, 256>>>((const double2 *)A, n2, out);
                CK(cudaDeviceSynchronize());
                cudaEventRecord(e0);
                for (int w = 0; w < 5; w++) read_bw<<<148 * g, 256>>>((const double2 *)A, n2, out);
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float ms = time_ms(e0, e1) / 5;
                printf("read_bw grid=148x%d : %8.3f ms  %8.1f GB/s\n", g, ms, (double)N * N * 8 / ms / 1e6);
            }
        }
        // (4) GEMV prototype
        {
            double *v, *part;
            CK(cudaMalloc(&v, N * 8)); CK(cudaMemset(v, 0, N * 8));
            CK(cudaMalloc(&part, N * 1024 * 8));
            int ms_[] = {16384, 16001, 8192, 4096, 2048};
            for (int mm : ms_) {
                for (int occ : {4, 8, 12}) {
                    run_gemv<2>(A, (int)N, mm, mm, v, part, occ, 16);
                    run_gemv<1>(A, (int)N, mm, mm, v, part, occ, 16);
                }
                run_gemv<4>(A, (int)N, mm, mm, v, part, 4, 16);
            }
        }
    }
    printf("probe done\n");
    return 0;
}
