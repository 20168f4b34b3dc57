// Development tool (not part of the product): tile-configuration sweep of the DMMA GEMM (starneig_b200/csrc/dgemm.cuh)
// on the shapes of the Hessenberg updates, next to cublasDgemm on the same operands as the yardstick (cuBLAS is only
// linked here, never by the library). For every shape and configuration it prints the CUDA-event time, TFLOP/s, the
// fraction of the measured DMMA issue peak (37.0 TFLOP/s, profiles/r1_probe_peaks.log) and -- for the first
// repetition -- the largest deviation from the cuBLAS result.
//
// Shapes (n = 20000, nb = 312, panel index p => i = p * nb, m = n - i - 1):
//   NT  C(m x m-w) -= Y(m x w) V(m-w x w)^T             trailing right update       engine.cuh reduce(): GEMM_NT
//   TN  W(m-w x w)  = A(m x m-w)^T VT(m x w)            trailing left update, a     GEMM_TN (+ split-K)
//   NN  W(n x w)    = Q(n x m) VT(m x w)                Q / top-row update, a       GEMM_NN (+ split-K)
// build: make -C tools            run: tools/bin/gemm_sweep [n] [panel] [reps]
#include "../starneig_b200/csrc/dgemm_tma.cuh"
#include <cublas_v2.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

using namespace sb200;

static const double DMMA_PEAK_TFLOPS = 37.0;

struct Shape { char kind; int M, N, K; };     // kind: 'T' = NT, 't' = TN, 'n' = NN

struct Buffers {
    double *A, *B, *C, *Cref, *part;
    int lda, ldb, ldc;
    size_t part_cap;
};

__global__ void k_fill(double *p, size_t count, unsigned seed)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) {
        unsigned long long x = (i + 1) * 6364136223846793005ull + seed * 1442695040888963407ull;
        x ^= x >> 29; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 32;
        p[i] = (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    }
}

__global__ void k_maxdiff(const double *a, const double *b, int rows, int cols, int ld, double *out)
{
    double mx = 0.0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)rows * cols; e += (size_t)gridDim.x * blockDim.x) {
        const size_t r = e % rows, c = e / rows;
        mx = fmax(mx, fabs(a[c * ld + r] - b[c * ld + r]));
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax((unsigned long long *)out, (unsigned long long)__double_as_longlong(mx));
}

// one timed configuration: `launch(splits, klen, stride, out)` issues the GEMM (and nothing else)
template <class Cfg>
static void run_config(const char *name, const Shape &s, Buffers &b, int reps, int splits_req, cudaStream_t st, int raster = 0)
{
    // GEMM_ONLY=substring: run only the configurations whose label contains it (one process per suspect configuration)
    static const char *only = getenv("GEMM_ONLY");
    if (only && !strstr(name, only)) return;
    Cfg::prepare();
    const bool nt = s.kind == 'T';
    int splits = nt ? 1 : std::max(1, splits_req);
    int klen = round_up(std::max(1, ceil_div(s.K + 1, splits)), GEMM_BK);
    splits = std::max(1, ceil_div(s.K + 1, klen));
    while (splits > 1 && (size_t)splits * b.ldc * s.N > b.part_cap) { splits--; klen = round_up(ceil_div(s.K + 1, splits), GEMM_BK); splits = ceil_div(s.K + 1, klen); }
    double *out = splits > 1 ? b.part : b.C;
    const size_t stride = splits > 1 ? (size_t)b.ldc * s.N : 0;
    const double alpha = nt ? -1.0 : 1.0, beta = nt ? 1.0 : 0.0;
    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0)); SB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    double err = -1.0;
    for (int it = 0; it < reps + 1; it++) {
        if (nt) SB_CUDA(cudaMemcpyAsync(b.C, b.Cref + (size_t)b.ldc * s.N, (size_t)b.ldc * s.N * 8, cudaMemcpyDeviceToDevice, st));
        SB_CUDA(cudaEventRecord(e0, st));
        if constexpr (Cfg::IS_TMA) {
            // operands as the engine hands them over: zero guard in front of the K-major workspace operand (k_guard)
            if (!Cfg::launch(st, s.M, s.N, s.K, alpha, b.A, b.lda, b.B, b.ldb, splits > 1 ? 0.0 : beta, out, b.ldc, splits, klen, stride, raster, true)) {
                printf("  %-34s cannot be framed for TMA\n", name);
                return;
            }
        } else
            Cfg::launch(st, s.M, s.N, s.K, alpha, b.A, b.lda, b.B, b.ldb, splits > 1 ? 0.0 : beta, out, b.ldc, splits, klen, stride, raster);
        if (splits > 1) {
            dim3 grid(ceil_div(s.M, 256), s.N);
            splitk_reduce_kernel<<<grid, 256, 0, st>>>(s.M, s.N, splits, out, b.ldc, stride, b.C, b.ldc);
        }
        SB_CUDA(cudaEventRecord(e1, st));
        SB_CUDA(cudaStreamSynchronize(st));
        SB_CUDA(cudaGetLastError());
        float ms = 0.f;
        SB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0) best = std::min(best, ms);
        if (it == 0) {
            double *d;
            SB_CUDA(cudaMalloc(&d, 8)); SB_CUDA(cudaMemset(d, 0, 8));
            k_maxdiff<<<592, 256, 0, st>>>(b.C, b.Cref, s.M, s.N, b.ldc, d);
            SB_CUDA(cudaMemcpy(&err, d, 8, cudaMemcpyDeviceToHost));
            cudaFree(d);
        }
    }
    const double tf = 2.0 * s.M * s.N * (double)s.K / (best * 1e-3) / 1e12;
    printf("  %-34s splits %2d  %8.3f ms  %6.2f TFLOP/s  %5.1f %% of DMMA peak   max|C - C_cublas| = %.2e\n", name, splits, best, tf,
           100.0 * tf / DMMA_PEAK_TFLOPS, err);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

static void run_cublas(cublasHandle_t h, const Shape &s, Buffers &b, int reps, cudaStream_t st)
{
    const bool nt = s.kind == 'T';
    const double alpha = nt ? -1.0 : 1.0, beta = nt ? 1.0 : 0.0;
    const cublasOperation_t ta = s.kind == 't' ? CUBLAS_OP_T : CUBLAS_OP_N, tb = nt ? CUBLAS_OP_T : CUBLAS_OP_N;
    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0)); SB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < reps + 1; it++) {
        // Cref holds [result | pristine C] back to back
        if (nt) SB_CUDA(cudaMemcpyAsync(b.Cref, b.Cref + (size_t)b.ldc * s.N, (size_t)b.ldc * s.N * 8, cudaMemcpyDeviceToDevice, st));
        SB_CUDA(cudaEventRecord(e0, st));
        if (cublasDgemm(h, ta, tb, s.M, s.N, s.K, &alpha, b.A, b.lda, b.B, b.ldb, &beta, b.Cref, b.ldc) != CUBLAS_STATUS_SUCCESS) {
            printf("cublasDgemm failed\n");
            exit(1);
        }
        SB_CUDA(cudaEventRecord(e1, st));
        SB_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        SB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0) best = std::min(best, ms);
    }
    const double tf = 2.0 * s.M * s.N * (double)s.K / (best * 1e-3) / 1e12;
    printf("  %-34s            %8.3f ms  %6.2f TFLOP/s  %5.1f %% of DMMA peak   (yardstick)\n", "cublasDgemm", best, tf, 100.0 * tf / DMMA_PEAK_TFLOPS);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

// <A K-major, B K-major, warps along M, warps along N, 8-row blocks per warp, 8-col blocks per warp, stages, CTAs/SM>
#define CFG(AK, BK_, WM, WN, MB, NB, ST, MINB) GemmConfig<AK, BK_, WM, WN, MB, NB, ST, MINB>
#define RUN(AK, BK_, WM, WN, MB, NB, ST, MINB, SPL)                                                                    \
    run_config<CFG(AK, BK_, WM, WN, MB, NB, ST, MINB)>(#WM "x" #WN " warps, " #MB "x" #NB " blocks, " #ST " st, " #MINB "/SM", s, b, reps, SPL, st)

// the same with the loader options of dgemm.cuh: OPT 1 = interleaved cp.async, 2 = 16-byte cp.async, 3 = both
#define RUNO(OPT, AK, BK_, WM, WN, MB, NB, ST, MINB, SPL)                                                              \
    run_config<GemmConfig<AK, BK_, WM, WN, MB, NB, ST, MINB, OPT>>(#WM "x" #WN " warps, " #MB "x" #NB " blocks, " #ST " st, " #MINB "/SM opt" #OPT, s, b, reps, SPL, st)
#define RUNI(AK, BK_, WM, WN, MB, NB, ST, MINB, SPL) RUNO(1, AK, BK_, WM, WN, MB, NB, ST, MINB, SPL)
// rasterised with the column tiles of one row tile next to each other (what the engine does for the skinny products)
#define RUNR(OPT, AK, BK_, WM, WN, MB, NB, ST, MINB, SPL)                                                              \
    run_config<GemmConfig<AK, BK_, WM, WN, MB, NB, ST, MINB, OPT>>(#WM "x" #WN " warps, " #MB "x" #NB " blocks, " #ST " st, " #MINB "/SM opt" #OPT " raster", s, b, reps, SPL, st, 1)
// tiles moved by TMA (dgemm_tma.cuh); PW 1 = dedicated producer warp + empty barriers, 0 = thread 0 after a CTA barrier
#define RUNK(KBOX, PW, AK, BK_, WM, WN, MB, NB, ST, MINB, SPL, RASTER)                                                 \
    run_config<GemmTmaConfig<AK, BK_, WM, WN, MB, NB, ST, MINB, PW, KBOX>>("TMA kbox" #KBOX " pw" #PW " " #WM "x" #WN " warps, " #MB "x" #NB " blocks, " #ST " st, " #MINB "/SM r" #RASTER, s, b, reps, SPL, st, RASTER)
#define RUNT(PW, AK, BK_, WM, WN, MB, NB, ST, MINB, SPL, RASTER)                                                       \
    run_config<GemmTmaConfig<AK, BK_, WM, WN, MB, NB, ST, MINB, PW>>("TMA pw" #PW " " #WM "x" #WN " warps, " #MB "x" #NB " blocks, " #ST " st, " #MINB "/SM r" #RASTER, s, b, reps, SPL, st, RASTER)

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 20000;
    const int panel = argc > 2 ? atoi(argv[2]) : 2;
    const int reps = argc > 3 ? atoi(argv[3]) : 3;
    const int nb = std::max(64, (int)std::ceil((0.001875596476 * n + 273.5908216) / 8.0) * 8);
    const int i = panel * nb, m = n - i - 1, w = std::min(nb, m);
    if (m < 2 * w) { printf("panel index too large\n"); return 1; }
    const int ld = round_up(n, 16);
    cudaStream_t st;
    SB_CUDA(cudaStreamCreate(&st));
    cublasHandle_t h;
    cublasCreate(&h);
    cublasSetStream(h, st);

    // operands: the big matrix (n x n), two skinny ones (n x w), C/W outputs
    double *big, *sk1, *sk2, *out, *ref, *part;
    const size_t nbig = (size_t)ld * n, nsk = (size_t)ld * round_up(w, 8);
    SB_CUDA(cudaMalloc(&big, nbig * 8)); SB_CUDA(cudaMalloc(&sk1, nsk * 8)); SB_CUDA(cudaMalloc(&sk2, nsk * 8));
    SB_CUDA(cudaMalloc(&out, nbig * 8)); SB_CUDA(cudaMalloc(&ref, 2 * nbig * 8));
    const size_t part_cap = 16 * nsk;
    SB_CUDA(cudaMalloc(&part, part_cap * 8));
    k_fill<<<1184, 256>>>(big, nbig, 1); k_fill<<<1184, 256>>>(sk1, nsk, 2); k_fill<<<1184, 256>>>(sk2, nsk, 3);
    SB_CUDA(cudaDeviceSynchronize());
    printf("gemm_sweep: n = %d, nb = %d, panel %d (i = %d, m = %d, w = %d)\n", n, nb, panel, i, m, w);
    // the skinny operands play V / VT / Y: stored with the row parity of the panel in A, behind a zero guard row
    const int par = (i + 1) & 1;
    SB_CUDA(cudaMemset2D(sk1, (size_t)ld * 8, 0, 8, round_up(w, 8)));
    SB_CUDA(cudaMemset2D(sk2, (size_t)ld * 8, 0, 8, round_up(w, 8)));

    {   // ---- NT: C(m x m-w) -= Y V^T
        Shape s{'T', m, m - w, w};
        Buffers b{sk1, sk2 + par + (w - 1), out + (size_t)w * ld + i + 1, ref, part, ld, ld, ld, part_cap};        // Y, rows w-1.. of V, C = A(i+1:, i+w:)
        k_fill<<<1184, 256>>>(ref + (size_t)ld * s.N, (size_t)ld * s.N, 4);     // pristine C
        printf("NT  C(%d x %d) -= A(%d x %d) B(%d x %d)^T\n", s.M, s.N, s.M, s.K, s.N, s.K);
        run_cublas(h, s, b, reps, st);
        RUNO(1, false, false, 2, 2, 8, 4, 4, 2, 1);     // cp.async product configuration (GemmNT: interleaved cp.async)
        RUN(false, false, 2, 2, 4, 4, 4, 4, 1);         // 64 x 64, 4 CTAs/SM (best cp.async tile of the round-2 sweep)
        RUNT(0, false, false, 2, 2, 8, 4, 4, 2, 1, 0);  // TMA product configuration (TmaNT)
        RUNT(1, false, false, 2, 2, 8, 4, 4, 2, 1, 0);
        RUNT(0, false, false, 2, 2, 8, 4, 3, 2, 1, 0);
        RUNT(0, false, false, 2, 2, 4, 4, 4, 3, 1, 0);  // 64 x 64, 3 CTAs/SM
        RUNT(0, false, false, 2, 2, 4, 4, 3, 4, 1, 0);  // 64 x 64, 4 CTAs/SM
        RUNT(1, false, false, 2, 2, 4, 4, 4, 3, 1, 0);
        RUNT(0, false, false, 2, 4, 8, 4, 4, 1, 1, 0);  // 128 x 128, 8 warps of 64 x 32, 1 CTA/SM
        RUNT(1, false, false, 2, 4, 8, 4, 4, 1, 1, 0);
        RUNT(1, false, false, 2, 4, 8, 4, 3, 1, 1, 0);
        RUNT(1, false, false, 4, 2, 4, 8, 4, 1, 1, 0);  // 128 x 128, 8 warps of 32 x 64 (cuBLAS's d884gemm_128x128 shape)
        RUNT(0, false, false, 4, 2, 4, 4, 4, 2, 1, 0);  // 128 x 64, 8 warps of 32 x 32, 2 CTAs/SM
    }
    {   // ---- TN: W(m-w x w) = A(m x m-w)^T VT(m x w)
        Shape s{'t', m - w, w, m};
        Buffers b{big + (size_t)(i + w) * ld + i + 1, sk1 + par, out, ref, part, ld, ld, ld, part_cap};
        printf("TN  W(%d x %d) = A(%d x %d)^T B(%d x %d)\n", s.M, s.N, s.K, s.M, s.K, s.N);
        run_cublas(h, s, b, reps, st);
        const int t13 = ceil_div(s.M, 64) * ceil_div(s.N, 104);
        const int spl = std::min(32, std::max(1, std::min(ceil_div(8 * 2 * 148, t13), s.K / 512)));
        RUN(true, true, 4, 1, 2, 13, 3, 2, spl);        // cp.async product configuration (GemmTN13), rows fastest
        RUNR(0, true, true, 4, 1, 2, 13, 3, 2, spl);    // ... as the engine launches it
        RUNK(8, 0, true, true, 4, 1, 2, 13, 4, 2, spl, 1);  // K-major tiles in boxes of 8 rows
        RUNK(8, 1, true, true, 4, 1, 2, 13, 4, 2, spl, 1);
        RUNK(8, 0, true, true, 4, 1, 2, 13, 3, 2, spl, 1);
        RUNK(8, 0, true, true, 4, 1, 2, 8, 4, 3, spl, 1);   // 64 x 64
        RUNK(16, 0, true, true, 4, 1, 2, 8, 4, 3, spl, 1);
        RUNK(64, 0, true, true, 4, 1, 2, 8, 4, 3, spl, 1);
        RUNT(0, true, true, 4, 1, 2, 13, 4, 2, spl, 1); // TMA product configuration (TmaTN13)
        RUNT(0, true, true, 4, 1, 2, 13, 4, 2, spl, 0);
        RUNT(1, true, true, 4, 1, 2, 13, 4, 2, spl, 1);
        RUNT(0, true, true, 4, 1, 2, 13, 3, 2, spl, 1);
        RUNT(0, true, true, 4, 1, 2, 13, 5, 2, spl, 1);
        RUNT(0, true, true, 4, 1, 2, 13, 4, 2, 1, 1);   // no split-K
        RUNT(1, true, true, 8, 1, 2, 13, 4, 1, spl, 1); // 128 x 104, 8 warps
        RUNT(0, true, true, 4, 1, 4, 13, 3, 1, spl, 1); // 128 x 104, 4 warps
        RUNT(0, true, true, 4, 1, 2, 10, 4, 2, spl, 1); // 64 x 80
        RUNT(0, true, true, 4, 1, 2, 8, 4, 3, spl, 1);  // 64 x 64
    }
    {   // ---- NN: W(n x w) = Q(n x m) VT(m x w)
        Shape s{'n', n, w, m};
        Buffers b{big + (size_t)(i + 1) * ld, sk1 + par, out, ref, part, ld, ld, ld, part_cap};
        printf("NN  W(%d x %d) = A(%d x %d) B(%d x %d)\n", s.M, s.N, s.M, s.K, s.K, s.N);
        run_cublas(h, s, b, reps, st);
        const int t13 = ceil_div(s.M, 64) * ceil_div(s.N, 104);
        const int spl = std::min(32, std::max(1, std::min(ceil_div(8 * 2 * 148, t13), s.K / 512)));
        RUN(false, true, 4, 1, 2, 13, 3, 2, spl);       // cp.async product configuration (GemmNN13), rows fastest
        RUNR(0, false, true, 4, 1, 2, 13, 3, 2, spl);   // ... as the engine launches it
        RUNK(8, 0, false, true, 4, 1, 2, 13, 4, 2, spl, 1);
        RUNK(8, 1, false, true, 4, 1, 2, 13, 4, 2, spl, 1);
        RUNK(8, 0, false, true, 4, 1, 2, 13, 3, 2, spl, 1);
        RUNK(8, 0, false, true, 4, 1, 2, 8, 4, 3, spl, 1);
        RUNT(0, false, true, 4, 1, 2, 13, 4, 2, spl, 1);// TMA product configuration (TmaNN13)
        RUNT(0, false, true, 4, 1, 2, 13, 4, 2, spl, 0);
        RUNT(1, false, true, 4, 1, 2, 13, 4, 2, spl, 1);
        RUNT(0, false, true, 4, 1, 2, 13, 3, 2, spl, 1);
        RUNT(0, false, true, 4, 1, 2, 13, 5, 2, spl, 1);
        RUNT(1, false, true, 8, 1, 2, 13, 4, 1, spl, 1);// 128 x 104, 8 warps
        RUNT(0, false, true, 4, 1, 4, 13, 3, 1, spl, 1);// 128 x 104, 4 warps
        RUNT(0, false, true, 4, 1, 2, 8, 4, 3, spl, 1); // 64 x 64
    }
    cublasDestroy(h);
    return 0;
}
