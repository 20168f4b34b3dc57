// Development micro-benchmark of the panel kernels (not part of the product): back-to-back launches with
// fixed arguments, CUDA-event timed.
#include "../starneig_b200/csrc/panel.cuh"
#include <vector>
#include <algorithm>
using namespace sb200;

__global__ void k_empty(PanelArgs a, int j) {}
__global__ void k_fence_only(PanelArgs a, int j)
{
    if (last_block_done(a.counter, gridDim.x)) { if (threadIdx.x == 0) a.w2[0] = 1.0; }
}

int main(int argc, char **argv)
{
    int m = argc > 1 ? atoi(argv[1]) : 1500, j = argc > 2 ? atoi(argv[2]) : 150;
    int ld = (m + 63) / 64 * 64, nbp = 320;
    PanelArgs a{};
    a.m = m; a.ld = ld;
    auto dalloc = [](size_t n) { double *p; SB_CUDA(cudaMalloc(&p, n * 8)); SB_CUDA(cudaMemset(p, 0, n * 8)); return p; };
    a.V = dalloc((size_t)ld * nbp); a.Y = dalloc((size_t)ld * nbp); a.VT = dalloc((size_t)ld * nbp);
    a.pcol = dalloc(ld); a.ldp = ld; a.ypart = dalloc((size_t)ld * 256);
    a.s = dalloc(nbp); a.w2 = dalloc(nbp); a.colpart = dalloc((size_t)PANEL_LDB * nbp); a.ldt = nbp; a.sqpart = dalloc(PANEL_LDB);
    SB_CUDA(cudaMalloc(&a.scal, nbp * sizeof(ColScal))); SB_CUDA(cudaMemset(a.scal, 0, nbp * sizeof(ColScal)));
    SB_CUDA(cudaMalloc(&a.counter, 16)); SB_CUDA(cudaMemset(a.counter, 0, 16));
    double *acol = dalloc(ld);
    SB_CUDA(cudaFuncSetAttribute(k_col_finish_update<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SB_CUDA(cudaFuncSetAttribute(k_col_reflector<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

    TileGeom tg;
    tg.nsub = std::max(1, (m + 32 * 148 - 1) / (32 * 148));
    int blocks = (m + 32 * tg.nsub - 1) / (32 * tg.nsub);
    tg.NW = std::max(1, (j + 31) / 32);
    tg.RS = std::max(1, std::min(tg.nsub, 16 / tg.NW));
    while (tg.NW * tg.RS < 4 && tg.RS < 4) tg.RS++;
    size_t smem_fu = (size_t)(2 * j + tg.nsub * 4 * tg.NW * 32 + 2 * tg.nsub * 32 + tg.RS * tg.NW * 32) * 8;
    size_t smem_rf = (size_t)(j + tg.nsub * tg.NW * 32 + tg.nsub * 32 + tg.RS * tg.NW * 32 + 32) * 8;
    int threads = 32 * tg.NW * tg.RS;
    printf("m=%d j=%d blocks=%d threads=%d nsub=%d NW=%d RS=%d smem_fu=%zu smem_rf=%zu\n", m, j, blocks, threads, tg.nsub, tg.NW, tg.RS, smem_fu, smem_rf);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 200;
    auto timeit = [&](const char *name, auto launch) {
        for (int i = 0; i < 10; i++) launch();
        SB_CUDA(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < reps; i++) launch();
        cudaEventRecord(e1);
        SB_CUDA(cudaDeviceSynchronize());
        SB_CUDA(cudaGetLastError());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("  %-34s %8.2f us per launch\n", name, 1e3 * ms / reps);
    };
    timeit("empty kernel", [&] { k_empty<<<blocks, threads>>>(a, j); });
    timeit("fence+atomic only", [&] { k_fence_only<<<blocks, threads>>>(a, j); });
    timeit("finish_update (do_update=1)", [&] { k_col_finish_update<512><<<blocks, threads, smem_fu>>>(a, j, 20, acol, 1, tg); });
    timeit("finish_update (do_update=0)", [&] { k_col_finish_update<512><<<blocks, threads, smem_fu>>>(a, j, 20, acol, 0, tg); });
    timeit("reflector", [&] { k_col_reflector<512><<<blocks, threads, smem_rf>>>(a, j, acol, tg); });
    timeit("fu + rf alternating", [&] { k_col_finish_update<512><<<blocks, threads, smem_fu>>>(a, j, 20, acol, 1, tg); k_col_reflector<512><<<blocks, threads, smem_rf>>>(a, j, acol, tg); });
    int RB = (m + 255) / 256, S = std::max(1, 1480 / RB); int ncols = m - j; int kc = std::max(16, (ncols + S - 1) / S); kc = (kc + 3) / 4 * 4; S = (ncols + kc - 1) / kc;
    double *Amat = dalloc((size_t)ld * (m + 8));
    timeit("gemv", [&] { k_col_gemv<<<RB * S, GEMV_THREADS, kc * 8>>>(a, j, ncols, Amat, ld, 0, kc, RB, acol); });
    timeit("fu + rf + gemv", [&] { k_col_finish_update<512><<<blocks, threads, smem_fu>>>(a, j, S, acol, 1, tg); k_col_reflector<512><<<blocks, threads, smem_rf>>>(a, j, acol, tg);
                                   k_col_gemv<<<RB * S, GEMV_THREADS, kc * 8>>>(a, j, ncols, Amat, ld, 0, kc, RB, acol); });
    return 0;
}
