// stream_bench.cu -- development micro-benchmark: what DRAM bandwidth does the ACCESS PATTERN of
// the subcycle kernel allow, without its arithmetic?  NR planes read, NW planes written, fp64.
//   A: flat 1-D, 8-byte accesses      B: flat 1-D, 16-byte accesses
//   C: marching strips (one column per thread, rows in a loop, next-row prefetch), 8-byte accesses
//   D: like C, but on a strip-tiled layout [strip][row][plane][column]: all planes of a strip row are
//      contiguous and a CTA walks one contiguous region (what an AoSoA plane pool would give)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/stream_bench.cu -o gpurun_out/stream_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
constexpr int NR = 37, NW = 14;
struct P { const double *r[NR]; double *w[NW]; };
__global__ void kA(P p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0;
#pragma unroll
    for (int k = 0; k < NR; ++k) s += __ldg(p.r[k] + i);
#pragma unroll
    for (int k = 0; k < NW; ++k) p.w[k][i] = s + k;
}
__global__ void kB(P p, size_t n2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    double2 s = make_double2(0, 0);
#pragma unroll
    for (int k = 0; k < NR; ++k) { double2 v = __ldg((const double2 *)p.r[k] + i); s.x += v.x; s.y += v.y; }
#pragma unroll
    for (int k = 0; k < NW; ++k) ((double2 *)p.w[k])[i] = make_double2(s.x + k, s.y + k);
}
template <int NT>
__global__ void __launch_bounds__(NT) kC(P p, int nx, int ny, int pitch, int strip, int rows, int ibase = 1) {
    int i = ibase + blockIdx.x * strip + threadIdx.x;
    int j0 = 1 + blockIdx.y * rows, j1 = min(j0 + rows, ny + 1);
    if (threadIdx.x >= strip || i > nx) return;
    double v[NR];
    size_t idx = (size_t)j0 * pitch + i;
#pragma unroll
    for (int k = 0; k < NR; ++k) v[k] = __ldg(p.r[k] + idx);
    for (int j = j0; j < j1; ++j) {
        double nv[NR];
        size_t nidx = (size_t)(j + 1) * pitch + i;
        if (j + 1 < j1) {
#pragma unroll
            for (int k = 0; k < NR; ++k) nv[k] = __ldg(p.r[k] + nidx);
        }
        double s = 0;
#pragma unroll
        for (int k = 0; k < NR; ++k) s += v[k];
        idx = (size_t)j * pitch + i;
#pragma unroll
        for (int k = 0; k < NW; ++k) p.w[k][idx] = s + k;
#pragma unroll
        for (int k = 0; k < NR; ++k) v[k] = nv[k];
    }
}
template <int NT>
__global__ void __launch_bounds__(NT) kD(const double *rd, double *wr, int strip, int rows_total, int rows) {
    // CTA (x, y): strip x, rows y*rows .. ; read region [x][row][NR][strip], write region [x][row][NW][strip]
    const int t = threadIdx.x;
    if (t >= strip) return;
    const int j0 = blockIdx.y * rows, j1 = min(j0 + rows, rows_total);
    const double *r0 = rd + (size_t)blockIdx.x * rows_total * NR * strip;
    double *w0 = wr + (size_t)blockIdx.x * rows_total * NW * strip;
    double v[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) v[k] = __ldg(r0 + ((size_t)j0 * NR + k) * strip + t);
    for (int j = j0; j < j1; ++j) {
        double nv[NR];
        if (j + 1 < j1) {
#pragma unroll
            for (int k = 0; k < NR; ++k) nv[k] = __ldg(r0 + ((size_t)(j + 1) * NR + k) * strip + t);
        }
        double s = 0;
#pragma unroll
        for (int k = 0; k < NR; ++k) s += v[k];
#pragma unroll
        for (int k = 0; k < NW; ++k) w0[((size_t)j * NW + k) * strip + t] = s + k;
#pragma unroll
        for (int k = 0; k < NR; ++k) v[k] = nv[k];
    }
}
int main() {
    const int nx = 1440, ny = 1080, pitch = 1456;
    const size_t cells = (size_t)pitch * (ny + 2);
    double *pool;
    cudaMalloc(&pool, sizeof(double) * cells * (NR + NW));
    cudaMemset(pool, 0, sizeof(double) * cells * (NR + NW));
    P p;
    for (int k = 0; k < NR; ++k) p.r[k] = pool + (size_t)k * cells;
    for (int k = 0; k < NW; ++k) p.w[k] = pool + (size_t)(NR + k) * cells;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = 8.0 * cells * (NR + NW);
    auto time = [&](const char *name, auto launch, double b) {
        for (int w = 0; w < 3; ++w) launch();
        cudaEventRecord(e0);
        const int reps = 20;
        for (int r = 0; r < reps; ++r) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-40s %8.2f us  %8.1f GB/s  (%s)\n", name, 1e3 * ms / reps, b / (ms / reps * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    time("A flat 8B, 256 thr", [&] { kA<<<(unsigned)((cells + 255) / 256), 256>>>(p, cells); }, bytes);
    time("B flat 16B, 256 thr", [&] { kB<<<(unsigned)((cells / 2 + 255) / 256), 256>>>(p, cells / 2); }, bytes);
    const double bc = 8.0 * (double)nx * ny * (NR + NW);
    time("C march 128 thr, strip 120, rows 45", [&] { kC<128><<<dim3(12, 24), 128>>>(p, nx, ny, pitch, 120, 45); }, bc);
    time("C march 128 thr, strip 120, rows 23", [&] { kC<128><<<dim3(12, 47), 128>>>(p, nx, ny, pitch, 120, 23); }, bc);
    time("C march 256 thr, strip 240, rows 45", [&] { kC<256><<<dim3(6, 24), 256>>>(p, nx, ny, pitch, 240, 45); }, bc);
    time("C march 64 thr, strip 60, rows 45", [&] { kC<64><<<dim3(24, 24), 64>>>(p, nx, ny, pitch, 60, 45); }, bc);
    time("C march 128 thr, strip 128 ALIGNED (i0 = 0), rows 45", [&] { kC<128><<<dim3(12, 24), 128>>>(p, nx, ny, pitch, 128, 45, 0); }, bc * 1456.0 / 1440.0);
    time("C march 128 thr, strip 128 aligned+1 (i0 = 1), rows 45", [&] { kC<128><<<dim3(12, 24), 128>>>(p, nx, ny, pitch, 128, 45, 1); }, bc * 1456.0 / 1440.0);
    time("C march 128 thr, strip 112 aligned (i0 = 16), rows 45", [&] { kC<128><<<dim3(13, 24), 128>>>(p, nx, ny, pitch, 112, 45, 16); }, bc * 1456.0 / 1440.0);
    time("C march 128 thr, strip 120, rows 12", [&] { kC<128><<<dim3(12, 90), 128>>>(p, nx, ny, pitch, 120, 12); }, bc);
    // D: the same 12 x 24 CTAs and bytes as the first C line, strip-tiled layout (pool reused: 12 strips x
    // 1080 rows x (37 + 14) planes x 120 columns <= the pool)
    time("D tiled layout, 128 thr, strip 120, rows 45", [&] {
        kD<128><<<dim3(12, 24), 128>>>(pool, pool + (size_t)12 * 1080 * NR * 120, 120, 1080, 45); }, bc);
    time("D tiled layout, 128 thr, strip 128, rows 45", [&] {
        kD<128><<<dim3(11, 24), 128>>>(pool, pool + (size_t)11 * 1080 * NR * 128, 128, 1080, 45); }, bc * (11.0 * 128) / 1440.0);
    time("D tiled layout, 256 thr, strip 240, rows 45", [&] {
        kD<256><<<dim3(6, 24), 256>>>(pool, pool + (size_t)6 * 1080 * NR * 240, 240, 1080, 45); }, bc);
    return 0;
}
