// coop_microbench.cu -- measurement behind DESIGN section 9's "thread per env, not warp per env": the motor-row sweep of the solver
// (9 rows per sweep: di = clamp(rhs - dv[D] / Mdd, -mx - app, mx - app); dv += Minv[:, D] di) written both ways on the same data:
//
//   thread  one env per thread, M^-1 / rows / dv in registers, rows unrolled (what step_kernel does);
//   coop    one env per 8-lane group (4 envs per warp, the north star's sketch): lane k owns dv[k] (lane 0 also dv[8]) and row k of M^-1,
//           a row broadcasts dv[D] with one shuffle, every lane computes di redundantly and updates its own component.
//
// Reports (a) the latency of one sweep for a single resident warp (clock64) and (b) the throughput of a full grid, in env-sweeps per
// microsecond.  Same arithmetic, same operation order per component, so both produce bit-identical dv (checked).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/coop_microbench scripts/coop_microbench.cu && gpurun_out/coop_microbench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int ND = 9, SWEEPS = 50;
struct EnvData { float Minv[ND][ND], rhs[ND], mx[ND]; };      // AoS on purpose for the host; the kernels read it once

__global__ void thread_kernel(const EnvData* __restrict__ in, float* __restrict__ out, long long* __restrict__ cyc, int n, int reps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float Mi[ND][ND], rhs[ND], mx[ND], invD[ND], dv[ND], app[ND];
#pragma unroll
    for (int a = 0; a < ND; a++) {
#pragma unroll
        for (int b = 0; b < ND; b++) Mi[a][b] = in[i].Minv[a][b];
        rhs[a] = in[i].rhs[a]; mx[a] = in[i].mx[a]; invD[a] = 1.0f / Mi[a][a];
    }
    const long long t0 = clock64();
    float acc = 0.f;
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int a = 0; a < ND; a++) { dv[a] = 0.f; app[a] = 0.f; }
        for (int s = 0; s < SWEEPS; s++) {
#pragma unroll
            for (int D = 0; D < ND; D++) {
                const float di = fminf(fmaxf(rhs[D] - dv[D] * invD[D], -mx[D] - app[D]), mx[D] - app[D]);
                app[D] += di;
#pragma unroll
                for (int k = 0; k < ND; k++) dv[k] += Mi[k][D] * di;
            }
        }
#pragma unroll
        for (int a = 0; a < ND; a++) acc += dv[a];
        rhs[0] += 1e-7f * acc;          // keep the repetitions dependent
    }
    if (cyc && threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
#pragma unroll
    for (int a = 0; a < ND; a++) out[(size_t)i * ND + a] = dv[a];
}

// 8 lanes per env: lane k (0..7) owns dv[k], app/rhs/mx/invD of row k and row k of M^-1; lane 0 additionally owns component 8
__global__ void coop_kernel(const EnvData* __restrict__ in, float* __restrict__ out, long long* __restrict__ cyc, int n, int reps) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, i = t >> 3, k = threadIdx.x & 7, lane0 = (threadIdx.x & 31) & ~7;
    if (i >= n) return;     // n is a multiple of 4 per warp in this benchmark
    float Mk[ND], M8[ND], rhs_k = in[i].rhs[k], mx_k = in[i].mx[k], rhs8 = in[i].rhs[8], mx8 = in[i].mx[8];
#pragma unroll
    for (int b = 0; b < ND; b++) { Mk[b] = in[i].Minv[k][b]; M8[b] = in[i].Minv[8][b]; }
    const float invD_k = 1.0f / Mk[k], invD8 = 1.0f / M8[8];
    const long long t0 = clock64();
    float acc = 0.f, dvk = 0.f, dv8 = 0.f;
    for (int r = 0; r < reps; r++) {
        dvk = 0.f; dv8 = 0.f;
        float app_k = 0.f, app8 = 0.f;
        for (int s = 0; s < SWEEPS; s++) {
#pragma unroll
            for (int D = 0; D < 8; D++) {
                // the row's owner holds rhs / app / bounds: it computes di, one shuffle broadcasts it (a second one would be needed if every
                // lane recomputed di from a broadcast dv[D]; broadcasting di is the shorter chain)
                float di = fminf(fmaxf(rhs_k - dvk * invD_k, -mx_k - app_k), mx_k - app_k);
                di = __shfl_sync(0xffffffffu, di, lane0 + D);
                if (k == D) app_k += di;
                dvk += Mk[D] * di; dv8 += M8[D] * di;
            }
            {   // row 8 lives in lane 0 of the group (all lanes carry dv8 redundantly: no shuffle needed)
                const float di = fminf(fmaxf(rhs8 - dv8 * invD8, -mx8 - app8), mx8 - app8);
                app8 += di;
                dvk += Mk[8] * di; dv8 += M8[8] * di;
            }
        }
        acc += dvk;
        rhs_k += 1e-7f * acc;
    }
    if (cyc && threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
    out[(size_t)i * ND + k] = dvk;
    if (k == 0) out[(size_t)i * ND + 8] = dv8;
}

int main() {
    const int n = 65536;
    std::vector<EnvData> h(n);
    srand(1);
    for (int i = 0; i < n; i++) {
        // a symmetric positive definite, diagonally dominant "M^-1" and motor rows in the product's value range
        for (int a = 0; a < ND; a++) for (int b = 0; b <= a; b++) { float v = (a == b) ? 2.0f + (rand() % 100) * 0.02f : ((rand() % 200) - 100) * 0.002f; h[i].Minv[a][b] = h[i].Minv[b][a] = v; }
        for (int a = 0; a < ND; a++) { h[i].rhs[a] = ((rand() % 200) - 100) * 0.001f; h[i].mx[a] = 0.05f + (rand() % 100) * 0.002f; }
    }
    EnvData* d_in; float *d_a, *d_b; long long* d_c;
    cudaMalloc(&d_in, n * sizeof(EnvData)); cudaMalloc(&d_a, (size_t)n * ND * 4); cudaMalloc(&d_b, (size_t)n * ND * 4); cudaMalloc(&d_c, 8192 * 8);
    cudaMemcpy(d_in, h.data(), n * sizeof(EnvData), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timed = [&](auto launch) { launch(); cudaDeviceSynchronize(); cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); return ms; };
    // (a) latency: one warp on the whole GPU
    long long c = 0;
    thread_kernel<<<1, 32>>>(d_in, d_a, d_c, 32, 20); cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
    printf("thread: one warp (32 envs), %d sweeps x 20: %.1f cycles per sweep (9 rows) = %.1f per row\n", SWEEPS, c / (20.0 * SWEEPS), c / (20.0 * SWEEPS * 9));
    coop_kernel<<<1, 32>>>(d_in, d_b, d_c, 4, 20); cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
    printf("coop:   one warp (4 envs),  %d sweeps x 20: %.1f cycles per sweep (9 rows) = %.1f per row\n", SWEEPS, c / (20.0 * SWEEPS), c / (20.0 * SWEEPS * 9));
    // (b) throughput: all 65,536 envs, 20 solves of 50 sweeps each
    const int reps = 20;
    for (int bs : {128, 256}) {
        float ms = timed([&] { thread_kernel<<<(n + bs - 1) / bs, bs>>>(d_in, d_a, nullptr, n, reps); });
        printf("thread: %d envs, block %d: %.3f ms -> %.1f env-sweeps per us\n", n, bs, ms, (double)n * reps * SWEEPS / (ms * 1e3));
        ms = timed([&] { coop_kernel<<<(n * 8 + bs - 1) / bs, bs>>>(d_in, d_b, nullptr, n, reps); });
        printf("coop:   %d envs, block %d: %.3f ms -> %.1f env-sweeps per us\n", n, bs, ms, (double)n * reps * SWEEPS / (ms * 1e3));
    }
    // same results?  (one solve each: the repetitions above perturb rhs differently in the two kernels)
    thread_kernel<<<(n + 127) / 128, 128>>>(d_in, d_a, nullptr, n, 1); coop_kernel<<<(n * 8 + 127) / 128, 128>>>(d_in, d_b, nullptr, n, 1); cudaDeviceSynchronize();
    std::vector<float> a((size_t)n * ND), b((size_t)n * ND);
    cudaMemcpy(a.data(), d_a, a.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), d_b, b.size() * 4, cudaMemcpyDeviceToHost);
    size_t diff = 0; for (size_t i = 0; i < a.size(); i++) diff += a[i] != b[i];
    printf("components that differ between the two mappings: %zu of %zu\n", diff, a.size());
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, thread_kernel); printf("registers: thread %d", fa.numRegs); cudaFuncGetAttributes(&fa, coop_kernel); printf(", coop %d\n", fa.numRegs);
    return 0;
}
