// FP64 micro-benchmarks for the B200 (sm_100a): the denominators behind the "FP64-bound" kernels of this repository
// (pm_pet_fast_kernel, kge_pop_pass_kernel) and the per-instruction latencies behind the MRTM latency model.
//   throughput: DFMA, DADD + DMUL issued separately (what -fmad=false code executes), DADD alone
//   latency   : dependent chains of DADD, DMUL, DFMA, a 64-bit SHFL.IDX exchange (2 x SHFL), LOP3, and the chain of
//               one MRTM sub-step (DMUL -> DADD -> DMUL -> 2 x SHFL -> XOR -> NT x DADD -> DADD), one warp per SM
// Built by tools/microbench/Makefile into tools/microbench/libxan_microbench.so; bench.py loads it with ctypes
// (xan_mb_fp64) and reports the numbers next to the kernels' measured rates.  Not part of libxanthos_b200.so.
#include <cuda_runtime.h>
#include <cstdio>

namespace {

constexpr int CHAINS = 8;

template <int MODE>   // 0 = DFMA, 1 = DMUL + DADD (no contraction), 2 = DADD
__global__ void __launch_bounds__(1024) throughput_kernel(double *out, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) x[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            if (MODE == 0) x[k] = __fma_rn(x[k], a, b);
            else if (MODE == 1) x[k] = __dadd_rn(__dmul_rn(x[k], a), b);
            else x[k] = __dadd_rn(x[k], b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += x[k];
    if (s == 12345.678) out[0] = s;   // never true: keeps the loop alive
}

// dependent chain, one warp per block; cycles per operation from clock64 of lane 0
template <int MODE, int NT = 1>   // 0 DADD, 1 DMUL, 2 DFMA, 3 SHFL pair, 4 LOP (xor), 5 MRTM sub-step chain with NT row terms
__global__ void __launch_bounds__(32) latency_kernel(double *out, long long *cyc, int iters, double a, double b, int key) {
    double x = (double)(threadIdx.x + 1) * 1e-3;
    const unsigned full = 0xffffffffu;
    const int src = (threadIdx.x + 1) & 31;
    double S = x, erl = 1e-3;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) x = __dadd_rn(x, b);
        else if (MODE == 1) x = __dmul_rn(x, a);
        else if (MODE == 2) x = __fma_rn(x, a, b);
        else if (MODE == 3) {
            const int lo = __shfl_sync(full, __double2loint(x), src), hi = __shfl_sync(full, __double2hiint(x), src);
            x = __hiloint2double(hi, lo);
        } else if (MODE == 4) {
            x = __longlong_as_double(__double_as_longlong(x) ^ (long long)(key + i));
        } else {
            // d -> d * dt -> S + . -> . * tauinv -> gather (2 SHFL + sign flip) -> nt dependent adds -> + erl
            const double ddt = __dmul_rn(x, a);
            S = __dadd_rn(S, ddt);
            const double F = __dmul_rn(S, b);
            const int lo = __shfl_sync(full, __double2loint(F), src);
            const int hi = __shfl_sync(full, __double2hiint(F), src) ^ (int)0x80000000u;
            double d = __hiloint2double(hi, lo);
#pragma unroll
            for (int s = 1; s < NT; ++s) d = __dadd_rn(d, F);
            x = __dadd_rn(d, erl);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (x == 12345.678) out[0] = x + S;
}

template <typename K, typename... A>
float time_ms(K kernel, dim3 grid, dim3 block, int reps, A... args) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kernel<<<grid, block>>>(args...);   // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        kernel<<<grid, block>>>(args...);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

}  // namespace

// out[0..2]  : TFLOP/s of DFMA (2 flops each), of DMUL + DADD pairs (2 flops per pair), of DADD alone (1 flop each)
// out[3..2+k]: cycles per dependent DADD, DMUL, DFMA, 64-bit shuffle exchange, 64-bit XOR
// out[8..16] : cycles of one MRTM sub-step chain for NT = 1..9
// out[17]    : SM clock (MHz) derived from clock64 against the event time of the DADD chain
// returns 0 on success
extern "C" int xan_mb_fp64(double *out, int n_out) {
    if (!out || n_out < 18) return -1;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -2;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *d_out = nullptr;
    long long *d_cyc = nullptr;
    if (cudaMalloc(&d_out, 64) != cudaSuccess || cudaMalloc(&d_cyc, sizeof(long long) * 1024) != cudaSuccess) return -2;
    const int iters = 4096;
    const dim3 grid(sms * 2), block(1024);
    const double ops = (double)grid.x * block.x * (double)iters * CHAINS;
    const float t0 = time_ms(throughput_kernel<0>, grid, block, 5, d_out, iters, 1.0000001, 1e-9);
    const float t1 = time_ms(throughput_kernel<1>, grid, block, 5, d_out, iters, 1.0000001, 1e-9);
    const float t2 = time_ms(throughput_kernel<2>, grid, block, 5, d_out, iters, 1.0000001, 1e-9);
    out[0] = 2.0 * ops / (t0 * 1e-3) / 1e12;
    out[1] = 2.0 * ops / (t1 * 1e-3) / 1e12;
    out[2] = 1.0 * ops / (t2 * 1e-3) / 1e12;
    const int lit = 20000;
    long long c = 0;
    auto lat = [&](auto kernel, int nt) {
        kernel<<<1, 32>>>(d_out, d_cyc, lit, 1.0000001, 0.9999999, nt);
        cudaDeviceSynchronize();
        kernel<<<1, 32>>>(d_out, d_cyc, lit, 1.0000001, 0.9999999, nt);
        cudaMemcpy(&c, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost);
        return (double)c / lit;
    };
    out[3] = lat(latency_kernel<0, 1>, 0);
    out[4] = lat(latency_kernel<1, 1>, 0);
    out[5] = lat(latency_kernel<2, 1>, 0);
    out[6] = lat(latency_kernel<3, 1>, 0);
    out[7] = lat(latency_kernel<4, 1>, 3);
    out[8] = lat(latency_kernel<5, 1>, 0);
    out[9] = lat(latency_kernel<5, 2>, 0);
    out[10] = lat(latency_kernel<5, 3>, 0);
    out[11] = lat(latency_kernel<5, 4>, 0);
    out[12] = lat(latency_kernel<5, 5>, 0);
    out[13] = lat(latency_kernel<5, 6>, 0);
    out[14] = lat(latency_kernel<5, 7>, 0);
    out[15] = lat(latency_kernel<5, 8>, 0);
    out[16] = lat(latency_kernel<5, 9>, 0);
    // clock: a long DADD chain timed with events against its own clock64 count
    {
        const int it2 = 2000000;
        const float ms = time_ms(latency_kernel<0, 1>, dim3(1), dim3(32), 3, d_out, d_cyc, it2, 1.0, 1e-9, 0);
        cudaMemcpy(&c, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost);
        out[17] = (double)c / (ms * 1e-3) / 1e6;
    }
    cudaFree(d_out);
    cudaFree(d_cyc);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

#ifdef XAN_MB_MAIN
int main() {
    double o[18];
    const int rc = xan_mb_fp64(o, 18);
    if (rc) {
        printf("{\"error\": %d}\n", rc);
        return 1;
    }
    printf("{\"dfma_tflops\": %.3f, \"dmul_dadd_tflops\": %.3f, \"dadd_tflops\": %.3f, \"lat_dadd\": %.2f, \"lat_dmul\": %.2f, "
           "\"lat_dfma\": %.2f, \"lat_shfl64\": %.2f, \"lat_xor64\": %.2f, \"mrtm_chain_cycles_nt1_9\": [%.1f, %.1f, %.1f, %.1f, "
           "%.1f, %.1f, %.1f, %.1f, %.1f], \"sm_mhz\": %.0f}\n",
           o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9], o[10], o[11], o[12], o[13], o[14], o[15], o[16], o[17]);
    return 0;
}
#endif
