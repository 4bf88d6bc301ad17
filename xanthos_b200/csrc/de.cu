// Differential evolution on the device: the population, the energies and the whole
// generation logic of scipy.optimize.differential_evolution as the reference uses it (calibrate_abcd.py:103-110:
// strategy best1bin, Latin-hypercube init, dither in [0.5, 1), recombination 0.7, tol 0.01, no polish) stay in
// HBM, so that a calibration generation is "trial kernel -> objective kernels -> selection kernel" with no host
// round trip.  Many independent problems (basins) advance in lock step, one thread block per problem; updating is
// deferred (a whole generation is evaluated as one batch), as in the host driver of calibrate/calibrate_abcd.py.
//
// Random numbers: Philox4x32-10, counter = (generation, problem, member, stream), key = seed - every draw is a
// pure function of its coordinates, so a run is reproducible whatever the launch geometry.
#include "common.cuh"

#include <math_constants.h>

namespace xan {

struct Philox {
    unsigned k0, k1;
    __device__ __forceinline__ uint4 operator()(unsigned c0, unsigned c1, unsigned c2, unsigned c3) const {
        unsigned a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ a;
            c1 = lo1;
            c2 = hi0 ^ c3 ^ b;
            c3 = lo0;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

__device__ __forceinline__ double u01(unsigned hi, unsigned lo) {   // uniform in [0, 1), 53 bits
    return (double)((((unsigned long long)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

enum { STREAM_INIT = 1, STREAM_PICK = 2, STREAM_CROSS = 3, STREAM_OOB = 4, STREAM_DITHER = 5 };

// Latin hypercube in [0, 1]^D: per (problem, dimension) the S strata are visited in a random order.
// thread = (problem, dimension); Fisher-Yates over the S strata in local memory (S <= DE_MAX_S).
constexpr int DE_MAX_S = 256;
constexpr int DE_MAX_D = 8;

__global__ void __launch_bounds__(64)
    de_init_kernel(double *__restrict__ pop, int n, int S, int D, unsigned seed_lo, unsigned seed_hi) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * D) return;
    const int prob = t / D, dim = t - prob * D;
    const Philox rng{seed_lo, seed_hi};
    unsigned char perm[DE_MAX_S];
    for (int k = 0; k < S; ++k) perm[k] = (unsigned char)k;
    for (int k = S - 1; k > 0; --k) {
        const uint4 r = rng(0u, (unsigned)prob, (unsigned)(dim * DE_MAX_S + k), STREAM_INIT);
        const int j = (int)(u01(r.x, r.y) * (k + 1));
        const unsigned char tmp = perm[k];
        perm[k] = perm[j];
        perm[j] = tmp;
    }
    const double seg = 1.0 / S;
    for (int k = 0; k < S; ++k) {
        const uint4 r = rng(1u, (unsigned)prob, (unsigned)(dim * DE_MAX_S + k), STREAM_INIT);
        pop[((size_t)prob * S + k) * D + dim] = seg * u01(r.x, r.y) + seg * perm[k];
    }
}

// One generation of trial vectors, best1bin.  block = active problem, thread = member.
//   trial_x [na][S][D]   scaled trial vectors
//   trial_p [na][S][DP]  the same in parameter units (lo + x * span), DP >= D, columns D.. are 0
__global__ void __launch_bounds__(DE_MAX_S)
    de_trial_kernel(const double *__restrict__ pop, const double *__restrict__ E, const int *__restrict__ active,
                    int S, int D, int DP, const double *__restrict__ lo, const double *__restrict__ span,
                    unsigned seed_lo, unsigned seed_hi, unsigned gen, double mut_lo, double mut_hi, double cr,
                    double *__restrict__ trial_x, double *__restrict__ trial_p) {
    __shared__ double s_e[DE_MAX_S];
    __shared__ int s_i[DE_MAX_S];
    const int a = blockIdx.x, prob = active[a], i = threadIdx.x;
    const Philox rng{seed_lo, seed_hi};
    // argmin of the energies, first index on ties (numpy.argmin); NaN counts as +inf
    double e = CUDART_INF;
    if (i < S) {
        e = E[(size_t)prob * S + i];
        if (isnan(e)) e = CUDART_INF;
    }
    s_e[i] = e;
    s_i[i] = i;
    __syncthreads();
    for (int off = DE_MAX_S / 2; off > 0; off >>= 1) {
        if (i < off) {
            const double eo = s_e[i + off];
            const int io = s_i[i + off];
            if (eo < s_e[i] || (eo == s_e[i] && io < s_i[i])) {
                s_e[i] = eo;
                s_i[i] = io;
            }
        }
        __syncthreads();
    }
    const int best = s_i[0];
    if (i >= S) return;
    const uint4 rd = rng(gen, (unsigned)prob, 0u, STREAM_DITHER);
    const double scale = mut_lo + (mut_hi - mut_lo) * u01(rd.x, rd.y);           // dither: one draw per generation
    // two members, different from i and from each other
    const uint4 rp = rng(gen, (unsigned)prob, (unsigned)i, STREAM_PICK);
    const int r0 = (i + 1 + (int)(u01(rp.x, rp.y) * (S - 1))) % S;
    int r1 = (i + 1 + (int)(u01(rp.z, rp.w) * (S - 2))) % S;
    if (r1 == r0) r1 = (r1 + 1) % S;
    if (r1 == i) r1 = (r1 + 1) % S;
    if (r1 == r0) r1 = (r1 + 1) % S;
    const uint4 rc0 = rng(gen, (unsigned)prob, (unsigned)i, STREAM_CROSS);
    const uint4 rc1 = rng(gen, (unsigned)prob, (unsigned)(i + DE_MAX_S), STREAM_CROSS);
    const uint4 rc2 = rng(gen, (unsigned)prob, (unsigned)(i + 2 * DE_MAX_S), STREAM_CROSS);
    const unsigned cr_bits[12] = {rc0.x, rc0.y, rc0.z, rc0.w, rc1.x, rc1.y, rc1.z, rc1.w, rc2.x, rc2.y, rc2.z, rc2.w};
    const int forced = (int)(u01(cr_bits[10], cr_bits[11]) * D);               // one gene always comes from the mutant
    const double *P = pop + (size_t)prob * S * D;
    for (int j = 0; j < D; ++j) {
        const double own = P[(size_t)i * D + j];
        const double mutant = P[(size_t)best * D + j] + scale * (P[(size_t)r0 * D + j] - P[(size_t)r1 * D + j]);
        const bool cross = (j == forced) || ((double)cr_bits[j] * (1.0 / 4294967296.0) < cr);
        double x = cross ? mutant : own;
        if (x < 0.0 || x > 1.0) {                                                 // out of bounds: redrawn uniformly
            const uint4 ro = rng(gen, (unsigned)prob, (unsigned)(i * DE_MAX_D + j), STREAM_OOB);
            x = u01(ro.x, ro.y);
        }
        trial_x[((size_t)a * S + i) * D + j] = x;
        trial_p[((size_t)a * S + i) * DP + j] = lo[j] + x * span[j];
    }
    for (int j = D; j < DP; ++j) trial_p[((size_t)a * S + i) * DP + j] = 0.0;
}

// Deferred selection + convergence test.  block = active problem, thread = member.
//   conv[prob] = generation at which std(E) <= atol + tol |mean(E)| first held with all energies finite (0 = not yet);
//   a converged problem is frozen (scipy stops it there).  trial_x == nullptr: only the test (after the initial
//   evaluation, when E itself was just written).
__global__ void __launch_bounds__(DE_MAX_S)
    de_select_kernel(double *__restrict__ pop, double *__restrict__ E, const int *__restrict__ active, int S, int D,
                     const double *__restrict__ trial_x, const double *__restrict__ trial_e, double tol, double atol,
                     int gen, int *__restrict__ conv) {
    __shared__ double s_sum[DE_MAX_S];
    __shared__ int s_bad;
    const int a = blockIdx.x, prob = active[a], i = threadIdx.x;
    if (i == 0) s_bad = 0;
    __syncthreads();
    const bool frozen = conv[prob] != 0;
    double e = 0.0;
    if (i < S) {
        e = E[(size_t)prob * S + i];
        if (isnan(e)) e = CUDART_INF;
        if (trial_x && !frozen) {
            double et = trial_e[(size_t)a * S + i];
            if (isnan(et)) et = CUDART_INF;
            if (et <= e) {
                e = et;
                for (int j = 0; j < D; ++j) pop[((size_t)prob * S + i) * D + j] = trial_x[((size_t)a * S + i) * D + j];
            }
        }
        E[(size_t)prob * S + i] = e;
        if (!isfinite(e)) atomicOr(&s_bad, 1);
    }
    // mean and population standard deviation of the energies (numpy.std, ddof = 0)
    s_sum[i] = (i < S && isfinite(e)) ? e : 0.0;
    __syncthreads();
    for (int off = DE_MAX_S / 2; off > 0; off >>= 1) {
        if (i < off) s_sum[i] += s_sum[i + off];
        __syncthreads();
    }
    const double mean = s_sum[0] / S;
    __syncthreads();
    s_sum[i] = (i < S && isfinite(e)) ? (e - mean) * (e - mean) : 0.0;
    __syncthreads();
    for (int off = DE_MAX_S / 2; off > 0; off >>= 1) {
        if (i < off) s_sum[i] += s_sum[i + off];
        __syncthreads();
    }
    if (i == 0 && !frozen && !s_bad) {
        const double sd = sqrt(s_sum[0] / S);
        if (sd <= atol + tol * fabs(mean)) conv[prob] = gen > 0 ? gen : -1;      // -1: converged at initialisation
    }
}

}  // namespace xan

using namespace xan;

extern "C" {

int xan_de_init(double *d_pop, int n_problems, int pop_size, int n_dims, unsigned long long seed, void *stream) {
    XAN_REQUIRE(d_pop && n_problems > 0 && pop_size >= 4 && pop_size <= DE_MAX_S && n_dims >= 1 && n_dims <= DE_MAX_D,
                "xan_de_init: bad arguments (population 4..%d, dimensions 1..%d)", DE_MAX_S, DE_MAX_D);
    de_init_kernel<<<ceil_div(n_problems * n_dims, 64), 64, 0, (cudaStream_t)stream>>>(
        d_pop, n_problems, pop_size, n_dims, (unsigned)(seed & 0xffffffffu), (unsigned)(seed >> 32));
    XAN_CUDA_CHECK(cudaGetLastError());
    return XAN_OK;
}

int xan_de_trial(const double *d_pop, const double *d_energy, const int *d_active, int n_active, int pop_size, int n_dims,
                 int n_par_cols, const double *d_lo, const double *d_span, unsigned long long seed, int generation,
                 double mutation_lo, double mutation_hi, double recombination, double *d_trial_x, double *d_trial_par,
                 void *stream) {
    XAN_REQUIRE(d_pop && d_energy && d_active && d_lo && d_span && d_trial_x && d_trial_par, "xan_de_trial: null pointer");
    XAN_REQUIRE(n_active > 0 && pop_size >= 4 && pop_size <= DE_MAX_S && n_dims >= 1 && n_dims <= DE_MAX_D &&
                    n_par_cols >= n_dims && generation >= 1,
                "xan_de_trial: bad arguments");
    de_trial_kernel<<<n_active, DE_MAX_S, 0, (cudaStream_t)stream>>>(
        d_pop, d_energy, d_active, pop_size, n_dims, n_par_cols, d_lo, d_span, (unsigned)(seed & 0xffffffffu),
        (unsigned)(seed >> 32), (unsigned)generation, mutation_lo, mutation_hi, recombination, d_trial_x, d_trial_par);
    XAN_CUDA_CHECK(cudaGetLastError());
    return XAN_OK;
}

int xan_de_select(double *d_pop, double *d_energy, const int *d_active, int n_active, int pop_size, int n_dims,
                  const double *d_trial_x, const double *d_trial_energy, double tol, double atol, int generation,
                  int *d_converged, void *stream) {
    XAN_REQUIRE(d_pop && d_energy && d_active && d_converged, "xan_de_select: null pointer");
    XAN_REQUIRE((d_trial_x == nullptr) == (d_trial_energy == nullptr), "xan_de_select: trial vectors and energies go together");
    XAN_REQUIRE(n_active > 0 && pop_size >= 4 && pop_size <= DE_MAX_S && n_dims >= 1 && n_dims <= DE_MAX_D && generation >= 0,
                "xan_de_select: bad arguments");
    de_select_kernel<<<n_active, DE_MAX_S, 0, (cudaStream_t)stream>>>(d_pop, d_energy, d_active, pop_size, n_dims, d_trial_x,
                                                                      d_trial_energy, tol, atol, generation, d_converged);
    XAN_CUDA_CHECK(cudaGetLastError());
    return XAN_OK;
}

}  // extern "C"
