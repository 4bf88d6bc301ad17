// Step-wise (legacy v1) path of Xanthos: Hargreaves PET (xanthos/pet/hargreaves.py) and GWAM runoff
// (xanthos/runoff/gwam.py), which the reference calls once per month from Components.simulation
// (components.py:329-366).  fp64, month-major fields [nmonths][ld], compiled with -fmad=false.
//
//   hargreaves_pet_kernel - thread = (cell, month); every month is independent, so the whole series is
//                           one launch; the two per-month scalars (solar declination, inverse relative
//                           Earth-Sun distance; utils/general.py:53-90) and the days of the month come
//                           from a small device table.
//   gwam_kernel           - thread = cell; the soil moisture carried from month to month (sm_prev,
//                           components.py:361-366) lives in a register for the optional spin-up pass over
//                           the first `spinup` months and for the simulation that follows it
//                           (configurations.py:106-123); forcing is prefetched 4 months ahead.
#include "common.cuh"

namespace xan {

// numpy.maximum / numpy.minimum propagate NaN
__device__ __forceinline__ double npmax(double a, double b) { return (isnan(a) || isnan(b)) ? (a + b) : (a >= b ? a : b); }
__device__ __forceinline__ double npmin(double a, double b) { return (isnan(a) || isnan(b)) ? (a + b) : (a <= b ? a : b); }

__global__ void __launch_bounds__(256)
    hargreaves_pet_kernel(const double *__restrict__ temp, const double *__restrict__ dtr,
                          const double *__restrict__ lat_rad, const double *__restrict__ month_tab /* [M][3] */,
                          double *__restrict__ pet, int ncell, int ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (c >= ncell) return;
    const double Y = month_tab[3 * m], dr = month_tab[3 * m + 1], days = month_tab[3 * m + 2];
    const double X = lat_rad[c];
    const size_t off = (size_t)m * ld + c;
    // prep_arrays / prep_pet: NaN -> 0 (components.py:143-176)
    const double t = nan_to_num(ldg_stream(temp + off));
    double d = nan_to_num(ldg_stream(dtr + off));
    // sunset hour angle with the reference's clipping acos (hargreaves.py:48-73); NaN -> 0
    const double a = (-tan(X)) * tan(Y);
    double ws = 0.0;
    if (a <= 1.0 && a >= -1.0) ws = acos(a);
    else if (a < -1.0) ws = acos(-1.0);
    else if (a > 1.0) ws = acos(1.0);
    const double ra = (15.392 * dr) * (((ws * sin(X)) * sin(Y)) + ((cos(X) * cos(Y)) * sin(ws)));   // :42-44
    if (d < 0) d = 0.;                                                                             // :32
    const double evap = (((days * 0.0023) * ra) * (t + 17.8)) * sqrt(d);                           // :35
    stg_stream(pet + off, npmax(evap, 0.0));                                                       // :36
}

// One month of runoffgen (gwam.py:18-88) for one cell; `ch` = soil moisture carried from last month.
__device__ __forceinline__ void gwam_step(double pet, double p, double sm, double ch, double one_minus_e,
                                          double &aet, double &q, double &sav) {
    aet = 0.0;
    q = 0.0;
    sav = 0.0;
    const double b = (ch + p) - pet;                                   // :42
    if (sm == 999.0) {                                                 // :52 water bodies
        q = npmax(0.0, p - pet);                                       // :61
        if (isnan(q)) q = 0.0;                                         // :62
        aet = npmin(p, pet);                                           // :63
        if (isnan(aet)) aet = pet;                                     // :64
    } else if (sm != 0.0) {
        if (b >= sm) {                                                 // :67-69 (false for NaN)
            q = b - sm;
            sav = sm;
            aet = pet;
        } else if (b < sm) {                                           // :72-86 (false for NaN)
            const double r = ch / sm;
            const double t3 = ch + p;
            const double t5 = (((5. * ch) / sm) - (2. * (r * r))) / 3.;
            const double t6 = npmin(1.0, t5);
            const double t7 = pet * npmax(0.1, t6);
            aet = npmin(t3, t7);
            // alpha = 1 (gwam.py:71): chstor * (1 - exp(-chstor / Sm)) / (1 - exp(-1)) + (P - AET)
            const double t8 = ((ch * (1 - exp((-ch) / sm))) / one_minus_e) + (p - aet);
            sav = npmin(sm, t8);
            if (sav <= 0) {                                            // :83-85
                sav = 0.0;
                aet = p + ch;
            }
            q = npmax(0.0, ((ch + p) - aet) - sav);                    // :86
        }
    }
}

constexpr int GWAM_U = 4;

__global__ void __launch_bounds__(128)
    gwam_kernel(const double *__restrict__ pet, const double *__restrict__ precip, const double *__restrict__ sm_max,
                const double *__restrict__ sm_prev, int ncell, int nmonths, int spinup, int ld, double one_minus_e,
                double *__restrict__ aet_o, double *__restrict__ q_o, double *__restrict__ sav_o,
                double *__restrict__ sm_spun_o, double *__restrict__ sm_last_o) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const double sm = sm_max[c];
    double ch = sm_prev[c];
    for (int pass = 0; pass < 2; ++pass) {
        const int n = pass == 0 ? spinup : nmonths;
        const bool store = pass == 1;
        int i = 0;
        for (; i + GWAM_U <= n; i += GWAM_U) {
            double e[GWAM_U], p[GWAM_U];
#pragma unroll
            for (int u = 0; u < GWAM_U; ++u) {
                const size_t off = (size_t)(i + u) * ld + c;
                e[u] = ldg_stream(pet + off);
                p[u] = ldg_stream(precip + off);
            }
#pragma unroll
            for (int u = 0; u < GWAM_U; ++u) {
                double aet, q, sav;
                gwam_step(e[u], p[u], sm, ch, one_minus_e, aet, q, sav);
                ch = sav;                                              // components.py:366
                if (store) {
                    const size_t off = (size_t)(i + u) * ld + c;
                    if (aet_o) stg_stream(aet_o + off, aet);
                    if (q_o) stg_stream(q_o + off, q);
                    if (sav_o) stg_stream(sav_o + off, sav);
                }
            }
        }
        for (; i < n; ++i) {
            const size_t off = (size_t)i * ld + c;
            double aet, q, sav;
            gwam_step(pet[off], precip[off], sm, ch, one_minus_e, aet, q, sav);
            ch = sav;
            if (store) {
                if (aet_o) stg_stream(aet_o + off, aet);
                if (q_o) stg_stream(q_o + off, q);
                if (sav_o) stg_stream(sav_o + off, sav);
            }
        }
        if (pass == 0 && sm_spun_o) sm_spun_o[c] = ch;
    }
    if (sm_last_o) sm_last_o[c] = ch;
}

}  // namespace xan

using namespace xan;

extern "C" {

int xan_hargreaves_pet(const double *d_temp, const double *d_dtr, const double *d_lat_rad, const double *h_solar_dec,
                       const double *h_dr, const int *h_days, double *d_pet, int ncell, int nmonths, int ld,
                       void *stream) {
    XAN_REQUIRE(d_temp && d_dtr && d_lat_rad && h_solar_dec && h_dr && h_days && d_pet, "xan_hargreaves_pet: null pointer");
    XAN_REQUIRE(ncell > 0 && nmonths > 0 && nmonths <= 65535 && ld >= ncell,
                "xan_hargreaves_pet: bad shape ncell=%d nmonths=%d ld=%d", ncell, nmonths, ld);
    cudaStream_t s = (cudaStream_t)stream;
    double *tab = nullptr, *h_tab = nullptr;
    XAN_CUDA_CHECK(cudaMallocHost(&h_tab, sizeof(double) * 3 * (size_t)nmonths));
    for (int m = 0; m < nmonths; ++m) {
        h_tab[3 * m] = h_solar_dec[m];
        h_tab[3 * m + 1] = h_dr[m];
        h_tab[3 * m + 2] = (double)h_days[m];
    }
    XAN_CUDA_CHECK(scratch_alloc(&tab, sizeof(double) * 3 * (size_t)nmonths, s));
    XAN_CUDA_CHECK(cudaMemcpyAsync(tab, h_tab, sizeof(double) * 3 * (size_t)nmonths, cudaMemcpyHostToDevice, s));
    hargreaves_pet_kernel<<<dim3(ceil_div(ncell, 256), nmonths), 256, 0, s>>>(d_temp, d_dtr, d_lat_rad, tab, d_pet,
                                                                             ncell, ld);
    XAN_CUDA_CHECK(cudaGetLastError());
    XAN_CUDA_CHECK(cudaFreeAsync(tab, s));
    XAN_CUDA_CHECK(cudaStreamSynchronize(s));   // the staging table is read by the copy above
    XAN_CUDA_CHECK(cudaFreeHost(h_tab));
    return XAN_OK;
}

int xan_gwam_run(const double *d_pet, const double *d_precip, const double *d_sm_max, const double *d_sm_prev,
                 int ncell, int nmonths, int spinup_months, int ld, double *d_aet, double *d_q, double *d_sav,
                 double *d_sm_after_spinup, double *d_sm_last, void *stream) {
    XAN_REQUIRE(d_pet && d_precip && d_sm_max && d_sm_prev, "xan_gwam_run: null pointer");
    XAN_REQUIRE(ncell > 0 && nmonths > 0 && spinup_months >= 0 && spinup_months <= nmonths && ld >= ncell,
                "xan_gwam_run: bad arguments ncell=%d nmonths=%d spinup=%d ld=%d", ncell, nmonths, spinup_months, ld);
    const double one_minus_e = 1 - exp(-1.0);   // (1 - np.exp(-alpha)), alpha = 1; glibc and numpy agree on exp(-1)
    gwam_kernel<<<ceil_div(ncell, 128), 128, 0, (cudaStream_t)stream>>>(d_pet, d_precip, d_sm_max, d_sm_prev, ncell,
                                                                        nmonths, spinup_months, ld, one_minus_e, d_aet,
                                                                        d_q, d_sav, d_sm_after_spinup, d_sm_last);
    XAN_CUDA_CHECK(cudaGetLastError());
    return XAN_OK;
}

}  // extern "C"
