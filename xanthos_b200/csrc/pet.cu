// PET kernels: Hargreaves-Samani, Thornthwaite, Penman-Monteith.  fp64, month-major fields.
// Compiled with -fmad=false: the operation order of the reference formulas is kept so that
// results differ from numpy only through the <= 2 ulp transcendental functions.
#include "common.cuh"
#include "pm_common.cuh"

#include <vector>
#include <cstring>

namespace xan {

// =============================================================================================
// Hargreaves-Samani (xanthos/pet/hargreaves_samani.py:31-65, 91-119)
// =============================================================================================
// ra depends on (latitude, month of year) only: table [12][ncell].
__global__ void hs_ra_table_kernel(const double *__restrict__ lat_deg, double *__restrict__ ra_tab,
                                   int ncell) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int moy = blockIdx.y;
    if (c >= ncell) return;
    const double pi = 3.141592653589793;
    const double dy = 15.0 + 30.0 * moy;                                  // :36
    const double delta = 0.4102 * sin(2 * (pi / 365) * (dy - 80));        // :47
    const double phi = (lat_deg[c] * pi / 180);                           // :49
    const double tn = -tan(delta) * tan(phi);                             // :51
    double acs;
    if ((tn < -1.) || (tn > 1.)) acs = 0;                                 // :53-56
    else acs = acos(tn);
    ra_tab[(size_t)moy * ncell + c] = 118 / pi * acs + cos(phi) * cos(delta) * sin(acs);   // :59
}

// thread = (cell, year): 36 independent streaming loads in flight per thread.
__global__ void __launch_bounds__(256)
    hs_pet_kernel(const double *__restrict__ tas, const double *__restrict__ tmax,
                  const double *__restrict__ tmin, const double *__restrict__ ra_tab,
                  double *__restrict__ pet, int ncell, int ld, int start_year) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (c >= ncell) return;
    const bool leap = is_leap_gregorian(start_year + y);
    double t[12], hi[12], lo[12], ra[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const size_t off = (size_t)(y * 12 + k) * ld + c;
        t[k] = ldg_stream(tas + off);
        hi[k] = ldg_stream(tmax + off);
        lo[k] = ldg_stream(tmin + off);
        ra[k] = __ldg(ra_tab + (size_t)k * ncell + c);
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        double v = 0.408 * 0.0023 * ra[k] * (t[k] + 17.8) * sqrt(fabs(hi[k] - lo[k]));   // :62
        if (t[k] < 0) v = 0;                                                             // :33
        v = v * (double)month_days(k, leap);                                             // :114
        stg_stream(pet + (size_t)(y * 12 + k) * ld + c, v);
    }
}

// =============================================================================================
// Thornthwaite (xanthos/pet/thornthwaite.py:18-44, 47-130)
// =============================================================================================
// Monthly mean day length per cell for a 365-day and a 366-day year: table [24][ncell].
__global__ void __launch_bounds__(128)
    tw_daylight_kernel(const double *__restrict__ lat_rad, double *__restrict__ L_tab, int ncell) {
    __shared__ double tan_dec[366];
    const double pi = 3.141592653589793;
    for (int d = threadIdx.x; d < 366; d += blockDim.x) {
        const double dec = 0.409 * sin(((2 * pi / 365.0) * (double)(d + 1) - 1.39));   // :30
        tan_dec[d] = tan(dec);
    }
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const double mtl = -tan(lat_rad[c]);                                               // :34
    for (int leap = 0; leap < 2; ++leap) {
        int first = 0;
        for (int moy = 0; moy < 12; ++moy) {
            const int n = month_days(moy, leap != 0);
            auto hours = [&](int i) {
                double x = mtl * tan_dec[first + i];
                x = fmin(fmax(x, -1.0), 1.0);                                          // :35 (clip)
                return acos(x) * (24.0 / pi);                                          // :38
            };
            // np.add.reduceat (:41): first element + pairwise sum of the remaining n-1
            const double h0 = hours(0);
            const double rest = numpy_pairwise_sum(n - 1, [&](int i) { return hours(i + 1); });
            L_tab[(size_t)(leap * 12 + moy) * ncell + c] = (h0 + rest) / (double)n;
            first += n;
        }
    }
}

// thread = (cell, year)
__global__ void __launch_bounds__(256)
    tw_pet_kernel(const double *__restrict__ tas, const double *__restrict__ L_tab,
                  double *__restrict__ pet, int ncell, int ld, int start_year, int nyears) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (c >= ncell) return;
    const bool leap = is_leap_gregorian(start_year + y);
    double t[12], hi[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        double v = ldg_stream(tas + (size_t)(y * 12 + k) * ld + c);
        if (isnan(v) || v < 0) v = 0;                                                  // :82
        t[k] = v;
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) hi[k] = pow(t[k] / 5.0, 1.514);                       // :88
    // np.add.reduceat over the 12 months of the year (:91): i0 + pairwise(i1..i11)
    double I = ((hi[1] + hi[2]) + (hi[3] + hi[4])) + ((hi[5] + hi[6]) + (hi[7] + hi[8]));
    I = I + hi[9];
    I = I + hi[10];
    I = I + hi[11];
    I = hi[0] + I;
    const double a = (.000000675 * pow(I, 3.0)) - (.0000771 * (I * I)) + (.0179 * I) + .492;   // :94
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const int kabs = y * 12 + k;
        double ratio = 0.0;
        if (I != 0) ratio = (10 * t[k]) / I;                                           // :104
        const double pu = 16 * pow(ratio, a);                                          // :105
        // day length: leap years use the 366-day table by calendar month; other years inherit the
        // np.repeat tiling of :110 (column kabs -> month kabs / nyears of the 365-day table)
        const double L = leap ? __ldg(L_tab + (size_t)(12 + k) * ncell + c)
                              : __ldg(L_tab + (size_t)(kabs / nyears) * ncell + c);
        const double N = (double)month_days(k, leap);
        stg_stream(pet + (size_t)kabs * ld + c, pu * (L / 12) * (N / 30.0));           // :127
    }
}

// =============================================================================================
// Penman-Monteith (xanthos/pet/penman_monteith.py)
// =============================================================================================
// per-(class, month) quantities shared by every cell, built once per block in shared memory
struct PmShared {
    double fc[XAN_PM_MAX_CLASSES][12];      // vegetation cover fraction (:257-261)
    double oma[XAN_PM_MAX_CLASSES][12];     // 1 - alpha
    double lai[XAN_PM_MAX_CLASSES][12];
    double cL[XAN_PM_MAX_CLASSES], beta[XAN_PM_MAX_CLASSES], rslimit[XAN_PM_MAX_CLASSES],
        topen[XAN_PM_MAX_CLASSES], tclose[XAN_PM_MAX_CLASSES], vclose[XAN_PM_MAX_CLASSES],
        vopen[XAN_PM_MAX_CLASSES], rblmin[XAN_PM_MAX_CLASSES], rblmax[XAN_PM_MAX_CLASSES],
        rc[XAN_PM_MAX_CLASSES], inv_rc[XAN_PM_MAX_CLASSES], emiss[XAN_PM_MAX_CLASSES];
};

// thread = (cell, year); months outer, land classes inner.
__global__ void __launch_bounds__(128)
    pm_pet_kernel(const double *__restrict__ tair, const double *__restrict__ tmin_,
                  const double *__restrict__ rhs, const double *__restrict__ wind,
                  const double *__restrict__ rsds, const double *__restrict__ rlds,
                  const double *__restrict__ lct, const double *__restrict__ elev,
                  const int *__restrict__ prev_idx, const PmTab *__restrict__ tab,
                  double *__restrict__ pet, int ncell, int ld, int start_year) {
    __shared__ PmShared sh;
    const int nlcs = tab->nlcs;
    for (int i = threadIdx.x; i < nlcs * 12; i += blockDim.x) {
        const int l = i / 12, k = i % 12;
        const double emin = exp(-0.5 * tab->laimin[l][k]);
        double fcd = emin - exp(-0.5 * tab->laimax[l][k]);                // :257
        if (fcd == 0.0) fcd = 1;                                          // :258
        double fc = (emin - exp(-0.5 * tab->lai[l][k])) / fcd;            // :260
        if (fc > 1) fc = 1;                                               // :261
        sh.fc[l][k] = fc;
        sh.oma[l][k] = 1 - tab->alpha[l][k];
        sh.lai[l][k] = tab->lai[l][k];
    }
    for (int l = threadIdx.x; l < nlcs; l += blockDim.x) {
        sh.cL[l] = tab->cL[l];
        sh.beta[l] = tab->beta[l];
        sh.rslimit[l] = tab->rslimit[l];
        sh.topen[l] = tab->Tminopen[l];
        sh.tclose[l] = tab->Tminclose[l];
        sh.vclose[l] = tab->VPDclose[l];
        sh.vopen[l] = tab->VPDopen[l];
        sh.rblmin[l] = tab->RBLmin[l];
        sh.rblmax[l] = tab->RBLmax[l];
        sh.rc[l] = tab->rc[l];
        sh.inv_rc[l] = 1 / tab->rc[l];
        sh.emiss[l] = tab->emiss[l];
    }
    __syncthreads();

    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (c >= ncell) return;
    const int water_idx = tab->water_idx, snow_idx = tab->snow_idx;
    const bool leap = is_leap_gregorian(start_year + y);                  // :57
    const double *lct_y = lct + (size_t)tab->lc_index[y] * nlcs * ld + c; // class l at lct_y[l * ld]

    // land-cover total in numpy's pairwise order (:45-47)
    double totpct = numpy_pairwise_sum(nlcs, [&](int l) { return __ldg(lct_y + (size_t)l * ld); });
    if (totpct == 0) totpct = 0.01;

    const double p = 101325 * pow((1 - 0.0065 * elev[c] / 288.15), 5.2558);   // :187
    const double wind_k = pow(2.0 / 10, 0.11);                                // :99
    int pc = (prev_idx != nullptr) ? prev_idx[c] : c - 1;                     // data_load.py:128-129

    for (int k = 0; k < 12; ++k) {
        const size_t off = (size_t)(y * 12 + k) * ld;
        const double T = ldg_stream(tair + off + c);
        const double Tn = ldg_stream(tmin_ + off + c);
        const double RH = ldg_stream(rhs + off + c);
        const double W = ldg_stream(wind + off + c);
        const double Rs = ldg_stream(rsds + off + c);
        const double Rl = ldg_stream(rlds + off + c);
        const double Tp = (pc >= 0) ? __ldg(tair + off + pc) : 0.0;
        const double dz = (double)month_days(k, leap);
        const double dzs = 86400 * dz;

        // ---- class independent ------------------------------------------------------------
        const double esx = 6.10588 * exp(17.32491 * T / (T + 238.102));           // :83
        const double vap = esx * (RH / 100);                                      // :86
        const double tk = T + 238.1;
        const double sx = (238.1 * 17.325 * esx / (tk * tk));                     // :89
        const double rcorr = p / (101300 * pow((273.15 + T) / 293.15, 1.75));     // :229
        const double gcu = 0.00001 * rcorr;                                       // :230
        const double vpd = esx - vap;                                             // :121
        const double rh = (RH > 99.9999) ? 99.9 : RH;                             // :205-209
        double g = 1.6198 * (T - Tp);                                             // :212
        if (k == 0) g = 0;                                                        // :214
        const double rho = p / ((T + 273.15) * 287.058);                          // :271
        const double rr = rho * PM_CP / (4.0 * PM_SIGMA2 * pow((T + 273.15), 3.0));   // :273
        const double rh100 = rh / 100;
        double fwet = (rh < 70) ? 0 : rh;                                         // :165-172
        if (rh >= 70) fwet = pow(rh100, 8.0);
        if (rh >= 80) fwet = pow(rh100, 10.0);
        if (rh >= 90) fwet = pow(rh100, 12.0);
        if (rh >= 95) fwet = pow(rh100, 16.0);
        const double T4 = pow(T + 273, 4.0);                                      // :158
        const double rho_cp = rho * PM_CP;
        const double omf = 1 - fwet;

        // ---- open water, alpha row 0, emissivity 0.98 (:337-361) ---------------------------
        double wat;
        {
            const double oma0 = sh.oma[0][k];
            const double rnlx = PM_SIGMA * T4 * 0.98 * dz - Rl * 86400 * dz;
            double rnx = (oma0 * Rs) * 86400 * dz - rnlx;
            if (rnx < 0) rnx = 0.0;
            const double rsnx = oma0 * Rs * 86400 * dz;                           // :92
            const double qtx = 0.5 * rsnx - ((k <= 5) ? 0.8 : 1.3) * rnlx;        // :347-349
            double ax = (rnx - qtx) / dzs;
            if (ax < 0) ax = 0;
            const double rn2x = rnx / dzs;
            const double ewetx = rn2x * dz * 0.6 / 2845;
            const double wind2 = W * wind_k;
            const double ewety = dz * 86400 * (sx * ax + PM_GAMMA * 6.43 * (0.5 + 0.54 * wind2) * (esx - vap)) /
                                 ((sx + PM_GAMMA) * PM_LAMBDA1);
            wat = (T < -1) ? ewetx : ewety;
            if (wat < 0.0) wat = 0.0;
        }
        // ---- snow, alpha row 6, emissivity 0.85 (:364-377) ---------------------------------
        double snow;
        {
            const double rnlx = PM_SIGMA * T4 * 0.85 * dz - Rl * 86400 * dz;
            double rnx = (sh.oma[6][k] * Rs) * 86400 * dz - rnlx;
            if (rnx < 0) rnx = 0.0;
            snow = (rnx / dzs) * dz * 0.6 / 2845;
            if (snow < 0.0) snow = 0.0;
        }

        double acc = 0.0;
        for (int l = 0; l < nlcs; ++l) {
            double eet;
            if (l == snow_idx) {
                eet = snow;                                                       // :462-464
            } else if (l == water_idx) {
                eet = wat;                                                        // :459-460
            } else {
                const double topen = sh.topen[l], tclose = sh.tclose[l];
                const double vopen = sh.vopen[l], vclose = sh.vclose[l];
                const double rc = sh.rc[l], rslimit = sh.rslimit[l];
                const double LAI = sh.lai[l][k], fc = sh.fc[l][k];

                double mtmin = 0.0;                                               // :102-114
                if (Tn <= tclose) mtmin = 0.1;
                else if (Tn >= topen) mtmin = 1.0;
                else if (Tn < topen && Tn > tclose) mtmin = (Tn - tclose) / (topen - tclose);

                const bool between = (vpd > vopen) && (vpd < vclose);
                double mvpd = vpd, rtotc = 0.0;                                   // :117-145
                if (vpd >= vclose) {
                    mvpd = 0.1;
                    rtotc = sh.rblmin[l];
                } else if (vpd <= vopen) {
                    mvpd = 1.0;
                    rtotc = sh.rblmax[l];
                } else if (between) {
                    mvpd = (vclose - vpd) / (vclose - vopen);
                    rtotc = sh.rblmax[l] - (sh.rblmax[l] - sh.rblmin[l]) * (vclose - vpd) / (vclose - vopen);
                }
                const double gs1 = sh.cL[l] * mtmin * mvpd * rcorr;               // :242

                const double rnl = PM_SIGMA * T4 * sh.emiss[l] * dz - Rl * 86400 * dz;   // :158
                const double rn = (sh.oma[l][k] * Rs) * 86400 * dz - rnl;                // :159
                const double a = rn / dzs;                                               // :160
                const double ac = fc * a;                                                // :263
                const double asoil = (1 - fc) * a - g;                                   // :266

                double rtot = rtotc * rcorr;                                             // :268-269
                if (rtot > 80) rtot = 80;
                double ra = rc * rr / (rc + rr);                                         // :277-278
                if (ra > rtot) ra = rtot;

                const double den = gs1 + sh.inv_rc[l] + gcu;                             // :192-197
                double cc;
                if (den < 0.0001) cc = 10000;
                else if (fwet == 1) cc = 0.00001;
                else if (LAI < 0.0001) cc = 0.00001;
                else cc = 0;
                if (cc == 0) cc = sh.inv_rc[l] * (gs1 + gcu) * LAI * omf / den;
                double rs = (cc == 0) ? 100000 : 1 / cc;                                 // :285
                if (rs > rslimit) rs = rslimit;                                          // :291

                double lai_fwet = LAI * fwet;                                            // :296
                if (lai_fwet == 0) lai_fwet = 1;
                double rhc = (LAI > 0.00001) ? rc / lai_fwet : rslimit;                  // :297
                if (rhc > rslimit) rhc = rslimit;                                        // :300
                double rhrc = rhc * rr / (rhc + rr);                                     // :303-304
                if (rhrc > rtot) rhrc = rtot;

                const double apres = dz * 86400 * (sx * ac + rho_cp * vpd * fc / rhrc) * fwet /
                                     ((sx + p * 0.01 * PM_CP * rhc / (PM_LAMBDA1 * 0.622 * rhrc)) * PM_LAMBDA1);
                const double ewet_c = (rh >= 70) ? apres : 0.0;                          // :309-310
                const double rasoil = rtot * rr / (rtot + rr);                           // :312
                const double nsoil = 86400 * dz * (sx * asoil + rho_cp * (1 - fc) * vpd / rasoil);
                const double dsoil = (sx + PM_GAMMA * rtot / rasoil) * PM_LAMBDA1;
                const double ewet_soil = nsoil * fwet / dsoil;                           // :314-315
                const double esoilpot = nsoil * omf / dsoil;                             // :316-317
                const double esoil = ewet_soil + esoilpot * pow(rh100, vpd / sh.beta[l]);   // :323
                double trans = dz * 86400 * (sx * ac + rho_cp * vpd * fc / ra) * omf /
                               ((sx + PM_GAMMA * (1 + rs / ra)) * PM_LAMBDA1);           // :326-327
                if (fc == 0) trans = 0;                                                  // :328
                eet = trans + ewet_c + esoil;                                            // :330
                if (eet < 0.0) eet = 0.0;                                                // :332
            }
            acc = acc + eet * __ldg(lct_y + (size_t)l * ld);                             // :467-470
        }
        stg_stream(pet + off + c, acc / totpct);                                         // :470
    }
}

}  // namespace xan

using namespace xan;

extern "C" {

int xan_hs_pet(const double *d_tas, const double *d_tmax, const double *d_tmin, const double *d_lat_deg,
               double *d_pet, int ncell, int nmonths, int ld, int start_year, void *stream) {
    XAN_REQUIRE(d_tas && d_tmax && d_tmin && d_lat_deg && d_pet, "xan_hs_pet: null pointer");
    XAN_REQUIRE(ncell > 0 && nmonths > 0 && nmonths % 12 == 0 && ld >= ncell,
                "xan_hs_pet: bad shape ncell=%d nmonths=%d ld=%d (nmonths must be whole years)", ncell,
                nmonths, ld);
    cudaStream_t s = (cudaStream_t)stream;
    double *ra_tab = nullptr;
    XAN_CUDA_CHECK(scratch_alloc(&ra_tab, sizeof(double) * 12 * (size_t)ncell, s));
    hs_ra_table_kernel<<<dim3(ceil_div(ncell, 128), 12), 128, 0, s>>>(d_lat_deg, ra_tab, ncell);
    hs_pet_kernel<<<dim3(ceil_div(ncell, 256), nmonths / 12), 256, 0, s>>>(d_tas, d_tmax, d_tmin, ra_tab,
                                                                          d_pet, ncell, ld, start_year);
    XAN_CUDA_CHECK(cudaGetLastError());
    XAN_CUDA_CHECK(cudaFreeAsync(ra_tab, s));
    return XAN_OK;
}

int xan_thornthwaite_pet(const double *d_tas, const double *d_lat_rad, double *d_pet, int ncell,
                         int nmonths, int ld, int start_year, void *stream) {
    XAN_REQUIRE(d_tas && d_lat_rad && d_pet, "xan_thornthwaite_pet: null pointer");
    XAN_REQUIRE(ncell > 0 && nmonths > 0 && nmonths % 12 == 0 && ld >= ncell,
                "xan_thornthwaite_pet: bad shape ncell=%d nmonths=%d ld=%d", ncell, nmonths, ld);
    cudaStream_t s = (cudaStream_t)stream;
    double *L_tab = nullptr;
    XAN_CUDA_CHECK(scratch_alloc(&L_tab, sizeof(double) * 24 * (size_t)ncell, s));
    tw_daylight_kernel<<<ceil_div(ncell, 128), 128, 0, s>>>(d_lat_rad, L_tab, ncell);
    const int nyears = nmonths / 12;
    tw_pet_kernel<<<dim3(ceil_div(ncell, 256), nyears), 256, 0, s>>>(d_tas, L_tab, d_pet, ncell, ld,
                                                                    start_year, nyears);
    XAN_CUDA_CHECK(cudaGetLastError());
    XAN_CUDA_CHECK(cudaFreeAsync(L_tab, s));
    return XAN_OK;
}

int xan_thornthwaite_daylight(const double *d_lat_rad, double *d_hours, int ncell, void *stream) {
    XAN_REQUIRE(d_lat_rad && d_hours && ncell > 0, "xan_thornthwaite_daylight: bad arguments");
    tw_daylight_kernel<<<ceil_div(ncell, 128), 128, 0, (cudaStream_t)stream>>>(d_lat_rad, d_hours, ncell);
    XAN_CUDA_CHECK(cudaGetLastError());
    return XAN_OK;
}

int xan_pm_pet(const double *d_tair, const double *d_tmin, const double *d_rhs, const double *d_wind,
               const double *d_rsds, const double *d_rlds, const double *d_lct, const double *d_elev,
               const int *d_prev_idx, const xan_pm_tables *t, const int *h_lc_index, double *d_pet,
               int ncell, int nmonths, int ld, int start_year, void *stream) {
    XAN_REQUIRE(d_tair && d_tmin && d_rhs && d_wind && d_rsds && d_rlds && d_lct && d_elev && d_pet && t &&
                    h_lc_index,
                "xan_pm_pet: null pointer");
    XAN_REQUIRE(ncell > 0 && nmonths > 0 && nmonths % 12 == 0 && ld >= ncell,
                "xan_pm_pet: bad shape ncell=%d nmonths=%d ld=%d", ncell, nmonths, ld);
    XAN_REQUIRE(t->nlcs >= 7 && t->nlcs <= XAN_PM_MAX_CLASSES,
                "xan_pm_pet: nlcs=%d outside 7..%d (albedo rows 0 and 6 are hard-wired to water and snow)",
                t->nlcs, XAN_PM_MAX_CLASSES);
    XAN_REQUIRE(t->water_idx >= 0 && t->water_idx < t->nlcs && t->snow_idx >= 0 && t->snow_idx < t->nlcs,
                "xan_pm_pet: water/snow index out of range");
    const int nyears = nmonths / 12;
    XAN_REQUIRE(nyears <= 512, "xan_pm_pet: more than 512 years");
    cudaStream_t s = (cudaStream_t)stream;

    // Pack the small tables on the host.  The source is pageable memory, so cudaMemcpyAsync has
    // staged it by the time it returns and the buffer can be released right after the call.
    std::vector<unsigned char> h_buf(sizeof(PmTab), 0);
    PmTab *h_tab = reinterpret_cast<PmTab *>(h_buf.data()), *d_tab = nullptr;
    h_tab->nlcs = t->nlcs;
    h_tab->water_idx = t->water_idx;
    h_tab->snow_idx = t->snow_idx;
    for (int l = 0; l < t->nlcs; ++l) {
        h_tab->cL[l] = t->cL[l];
        h_tab->beta[l] = t->beta[l];
        h_tab->rslimit[l] = t->rslimit[l];
        h_tab->Tminopen[l] = t->Tminopen[l];
        h_tab->Tminclose[l] = t->Tminclose[l];
        h_tab->VPDclose[l] = t->VPDclose[l];
        h_tab->VPDopen[l] = t->VPDopen[l];
        h_tab->RBLmin[l] = t->RBLmin[l];
        h_tab->RBLmax[l] = t->RBLmax[l];
        h_tab->rc[l] = t->rc[l];
        h_tab->emiss[l] = t->emiss[l];
        for (int k = 0; k < 12; ++k) {
            h_tab->alpha[l][k] = t->alpha[l * 12 + k];
            h_tab->lai[l][k] = t->lai[l * 12 + k];
            h_tab->laimin[l][k] = t->laimin[l * 12 + k];
            h_tab->laimax[l][k] = t->laimax[l * 12 + k];
        }
    }
    for (int y = 0; y < nyears; ++y) {
        if (h_lc_index[y] < 0 || h_lc_index[y] > 255) {
            set_error("xan_pm_pet: land-cover index %d for year %d out of range", h_lc_index[y], y);
            return XAN_E_INVALID;
        }
        h_tab->lc_index[y] = (unsigned char)h_lc_index[y];
    }
    XAN_CUDA_CHECK(scratch_alloc(&d_tab, sizeof(PmTab), s));
    XAN_CUDA_CHECK(cudaMemcpyAsync(d_tab, h_tab, sizeof(PmTab), cudaMemcpyHostToDevice, s));
    const char *exact = getenv("XANTHOS_PM_EXACT");
    if (exact && exact[0] == '1')
        pm_pet_kernel<<<dim3(ceil_div(ncell, 128), nyears), 128, 0, s>>>(d_tair, d_tmin, d_rhs, d_wind, d_rsds,
                                                                        d_rlds, d_lct, d_elev, d_prev_idx, d_tab,
                                                                        d_pet, ncell, ld, start_year);
    else
        launch_pm_pet_fast(d_tair, d_tmin, d_rhs, d_wind, d_rsds, d_rlds, d_lct, d_elev, d_prev_idx, d_tab, d_pet,
                           ncell, nyears, ld, start_year, s);
    XAN_CUDA_CHECK(cudaGetLastError());
    XAN_CUDA_CHECK(cudaFreeAsync(d_tab, s));
    return XAN_OK;
}

}  // extern "C"
