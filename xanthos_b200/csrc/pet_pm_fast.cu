// Penman-Monteith PET, throughput version (xanthos/pet/penman_monteith.py:394-477).
//
// Same formulas as pm_pet_kernel in pet.cu (which keeps the reference's exact operation order and
// is the parity anchor), restructured for the fp64 pipe, which is what bounds this stage:
//   * quotients that share a denominator are evaluated as one reciprocal and several products;
//   * integer powers (x^2, x^3, x^4, rh^8/10/12/16) are products, not pow();
//   * pow(rh/100, vpd/beta) = exp((vpd/beta) * log(rh/100)) with log() hoisted out of the class loop;
//   * per-class constants (1/(VPDclose-VPDopen), 1/(Tminopen-Tminclose), 1/beta, ...) are tabulated once
//     per block.
// FMA contraction stays OFF (like everywhere else): the soil-evaporation terms cancel twice, and with
// contraction the result moved by up to 1.1e-9 relative in 1 of 24 M cell-months (measured, tools/pm_compare.py);
// without it the kernel stays within 5e-13 of the exact-order kernel over the full 67,420 x 360 workload
// (2e-13 of the numpy oracle on 67,420 x 24) and is 2.6x faster than it (3.4 ms vs 8.9 ms on a B200).
#include "pm_common.cuh"

namespace xan {

struct PmFastShared {
    double fc[XAN_PM_MAX_CLASSES][12];      // vegetation cover fraction (:257-261)
    double oma[XAN_PM_MAX_CLASSES][12];     // 1 - alpha
    double lai[XAN_PM_MAX_CLASSES][12];
    double cL[XAN_PM_MAX_CLASSES], inv_beta[XAN_PM_MAX_CLASSES], rslimit[XAN_PM_MAX_CLASSES],
        topen[XAN_PM_MAX_CLASSES], tclose[XAN_PM_MAX_CLASSES], inv_dt[XAN_PM_MAX_CLASSES],
        vclose[XAN_PM_MAX_CLASSES], vopen[XAN_PM_MAX_CLASSES], inv_dv[XAN_PM_MAX_CLASSES],
        rblmin[XAN_PM_MAX_CLASSES], rblmax[XAN_PM_MAX_CLASSES], rc[XAN_PM_MAX_CLASSES],
        inv_rc[XAN_PM_MAX_CLASSES], emiss[XAN_PM_MAX_CLASSES];
};

__device__ __forceinline__ double rcp(double x) { return 1.0 / x; }

__global__ void __launch_bounds__(128)
    pm_pet_fast_kernel(const double *__restrict__ tair, const double *__restrict__ tmin_,
                       const double *__restrict__ rhs, const double *__restrict__ wind,
                       const double *__restrict__ rsds, const double *__restrict__ rlds,
                       const double *__restrict__ lct, const double *__restrict__ elev,
                       const int *__restrict__ prev_idx, const PmTab *__restrict__ tab,
                       double *__restrict__ pet, int ncell, int ld, int start_year) {
    __shared__ PmFastShared sh;
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
        sh.inv_beta[l] = 1 / tab->beta[l];
        sh.rslimit[l] = tab->rslimit[l];
        sh.topen[l] = tab->Tminopen[l];
        sh.tclose[l] = tab->Tminclose[l];
        sh.inv_dt[l] = 1 / (tab->Tminopen[l] - tab->Tminclose[l]);
        sh.vclose[l] = tab->VPDclose[l];
        sh.vopen[l] = tab->VPDopen[l];
        sh.inv_dv[l] = 1 / (tab->VPDclose[l] - tab->VPDopen[l]);
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
    const double inv_tot = rcp(totpct);

    const double p = 101325 * pow((1 - 0.0065 * elev[c] / 288.15), 5.2558);   // :187
    const double wind_k = 0.8377478067708294;                                 // pow(2/10, 0.11), :99
    const int pc = (prev_idx != nullptr) ? prev_idx[c] : c - 1;               // data_load.py:128-129
    const double p_cp_001_over_l622 = p * 0.01 * PM_CP / (PM_LAMBDA1 * 0.622);

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
        const double inv_dzs = rcp(dzs);

        // ---- class independent ------------------------------------------------------------
        const double esx = 6.10588 * exp(17.32491 * T / (T + 238.102));           // :83
        const double vap = esx * (RH / 100);                                      // :86 (exact: vpd = esx - vap cancels near RH = 100)
        const double tk = T + 238.1;
        const double sx = (238.1 * 17.325) * esx / (tk * tk);                     // :89
        const double tK = T + 273.15;
        const double rcorr = p / (101300 * pow(tK * (1.0 / 293.15), 1.75));       // :229
        const double gcu = 0.00001 * rcorr;                                       // :230
        const double vpd = esx - vap;                                             // :121
        const double rh = (RH > 99.9999) ? 99.9 : RH;                             // :205-209
        const double g = (k == 0) ? 0.0 : 1.6198 * (T - Tp);                      // :212-214
        const double inv_tK = rcp(tK);
        const double rho = p * inv_tK * (1.0 / 287.058);                          // :271
        const double rho_cp = rho * PM_CP;
        const double rr = rho_cp * (inv_tK * inv_tK * inv_tK) * (1.0 / (4.0 * PM_SIGMA2));   // :273
        const double rh100 = rh / 100;
        double fwet = 0.0;                                                        // :165-172
        if (rh >= 70) {
            const double x2 = rh100 * rh100, x4 = x2 * x2, x8 = x4 * x4;
            fwet = (rh >= 95) ? x8 * x8 : (rh >= 90) ? x8 * x4 : (rh >= 80) ? x8 * x2 : x8;
        }
        const double t273 = T + 273, t2 = t273 * t273;
        const double T4 = t2 * t2;                                                // :158
        const double omf = 1 - fwet;
        const double log_rh = log(rh100);                                         // for pow(rh/100, vpd/beta), :323
        const double sig_t4_dz = PM_SIGMA * T4 * dz;
        const double rl_dzs = Rl * dzs;
        const double rs_dzs = Rs * dzs;
        const double rho_cp_vpd = rho_cp * vpd;

        // ---- open water, alpha row 0, emissivity 0.98 (:337-361) ---------------------------
        double wat;
        {
            const double oma0 = sh.oma[0][k];
            const double rnlx = sig_t4_dz * 0.98 - rl_dzs;
            const double rsnx = oma0 * rs_dzs;
            double rnx = rsnx - rnlx;
            if (rnx < 0) rnx = 0.0;
            const double qtx = 0.5 * rsnx - ((k <= 5) ? 0.8 : 1.3) * rnlx;        // :347-349
            double ax = (rnx - qtx) * inv_dzs;
            if (ax < 0) ax = 0;
            const double ewetx = rnx * inv_dzs * dz * (0.6 / 2845);
            const double wind2 = W * wind_k;
            const double ewety = dzs * (sx * ax + (PM_GAMMA * 6.43) * (0.5 + 0.54 * wind2) * vpd) /
                                 ((sx + PM_GAMMA) * PM_LAMBDA1);
            wat = (T < -1) ? ewetx : ewety;
            if (wat < 0.0) wat = 0.0;
        }
        // ---- snow, alpha row 6, emissivity 0.85 (:364-377) ---------------------------------
        double snow;
        {
            const double rnlx = sig_t4_dz * 0.85 - rl_dzs;
            double rnx = sh.oma[6][k] * rs_dzs - rnlx;
            if (rnx < 0) rnx = 0.0;
            snow = rnx * inv_dzs * dz * (0.6 / 2845);
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
                const double rc = sh.rc[l], inv_rc = sh.inv_rc[l], rslimit = sh.rslimit[l];
                const double LAI = sh.lai[l][k], fc = sh.fc[l][k];

                double mtmin = 0.0;                                               // :102-114
                if (Tn <= tclose) mtmin = 0.1;
                else if (Tn >= topen) mtmin = 1.0;
                else if (Tn < topen && Tn > tclose) mtmin = (Tn - tclose) * sh.inv_dt[l];

                double mvpd = vpd, rtotc = 0.0;                                   // :117-145
                if (vpd >= vclose) {
                    mvpd = 0.1;
                    rtotc = sh.rblmin[l];
                } else if (vpd <= vopen) {
                    mvpd = 1.0;
                    rtotc = sh.rblmax[l];
                } else if ((vpd > vopen) && (vpd < vclose)) {
                    mvpd = (vclose - vpd) * sh.inv_dv[l];
                    rtotc = sh.rblmax[l] - (sh.rblmax[l] - sh.rblmin[l]) * mvpd;
                }
                const double gs1 = sh.cL[l] * mtmin * mvpd * rcorr;               // :242

                const double rnl = sig_t4_dz * sh.emiss[l] - rl_dzs;              // :158
                const double a = (sh.oma[l][k] * rs_dzs - rnl) * inv_dzs;         // :159-160
                const double ac = fc * a;                                         // :263
                const double asoil = (1 - fc) * a - g;                            // :266

                double rtot = rtotc * rcorr;                                      // :268-269
                if (rtot > 80) rtot = 80;
                double ra = rc * rr / (rc + rr);                                  // :277-278
                if (ra > rtot) ra = rtot;
                const double inv_ra = rcp(ra);

                const double den = gs1 + inv_rc + gcu;                            // :192-197
                double cc;
                if (den < 0.0001) cc = 10000;
                else if (fwet == 1) cc = 0.00001;
                else if (LAI < 0.0001) cc = 0.00001;
                else cc = inv_rc * (gs1 + gcu) * LAI * omf / den;
                double rs = (cc == 0) ? 100000 : 1 / cc;                          // :285
                if (rs > rslimit) rs = rslimit;                                   // :291

                double lai_fwet = LAI * fwet;                                     // :296
                if (lai_fwet == 0) lai_fwet = 1;
                double rhc = (LAI > 0.00001) ? rc / lai_fwet : rslimit;           // :297
                if (rhc > rslimit) rhc = rslimit;                                 // :300
                double rhrc = rhc * rr / (rhc + rr);                              // :303-304
                if (rhrc > rtot) rhrc = rtot;
                const double inv_rhrc = rcp(rhrc);

                double ewet_c = 0.0;                                              // :306-310
                if (rh >= 70)
                    ewet_c = dzs * (sx * ac + rho_cp_vpd * fc * inv_rhrc) * fwet /
                             ((sx + p_cp_001_over_l622 * rhc * inv_rhrc) * PM_LAMBDA1);
                const double inv_rasoil = (rtot + rr) / (rtot * rr);              // 1 / rasoil, :312
                const double nsoil = dzs * (sx * asoil + rho_cp_vpd * (1 - fc) * inv_rasoil);
                const double inv_dsoil = rcp((sx + PM_GAMMA * rtot * inv_rasoil) * PM_LAMBDA1);
                // ewet_soil + esoilpot * (rh/100)^(vpd/beta)  (:314-323)
                const double ex = vpd * sh.inv_beta[l];
                const double pw = (ex == 0.0) ? 1.0 : exp(ex * log_rh);
                const double esoil = nsoil * inv_dsoil * (fwet + omf * pw);
                double trans = 0.0;                                               // :326-328
                if (fc != 0)
                    trans = dzs * (sx * ac + rho_cp_vpd * fc * inv_ra) * omf /
                            ((sx + PM_GAMMA * (1 + rs * inv_ra)) * PM_LAMBDA1);
                eet = trans + ewet_c + esoil;                                     // :330
                if (eet < 0.0) eet = 0.0;                                         // :332
            }
            acc = acc + eet * __ldg(lct_y + (size_t)l * ld);                      // :467-470
        }
        stg_stream(pet + off + c, acc * inv_tot);                                 // :470
    }
}

void launch_pm_pet_fast(const double *tair, const double *tmin, const double *rhs, const double *wind,
                        const double *rsds, const double *rlds, const double *lct, const double *elev,
                        const int *prev_idx, const PmTab *tab, double *pet, int ncell, int nyears, int ld,
                        int start_year, cudaStream_t s) {
    pm_pet_fast_kernel<<<dim3(ceil_div(ncell, 128), nyears), 128, 0, s>>>(tair, tmin, rhs, wind, rsds, rlds, lct, elev,
                                                                         prev_idx, tab, pet, ncell, ld, start_year);
}

}  // namespace xan
