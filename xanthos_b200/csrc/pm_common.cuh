// Shared between pet.cu (reference-order kernel, -fmad=false) and pet_pm_fast.cu (optimised kernel).
#pragma once

#include "common.cuh"

namespace xan {

struct PmTab {
    int nlcs, water_idx, snow_idx, pad;
    double cL[XAN_PM_MAX_CLASSES], beta[XAN_PM_MAX_CLASSES], rslimit[XAN_PM_MAX_CLASSES],
        Tminopen[XAN_PM_MAX_CLASSES], Tminclose[XAN_PM_MAX_CLASSES], VPDclose[XAN_PM_MAX_CLASSES],
        VPDopen[XAN_PM_MAX_CLASSES], RBLmin[XAN_PM_MAX_CLASSES], RBLmax[XAN_PM_MAX_CLASSES],
        rc[XAN_PM_MAX_CLASSES], emiss[XAN_PM_MAX_CLASSES];
    double alpha[XAN_PM_MAX_CLASSES][12], lai[XAN_PM_MAX_CLASSES][12], laimin[XAN_PM_MAX_CLASSES][12],
        laimax[XAN_PM_MAX_CLASSES][12];
    unsigned char lc_index[512];   // land-cover slice per simulated year
};

constexpr double PM_LAMBDA1 = 2.46e6;   // :76-81
constexpr double PM_CP = 1006;
constexpr double PM_SIGMA = 4.9e-3;
constexpr double PM_SIGMA2 = 5.67e-8;
constexpr double PM_GAMMA = 0.67;


void launch_pm_pet_fast(const double *tair, const double *tmin, const double *rhs, const double *wind,
                        const double *rsds, const double *rlds, const double *lct, const double *elev,
                        const int *prev_idx, const PmTab *tab, double *pet, int ncell, int nyears, int ld,
                        int start_year, cudaStream_t s);

}  // namespace xan
