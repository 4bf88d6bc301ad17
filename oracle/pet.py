"""
PET oracle: Penman-Monteith, Hargreaves-Samani, Thornthwaite (test infrastructure only).

Plain numpy, one cell-month formula at a time, operation order kept identical
to the reference so that results are bitwise equal where the reference's own
numpy call is deterministic.  All arrays are [ncell, nmonths] like the
reference's.
"""

import calendar

import numpy as np

from .calendar_utils import MONTHDAYS, LEAP_MONTHDAYS, gregorian_days

# Penman-Monteith constants, xanthos/pet/penman_monteith.py:76-81
LAMBDA1 = 2.46e6
CP = 1006
SIGMA = 4.9e-3
SIGMA2 = 5.67e-8
GAMMA = 0.67


def numpy_pairwise_sum(x):
    """
    Sum over the last axis in the order numpy's pairwise summation uses for a
    contiguous reduction of length k <= 128 (numpy/core/src/umath/loops_utils.h,
    `pairwise_sum`): k < 8 sequential from 0; otherwise eight running partial
    sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail in order.

    Two reference reductions hit this path and are restated with it:
    the land-cover total `np.sum(self.c.lct, axis=0)` (penman_monteith.py:45; the
    class axis is the contiguous one of the swapped view) and the annual heat
    index `np.add.reduceat` (thornthwaite.py:91; first element + pairwise(rest)).
    """
    k = x.shape[-1]
    if k < 8:
        r = np.zeros(x.shape[:-1])
        for i in range(k):
            r = r + x[..., i]
        return r
    assert k <= 128
    r = [x[..., j] for j in range(8)]
    i = 8
    while i < k - (k % 8):
        for j in range(8):
            r[j] = r[j] + x[..., i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < k:
        res = res + x[..., i]
        i += 1
    return res


def pm_land_cover_year(target_yr, land_cover_years):
    """Index of the land-cover slice used for `target_yr` (penman_monteith.py:32-43)."""
    lc = sorted(land_cover_years)
    if target_yr >= lc[-1]:
        return lc.index(lc[-1])
    return lc.index([x for x in lc if x - target_yr >= -4][0])


def _pm_fwet(rh):
    """penman_monteith.py:165-172."""
    fwet = np.where(rh < 70, 0, rh)
    fwet = np.where(rh >= 70, np.power(rh / 100, 8), fwet)
    fwet = np.where(rh >= 80, np.power(rh / 100, 10), fwet)
    fwet = np.where(rh >= 90, np.power(rh / 100, 12), fwet)
    fwet = np.where(rh >= 95, np.power(rh / 100, 16), fwet)
    return fwet


def pm_pet(inp, ncells, nlcs, start_yr, end_yr, water_idx, snow_idx, land_cover_years):
    """
    Penman-Monteith PET, restating run_pmpet (penman_monteith.py:394-477) with
    SetData (:17-99), et_veg (:223-334), et_water (:337-361), et_snow (:364-377).

    `inp` is a mapping with the DataLoader attribute names (data_load.py:92-135):
    tair_load, TMIN_load, rhs_load, wind_load, rsds_load, rlds_load, tairprev_load
    [N, M]; lct_load [N, nlcs, nyears]; elev [N, 1]; per-class vectors cL, beta,
    rslimit, Tminopen, Tminclose, VPDclose, VPDopen, RBLmin, RBLmax, rc, emiss;
    tables alpha, lai, laimin, laimax [nlcs, 12].
    """
    nmonths_tot = (end_yr - start_yr + 1) * 12
    out = np.zeros((ncells, nmonths_tot))
    elev = np.asarray(inp['elev']).reshape(ncells, 1)
    wind_k = np.power(2 / 10, 0.11)                                       # :99

    for yi, y in enumerate(range(start_yr, end_yr + 1)):
        s, e = yi * 12, yi * 12 + 12
        T = inp['tair_load'][:, s:e]
        Tn = inp['TMIN_load'][:, s:e]
        RH = inp['rhs_load'][:, s:e]
        W = inp['wind_load'][:, s:e]
        Rs = inp['rsds_load'][:, s:e]
        Rl = inp['rlds_load'][:, s:e]
        Tp = inp['tairprev_load'][:, s:e]
        lct = inp['lct_load'][:, :, pm_land_cover_year(y, land_cover_years)]    # [N, nlcs]
        dz = np.array(LEAP_MONTHDAYS if calendar.isleap(y) else MONTHDAYS)    # :57-62 (int)

        # class independent terms
        esx = 6.10588 * np.exp(17.32491 * T / (T + 238.102))                  # :83
        vap = np.multiply(esx, RH / 100)                                       # :86, :235
        sx = (238.1 * 17.325 * esx / np.power((T + 238.1), 2))                 # :89
        p = 101325 * np.power((1 - 0.0065 * elev / 288.15), 5.2558)           # :187  [N,1]
        rcorr = p / (101300 * np.power((273.15 + T) / 293.15, 1.75))          # :229
        gcu = 0.00001 * rcorr                                                  # :230
        vpd = esx - vap                                                        # :121
        rh = RH.copy()
        rh[rh > 99.9999] = 99.9                                                # :205-209
        g = 1.6198 * (T - Tp)
        g[:, 0] = 0                                                            # :212-216
        rho = p / ((T + 273.15) * 287.058)                                     # :271
        rr = rho * CP / (4.0 * SIGMA2 * np.power((T + 273.15), 3))             # :273, :293
        fwet = _pm_fwet(rh)                                                    # :280
        T4 = np.power(T + 273, 4.0)                                            # :158
        wind2 = W * wind_k

        acc = np.zeros((ncells, 12))
        per_class = {}
        for l in range(nlcs):
            if l == water_idx or l == snow_idx:
                continue
            A = inp['alpha'][l][None, :]
            LAI = inp['lai'][l][None, :]
            LAImin = inp['laimin'][l][None, :]
            LAImax = inp['laimax'][l][None, :]
            topen, tclose = inp['Tminopen'][l], inp['Tminclose'][l]
            vopen, vclose = inp['VPDopen'][l], inp['VPDclose'][l]
            rblmin, rblmax = inp['RBLmin'][l], inp['RBLmax'][l]
            rc, cL, beta = inp['rc'][l], inp['cL'][l], inp['beta'][l]
            rslimit, emiss = inp['rslimit'][l], inp['emiss'][l]

            mtmin = np.zeros_like(Tn)                                          # :102-114
            mtmin[Tn >= topen] = 1.0
            mtmin[Tn <= tclose] = 0.1
            xi = (Tn < topen) & (Tn > tclose)
            mtmin = np.where(xi, (Tn - tclose) / (topen - tclose), mtmin)

            mvpd = vpd.copy()                                                  # :117-129
            mvpd = np.where(vpd <= vopen, 1.0, mvpd)
            mvpd = np.where(vpd >= vclose, 0.1, mvpd)
            vi = (vpd > vopen) & (vpd < vclose)
            mvpd = np.where(vi, (vclose - vpd) / (vclose - vopen), mvpd)

            gs1 = cL * mtmin * mvpd * rcorr                                    # :242

            rtotc = np.zeros_like(vpd)                                         # :132-145
            rtotc = np.where(vpd <= vopen, rblmax, rtotc)
            rtotc = np.where(vpd >= vclose, rblmin, rtotc)
            rtotc = np.where(vi, rblmax - (rblmax - rblmin) * (vclose - vpd) / (vclose - vopen), rtotc)

            rnl = SIGMA * T4 * emiss * dz - Rl * 86400 * dz                    # :158
            rn = ((1 - A) * Rs) * 86400 * dz - rnl                             # :159
            a = rn / (86400 * dz)                                              # :160

            fc_denom = np.exp(-0.5 * LAImin) - np.exp(-0.5 * LAImax)           # :257-258
            fc_denom = np.where(fc_denom == 0.0, 1, fc_denom)
            fc = (np.exp(-0.5 * LAImin) - np.exp(-0.5 * LAI)) / fc_denom       # :260
            fc = np.where(fc > 1, 1, fc)
            ac = fc * a                                                        # :263
            asoil = (1 - fc) * a - g                                           # :266

            rtot = rtotc * rcorr                                               # :268-269
            rtot = np.where(rtot > 80, 80, rtot)
            ra = rc * rr / (rc + rr)                                           # :277-278
            ra = np.where(ra > rtot, rtot, ra)

            den = gs1 + 1 / rc + gcu                                           # :192-197
            cc = np.where(den < 0.0001, 10000,
                          np.where(fwet == 1, 0.00001, np.where(LAI < 0.0001, 0.00001, 0)))
            ccx = np.where(cc == 0, 1 / rc * (gs1 + gcu) * LAI * (1 - fwet) / den, cc)
            with np.errstate(divide='ignore'):
                rs = np.where(ccx == 0, 100000, 1 / ccx)                       # :285
            rs = np.where(rs > rslimit, rslimit, rs)                           # :291

            lai_fwet = np.where(LAI * fwet == 0, 1, LAI * fwet)                # :296
            rhc = np.where(LAI > 0.00001, rc / lai_fwet, rslimit)              # :297
            rhc = np.where(rhc > rslimit, rslimit, rhc)                        # :300
            rvc = rhc
            rhrc = rhc * rr / (rhc + rr)                                       # :303-304
            rhrc = np.where(rhrc > rtot, rtot, rhrc)

            apres = dz * 86400 * (sx * ac + rho * CP * vpd * fc / rhrc) * fwet / (
                (sx + p * 0.01 * CP * rvc / (LAMBDA1 * 0.622 * rhrc)) * LAMBDA1)   # :306-307
            ewet_c = np.where(rh >= 70, apres, 0.0)                            # :309-310
            rasoil = rtot * rr / (rtot + rr)                                   # :312
            ewet_soil = 86400 * dz * (sx * asoil + rho * CP * (1 - fc) * vpd / rasoil) * fwet / (
                (sx + GAMMA * rtot / rasoil) * LAMBDA1)                        # :314-315
            esoilpot = 86400 * dz * (sx * asoil + rho * CP * (1 - fc) * vpd / rasoil) * (1 - fwet) / (
                (sx + GAMMA * rtot / rasoil) * LAMBDA1)                        # :316-317
            esoil = ewet_soil + esoilpot * np.power((rh / 100), vpd / beta)    # :323
            trans = dz * 86400 * (sx * ac + rho * CP * vpd * fc / ra) * (1 - fwet) / (
                (sx + GAMMA * (1 + rs / ra)) * LAMBDA1)                        # :326-327
            trans = np.where(fc == 0, 0, trans)                                # :328
            eet = trans + ewet_c + esoil                                       # :330
            eet = np.where(eet < 0.0, 0.0, eet)                                # :332
            per_class[l] = eet

        # open water: alpha row 0, emissivity 0.98 (:337-361); snow: alpha row 6, 0.85 (:364-377)
        def _rad(alpha_row, emiss):
            rnlx = SIGMA * T4 * emiss * dz - Rl * 86400 * dz
            rnx = ((1 - alpha_row) * Rs) * 86400 * dz - rnlx
            rnx = np.where(rnx < 0, 0.0, rnx)
            return rnlx, rnx

        A0 = inp['alpha'][0][None, :]
        rnlx, rnx = _rad(A0, 0.98)
        rsnx = (1 - A0) * Rs * 86400 * dz                                      # :92
        coef = np.where(np.arange(12) <= 5, 0.8, 1.3)[None, :]                 # :347-349
        qtx = 0.5 * rsnx - coef * rnlx
        ax = (rnx - qtx) / (86400 * dz)
        ax = np.where(ax < 0, 0, ax)
        rn2x = rnx / (86400 * dz)
        ewetx = rn2x * dz * 0.6 / 2845
        ewety = dz * 86400 * (sx * ax + GAMMA * 6.43 * (0.5 + 0.54 * wind2) * (esx - vap)) / (
            (sx + GAMMA) * LAMBDA1)
        wat = np.where(T < -1, ewetx, ewety)
        wat = np.where(wat < 0.0, 0.0, wat)

        a_snow = inp['alpha'][6][None, :]                                      # hard-coded row 6 (:377)
        _, rnx_s = _rad(a_snow, 0.85)
        snow = (rnx_s / (86400 * dz)) * dz * 0.6 / 2845
        snow = np.where(snow < 0.0, 0.0, snow)

        per_class[water_idx] = wat                                             # :459-460
        per_class[snow_idx] = snow                                             # :462-464 (written last)
        for l in range(nlcs):                                                  # :467-470, sum in class order
            acc = acc + per_class[l] * lct[:, l][:, None]
        tot = numpy_pairwise_sum(lct)                                          # :45 (pairwise order)
        tot = np.where(tot == 0, 0.01, tot)                                    # :46-47
        out[:, s:e] = acc / tot[:, None]
    return out


def hs_pet(tas, tmax, tmin, lat_deg, start_yr, end_yr):
    """
    Hargreaves-Samani PET (hargreaves_samani.py:31-65 `pet`, :91-119 `execute`).
    The reference is a scalar double loop; this is the same formula on arrays.
    """
    n, m = tas.shape
    j = np.array([15, 45, 75, 105, 135, 165, 195, 225, 255, 285, 315, 345])
    dy = j[np.arange(m) % 12][None, :]                                         # :37-45
    delta = 0.4102 * np.sin(2 * (np.pi / 365) * (dy - 80))                     # :47
    phi = (lat_deg * np.pi / 180)[:, None]                                     # :49
    tn = -np.tan(delta) * np.tan(phi)                                          # :51
    inside = ~((tn < -1.) | (tn > 1.))
    acs = np.where(inside, np.arccos(np.where(inside, tn, 0.0)), 0.0)          # :53-56
    ra = 118 / np.pi * acs + np.cos(phi) * np.cos(delta) * np.sin(acs)         # :59
    pet = 0.408 * 0.0023 * ra * (tas + 17.8) * np.sqrt(np.abs(tmax - tmin))    # :62
    pet = np.where(tas < 0, 0.0, pet)                                          # :33 (NaN tas -> NaN)
    return pet * gregorian_days(start_yr, end_yr)[None, :]                     # :114


def daylight_hours(mth_days, lat_radians):
    """Monthly mean day length, thornthwaite.py:18-44."""
    days = np.arange(sum(mth_days)) + 1
    solar_dec = 0.409 * np.sin(((2 * np.pi / 365.0) * days - 1.39))
    c = -np.tan(lat_radians[:, np.newaxis]) * np.tan(solar_dec[np.newaxis, :])
    sha = np.arccos(np.clip(c, -1, 1))
    hours = sha * (24.0 / np.pi)
    idx = np.roll(np.cumsum(mth_days), 1)
    idx[0] = 0
    return np.add.reduceat(hours, idx, axis=1) / mth_days


def thornthwaite_pet(tas, lat_radians, start_yr, end_yr):
    """
    Thornthwaite PET, thornthwaite.py:47-130, including the `np.repeat`
    day-length tiling of :110 (column k of a non-leap year uses L12[:, k // nyears]).
    Does not mutate `tas` (the reference does, :82).
    """
    tas = np.where(np.isnan(tas) | (tas < 0), 0.0, tas)                        # :82
    n, m = tas.shape
    nyears = end_yr - start_yr + 1
    i = np.power(tas / 5.0, 1.514)                                             # :88
    I = i.reshape(n, nyears, 12)
    Iy = I[:, :, 0] + numpy_pairwise_sum(I[:, :, 1:])                          # :91 (reduceat order)
    a = (.000000675 * Iy ** 3) - (.0000771 * Iy ** 2) + (.0179 * Iy) + .492   # :94
    Im = np.repeat(Iy, 12, axis=1)
    am = np.repeat(a, 12, axis=1)
    ratio = np.divide(10 * tas, Im, out=np.zeros_like(Im), where=(Im != 0))    # :104
    pu = 16 * np.power(ratio, am)                                              # :105
    L12 = daylight_hours(MONTHDAYS, lat_radians)
    L12_leap = daylight_hours(LEAP_MONTHDAYS, lat_radians)
    L = np.empty((n, m))
    N = np.empty(m)
    for k in range(m):
        yr = start_yr + k // 12
        if calendar.isleap(yr):                                                # :113-119
            L[:, k] = L12_leap[:, k % 12]
            N[k] = LEAP_MONTHDAYS[k % 12]
        else:
            L[:, k] = L12[:, k // nyears]                                      # :110 (np.repeat tiling)
            N[k] = MONTHDAYS[k % 12]
    return pu * (L / 12) * (N / 30.0)                                          # :127
