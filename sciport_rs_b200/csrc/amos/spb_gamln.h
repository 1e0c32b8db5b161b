// spb_gamln.h — ln Gamma(x), x > 0, as used inside the series / Miller
// normalisations (the DGAMLN companion of TOMS 644): exact table for integer
// 1..100, otherwise shift to x >= 7 and the Stirling series with Bernoulli
// coefficients CF.
#pragma once
#include "spb_common.h"

namespace spb {

// Tables live in function-local statics so that the same header works on the
// host and (as device globals) in CUDA kernels.
#define SPB_TABLE(name, n) static const double name[n]

SPB_FN real gamln(real z, int &ierr) {
#include "spb_tables_gln.inc"
    const real con = 1.83787706640934548;  // ln(2 pi)
    ierr = 0;
    if (!(z > real(0.0))) {
        ierr = 1;
        return real(0.0);
    }
    int nz = 0;
    if (z <= real(101.0)) {
        nz = (int)z;
        real fz = z - real(nz);
        if (fz <= real(0.0) && nz <= 100) return real(GLN[nz - 1]);
    }
    const real wdtol = SPB_TOL;
    // zmin = 7 for binary64 (1.8 + 0.3875*(min(53*log10 2,20)-3) -> int + 1)
    const real zmin = 7.0;
    real zdmy = z, zinc = 0.0;
    if (z < zmin) {
        zinc = zmin - real(nz);
        zdmy = z + zinc;
    }
    real zp = real(1.0) / zdmy;
    real t1 = real(CF[0]) * zp;
    real s = t1;
    if (zp >= wdtol) {
        real zsq = zp * zp;
        real tst = t1 * wdtol;
        for (int k = 1; k < 22; ++k) {
            zp = zp * zsq;
            real trm = real(CF[k]) * zp;
            if (rabs(trm) < tst) break;
            s = s + trm;
        }
    }
    // (library logarithm here on purpose: the Stirling form multiplies it by z ~ 10 and cancels to O(0.1) for the
    // arguments 1 + dnu of the Temme series, which amplifies half an ulp of difference to the CPU restatement)
    if (zinc == real(0.0)) {
        real tlg = log(z);
        return z * (tlg - real(1.0)) + real(0.5) * (con - tlg) + s;
    }
    zp = 1.0;
    int ninc = (int)zinc;
    for (int i = 0; i < ninc; ++i) zp = zp * (z + real(i));
    real tlg = log(zdmy);
    return zdmy * (tlg - real(1.0)) - log(zp) + real(0.5) * (con - tlg) + s;
}

SPB_HD real gamln(real z) {
    int ie;
    return gamln(z, ie);
}

// ln Gamma at the arguments that depend on the order alone: which = 0: 1 + fnu (power series), 1: 1 + fnf, 2: 1 + 2 fnf
// with fnf = fnu - int(fnu) (Neumann normalisation of the Miller routine).  A kernel of a broadcast-order job
// evaluates them once per CTA and hands them to the routine (`gl`), where every point would otherwise evaluate them
// again; the numbers are gamln() of the same arguments.
SPB_FN real order_gln(real fnu, int which) {
    const real fnf = fnu - real((int)fnu);
    const real arg = which == 0 ? fnu + real(1.0) : which == 1 ? real(1.0) + fnf : (fnf + fnf) + real(1.0);
    return gamln(arg);
}

}  // namespace spb
