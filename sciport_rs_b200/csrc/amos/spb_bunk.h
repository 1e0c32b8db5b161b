// spb_bunk.h — K_fnu(z) for fnu > fnul by the uniform asymptotic expansions
// (TOMS 644 "ZBUNK" -> "ZUNK1"/"ZUNK2").
#pragma once
#include "spb_common.h"
#include "spb_uniform.h"
#include "spb_airy.h"

namespace spb {

SPB_FN int bunk(cplx z, real fnu, int kode, int mr, int n, cplx *y, real tol, real elim, real alim) {
    // placeholder until unk1/unk2 land: report "no convergence"
    (void)z; (void)fnu; (void)kode; (void)mr; (void)tol; (void)elim; (void)alim;
    for (int i = 0; i < n; ++i) y[i] = cplx(0.0, 0.0);
    return -2;
}

}  // namespace spb
