// ev2b_math.h -- exact float64 division by a load-time constant (host + device).
//
// x / d with RN(1/d) known:  q = RN(x * rd);  r = x - d*q (exact, one FMA);  q' = RN(q + r*rd).
// This is the tail of the IEEE-754 division sequence ptxas itself emits for `x / d`
// (MUFU.RCP64H + Newton steps to get ~RN(1/d), then exactly these three operations, see
// `cuobjdump -sass`): with rd = RN(1/d) computed once (host `1.0 / d`) the quotient is the
// correctly rounded one for all normal-range operands (Markstein's correction step), so results
// are bit-identical to `/` while costing 3 fp64 issues instead of ~15 instructions.
// tests/test_div_const.py checks it against `/` on 10^7 operands per divisor used by the engine.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define EV2B_HD __host__ __device__ __forceinline__
#else
#define EV2B_HD static inline
#endif

EV2B_HD double ev2b_div_c(double x, double d, double rd) {
    const double q = x * rd;
    const double r = fma(-d, q, x);
    return fma(rd, r, q);
}
