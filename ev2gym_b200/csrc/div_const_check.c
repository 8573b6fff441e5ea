/* Host check of ev2b_div_c (ev2b_math.h) against IEEE division; built and run by tests/test_div_const.py.
 * usage: div_const_check <n> <d1> <d2> ...   prints the number of mismatches per divisor. */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include "ev2b_math.h"

static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t next(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

int main(int argc, char **argv) {
    long n = atol(argv[1]);
    int bad_total = 0;
    for (int a = 2; a < argc; ++a) {
        volatile double d = atof(argv[a]);
        volatile double rd = 1.0 / d;
        long bad = 0;
        for (long i = 0; i < n; ++i) {
            uint64_t u = next();
            double x;
            if (i % 3 == 0) { /* lattice points: k/100 style values as battery levels are */
                x = (double)(u % 10000000ull) / 100.0;
            } else if (i % 3 == 1) { /* random mantissa, moderate exponent */
                uint64_t bits = ((uint64_t)(1023 - 20 + (u >> 58)) << 52) | (u & 0xFFFFFFFFFFFFFull);
                memcpy(&x, &bits, 8);
            } else {
                x = ((double)(u >> 11) / 9007199254740992.0) * 200.0 - 100.0;
            }
            volatile double want = x / d;
            double got = ev2b_div_c(x, d, rd);
            if (got != want) { if (bad < 3) fprintf(stderr, "d=%g x=%.17g want=%.17g got=%.17g\n", d, x, want, got); ++bad; }
        }
        printf("%s %ld\n", argv[a], bad);
        bad_total += bad != 0;
    }
    return bad_total;
}
