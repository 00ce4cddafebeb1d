// Host check of crlk_crmath.h against the C library (the reference's arithmetic):
//   g++ -O2 -std=c++17 -ffp-contract=off -I crl_kino_b200/csrc tools/cr_check.cpp -o /tmp/cr_check -lm && /tmp/cr_check [n]
// Mismatches are adjudicated with the 64-bit-mantissa long double functions.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "crlk_crmath.h"

static const uint64_t atan_bits[] = CRLK_CR_ATAN_INIT;
static const uint64_t sc_bits[] = CRLK_CR_SC_INIT;
static double atan_tab[sizeof(atan_bits) / 8], sc_tab[sizeof(sc_bits) / 8];
static long slow_atan = 0, slow_sc = 0;

int main(int argc, char **argv)
{
    long n = argc > 1 ? atol(argv[1]) : 20000000;
    memcpy(atan_tab, atan_bits, sizeof(atan_bits));
    memcpy(sc_tab, sc_bits, sizeof(sc_bits));
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    long bad = 0, lib_wrong = 0, mine_wrong = 0;
    for (long i = 0; i < n; i++) {
        double y, x;
        switch (i & 3) {
        case 0: y = 40 * u(rng); x = 40 * u(rng); break;
        case 1: y = 1e-3 * u(rng); x = 40 * u(rng); break;
        case 2: { x = 40 * u(rng); int k = (int)(32.5 + 32.5 * u(rng)); y = x * (k / 64.0 + 1e-9 * u(rng)) * (u(rng) < 0 ? -1 : 1); break; }
        default: { double gx = 25 * u(rng), gy = 25 * u(rng); y = gy - 0.2 * u(rng); x = gx - 0.2 * u(rng); }
        }
        if (x == 0 && y == 0) continue;
        double a = crm_atan2_core(y, x, atan_tab), b = atan2(y, x);
        if (a != b) {
            bad++;
            long double t = atan2l((long double)y, (long double)x);
            double tr = (double)t;
            if (tr == a) lib_wrong++; else if (tr == b) mine_wrong++;
            if (bad <= 5) printf("atan2(%a, %a): mine %a lib %a ld %a\n", y, x, a, b, tr);
        }
    }
    printf("atan2: %ld args, %ld differ from libm (libm not correctly rounded: %ld, mine: %ld)\n", n, bad, lib_wrong, mine_wrong);
    bad = lib_wrong = mine_wrong = 0;
    for (long i = 0; i < n; i++) {
        double th;
        switch (i & 3) {
        case 0: th = 3.141592653589793 * u(rng); break;
        case 1: th = 1e-4 * u(rng); break;
        case 2: th = copysign(3.141592653589793 - 1e-4 * fabs(u(rng)), u(rng)); break;
        default: th = 1.5707963267948966 + 1e-3 * u(rng);
        }
        double s, c;
        if (!crm_sincos_core(th, sc_tab, &s, &c)) { printf("range %a\n", th); continue; }
        double ls = sin(th), lc = cos(th);
        if (s != ls || c != lc) {
            bad++;
            double ts = (double)sinl((long double)th), tc = (double)cosl((long double)th);
            bool mine_ok = ts == s && tc == c, lib_ok = ts == ls && tc == lc;
            if (mine_ok && !lib_ok) lib_wrong++; else if (lib_ok && !mine_ok) mine_wrong++;
            if (bad <= 5) printf("sincos(%a): mine %a %a lib %a %a ld %a %a\n", th, s, c, ls, lc, ts, tc);
        }
    }
    printf("sincos: %ld args, %ld differ from libm (libm not correctly rounded: %ld, mine: %ld)\n", n, bad, lib_wrong, mine_wrong);
    return 0;
}
