"""CPU check of the FP64 exp / constant-divisor division used by the kernel-matrix element generators
(gp-bayesopinf_b200/csrc/fastmath.h compiles for host and device): error in ulps against long-double references."""
import os
import shutil
import subprocess
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = textwrap.dedent(r'''
    #include "fastmath.h"
    #include <cmath>
    #include <cstdio>
    #include <random>
    int main() {
        std::mt19937_64 rng(7);
        double maxulp = 0; long bad = 0;
        auto check = [&](double x) {
            const double y = gpbo::gpbo_exp(x);
            const long double ref = expl((long double)x);
            if (x < -708.0005) { if (y != 0.0) ++bad; return; }   // flush decided on the high word of x
            const double refd = (double)ref;
            const double ulp = std::nextafter(refd, INFINITY) - refd;
            const double e = std::fabs((double)((long double)y - ref)) / ulp;
            if (e > maxulp) maxulp = e;
        };
        std::uniform_real_distribution<double> U(-700.0, 0.0), V(-40.0, 0.0), W(-1e-3, 0.0), Z(-760.0, -708.0);
        for (long i = 0; i < 2000000; ++i) { check(U(rng)); check(V(rng)); check(W(rng)); check(Z(rng)); }
        check(0.0); check(-0.0); check(-1e-300); check(-1e6); check(-INFINITY);
        long ndiff = 0; double maxd = 0;
        std::uniform_real_distribution<double> A(0.0, 50.0), B(1e-10, 1e4);
        for (long i = 0; i < 2000000; ++i) {
            const double a = A(rng) * A(rng), b = B(rng), q = gpbo::gpbo_div(a, b, 1.0 / b), ref = a / b;
            if (q != ref) { ++ndiff; maxd = std::fmax(maxd, std::fabs(q - ref) / (std::nextafter(ref, INFINITY) - ref)); }
        }
        // the kernels' fast path (constant-bank operands, integer flush): same arithmetic as gpbo_exp for x <= 0
        long neq = 0;
        for (long i = 0; i < 2000000; ++i) {
            const double x = U(rng), y = Z(rng);
            if (gpbo::gpbo_exp_neg(x) != gpbo::gpbo_exp(x)) ++neq;
            if (gpbo::gpbo_exp_neg(y) != gpbo::gpbo_exp(y)) ++neq;
        }
        if (gpbo::gpbo_exp_neg(0.0) != 1.0 || gpbo::gpbo_exp_neg(-1e9) != 0.0 || gpbo::gpbo_exp_neg(-INFINITY) != 0.0) ++neq;
        // the table-driven variant of the stand-alone assembly kernels (10 FP64 instructions): <= 1.5 ulp on [-708, 0]
        static const double TAB[64] = {GPBO_EXP2_TAB_VALUES};
        double tabulp = 0; long tabbad = 0;
        auto check_tab = [&](double x) {
            const double y = gpbo::gpbo_exp_neg_tab(x, TAB);
            if (x < -708.0005) { if (y != 0.0) ++tabbad; return; }
            if (x < -708.0) return;                                     // either flushed or not: both accepted
            const long double ref = expl((long double)x);
            const double refd = (double)ref;
            const double ulp = std::nextafter(refd, INFINITY) - refd;
            tabulp = std::fmax(tabulp, std::fabs((double)((long double)y - ref)) / ulp);
        };
        for (long i = 0; i < 2000000; ++i) { check_tab(U(rng)); check_tab(V(rng)); check_tab(W(rng)); check_tab(Z(rng)); }
        if (gpbo::gpbo_exp_neg_tab(0.0, TAB) != 1.0 || gpbo::gpbo_exp_neg_tab(-1e9, TAB) != 0.0 ||
            gpbo::gpbo_exp_neg_tab(-INFINITY, TAB) != 0.0) ++tabbad;
        std::printf("%.6f %ld %ld %.3f %d %d %ld %.6f %ld\n", maxulp, bad, ndiff, maxd, gpbo::gpbo_exp(0.0) == 1.0,
                    std::isnan(gpbo::gpbo_exp(NAN)) ? 1 : 0, neq, tabulp, tabbad);
    }
''')


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_fast_exp_and_div_error(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    inc = os.path.join(ROOT, "gp-bayesopinf_b200", "csrc")
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-I", inc, "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    maxulp, bad, ndiff, maxd, exp0, nan_ok = float(out[0]), int(out[1]), int(out[2]), float(out[3]), int(out[4]), int(out[5])
    assert int(out[6]) == 0      # gpbo_exp_neg == gpbo_exp bit for bit on x <= 0 (and 0 below the flush threshold)
    assert maxulp <= 1.0         # at most 1 ulp on [-708, 0]
    assert bad == 0              # exact 0 below the flush threshold
    assert ndiff <= 20 and maxd <= 1.0   # division correctly rounded up to rare 1-ulp cases
    assert exp0 == 1 and nan_ok == 1
    assert float(out[7]) <= 1.5 and int(out[8]) == 0      # table-driven exp of the assembly kernels
