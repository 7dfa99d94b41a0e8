// Bound-constrained limited-memory BFGS for a handful of variables, as a re-entrant state machine.
//
// Replaces, for the step2_fitgps hot path, the reference's per-start call
//   scipy.optimize.minimize(obj_func, theta0, method="L-BFGS-B", jac=True, bounds=bounds)
// (sklearn/gaussian_process/_gpr.py:658-668; defaults maxcor=10, ftol=2.22e-9, gtol=1e-5,
// maxfun=maxiter=15000, maxls=20 -- scipy/optimize/_lbfgsb_py.py:272-275).
//
// It is a from-scratch implementation of the published algorithm (Byrd, Lu, Nocedal, Zhu 1995;
// Morales & Nocedal 2011 for the projected subspace step; More & Thuente 1994 line search):
// generalized Cauchy point -> subspace minimisation over the free variables -> projection /
// backtracking -> More-Thuente line search (ftol 1e-3, gtol 0.9, xtol 0.1) -> BFGS pair update,
// with the same termination tests (projected-gradient <= pgtol; relative reduction <= factr*eps).
// Because the dimension is tiny (3 hyper-parameters) the quasi-Newton matrix B is formed densely
// by replaying the stored (s, y) pairs on B0 = theta*I, which is algebraically identical to the
// compact representation used by the Fortran/C code.
//
// Reverse communication lets the caller evaluate f and g for ALL live optimisers in one batched
// GPU launch per lock-step round (BASELINE.json north_star (3)).
//
// Every function is __host__ __device__: the lock-step driver of the large-matrix path runs the state machines on
// the host (gpbo_fit_host, gpbo_optpool_*), the persistent kernel of the small-matrix path (kernels_small.cuh)
// runs the very same code on thread 0 of each CTA.
#pragma once
#include <cmath>
#include <cfloat>

#ifdef __CUDACC__
#define GPBO_HD __host__ __device__
#else
#define GPBO_HD
#endif

namespace gpbo {

// exact, so host and device runs agree bit for bit
GPBO_HD inline double lb_max(double a, double b) { return a > b ? a : b; }
GPBO_HD inline double lb_min(double a, double b) { return a < b ? a : b; }
GPBO_HD inline double lb_abs(double a) { return a < 0.0 ? -a : (a == 0.0 ? 0.0 : a); }
GPBO_HD inline bool lb_finite(double a) { return (a - a) == 0.0; }
GPBO_HD inline double lb_sqrt(double a) { return sqrt(a); }
template <class T>
GPBO_HD inline void lb_swap(T& a, T& b) { T c = a; a = b; b = c; }

constexpr int LB_N = 3;
constexpr int LB_M = 10;

enum LbStatus {
    LB_RUNNING = -1,
    LB_CONV_PGTOL = 0,      // CONVERGENCE: NORM_OF_PROJECTED_GRADIENT_<=_PGTOL
    LB_CONV_FACTR = 1,      // CONVERGENCE: REL_REDUCTION_OF_F_<=_FACTR*EPSMCH
    LB_ABNORMAL = 2,        // ABNORMAL_TERMINATION_IN_LNSRCH
    LB_MAXITER = 3,
    LB_MAXFUN = 4,
    LB_BAD_START = 5        // objective not finite at the starting point
};

struct LbOptions {
    double factr = 2.220446049250313e-09 / DBL_EPSILON;  // ftol / eps
    double pgtol = 1e-5;
    int maxiter = 15000;
    int maxfun = 15000;
    int maxls = 20;
};

// ---- More-Thuente line search (dcsrch / dcstep) -------------------------------------------
struct LineSearch {
    bool brackt;
    int stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
    enum Task { FG, CONVERGED, WARNING, ERROR } task;

    GPBO_HD static void step(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp,
                     double fp, double dp, bool& brackt, double stpmin, double stpmax) {
        const double sgnd = dp * (dx / lb_abs(dx));
        double stpf, stpc, stpq, theta, s, gamma, p, q, r;
        if (!lb_finite(fp) || !lb_finite(dp)) {
            // Objective undefined at the trial point (Cholesky failed): bisect back towards stx.
            brackt = true;
            sty = stp; fy = fp; dy = dp;
            stp = stx + 0.5 * (stp - stx);
            return;
        }
        if (fp > fx) {
            theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            s = lb_max(lb_abs(theta), lb_max(lb_abs(dx), lb_abs(dp)));
            gamma = s * lb_sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
            if (stp < stx) gamma = -gamma;
            p = (gamma - dx) + theta;
            q = ((gamma - dx) + gamma) + dp;
            r = p / q;
            stpc = stx + r * (stp - stx);
            stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
            if (lb_abs(stpc - stx) < lb_abs(stpq - stx)) stpf = stpc;
            else stpf = stpc + (stpq - stpc) / 2.0;
            brackt = true;
        } else if (sgnd < 0.0) {
            theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            s = lb_max(lb_abs(theta), lb_max(lb_abs(dx), lb_abs(dp)));
            gamma = s * lb_sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
            if (stp > stx) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dx;
            r = p / q;
            stpc = stp + r * (stx - stp);
            stpq = stp + (dp / (dp - dx)) * (stx - stp);
            if (lb_abs(stpc - stp) > lb_abs(stpq - stp)) stpf = stpc;
            else stpf = stpq;
            brackt = true;
        } else if (lb_abs(dp) < lb_abs(dx)) {
            theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
            s = lb_max(lb_abs(theta), lb_max(lb_abs(dx), lb_abs(dp)));
            gamma = s * lb_sqrt(lb_max(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
            if (stp > stx) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = (gamma + (dx - dp)) + gamma;
            r = p / q;
            if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
            else if (stp > stx) stpc = stpmax;
            else stpc = stpmin;
            stpq = stp + (dp / (dp - dx)) * (stx - stp);
            if (brackt) {
                if (lb_abs(stpc - stp) < lb_abs(stpq - stp)) stpf = stpc;
                else stpf = stpq;
                if (stp > stx) stpf = lb_min(stp + 0.66 * (sty - stp), stpf);
                else stpf = lb_max(stp + 0.66 * (sty - stp), stpf);
            } else {
                if (lb_abs(stpc - stp) > lb_abs(stpq - stp)) stpf = stpc;
                else stpf = stpq;
                stpf = lb_min(stpmax, stpf);
                stpf = lb_max(stpmin, stpf);
            }
        } else {
            if (brackt) {
                theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
                s = lb_max(lb_abs(theta), lb_max(lb_abs(dy), lb_abs(dp)));
                gamma = s * lb_sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
                if (stp > sty) gamma = -gamma;
                p = (gamma - dp) + theta;
                q = ((gamma - dp) + gamma) + dy;
                r = p / q;
                stpc = stp + r * (sty - stp);
                stpf = stpc;
            } else if (stp > stx) stpf = stpmax;
            else stpf = stpmin;
        }
        if (fp > fx) {
            sty = stp; fy = fp; dy = dp;
        } else {
            if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
            stx = stp; fx = fp; dx = dp;
        }
        stp = stpf;
    }

    GPBO_HD void start(double f, double g, double& stp, double stpmax_) {
        (void)stpmax_;
        brackt = false;
        stage = 1;
        finit = f; ginit = g; gtest = ftol * ginit;
        width = stpmax_ - 0.0;
        width1 = width / 0.5;
        stx = 0.0; fx = finit; gx = ginit;
        sty = 0.0; fy = finit; gy = ginit;
        stmin = 0.0;
        stmax = stp + 4.0 * stp;
        task = FG;
    }

    static constexpr double ftol = 1e-3, gtol = 0.9, xtol = 0.1;

    GPBO_HD void iterate(double f, double g, double& stp, double stpmin, double stpmax) {
        const double ftest = finit + stp * gtest;
        if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
        task = FG;
        if (brackt && (stp <= stmin || stp >= stmax)) task = WARNING;          // rounding errors prevent progress
        if (brackt && stmax - stmin <= xtol * stmax) task = WARNING;           // xtol test satisfied
        if (stp == stpmax && f <= ftest && g <= gtest) task = WARNING;         // stp = stpmax
        if (stp == stpmin && (f > ftest || g >= gtest)) task = WARNING;        // stp = stpmin
        if (f <= ftest && lb_abs(g) <= gtol * (-ginit)) task = CONVERGED;
        if (task != FG) return;
        if (stage == 1 && f <= fx && f > ftest) {
            const double fm = f - stp * gtest;
            double fxm = fx - stx * gtest, fym = fy - sty * gtest;
            const double gm = g - gtest;
            double gxm = gx - gtest, gym = gy - gtest;
            step(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax);
            fx = fxm + stx * gtest; fy = fym + sty * gtest;
            gx = gxm + gtest; gy = gym + gtest;
        } else {
            step(stx, fx, gx, sty, fy, gy, stp, f, g, brackt, stmin, stmax);
        }
        if (brackt) {
            if (lb_abs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
            width1 = width;
            width = lb_abs(sty - stx);
        }
        if (brackt) {
            stmin = lb_min(stx, sty);
            stmax = lb_max(stx, sty);
        } else {
            stmin = stp + 1.1 * (stp - stx);
            stmax = stp + 4.0 * (stp - stx);
        }
        stp = lb_max(stp, stpmin);
        stp = lb_min(stp, stpmax);
        if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= xtol * stmax)) stp = stx;
    }
};

// ---- the optimiser ----------------------------------------------------------------------------
struct Lbfgsb {
    // problem
    double l[LB_N], u[LB_N];
    LbOptions opt;
    // iterate
    double x[LB_N], f, g[LB_N];
    int status = LB_RUNNING;
    int nfev = 0, nit = 0;
    // memory
    double S[LB_M][LB_N], Y[LB_M][LB_N];
    int col = 0, head = 0;
    double theta = 1.0;
    // line-search state
    LineSearch ls;
    double d[LB_N], z[LB_N], t[LB_N], r[LB_N];
    double fold, gd, gdold, stp, stpmx, dnorm;
    int ifun, iback;
    bool first_eval = true;

    // Start at x0 (clipped into the box like scipy, _lbfgsb_py.py:359). The caller must then
    // evaluate f, g at `x` and call feed().
    GPBO_HD void init(const double* x0, const double* lo, const double* hi, const LbOptions& o) {
        opt = o;
        if (opt.maxls < 1) opt.maxls = 1;
        for (int i = 0; i < LB_N; ++i) {
            l[i] = lo[i]; u[i] = hi[i];
            x[i] = lb_min(lb_max(x0[i], l[i]), u[i]);
        }
        status = LB_RUNNING; nfev = 0; nit = 0; col = 0; head = 0; theta = 1.0; first_eval = true;
    }
    GPBO_HD bool running() const { return status == LB_RUNNING; }

    // Best point so far: the result once terminated; while running (a driver that stops after a fixed number of
    // rounds) the last accepted iterate, i.e. the start of the current line search.
    GPBO_HD void current_best(double* xb, double* fb) const {
        const bool at_iterate = status == LB_RUNNING && !first_eval;
        for (int i = 0; i < LB_N; ++i) xb[i] = at_iterate ? t[i] : x[i];
        *fb = at_iterate ? fold : (first_eval ? HUGE_VAL : f);
    }

    GPBO_HD double projgr() const {
        double nrm = 0.0;
        for (int i = 0; i < LB_N; ++i) {
            double gi = g[i];
            if (gi < 0.0) gi = lb_max(x[i] - u[i], gi);
            else gi = lb_min(x[i] - l[i], gi);
            nrm = lb_max(nrm, lb_abs(gi));
        }
        return nrm;
    }

    GPBO_HD void dense_B(double B[LB_N][LB_N]) const {
        for (int i = 0; i < LB_N; ++i)
            for (int j = 0; j < LB_N; ++j) B[i][j] = (i == j) ? theta : 0.0;
        for (int k = 0; k < col; ++k) {
            const int idx = (head + k) % LB_M;
            const double* s = S[idx];
            const double* y = Y[idx];
            double Bs[LB_N], sBs = 0.0, ys = 0.0;
            for (int i = 0; i < LB_N; ++i) {
                Bs[i] = 0.0;
                for (int j = 0; j < LB_N; ++j) Bs[i] += B[i][j] * s[j];
            }
            for (int i = 0; i < LB_N; ++i) { sBs += s[i] * Bs[i]; ys += y[i] * s[i]; }
            for (int i = 0; i < LB_N; ++i)
                for (int j = 0; j < LB_N; ++j) B[i][j] += -Bs[i] * Bs[j] / sBs + y[i] * y[j] / ys;
        }
    }

    // Generalized Cauchy point along the projected steepest-descent path. iwhere: 0 free, 1 at lower, 2 at upper.
    GPBO_HD void cauchy(const double B[LB_N][LB_N], double sbgnrm, double* xcp, int* iwhere) const {
        for (int i = 0; i < LB_N; ++i) { xcp[i] = x[i]; iwhere[i] = 0; }
        if (sbgnrm <= 0.0) return;
        double dd[LB_N], tb[LB_N];
        int order[LB_N], nbreak = 0;
        for (int i = 0; i < LB_N; ++i) {
            const double neggi = -g[i];
            const double tl = x[i] - l[i], tu = u[i] - x[i];
            const bool xlower = tl <= 0.0, xupper = tu <= 0.0;
            if (xlower && neggi <= 0.0) iwhere[i] = 1;
            else if (xupper && neggi >= 0.0) iwhere[i] = 2;
            else if (lb_abs(neggi) <= 0.0) iwhere[i] = -3;
            if (iwhere[i] != 0) { dd[i] = 0.0; tb[i] = HUGE_VAL; continue; }
            dd[i] = neggi;
            tb[i] = neggi < 0.0 ? tl / (-neggi) : tu / neggi;
            order[nbreak++] = i;
        }
        for (int a = 1; a < nbreak; ++a)                       // insertion sort of <= 3 breakpoints
            for (int b = a; b > 0 && tb[order[b]] < tb[order[b - 1]]; --b) lb_swap(order[b], order[b - 1]);
        auto derivs = [&](const double* zz, double& f1, double& f2) {
            f1 = 0.0; f2 = 0.0;
            for (int i = 0; i < LB_N; ++i) {
                double Bz = 0.0, Bd = 0.0;
                for (int j = 0; j < LB_N; ++j) { Bz += B[i][j] * zz[j]; Bd += B[i][j] * dd[j]; }
                f1 += dd[i] * (g[i] + Bz);
                f2 += dd[i] * Bd;
            }
        };
        double zz[LB_N] = {0.0, 0.0, 0.0};
        double f1, f2;
        derivs(zz, f1, f2);
        const double f2_org = f2;
        double dtm = -f1 / f2, tsum = 0.0;
        bool all_fixed = false;
        for (int b = 0; b < nbreak; ++b) {
            const int ibp = order[b];
            const double dt = tb[ibp] - tsum;
            if (dtm < dt) break;
            tsum += dt;
            for (int i = 0; i < LB_N; ++i) zz[i] += dt * dd[i];
            const double dibp = dd[ibp];
            dd[ibp] = 0.0;
            if (dibp > 0.0) { xcp[ibp] = u[ibp]; iwhere[ibp] = 2; }
            else { xcp[ibp] = l[ibp]; iwhere[ibp] = 1; }
            zz[ibp] = xcp[ibp] - x[ibp];
            if (b == nbreak - 1) { all_fixed = true; dtm = 0.0; break; }
            derivs(zz, f1, f2);
            f2 = lb_max(DBL_EPSILON * f2_org, f2);
            dtm = -f1 / f2;
        }
        if (!all_fixed) {
            dtm = lb_max(dtm, 0.0);
            tsum += dtm;
        }
        for (int i = 0; i < LB_N; ++i)
            if (dd[i] != 0.0) xcp[i] = x[i] + tsum * dd[i];
    }

    // Subspace minimisation over the free variables, followed by projection (or backtracking).
    GPBO_HD void subsm(const double B[LB_N][LB_N], const int* iwhere, double* xcp) const {
        int fr[LB_N], nf = 0;
        for (int i = 0; i < LB_N; ++i)
            if (iwhere[i] <= 0) fr[nf++] = i;
        if (nf == 0) return;
        double rr[LB_N], M[LB_N][LB_N + 1];
        for (int a = 0; a < nf; ++a) {
            const int i = fr[a];
            double Bz = 0.0;
            for (int j = 0; j < LB_N; ++j) Bz += B[i][j] * (xcp[j] - x[j]);
            rr[a] = -(g[i] + Bz);
            for (int b = 0; b < nf; ++b) M[a][b] = B[i][fr[b]];
            M[a][nf] = rr[a];
        }
        // Gaussian elimination with partial pivoting (nf <= 3)
        for (int k = 0; k < nf; ++k) {
            int piv = k;
            for (int a = k + 1; a < nf; ++a)
                if (lb_abs(M[a][k]) > lb_abs(M[piv][k])) piv = a;
            if (M[piv][k] == 0.0) return;
            if (piv != k)
                for (int b = 0; b <= nf; ++b) lb_swap(M[k][b], M[piv][b]);
            for (int a = k + 1; a < nf; ++a) {
                const double fct = M[a][k] / M[k][k];
                for (int b = k; b <= nf; ++b) M[a][b] -= fct * M[k][b];
            }
        }
        double dsub[LB_N];
        for (int a = nf - 1; a >= 0; --a) {
            double s = M[a][nf];
            for (int b = a + 1; b < nf; ++b) s -= M[a][b] * dsub[b];
            dsub[a] = s / M[a][a];
        }
        for (int a = 0; a < nf; ++a)
            if (!lb_finite(dsub[a])) return;
        // projected point
        double xp[LB_N];
        for (int i = 0; i < LB_N; ++i) xp[i] = xcp[i];
        bool projected = false;
        double xn[LB_N];
        for (int i = 0; i < LB_N; ++i) xn[i] = xcp[i];
        for (int a = 0; a < nf; ++a) {
            const int k = fr[a];
            const double v = xcp[k] + dsub[a];
            xn[k] = lb_max(l[k], lb_min(u[k], v));
            if (xn[k] != v) projected = true;
        }
        double ddp = 0.0;
        if (projected)
            for (int i = 0; i < LB_N; ++i) ddp += (xn[i] - x[i]) * g[i];
        if (projected && ddp > 0.0) {
            // projection is not a descent direction: truncate the Newton step at the first bound hit
            double alpha = 1.0, temp1 = alpha;
            int ibd = -1;
            for (int a = 0; a < nf; ++a) {
                const int k = fr[a];
                const double dk = dsub[a];
                if (dk < 0.0) {
                    const double temp2 = l[k] - xp[k];
                    if (temp2 >= 0.0) temp1 = 0.0;
                    else if (dk * alpha < temp2) temp1 = temp2 / dk;
                } else if (dk > 0.0) {
                    const double temp2 = u[k] - xp[k];
                    if (temp2 <= 0.0) temp1 = 0.0;
                    else if (dk * alpha > temp2) temp1 = temp2 / dk;
                }
                if (temp1 < alpha) { alpha = temp1; ibd = a; }
            }
            for (int i = 0; i < LB_N; ++i) xn[i] = xp[i];
            if (alpha < 1.0 && ibd >= 0) {
                const int k = fr[ibd];
                if (dsub[ibd] > 0.0) { xn[k] = u[k]; dsub[ibd] = 0.0; }
                else if (dsub[ibd] < 0.0) { xn[k] = l[k]; dsub[ibd] = 0.0; }
            }
            for (int a = 0; a < nf; ++a) xn[fr[a]] += alpha * dsub[a];
        }
        for (int i = 0; i < LB_N; ++i) xcp[i] = xn[i];
    }

    GPBO_HD void reset_memory() { col = 0; head = 0; theta = 1.0; }

    // Build the search direction for the current iterate and start the line search.
    // Returns false when the optimiser terminated instead.
    GPBO_HD bool new_iteration() {
        for (int attempt = 0; attempt < 2; ++attempt) {
            double B[LB_N][LB_N];
            dense_B(B);
            bool okB = true;
            for (int i = 0; i < LB_N; ++i)
                for (int j = 0; j < LB_N; ++j)
                    if (!lb_finite(B[i][j])) okB = false;
            if (!okB) { reset_memory(); dense_B(B); }
            int iwhere[LB_N];
            cauchy(B, projgr(), z, iwhere);
            if (col > 0) subsm(B, iwhere, z);
            double dtd = 0.0;
            for (int i = 0; i < LB_N; ++i) { d[i] = z[i] - x[i]; dtd += d[i] * d[i]; }
            dnorm = lb_sqrt(dtd);
            stpmx = 1e10;
            if (nit == 0) stpmx = 1.0;
            else
                for (int i = 0; i < LB_N; ++i) {
                    const double a1 = d[i];
                    if (a1 < 0.0) {
                        const double a2 = l[i] - x[i];
                        if (a2 >= 0.0) stpmx = 0.0;
                        else if (a1 * stpmx < a2) stpmx = a2 / a1;
                    } else if (a1 > 0.0) {
                        const double a2 = u[i] - x[i];
                        if (a2 <= 0.0) stpmx = 0.0;
                        else if (a1 * stpmx > a2) stpmx = a2 / a1;
                    }
                }
            stp = 1.0;   // every variable is boxed
            for (int i = 0; i < LB_N; ++i) { t[i] = x[i]; r[i] = g[i]; }
            fold = f;
            ifun = 0; iback = 0;
            gd = 0.0;
            for (int i = 0; i < LB_N; ++i) gd += g[i] * d[i];
            gdold = gd;
            if (gd >= 0.0 || !(dnorm > 0.0)) {
                // not a descent direction
                if (col == 0) { status = LB_ABNORMAL; return false; }
                reset_memory();
                continue;
            }
            ls.start(f, gd, stp, stpmx);
            // first trial point of the line search: advance_trial() with ifun = 0, written out so that the call
            // graph stays acyclic (the device build then needs no recursion stack); maxls >= 1 is enforced in init()
            ifun = 1; iback = 0;
            if (stp == 1.0) for (int i = 0; i < LB_N; ++i) x[i] = z[i];
            else for (int i = 0; i < LB_N; ++i) x[i] = stp * d[i] + t[i];
            return true;
        }
        status = LB_ABNORMAL;
        return false;
    }

    // Move x to the next trial point of the line search; false if the evaluation budget is spent.
    GPBO_HD bool advance_trial() {
        ifun += 1;
        iback = ifun - 1;
        if (iback >= opt.maxls) return line_search_failed();
        if (stp == 1.0) for (int i = 0; i < LB_N; ++i) x[i] = z[i];
        else for (int i = 0; i < LB_N; ++i) x[i] = stp * d[i] + t[i];
        return true;
    }

    GPBO_HD bool line_search_failed() {
        for (int i = 0; i < LB_N; ++i) { x[i] = t[i]; g[i] = r[i]; }
        f = fold;
        if (col == 0) { status = LB_ABNORMAL; return false; }
        reset_memory();
        return new_iteration();
    }

    // Provide f(x), g(x) for the current x. Afterwards either running() is false or x holds the next
    // point to evaluate.
    GPBO_HD void feed(double fv, const double* gv) {
        nfev += 1;
        if (first_eval) {
            first_eval = false;
            f = fv;
            for (int i = 0; i < LB_N; ++i) g[i] = gv[i];
            if (!lb_finite(fv)) { status = LB_BAD_START; return; }
            if (projgr() <= opt.pgtol) { status = LB_CONV_PGTOL; return; }
            new_iteration();
            return;
        }
        f = fv;
        for (int i = 0; i < LB_N; ++i) g[i] = gv[i];
        gd = 0.0;
        for (int i = 0; i < LB_N; ++i) gd += g[i] * d[i];
        ls.iterate(fv, gd, stp, 0.0, stpmx);
        if (ls.task == LineSearch::FG) {
            advance_trial();
            return;
        }
        if (!lb_finite(fv)) { line_search_failed(); return; }
        // line search finished: x is the new iterate
        nit += 1;
        const double sbgnrm = projgr();
        if (sbgnrm <= opt.pgtol) { status = LB_CONV_PGTOL; return; }
        const double ddum = lb_max(lb_max(lb_abs(fold), lb_abs(f)), 1.0);
        if (fold - f <= DBL_EPSILON * opt.factr * ddum) { status = LB_CONV_FACTR; return; }
        if (nit >= opt.maxiter) { status = LB_MAXITER; return; }
        if (nfev > opt.maxfun) { status = LB_MAXFUN; return; }
        // BFGS pair
        double yv[LB_N], sv[LB_N], rrn = 0.0, dr, dd_;
        for (int i = 0; i < LB_N; ++i) { yv[i] = g[i] - r[i]; rrn += yv[i] * yv[i]; }
        if (stp == 1.0) { dr = gd - gdold; dd_ = -gdold; for (int i = 0; i < LB_N; ++i) sv[i] = d[i]; }
        else { dr = (gd - gdold) * stp; dd_ = -gdold * stp; for (int i = 0; i < LB_N; ++i) sv[i] = stp * d[i]; }
        if (dr > DBL_EPSILON * dd_) {
            int slot;
            if (col < LB_M) { slot = (head + col) % LB_M; col += 1; }
            else { slot = head; head = (head + 1) % LB_M; }
            for (int i = 0; i < LB_N; ++i) { S[slot][i] = sv[i]; Y[slot][i] = yv[i]; }
            theta = rrn / dr;
        }
        new_iteration();
    }
};

}  // namespace gpbo
