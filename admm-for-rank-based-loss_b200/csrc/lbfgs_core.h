// lbfgs_core.h — L-BFGS-B for an UNCONSTRAINED problem, as scipy.optimize.minimize(method="L-BFGS-B") runs it for the
// reference's smooth w-steps (src/util/w_LBFGS.py:48-62: w_solver -> minimize(fun, w0, jac, method='L-BFGS-B',
// options={'maxiter': 1000}); scipy defaults maxcor = 10, ftol = 2.2e-9, gtol = 1e-5, maxls = 20, maxfun = 15000).
//
// scipy / L-BFGS-B 3.0 are third-party and un-vendored (README.md pins scipy 1.10.1).  What is restated here, from the
// published algorithm (Byrd, Lu, Nocedal, Zhu 1995; Zhu, Byrd, Lu, Nocedal 1997; Morales, Nocedal 2011) and the
// MINPACK-2 line search (More', Thuente 1994: dcsrch / dcstep), is its control flow WITHOUT bounds:
//   * first direction -g with first trial step min(1/||d||, stpmx), later directions -H g with trial step 1, where H
//     is the limited-memory BFGS inverse Hessian of the last m accepted pairs with initial scaling 1/theta,
//     theta = y.y / s.y of the latest pair (the compact representation L-BFGS-B uses equals this two-loop
//     recursion in exact arithmetic when no bound is active);
//   * line search dcsrch(ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0, stpmax = 1e10), at most maxls steps;
//   * stop when max|g_i| <= pgtol, or (f_old - f) <= ftol * max(|f_old|, |f|, 1), or iter >= maxiter, or
//     evaluations >= maxfun; pairs with s.y <= eps * (-g_old.d * stp) are skipped.
// Pure host C++ (no CUDA): the library drives it with device f/g evaluations (api.cu), the CPU tests drive it with a
// callback and compare with the installed scipy (tests/test_host.py).
#pragma once
#include <math.h>
#include <string.h>

#include <vector>

namespace rbl_lbfgs {

enum Task { TASK_START, TASK_FG, TASK_CONVERGENCE, TASK_WARNING, TASK_ERROR };

struct LineSearchState {
    bool brackt;
    int stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
};

// MINPACK-2 dcstep: safeguarded step for the line search
inline void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp, double fp,
                   double dp, bool& brackt, double stpmin, double stpmax) {
    const double sgnd = dp * (dx / fabs(dx));
    double stpf, stpc, stpq, theta, s, gamma, p, q, r;
    if (fp > fx) {  // case 1: higher function value — the minimum is bracketed
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        p = (gamma - dx) + theta;
        q = ((gamma - dx) + gamma) + dp;
        r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        if (fabs(stpc - stx) < fabs(stpq - stx)) stpf = stpc;
        else stpf = stpc + (stpq - stpc) / 2.0;
        brackt = true;
    } else if (sgnd < 0.0) {  // case 2: lower value, derivatives of opposite sign — bracketed
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + dx;
        r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
        else stpf = stpq;
        brackt = true;
    } else if (fabs(dp) < fabs(dx)) {  // case 3: lower value, same sign, derivative magnitude decreases
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc;
            else stpf = stpq;
            if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
            else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
        } else {
            if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
            else stpf = stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {  // case 4: lower value, same sign, derivative magnitude does not decrease
        if (brackt) {
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
            gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dy;
            r = p / q;
            stpc = stp + r * (sty - stp);
            stpf = stpc;
        } else if (stp > stx) {
            stpf = stpmax;
        } else {
            stpf = stpmin;
        }
    }
    if (fp > fx) {
        sty = stp;
        fy = fp;
        dy = dp;
    } else {
        if (sgnd < 0.0) {
            sty = stx;
            fy = fx;
            dy = dx;
        }
        stx = stp;
        fx = fp;
        dx = dp;
    }
    stp = stpf;
}

// MINPACK-2 dcsrch: one reverse-communication step of the More'-Thuente line search
inline void dcsrch(double f, double g, double& stp, double ftol, double gtol, double xtol, double stpmin,
                   double stpmax, Task& task, LineSearchState& S) {
    const double xtrapl = 1.1, xtrapu = 4.0, p5 = 0.5, p66 = 0.66;
    if (task == TASK_START) {
        if (stp < stpmin || stp > stpmax || g >= 0.0 || ftol < 0.0 || gtol < 0.0 || xtol < 0.0 || stpmin < 0.0 ||
            stpmax < stpmin) {
            task = TASK_ERROR;
            return;
        }
        S.brackt = false;
        S.stage = 1;
        S.finit = f;
        S.ginit = g;
        S.gtest = ftol * S.ginit;
        S.width = stpmax - stpmin;
        S.width1 = S.width / p5;
        S.stx = 0.0;
        S.fx = S.finit;
        S.gx = S.ginit;
        S.sty = 0.0;
        S.fy = S.finit;
        S.gy = S.ginit;
        S.stmin = 0.0;
        S.stmax = stp + xtrapu * stp;
        task = TASK_FG;
        return;
    }
    const double ftest = S.finit + stp * S.gtest;
    if (S.stage == 1 && f <= ftest && g >= 0.0) S.stage = 2;
    if (S.brackt && (stp <= S.stmin || stp >= S.stmax)) task = TASK_WARNING;   // rounding errors prevent progress
    if (S.brackt && S.stmax - S.stmin <= xtol * S.stmax) task = TASK_WARNING;  // xtol test satisfied
    if (stp == stpmax && f <= ftest && g <= S.gtest) task = TASK_WARNING;      // stp = stpmax
    if (stp == stpmin && (f > ftest || g >= S.gtest)) task = TASK_WARNING;     // stp = stpmin
    if (f <= ftest && fabs(g) <= gtol * (-S.ginit)) task = TASK_CONVERGENCE;
    if (task == TASK_WARNING || task == TASK_CONVERGENCE) return;
    if (S.stage == 1 && f <= S.fx && f > ftest) {  // modified function while the sufficient decrease fails
        double fm = f - stp * S.gtest, fxm = S.fx - S.stx * S.gtest, fym = S.fy - S.sty * S.gtest;
        double gm = g - S.gtest, gxm = S.gx - S.gtest, gym = S.gy - S.gtest;
        dcstep(S.stx, fxm, gxm, S.sty, fym, gym, stp, fm, gm, S.brackt, S.stmin, S.stmax);
        S.fx = fxm + S.stx * S.gtest;
        S.fy = fym + S.sty * S.gtest;
        S.gx = gxm + S.gtest;
        S.gy = gym + S.gtest;
    } else {
        dcstep(S.stx, S.fx, S.gx, S.sty, S.fy, S.gy, stp, f, g, S.brackt, S.stmin, S.stmax);
    }
    if (S.brackt) {
        if (fabs(S.sty - S.stx) >= p66 * S.width1) stp = S.stx + p5 * (S.sty - S.stx);
        S.width1 = S.width;
        S.width = fabs(S.sty - S.stx);
    }
    if (S.brackt) {
        S.stmin = fmin(S.stx, S.sty);
        S.stmax = fmax(S.stx, S.sty);
    } else {
        S.stmin = stp + xtrapl * (stp - S.stx);
        S.stmax = stp + xtrapu * (stp - S.stx);
    }
    stp = fmax(stp, stpmin);
    stp = fmin(stp, stpmax);
    if ((S.brackt && (stp <= S.stmin || stp >= S.stmax)) || (S.brackt && S.stmax - S.stmin <= xtol * S.stmax))
        stp = S.stx;
    task = TASK_FG;
}

struct Result {
    int nit;
    int nfev;
    int status;  // 0 converged (gradient or function test), 1 iteration / evaluation limit, 2 abnormal line search
    double f;
};

inline double dot(const double* a, const double* b, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

// x: start on entry, solution on exit.  FG: int fg(const double* x, double* f, double* g) (non-zero = failure).
template <class FG>
Result minimize(int n, double* x, FG&& fg, int m = 10, int maxiter = 1000, int maxfun = 15000, double ftol = 2.2204460492503131e-09,
                double pgtol = 1e-5, int maxls = 20) {
    const double epsmch = 2.220446049250313e-16, stpmx = 1e10;
    Result res{0, 0, 0, 0.0};
    std::vector<double> g(n), d(n), t(n), r(n), q(n), S((size_t)m * n), Y((size_t)m * n), rho(m), alpha(m);
    double f = 0.0;
    if (fg(x, &f, g.data())) { res.status = 2; return res; }
    res.nfev = 1;
    int col = 0, head = 0;  // pairs live at (head + i) % m, i = 0 (oldest) .. col - 1 (newest)
    double theta = 1.0;
    auto inf_norm = [&](const std::vector<double>& v) {
        double a = 0.0;
        for (int i = 0; i < n; ++i) a = fmax(a, fabs(v[i]));
        return a;
    };
    if (inf_norm(g) <= pgtol) { res.f = f; return res; }
    for (;;) {
        // ---- direction d = -H g
        if (col == 0) {
            for (int i = 0; i < n; ++i) d[i] = -g[i] / theta;
        } else {
            for (int i = 0; i < n; ++i) q[i] = g[i];
            for (int k = col - 1; k >= 0; --k) {
                const int p = (head + k) % m;
                alpha[k] = rho[p] * dot(&S[(size_t)p * n], q.data(), n);
                const double* y = &Y[(size_t)p * n];
                for (int i = 0; i < n; ++i) q[i] -= alpha[k] * y[i];
            }
            for (int i = 0; i < n; ++i) q[i] /= theta;
            for (int k = 0; k < col; ++k) {
                const int p = (head + k) % m;
                const double beta = rho[p] * dot(&Y[(size_t)p * n], q.data(), n);
                const double* s = &S[(size_t)p * n];
                for (int i = 0; i < n; ++i) q[i] += s[i] * (alpha[k] - beta);
            }
            for (int i = 0; i < n; ++i) d[i] = -q[i];
        }
        // ---- line search (lnsrlb)
        const double dnorm = sqrt(dot(d.data(), d.data(), n));
        double stp = (res.nit == 0) ? fmin(1.0 / dnorm, stpmx) : 1.0;
        for (int i = 0; i < n; ++i) { t[i] = x[i]; r[i] = g[i]; }
        const double fold = f;
        double gd = 0.0, gdold = 0.0;
        int ifun = 0;
        Task task = TASK_START;
        LineSearchState ls;
        bool bad = false;
        for (;;) {
            gd = dot(g.data(), d.data(), n);
            if (ifun == 0) {
                gdold = gd;
                if (gd >= 0.0) { bad = true; break; }  // not a descent direction
            }
            dcsrch(f, gd, stp, 1e-3, 0.9, 0.1, 0.0, stpmx, task, ls);
            if (task == TASK_CONVERGENCE || task == TASK_WARNING) break;
            if (task == TASK_ERROR) { bad = true; break; }
            if (ifun >= maxls) { bad = true; break; }
            ++ifun;
            if (stp == 1.0) for (int i = 0; i < n; ++i) x[i] = t[i] + d[i];
            else for (int i = 0; i < n; ++i) x[i] = stp * d[i] + t[i];
            if (fg(x, &f, g.data())) { bad = true; break; }
            ++res.nfev;
        }
        if (bad) {
            // L-BFGS-B restarts from the last good point with an empty memory; a second failure ends the run
            for (int i = 0; i < n; ++i) { x[i] = t[i]; g[i] = r[i]; }
            f = fold;
            if (col == 0) { res.status = 2; break; }
            col = 0;
            head = 0;
            theta = 1.0;
            continue;
        }
        ++res.nit;
        // ---- stopping tests
        if (inf_norm(g) <= pgtol) break;
        const double ddum0 = fmax(fmax(fabs(fold), fabs(f)), 1.0);
        if (fold - f <= ftol * ddum0) break;
        if (res.nit >= maxiter || res.nfev >= maxfun) { res.status = 1; break; }
        // ---- BFGS pair: s = stp d, y = g - g_old
        for (int i = 0; i < n; ++i) r[i] = g[i] - r[i];
        double dr, ddum;
        if (stp == 1.0) {
            dr = gd - gdold;
            ddum = -gdold;
        } else {
            dr = (gd - gdold) * stp;
            for (int i = 0; i < n; ++i) d[i] *= stp;
            ddum = -gdold * stp;
        }
        const double rr = dot(r.data(), r.data(), n);
        if (dr <= epsmch * ddum) continue;  // skip the update, keep the memory
        int p;
        if (col < m) {
            p = (head + col) % m;
            ++col;
        } else {
            p = head;
            head = (head + 1) % m;
        }
        memcpy(&S[(size_t)p * n], d.data(), sizeof(double) * n);
        memcpy(&Y[(size_t)p * n], r.data(), sizeof(double) * n);
        rho[p] = 1.0 / dr;
        theta = rr / dr;
    }
    res.f = f;
    return res;
}

}  // namespace rbl_lbfgs
