/*
 * CPU oracle (C twin of oracle/gml_oracle.py) for the learn() hot path of
 * lanl-ansi/GraphicalModelLearning.jl v0.2.2.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker and the timed CPU baseline.  The product
 * (graphicalmodellearning.jl_b200/) never links, loads or calls it.
 *
 * It restates, in float64, what the reference does per node (all paths relative to
 * /root/reference/):
 *   - nodal statistics               src/GraphicalModelLearning.jl:162   (pairwise)
 *                                    src/GraphicalModelLearning.jl:94-108 (multiRISE keys/products)
 *   - objectives RISE/logRISE/RPLE   src/GraphicalModelLearning.jl:169-172 / 278-281 / 316-319
 *   - row placement + symmetrise     src/GraphicalModelLearning.jl:181-186
 * and solves each node's convex problem with a second-order method that, like the reference's
 * Ipopt (exact AD Hessians, one dense Newton system per iteration), evaluates f, grad f and the
 * dense Hessian over the whole histogram every iteration:
 *   mode 0 "exact"   proximal Newton + coordinate descent  -> exact L1 minimiser
 *   mode 1 "barrier" damped Newton on f + sum phi_mu       -> the log-barrier point Ipopt returns
 *                    (phi_mu = slack-eliminated barrier term, see gml_oracle.py docstring)
 * Pinned against the reference's 12 stored known-answer matrices by tests/test_oracle_golden.py.
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { FORM_RISE = 0, FORM_LOGRISE = 1, FORM_RPLE = 2 };

typedef struct {
    int64_t K;
    int F;
    const int8_t* stat; /* K x F row-major, entries +-1 */
    const double* w;    /* K weights c_k / M */
    int form;
} node_problem;

/* f only.  t_k = sum_f x_f stat[k,f] */
static double eval_f(const node_problem* p, const double* x) {
    const int F = p->F;
    double acc = 0.0;
    if (p->form == FORM_LOGRISE) {
        double tmax = -INFINITY;
        for (int64_t k = 0; k < p->K; ++k) {
            const int8_t* s = p->stat + k * F;
            double t = 0.0;
            for (int f = 0; f < F; ++f) t += x[f] * s[f];
            if (-t > tmax) tmax = -t;
        }
        for (int64_t k = 0; k < p->K; ++k) {
            const int8_t* s = p->stat + k * F;
            double t = 0.0;
            for (int f = 0; f < F; ++f) t += x[f] * s[f];
            acc += p->w[k] * exp(-t - tmax);
        }
        return log(acc) + tmax;
    }
    for (int64_t k = 0; k < p->K; ++k) {
        const int8_t* s = p->stat + k * F;
        double t = 0.0;
        for (int f = 0; f < F; ++f) t += x[f] * s[f];
        if (p->form == FORM_RISE) {
            acc += p->w[k] * exp(-t);
        } else {
            double a = -2.0 * t;
            acc += p->w[k] * ((a > 0 ? a : 0.0) + log1p(exp(-fabs(a))));
        }
    }
    return acc;
}

/* Threads used INSIDE one node's f/g/H accumulation (split over histogram rows).  The node loops of the
 * learn_* entry points set it to (host threads) / (nodes solved concurrently), so that a test which asks for a
 * few nodes of a large problem still uses the whole box.  Partial sums are combined in thread order. */
static int g_inner_threads = 1;

static double eval_fgh_rows(const node_problem* p, const double* x, double* g, double* H, double* row,
                            int64_t k0, int64_t k1, double tmax);

/* f, g (F), H (F x F, full symmetric).  `row` is scratch of F doubles. */
static double eval_fgh(const node_problem* p, const double* x, double* g, double* H, double* row) {
    const int F = p->F;
#ifdef _OPENMP
    if (g_inner_threads > 1 && p->K >= 8192) {
        const int T = g_inner_threads;
        double tmax = 0.0;
        if (p->form == FORM_LOGRISE) {
            tmax = -INFINITY;
#pragma omp parallel for num_threads(T) reduction(max : tmax)
            for (int64_t k = 0; k < p->K; ++k) {
                const int8_t* s = p->stat + k * F;
                double t = 0.0;
                for (int f = 0; f < F; ++f) t += x[f] * s[f];
                if (-t > tmax) tmax = -t;
            }
        }
        double* gs = malloc(sizeof(double) * (size_t)T * F);
        double* Hs = malloc(sizeof(double) * (size_t)T * F * F);
        double* accs = malloc(sizeof(double) * T);
        int granted = 1;                      /* the runtime may grant fewer threads than asked for (nested regions) */
#pragma omp parallel num_threads(T)
        {
            const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#pragma omp single
            granted = nt;
            double* rw = malloc(sizeof(double) * F);
            const int64_t k0 = p->K * t / nt, k1 = p->K * (t + 1) / nt;
            accs[t] = eval_fgh_rows(p, x, gs + (size_t)t * F, Hs + (size_t)t * F * F, rw, k0, k1, tmax);
            free(rw);
        }
        double acc = 0.0;
        memset(g, 0, sizeof(double) * F);
        memset(H, 0, sizeof(double) * (size_t)F * F);
        for (int t = 0; t < granted; ++t) {
            acc += accs[t];
            for (int f = 0; f < F; ++f) g[f] += gs[(size_t)t * F + f];
            for (size_t i = 0; i < (size_t)F * F; ++i) H[i] += Hs[(size_t)t * F * F + i];
        }
        free(gs); free(Hs); free(accs);
        double fval = acc;
        if (p->form == FORM_LOGRISE) {
            const double Z = acc;
            for (int f = 0; f < F; ++f) g[f] /= Z;
            for (int a = 0; a < F; ++a)
                for (int b = 0; b <= a; ++b) H[(size_t)a * F + b] = H[(size_t)a * F + b] / Z - g[a] * g[b];
            fval = log(Z) + tmax;
        }
        for (int a = 0; a < F; ++a)
            for (int b = 0; b < a; ++b) H[(size_t)b * F + a] = H[(size_t)a * F + b];
        (void)row;
        return fval;
    }
#endif
    memset(g, 0, sizeof(double) * F);
    memset(H, 0, sizeof(double) * (size_t)F * F);
    double acc = 0.0, tmax = 0.0;
    if (p->form == FORM_LOGRISE) {
        tmax = -INFINITY;
        for (int64_t k = 0; k < p->K; ++k) {
            const int8_t* s = p->stat + k * F;
            double t = 0.0;
            for (int f = 0; f < F; ++f) t += x[f] * s[f];
            if (-t > tmax) tmax = -t;
        }
    }
    for (int64_t k = 0; k < p->K; ++k) {
        const int8_t* s = p->stat + k * F;
        double t = 0.0;
        for (int f = 0; f < F; ++f) t += x[f] * s[f];
        double gw, hw; /* d f/d t contribution (negated) and curvature weight */
        if (p->form == FORM_RISE) {
            double e = p->w[k] * exp(-t);
            acc += e; gw = e; hw = e;
        } else if (p->form == FORM_LOGRISE) {
            double e = p->w[k] * exp(-t - tmax);
            acc += e; gw = e; hw = e;
        } else {
            double a = -2.0 * t;
            acc += p->w[k] * ((a > 0 ? a : 0.0) + log1p(exp(-fabs(a))));
            double sig = 0.5 * (1.0 - tanh(t));
            gw = 2.0 * p->w[k] * sig;
            hw = 4.0 * p->w[k] * sig * (1.0 - sig);
        }
        for (int f = 0; f < F; ++f) { g[f] -= gw * s[f]; row[f] = hw * s[f]; }
        for (int a = 0; a < F; ++a) {
            double* Ha = H + (size_t)a * F;
            const double sa = (double)s[a];
            for (int b = 0; b <= a; ++b) Ha[b] += sa * row[b];
        }
    }
    double fval = acc;
    if (p->form == FORM_LOGRISE) {
        const double Z = acc;
        for (int f = 0; f < F; ++f) g[f] /= Z;
        for (int a = 0; a < F; ++a)
            for (int b = 0; b <= a; ++b) H[(size_t)a * F + b] = H[(size_t)a * F + b] / Z - g[a] * g[b];
        fval = log(Z) + tmax;
    }
    for (int a = 0; a < F; ++a)
        for (int b = 0; b < a; ++b) H[(size_t)b * F + a] = H[(size_t)a * F + b];
    return fval;
}

/* un-normalised sums of eval_fgh over the histogram rows [k0, k1) (lower triangle of H only) */
static double eval_fgh_rows(const node_problem* p, const double* x, double* g, double* H, double* row,
                            int64_t k0, int64_t k1, double tmax) {
    const int F = p->F;
    memset(g, 0, sizeof(double) * F);
    memset(H, 0, sizeof(double) * (size_t)F * F);
    double acc = 0.0;
    for (int64_t k = k0; k < k1; ++k) {
        const int8_t* s = p->stat + k * F;
        double t = 0.0;
        for (int f = 0; f < F; ++f) t += x[f] * s[f];
        double gw, hw;
        if (p->form == FORM_RISE) {
            double e = p->w[k] * exp(-t);
            acc += e; gw = e; hw = e;
        } else if (p->form == FORM_LOGRISE) {
            double e = p->w[k] * exp(-t - tmax);
            acc += e; gw = e; hw = e;
        } else {
            double a = -2.0 * t;
            acc += p->w[k] * ((a > 0 ? a : 0.0) + log1p(exp(-fabs(a))));
            double sig = 0.5 * (1.0 - tanh(t));
            gw = 2.0 * p->w[k] * sig;
            hw = 4.0 * p->w[k] * sig * (1.0 - sig);
        }
        for (int f = 0; f < F; ++f) { g[f] -= gw * s[f]; row[f] = hw * s[f]; }
        for (int a = 0; a < F; ++a) {
            double* Ha = H + (size_t)a * F;
            const double sa = (double)s[a];
            for (int b = 0; b <= a; ++b) Ha[b] += sa * row[b];
        }
    }
    return acc;
}

static double l1_pen(const double* x, const uint8_t* pen, int F) {
    double s = 0.0;
    for (int f = 0; f < F; ++f) if (pen[f]) s += fabs(x[f]);
    return s;
}

/* exact L1 minimiser: proximal Newton with cyclic coordinate descent */
static int solve_exact(const node_problem* p, const uint8_t* pen, double lam, double* x,
                       double tol, int max_outer, int64_t* n_fgh, int64_t* n_f) {
    const int F = p->F;
    double* g = malloc(sizeof(double) * F);
    double* H = malloc(sizeof(double) * (size_t)F * F);
    double* row = malloc(sizeof(double) * F);
    double* d = malloc(sizeof(double) * F);
    double* Hd = malloc(sizeof(double) * F);
    double* xn = malloc(sizeof(double) * F);
    int it = 0;
    double F_prev = INFINITY, step_prev = INFINITY;
    int idle = 0;
    for (; it < max_outer; ++it) {
        double f = eval_fgh(p, x, g, H, row); ++*n_fgh;
        double Fx = f + lam * l1_pen(x, pen, F);
        /* A coordinate that sits on the activation threshold to rounding keeps flipping between 0 and ~1e-11: the step
         * never drops below tol although the objective has stopped moving.  Stop there (the point is converged to the
         * precision float64 sums over K rows can resolve) instead of spending all max_outer Hessian sweeps. */
        idle = (it > 1 && step_prev < 1e-9 && fabs(F_prev - Fx) <= 1e-13 * fmax(fabs(Fx), 1.0)) ? idle + 1 : 0;
        if (idle >= 4) break;
        F_prev = Fx;
        memset(d, 0, sizeof(double) * F);
        memset(Hd, 0, sizeof(double) * F);
        for (int sweep = 0; sweep < 10000; ++sweep) {
            double maxchg = 0.0;
            for (int j = 0; j < F; ++j) {
                double a = H[(size_t)j * F + j];
                if (a < 1e-300) a = 1e-300;
                double cur = x[j] + d[j];
                double v = cur - (g[j] + Hd[j]) / a;
                if (pen[j]) {
                    double thr = lam / a, av = fabs(v) - thr;
                    v = av > 0 ? copysign(av, v) : 0.0;
                }
                double delta = v - cur;
                if (delta != 0.0) {
                    d[j] += delta;
                    const double* Hj = H + (size_t)j * F;
                    for (int i = 0; i < F; ++i) Hd[i] += Hj[i] * delta;
                    if (fabs(delta) > maxchg) maxchg = fabs(delta);
                }
            }
            if (maxchg < 1e-16 + 1e-3 * tol) break;
        }
        double step = 0.0;
        for (int j = 0; j < F; ++j) if (fabs(d[j]) > step) step = fabs(d[j]);
        step_prev = step;
        if (step < tol) { for (int j = 0; j < F; ++j) x[j] += d[j]; ++it; break; }
        double gd = 0.0;
        for (int j = 0; j < F; ++j) { gd += g[j] * d[j]; xn[j] = x[j] + d[j]; }
        double delta_model = gd + lam * (l1_pen(xn, pen, F) - l1_pen(x, pen, F));
        double alpha = 1.0;
        for (;;) {
            for (int j = 0; j < F; ++j) xn[j] = x[j] + alpha * d[j];
            double Fn = eval_f(p, xn) + lam * l1_pen(xn, pen, F); ++*n_f;
            /* slack: float64 summation noise of eval_f over K terms */
            if (Fn <= Fx + 1e-4 * alpha * delta_model + 1e-13 * fmax(fabs(Fx), 1.0) || alpha < 1e-10) break;
            alpha *= 0.5;
        }
        memcpy(x, xn, sizeof(double) * F);
    }
    free(g); free(H); free(row); free(d); free(Hd); free(xn);
    return it;
}

/* in-place Cholesky solve of A y = b (A symmetric positive definite, F x F); returns 0 on success */
static int chol_solve(double* A, double* b, int F) {
    for (int j = 0; j < F; ++j) {
        double s = A[(size_t)j * F + j];
        for (int k = 0; k < j; ++k) s -= A[(size_t)j * F + k] * A[(size_t)j * F + k];
        if (!(s > 0.0)) return 1;
        double l = sqrt(s);
        A[(size_t)j * F + j] = l;
        for (int i = j + 1; i < F; ++i) {
            double t = A[(size_t)i * F + j];
            for (int k = 0; k < j; ++k) t -= A[(size_t)i * F + k] * A[(size_t)j * F + k];
            A[(size_t)i * F + j] = t / l;
        }
    }
    for (int i = 0; i < F; ++i) {
        double t = b[i];
        for (int k = 0; k < i; ++k) t -= A[(size_t)i * F + k] * b[k];
        b[i] = t / A[(size_t)i * F + i];
    }
    for (int i = F - 1; i >= 0; --i) {
        double t = b[i];
        for (int k = i + 1; k < F; ++k) t -= A[(size_t)k * F + i] * b[k];
        b[i] = t / A[(size_t)i * F + i];
    }
    return 0;
}

static double barrier_sum(const double* x, const uint8_t* pen, int F, double lam, double mu) {
    const double eps = mu / lam;
    double s = 0.0;
    for (int f = 0; f < F; ++f) if (pen[f]) {
        double r = sqrt(eps * eps + x[f] * x[f]), z = eps + r;
        s += lam * z - mu * log(2.0 * eps * z);
    }
    return s;
}

/* damped Newton on f(x) + sum_pen phi_mu(x_j) from the warm start x */
static int solve_barrier(const node_problem* p, const uint8_t* pen, double lam, double mu, double* x,
                         double tol, int max_iter, int64_t* n_fgh, int64_t* n_f) {
    const int F = p->F;
    const double eps = mu / lam;
    double* g = malloc(sizeof(double) * F);
    double* H = malloc(sizeof(double) * (size_t)F * F);
    double* row = malloc(sizeof(double) * F);
    double* d = malloc(sizeof(double) * F);
    double* xn = malloc(sizeof(double) * F);
    int it = 0;
    for (; it < max_iter; ++it) {
        double f = eval_fgh(p, x, g, H, row); ++*n_fgh;
        for (int j = 0; j < F; ++j) if (pen[j]) {
            double r = sqrt(eps * eps + x[j] * x[j]), z = eps + r;
            g[j] += lam * x[j] / z;
            H[(size_t)j * F + j] += lam * (z - x[j] * x[j] / r) / (z * z);
        }
        double m0 = f + barrier_sum(x, pen, F, lam, mu);
        for (int j = 0; j < F; ++j) d[j] = -g[j];
        if (chol_solve(H, d, F)) break; /* not PD: keep current point */
        double slope = 0.0;
        for (int j = 0; j < F; ++j) slope += g[j] * d[j];
        double alpha = 1.0;
        for (;;) {
            for (int j = 0; j < F; ++j) xn[j] = x[j] + alpha * d[j];
            double m = eval_f(p, xn) + barrier_sum(xn, pen, F, lam, mu); ++*n_f;
            if (m <= m0 + 1e-4 * alpha * slope + 1e-13 * fmax(fabs(m0), 1.0) || alpha < 1e-12) break;
            alpha *= 0.5;
        }
        double step = 0.0;
        for (int j = 0; j < F; ++j) { if (fabs(alpha * d[j]) > step) step = fabs(alpha * d[j]); x[j] = xn[j]; }
        if (step < tol) { ++it; break; }
    }
    free(g); free(H); free(row); free(d); free(xn);
    return it;
}

/* Generic single-node solve.  stat is K x F row-major int8 (+-1).  Returns iterations. */
int gml_oracle_solve_node(int form, const int8_t* stat, const double* w, int64_t K, int F,
                          const uint8_t* pen, double lam, int mode, double mu, double* x,
                          double* obj, int64_t* n_fgh, int64_t* n_f) {
    node_problem p = { K, F, stat, w, form };
    int64_t a = 0, b = 0;
    memset(x, 0, sizeof(double) * F);
    int it = solve_exact(&p, pen, lam, x, 1e-13, 200, &a, &b);
    if (mode == 1 && mu > 0.0 && lam > 0.0) {
        double* g = malloc(sizeof(double) * F);
        double* H = malloc(sizeof(double) * (size_t)F * F);
        double* row = malloc(sizeof(double) * F);
        eval_fgh(&p, x, g, H, row); ++a;
        for (int j = 0; j < F; ++j) if (pen[j] && x[j] == 0.0) {
            double den = lam * lam - g[j] * g[j];
            if (den < 1e-300) den = 1e-300;
            x[j] = -2.0 * mu * g[j] / den;
        }
        free(g); free(H); free(row);
        it += solve_barrier(&p, pen, lam, mu, x, 1e-15, 200, &a, &b);
    }
    if (obj) *obj = eval_f(&p, x) + lam * l1_pen(x, pen, F);
    if (n_fgh) *n_fgh = a;
    if (n_f) *n_f = b;
    return it;
}

/*
 * learn(samples, RISE/logRISE/RPLE) for nodes [node_begin, node_end).
 * counts[K] (double), spins: int8 spin-major (spins[i*ld + k] = s_i^k), the layout of a Julia
 * column-major samples[:,2:end] converted to Int8.  out: N x N row-major (row u = node u).
 * When the node range is the full range and symmetrize != 0, out is symmetrised (:184-186).
 * stats[0] = sum of f/g/H evaluations, stats[1] = sum of f-only evaluations (over nodes).
 */
/* host threads = (nodes solved concurrently) x (threads inside a node's accumulation) */
static int split_threads(int n_nodes) {
#ifdef _OPENMP
    const int T = omp_get_max_threads();
    int outer = n_nodes < T ? (n_nodes > 0 ? n_nodes : 1) : T;
    g_inner_threads = T / outer > 1 ? T / outer : 1;
    omp_set_max_active_levels(2);
    return outer;
#else
    (void)n_nodes;
    return 1;
#endif
}

int gml_oracle_learn_pairwise(const double* counts, const int8_t* spins, int64_t K, int N, int64_t ld,
                              int form, double lam, int symmetrize, int mode, double mu,
                              int node_begin, int node_end, double* out, double* obj, int64_t* stats) {
    double M = 0.0;
    for (int64_t k = 0; k < K; ++k) M += counts[k];
    double* w = malloc(sizeof(double) * K);
    for (int64_t k = 0; k < K; ++k) w[k] = counts[k] / M;
    int64_t tot_fgh = 0, tot_f = 0;
    const int outer = split_threads(node_end - node_begin);
    (void)outer;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tot_fgh, tot_f) num_threads(outer)
    for (int u = node_begin; u < node_end; ++u) {
        int8_t* stat = malloc((size_t)K * N);
        uint8_t* pen = malloc(N);
        double* x = malloc(sizeof(double) * N);
        for (int i = 0; i < N; ++i) {
            pen[i] = (i != u);
            for (int64_t k = 0; k < K; ++k)
                stat[k * N + i] = (i == u) ? spins[(int64_t)u * ld + k]
                                           : (int8_t)(spins[(int64_t)u * ld + k] * spins[(int64_t)i * ld + k]);
        }
        double o = 0.0; int64_t a = 0, b = 0;
        gml_oracle_solve_node(form, stat, w, K, N, pen, lam, mode, mu, x, &o, &a, &b);
        for (int i = 0; i < N; ++i) out[(size_t)u * N + i] = x[i];
        if (obj) obj[u] = o;
        tot_fgh += a; tot_f += b;
        free(stat); free(pen); free(x);
    }
    if (symmetrize && node_begin == 0 && node_end == N) {
        for (int i = 0; i < N; ++i)
            for (int j = i + 1; j < N; ++j) {
                double m = 0.5 * (out[(size_t)i * N + j] + out[(size_t)j * N + i]);
                out[(size_t)i * N + j] = m; out[(size_t)j * N + i] = m;
            }
    }
    if (stats) { stats[0] = tot_fgh; stats[1] = tot_f; }
    g_inner_threads = 1;
    free(w);
    return 0;
}

/*
 * learn(samples, multiRISE(c, sym, p)) un-symmetrised core: per node u the features are the
 * products over key members; keys are passed in as a flat table key_idx[n_keys * order]
 * (0-based spin ids, -1 padded), key_len[n_keys], listed in the reference's own order
 * ((u,), (u,j)..., (u,j<k)...).  out_vals[u * n_keys + f].  L1 on keys of length > 1 (:118).
 */
int gml_oracle_learn_multibody_nodes(const double* counts, const int8_t* spins, int64_t K, int N, int64_t ld,
                                     int order, int n_keys, const int32_t* key_idx, const int32_t* key_len,
                                     double lam, int mode, double mu, int node_begin, int node_end,
                                     double* out_vals, double* obj) {
    double M = 0.0;
    for (int64_t k = 0; k < K; ++k) M += counts[k];
    double* w = malloc(sizeof(double) * K);
    for (int64_t k = 0; k < K; ++k) w[k] = counts[k] / M;
    (void)N;
    const int outer = split_threads(node_end - node_begin);
    (void)outer;
#pragma omp parallel for schedule(dynamic, 1) num_threads(outer)
    for (int u = node_begin; u < node_end; ++u) {
        int8_t* stat = malloc((size_t)K * n_keys);
        uint8_t* pen = malloc(n_keys);
        double* x = malloc(sizeof(double) * n_keys);
        const int32_t* kidx = key_idx + (size_t)u * n_keys * order;
        const int32_t* klen = key_len + (size_t)u * n_keys;
        for (int f = 0; f < n_keys; ++f) {
            pen[f] = klen[f] > 1;
            for (int64_t k = 0; k < K; ++k) {
                int8_t pr = 1;
                for (int q = 0; q < klen[f]; ++q) pr = (int8_t)(pr * spins[(int64_t)kidx[f * order + q] * ld + k]);
                stat[k * n_keys + f] = pr;
            }
        }
        double o = 0.0;
        gml_oracle_solve_node(FORM_RISE, stat, w, K, n_keys, pen, lam, mode, mu, x, &o, 0, 0);
        for (int f = 0; f < n_keys; ++f) out_vals[(size_t)u * n_keys + f] = x[f];
        if (obj) obj[u] = o;
        free(stat); free(pen); free(x);
    }
    g_inner_threads = 1;
    free(w);
    return 0;
}

int gml_oracle_learn_multibody(const double* counts, const int8_t* spins, int64_t K, int N, int64_t ld,
                               int order, int n_keys, const int32_t* key_idx, const int32_t* key_len,
                               double lam, int mode, double mu, double* out_vals, double* obj) {
    return gml_oracle_learn_multibody_nodes(counts, spins, K, N, ld, order, n_keys, key_idx, key_len, lam, mode, mu,
                                            0, N, out_vals, obj);
}

/*
 * Objective and gradient of the smooth part f_u at caller-supplied points, WITHOUT solving: the float64 statement of
 *   RISE    f_u = sum_k w_k exp(-t_k)            src/GraphicalModelLearning.jl:170   (gradient: :199-208)
 *   logRISE f_u = log sum_k w_k exp(-t_k)        src/GraphicalModelLearning.jl:279
 *   RPLE    f_u = sum_k w_k log(1+exp(-2 t_k))   src/GraphicalModelLearning.jl:317
 * with t_k = s_u^k (sum_{i != u} x_i s_i^k + x_field)  (nodal_stat of :162 folded in), w_k = c_k / M (:79).
 * nodes[n_nodes]: 0-based node ids; x, g_out: n_nodes x (N+1) row-major in the order of the C ABI's
 * gml_b200_eval_pairwise (couplings to spins 0..N-1, the self entry ignored / returned as 0, then the field).
 * Used by the parity tests at sizes where a per-node K x N float64 nodal_stat matrix is too slow, and by bench.py's
 * KKT check of the solution it times.  The histogram is swept once per node in cache-sized row chunks, threads
 * split the rows, partial sums are combined in thread order (deterministic for a fixed thread count).
 */
int gml_oracle_eval_pairwise(const double* counts, const int8_t* spins, int64_t K, int N, int64_t ld, int form,
                             const int32_t* nodes, int n_nodes, const double* x, double* f_out, double* g_out) {
    double M = 0.0;
    for (int64_t k = 0; k < K; ++k) M += counts[k];
    int T = 1;
#ifdef _OPENMP
    T = omp_get_max_threads();
#endif
    const int F = N + 1;
    enum { CH = 2048 };
    double* gpart = malloc(sizeof(double) * (size_t)T * F);
    double* fpart = malloc(sizeof(double) * T);
    double* mpart = malloc(sizeof(double) * T);
    for (int q = 0; q < n_nodes; ++q) {
        const int u = nodes[q];
        const double* xu = x + (size_t)q * F;
        const int8_t* su = spins + (int64_t)u * ld;
        /* logRISE: shift by the largest exponent (same value of f, no overflow) */
        double tmax = 0.0;
        for (int t = 0; t < T; ++t) { mpart[t] = -INFINITY; fpart[t] = 0.0; }
        memset(gpart, 0, sizeof(double) * (size_t)T * F);
        if (form == FORM_LOGRISE) {
#pragma omp parallel num_threads(T)
            {
                int t = 0;
#ifdef _OPENMP
                t = omp_get_thread_num();
#endif
                double* e = malloc(sizeof(double) * CH);
                double mx = -INFINITY;
                int nt = 1;
#ifdef _OPENMP
                nt = omp_get_num_threads();
#endif
                const int64_t k0 = K * t / nt, k1 = K * (t + 1) / nt;
                for (int64_t c0 = k0; c0 < k1; c0 += CH) {
                    const int n = (int)((k1 - c0) < CH ? (k1 - c0) : CH);
                    for (int k = 0; k < n; ++k) e[k] = xu[N];
                    for (int i = 0; i < N; ++i) {
                        if (i == u || xu[i] == 0.0) continue;
                        const int8_t* si = spins + (int64_t)i * ld + c0;
                        const double xi = xu[i];
                        for (int k = 0; k < n; ++k) e[k] += xi * si[k];
                    }
                    for (int k = 0; k < n; ++k) { const double t_k = su[c0 + k] * e[k]; if (-t_k > mx) mx = -t_k; }
                }
                mpart[t] = mx;
                free(e);
            }
            tmax = -INFINITY;
            for (int t = 0; t < T; ++t) if (mpart[t] > tmax) tmax = mpart[t];     /* slots of threads not granted hold -inf */
        }
#pragma omp parallel num_threads(T)
        {
            int t = 0;
#ifdef _OPENMP
            t = omp_get_thread_num();
#endif
            double* e = malloc(sizeof(double) * CH);
            double* r = malloc(sizeof(double) * CH);
            double* g = gpart + (size_t)t * F;
            memset(g, 0, sizeof(double) * F);
            double acc = 0.0;
            int nt = 1;
#ifdef _OPENMP
            nt = omp_get_num_threads();
#endif
            const int64_t k0 = K * t / nt, k1 = K * (t + 1) / nt;
            for (int64_t c0 = k0; c0 < k1; c0 += CH) {
                const int n = (int)((k1 - c0) < CH ? (k1 - c0) : CH);
                for (int k = 0; k < n; ++k) e[k] = xu[N];
                for (int i = 0; i < N; ++i) {
                    if (i == u || xu[i] == 0.0) continue;
                    const int8_t* si = spins + (int64_t)i * ld + c0;
                    const double xi = xu[i];
                    for (int k = 0; k < n; ++k) e[k] += xi * si[k];
                }
                for (int k = 0; k < n; ++k) {
                    const double s = su[c0 + k], t_k = s * e[k], w = counts[c0 + k] / M;
                    double gw;
                    if (form == FORM_RISE) { gw = w * exp(-t_k); acc += gw; }
                    else if (form == FORM_LOGRISE) { gw = w * exp(-t_k - tmax); acc += gw; }
                    else {
                        const double a = -2.0 * t_k;
                        acc += w * ((a > 0 ? a : 0.0) + log1p(exp(-fabs(a))));
                        gw = 2.0 * w * 0.5 * (1.0 - tanh(t_k));
                    }
                    r[k] = -gw * s;                       /* d f / d (sum_i x_i s_i + h) for this row */
                }
                if (g_out) {
                    for (int i = 0; i < N; ++i) {
                        const int8_t* si = spins + (int64_t)i * ld + c0;
                        double d = 0.0;
                        for (int k = 0; k < n; ++k) d += r[k] * si[k];
                        g[i] += d;
                    }
                    double d = 0.0;
                    for (int k = 0; k < n; ++k) d += r[k];
                    g[N] += d;
                }
            }
            fpart[t] = acc;
            free(e); free(r);
        }
        double acc = 0.0;
        for (int t = 0; t < T; ++t) acc += fpart[t];
        f_out[q] = (form == FORM_LOGRISE) ? log(acc) + tmax : acc;
        if (g_out) {
            double* go = g_out + (size_t)q * F;
            for (int f = 0; f < F; ++f) {
                double v = 0.0;
                for (int t = 0; t < T; ++t) v += gpart[(size_t)t * F + f];
                go[f] = (form == FORM_LOGRISE) ? v / acc : v;
            }
            go[u] = 0.0;
        }
    }
    free(gpart); free(fpart); free(mpart);
    return 0;
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline asks for the host cores explicitly */
void gml_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int gml_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
