// K4: batched FISTA driver.  All Nn node problems advance together; every node carries its own
// Lipschitz estimate L_u, momentum t_u, restart and convergence state, so there is no host decision
// inside a round and every round costs exactly ONE objective+gradient pass over the histogram:
//
//     trial    Z_u  = prox_{lambda/L_u}(Y_u - G_u/L_u)          (soft-threshold on PEN_L1 coordinates)
//              Y'_u = Z_u + beta_u (Z_u - X_u)                   (momentum; beta = 0 on a gradient restart)
//     pass     f(Y'), G(Y')                                     (energy + gradient contractions)
//     accept   descent test at the NEW point against the model built at the old one,
//                  f(Y') <= f(Y) + <G, Y'-Y> + L/2 |Y'-Y|^2 ,
//              which must hold for any pair of points once L bounds the curvature.  Passed: X <- Z,
//              (Y, G, f) <- (Y', G', f').  Failed: L_u *= 2 and the node retries from the same (Y, G).
//
// Nodes drop out of the passes as they finish: on the coarse precision level of a lattice backend a node that has
// reached the tolerance or the resolution of the coarse lattice parks (status 3) until all have, on the fine level a
// converged node retires (status 1); the passes are then restricted to an ordered list of the active nodes
// (fista_compact_kernel -> EvalBackend::set_active), so late rounds cost what their active nodes cost.
//
// Testing the lemma at Y' instead of at Z saves the separate objective-only pass of textbook
// backtracking FISTA; it probes the curvature along the direction the iterate actually moves.
// The problem solved per node is the reference's (src/GraphicalModelLearning.jl:169-177 etc.) with
// the slack variables z eliminated: min f_u(x) + lambda*sum_{pen} |x_j|.
#include "common.cuh"

#include <chrono>
#include <cmath>
#include <memory>

namespace gml {
namespace {

struct FistaState {
    int Nn, Fp, form;
    double lambda, tol, lattice_inv, lattice, eps_f, eps_g;
    double x_range;   // lattice backends: largest representable |x| at the current precision level (0 = unbounded)
    const uint8_t* pen;
    double *X, *Z, *Y, *Yn, *G, *Gn;
    double *fY, *fYn, *L, *t, *tn, *q, *c, *gmap, *obj;
    double* gtrue;    // gradient-mapping norm of the UNSNAPPED prox point (what the stopping rule is about; gmap is the snapped one)
    int coarse_retire;   // 1: nodes may retire on the coarse level (its gradient noise is far below tol)
    double* best;     // smallest gradient-mapping norm seen
    int *stall, *streak;
    int* status;      // 0 active, 1 converged, 2 stalled at the gradient noise floor, 3 parked (done with the coarse level),
                      // 4 left the fixed-point range of the backend on the fine level (re-solved by the unbounded backend)
    int* clamped;     // [Nn] raised by the trial kernel when a prox point had to be clamped to the backend's range
    double park_tol;  // coarse level: a node whose gradient-mapping norm is below this waits for the fine level
    int* n_active;
    unsigned long long* gmax;   // bits of the largest gradient-mapping norm over the active nodes of this round
    int fine;                   // 1: the backend runs at its fine precision level
};

__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    return s;
}
__device__ __forceinline__ double block_max(double v, double* red) {
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s = fmax(s, red[i]);
    return s;
}

__device__ __forceinline__ double snap(double v, const FistaState& s) {
    // lattice backends hold x as a fixed-point number with a bounded range (|x| < 7.9 on the fine level of the
    // tensor-core backend).  A point that had to be clamped is NOT the prox point: callers record it (snap_flag) and the
    // node is never retired with it.
    return s.lattice_inv > 0.0 ? rint(fmin(fmax(v, -s.x_range), s.x_range) * s.lattice_inv) * s.lattice : v;
}
__device__ __forceinline__ double snap_flag(double v, const FistaState& s, bool& clamped) {
    if (s.lattice_inv > 0.0 && fabs(v) > s.x_range) clamped = true;
    return snap(v, s);
}

// Z = prox step from (Y, G);  Y' = Z + beta (Z - X);  model terms of the descent test at Y'
__global__ void __launch_bounds__(128) fista_trial_kernel(FistaState s) {
    const int u = blockIdx.x;
    if (s.status[u]) return;
    __shared__ double red[4];
    const double L = s.L[u], thr = s.lambda / L;
    const int64_t o = (int64_t)u * s.Fp;
    double dm = 0.0, r = 0.0, dm_true = 0.0;
    bool clamped = false;
    for (int f = threadIdx.x; f < s.Fp; f += blockDim.x) {
        const uint8_t pc = s.pen[o + f];
        const double y = s.Y[o + f], g = s.G[o + f];
        double z = 0.0;
        if (pc != PEN_ZERO) {
            z = y - g / L;
            if (pc == PEN_L1) { const double a = fabs(z) - thr; z = a > 0.0 ? copysign(a, z) : 0.0; }
            dm_true = fmax(dm_true, fabs(z - y));
            z = snap_flag(z, s, clamped);
        }
        s.Z[o + f] = z;
        dm = fmax(dm, fabs(z - y));
        r += (y - z) * (z - s.X[o + f]);          // gradient-scheme adaptive restart: <Y - Z, Z - X> > 0
    }
    dm = block_max(dm, red);
    dm_true = block_max(dm_true, red);
    r = block_sum(r, red);
    const double t = s.t[u];
    const bool restart = r > 0.0;
    const double tn = restart ? 1.0 : 0.5 * (1.0 + sqrt(1.0 + 4.0 * t * t));
    const double beta = restart ? 0.0 : (t - 1.0) / tn;
    double q1 = 0.0, q2 = 0.0;
    for (int f = threadIdx.x; f < s.Fp; f += blockDim.x) {
        const double z = s.Z[o + f], y = s.Y[o + f];
        const double yn = snap_flag(z + beta * (z - s.X[o + f]), s, clamped);
        s.Yn[o + f] = yn;
        const double d = yn - y;
        q1 += s.G[o + f] * d; q2 += d * d;
    }
    q1 = block_sum(q1, red); q2 = block_sum(q2, red);
    if (__syncthreads_or(clamped ? 1 : 0) && threadIdx.x == 0) s.clamped[u] = 1;
    if (threadIdx.x == 0) {
        s.q[u] = q1 + 0.5 * L * q2;
        s.c[u] = 0.5 * L * q2;
        s.gmap[u] = L * dm;          // max-norm of the prox-gradient mapping at Y (to the snapped prox point)
        s.gtrue[u] = L * dm_true;    // ... to the exact prox point
        s.tn[u] = tn;
        atomicMax(s.gmax, (unsigned long long)__double_as_longlong(L * dm));   // non-negative doubles order like their bits
    }
}

__global__ void __launch_bounds__(128) fista_accept_kernel(FistaState s) {
    const int u = blockIdx.x;
    // every thread reads the status before any thread of the block may change it
    const int status0 = s.status[u];
    const int was_clamped = s.clamped[u];
    __syncthreads();
    if (status0) return;
    const int64_t o = (int64_t)u * s.Fp;
    const double gm = s.gmap[u];
    // The iterate left the representable range of the lattice backend: its trial points are clamped, not prox points.
    // Coarse level: handled by the driver (range overflow -> fine level).  Fine level: the node leaves the passes with
    // status 4 at its last accepted point and is re-solved by the unbounded CUDA-core backend (solve_fista).
    if (was_clamped && s.fine) {
        if (threadIdx.x == 0) s.status[u] = 4;
        return;
    }
    // ---- convergence is decided at Y (whose gradient is known): the prox point Z is the answer
    // (on a lattice backend a prox step of at most one lattice unit is the finest resolvable fixed point; nodes never
    // retire on the coarse precision level, whose lattice and gradient noise are above the tolerance)
    const bool conv = s.fine && (gm <= s.tol || (s.lattice > 0.0 && gm <= 1.01 * s.L[u] * s.lattice && gm <= 4.0 * s.tol));
    if (conv) {
        for (int f = threadIdx.x; f < s.Fp; f += blockDim.x) s.X[o + f] = s.Z[o + f];
        if (threadIdx.x == 0) s.status[u] = 1;
        return;
    }
    // Coarse level, when its gradient noise is far below the tolerance (16-bit residuals: ~6e-8 at C3): the stopping rule
    // -- prox-gradient mapping at Y <= tol -- is evaluated with the EXACT prox point; a node that meets it is done and
    // returns that point (not its snap to the coarse lattice).  Only nodes that the coarse lattice keeps above tol (the
    // snapped iteration stops within half a lattice unit of its fixed point: a gradient mapping of up to L * 2^-23) go on to
    // the fine level.
    if (!s.fine && s.coarse_retire && s.gtrue[u] <= s.tol && !was_clamped) {
        const double Lu = s.L[u], thr = s.lambda / Lu;
        for (int f = threadIdx.x; f < s.Fp; f += blockDim.x) {
            const uint8_t pc = s.pen[o + f];
            double z = 0.0;
            if (pc != PEN_ZERO) {
                z = s.Y[o + f] - s.G[o + f] / Lu;
                if (pc == PEN_L1) { const double a = fabs(z) - thr; z = a > 0.0 ? copysign(a, z) : 0.0; }
            }
            s.X[o + f] = z;
        }
        if (threadIdx.x == 0) { s.status[u] = 1; s.gmap[u] = s.gtrue[u]; }
        return;
    }
    // Coarse level: nothing retires, but a node that has reached the resolution of the coarse lattice stops taking
    // part in the passes (its state is frozen at the last accepted point) until the fine level resumes it.
    if (!s.fine && (gm <= s.park_tol || (s.lattice > 0.0 && gm <= 1.01 * s.L[u] * s.lattice))) {
        if (threadIdx.x == 0) s.status[u] = 3;
        return;
    }
    const double fY = s.fY[u], fN = s.fYn[u];
    // Upper bound of the evaluation noise of f (fp32 per-sample terms): relative for the RISE/RPLE sums,
    // absolute for logRISE (f = log Z).  D is the slack of the descent test; for a locally quadratic f,
    // D/c = 1 - L_dir/L with L_dir the curvature along Y' - Y.
    const double noise = s.eps_f * (s.form == GML_B200_LOGRISE ? fmax(fabs(fY), 1.0) : fmax(fabs(fY), 1e-300));
    const double c = s.c[u];
    const bool measurable = c > 10.0 * noise;
    // Inside the noise floor of f the same curvature is read from the GRADIENTS, which stay accurate for tiny
    // steps: <G' - G, Y' - Y> = dY^T H dY, so D = c (1 - L_dir / L) with L_dir = <G'-G, dY> / |dY|^2.
    __shared__ double red[4];
    double dot = 0.0;
    for (int f = threadIdx.x; f < s.Fp; f += blockDim.x) dot += (s.Gn[o + f] - s.G[o + f]) * (s.Yn[o + f] - s.Y[o + f]);
    dot = block_sum(dot, red);
    // ... as long as the gradient change L|dY| is well above the gradient noise of the backend
    const bool secant_ok = !measurable && s.L[u] * sqrt(2.0 * c / s.L[u]) > 20.0 * s.eps_g;
    const double D = measurable ? fY + s.q[u] - fN : (secant_ok ? c - 0.5 * dot : 0.0);
    // Reject only a violation that is significant against c (L more than ~10% below the directional
    // curvature) and, for the function-value form, against the noise.
    const bool reject = !isfinite(fN) || !isfinite(dot) || (c > 0.0 && D < -(0.1 * c + (measurable ? noise : 0.0)));
    if (reject) {
        // L doubles (a discrete rule on purpose: a growth factor computed from the measured violation made L -- and with it
        // the whole path -- depend continuously on the last bits of f, whose summation order differs between partitions
        // of the same problem; measured, it also cost 6 more rounds at C3 than doubling)
        if (threadIdx.x == 0) { s.L[u] *= 2.0; s.streak[u] = 0; atomicAdd(s.n_active, 1); }
        return;
    }
    int stall = s.stall[u];
    double best = s.best[u];
    if (gm < 0.9 * best) { best = gm; stall = 0; } else ++stall;
    // the gradient mapping stopped improving: gradient noise floor (on the coarse level: its lattice; park after 12 rounds)
    const bool stalled = stall >= (s.fine ? 200 : 12);
    for (int f = threadIdx.x; f < s.Fp; f += blockDim.x) {
        s.X[o + f] = s.Z[o + f];
        s.Y[o + f] = s.Yn[o + f];
        s.G[o + f] = s.Gn[o + f];
    }
    if (threadIdx.x == 0) {
        s.t[u] = s.tn[u];
        s.fY[u] = fN;
        s.best[u] = best; s.stall[u] = stall;
        // L is relaxed after three consecutive measurable steps that passed with L > 1.33 L_dir, so it
        // tracks the local curvature (which drops along the path for RPLE) within [0.9, 1.33] L_dir
        int streak = s.streak[u];
        streak = (c > 0.0 && D > 0.25 * c) ? streak + 1 : 0;
        if (streak >= 3) { s.L[u] *= 0.85; streak = 0; }
        s.streak[u] = streak;
        if (stalled) s.status[u] = s.fine ? 2 : 3; else atomicAdd(s.n_active, 1);
    }
}

__global__ void fista_init_kernel(FistaState s, double L0) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= s.Nn) return;
    s.L[u] = L0 > 0.0 ? L0 : s.L[u] * -L0;      // first level: initial estimate; later levels: rescale by rho ratio
    s.t[u] = 1.0; s.status[u] = 0; s.gmap[u] = 1e300; s.obj[u] = 0.0;
    s.best[u] = 1e300; s.stall[u] = 0; s.streak[u] = 0; s.clamped[u] = 0;
}

// start point of a warm-started solve -> the lattice / range of the precision level it is first evaluated on
__global__ void fista_snap_kernel(FistaState s, double* __restrict__ x, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = snap(x[i], s);
}

// parked nodes rejoin (fine level): state kept, stall counters reset
__global__ void fista_unpark_kernel(FistaState s) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= s.Nn) return;
    if (s.status[u] == 3) { s.status[u] = 0; s.best[u] = 1e300; s.stall[u] = 0; }
}

// ordered list of the active nodes (single block; the host already knows the count)
__global__ void __launch_bounds__(1024) fista_compact_kernel(FistaState s, int* __restrict__ idx) {
    __shared__ int warp_tot[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int u0 = 0; u0 < s.Nn; u0 += 1024) {
        const int u = u0 + threadIdx.x;
        const bool on = u < s.Nn && s.status[u] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, on);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_tot[w] = __popc(m);
        __syncthreads();
        int off = base;
        for (int i = 0; i < w; ++i) off += warp_tot[i];
        if (on) idx[off + __popc(m & ((1u << lane) - 1))] = u;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < 32; ++i) t += warp_tot[i]; base += t; }
        __syncthreads();
    }
}

// obj_u = f_u(X) + lambda |X_pen|_1
__global__ void __launch_bounds__(128) fista_objective_kernel(FistaState s, const double* __restrict__ fX) {
    const int u = blockIdx.x;
    __shared__ double red[4];
    const int64_t o = (int64_t)u * s.Fp;
    double l1 = 0.0;
    for (int f = threadIdx.x; f < s.Fp; f += blockDim.x)
        if (s.pen[o + f] == PEN_L1) l1 += fabs(s.X[o + f]);
    l1 = block_sum(l1, red);
    if (threadIdx.x == 0) s.obj[u] = fX[u] + s.lambda * l1;
}

}  // namespace

namespace {
void solve_fista_impl(const NodeProblem& prob, const gml_b200_opts& o, int backend, SolveResult& r, std::vector<int>& out_of_range,
                      cudaStream_t st) {
    const int Nn = prob.Nn, Fp = prob.Fp;
    auto tick = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = tick();
    std::unique_ptr<EvalBackend> be(backend == GML_B200_SOLVER_FISTA_TC ? make_backend_tc(prob, st)
                                                                        : make_backend_cc(prob, st));
    be->set_profiling(o.reserved[0] != 0);
    const size_t nx = (size_t)Nn * Fp;
    DevBuf<double> Z, Y, Yn, G, Gn, fY, fYn, L, t, tn, q, c, gmap, best, gtrue;
    DevBuf<int> status, n_active, stall, streak, clamped;
    DevBuf<unsigned long long> gmax;
    DevBuf<int> act_idx;     // compacted list of the active nodes
    act_idx.alloc(Nn);
    r.x.alloc(nx); r.objective.alloc(Nn);
    Z.alloc(nx); Y.alloc(nx); Yn.alloc(nx); G.alloc(nx); Gn.alloc(nx);
    fY.alloc(Nn); fYn.alloc(Nn); L.alloc(Nn); t.alloc(Nn); tn.alloc(Nn); q.alloc(Nn); c.alloc(Nn); gmap.alloc(Nn);
    gtrue.alloc(Nn);
    status.alloc(Nn); n_active.alloc(1); best.alloc(Nn); stall.alloc(Nn); streak.alloc(Nn); gmax.alloc(1); clamped.alloc(Nn);
    if (prob.x0) GML_CUDA(cudaMemcpyAsync(r.x.p, prob.x0, nx * sizeof(double), cudaMemcpyDeviceToDevice, st));
    else GML_CUDA(cudaMemsetAsync(r.x.p, 0, nx * sizeof(double), st));
    GML_CUDA(cudaMemsetAsync(Y.p, 0, nx * sizeof(double), st));
    GML_CUDA(cudaMemsetAsync(Z.p, 0, nx * sizeof(double), st));

    FistaState s{};
    s.Nn = Nn; s.Fp = Fp; s.form = prob.form; s.lambda = prob.lambda;
    s.tol = o.tol > 0 ? o.tol : 1e-6;   // overwritten per level
    s.lattice = be->lattice();
    s.lattice_inv = s.lattice > 0 ? 1.0 / s.lattice : 0.0;
    s.eps_f = 1e-6;   // generous upper bound of the evaluation noise of f
    s.best = best.p; s.stall = stall.p; s.streak = streak.p; s.clamped = clamped.p; s.gtrue = gtrue.p;
    s.pen = prob.pen.p;
    s.X = r.x.p; s.Z = Z.p; s.Y = Y.p; s.Yn = Yn.p; s.G = G.p; s.Gn = Gn.p;
    s.fY = fY.p; s.fYn = fYn.p; s.L = L.p; s.t = t.p; s.tn = tn.p; s.q = q.p; s.c = c.p; s.gmap = gmap.p;
    s.obj = r.objective.p; s.status = status.p; s.n_active = n_active.p; s.gmax = gmax.p; s.fine = 1;

    double t_setup = 0, t_loop = 0;
    if (o.verbose > 0) { GML_CUDA(cudaStreamSynchronize(st)); t_setup = tick(); }
    // ---- multilevel continuation: solve on strided subsets of the histogram first (every 256th, then every
    // 16th block of 128 samples), warm-starting each level from the previous one.  A coarse level only needs
    // the accuracy of its own sampling error (~1/sqrt(M_level)), costs 1/stride of a pass per round, and leaves
    // the full-data level a start that is ~1e-3 from its optimum instead of |x| ~ 0.4.
    const Histogram& hist = *prob.hist;
    const int64_t SB = hist.Kp / 128;
    std::vector<int64_t> strides;
    // Opt-in (opts.reserved[1]): on the well-conditioned benchmark problems a cold start converges in ~35-45
    // rounds and the warm start from a 16x subsample saves fewer rounds than its momentum restart costs
    // (measured at N=1000, K=1e6: 29 coarse + 40 fine rounds vs 35 cold).
    const bool can_subsample = o.reserved[1] != 0 && be->set_subsample(1, st) > 0.0;
    if (can_subsample && SB >= 256 * 64) strides.push_back(256);
    if (can_subsample && SB >= 16 * 64) strides.push_back(16);
    strides.push_back(1);
    const double user_tol = o.tol > 0 ? o.tol : 1e-6;
    const bool scale_free = prob.form == GML_B200_LOGRISE;     // grad log Z does not scale with the weight mass
    int n_fg = 0, n_f = 0, it = 0, active = Nn;
    double fg_units = 0.0, rho_prev = 1.0;
    const int max_iter = o.max_iter > 0 ? o.max_iter : 5000;
    for (size_t li = 0; li < strides.size(); ++li) {
        const int64_t stride = strides[li];
        const bool last = li + 1 == strides.size();
        const double rho = stride > 1 ? be->set_subsample(stride, st) : (can_subsample ? be->set_subsample(1, st) : 1.0);
        const double scale = scale_free ? 1.0 : rho;
        // weights are not renormalised: f_level = rho * (normalised f), so lambda, L and the tolerance scale by rho
        s.lambda = prob.lambda * scale;
        s.tol = scale * (last ? user_tol : std::max(user_tol, 0.1 / std::sqrt(std::max(hist.M * rho, 1.0))));
        fista_init_kernel<<<(unsigned)ceil_div(Nn, 128), 128, 0, st>>>(s, li == 0 ? scale : -(scale_free ? 1.0 : rho / rho_prev));
        GML_LAUNCHED();
        rho_prev = rho;
        GML_CUDA(cudaMemcpyAsync(Y.p, r.x.p, nx * sizeof(double), cudaMemcpyDeviceToDevice, st));   // warm start
        // Precision levels: every node runs with a 3-limb iterate (lattice 2^-22) and 16-bit residuals; it retires there when
        // the exact prox-gradient mapping meets the tolerance, or parks at the resolution of that lattice; when no node is
        // active any more, the parked ones finish at full precision.  Coarse-lattice points are fine-lattice points.
        // Levels of a lattice backend: 0 coarse, 1 fine; -1 rough (2-limb iterate on 2^-14, one 8-bit residual plane: 3 executed
        // GEMM units per pass instead of 5) only on request (opts.reserved[3] == 2).  A node runs on a level until it reaches
        // the tolerance or the resolution of that level's lattice, parks, and when all have parked the solve moves one level
        // up.  The rough level is NOT part of the default path.  Its kernels are exact for what they are given (parity-tested
        // like the others) and a rough pass costs 23.3 ms against 31.3 ms (ncu), but an 8-bit residual is not a NOISY
        // gradient, it is a BIASED one: t = s_u E takes few distinct values per node (the configurations of its strong
        // neighbours), so the rounding error is a deterministic function of exactly the spins the gradient correlates it
        // with.  Measured at C3: the gradient mapping stalls at ~3e-3 (100x the random-rounding estimate), the nodes park
        // there and the solve needs 93 rounds / 2.41 s instead of 35 / 1.16 s; the same happens with the 3-limb iterate and
        // one residual plane (176 rounds).  profiles/r2_rough_level_experiment.json.
        int level = 1;
        if ((o.reserved[3] == 0 || o.reserved[3] == 2) && user_tol <= 1e-4) {
            if (o.reserved[3] == 2 && !prob.x0 && be->set_level(-1, st)) level = -1;
            else if (be->set_level(0, st)) level = 0;
        }
        be->set_level(1, st);
        s.x_range = be->x_range();          // the clamp is the FINE level's range on all levels: the lower levels' own
        if (level != 1) be->set_level(level, st);   // (smaller) range is watched by the quantiser, which raises device_flags()
        auto sync_level = [&] {
            s.fine = level == 1 ? 1 : 0; s.lattice = be->lattice(); s.lattice_inv = s.lattice > 0 ? 1.0 / s.lattice : 0.0;
            s.eps_g = be->grad_noise() * scale;
            // retirement on the coarse level needs its gradient noise (conservative estimate) well below the tolerance
            s.coarse_retire = (level == 0 && s.eps_g <= 0.25 * s.tol) ? 1 : 0;
        };
        sync_level();
        if (prob.x0 && li == 0 && s.lattice > 0.0) {
            // a warm start (e.g. the previous point of a lambda path, on the fine lattice) is first evaluated on this
            // level's lattice: move X and Y there, so that the stored (f, G) belong to the point the driver holds
            fista_snap_kernel<<<(unsigned)ceil_div((int64_t)nx, 256), 256, 0, st>>>(s, r.x.p, (int64_t)nx);
            GML_LAUNCHED();
            GML_CUDA(cudaMemcpyAsync(Y.p, r.x.p, nx * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        // Active-set compaction: once the active nodes fill at most 7/8 of the slots of the current pass, the passes are
        // restricted to the active nodes (ordered list built on the device; the host knows the count from its per-round
        // read).  Parked / converged nodes in the list stay as dead slots until the next compaction.
        int cap = 0;                    // slots of the current passes (0: the backend evaluates the whole shard)
        bool can_compact = o.reserved[6] == 0;
        auto pass_units = [&] { return (cap ? (double)cap / Nn : 1.0) / stride; };
        auto whole_shard = [&] { if (cap) { be->set_active(nullptr, 0, st); cap = 0; } };
        auto maybe_compact = [&](int n_on) {
            // worth it as soon as the padded slot count (128-node gradient tiles) shrinks by an eighth: the gather of the
            // active nodes' spins costs a small fraction of one pass
            const int64_t cur_pad = round_up(cap ? cap : Nn, 128), new_pad = round_up(n_on, 128);
            const bool shrinks = new_pad * 8 <= cur_pad * 7 || (cur_pad == 128 && round_up(n_on, 64) < round_up(cap ? cap : Nn, 64));
            if (!can_compact || n_on <= 0 || !shrinks) return;
            fista_compact_kernel<<<1, 1024, 0, st>>>(s, act_idx.p);
            GML_LAUNCHED();
            if (be->set_active(act_idx.p, n_on, st)) cap = n_on; else can_compact = false;
        };
        s.park_tol = s.tol;
        be->set_active(nullptr, 0, st);
        be->eval(Y.p, true, fY.p, G.p, st); ++n_fg; fg_units += pass_units();
        // Cold full pairwise solve -> start from the mean-field couplings read off the gradient at 0 (= minus the pair
        // correlations), see warmstart.cu.  One more pass; measured at C3: 43 -> 35 rounds, learn() 1.29 -> 1.16 s.
        // Default for cold solves of ALL nodes of a pairwise problem with 128 <= N <= 2048 (the N x N inverse is ~N tiny
        // launches); opts.reserved[7] & 1 forces it on, & 4 switches it off.  Node shards (Nn < N) cannot use it: the
        // inverse needs every row of the correlation matrix.
        const bool mf_default = hist.N >= 128 && hist.N <= 2048 && (o.reserved[7] & 4) == 0;
        if (li == 0 && strides.size() == 1 && !prob.x0 && ((o.reserved[7] & 1) != 0 || mf_default) && Nn == hist.N && prob.Q == hist.base.p &&
            prob.F == hist.N + 1) {
            const double xmax = level <= 0 ? 1.9 : 7.0;
            if (meanfield_start(G.p, Nn, Fp, prob.pen.p, xmax, s.lattice, Y.p, st)) {
                GML_CUDA(cudaMemcpyAsync(r.x.p, Y.p, nx * sizeof(double), cudaMemcpyDeviceToDevice, st));
                GML_CUDA(cudaMemcpyAsync(Z.p, Y.p, nx * sizeof(double), cudaMemcpyDeviceToDevice, st));
                be->eval(Y.p, true, fY.p, G.p, st); ++n_fg; fg_units += pass_units();
                if (o.verbose > 0) fprintf(stderr, "[gml_b200] fista: mean-field warm start\n");
            } else {
                GML_CUDA(cudaMemsetAsync(Y.p, 0, nx * sizeof(double), st));      // singular correlation matrix: stay cold
            }
        }
        active = Nn;
        bool stale_g = false;
        for (; it < max_iter; ++it) {
            const bool trace = o.verbose > 2 && it == 5;     // one round dissected with events
            cudaEvent_t ev[4];
            if (trace) for (auto& e : ev) cudaEventCreate(&e);
            if (trace) cudaEventRecord(ev[0], st);
            GML_CUDA(cudaMemsetAsync(gmax.p, 0, sizeof(unsigned long long), st));
            fista_trial_kernel<<<Nn, 128, 0, st>>>(s);
            GML_LAUNCHED();
            if (trace) cudaEventRecord(ev[1], st);
            be->eval(Yn.p, true, fYn.p, Gn.p, st); ++n_fg; fg_units += pass_units();
            if (trace) cudaEventRecord(ev[2], st);
            GML_CUDA(cudaMemsetAsync(n_active.p, 0, sizeof(int), st));
            fista_accept_kernel<<<Nn, 128, 0, st>>>(s);
            GML_LAUNCHED();
            if (trace) {
                cudaEventRecord(ev[3], st); cudaEventSynchronize(ev[3]);
                float a, b, c;
                cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]); cudaEventElapsedTime(&c, ev[2], ev[3]);
                fprintf(stderr, "[gml_b200] round %d: trial %.3f ms, pass %.3f ms, accept %.3f ms\n", it, a, b, c);
                for (auto& e : ev) cudaEventDestroy(e);
            }
            double h_gmax = 0.0;
            int h_flags = 0;
            GML_CUDA(cudaMemcpyAsync(&active, n_active.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            GML_CUDA(cudaMemcpyAsync(&h_gmax, gmax.p, sizeof(double), cudaMemcpyDeviceToHost, st));
            if (level < 1 && be->device_flags())
                GML_CUDA(cudaMemcpyAsync(&h_flags, be->device_flags(), sizeof(int), cudaMemcpyDeviceToHost, st));
            GML_CUDA(cudaStreamSynchronize(st));       // the only host sync of the round
            if (stale_g && level == 1) { stale_g = false; sync_level(); }      // every stored G is a fine-level one from here on (rejected nodes keep theirs: conservative eps for one round only matters for the test's noise gate)
            if (level < 1) {
                const bool overflow = (h_flags & 2) != 0;           // some |x| reached 1: the range of the lower levels is exhausted
                if (overflow) be->note_coarse_overflow();
                // every node has parked (reached the tolerance or the resolution of the coarse lattice): the stragglers
                // finish on the coarse level in compacted -- cheap -- passes instead of dragging all nodes to the fine one
                if (overflow || active == 0) {
                    const int from_level = level;
                    const double eps_g_from = s.eps_g;
                    level = overflow ? 1 : level + 1;
                    if (!be->set_level(level, st)) { level = 1; be->set_level(1, st); }
                    sync_level();
                    if (o.verbose > 0) fprintf(stderr, "[gml_b200] fista: %s precision from round %d (gmap max %.3g, %d nodes still active%s)\n", level == 1 ? "fine" : "coarse", it + 1, h_gmax, active, overflow ? ", range overflow" : "");
                    whole_shard();
                    if (overflow) {
                        // restart from the last accepted iterate at full precision
                        fista_init_kernel<<<(unsigned)ceil_div(Nn, 128), 128, 0, st>>>(s, -1.0);
                        GML_LAUNCHED();
                        GML_CUDA(cudaMemcpyAsync(Y.p, r.x.p, nx * sizeof(double), cudaMemcpyDeviceToDevice, st));
                    } else {
                        fista_unpark_kernel<<<(unsigned)ceil_div(Nn, 128), 128, 0, st>>>(s);      // parked nodes rejoin with their state
                        GML_LAUNCHED();
                    }
                    active = Nn;
                    // After a range overflow Y moved (restart from X); after the rough level the stored G carries an 8-bit
                    // residual's bias: both need a fresh (f, G).  Coarse -> fine needs none: Y is a point of both lattices,
                    // the energies are exact on both levels and the per-sample arithmetic is the same, so f(Y) is the same
                    // number, and the 16-bit residual rounding left ~6e-8 in G -- far below any tolerance the fine level is
                    // asked for.  (Round 1 spent a whole 4-limb / 3-plane pass here: 46 ms of 1.16 s at C3.)
                    if (overflow || from_level < 0) { be->eval(Y.p, true, fY.p, G.p, st); ++n_fg; fg_units += pass_units(); }
                    else { s.eps_g = std::max(s.eps_g, eps_g_from); stale_g = true; }     // the next round still compares against the coarse G
                } else {
                    maybe_compact(active);
                }
            } else {
                maybe_compact(active);
            }
            if (o.verbose > 1) {
                std::vector<double> hg(Nn), hL(Nn);
                GML_CUDA(cudaMemcpy(hg.data(), gmap.p, sizeof(double) * Nn, cudaMemcpyDeviceToHost));
                GML_CUDA(cudaMemcpy(hL.data(), L.p, sizeof(double) * Nn, cudaMemcpyDeviceToHost));
                double gmax = 0, gsum = 0, lmax = 0, lmin = 1e300;
                for (int u = 0; u < Nn; ++u) { gmax = std::max(gmax, hg[u]); gsum += hg[u]; lmax = std::max(lmax, hL[u]); lmin = std::min(lmin, hL[u]); }
                fprintf(stderr, "[gml_b200] fista level %zu (stride %lld) round %d active %d gmap max %.3e mean %.3e L [%.3g, %.3g]\n",
                        li, (long long)stride, it, active, gmax / scale, gsum / Nn / scale, lmin / scale, lmax / scale);
            }
            if (active == 0) { ++it; break; }
        }
        if (o.verbose > 0) fprintf(stderr, "[gml_b200] fista level %zu: stride %lld rho %.6g tol %.3g, rounds so far %d\n", li, (long long)stride, rho, s.tol, it);
    }
    s.lambda = prob.lambda;
    if (o.verbose > 0) { GML_CUDA(cudaStreamSynchronize(st)); t_loop = tick(); }
    // objective at the returned point (all nodes)
    be->set_active(nullptr, 0, st);
    if (!r.want_objective && !r.want_grad_at_x) {
        // nobody asked for the objective values (the reference's learn() returns none): no final pass
        GML_CUDA(cudaMemsetAsync(r.objective.p, 0, sizeof(double) * Nn, st));
        r.f_units = 0.0;
    } else if (r.want_grad_at_x) {          // the support polish reads the gradient at the returned point
        r.grad.alloc(nx);
        be->eval(r.x.p, true, fYn.p, r.grad.p, st); ++n_fg; fg_units += 1.0;
        r.f_units = 0.0;
    } else {
        be->eval(r.x.p, false, fYn.p, nullptr, st); ++n_f;
        r.f_units = 1.0;
    }
    r.fg_units = fg_units;
    if (r.want_objective || r.want_grad_at_x) {
        fista_objective_kernel<<<Nn, 128, 0, st>>>(s, fYn.p);
        GML_LAUNCHED();
    }

    std::vector<double> hg(Nn);
    std::vector<int> hs(Nn);
    GML_CUDA(cudaMemcpyAsync(hg.data(), gmap.p, sizeof(double) * Nn, cudaMemcpyDeviceToHost, st));
    GML_CUDA(cudaMemcpyAsync(hs.data(), status.p, sizeof(int) * Nn, cudaMemcpyDeviceToHost, st));
    GML_CUDA(cudaStreamSynchronize(st));
    if (o.verbose > 0) {
        fprintf(stderr, "[gml_b200] fista: setup %.2f ms, %d rounds %.2f ms, final %.2f ms\n", t_setup - t_begin, it,
                t_loop - t_setup, tick() - t_loop);
        std::vector<double> hL(Nn), hq(Nn), hc(Nn), hfY(Nn), hfN(Nn), ht(Nn);
        GML_CUDA(cudaMemcpy(hL.data(), L.p, sizeof(double) * Nn, cudaMemcpyDeviceToHost));
        GML_CUDA(cudaMemcpy(hq.data(), q.p, sizeof(double) * Nn, cudaMemcpyDeviceToHost));
        GML_CUDA(cudaMemcpy(hc.data(), c.p, sizeof(double) * Nn, cudaMemcpyDeviceToHost));
        GML_CUDA(cudaMemcpy(hfY.data(), fY.p, sizeof(double) * Nn, cudaMemcpyDeviceToHost));
        GML_CUDA(cudaMemcpy(hfN.data(), fYn.p, sizeof(double) * Nn, cudaMemcpyDeviceToHost));
        GML_CUDA(cudaMemcpy(ht.data(), t.p, sizeof(double) * Nn, cudaMemcpyDeviceToHost));
        for (int u = 0; u < Nn; ++u)
            if (hs[u] != 1 || o.verbose > 2)
                fprintf(stderr, "[gml_b200] node %d status %d L %.6g t %.4g gmap %.3e q %.3e c %.3e fY %.17g fX %.17g\n",
                        u, hs[u], hL[u], ht[u], hg[u], hq[u], hc[u], hfY[u], hfN[u]);
    }
    double mr = 0.0; int unconv = 0, stalled = 0;
    // A node whose gradient mapping stopped improving above tol (status 2: the noise floor of the backend's gradient)
    // is accepted when it is within 10x tol and REPORTED in n_stalled / max_residual (gml_b200_stats); beyond that, or
    // when max_iter ran out, it is unconverged (GML_B200_ENOTCONV).  Status 4 nodes are handed to the caller.
    for (int u = 0; u < Nn; ++u) {
        if (hs[u] == 4) { out_of_range.push_back(u); continue; }
        mr = std::max(mr, hg[u]);
        if (hs[u] == 2 && hg[u] <= 10.0 * s.tol) ++stalled;
        else if (hs[u] != 1) ++unconv;
    }
    be->collect_profile(r.profile);
    r.iterations = it; r.n_fg = n_fg; r.n_f = n_f; r.n_unconverged = unconv; r.n_stalled = stalled; r.max_residual = mr;
}

__global__ void gather_subproblem_kernel(const int* __restrict__ idx, int Fp, const int32_t* __restrict__ spin_row, const uint8_t* __restrict__ pen,
                                         const double* __restrict__ x, int32_t* __restrict__ spin_row_s, uint8_t* __restrict__ pen_s,
                                         double* __restrict__ x_s) {
    const int j = blockIdx.x, u = idx[j];
    if (threadIdx.x == 0) spin_row_s[j] = spin_row[u];
    for (int f = threadIdx.x; f < Fp; f += blockDim.x) {
        pen_s[(int64_t)j * Fp + f] = pen[(int64_t)u * Fp + f];
        x_s[(int64_t)j * Fp + f] = x[(int64_t)u * Fp + f];
    }
}
__global__ void scatter_subsolution_kernel(const int* __restrict__ idx, int Fp, const double* __restrict__ x_s, const double* __restrict__ obj_s,
                                           double* __restrict__ x, double* __restrict__ obj) {
    const int j = blockIdx.x, u = idx[j];
    if (threadIdx.x == 0) obj[u] = obj_s[j];
    for (int f = threadIdx.x; f < Fp; f += blockDim.x) x[(int64_t)u * Fp + f] = x_s[(int64_t)j * Fp + f];
}
}  // namespace

// Front end: the tensor-core backend holds the iterate as a fixed-point number (|x| < 7.9).  Nodes whose optimum lies
// beyond that (near-deterministic couplings: the L1 optimum grows like -log lambda) leave its passes with status 4 and
// are re-solved here, warm-started from their last accepted point, by the CUDA-core backend, which has no such bound --
// the reference's Ipopt has none either (src/GraphicalModelLearning.jl:169-177).
void solve_fista(const NodeProblem& prob, const gml_b200_opts& o, int backend, SolveResult& r, cudaStream_t st) {
    std::vector<int> oor;
    solve_fista_impl(prob, o, backend, r, oor, st);
    if (oor.empty()) return;
    GML_REQUIRE(backend == GML_B200_SOLVER_FISTA_TC && prob.comm == nullptr,
                "an iterate left the representable range of the solver backend");
    if (o.verbose > 0) fprintf(stderr, "[gml_b200] fista: %zu node(s) left the fixed-point range |x| < 7.9: re-solving them with the CUDA-core backend\n", oor.size());
    const int n = (int)oor.size(), Fp = prob.Fp;
    DevBuf<int> idx;
    DevBuf<double> x0;
    idx.alloc(n); x0.alloc((size_t)n * Fp);
    GML_CUDA(cudaMemcpyAsync(idx.p, oor.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
    NodeProblem sub;
    sub.hist = prob.hist; sub.Q = prob.Q; sub.F = prob.F; sub.Fp = Fp; sub.form = prob.form; sub.lambda = prob.lambda; sub.Nn = n;
    sub.spin_row.alloc(n); sub.pen.alloc((size_t)n * Fp);
    gather_subproblem_kernel<<<n, 128, 0, st>>>(idx.p, Fp, prob.spin_row.p, prob.pen.p, r.x.p, sub.spin_row.p, sub.pen.p, x0.p);
    GML_LAUNCHED();
    sub.x0 = x0.p;
    SolveResult rs;
    std::vector<int> none;
    solve_fista_impl(sub, o, GML_B200_SOLVER_FISTA_CC, rs, none, st);
    scatter_subsolution_kernel<<<n, 128, 0, st>>>(idx.p, Fp, rs.x.p, rs.objective.p, r.x.p, r.objective.p);
    GML_LAUNCHED();
    GML_CUDA(cudaStreamSynchronize(st));
    r.n_fg += rs.n_fg; r.n_f += rs.n_f;
    if (r.fg_units >= 0.0) r.fg_units += (rs.fg_units >= 0.0 ? rs.fg_units : rs.n_fg) * (double)n / prob.Nn;
    r.iterations += rs.iterations;
    r.n_unconverged += rs.n_unconverged; r.n_stalled += rs.n_stalled;
    r.n_out_of_range = n;
    r.max_residual = std::max(r.max_residual, rs.max_residual);
}

}  // namespace gml
