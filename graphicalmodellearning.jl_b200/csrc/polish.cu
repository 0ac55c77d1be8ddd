// Reduced-space Newton polish: a first-order solve to tolerance tol leaves every node ~tol/mu from its
// optimum with the SUPPORT already identified (degree << N).  On that support the node problem has at most a few dozen
// coordinates, so the fp64 Newton solver of newton.cu (dense Hessian over the histogram, coordinate descent on the L1
// model) finishes it quadratically: the exact L1 minimiser to ~1e-12 instead of ~1e-6, for any number of features per
// node -- what the reference gets from Ipopt's second-order iterations (src/GraphicalModelLearning.jl:164-181).
// The support of node u: coordinates that are free, nonzero, or whose gradient is within 2 % of the threshold lambda
// (they may enter).  After the reduced solve the optimality conditions off the support are re-checked with a full
// gradient pass at the polished point and violators are added (one repeat).  Nodes whose support exceeds the reduced
// solver's feature limit keep their first-order solution and are reported (n_unpolished).
// (The Ipopt-compatible barrier point couples ALL coordinates at the 1e-6 level -- no coordinate is exactly zero there --
// so it is not a reduced-space problem: barrier_mu is served by the full Newton solver up to NEWTON_MAX_F features.)
#include "common.cuh"

#include <memory>

namespace gml {
namespace {

constexpr double SUPPORT_MARGIN = 0.98;     // |g_j| >= 0.98 lambda: j may enter the support

// one block per node: ordered support list, reduced start point / penalty classes
__global__ void __launch_bounds__(256) polish_select_kernel(const double* __restrict__ x, const double* __restrict__ g, const uint8_t* __restrict__ pen,
                                                          int Fp, int Fr, double lambda, int32_t* __restrict__ feat, uint8_t* __restrict__ pen_r,
                                                          double* __restrict__ x_r, int* __restrict__ count) {
    const int u = blockIdx.x;
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int64_t o = (int64_t)u * Fp, orr = (int64_t)u * Fr;
    // ordered compaction, one warp-sized chunk of features at a time (a few hundred features at most: serial chunks)
    for (int f0 = 0; f0 < Fp; f0 += blockDim.x) {
        const int f = f0 + threadIdx.x;
        bool in = false;
        if (f < Fp) {
            const uint8_t pc = pen[o + f];
            in = pc == PEN_FREE || (pc == PEN_L1 && (x[o + f] != 0.0 || fabs(g[o + f]) >= SUPPORT_MARGIN * lambda));
        }
        // block-wide ordered rank through ballots
        __shared__ int warp_cnt[8];
        const unsigned m = __ballot_sync(0xffffffffu, in);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_cnt[w] = __popc(m);
        __syncthreads();
        int off = s_n;
        for (int i = 0; i < w; ++i) off += warp_cnt[i];
        const int pos = off + __popc(m & ((1u << lane) - 1));
        if (in && pos < Fr) {
            feat[orr + pos] = f;
            pen_r[orr + pos] = pen[o + f];
            x_r[orr + pos] = x[o + f];
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += warp_cnt[i]; s_n += t; }
        __syncthreads();
    }
    const int n = s_n;
    for (int j = threadIdx.x; j < Fr; j += blockDim.x)
        if (j >= min(n, Fr)) { feat[orr + j] = 0; pen_r[orr + j] = PEN_ZERO; x_r[orr + j] = 0.0; }
    if (threadIdx.x == 0) count[u] = n;
}

// nodes whose support does not fit keep their first-order point: switch every reduced coordinate off
__global__ void polish_disable_kernel(const int* __restrict__ count, int Fr, uint8_t* __restrict__ pen_r, double* __restrict__ x_r) {
    const int u = blockIdx.x;
    if (count[u] <= Fr) return;
    for (int j = threadIdx.x; j < Fr; j += blockDim.x) { pen_r[(int64_t)u * Fr + j] = PEN_ZERO; x_r[(int64_t)u * Fr + j] = 0.0; }
}

// polished reduced solution -> full coordinates; barrier mode: closed-form values off the support
__global__ void __launch_bounds__(256) polish_scatter_kernel(const double* __restrict__ x_r, const int32_t* __restrict__ feat, const uint8_t* __restrict__ pen_r,
                                                           const int* __restrict__ count, int Fp, int Fr, const uint8_t* __restrict__ pen,
                                                           const double* __restrict__ g, double lambda, double mu, double* __restrict__ x,
                                                           double* __restrict__ obj, const double* __restrict__ obj_r) {
    const int u = blockIdx.x;
    if (count[u] > Fr) return;                          // unpolished node: untouched
    const int64_t o = (int64_t)u * Fp, orr = (int64_t)u * Fr;
    __shared__ double red[8];
    double l1_off = 0.0;
    for (int f = threadIdx.x; f < Fp; f += blockDim.x) {
        double v = 0.0;
        if (mu > 0.0 && pen[o + f] == PEN_L1) {
            const double gj = g[o + f];
            v = -2.0 * mu * gj / fmax(lambda * lambda - gj * gj, 1e-300);
            l1_off += fabs(v);
        }
        x[o + f] = v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < count[u]; j += blockDim.x) {
        const int f = feat[orr + j];
        if (mu > 0.0 && pen_r[orr + j] == PEN_L1) l1_off -= fabs(x[o + f]);
        x[o + f] = x_r[orr + j];
    }
    for (int s = 16; s; s >>= 1) l1_off += __shfl_xor_sync(0xffffffffu, l1_off, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l1_off;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        obj[u] = obj_r[u] + lambda * t;
    }
}

// optimality check of the polished point off the support: |g_j| <= lambda (exact mode).  Violators get a tiny value of the
// right sign so that the next selection picks them up.
__global__ void __launch_bounds__(256) polish_verify_kernel(double* __restrict__ x, const double* __restrict__ g, const uint8_t* __restrict__ pen, int Fp,
                                                          double lambda, double slack, int* __restrict__ n_viol) {
    const int u = blockIdx.x;
    const int64_t o = (int64_t)u * Fp;
    int v = 0;
    for (int f = threadIdx.x; f < Fp; f += blockDim.x)
        if (pen[o + f] == PEN_L1 && x[o + f] == 0.0 && fabs(g[o + f]) > lambda + slack) ++v;
    if (v) atomicAdd(n_viol, v);
}

}  // namespace

void polish_on_support(const NodeProblem& prob, const gml_b200_opts& o, SolveResult& r, cudaStream_t st) {
    GML_REQUIRE(prob.comm == nullptr, "the support polish is not available in the sample-sharded mode");
    GML_REQUIRE(o.barrier_mu == 0.0, "the support polish returns the exact L1 minimiser (barrier_mu needs the full Newton solver)");
    GML_REQUIRE(r.grad.p != nullptr, "internal: polish needs the gradient at the first-order solution");
    const int Nn = prob.Nn, Fp = prob.Fp, Fr = NEWTON_MAX_F;
    DevBuf<int32_t> feat;
    DevBuf<uint8_t> pen_r;
    DevBuf<double> x_r;
    DevBuf<int> count, n_viol;
    feat.alloc((size_t)Nn * Fr); pen_r.alloc((size_t)Nn * Fr); x_r.alloc((size_t)Nn * Fr); count.alloc(Nn); n_viol.alloc(1);
    std::vector<int> h_count(Nn);
    std::unique_ptr<EvalBackend> be;
    gml_b200_opts on = o;
    // The reduced solve stops at a Newton step of 1e-10: the support deliberately contains coordinates AT the threshold
    // (|g_j| within 2 % of lambda), some of which are degenerate to rounding and keep flipping between 0 and ~1e-11.
    on.tol = 1e-10;
    on.max_iter = 60;
    for (int round = 0; round < 2; ++round) {
        polish_select_kernel<<<Nn, 256, 0, st>>>(r.x.p, r.grad.p, prob.pen.p, Fp, Fr, prob.lambda, feat.p, pen_r.p, x_r.p, count.p);
        GML_LAUNCHED();
        polish_disable_kernel<<<Nn, 64, 0, st>>>(count.p, Fr, pen_r.p, x_r.p);
        GML_LAUNCHED();
        GML_CUDA(cudaMemcpyAsync(h_count.data(), count.p, sizeof(int) * Nn, cudaMemcpyDeviceToHost, st));
        GML_CUDA(cudaStreamSynchronize(st));
        int fmax = 1, unpolished = 0;
        for (int u = 0; u < Nn; ++u) { if (h_count[u] > Fr) ++unpolished; else fmax = std::max(fmax, h_count[u]); }
        r.n_unpolished = unpolished;
        if (o.verbose > 0) fprintf(stderr, "[gml_b200] polish round %d: largest support %d, %d node(s) beyond %d features\n", round, fmax, unpolished, Fr);
        NodeProblem red;
        red.hist = prob.hist; red.Q = prob.Q; red.F = fmax; red.Fp = Fr; red.form = prob.form; red.lambda = prob.lambda; red.Nn = Nn;
        red.spin_row.alloc(Nn); red.pen.alloc((size_t)Nn * Fr);
        GML_CUDA(cudaMemcpyAsync(red.spin_row.p, prob.spin_row.p, sizeof(int32_t) * Nn, cudaMemcpyDeviceToDevice, st));
        GML_CUDA(cudaMemcpyAsync(red.pen.p, pen_r.p, (size_t)Nn * Fr, cudaMemcpyDeviceToDevice, st));
        red.x0 = x_r.p; red.feat = feat.p;
        SolveResult rr;
        solve_newton(red, on, rr, st);
        polish_scatter_kernel<<<Nn, 256, 0, st>>>(rr.x.p, feat.p, pen_r.p, count.p, Fp, Fr, prob.pen.p, r.grad.p, prob.lambda, 0.0,
                                                  r.x.p, r.objective.p, rr.objective.p);
        GML_LAUNCHED();
        r.iterations += rr.iterations; r.n_fg += rr.n_fg; r.n_f += rr.n_f;
        // a reduced solve that ran out of iterations with steps already below 1e-8 (flipping threshold coordinates) is
        // still a point far inside the first-order tolerance: reported through max_residual, not as a failure
        if (rr.max_residual > 1e-8) r.n_unconverged += rr.n_unconverged;
        r.max_residual = rr.max_residual;
        if (round == 1) break;
        // re-check the optimality conditions off the support with a gradient at the polished point (CUDA-core backend:
        // the polished point is not on the tensor-core backend's lattice)
        if (!be) be.reset(make_backend_cc(prob, st));
        DevBuf<double> f_tmp;
        f_tmp.alloc(Nn);
        be->eval(r.x.p, true, f_tmp.p, r.grad.p, st); ++r.n_fg;
        if (r.fg_units >= 0.0) r.fg_units += 1.0;
        GML_CUDA(cudaMemsetAsync(n_viol.p, 0, sizeof(int), st));
        polish_verify_kernel<<<Nn, 256, 0, st>>>(r.x.p, r.grad.p, prob.pen.p, Fp, prob.lambda, 5e-6, n_viol.p);
        GML_LAUNCHED();
        int hv = 0;
        GML_CUDA(cudaMemcpyAsync(&hv, n_viol.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        GML_CUDA(cudaStreamSynchronize(st));
        if (o.verbose > 0) fprintf(stderr, "[gml_b200] polish: %d coordinate(s) off the support violate |g| <= lambda\n", hv);
        if (hv == 0) break;
    }
    GML_CUDA(cudaStreamSynchronize(st));
}

}  // namespace gml
