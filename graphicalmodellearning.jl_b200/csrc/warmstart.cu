// Mean-field warm start of the pairwise FISTA solves (opt-in: opts.reserved[7]).
//
// For all three pairwise formulations the gradient of the smooth part at x = 0 is minus the weighted pair-correlation
// matrix of the histogram (psi == 1 there): G0[u,f] = -sum_k w_k s_u[k] Q[k,f], i.e. -C[u,f] for the spin features and
// -m_u for the constant one.  The solver's first pass therefore already holds C and m, and the naive mean-field
// estimate  J = -(C - m m^T)^-1 (off-diagonal),  h_u = atanh(m_u) - sum_j J_uj m_j  is a starting point ~0.1 from the
// optimum instead of |x| ~ 0.4 (float64 emulation of the driver: ~30 % fewer rounds, scripts/dev_warmstart_study.py).
// The problem is convex, so the start changes the path, not the answer (src/GraphicalModelLearning.jl:169-177).
//
// Everything runs on the device: blocked N x N Gauss-Jordan inversion of the (ridge-stabilised) connected correlation
// matrix over a ping-pong pair of matrices (symmetric positive definite: no pivoting needed).
#include "common.cuh"

#include <cmath>

namespace gml {
namespace {

// A = C - m m^T + ridge I from the first-pass gradient rows (row u of G0 belongs to node u; diagonal of C is 1)
__global__ void mf_build_kernel(const double* __restrict__ G0, int N, int Fp, double ridge, double* __restrict__ A) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * N) return;
    const int i = (int)(idx / N), j = (int)(idx % N);
    const double mi = -G0[(int64_t)i * Fp + N], mj = -G0[(int64_t)j * Fp + N];
    // the two directed sums are the same exact integer in the tensor-core backend; average for the fp32 backend
    const double c = (i == j) ? 1.0 : -0.5 * (G0[(int64_t)i * Fp + j] + G0[(int64_t)j * Fp + i]);
    A[idx] = c - mi * mj + (i == j ? ridge : 0.0);
}

// Blocked Gauss-Jordan inversion, out of place over a ping-pong pair, MF_B pivots per step (symmetric positive definite +
// ridge: no pivoting needed).  With K = the MF_B rows / columns of the step, D = A[K,K]:
//     out[K,K]   = D^-1                         out[K,j]   = D^-1 A[K,j]                  (j not in K)
//     out[i,K]   = -A[i,K] D^-1                 out[i,j]   = A[i,j] - A[i,K] D^-1 A[K,j]  (i, j not in K)
// i.e. MF_B steps of the scalar elimination at once: ~3 N / MF_B launches instead of N (the scalar form, one launch per
// pivot, was 8.5 ms of every C3 solve -- and, not being divided by the GPU count, most of what separated the 8-GPU run from
// perfect scaling).
constexpr int MF_B = 32;

// D^-1 of the diagonal block (one CTA, scalar Gauss-Jordan in shared memory); rows / columns beyond N act as identity
__global__ void __launch_bounds__(MF_B * MF_B) mf_block_inverse_kernel(const double* __restrict__ A, int N, int k0, double* __restrict__ Dinv,
                                                                     int* __restrict__ bad) {
    __shared__ double S[MF_B][MF_B + 1];
    const int r = threadIdx.x / MF_B, c = threadIdx.x % MF_B;
    const int gr = k0 + r, gc = k0 + c;
    S[r][c] = (gr < N && gc < N) ? A[(int64_t)gr * N + gc] : (r == c ? 1.0 : 0.0);
    __syncthreads();
    for (int p = 0; p < MF_B; ++p) {
        const double piv = S[p][p], f = S[r][p], prow = S[p][c], v = S[r][c];
        __syncthreads();
        if (!(piv > 1e-12) || !isfinite(piv)) { if (threadIdx.x == 0) *bad = 1; }
        else {
            const double rp = 1.0 / piv;
            S[r][c] = (r == p) ? (c == p ? rp : prow * rp) : (c == p ? -f * rp : v - f * prow * rp);
        }
        __syncthreads();
    }
    Dinv[r * MF_B + c] = S[r][c];
}

// T[r, j] = sum_c D^-1[r, c] A[k0 + c, j]   (MF_B x N row panel)
__global__ void mf_row_panel_kernel(const double* __restrict__ A, int N, int k0, const double* __restrict__ Dinv, double* __restrict__ T) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (j >= N) return;
    double acc = 0.0;
    for (int c = 0; c < MF_B && k0 + c < N; ++c) acc += Dinv[r * MF_B + c] * A[(int64_t)(k0 + c) * N + j];
    T[(int64_t)r * N + j] = acc;
}

__global__ void mf_block_update_kernel(const double* __restrict__ Ain, double* __restrict__ Aout, int N, int k0, const double* __restrict__ Dinv,
                                       const double* __restrict__ T) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= N) return;
    const bool iK = i >= k0 && i < k0 + MF_B, jK = j >= k0 && j < k0 + MF_B;
    double out;
    if (iK && jK) out = Dinv[(i - k0) * MF_B + (j - k0)];
    else if (iK) out = T[(int64_t)(i - k0) * N + j];
    else {
        const double* arow = Ain + (int64_t)i * N + k0;
        double acc = 0.0;
        if (jK) { for (int c = 0; c < MF_B && k0 + c < N; ++c) acc += arow[c] * Dinv[c * MF_B + (j - k0)]; out = -acc; }
        else { for (int c = 0; c < MF_B && k0 + c < N; ++c) acc += arow[c] * T[(int64_t)c * N + j]; out = Ain[(int64_t)i * N + j] - acc; }
    }
    Aout[(int64_t)i * N + j] = out;
}

// x0[u, j] = -Ainv[u, j] (couplings), x0[u, N] = atanh(m_u) - sum_j J_uj m_j (field); clamped, snapped to the lattice,
// zero on the coordinates the problem fixes
__global__ void __launch_bounds__(128) mf_start_kernel(const double* __restrict__ Ainv, const double* __restrict__ G0, int N, int Fp,
                                                     const uint8_t* __restrict__ pen, double xmax, double lattice,
                                                     double* __restrict__ x0, int* __restrict__ bad) {
    const int u = blockIdx.x;
    __shared__ double red[4];
    double acc = 0.0;
    for (int f = threadIdx.x; f < Fp; f += blockDim.x) {
        double v = 0.0;
        if (f < N && pen[(int64_t)u * Fp + f] != PEN_ZERO) {
            v = -Ainv[(int64_t)u * N + f];
            if (!isfinite(v)) { *bad = 1; v = 0.0; }
            v = fmin(fmax(v, -xmax), xmax);
            if (lattice > 0.0) v = rint(v / lattice) * lattice;
            acc += v * -G0[(int64_t)f * Fp + N];          // J_uf m_f
        }
        if (f != N) x0[(int64_t)u * Fp + f] = v;
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        const double mu = fmin(fmax(-G0[(int64_t)u * Fp + N], -0.999), 0.999);
        double h = atanh(mu) - (red[0] + red[1] + red[2] + red[3]);
        h = fmin(fmax(h, -xmax), xmax);
        if (lattice > 0.0) h = rint(h / lattice) * lattice;
        x0[(int64_t)u * Fp + N] = (pen[(int64_t)u * Fp + N] != PEN_ZERO && isfinite(h)) ? h : 0.0;
    }
}

}  // namespace

// G0: gradient rows at x = 0 of ALL N nodes [N x Fp]; writes the starting point x0 [N x Fp].  Returns false (x0 is
// then unusable) when the correlation matrix could not be inverted.
bool meanfield_start(const double* G0, int N, int Fp, const uint8_t* pen, double xmax, double lattice, double* x0, cudaStream_t st) {
    DevBuf<double> A, B;
    DevBuf<int> bad;
    const int64_t nn = (int64_t)N * N;
    A.alloc(nn); B.alloc(nn); bad.alloc(1);
    GML_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    const unsigned grid = (unsigned)ceil_div(nn, 256);
    mf_build_kernel<<<grid, 256, 0, st>>>(G0, N, Fp, 1e-6, A.p);
    GML_LAUNCHED();
    DevBuf<double> Dinv, T;
    Dinv.alloc(MF_B * MF_B); T.alloc((size_t)MF_B * N);
    double *in = A.p, *out = B.p;
    for (int k0 = 0; k0 < N; k0 += MF_B) {
        mf_block_inverse_kernel<<<1, MF_B * MF_B, 0, st>>>(in, N, k0, Dinv.p, bad.p);
        GML_LAUNCHED();
        mf_row_panel_kernel<<<dim3((unsigned)ceil_div(N, 256), MF_B), 256, 0, st>>>(in, N, k0, Dinv.p, T.p);
        GML_LAUNCHED();
        mf_block_update_kernel<<<dim3((unsigned)ceil_div(N, 256), (unsigned)N), 256, 0, st>>>(in, out, N, k0, Dinv.p, T.p);
        GML_LAUNCHED();
        std::swap(in, out);
    }
    mf_start_kernel<<<N, 128, 0, st>>>(in, G0, N, Fp, pen, xmax, lattice, x0, bad.p);
    GML_LAUNCHED();
    int h_bad = 0;
    GML_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    GML_CUDA(cudaStreamSynchronize(st));
    return h_bad == 0;
}

}  // namespace gml
