// K2/K3, CUDA-core backend: fp32 tiled contractions with fp64 reductions.
//   energy pass   E[k,u] = sum_f Q[f,k] x_u[f];  t = s_u[k] E;  f_u += w_k g(t);  R[u,k] = s_u[k] w_k g'(t)
//   gradient pass G[u,f] = -sum_k R[u,k] Q[f,k]         (split over sample ranges, fp64 atomics)
// nodal_stat (src/GraphicalModelLearning.jl:162) is never materialised: stat[k,f] = s_u[k]*Q[f,k].
// This backend serves feature counts the Newton path does not take and is the independent on-device
// cross-check of the tensor-core backend.
#include "common.cuh"

namespace gml {
namespace {

constexpr int TK = 128, TU = 32, FC = 32;

template <int FORM>
__device__ __forceinline__ void terms32(float t, float w, float& fterm, float& gw) {
    if (FORM == GML_B200_RPLE) {
        const float a = -2.f * t;
        fterm = w * (fmaxf(a, 0.f) + log1pf(expf(-fabsf(a))));
        gw = 2.f * w / (1.f + expf(2.f * t));
    } else {
        const float e = w * expf(fminf(-t, 80.f));
        fterm = e; gw = e;
    }
}

template <int FORM, bool GRAD>
__global__ void __launch_bounds__(256) cc_energy_kernel(const int8_t* __restrict__ Q, const int8_t* __restrict__ base,
                                                       const float* __restrict__ w32, const int32_t* __restrict__ spin_row,
                                                       const double* __restrict__ x, int F, int Fp, int64_t Kp, int Nn,
                                                       float* __restrict__ R, double* __restrict__ fsum) {
    __shared__ __align__(16) float qs[FC][TK];
    __shared__ float xs[TU][FC + 1];
    const int t = threadIdx.x;
    const int tk = t & 31, tu = t >> 5;
    const int64_t k0 = (int64_t)blockIdx.x * TK;
    const int u0 = blockIdx.y * TU;
    float acc[4][4];   // [node j][sample i]
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

    const int Fr = (F + FC - 1) / FC * FC;
    for (int f0 = 0; f0 < Fr; f0 += FC) {
        {   // stage Q chunk: 32 rows x 128 bytes, 16 bytes per thread
            const int row = t >> 3, col = (t & 7) * 16;
            const int4 v = *reinterpret_cast<const int4*>(Q + (int64_t)(f0 + row) * Kp + k0 + col);
            int8_t b[16];
            memcpy(b, &v, 16);
#pragma unroll
            for (int i = 0; i < 16; ++i) qs[row][col + i] = (float)b[i];
        }
        {   // stage x chunk: 32 nodes x 32 features, 4 per thread
            const int node = t >> 3, ff = (t & 7) * 4;
            const int u = u0 + node;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                xs[node][ff + i] = (u < Nn) ? (float)x[(int64_t)u * Fp + f0 + ff + i] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int ff = 0; ff < FC; ++ff) {
            const float4 a = *reinterpret_cast<const float4*>(&qs[ff][tk * 4]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float b = xs[tu * 4 + j][ff];
                acc[j][0] = fmaf(a.x, b, acc[j][0]); acc[j][1] = fmaf(a.y, b, acc[j][1]);
                acc[j][2] = fmaf(a.z, b, acc[j][2]); acc[j][3] = fmaf(a.w, b, acc[j][3]);
            }
        }
        __syncthreads();
    }
    const float4 w4 = *reinterpret_cast<const float4*>(w32 + k0 + tk * 4);
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int u = u0 + tu * 4 + j;
        if (u >= Nn) continue;   // warp-uniform
        const char4 s4 = *reinterpret_cast<const char4*>(base + (int64_t)spin_row[u] * Kp + k0 + tk * 4);
        const float sv[4] = {(float)s4.x, (float)s4.y, (float)s4.z, (float)s4.w};
        float fs = 0.f, rv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float ft, gw;
            terms32<FORM>(sv[i] * acc[j][i], wv[i], ft, gw);
            fs += ft;
            rv[i] = sv[i] * gw;
        }
        if (GRAD) *reinterpret_cast<float4*>(R + (int64_t)u * Kp + k0 + tk * 4) = make_float4(rv[0], rv[1], rv[2], rv[3]);
        double fd = fs;
        for (int o = 16; o; o >>= 1) fd += __shfl_xor_sync(0xffffffffu, fd, o);
        if (tk == 0) atomicAdd(fsum + u, fd);
    }
}

constexpr int GF = 64, GU = 32, GK = 32;

__global__ void __launch_bounds__(256) cc_grad_kernel(const int8_t* __restrict__ Q, const float* __restrict__ R, int Fp,
                                                     int64_t Kp, int Nn, int64_t kchunk, double* __restrict__ G) {
    __shared__ __align__(16) float qs[GK][GF];
    __shared__ float rs[GU][GK + 1];
    const int t = threadIdx.x;
    const int tf = t & 15, tu = t >> 4;   // features tf*4..+3, nodes tu*2..+1
    const int f0 = blockIdx.x * GF, u0 = blockIdx.y * GU;
    const int64_t kb = (int64_t)blockIdx.z * kchunk, ke = min(kb + kchunk, Kp);
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int64_t k0 = kb; k0 < ke; k0 += GK) {
        {   // Q chunk: 64 rows x 32 bytes, 8 bytes per thread, stored transposed
            const int row = t >> 2, col = (t & 3) * 8;
            const int2 v = *reinterpret_cast<const int2*>(Q + (int64_t)(f0 + row) * Kp + k0 + col);
            int8_t b[8];
            memcpy(b, &v, 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) qs[col + i][row] = (float)b[i];
        }
        {   // R chunk: 32 nodes x 32 samples, 4 floats per thread
            const int node = t >> 3, kk = (t & 7) * 4;
            const int u = u0 + node;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u < Nn) v = *reinterpret_cast<const float4*>(R + (int64_t)u * Kp + k0 + kk);
            rs[node][kk] = v.x; rs[node][kk + 1] = v.y; rs[node][kk + 2] = v.z; rs[node][kk + 3] = v.w;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < GK; ++kk) {
            const float4 b = *reinterpret_cast<const float4*>(&qs[kk][tf * 4]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float a = rs[tu * 2 + j][kk];
                acc[j][0] = fmaf(a, b.x, acc[j][0]); acc[j][1] = fmaf(a, b.y, acc[j][1]);
                acc[j][2] = fmaf(a, b.z, acc[j][2]); acc[j][3] = fmaf(a, b.w, acc[j][3]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int u = u0 + tu * 2 + j;
        if (u >= Nn) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) atomicAdd(G + (int64_t)u * Fp + f0 + tf * 4 + i, -(double)acc[j][i]);
    }
}

// logRISE: f = log Z, G /= Z   (src/GraphicalModelLearning.jl:279)
__global__ void cc_finalize_kernel(int form, int Nn, int Fp, const double* __restrict__ fsum, double* __restrict__ f_out,
                                   double* __restrict__ G, int want_grad) {
    const int u = blockIdx.x;
    const double s = fsum[u];
    if (form == GML_B200_LOGRISE) {
        if (threadIdx.x == 0) f_out[u] = log(s);
        if (want_grad) for (int f = threadIdx.x; f < Fp; f += blockDim.x) G[(int64_t)u * Fp + f] /= s;
    } else if (threadIdx.x == 0) f_out[u] = s;
}

struct BackendCC : EvalBackend {
    const NodeProblem& p;
    DevBuf<float> R;
    DevBuf<double> fsum;
    explicit BackendCC(const NodeProblem& prob) : p(prob) {
        R.alloc((size_t)p.Nn * p.hist->Kp);
        fsum.alloc(p.Nn);
    }
    void eval(const double* x, bool want_grad, double* f_out, double* g_out, cudaStream_t st) override {
        const Histogram& h = *p.hist;
        GML_CUDA(cudaMemsetAsync(fsum.p, 0, sizeof(double) * p.Nn, st));
        dim3 grid((unsigned)(h.Kp / TK), (unsigned)ceil_div(p.Nn, TU));
#define GML_CC_LAUNCH(FORM)                                                                                   \
        if (want_grad) cc_energy_kernel<FORM, true><<<grid, 256, 0, st>>>(p.Q, h.base.p, h.w32.p, p.spin_row.p, x, \
                                                                          p.F, p.Fp, h.Kp, p.Nn, R.p, fsum.p); \
        else cc_energy_kernel<FORM, false><<<grid, 256, 0, st>>>(p.Q, h.base.p, h.w32.p, p.spin_row.p, x,     \
                                                                 p.F, p.Fp, h.Kp, p.Nn, R.p, fsum.p)
        if (p.form == GML_B200_RPLE) { GML_CC_LAUNCH(GML_B200_RPLE); }
        else { GML_CC_LAUNCH(GML_B200_RISE); }
#undef GML_CC_LAUNCH
        GML_LAUNCHED();
        if (want_grad) {
            GML_CUDA(cudaMemsetAsync(g_out, 0, sizeof(double) * p.Nn * p.Fp, st));
            const int64_t tiles = (int64_t)(p.Fp / GF) * ceil_div(p.Nn, GU);
            int64_t splits = std::max<int64_t>(1, std::min<int64_t>(ceil_div(h.Kp, 4096), ceil_div(148 * 8, tiles)));
            const int64_t kchunk = round_up(ceil_div(h.Kp, splits), GK);
            splits = ceil_div(h.Kp, kchunk);
            dim3 gg(p.Fp / GF, (unsigned)ceil_div(p.Nn, GU), (unsigned)splits);
            cc_grad_kernel<<<gg, 256, 0, st>>>(p.Q, R.p, p.Fp, h.Kp, p.Nn, kchunk, g_out);
            GML_LAUNCHED();
        }
        cc_finalize_kernel<<<p.Nn, 128, 0, st>>>(p.form, p.Nn, p.Fp, fsum.p, f_out, g_out, want_grad ? 1 : 0);
        GML_LAUNCHED();
    }
};

}  // namespace

EvalBackend* make_backend_cc(const NodeProblem& p, cudaStream_t) { return new BackendCC(p); }

}  // namespace gml
