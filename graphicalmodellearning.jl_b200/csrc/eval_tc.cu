// K2/K3, tensor-core backend: the two dense contractions of an objective/gradient pass run on the
// sm_100a 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM, operands staged
// by TMA into 128B-swizzled shared memory), in EXACT fixed-point arithmetic:
//
//   * the iterate lives on the lattice 2^-24 (the FISTA driver snaps to it), so x = q * 2^-24 with
//     |q| < 2^27, split into 4 balanced base-128 int8 limbs.  The energy contraction
//         E[k,u] = sum_f S[k,f] * x_u[f]            (S = +-1 features, int8)
//     is then four int8 GEMM column groups with int32 accumulation -- no rounding at all.
//   * the epilogue turns E into t = s_u E, psi = exp(-t) (or the RPLE logistic terms) in fp32, and
//     quantises  r = s_u w psi  to nR balanced int8 limbs with a per-node scale derived from the
//     bound |t| <= |x_u|_1; the objective term is accumulated as an int64 sum of the same grid.
//   * the gradient contraction  G[u,f] = -sum_k r[u,k] S[k,f]  is again an int8 GEMM, split over
//     sample ranges; partial tiles are combined with int64 atomics, so the result is independent
//     of the reduction order (bitwise reproducible).
//
// Replaces the per-node JuMP expression evaluation of src/GraphicalModelLearning.jl:162-172 (and
// :271-281, :309-319, :106-119) for all nodes at once.
#include <cuda.h>

#include <cmath>

#include "common.cuh"

namespace gml {
namespace {

// ------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32, M = 128, K = 32 per instruction
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major operand tile, rows of 128 bytes, 128B swizzle (what TMA SWIZZLE_128B writes):
// start address >> 4, LBO = 1 (ignored for swizzled K-major), SBO = 1024 B (8 rows), version 1, layout 2
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor, kind::i8: D = s32, A = B = s8, both K-major
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// fixed-point helpers
// ------------------------------------------------------------------------------------------
constexpr double X_LATTICE = 1.0 / 16777216.0;   // 2^-24
constexpr int X_LIMBS = 4;
constexpr int NODE_TILE1 = 64;                   // nodes per energy tile (4 limbs -> N = 256)
constexpr int NODE_TILE2 = 128;                  // nodes per gradient tile (M = 128)
constexpr int QF_MAX = 1 << 20;                  // objective-term grid

__host__ __device__ constexpr int r_qmax(int nR) { return nR == 2 ? 8000 : (nR == 3 ? 1020000 : 130000000); }

__device__ __forceinline__ int balanced_digit(int& q) {   // returns q mod 128 in [-64, 63], q <- (q - d) / 128
    const int d = ((q + 64) & 127) - 64;
    q = (q - d) >> 7;
    return d;
}

struct NodeScale { float inv_df, inv_dr; };      // 1/deltaF, 1/deltaR per node (device)

// x [Nn x Fp] (double, on the lattice) -> limb tiles X4 [(tile*4 + limb)*64 + i][Fp] and per-node scales
__global__ void __launch_bounds__(128) tc_quantize_x_kernel(const double* __restrict__ x, int Nn, int Fp, int form, double wmax,
                                                           int nR, int8_t* __restrict__ X4, NodeScale* __restrict__ scale,
                                                           double* __restrict__ delta /* [2*Nn_pad]: dF, dR */, int* __restrict__ flags) {
    const int u = blockIdx.x;
    const int tile = u / NODE_TILE1, i = u % NODE_TILE1;
    __shared__ double red[4];
    double l1 = 0.0;
    for (int f = threadIdx.x; f < Fp; f += blockDim.x) {
        const double v = (u < Nn) ? x[(int64_t)u * Fp + f] : 0.0;
        l1 += fabs(v);
        long long ql = llrint(v * 16777216.0);
        if (ql > 134000000LL || ql < -134000000LL) { atomicOr(flags, 1); ql = ql > 0 ? 134000000LL : -134000000LL; }
        int q = (int)ql;
        const int d3 = balanced_digit(q), d2 = balanced_digit(q), d1 = balanced_digit(q), d0 = q;
        const int64_t row = ((int64_t)tile * X_LIMBS) * NODE_TILE1 + i;
        X4[(row + 0 * NODE_TILE1) * Fp + f] = (int8_t)d0;
        X4[(row + 1 * NODE_TILE1) * Fp + f] = (int8_t)d1;
        X4[(row + 2 * NODE_TILE1) * Fp + f] = (int8_t)d2;
        X4[(row + 3 * NODE_TILE1) * Fp + f] = (int8_t)d3;
    }
    for (int o = 16; o; o >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l1;
    __syncthreads();
    if (threadIdx.x == 0) {
        const double B = red[0] + red[1] + red[2] + red[3];   // |t| <= B for every sample
        double dF, dR;
        if (form == GML_B200_RPLE) {
            dF = wmax * (2.0 * B + 0.6931471805599453) * 1.000001 / QF_MAX;
            dR = 2.0 * wmax * 1.000001 / r_qmax(nR);
        } else {
            const double top = wmax * exp(fmin(B, 80.0)) * 1.000001;
            dF = top / QF_MAX;
            dR = top / r_qmax(nR);
        }
        scale[u].inv_df = (float)(1.0 / dF);
        scale[u].inv_dr = (float)(1.0 / dR);
        // the exact reciprocal of what the epilogue multiplies with, so that value = q * delta holds
        delta[2 * u] = 1.0 / (double)scale[u].inv_df;
        delta[2 * u + 1] = 1.0 / (double)scale[u].inv_dr;
    }
}

// ------------------------------------------------------------------------------------------
// GEMM-1: energies + fused epilogue
// ------------------------------------------------------------------------------------------
struct EnergyParams {
    int64_t Kp;
    int Fp, Nn, n_tiles, node_begin_row;   // node tile t covers base rows node_begin_row + 64 t ...
    int64_t sample_blocks;                 // Kp / 128
    int64_t r_rows_per_limb;               // Nn_pad2
    int nR, form;
    const float* w32;
    const NodeScale* scale;
    double* fsum;                          // [Nn_pad1] objective sums
};

constexpr int E_STAGES = 3;
constexpr int E_A_BYTES = 128 * 128, E_B_BYTES = 256 * 128, E_STAGE_BYTES = E_A_BYTES + E_B_BYTES;
constexpr int E_S_BYTES = NODE_TILE1 * 128;          // spins tile of the node block
constexpr int E_R_BYTES_PER_LIMB = NODE_TILE1 * 128; // staging for the R limbs
constexpr int E_SMEM = E_STAGES * E_STAGE_BYTES + 2 * E_S_BYTES + 4 * E_R_BYTES_PER_LIMB + 1024 /*align*/ + 256 /*barriers*/;

template <int FORM, bool GRAD>
__global__ void __launch_bounds__(192, 1) tc_energy_kernel(const __grid_constant__ CUtensorMap tmA,   // P  [Kp x Fp]
                                                          const __grid_constant__ CUtensorMap tmB,   // X4 [tiles*256 x Fp]
                                                          const __grid_constant__ CUtensorMap tmS,   // base [Fb x Kp], box 64 rows
                                                          const __grid_constant__ CUtensorMap tmR,   // R  [nR*Nn_pad2 x Kp], box 64 rows
                                                          EnergyParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_stage = smem;
    uint8_t* s_spin = smem + E_STAGES * E_STAGE_BYTES;
    uint8_t* s_r = s_spin + 2 * E_S_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_r + 4 * E_R_BYTES_PER_LIMB);
    uint64_t* full = bars;                    // [E_STAGES]
    uint64_t* empty = bars + E_STAGES;        // [E_STAGES]
    uint64_t* tfull = bars + 2 * E_STAGES;    // [2]
    uint64_t* tempty = tfull + 2;             // [2]
    uint64_t* sfull = tempty + 2;             // [2]
    uint64_t* sempty = sfull + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = p.Fp / 128;
    const int64_t total = (int64_t)p.n_tiles * p.sample_blocks;
    const int64_t per = total / gridDim.x, rem = total % gridDim.x;
    const int64_t u_begin = (int64_t)blockIdx.x * per + ((int64_t)blockIdx.x < rem ? (int64_t)blockIdx.x : rem);
    const int64_t u_end = u_begin + per + (blockIdx.x < rem ? 1 : 0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < E_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); mbar_init(&sfull[i], 1); mbar_init(&sempty[i], 4); }
        fence_barrier_init();
        prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmS);
        if (GRAD) prefetch_tmap(&tmR);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int slot = 0; uint32_t sphase = 0;
            for (int64_t idx = u_begin; idx < u_end; ++idx) {
                const int nt = (int)(idx / p.sample_blocks);
                const int64_t sb = idx % p.sample_blocks;
                mbar_wait(&sempty[slot], sphase ^ 1);
                mbar_expect_tx(&sfull[slot], E_S_BYTES);
                tma_load_2d(s_spin + slot * E_S_BYTES, &tmS, &sfull[slot], (int)(sb * 128), p.node_begin_row + nt * NODE_TILE1);
                if (++slot == 2) { slot = 0; sphase ^= 1; }
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], E_STAGE_BYTES);
                    uint8_t* a = s_stage + stage * E_STAGE_BYTES;
                    tma_load_2d(a, &tmA, &full[stage], kb * 128, (int)(sb * 128));
                    tma_load_2d(a + E_A_BYTES, &tmB, &full[stage], kb * 128, nt * 256);
                    if (++stage == E_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_i8(128, 256);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int64_t idx = u_begin; idx < u_end; ++idx) {
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + as * 256;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(s_stage + stage * E_STAGE_BYTES);
                    const uint64_t da = make_kmajor_desc(a_addr), db = make_kmajor_desc(a_addr + E_A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_i8(d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&empty[stage]);
                    if (++stage == E_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ================= epilogue warps (2..5): TMEM lane quarter = warp % 4 =================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;            // sample within the block == TMEM lane
        const int et = threadIdx.x - 64;                // 0..127
        // objective terms: fp32 per-thread partial sums over at most F_FLUSH sample blocks, then fp64
        float facc[NODE_TILE1];
#pragma unroll
        for (int i = 0; i < NODE_TILE1; ++i) facc[i] = 0.f;
        int as = 0; uint32_t aphase = 0;
        int slot = 0; uint32_t sphase = 0;
        int cur_nt = -1, since_flush = 0;
        auto flush = [&](int nt) {
#pragma unroll
            for (int i = 0; i < NODE_TILE1; ++i) {
                double v = (double)facc[i];
                for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) atomicAdd(p.fsum + (int64_t)nt * NODE_TILE1 + i, v);
                facc[i] = 0.f;
            }
        };
        for (int64_t idx = u_begin; idx < u_end; ++idx) {
            const int nt = (int)(idx / p.sample_blocks);
            const int64_t sb = idx % p.sample_blocks;
            if (nt != cur_nt || since_flush >= 32) {
                if (cur_nt >= 0) flush(cur_nt);
                cur_nt = nt; since_flush = 0;
            }
            ++since_flush;
            const float wk = p.w32[sb * 128 + row];
            mbar_wait(&sfull[slot], sphase);
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            if (GRAD) {
                if (et == 0) tma_store_wait_read();   // previous tile's stores have drained the staging buffer
                named_bar_sync(1, 128);
            }
            const uint8_t* spin = s_spin + slot * E_S_BYTES;
            const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * 256;
#pragma unroll
            for (int c = 0; c < NODE_TILE1 / 16; ++c) {
                int32_t a0[16], a1[16], a2[16], a3[16];
                tmem_ld16(tbase + 0 * NODE_TILE1 + c * 16, a0);
                tmem_ld16(tbase + 1 * NODE_TILE1 + c * 16, a1);
                tmem_ld16(tbase + 2 * NODE_TILE1 + c * 16, a2);
                tmem_ld16(tbase + 3 * NODE_TILE1 + c * 16, a3);
                tmem_ld_wait();
                if (c == NODE_TILE1 / 16 - 1) {       // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[as]);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int node_in_tile = c * 16 + i;
                    const int hi = a0[i] * 128 + a1[i], lo = a2[i] * 128 + a3[i];
                    const float e = fmaf((float)hi, 16384.f, (float)lo) * (float)X_LATTICE;
                    const float su = (float)(int8_t)spin[node_in_tile * 128 + row];
                    const float t = su * e;
                    const NodeScale sc = p.scale[nt * NODE_TILE1 + node_in_tile];
                    float fterm, gterm;
                    if (FORM == GML_B200_RPLE) {
                        const float a = -2.f * t;
                        const float ex = exp2f(-fabsf(a) * 1.4426950408889634f);
                        fterm = wk * (fmaxf(a, 0.f) + log1pf(ex));
                        gterm = 2.f * wk * (a > 0.f ? 1.f : ex) / (1.f + ex);       // 2 w sigma(-2t)
                    } else {
                        const float psi = exp2f(fminf(-t, 80.f) * 1.4426950408889634f);
                        fterm = wk * psi; gterm = fterm;
                    }
                    facc[node_in_tile] += fterm;
                    if (GRAD) {
                        int q = __float2int_rn(su * gterm * sc.inv_dr);
                        const int d_lo = balanced_digit(q);
                        if (p.nR == 2) {
                            s_r[(0 * NODE_TILE1 + node_in_tile) * 128 + row] = (uint8_t)(int8_t)q;
                            s_r[(1 * NODE_TILE1 + node_in_tile) * 128 + row] = (uint8_t)(int8_t)d_lo;
                        } else {
                            const int d_mid = balanced_digit(q);
                            if (p.nR == 3) {
                                s_r[(0 * NODE_TILE1 + node_in_tile) * 128 + row] = (uint8_t)(int8_t)q;
                                s_r[(1 * NODE_TILE1 + node_in_tile) * 128 + row] = (uint8_t)(int8_t)d_mid;
                                s_r[(2 * NODE_TILE1 + node_in_tile) * 128 + row] = (uint8_t)(int8_t)d_lo;
                            } else {
                                const int d_2 = balanced_digit(q);
                                s_r[(0 * NODE_TILE1 + node_in_tile) * 128 + row] = (uint8_t)(int8_t)q;
                                s_r[(1 * NODE_TILE1 + node_in_tile) * 128 + row] = (uint8_t)(int8_t)d_2;
                                s_r[(2 * NODE_TILE1 + node_in_tile) * 128 + row] = (uint8_t)(int8_t)d_mid;
                                s_r[(3 * NODE_TILE1 + node_in_tile) * 128 + row] = (uint8_t)(int8_t)d_lo;
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sempty[slot]);
            if (GRAD) {
                fence_proxy_async();
                named_bar_sync(1, 128);
                if (et == 0) {
                    for (int j = 0; j < p.nR; ++j)
                        tma_store_2d(&tmR, s_r + j * E_R_BYTES_PER_LIMB, (int)(sb * 128),
                                     (int)(j * p.r_rows_per_limb) + nt * NODE_TILE1);
                    tma_store_commit();
                }
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
            if (++slot == 2) { slot = 0; sphase ^= 1; }
        }
        if (cur_nt >= 0) flush(cur_nt);
        if (GRAD && et == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// GEMM-2: gradient contraction, split over sample ranges
// ------------------------------------------------------------------------------------------
struct GradParams {
    int Fp, m_tiles, f_tiles, nR;
    int64_t r_rows_per_limb;      // Nn_pad2
    int64_t chunks, chunk_blocks, sample_blocks;
    long long* G;                 // [Nn_pad2 x Fp] int64
};

constexpr int G_TILE_BYTES = 128 * 128;
__host__ __device__ constexpr int g_stages(int nr) { return nr >= 4 ? 2 : 3; }

template <int NR>
__global__ void __launch_bounds__(192, 1) tc_grad_kernel(const __grid_constant__ CUtensorMap tmRa,   // R [nR*Nn_pad2 x Kp], box 128 rows, SW128
                                                        const __grid_constant__ CUtensorMap tmQ,    // Q [Fp x Kp], box 128 rows, SW128
                                                        GradParams p) {
    constexpr int STAGE_BYTES = (NR + 1) * G_TILE_BYTES;
    constexpr int G_STAGES = g_stages(NR);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + G_STAGES;
    uint64_t* tfull = bars + 2 * G_STAGES;
    uint64_t* tempty = tfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int64_t tiles = (int64_t)p.m_tiles * p.f_tiles;
    const int64_t total = tiles * p.chunks;
    const int64_t per = total / gridDim.x, rem = total % gridDim.x;
    const int64_t u_begin = (int64_t)blockIdx.x * per + ((int64_t)blockIdx.x < rem ? (int64_t)blockIdx.x : rem);
    const int64_t u_end = u_begin + per + (blockIdx.x < rem ? 1 : 0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < G_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tfull, 1); mbar_init(tempty, 4);
        fence_barrier_init();
        prefetch_tmap(&tmRa); prefetch_tmap(&tmQ);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // unit -> (chunk, m tile, f tile); chunk-major so that concurrently running CTAs share operand tiles in L2
    auto decode = [&](int64_t idx, int64_t& chunk, int& mt, int& ft) {
        chunk = idx / tiles;
        const int r = (int)(idx % tiles);
        mt = r / p.f_tiles; ft = r % p.f_tiles;
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t idx = u_begin; idx < u_end; ++idx) {
                int64_t chunk; int mt, ft;
                decode(idx, chunk, mt, ft);
                const int64_t b0 = chunk * p.chunk_blocks, b1 = min(b0 + p.chunk_blocks, p.sample_blocks);
                for (int64_t b = b0; b < b1; ++b) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    uint8_t* s = smem + stage * STAGE_BYTES;
#pragma unroll
                    for (int j = 0; j < NR; ++j)
                        tma_load_2d(s + j * G_TILE_BYTES, &tmRa, &full[stage], (int)(b * 128), (int)(j * p.r_rows_per_limb) + mt * 128);
                    tma_load_2d(s + NR * G_TILE_BYTES, &tmQ, &full[stage], (int)(b * 128), ft * 128);
                    if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_i8(128, 128);
            int stage = 0; uint32_t phase = 0, aphase = 0;
            for (int64_t idx = u_begin; idx < u_end; ++idx) {
                int64_t chunk; int mt, ft;
                decode(idx, chunk, mt, ft);
                const int64_t b0 = chunk * p.chunk_blocks, b1 = min(b0 + p.chunk_blocks, p.sample_blocks);
                mbar_wait(tempty, aphase ^ 1);
                tc_fence_after();
                for (int64_t b = b0; b < b1; ++b) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t s_addr = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t db = make_kmajor_desc(s_addr + NR * G_TILE_BYTES);
#pragma unroll
                    for (int j = 0; j < NR; ++j) {
                        const uint64_t da = make_kmajor_desc(s_addr + j * G_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_i8(tmem_base + j * 128, da + 2 * k, db + 2 * k, idesc, (b > b0 || k) ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull);
                aphase ^= 1;
            }
        }
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        uint32_t aphase = 0;
        for (int64_t idx = u_begin; idx < u_end; ++idx) {
            int64_t chunk; int mt, ft;
            decode(idx, chunk, mt, ft);
            mbar_wait(tfull, aphase);
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16);
            long long* dst = p.G + ((int64_t)mt * 128 + row) * p.Fp + ft * 128;
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                int32_t acc[NR][16];
#pragma unroll
                for (int j = 0; j < NR; ++j) tmem_ld16(tbase + j * 128 + c * 16, acc[j]);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    long long v = 0;
#pragma unroll
                    for (int j = 0; j < NR; ++j) v = v * 128 + (long long)acc[j][i];
                    if (v != 0) atomicAdd(reinterpret_cast<unsigned long long*>(dst + c * 16 + i), (unsigned long long)v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
            aphase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// int64 sums -> double objective / gradient (and the logRISE normalisation, :279)
__global__ void tc_finalize_kernel(int form, int Nn, int Fp, const double* __restrict__ fsum,
                                   const long long* __restrict__ G64, const double* __restrict__ delta,
                                   double* __restrict__ f_out, double* __restrict__ g_out, int want_grad) {
    const int u = blockIdx.x;
    const double fs = fsum[u];
    if (threadIdx.x == 0) f_out[u] = (form == GML_B200_LOGRISE) ? log(fs) : fs;
    if (!want_grad) return;
    const double sc = -delta[2 * u + 1] / (form == GML_B200_LOGRISE ? fs : 1.0);
    for (int f = threadIdx.x; f < Fp; f += blockDim.x) g_out[(int64_t)u * Fp + f] = sc * (double)G64[(int64_t)u * Fp + f];
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
CUtensorMap make_map_2d(const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer,
                        CUtensorMapSwizzle swz) {
    CUtensorMap m;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {inner};   // bytes (int8 elements)
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    // resolved through the runtime so that libgml_b200.so does not link libcuda (it must still load on a
    // box without a driver, where every compute entry point reports GML_B200_ECUDA)
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        GML_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        GML_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
        encode = reinterpret_cast<encode_fn>(fn);
    }
    const CUresult rc = encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                                               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
        throw CudaError{GML_B200_ECUDA};
    }
    return m;
}

struct BackendTC : EvalBackend {
    const NodeProblem& p;
    int Nn_pad1, Nn_pad2, nR, n_sms;
    int32_t first_row = 0;   // spin rows of this shard are contiguous in `base`: spin_row[u] = first_row + u
    DevBuf<int8_t> X4, R;
    DevBuf<NodeScale> scale;
    DevBuf<double> delta;
    DevBuf<double> fsum;
    DevBuf<long long> G64;
    DevBuf<int> flags;
    const int8_t* P;
    CUtensorMap tmA, tmB, tmS, tmR, tmRa, tmQ;

    BackendTC(const NodeProblem& prob, cudaStream_t st) : p(prob) {
        Histogram& h = *p.hist;
        Nn_pad1 = (int)round_up(p.Nn, NODE_TILE1);
        Nn_pad2 = (int)round_up(p.Nn, NODE_TILE2);
        // Residual limbs: the rounding noise of the gradient is ~0.3 sqrt(K) wmax e^B / qmax(nR).  3 limbs
        // (qmax 1e6) keep it below 1e-8 for near-uniform counts; strongly weighted histograms get 4.
        nR = (std::sqrt((double)h.K) * h.wmax > 2e-3) ? 4 : 3;
        int dev = 0;
        GML_CUDA(cudaGetDevice(&dev));
        GML_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
        P = ensure_P(h, p.Q, p.Fp, st);
        X4.alloc((size_t)Nn_pad1 * X_LIMBS * p.Fp);
        R.alloc((size_t)nR * Nn_pad2 * h.Kp);
        scale.alloc(Nn_pad1); delta.alloc(2 * (size_t)Nn_pad1);
        fsum.alloc(Nn_pad1);
        G64.alloc((size_t)Nn_pad2 * p.Fp);
        flags.alloc(1);
        GML_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(int), st));
        // rows of R that belong to padding nodes are never written by GEMM-1 tiles beyond Nn_pad1: clear once
        GML_CUDA(cudaMemsetAsync(R.p, 0, (size_t)nR * Nn_pad2 * h.Kp, st));
        tmA = make_map_2d(P, p.Fp, h.Kp, 128, 128, CU_TENSOR_MAP_SWIZZLE_128B);
        tmB = make_map_2d(X4.p, p.Fp, (uint64_t)Nn_pad1 * X_LIMBS, 128, 256, CU_TENSOR_MAP_SWIZZLE_128B);
        tmS = make_map_2d(h.base.p, h.Kp, h.Fb, 128, NODE_TILE1, CU_TENSOR_MAP_SWIZZLE_NONE);
        tmR = make_map_2d(R.p, h.Kp, (uint64_t)nR * Nn_pad2, 128, NODE_TILE1, CU_TENSOR_MAP_SWIZZLE_NONE);
        tmRa = make_map_2d(R.p, h.Kp, (uint64_t)nR * Nn_pad2, 128, 128, CU_TENSOR_MAP_SWIZZLE_128B);
        tmQ = make_map_2d(p.Q, h.Kp, p.Fp, 128, 128, CU_TENSOR_MAP_SWIZZLE_128B);
        GML_CUDA(cudaMemcpyAsync(&first_row, p.spin_row.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        GML_CUDA(cudaStreamSynchronize(st));
        configure();
    }

    template <class K> static void set_smem(K kernel, int bytes) {
        GML_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    }
    void configure() {
        set_smem(tc_energy_kernel<GML_B200_RISE, true>, E_SMEM);
        set_smem(tc_energy_kernel<GML_B200_RISE, false>, E_SMEM);
        set_smem(tc_energy_kernel<GML_B200_RPLE, true>, E_SMEM);
        set_smem(tc_energy_kernel<GML_B200_RPLE, false>, E_SMEM);
        set_smem(tc_grad_kernel<3>, grad_smem(3));
        set_smem(tc_grad_kernel<4>, grad_smem(4));
    }
    static int grad_smem(int nr) { return g_stages(nr) * (nr + 1) * G_TILE_BYTES + 1024 + 128; }

    double lattice() const override { return X_LATTICE; }

    void eval(const double* x, bool want_grad, double* f_out, double* g_out, cudaStream_t st) override {
        const Histogram& h = *p.hist;
        tc_quantize_x_kernel<<<Nn_pad1, 128, 0, st>>>(x, p.Nn, p.Fp, p.form, h.wmax, nR, X4.p, scale.p, delta.p, flags.p);
        GML_LAUNCHED();
        GML_CUDA(cudaMemsetAsync(fsum.p, 0, sizeof(double) * Nn_pad1, st));
        EnergyParams ep{};
        ep.Kp = h.Kp; ep.Fp = p.Fp; ep.Nn = p.Nn; ep.n_tiles = Nn_pad1 / NODE_TILE1;
        ep.sample_blocks = h.Kp / 128; ep.r_rows_per_limb = Nn_pad2; ep.nR = nR; ep.form = p.form;
        ep.w32 = h.w32.p; ep.scale = scale.p; ep.fsum = fsum.p;
        ep.node_begin_row = first_row;
        const int64_t units = (int64_t)ep.n_tiles * ep.sample_blocks;
        const int grid1 = (int)std::min<int64_t>(units, n_sms);
        const bool rple = p.form == GML_B200_RPLE;
        if (want_grad) {
            if (rple) tc_energy_kernel<GML_B200_RPLE, true><<<grid1, 192, E_SMEM, st>>>(tmA, tmB, tmS, tmR, ep);
            else tc_energy_kernel<GML_B200_RISE, true><<<grid1, 192, E_SMEM, st>>>(tmA, tmB, tmS, tmR, ep);
        } else {
            if (rple) tc_energy_kernel<GML_B200_RPLE, false><<<grid1, 192, E_SMEM, st>>>(tmA, tmB, tmS, tmR, ep);
            else tc_energy_kernel<GML_B200_RISE, false><<<grid1, 192, E_SMEM, st>>>(tmA, tmB, tmS, tmR, ep);
        }
        GML_LAUNCHED();
        if (want_grad) {
            GML_CUDA(cudaMemsetAsync(G64.p, 0, sizeof(long long) * Nn_pad2 * p.Fp, st));
            GradParams gp{};
            gp.Fp = p.Fp; gp.m_tiles = Nn_pad2 / 128; gp.f_tiles = p.Fp / 128; gp.nR = nR;
            gp.r_rows_per_limb = Nn_pad2; gp.sample_blocks = h.Kp / 128; gp.G = G64.p;
            const int64_t tiles = (int64_t)gp.m_tiles * gp.f_tiles;
            // split the sample axis so that there are ~4 units per SM, at most 2^17 samples per unit
            int64_t chunks = std::max<int64_t>(1, ceil_div((int64_t)n_sms * 4, tiles));
            chunks = std::min<int64_t>(chunks, gp.sample_blocks);
            chunks = std::max<int64_t>(chunks, ceil_div(gp.sample_blocks, 1024));
            gp.chunk_blocks = ceil_div(gp.sample_blocks, chunks);
            gp.chunks = ceil_div(gp.sample_blocks, gp.chunk_blocks);
            const int grid2 = (int)std::min<int64_t>(tiles * gp.chunks, n_sms);
            if (nR == 3) tc_grad_kernel<3><<<grid2, 192, grad_smem(3), st>>>(tmRa, tmQ, gp);
            else tc_grad_kernel<4><<<grid2, 192, grad_smem(4), st>>>(tmRa, tmQ, gp);
            GML_LAUNCHED();
        }
        tc_finalize_kernel<<<p.Nn, 128, 0, st>>>(p.form, p.Nn, p.Fp, fsum.p, G64.p, delta.p, f_out, g_out, want_grad ? 1 : 0);
        GML_LAUNCHED();
    }
};

}  // namespace

EvalBackend* make_backend_tc(const NodeProblem& p, cudaStream_t st) { return new BackendTC(p, st); }

}  // namespace gml
