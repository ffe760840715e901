// TD pass of the wide states (d a multiple of 16, float streams) on the FP64 tensor-core path (DMMA.8x8x4).
//
// At d = 64 / 256 the critic of mfg_ac2.py:290-344 is two dense contractions over the recorded batch:
//   values   V_n = pi_n^T U pi_n + l . pi_n + b       = row sums of (Pi U) o Pi,     Pi [N, d],  U [d, d] upper triangular
//   gradient G   = sum_n delta_n pi_n pi_n^T (upper triangle)  = (delta o Pi)^T Pi,  K = N transitions
// (mfg_ac2.py:505-514 summed over the batch).  td_delta_kernel / td_gw_kernel walk them with scalar loads (3.2 ms +
// 1.4 ms at d = 64, B = 2^16, T = 16); here both run as m8n8k4 FP64 MMAs:
//   td_values_dmma_kernel : warp = 8 states (staged in shared memory as doubles), loops the 8-column blocks of U and
//                           only the k blocks on or above the diagonal;
//   td_gram_dmma_kernel   : warp = one 16x16 block (I <= J) of G over a K range (split-K), 3 or 4 MMAs per 4 samples,
//                           partial blocks summed in fixed order by td_gram_reduce_kernel (deterministic).
// The linear / bias features and the two scalars stay in td_gw_kernel (quadratic part switched off).
#pragma once
#include "dmfg_rollout2.cuh"

namespace dmfg {

// U[i][j] = w[quad(i, j)] for j >= i, 0 below the diagonal
__global__ void td_unpack_w_kernel(int d, const double* __restrict__ w, double* __restrict__ U) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= d * d) return;
    const int i = idx / d, j = idx - i * d;
    U[idx] = j >= i ? w[quad_index(d, i, j)] : 0.0;
}

constexpr int kTdDmmaThreads = 128;

__global__ void __launch_bounds__(kTdDmmaThreads)
td_values_dmma_kernel(int d, long long N, const float* __restrict__ states, const double* __restrict__ U,
                      const double* __restrict__ w, double* __restrict__ vbuf) {
    extern __shared__ __align__(16) double vsm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const int stride = d + 4;                                   // 4 mod 16 doubles: conflict-free fragment loads
    double* S = vsm + (size_t)wib * 8 * stride;
    const int Q = d * (d + 1) / 2;
    const double* lin = w + Q;
    const double bias = w[Q + d];
    const long long ntiles = (N + 7) / 8;
    const long long wstride = (long long)gridDim.x * (kTdDmmaThreads / 32);
    for (long long tile = (long long)blockIdx.x * (kTdDmmaThreads / 32) + wib; tile < ntiles; tile += wstride) {
        const long long n0 = tile * 8;
        __syncwarp();
        for (int e = lane; e < 8 * d; e += 32) {
            const int r = e / d, c = e - r * d;
            S[r * stride + c] = (n0 + r < N) ? (double)states[(n0 + r) * d + c] : 0.0;
        }
        __syncwarp();
        const double* Sg = S + gid * stride;
        double vth = 0.0;
        for (int cb = 0; cb < d / 8; ++cb) {
            double c0 = 0.0, c1 = 0.0;
            const double* Ub = U + 8 * cb + gid;
#pragma unroll 4
            for (int kb = 0; kb <= 2 * cb + 1; ++kb)             // k = 4 kb .. 4 kb + 3 <= 8 cb + 7: on or above the diagonal
                dmma884(c0, c1, Sg[4 * kb + tig], Ub[(size_t)(4 * kb + tig) * d]);
            const int j0 = 8 * cb + 2 * tig;
            vth = fma(c0 + lin[j0], Sg[j0], vth);
            vth = fma(c1 + lin[j0 + 1], Sg[j0 + 1], vth);
        }
        vth += __shfl_xor_sync(0xffffffffu, vth, 1);
        vth += __shfl_xor_sync(0xffffffffu, vth, 2);
        if (tig == 0 && n0 + gid < N) vbuf[n0 + gid] = vth + bias;
    }
}

// delta[t][b] = r + gfac V(pi_{t+1}) - V(pi_t)   (mfg_ac2.py:505; ac_irl.py:691 with the running discount)
__global__ void td_delta_from_values_kernel(int T, long long B, double gamma, int discount_kind,
                                            const float* __restrict__ rewards, const double* __restrict__ vbuf,
                                            float* __restrict__ deltas, double* __restrict__ delta_buf) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long long)T * B) return;
    const int t = (int)(n / B);
    double gfac = gamma;
    if (discount_kind != DMFG_DISCOUNT_STEP) {
        gfac = 1.0;
        for (int k = 0; k < t; ++k) gfac *= gamma;
    }
    const double delta = (double)rewards[n] + gfac * vbuf[n + B] - vbuf[n];
    if (deltas != nullptr) deltas[n] = (float)delta;
    delta_buf[n] = delta;
}

// block (I <= J) of the Gram matrix over the K range of this warp; partials [ksplit][nblk][16][16]
__global__ void __launch_bounds__(kTdDmmaThreads)
td_gram_dmma_kernel(int d, long long Nt, const float* __restrict__ states, const double* __restrict__ delta,
                    int nblk, int ksplit, long long kchunk, double* __restrict__ partials) {
    const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    const long long wid = (long long)blockIdx.x * (kTdDmmaThreads / 32) + (threadIdx.x >> 5);
    if (wid >= (long long)nblk * ksplit) return;
    const int blk = (int)(wid % nblk), ks = (int)(wid / nblk);
    const int nb = d / 16;
    int I = 0, rem = blk;
    while (rem >= nb - I) { rem -= nb - I; ++I; }
    const int J = I + rem;
    const long long k_begin = (long long)ks * kchunk;
    long long k_end = k_begin + kchunk;
    if (k_end > Nt) k_end = Nt;
    double c00a = 0, c00b = 0, c01a = 0, c01b = 0, c10a = 0, c10b = 0, c11a = 0, c11b = 0;
    const bool offdiag = I != J;
#pragma unroll 4
    for (long long k0 = k_begin; k0 < k_end; k0 += 4) {
        const long long n = k0 + tig;
        double dl = 0.0, a_lo = 0.0, a_hi = 0.0, b_lo = 0.0, b_hi = 0.0;
        if (n < k_end) {
            const float* s = states + n * d;
            dl = delta[n];
            b_lo = (double)s[16 * J + gid];
            b_hi = (double)s[16 * J + 8 + gid];
            a_lo = dl * (double)s[16 * I + gid];
            a_hi = dl * (double)s[16 * I + 8 + gid];
        }
        dmma884(c00a, c00b, a_lo, b_lo);
        dmma884(c01a, c01b, a_lo, b_hi);
        dmma884(c11a, c11b, a_hi, b_hi);
        if (offdiag) dmma884(c10a, c10b, a_hi, b_lo);            // below the diagonal inside a diagonal block: not a feature
    }
    double* out = partials + ((size_t)ks * nblk + blk) * 256;
    out[gid * 16 + 2 * tig] = c00a;          out[gid * 16 + 2 * tig + 1] = c00b;
    out[gid * 16 + 8 + 2 * tig] = c01a;      out[gid * 16 + 8 + 2 * tig + 1] = c01b;
    out[(8 + gid) * 16 + 2 * tig] = c10a;    out[(8 + gid) * 16 + 2 * tig + 1] = c10b;
    out[(8 + gid) * 16 + 8 + 2 * tig] = c11a; out[(8 + gid) * 16 + 8 + 2 * tig + 1] = c11b;
}

// acc[1 + f] for the quadratic features f < Q: sum of the partial blocks over the K splits, in split order
__global__ void td_gram_reduce_kernel(int d, int nblk, int ksplit, const double* __restrict__ partials,
                                      double* __restrict__ acc) {
    const int Q = d * (d + 1) / 2;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= Q) return;
    int i = 0, rem = f;
    while (rem >= d - i) { rem -= d - i; ++i; }
    const int j = i + rem;
    const int nb = d / 16, I = i / 16, J = j / 16;
    const int blk = I * nb - (I * (I - 1)) / 2 + (J - I);
    const double* src = partials + (size_t)blk * 256 + (i % 16) * 16 + (j % 16);
    double s = 0.0;
    for (int ks = 0; ks < ksplit; ++ks) s += src[(size_t)ks * nblk * 256];
    acc[1 + f] = s;
}

}  // namespace dmfg
