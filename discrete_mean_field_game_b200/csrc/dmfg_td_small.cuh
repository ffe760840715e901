// TD errors from a recorded batch at small d (15 / 16, float streams): V(pi_t) for t = 0..T and
// delta_t = r_t + gamma_t V(pi_{t+1}) - V(pi_t) (mfg_ac2.py:505 / ac_irl.py:691) with ONE THREAD per population.
//
// td_delta_kernel (dmfg_rollout.cuh) is the any-d form: a warp per population, lane = row of the quadratic form, weights
// read from global memory -- at d = 15 half the lanes idle and every lane walks a serial chain of global loads: 6.2 ms for
// the 2^20 x 17 states of a config-5 step (173 GB/s of a 1.07 GB record).  Here the 136 critic weights sit in shared
// memory (every access is a broadcast with an immediate address: the loops are compile-time), a thread keeps its state in
// registers and evaluates the upper-triangular quadratic form row by row in double; consecutive threads read consecutive
// 60-byte states, so the record is still read exactly once through L1.
#pragma once
#include "dmfg_rollout.cuh"

namespace dmfg {

constexpr int kTdSmallThreads = 128;

template <int D>
__global__ void __launch_bounds__(kTdSmallThreads) td_delta_small_kernel(const TdParams<float> p) {
    constexpr int Q = D * (D + 1) / 2, F = Q + D + 1;
    __shared__ double ws[F];
    for (int f = threadIdx.x; f < F; f += kTdSmallThreads) ws[f] = p.w[f];
    __syncthreads();
    const long long b = (long long)blockIdx.x * kTdSmallThreads + threadIdx.x;
    if (b >= p.B) return;
    // the weights are read where they are used, with explicit (volatile) shared loads: left to itself the compiler hoists
    // all 136 loop-invariant doubles out of the step loop into registers and spills 1.8 KB per thread
    const uint32_t wa = (uint32_t)__cvta_generic_to_shared(ws);
    auto W = [&](int idx) -> double {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(wa + 8u * (uint32_t)idx));
        return v;
    };
    double v_prev = 0.0, disc = 1.0;
#pragma unroll 1
    for (int t = 0; t <= p.T; ++t) {
        const float* s = p.states + ((long long)t * p.B + b) * D;
        double x[D];
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = (double)s[k];
        // V = sum_i x_i (w_lin_i + sum_{j >= i} w_ij x_j) + bias, two independent chains over the rows
        double v0 = W(Q + D), v1 = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const int base = quad_index(D, i, i);
            double a = W(Q + i);
#pragma unroll
            for (int j = i; j < D; ++j) a = fma(W(base + (j - i)), x[j], a);
            if (i & 1) v1 = fma(a, x[i], v1); else v0 = fma(a, x[i], v0);
        }
        const double v = v0 + v1;
        if (t > 0) {
            const long long tb = (long long)(t - 1) * p.B + b;
            const double gfac = p.discount_kind == DMFG_DISCOUNT_STEP ? p.gamma : disc;
            const double delta = (double)p.rewards[tb] + gfac * v - v_prev;
            if (p.deltas) p.deltas[tb] = (float)delta;
            if (p.delta_buf) p.delta_buf[tb] = delta;
            disc *= p.gamma;
        }
        v_prev = v;
    }
}

}  // namespace dmfg
