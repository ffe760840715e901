// rollout_v2_kernel -- the throughput version of the fused population step for d = 15 / 16, float
// streams (sm_100a).  Same arithmetic contract as rollout_fast_kernel (dmfg_rollout.cuh), restructured
// around what ncu showed on the first version (profiles/r1_rollout_fast_train_ncu_summary.md: 2.8 k warp
// instructions per population-step, a fully unrolled 120 KB body thrashing the instruction cache, 21 % of
// the XU pipe spent on float<->double conversions and precise libm calls):
//
//   * a group of 16 lanes owns one population, lane r owns ROW r of P (as before), but the row is walked
//     by two ROLLED loops, so the hot body is a few KB:
//       pass 1 (per column pair): alpha, alpha' (one ex2, one lg2, one rcp), psi(alpha) (one rcp, one lg2),
//               the Gamma pair (one Philox call, Box-Muller on MUFU, squeeze-accepted Marsaglia-Tsang);
//               {y, alpha'} parked in an 8-byte shared slot, row sum of y in double;
//       pass 2 (per column, after the row sum is known): P = y / s, ln P, the reward term P^2 (pi_j - pi_i)
//               and the flux pi_i P_ij in double, written back INTO the same slot;
//   * pi' = P^T pi is a transposed read of those slots (15 LDS.64 + 15 DADD per lane) instead of 60
//     shuffles; the state vector lives in a double-buffered shared array (double + float copies) that
//     all lanes of the group read as broadcasts;
//   * float everywhere except where the TD error needs it: row sums, P, reward, flux, state, critic
//     value and all reductions stay double, every float -> double conversion that
//     remains is per element of pass 2 or per row;
//   * critic weights are staged once per CTA in a [slot][lane] table shared by all groups (2 KB instead
//     of 35 KB), so three CTAs fit an SM.
#pragma once
#include "dmfg_rollout.cuh"

namespace dmfg {

constexpr int kV2Threads = 256;
constexpr int kV2G = 16;
constexpr int kV2Slots = 17;          // 8-byte slots per lane row (16 columns + 1 pad => odd stride)

template <int D>
struct V2Smem {
    static constexpr int GPB = kV2Threads / kV2G;
    static constexpr int NSLOT = D + 2;
    // offsets in doubles
    static constexpr int tile = 0;                                  // [GPB][16][17] slots
    static constexpr int pid = tile + GPB * kV2G * kV2Slots;        // [2][GPB][16] state, double
    static constexpr int wl = pid + 2 * GPB * kV2G;                 // [NSLOT][16] critic slots
    static constexpr int pif = wl + NSLOT * kV2G;                   // [2][GPB][16] state, float (GPB*16 doubles)
    // per-group gradient accumulators (train only), TRIANGULAR: lane r of a group only ever touches the
    // quadratic features (r,k) with k >= r -> [k(k+1)/2 + r] for k < D, then the D linear slots (padded to
    // 16) and the bias: 137 / 153 doubles per group instead of 17 x 16 (3 CTAs per SM instead of 2)
    static constexpr int Q = D * (D + 1) / 2;
    static constexpr int acc_group = (Q + kV2G + 1 + 1) & ~1;
    static constexpr int acc = pif + GPB * kV2G;
    static constexpr int total_notd = acc;
    static constexpr int total_td = acc + GPB * acc_group;
};

// critic value from the shared state buffer: lane r sums w[r,k] pi_r pi_k over k >= r (zeros below)
template <int D>
__device__ __forceinline__ double critic_value_v2(const double* __restrict__ wl, const double* __restrict__ pi,
                                                  double pi_self, int r) {
    double v = 0.0;
#pragma unroll 5
    for (int k = 0; k < D; ++k) v = fma(wl[k * kV2G + r], pi[k], v);
    v = fma(v, pi_self, wl[D * kV2G + r] * pi_self) + wl[(D + 1) * kV2G + r];
    return group_sum<kV2G>(v);
}

// TRAIN = the batched train step: accumulators wanted and NO per-step output stream requested -- every
// output test is compiled out of the step loop.  REC = some per-element stream (actions / alpha) is written.
template <int D, int NOISE, bool REC, bool TRAIN>
__global__ void __launch_bounds__(kV2Threads, 2)
rollout_v2_kernel(const RolloutParams<float> p) {
    using S = V2Smem<D>;
    constexpr int G = kV2G, NT = kV2Threads, GPB = S::GPB, NSLOT = S::NSLOT, PD = (D + 1) / 2;
    constexpr int F = num_features_c(D);
    extern __shared__ double smem[];
    const int tid = threadIdx.x, r = tid & (G - 1), grp = tid / G;
    double* slots = smem + S::tile + (grp * G + r) * kV2Slots;            // this lane's row of slots
    const double* col = smem + S::tile + grp * G * kV2Slots + r;          // column r of the group's tile
    double* pid = smem + S::pid + grp * G;                                 // [2] buffers, stride GPB*G
    float* pif = reinterpret_cast<float*>(smem + S::pif) + grp * G;        // [2] buffers, stride GPB*G
    const double* wl = smem + S::wl;
    double* acc = smem + S::acc + grp * S::acc_group;                      // this group's triangular block
    const bool td = TRAIN || p.w != nullptr;
    const bool want_acc = TRAIN || (td && p.partials != nullptr);
    const bool row_ok = r < D;
    if (td) {
        if (tid < G) stage_critic_slots<D>(smem + S::wl + tid, G, p.w, tid);
        if (want_acc)
            for (int k = r; k < S::acc_group; k += G) acc[k] = 0.0;
    }
    __syncthreads();
    const float theta = (float)(p.theta_dev ? *p.theta_dev : p.theta);
    const float shift = p.shift_f, scale = p.scale_f;
    const bool ac2 = p.reward_kind == DMFG_REWARD_AC2;
    double sum_dg = 0.0, sum_r = 0.0;
    const long long ntiles = (p.B + GPB - 1) / GPB;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        long long b = tile * GPB + grp;
        const bool live = b < p.B;                 // dead groups shadow the last population, writes masked
        if (!live) b = p.B - 1;
        const bool wr = live && row_ok;
        const NoiseKey nk = make_noise_key(p.seed, (unsigned long long)(p.pop_offset + b));
        double pi_self = row_ok ? (double)p.pi0[b * D + r] : 0.0;
        int cur = 0;
        pid[r] = pi_self;
        pif[r] = (float)pi_self;
        __syncwarp();
        double v_cur = td ? critic_value_v2<D>(wl, pid, pi_self, r) : 0.0;
        double disc = 1.0;
        if (!TRAIN && p.states != nullptr && wr) p.states[b * D + r] = (float)pi_self;
        for (int t = 0; t < p.T; ++t) {
            const long long tb = (long long)t * p.B + b;
            const long long row = (tb * D + r) * D;
            const double* pic = pid + cur * (GPB * G);
            const float* pfc = pif + cur * (GPB * G);
            // ------------------------------------------------------------------ pass 1
            const float xi = (float)pi_self + shift;
            float2 asum2 = make_float2(0.f, 0.f), dsum2 = asum2, g12 = asum2;
            double ysum = 0.0;
#pragma unroll 4
            for (int pp = 0; pp < PD; ++pp) {
                const float2 pj = *reinterpret_cast<const float2*>(pfc + 2 * pp);
                const bool ok1 = (2 * pp + 1) < D;                     // only the last pair of an odd D
                // the pair (columns 2pp, 2pp+1) runs as packed f32x2 math: one issue slot per two elements
                float2 a, dv, psi;
                alpha_psi_fast2(theta, __fadd2_rn(pj, splat2(-xi)), a, dv, psi);
                if (!ok1) { a.y = 1.0f; dv.y = 0.0f; }
                g12 = __ffma2_rn(psi, neg2(dv), g12);
                asum2 = __fadd2_rn(asum2, make_float2(a.x, ok1 ? a.y : 0.0f));
                dsum2 = __fadd2_rn(dsum2, dv);
                float y0, y1;
                if (NOISE == DMFG_NOISE_PHILOX) {
                    gamma_pair_fast(nk, p.rk, gamma_slot((uint32_t)(p.step_offset + t), D, r, pp), a, scale, y0, y1);
                } else {
                    y0 = row_ok ? p.noise_y[row + 2 * pp] : 1.0f;
                    y1 = (row_ok && ok1) ? p.noise_y[row + 2 * pp + 1] : 1.0f;
                }
                if (y0 == 0.0f) y0 = 1e-20f;                             // mfg_ac2.py:244
                if (y1 == 0.0f) y1 = 1e-20f;
                if (!ok1) y1 = 0.0f;
                ysum += (double)(y0 + y1);       // pair sum in float: <= 6e-8 relative on the pair, one conversion
                *reinterpret_cast<float2*>(slots + 2 * pp) = make_float2(y0, dv.x);
                *reinterpret_cast<float2*>(slots + 2 * pp + 1) = make_float2(y1, dv.y);
                if (REC && p.alpha != nullptr && wr) {
                    p.alpha[row + 2 * pp] = a.x;
                    p.alpha_deriv[row + 2 * pp] = dv.x;
                    if (ok1) { p.alpha[row + 2 * pp + 1] = a.y; p.alpha_deriv[row + 2 * pp + 1] = dv.y; }
                }
            }
            const float asum = asum2.x + asum2.y, dsum = dsum2.x + dsum2.y, g1 = g12.x + g12.y;
            // ------------------------------------------------------------------ row level
            double inv = (double)rcp_approx((float)ysum);              // 1/s: float seed + 2 Newton steps
            inv = inv * (2.0 - ysum * inv);
            inv = inv * (2.0 - ysum * inv);
            const float inv_f = (float)inv;
            const double q = pi_self * inv;
            const float psi_row = digamma_fast(asum);
            // ------------------------------------------------------------------ pass 2
            double racc = 0.0;
            float g2 = 0.f;
            float* act_row = (REC && p.actions != nullptr && wr) ? p.actions + row : nullptr;
#pragma unroll 5
            for (int j = 0; j < D; ++j) {
                const float2 yd = *reinterpret_cast<const float2*>(slots + j);
                const double yv = (double)yd.x;
                const double P = yv * inv;
                const float Pf = yd.x * inv_f;
                g2 = fmaf(lg2_approx(Pf), yd.y, g2);
                // AC2: P^2 (pi_j - pi_i); synthetic: P^2 (the factor is hoisted out of the element loop)
                const double dlt = ac2 ? pic[j] - pi_self : 1.0;
                racc = fma(P * P, dlt, racc);
                slots[j] = yv * q;                                       // flux pi_i P_ij
                if (REC && act_row != nullptr) act_row[j] = Pf;
            }
            __syncwarp();
            // ------------------------------------------------------------------ pi' = P^T pi (transposed read)
            double next_self = 0.0;
#pragma unroll 5
            for (int i = 0; i < D; ++i) next_self += col[i * kV2Slots];
            if (!row_ok) next_self = 0.0;
            double rew = 0.0;
            if (ac2) rew = pi_self * racc;
            else if (p.reward_kind == DMFG_REWARD_SYNTHETIC) rew = -0.5 * pi_self * racc;
            const double glane = row_ok ? (double)(g1 + DMFG_LN2 * g2 + psi_row * dsum) : 0.0;
            rew = group_sum<G>(rew);
            const double grad = group_sum<G>(glane);
            const int nxt = cur ^ 1;
            pid[nxt * (GPB * G) + r] = next_self;
            pif[nxt * (GPB * G) + r] = (float)next_self;
            __syncwarp();
            if (!TRAIN && p.rewards_in != nullptr) rew = (double)p.rewards_in[tb];
            if (td) {
                const double v_next = critic_value_v2<D>(wl, pid + nxt * (GPB * G), next_self, r);
                const double gfac = p.discount_kind == DMFG_DISCOUNT_STEP ? p.gamma : disc;
                const double delta = rew + gfac * v_next - v_cur;
                if (want_acc && live) {
                    const double dp = delta * pi_self;
                    // quadratic features (r,k), k >= r, at [k(k+1)/2 + r]; linear at [Q + r]; bias at [Q + 16]
#pragma unroll
                    for (int k = 0; k < D; ++k)
                        if (k >= r) acc[k * (k + 1) / 2 + r] = fma(dp, pic[k], acc[k * (k + 1) / 2 + r]);
                    if (row_ok) acc[S::Q + r] += dp;
                    if (r == 0) {
                        acc[S::Q + G] += delta;
                        sum_dg = fma(delta, grad, sum_dg);
                    }
                }
                if (!TRAIN && p.deltas != nullptr && live && r == 0) p.deltas[tb] = (float)delta;
                v_cur = v_next;
            }
            if (live && r == 0) {
                sum_r += rew;
                if (!TRAIN) {
                    if (p.rewards != nullptr) p.rewards[tb] = (float)rew;
                    if (p.grads != nullptr) p.grads[tb] = (float)grad;
                }
            }
            disc *= p.gamma;
            pi_self = next_self;
            cur = nxt;
            if (!TRAIN && p.states != nullptr && wr) p.states[(tb + p.B) * D + r] = (float)pi_self;
        }
        if (!TRAIN && p.pi_final != nullptr && wr) p.pi_final[b * D + r] = (float)pi_self;
        // make the next population start from buffer 0 again
        __syncwarp();
    }
    if (!want_acc) return;
    // ---- per-CTA partial sums, fixed order (deterministic) -- same layout as rollout_fast_kernel -------
    __shared__ double red[2][NT / 32];
    {
        double a = sum_dg, c = sum_r;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if ((tid & 31) == 0) { red[0][tid >> 5] = a; red[1][tid >> 5] = c; }
    }
    __syncthreads();
    double* out = p.partials + (long long)blockIdx.x * (2 + F);
    const double* accbase = smem + S::acc;
    for (int f = tid; f < F; f += NT) {
        // feature f of the reference order -> slot of the triangular per-group block
        int slot;
        constexpr int Q = S::Q;
        if (f < Q) {
            int rw = 0, rem = f;
            while (rem >= D - rw) { rem -= D - rw; ++rw; }
            const int k = rw + rem;
            slot = k * (k + 1) / 2 + rw;
        } else if (f < Q + D) {
            slot = Q + (f - Q);
        } else {
            slot = Q + G;
        }
        double s = 0.0;
        for (int g = 0; g < GPB; ++g) s += accbase[g * S::acc_group + slot];
        out[1 + f] = s;
    }
    if (tid == 0) {
        double a = 0.0, c = 0.0;
        for (int wv = 0; wv < NT / 32; ++wv) { a += red[0][wv]; c += red[1][wv]; }
        out[0] = a;
        out[1 + F] = c;
    }
}

}  // namespace dmfg
