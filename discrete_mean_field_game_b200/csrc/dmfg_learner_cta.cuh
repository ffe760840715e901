// One serial learner per CTA: the latency form of learners_v2_kernel (mfg_ac2.py:448-539, per-step online updates).
//
// learners_v2_kernel gives a learner 16 lanes (lane = row of P, 8 column pairs walked in sequence): right when there
// are thousands of learners, but ONE learner -- the reference's own run, BASELINE configs[0] -- is then a serial chain
// of ~7000 cycles per step.  Here a learner owns a CTA (128 threads at d = 15 / 16: 16 rows x 8 pair slots; 352 at the
// reference's default d = 21: 22 rows x 16 pair slots, 11 of them live): thread (i, jp) = (row, column pair) samples ONE
// Gamma pair per step, row sums are 8- / 16-lane shuffles, pi' = sum_i q_i y_ij is a shuffle over the rows of a warp plus
// an exchange through shared memory, and every thread keeps "its" two quadratic critic weights w[i,2jp], w[i,2jp+1]
// (thread (i,0) also the linear weight of pi_i, thread 0 the bias) in registers for the whole run: TD error and
// d log F/d theta come from one block reduction per step, the updates are thread-local.  Two __syncthreads per step.
// Same math, same Philox slots (seed, learner, step, row, pair) as every other kernel: a learner draws the same Gamma
// variates here and in learners_v2_kernel; sums are taken in a different (fixed) order.
// float streams, d = 15 / 16 / 21.
#pragma once
#include "dmfg_rollout2.cuh"

namespace dmfg {

template <int D>
struct LearnerCtaGeom {
    static constexpr int PD = (D + 1) / 2;             // column pairs of a row
    static constexpr int PS = PD <= 8 ? 8 : 16;        // pair slots per row (a power of two: shuffle reductions)
    static constexpr int RPW = 32 / PS;                // rows per warp
    static constexpr int ROWS = (D + RPW - 1) / RPW * RPW;
    static constexpr int NT = ROWS * PS;
    static constexpr int NW = NT / 32;
    static constexpr int NC = 2 * PS;                  // column slots
    static_assert(PD <= 16 && NT <= 1024, "d <= 32");
};

template <int D, int NOISE>
__global__ void __launch_bounds__(LearnerCtaGeom<D>::NT)
learner_cta_kernel(const LearnerParams<float> p, const PhiloxKeys rk) {
    using Gm = LearnerCtaGeom<D>;
    constexpr int F = num_features_c(D), Q = D * (D + 1) / 2, NW = Gm::NW, PS = Gm::PS;
    __shared__ double colpart[NW][Gm::NC];     // per-warp partial of pi'_j (the RPW rows of the warp)
    __shared__ double red[NW][3];              // per-warp partials of (delta, grad, episode reward)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = tid / PS, jp = tid % PS, ja = 2 * jp, jb = 2 * jp + 1;
    const bool row_ok = i < D, pair_ok = jp < Gm::PD, ok_b = jb < D, head = jp == 0;
    const long long l = blockIdx.x;
    // upper-triangular features owned here
    const bool va = row_ok && pair_ok && ja >= i, vb = row_ok && pair_ok && ok_b && jb >= i;
    double* wg = p.w + l * F;
    double w_a = va ? wg[quad_index(D, i, ja)] : 0.0;
    double w_b = vb ? wg[quad_index(D, i, jb)] : 0.0;
    double w_lin = (head && row_ok) ? wg[Q + i] : 0.0;
    double w_bias = tid == 0 ? wg[Q + D] : 0.0;
    double theta = p.theta[l];
    const float shift = (float)(p.shift ? p.shift[l] : p.shift_scalar);
    const float scale = (float)(p.alpha_scale ? p.alpha_scale[l] : p.alpha_scale_scalar);
    const bool ac2 = p.reward_kind == DMFG_REWARD_AC2;
    const bool has_reward = p.reward_kind != DMFG_REWARD_NONE;
    const double rew_scale = ac2 ? 1.0 : -0.5;
    const NoiseKey nk = make_noise_key(p.seed, (unsigned long long)(p.learner_offset + l));
    double pi_i = 0.0, pi_a = 0.0, pi_b = 0.0;
    for (int e = 0; e < p.E; ++e) {
        const int episode = p.episode0 + e;
        int start;
        if (p.start_rows != nullptr) {
            start = p.start_rows[l * p.E + e];
        } else {
            const uint4 wv = philox4x32_10(nk.p0, nk.p1, (uint32_t)(episode + p.noise_episode_offset),
                                           DMFG_CTR_START, nk.k0, nk.k1);
            start = (int)__umulhi(wv.x, (uint32_t)p.S);          // randint(S), mfg_ac2.py:466
        }
        const float* s0 = p.mat_pi0 + (long long)start * D;
        pi_i = row_ok ? (double)s0[i] : 0.0;
        pi_a = pair_ok ? (double)s0[ja] : 0.0;
        pi_b = (pair_ok && ok_b) ? (double)s0[jb] : 0.0;
        const double lr_c = p.constant_lr ? p.lr_critic : p.lr_critic / (episode + 1.0);
        const double lr_a = p.constant_lr ? p.lr_actor : p.lr_actor / ((episode + 1.0) * log(log(episode + 20.0)));
        double disc = 1.0, total_thread = 0.0;
        for (int t = 0; t < p.T; ++t) {
            const long long et = ((long long)l * p.E + e) * p.T + t;
            const float thf = (float)theta;
            // ---------------------------------------------------------------- this thread's pair of row i
            const float xi = (float)pi_i + shift;
            float2 a, dv, psi;
            alpha_psi_fast2(thf, __fadd2_rn(make_float2((float)pi_a, (float)pi_b), splat2(-xi)), a, dv, psi);
            if (!ok_b) dv.y = 0.0f;                              // phantom column of an odd D
            if (!pair_ok) dv = make_float2(0.f, 0.f);            // idle pair slot (d = 21: slots 11 .. 15)
            float y0 = 1.0f, y1 = 1.0f;
            if (NOISE == DMFG_NOISE_PHILOX) {
                if (pair_ok) {
                    const uint32_t slot = gamma_slot((uint32_t)((episode + p.noise_episode_offset) * p.T + t), D, i, jp);
                    gamma_pair_fast(nk, rk, slot, a, scale, y0, y1);
                }
            } else {
                const float* nr = p.noise_y + (et * D + i) * D;
                y0 = gamma_floor((row_ok && pair_ok) ? nr[ja] : 1.0f);             // mfg_ac2.py:244
                y1 = gamma_floor((row_ok && pair_ok && ok_b) ? nr[jb] : 1.0f);
            }
            float g1 = -fmaf(psi.x, dv.x, psi.y * dv.y);
            float g2 = fmaf(lg2_approx(y0), dv.x, lg2_approx(y1) * dv.y);
            float asum = pair_ok ? a.x + (ok_b ? a.y : 0.0f) : 0.0f, dsum = dv.x + dv.y;
            const double yd0 = pair_ok ? (double)y0 : 0.0, yd1 = (pair_ok && ok_b) ? (double)y1 : 0.0;
            double ysum = yd0 + yd1;
            // reward weights: AC2 sum_j y^2 (pi_j - pi_i);  synthetic sum_j y^2
            double racc = 0.0;
            if (has_reward) {
                const double c1 = ac2 ? 1.0 : 0.0, c0 = ac2 ? -pi_i : 1.0;
                racc = fma(yd0 * yd0, fma(c1, pi_a, c0), (yd1 * yd1) * fma(c1, pi_b, c0));
            }
            // ---------------------------------------------------------------- row sums over the pair slots
#pragma unroll
            for (int o = 1; o < PS; o <<= 1) {
                ysum += __shfl_xor_sync(0xffffffffu, ysum, o);
                racc += __shfl_xor_sync(0xffffffffu, racc, o);
                asum += __shfl_xor_sync(0xffffffffu, asum, o);
                dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
                g1 += __shfl_xor_sync(0xffffffffu, g1, o);
                g2 += __shfl_xor_sync(0xffffffffu, g2, o);
            }
            const float ysum_f = (float)ysum;
            double inv = (double)rcp_approx(ysum_f);                   // 1/s: float seed + 2 Newton steps
            inv = inv * (2.0 - ysum * inv);
            inv = inv * (2.0 - ysum * inv);
            const double q = pi_i * inv;
            const float psi_row = digamma_fast(asum);
            const float lnp_term = DMFG_LN2 * fmaf(-lg2_approx(ysum_f), dsum, g2);
            const double glane = (row_ok && head) ? (double)(g1 + lnp_term + psi_row * dsum) : 0.0;
            const double rew_row = (has_reward && head) ? rew_scale * (q * inv) * racc : 0.0;
            // ---------------------------------------------------------------- pi'_j = sum_i q_i y_ij
            double ca = q * yd0, cb = q * yd1;
#pragma unroll
            for (int o = PS; o < 32; o <<= 1) {
                ca += __shfl_xor_sync(0xffffffffu, ca, o);
                cb += __shfl_xor_sync(0xffffffffu, cb, o);
            }
            if (lane < PS) { colpart[warp][ja] = ca; colpart[warp][jb] = cb; }
            __syncthreads();
            double nx_i = 0.0, nx_a = 0.0, nx_b = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                nx_i += colpart[w][i];
                nx_a += colpart[w][ja];
                nx_b += colpart[w][jb];
            }
            if (!row_ok) nx_i = 0.0;
            if (!ok_b) nx_b = 0.0;
            // ---------------------------------------------------------------- TD error with the CURRENT w
            double vn = nx_i * fma(w_a, nx_a, w_b * nx_b);
            double vc = pi_i * fma(w_a, pi_a, w_b * pi_b);
            if (head) { vn = fma(w_lin, nx_i, vn); vc = fma(w_lin, pi_i, vc); }
            if (tid == 0) { vn += w_bias; vc += w_bias; }
            const double gfac = p.discount_kind == DMFG_DISCOUNT_STEP ? p.gamma : disc;
            double dpart = rew_row + fma(gfac, vn, -vc), gpart = glane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                dpart += __shfl_xor_sync(0xffffffffu, dpart, o);
                gpart += __shfl_xor_sync(0xffffffffu, gpart, o);
            }
            if (lane == 0) { red[warp][0] = dpart; red[warp][1] = gpart; }
            __syncthreads();
            double delta = 0.0, grad = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) { delta += red[w][0]; grad += red[w][1]; }
            // critic first, then actor, both with the same delta (mfg_ac2.py:505-522)
            const double step_w = lr_c * delta;
            const double dp = step_w * pi_i;
            if (va) w_a = fma(dp, pi_a, w_a);
            if (vb) w_b = fma(dp, pi_b, w_b);
            if (head && row_ok) w_lin += dp;
            if (tid == 0) w_bias += step_w;
            theta = fma(lr_a * delta, grad, theta);
            if (tid == 0) {
                if (p.theta_trace) p.theta_trace[et] = theta;
                if (p.delta_trace) p.delta_trace[et] = delta;
            }
            total_thread += rew_row;
            disc *= p.gamma;
            pi_i = nx_i; pi_a = nx_a; pi_b = nx_b;
        }
        if (p.total_reward) {
            double tot = total_thread;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0) red[warp][2] = tot;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int w = 0; w < NW; ++w) s += red[w][2];
                p.total_reward[l * p.E + e] = s;
            }
        }
    }
    if (tid == 0) p.theta[l] = theta;
    if (p.pi_final && row_ok && head) p.pi_final[l * D + i] = (float)pi_i;
    if (va) wg[quad_index(D, i, ja)] = w_a;
    if (vb) wg[quad_index(D, i, jb)] = w_b;
    if (head && row_ok) wg[Q + i] = w_lin;
    if (tid == 0) wg[Q + D] = w_bias;
}

}  // namespace dmfg
