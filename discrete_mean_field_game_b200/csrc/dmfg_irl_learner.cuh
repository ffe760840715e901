// AC_IRL.train (ac_irl.py:634-732) as ONE kernel: a serial learner per CTA with the reward network IN the loop.
//
// The reference queries the reward net through sess.run for every transition (ac_irl.py:683); round 1 ran that as a
// chain of 4 launches per transition (rollout(T=1) -> rnet forward -> TD -> update, 37 us per transition as a CUDA
// graph).  Here a CTA of 128 threads owns the learner for all its episodes: thread (i, jp) = (row of P, column pair)
// as in learner_cta_kernel samples its Gamma pair, the row sums / pi' / TD error use the same shuffles and exchanges,
// and r = r_net(pi, P) (networks.py:13-157) is evaluated by the same threads in between:
//     P (normalised, float32 like the TF graph) -> zero-haloed shared tile
//     conv 5x5 + ReLU: thread (i, jp) computes outputs (i, 2jp), (i, 2jp+1)      -> second tile
//     conv 3x3 (2 channels) + ReLU: the same two positions, both channels          (registers)
//     fc3: each thread's 4 activations x its 4 rows of W3, block reduction of the n3 sums
//     fc4 over [h3, pi] + ReLU, out + tanh: lanes 0..7 of every warp; dropout masks from Philox keyed by the transition id,
//     exactly as rnet_kernel draws them (the reference's dropout is active at inference too, networks.py:70)
// The TD partial sums that do not depend on r cross the same barrier as the fc3 partial sums and every warp evaluates
// the head itself.  Weights (15 KB) are staged in shared memory once per CTA.  3 __syncthreads per transition.  Cumulative discount gamma^t on V(pi') and episodes counted from
// 1 are the caller's choice (discount_kind / episode0), as in the other learner kernels.
// float streams, d <= 16 (d = 15 is AC_IRL's default), n_fc3, n_fc4 <= 8.
#pragma once
#include "dmfg_learner_cta.cuh"
#include "dmfg_rnet.cuh"

namespace dmfg {

struct IrlLearnerNet {
    const float* params;          // flat reward-net parameters (rnet_layout order)
    int n3, n4;
    int dropout;                  // DMFG_DROPOUT_NONE | DMFG_DROPOUT_PHILOX
    float keep_prob;
    unsigned long long seed, sample_offset;   // transition (l, e, t) draws the masks of sample_offset + (l*E + e)*T + t
    float* reward_trace;          // optional [L][E][T]
};

template <int D>
struct IrlLearnerSmem {
    static constexpr int SP = 21, SC = 19;                 // odd row strides of the two zero-haloed tiles
    static constexpr int pt = 0;                            // [D+4][SP] action tile (halo 2)
    static constexpr int c1t = pt + 20 * SP;                // [D+2][SC] conv1 tile (halo 1)
    static constexpr int pis = c1t + 18 * SC;               // [2][16] state, double-buffered by step parity (the head reads it
                                                            // after the last barrier of a step; the next step rewrites it)
    static constexpr int z3p = pis + 32;                    // [4 warps][8]
    static constexpr int w3a = (z3p + 32 + 3) & ~3;         // [2 d^2][8] fc3 weights, 16-byte aligned rows (zero padded columns)
    // row r of W3 sits at 8 r + 4 (r / 4): thread (i, jp) reads rows 30 i + 4 jp + u, so without the skew the 8 lanes of
    // a quarter warp (one i, jp = 0..7) are 32 words apart -- the same banks, an 8-way conflict on every LDS.128 (ncu:
    // 26-30 wavefronts per load, 16 % of the kernel's samples); with it they are 36 words apart: 8 distinct bank quads
    static __host__ __device__ constexpr int w3row(int r) { return 8 * r + 4 * (r >> 2); }
    static constexpr int w3size = (w3row(2 * D * D - 1) + 8 + 3) & ~3;
    static_assert(D <= 16, "one 128-thread CTA covers d <= 16");
};

template <int D, int NOISE>
__global__ void __launch_bounds__(128)
irl_learner_cta_kernel(const LearnerParams<float> p, const PhiloxKeys rk, const IrlLearnerNet net) {
    using Gm = LearnerCtaGeom<D>;
    using SMp = IrlLearnerSmem<D>;
    static_assert(Gm::NT == 128 && Gm::PS == 8, "thread = (row, column pair), 16 x 8");
    constexpr int F = num_features_c(D), Q = D * (D + 1) / 2, NW = Gm::NW, PS = Gm::PS, SP = SMp::SP, SC = SMp::SC;
    extern __shared__ __align__(16) float ism[];
    __shared__ double colpart[NW][Gm::NC];
    __shared__ double red[NW][3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = tid / PS, jp = tid % PS, ja = 2 * jp, jb = 2 * jp + 1;
    const bool row_ok = i < D, pair_ok = jp < Gm::PD, ok_b = jb < D, head = jp == 0;
    const long long l = blockIdx.x;
    const int n3 = net.n3, n4 = net.n4;
    const RnetLayout L = rnet_layout(D, n3, n4);
    float* Pt = ism + SMp::pt;
    float* C1t = ism + SMp::c1t;
    float* pis = ism + SMp::pis;
    float* z3p = ism + SMp::z3p;
    float* w3a = ism + SMp::w3a;
    float* wf = w3a + SMp::w3size;                          // the flat parameter vector
    for (int k = tid; k < SMp::w3a; k += 128) ism[k] = 0.f;  // tiles (halos stay zero for the whole kernel)
    for (int k = tid; k < L.total; k += 128) wf[k] = net.params[k];
    __syncthreads();
    for (int k = tid; k < 2 * D * D * 8; k += 128) {
        const int row = k >> 3, j = k & 7;
        w3a[SMp::w3row(row) + j] = j < n3 ? wf[L.w3 + row * n3 + j] : 0.f;
    }
    const float inv_keep = net.dropout ? 1.0f / net.keep_prob : 1.0f;
    // conv weights of this thread's use in registers
    float k1[25], k2[18];
#pragma unroll
    for (int k = 0; k < 25; ++k) k1[k] = wf[L.k1 + k];
#pragma unroll
    for (int k = 0; k < 18; ++k) k2[k] = wf[L.k2 + k];
    const float b1 = wf[L.b1], b20 = wf[L.b2], b21 = wf[L.b2 + 1];
    // critic weights owned here (as in learner_cta_kernel)
    const bool va = row_ok && pair_ok && ja >= i, vb = row_ok && pair_ok && ok_b && jb >= i;
    double* wg = p.w + l * F;
    double w_a = va ? wg[quad_index(D, i, ja)] : 0.0;
    double w_b = vb ? wg[quad_index(D, i, jb)] : 0.0;
    double w_lin = (head && row_ok) ? wg[Q + i] : 0.0;
    double w_bias = tid == 0 ? wg[Q + D] : 0.0;
    double theta = p.theta[l];
    const float shift = (float)(p.shift ? p.shift[l] : p.shift_scalar);
    const float scale = (float)(p.alpha_scale ? p.alpha_scale[l] : p.alpha_scale_scalar);
    const NoiseKey nk = make_noise_key(p.seed, (unsigned long long)(p.learner_offset + l));
    double pi_i = 0.0, pi_a = 0.0, pi_b = 0.0;
    int par = 0;                                            // parity of the running step count (state double buffer)
    __syncthreads();
    for (int e = 0; e < p.E; ++e) {
        const int episode = p.episode0 + e;
        int start;
        if (p.start_rows != nullptr) {
            start = p.start_rows[l * p.E + e];
        } else {
            const uint4 wv = philox4x32_10(nk.p0, nk.p1, (uint32_t)(episode + p.noise_episode_offset),
                                           DMFG_CTR_START, nk.k0, nk.k1);
            start = (int)__umulhi(wv.x, (uint32_t)p.S);          // randint(S), ac_irl.py:655
        }
        const float* s0 = p.mat_pi0 + (long long)start * D;
        pi_i = row_ok ? (double)s0[i] : 0.0;
        pi_a = pair_ok ? (double)s0[ja] : 0.0;
        pi_b = (pair_ok && ok_b) ? (double)s0[jb] : 0.0;
        const double lr_c = p.constant_lr ? p.lr_critic : p.lr_critic / (episode + 1.0);
        const double lr_a = p.constant_lr ? p.lr_actor : p.lr_actor / ((episode + 1.0) * log(log(episode + 20.0)));
        double disc = 1.0, total = 0.0;
        for (int t = 0; t < p.T; ++t) {
            const long long et = ((long long)l * p.E + e) * p.T + t;
            const float thf = (float)theta;
            // ---------------------------------------------------------------- this thread's pair of row i
            const float xi = (float)pi_i + shift;
            float2 a, dv, psi;
            alpha_psi_fast2(thf, __fadd2_rn(make_float2((float)pi_a, (float)pi_b), splat2(-xi)), a, dv, psi);
            if (!ok_b) dv.y = 0.0f;
            if (!pair_ok) dv = make_float2(0.f, 0.f);
            float y0 = 1.0f, y1 = 1.0f;
            if (NOISE == DMFG_NOISE_PHILOX) {
                if (pair_ok) {
                    const uint32_t slot = gamma_slot((uint32_t)((episode + p.noise_episode_offset) * p.T + t), D, i, jp);
                    gamma_pair_fast(nk, rk, slot, a, scale, y0, y1);
                }
            } else {
                const float* nr = p.noise_y + (et * D + i) * D;
                y0 = gamma_floor((row_ok && pair_ok) ? nr[ja] : 1.0f);             // ac_irl.py:536
                y1 = gamma_floor((row_ok && pair_ok && ok_b) ? nr[jb] : 1.0f);
            }
            float g1 = -fmaf(psi.x, dv.x, psi.y * dv.y);
            float g2 = fmaf(lg2_approx(y0), dv.x, lg2_approx(y1) * dv.y);
            float asum = pair_ok ? a.x + (ok_b ? a.y : 0.0f) : 0.0f, dsum = dv.x + dv.y;
            const double yd0 = pair_ok ? (double)y0 : 0.0, yd1 = (pair_ok && ok_b) ? (double)y1 : 0.0;
            double ysum = yd0 + yd1;
#pragma unroll
            for (int o = 1; o < PS; o <<= 1) {
                ysum += __shfl_xor_sync(0xffffffffu, ysum, o);
                asum += __shfl_xor_sync(0xffffffffu, asum, o);
                dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
                g1 += __shfl_xor_sync(0xffffffffu, g1, o);
                g2 += __shfl_xor_sync(0xffffffffu, g2, o);
            }
            const float ysum_f = (float)ysum;
            double inv = (double)rcp_approx(ysum_f);
            inv = inv * (2.0 - ysum * inv);
            inv = inv * (2.0 - ysum * inv);
            const double q = pi_i * inv;
            const float psi_row = digamma_fast(asum);
            const float lnp_term = DMFG_LN2 * fmaf(-lg2_approx(ysum_f), dsum, g2);
            const double glane = (row_ok && head) ? (double)(g1 + lnp_term + psi_row * dsum) : 0.0;
            // ---------------------------------------------------------------- P (float32, as recorded) -> tile; pi -> shared
            {
                const float inv_f = (float)inv;
                if (row_ok && pair_ok) {
                    Pt[(i + 2) * SP + ja + 2] = (float)yd0 * inv_f;
                    if (ok_b) Pt[(i + 2) * SP + jb + 2] = (float)yd1 * inv_f;
                }
                if (head) pis[par * 16 + i] = (float)pi_i;        // rows >= D hold 0
            }
            // ---------------------------------------------------------------- pi'_j = sum_i q_i y_ij
            double ca = q * yd0, cb = q * yd1;
#pragma unroll
            for (int o = PS; o < 32; o <<= 1) {
                ca += __shfl_xor_sync(0xffffffffu, ca, o);
                cb += __shfl_xor_sync(0xffffffffu, cb, o);
            }
            if (lane < PS) { colpart[warp][ja] = ca; colpart[warp][jb] = cb; }
            __syncthreads();                                                      // (1) P tile, pi, column partials
            // dropout masks of this transition (lane m of every warp = unit m): drawn here, two barriers ahead of the head that
            // uses them, so the two Philox chains run under the shared-memory latency of the conv phases
            float m3 = 1.f, m4 = 1.f;
            if (net.dropout == DMFG_DROPOUT_PHILOX) {
                const int m = lane & 7;
                const unsigned long long sid = net.sample_offset + (unsigned long long)et;
                const uint32_t k0 = (uint32_t)net.seed, k1s = (uint32_t)(net.seed >> 32);
                const uint4 wa = philox4x32_10((uint32_t)sid, (uint32_t)(sid >> 32), (uint32_t)(m >> 2), DMFG_CTR_DROPOUT, k0, k1s);
                const uint4 wb = philox4x32_10((uint32_t)sid, (uint32_t)(sid >> 32), 64u + (uint32_t)(m >> 2), DMFG_CTR_DROPOUT, k0, k1s);
                const uint32_t sa = (m & 3) == 0 ? wa.x : (m & 3) == 1 ? wa.y : (m & 3) == 2 ? wa.z : wa.w;
                const uint32_t sb = (m & 3) == 0 ? wb.x : (m & 3) == 1 ? wb.y : (m & 3) == 2 ? wb.z : wb.w;
                m3 = u01(sa) < net.keep_prob ? 1.f : 0.f;
                m4 = u01(sb) < net.keep_prob ? 1.f : 0.f;
            }
            double nx_i = 0.0, nx_a = 0.0, nx_b = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                nx_i += colpart[w][i];
                nx_a += colpart[w][ja];
                nx_b += colpart[w][jb];
            }
            if (!row_ok) nx_i = 0.0;
            if (!ok_b) nx_b = 0.0;
            // ---------------------------------------------------------------- conv1 at (i, ja), (i, jb)
            float c1a = b1, c1b = b1;
#pragma unroll
            for (int dh = 0; dh < 5; ++dh) {
                float v[6];
#pragma unroll
                for (int c = 0; c < 6; ++c) v[c] = Pt[(i + dh) * SP + ja + c];
#pragma unroll
                for (int dw = 0; dw < 5; ++dw) {
                    c1a = fmaf(v[dw], k1[dh * 5 + dw], c1a);
                    c1b = fmaf(v[dw + 1], k1[dh * 5 + dw], c1b);
                }
            }
            if (row_ok && pair_ok) {
                C1t[(i + 1) * SC + ja + 1] = fmaxf(c1a, 0.f);
                if (ok_b) C1t[(i + 1) * SC + jb + 1] = fmaxf(c1b, 0.f);
            }
            __syncthreads();                                                      // (2) conv1 tile
            // ---------------------------------------------------------------- conv2 (2 channels) at the same positions, fc3
            float z3[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) z3[j] = 0.f;
            if (row_ok && pair_ok) {
                float o0a = b20, o1a = b21, o0b = b20, o1b = b21;
#pragma unroll
                for (int dh = 0; dh < 3; ++dh) {
                    float v[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) v[c] = C1t[(i + dh) * SC + ja + c];
#pragma unroll
                    for (int dw = 0; dw < 3; ++dw) {
                        const float ka = k2[(dh * 3 + dw) * 2], kb = k2[(dh * 3 + dw) * 2 + 1];
                        o0a = fmaf(v[dw], ka, o0a); o1a = fmaf(v[dw], kb, o1a);
                        o0b = fmaf(v[dw + 1], ka, o0b); o1b = fmaf(v[dw + 1], kb, o1b);
                    }
                }
                const float act[4] = {fmaxf(o0a, 0.f), fmaxf(o1a, 0.f), ok_b ? fmaxf(o0b, 0.f) : 0.f, ok_b ? fmaxf(o1b, 0.f) : 0.f};
                const int kbase = (i * D + ja) * 2;                               // NHWC flatten: (row, col, channel)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (u < 2 || ok_b) {
                        const float4 wlo = *reinterpret_cast<const float4*>(w3a + SMp::w3row(kbase + u));
                        const float4 whi = *reinterpret_cast<const float4*>(w3a + SMp::w3row(kbase + u) + 4);
                        z3[0] = fmaf(act[u], wlo.x, z3[0]); z3[1] = fmaf(act[u], wlo.y, z3[1]);
                        z3[2] = fmaf(act[u], wlo.z, z3[2]); z3[3] = fmaf(act[u], wlo.w, z3[3]);
                        z3[4] = fmaf(act[u], whi.x, z3[4]); z3[5] = fmaf(act[u], whi.y, z3[5]);
                        z3[6] = fmaf(act[u], whi.z, z3[6]); z3[7] = fmaf(act[u], whi.w, z3[7]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) z3[j] += __shfl_xor_sync(0xffffffffu, z3[j], o);
            }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) z3p[warp * 8 + j] = z3[j];
            }
            // ---------------------------------------------------------------- TD error with the CURRENT w: everything but r
            // (gamma V(pi') - V(pi) and d log F / d theta do not depend on the reward net: their warp sums travel through the
            // same barrier as the fc3 partial sums, and every warp then evaluates the small head itself -- no fourth barrier,
            // no warp waiting for warp 0)
            {
                double vn = nx_i * fma(w_a, nx_a, w_b * nx_b);
                double vc = pi_i * fma(w_a, pi_a, w_b * pi_b);
                if (head) { vn = fma(w_lin, nx_i, vn); vc = fma(w_lin, pi_i, vc); }
                if (tid == 0) { vn += w_bias; vc += w_bias; }
                const double gfac = p.discount_kind == DMFG_DISCOUNT_STEP ? p.gamma : disc;
                double dpart = fma(gfac, vn, -vc), gpart = glane;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    dpart += __shfl_xor_sync(0xffffffffu, dpart, o);
                    gpart += __shfl_xor_sync(0xffffffffu, gpart, o);
                }
                if (lane == 0) { red[warp][0] = dpart; red[warp][1] = gpart; }
            }
            __syncthreads();                                                      // (3) fc3 partial sums, TD partial sums
            // ---------------------------------------------------------------- fc4, out, tanh: every warp, lane % 8 = unit
            float r = 0.f;
            {
                const int m = lane & 7;
                // lane m holds h3[m]
                float h3 = 0.f;
                if (m < n3) {
                    const float z = (z3p[m] + z3p[8 + m]) + (z3p[16 + m] + z3p[24 + m]) + wf[L.b3 + m];
                    h3 = fmaxf(z, 0.f) * m3 * inv_keep;
                }
                // lane m computes z4[m] = b4 + sum_j h3[j] W4[j][m] + sum_k pi_k W4[n3+k][m]
                // (two independent chains: the h3 part and the state part; the state part does not wait for fc3)
                float z4 = m < n4 ? wf[L.b4 + m] : 0.f, z4s = 0.f;
                if (m < n4) {
#pragma unroll
                    for (int k = 0; k < D; ++k) z4s = fmaf(pis[par * 16 + k], wf[L.w4 + (n3 + k) * n4 + m], z4s);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float hj = __shfl_sync(0xffffffffu, h3, j);
                    if (j < n3 && m < n4) z4 = fmaf(hj, wf[L.w4 + j * n4 + m], z4);
                }
                z4 += z4s;
                float part = m < n4 ? fmaxf(z4, 0.f) * m4 * inv_keep * wf[L.w5 + m] : 0.f;
                part += __shfl_xor_sync(0xffffffffu, part, 1);
                part += __shfl_xor_sync(0xffffffffu, part, 2);
                part += __shfl_xor_sync(0xffffffffu, part, 4);
                r = tanhf(part + wf[L.b5]);
            }
            double delta = (double)r, grad = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) { delta += red[w][0]; grad += red[w][1]; }
            // critic first, then actor, both with the same delta (ac_irl.py:691-708)
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
                if (net.reward_trace) net.reward_trace[et] = r;
                total += (double)r;
            }
            disc *= p.gamma;
            par ^= 1;
            pi_i = nx_i; pi_a = nx_a; pi_b = nx_b;
        }
        if (p.total_reward && tid == 0) p.total_reward[l * p.E + e] = total;
    }
    if (tid == 0) p.theta[l] = theta;
    if (p.pi_final && row_ok && head) p.pi_final[l * D + i] = (float)pi_i;
    if (va) wg[quad_index(D, i, ja)] = w_a;
    if (vb) wg[quad_index(D, i, jb)] = w_b;
    if (head && row_ok) wg[Q + i] = w_lin;
    if (tid == 0) wg[Q + D] = w_bias;
}

}  // namespace dmfg
