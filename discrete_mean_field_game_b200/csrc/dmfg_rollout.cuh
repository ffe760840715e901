// Fused population-step kernels (sm_100a): the batched rollout with frozen
// parameters, the independent serial learners, and the generic-d fallbacks.
//
// FAST variant -- "row per lane":
//   a group of G lanes (G = 16 for d = 15/16: half a warp) owns one population.
//   Lane r owns ROW r of the d x d transition matrix: it evaluates the d
//   concentrations alpha_rj, draws the d Gamma variates, normalises the row,
//   and keeps every row-wise quantity (row sum, sum_j alpha, sum_j alpha',
//   psi terms, ln P) thread-local with d-way instruction-level parallelism.
//   Only the mean-field step pi'_j = sum_i pi_i P_ij crosses lanes: a
//   recursive-halving reduce-scatter + all-gather of d doubles (60 SHFL per
//   step, ~2 % of the step).  The simplex state lives in registers for the
//   whole episode; P never leaves registers unless the caller asks for the
//   trajectory.
//   Precision: transcendental work (softplus, Gamma sampler, digamma, ln P)
//   in the stream dtype R; the state recurrence, row normalisation, reward,
//   critic value and every reduction in double (B200 FP64 = 1/2 FP32 rate,
//   ~3 % of the issue slots), which is what keeps the TD error -- a difference
//   of nearly equal terms -- inside 1e-5 of the float64 reference.
//
// GENERIC variant -- "warp per population", any d <= DMFG_MAX_D:
//   rows are processed one after the other, lanes stride over column pairs;
//   row reductions are warp shuffles; pi' accumulates lane-locally.
#pragma once
#include "dmfg_math.cuh"
#include "../../include/dmfg.h"

namespace dmfg {

constexpr int kFastThreads = 256;
constexpr int kGenericThreads = 128;

__host__ __device__ constexpr int num_features_c(int d) { return d * (d + 1) / 2 + d + 1; }
// index of the quadratic feature pi_i*pi_j, i <= j (itertools.combinations_with_replacement order)
__host__ __device__ constexpr int quad_index(int d, int i, int j) { return i * d - (i * (i - 1)) / 2 + (j - i); }

// math of the float stream = the MUFU-level variants (same accuracy class, 2-3x fewer instructions);
// the double stream keeps libm-accurate calls for the exact-parity mode
template <typename R> struct StreamMath;
template <> struct StreamMath<float> {
    static __device__ __forceinline__ void alpha(float th, float x, float& a, float& d) { policy_alpha_fast(th, x, a, d); }
    static __device__ __forceinline__ float psi(float x) { return digamma_fast(x); }
    // lg2.approx.ftz flushes a denormal argument to zero (-inf): a denormal P > 0 (a Gamma variate far below the row
    // sum, shapes << 1) is scaled into the normal range first -- ln(p) = ln(p 2^64) - 64 ln 2
    static __device__ __forceinline__ float lnp(float p) {
        if (p >= 1.17549435e-38f) return lg2_approx(p) * DMFG_LN2;
        return p > 0.0f ? (lg2_approx(p * 1.8446744073709552e19f) - 64.0f) * DMFG_LN2 : -230.25850929940458f;
    }
};
template <> struct StreamMath<double> {
    static __device__ __forceinline__ void alpha(double th, double x, double& a, double& d) { policy_alpha<double>(th, x, a, d); }
    static __device__ __forceinline__ double psi(double x) { return digamma(x); }
    static __device__ __forceinline__ double lnp(double p) { return log_prob(p); }
};

template <typename R>
struct RolloutParams {
    int d, T;
    long long B, pop_offset;
    double theta;
    const double* theta_dev;
    double shift, alpha_scale, gamma;
    int reward_kind, discount_kind;
    const R* noise_y;
    unsigned long long seed, step_offset;
    const unsigned long long* step_offset_dev;   // optional device scalar ADDED to step_offset (CUDA-graph replays)
    const R* pi0;
    const double* w;
    const R* rewards_in;
    R *states, *actions, *alpha, *alpha_deriv, *rewards, *deltas, *grads, *pi_final;
    double* partials;      // [gridDim.x][2+F] per-CTA sums (fast variant with w)
    PhiloxKeys rk;         // round keys of `seed` (v2 kernel)
    float shift_f, scale_f;
    // fused per-step update (dmfg_ac_step): the last CTA to finish reduces the per-CTA partials in fixed order and
    // applies theta += lr_a * scale * sum delta g, w += lr_c * scale * sum delta phi in the same launch
    unsigned int* fuse_counter;
    double* fuse_theta;
    double* fuse_w;
    double* fuse_acc;              // optional [2+F]: the step's reduced sums
    const double* fuse_lr_dev;     // optional {lr_critic_eff, lr_actor_eff} on the device
    double fuse_lr_c, fuse_lr_a, fuse_scale;
};

// first Philox step index of this launch: the by-value offset plus the optional device-side one
template <typename R>
__device__ __forceinline__ unsigned long long step_base(const RolloutParams<R>& p) {
    return p.step_offset + (p.step_offset_dev != nullptr ? *p.step_offset_dev : 0ull);
}

// ---------------------------------------------------------------------------
// group collectives over G lanes (G power of two <= 32)
// ---------------------------------------------------------------------------
// c[j] holds this lane's contribution to element j; on return lane r has sum_i c_i[r].
template <int G>
__device__ __forceinline__ double reduce_scatter(double (&c)[G], int r) {
#pragma unroll
    for (int h = G / 2; h >= 1; h >>= 1) {
        const bool up = (r & h) != 0;
#pragma unroll
        for (int k = 0; k < h; ++k) {
            const double send = up ? c[k] : c[k + h];
            const double keep = up ? c[k + h] : c[k];
            c[k] = keep + __shfl_xor_sync(0xffffffffu, send, h, G);
        }
    }
    return c[0];
}
template <int G>
__device__ __forceinline__ void all_gather(double self, double (&out)[G]) {
#pragma unroll
    for (int j = 0; j < G; ++j) out[j] = __shfl_sync(0xffffffffu, self, j, G);
}

// ---------------------------------------------------------------------------
// One transition of one population, row-per-lane.  Everything the two fast
// kernels share: a1 (sample_action), a2 (pi' = P^T pi), a3 (closed-form
// reward), a6 (d log F / d theta).
// ---------------------------------------------------------------------------
template <int D, int G, typename R, int NOISE>
struct StepCore {
    static constexpr int PD = (D + 1) / 2;

    // in:  pi[G] (replicated state, zero beyond D), pi_self (= pi[r], 0 for idle lanes r >= D)
    // out: pi_next[G], next_self, reward (group-uniform), grad (group-uniform)
    //      P row is streamed to `act_row`, alpha / alpha' to `alpha_row` / `deriv_row` when non-null
    static __device__ __forceinline__ void run(
        const double (&pi)[G], double pi_self, int r, R theta, R shift, R alpha_scale, int reward_kind,
        const R* __restrict__ y_row, const NoiseKey& nk, uint32_t step, R* __restrict__ act_row,
        R* __restrict__ alpha_row, R* __restrict__ deriv_row, double (&pi_next)[G], double& next_self,
        double& reward, double& grad) {
        const bool row_ok = r < D;
        R yv[2 * PD], dv[2 * PD];
        double ysum = 0.0, asum = 0.0, dsum = 0.0, g1 = 0.0;
#pragma unroll
        for (int p = 0; p < PD; ++p) {
            R a[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 2 * p + e;
                if (j < D) {
                    const R x = (R)(pi[j] - pi_self) - shift;
                    StreamMath<R>::alpha(theta, x, a[e], dv[j]);
                    asum += (double)a[e];
                    dsum += (double)dv[j];
                    g1 -= (double)(StreamMath<R>::psi(a[e]) * dv[j]);
                    if (alpha_row != nullptr && row_ok) { alpha_row[j] = a[e]; deriv_row[j] = dv[j]; }
                } else {
                    a[e] = R(1);
                    dv[j] = R(0);
                }
            }
            if (NOISE == DMFG_NOISE_PHILOX) {
                float y0, y1;
                gamma_pair(nk, gamma_slot(step, D, r, p), (float)a[0], (float)a[1], (float)alpha_scale, y0, y1);
                yv[2 * p] = (R)y0;
                yv[2 * p + 1] = (R)y1;
            } else {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = 2 * p + e;
                    yv[j] = (j < D && row_ok) ? y_row[j] : R(1);
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 2 * p + e;
                if (j < D) {
                    if (NOISE != DMFG_NOISE_ACTIONS && yv[j] == R(0)) yv[j] = R(1e-20);   // mfg_ac2.py:244
                    ysum += (double)yv[j];
                }
            }
        }
        const double inv = NOISE == DMFG_NOISE_ACTIONS ? 1.0 : 1.0 / ysum;
        const R psi_row = StreamMath<R>::psi((R)asum);
        double c[G];
        double racc = 0.0, g2 = 0.0;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            if (j < D) {
                const double P = (double)yv[j] * inv;
                const R Pr = (R)P;
                g2 += (double)(StreamMath<R>::lnp(Pr) * dv[j]);
                c[j] = pi_self * P;
                if (reward_kind == DMFG_REWARD_AC2) racc += P * P * (pi[j] - pi_self);
                else racc += P * P;
                if (act_row != nullptr && row_ok) act_row[j] = Pr;
            } else {
                c[j] = 0.0;
            }
        }
        double rew = 0.0;
        if (reward_kind == DMFG_REWARD_AC2) rew = pi_self * racc;
        else if (reward_kind == DMFG_REWARD_SYNTHETIC) rew = -0.5 * pi_self * racc;
        const double glane = row_ok ? (g1 + g2 + (double)psi_row * dsum) : 0.0;
        next_self = reduce_scatter<G>(c, r);
        all_gather<G>(next_self, pi_next);
        reward = group_sum<G>(rew);
        grad = group_sum<G>(glane);
    }
};

// per-lane critic slots: k < D -> quadratic (r,k) (valid for k >= r); D -> linear r; D+1 -> bias (lane 0)
template <int D, int G>
__device__ __forceinline__ double critic_value(const double* __restrict__ wl, int stride,
                                               const double (&pi)[G], double pi_self) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) v = fma(wl[k * stride], pi_self * pi[k], v);
    v = fma(wl[D * stride], pi_self, v);
    v += wl[(D + 1) * stride];
    return group_sum<G>(v);
}

template <int D>
__device__ __forceinline__ void stage_critic_slots(double* wl, int stride, const double* __restrict__ w, int r) {
    constexpr int Q = D * (D + 1) / 2;
#pragma unroll
    for (int k = 0; k < D; ++k) wl[k * stride] = (r < D && k >= r) ? w[quad_index(D, r, k)] : 0.0;
    wl[D * stride] = (r < D) ? w[Q + r] : 0.0;
    wl[(D + 1) * stride] = (r == 0) ? w[Q + D] : 0.0;
}

// ---------------------------------------------------------------------------
// FAST rollout kernel (frozen parameters)
// ---------------------------------------------------------------------------
template <int D, int G, typename R, int NOISE>
__global__ void __launch_bounds__(kFastThreads)
rollout_fast_kernel(const RolloutParams<R> p) {
    constexpr int NT = kFastThreads;
    constexpr int GPB = NT / G;
    constexpr int NSLOT = D + 2;
    constexpr int F = num_features_c(D);
    extern __shared__ double smem[];
    double* wl = smem + threadIdx.x;                  // [NSLOT][NT]
    double* acc = smem + NSLOT * NT + threadIdx.x;    // [NSLOT][NT]
    const int r = threadIdx.x % G;
    const int grp = threadIdx.x / G;
    const bool td = p.w != nullptr;
    const bool want_acc = td && p.partials != nullptr;
    if (td) {
        stage_critic_slots<D>(wl, NT, p.w, r);
#pragma unroll
        for (int k = 0; k < NSLOT; ++k) acc[k * NT] = 0.0;
    }
    const R theta = (R)(p.theta_dev ? *p.theta_dev : p.theta);
    const R shift = (R)p.shift, scale = (R)p.alpha_scale;
    double sum_dg = 0.0, sum_r = 0.0;
    const long long ntiles = (p.B + GPB - 1) / GPB;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        long long b = tile * GPB + grp;
        const bool live = b < p.B;                    // dead groups shadow the last population, writes masked
        if (!live) b = p.B - 1;
        const bool row_ok = r < D;
        const bool wr = live && row_ok;
        const NoiseKey nk = make_noise_key(p.seed, (unsigned long long)(p.pop_offset + b));
        double pi_self = row_ok ? (double)p.pi0[b * D + r] : 0.0;
        double pi[G];
        all_gather<G>(pi_self, pi);
        double v_cur = td ? critic_value<D, G>(wl, NT, pi, pi_self) : 0.0;
        double disc = 1.0;
        if (p.states != nullptr && wr) p.states[b * D + r] = (R)pi_self;
        for (int t = 0; t < p.T; ++t) {
            const long long tb = (long long)t * p.B + b;
            const long long row = (tb * D + r) * D;
            double pi_next[G], next_self, rew, grad;
            StepCore<D, G, R, NOISE>::run(
                pi, pi_self, r, theta, shift, scale, p.reward_kind,
                NOISE != DMFG_NOISE_PHILOX ? p.noise_y + row : nullptr, nk,
                (uint32_t)(step_base(p) + t),
                (p.actions && live) ? p.actions + row : nullptr,
                (p.alpha && live) ? p.alpha + row : nullptr,
                (p.alpha_deriv && live) ? p.alpha_deriv + row : nullptr,
                pi_next, next_self, rew, grad);
            if (p.rewards_in != nullptr) rew = (double)p.rewards_in[tb];
            if (td) {
                const double v_next = critic_value<D, G>(wl, NT, pi_next, next_self);
                const double gfac = p.discount_kind == DMFG_DISCOUNT_STEP ? p.gamma : disc;
                const double delta = rew + gfac * v_next - v_cur;
                if (want_acc && live) {
                    const double dp = delta * pi_self;
#pragma unroll
                    for (int k = 0; k < D; ++k) acc[k * NT] = fma(dp, pi[k], acc[k * NT]);
                    acc[D * NT] += dp;
                    acc[(D + 1) * NT] += delta;
                    if (r == 0) sum_dg = fma(delta, grad, sum_dg);
                }
                if (p.deltas != nullptr && live && r == 0) p.deltas[tb] = (R)delta;
                v_cur = v_next;
            }
            if (live && r == 0) {
                sum_r += rew;
                if (p.rewards != nullptr) p.rewards[tb] = (R)rew;
                if (p.grads != nullptr) p.grads[tb] = (R)grad;
            }
            disc *= p.gamma;
            pi_self = next_self;
#pragma unroll
            for (int j = 0; j < G; ++j) pi[j] = pi_next[j];
            if (p.states != nullptr && wr) p.states[(tb + p.B) * D + r] = (R)pi_self;
        }
        if (p.pi_final != nullptr && wr) p.pi_final[b * D + r] = (R)pi_self;
    }
    if (!want_acc) return;
    // ---- per-CTA partial sums, fixed order (deterministic) ------------------
    __shared__ double red[2][NT / 32];
    {
        double a = sum_dg, c = sum_r;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = c; }
    }
    __syncthreads();
    double* out = p.partials + (long long)blockIdx.x * (2 + F);
    const double* accbase = smem + NSLOT * NT;
    for (int f = threadIdx.x; f < F; f += NT) {
        int row, k;
        constexpr int Q = D * (D + 1) / 2;
        if (f < Q) {
            row = 0;
            int rem = f;
            while (rem >= D - row) { rem -= D - row; ++row; }
            k = row + rem;
        } else if (f < Q + D) {
            row = f - Q; k = D;
        } else {
            row = 0; k = D + 1;
        }
        double s = 0.0;
        for (int g = 0; g < GPB; ++g) s += accbase[k * NT + g * G + row];
        out[1 + f] = s;
    }
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int wv = 0; wv < NT / 32; ++wv) { a += red[0][wv]; c += red[1][wv]; }
        out[0] = a;
        out[1 + F] = c;
    }
}

// acc[f] = sum over CTAs of partials[cta][f], in CTA order
static __global__ void reduce_partials_kernel(const double* __restrict__ partials, int ncta, int n, double* __restrict__ acc) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    // four independent chains with a fixed assignment and combination order: deterministic, loads pipelined
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = 0;
#pragma unroll 2
    for (; c + 3 < ncta; c += 4) {
        s0 += partials[(long long)c * n + f];
        s1 += partials[(long long)(c + 1) * n + f];
        s2 += partials[(long long)(c + 2) * n + f];
        s3 += partials[(long long)(c + 3) * n + f];
    }
    for (; c < ncta; ++c) s0 += partials[(long long)c * n + f];
    acc[f] = (s0 + s1) + (s2 + s3);
}

// ---------------------------------------------------------------------------
// GENERIC rollout kernel: warp per population, runtime d.  No critic here --
// TD errors for generic d come from td_delta_kernel / td_gw_kernel below.
// ---------------------------------------------------------------------------
template <typename R, int NOISE>
__global__ void __launch_bounds__(kGenericThreads)
rollout_generic_kernel(const RolloutParams<R> p) {
    extern __shared__ double smem[];
    const int d = p.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int WPB = kGenericThreads / 32;
    // per warp: pi[d], pi_next[d] (double) then y[d], deriv[d] (R)
    double* pi_s = smem + (size_t)warp * (2 * d) ;
    double* nx_s = pi_s + d;
    R* y_s = reinterpret_cast<R*>(smem + (size_t)WPB * 2 * d) + (size_t)warp * 2 * d;
    R* dv_s = y_s + d;
    const int pd = (d + 1) >> 1;
    const R theta = (R)(p.theta_dev ? *p.theta_dev : p.theta);
    const R shift = (R)p.shift, scale = (R)p.alpha_scale;
    for (long long b = (long long)blockIdx.x * WPB + warp; b < p.B; b += (long long)gridDim.x * WPB) {
        const NoiseKey nk = make_noise_key(p.seed, (unsigned long long)(p.pop_offset + b));
        for (int j = lane; j < d; j += 32) {
            const R v = p.pi0[b * d + j];
            pi_s[j] = (double)v;
            if (p.states) p.states[b * d + j] = v;
        }
        __syncwarp();
        for (int t = 0; t < p.T; ++t) {
            const long long tb = (long long)t * p.B + b;
            for (int j = lane; j < d; j += 32) nx_s[j] = 0.0;
            double racc = 0.0, gacc = 0.0;
            __syncwarp();
            for (int i = 0; i < d; ++i) {
                const double pi_i = pi_s[i];
                const long long row = (tb * d + i) * d;
                double ysum = 0.0, asum = 0.0, dsum = 0.0;
                for (int pp = lane; pp < pd; pp += 32) {
                    R a[2], dv[2], yv[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = 2 * pp + e;
                        if (j < d) {
                            const R x = (R)(pi_s[j] - pi_i) - shift;
                            StreamMath<R>::alpha(theta, x, a[e], dv[e]);
                            asum += (double)a[e];
                            dsum += (double)dv[e];
                            gacc -= (double)(StreamMath<R>::psi(a[e]) * dv[e]);
                            if (p.alpha) { p.alpha[row + j] = a[e]; p.alpha_deriv[row + j] = dv[e]; }
                        } else {
                            a[e] = R(1); dv[e] = R(0);
                        }
                    }
                    if (NOISE == DMFG_NOISE_PHILOX) {
                        float y0, y1;
                        gamma_pair_fast(nk, p.rk, gamma_slot((uint32_t)(step_base(p) + t), d, i, pp),
                                        make_float2((float)a[0], (float)a[1]), (float)scale, y0, y1);
                        yv[0] = (R)y0; yv[1] = (R)y1;
                    } else {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int j = 2 * pp + e;
                            yv[e] = j < d ? p.noise_y[row + j] : R(1);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = 2 * pp + e;
                        if (j < d) {
                            if (NOISE != DMFG_NOISE_ACTIONS && yv[e] == R(0)) yv[e] = R(1e-20);
                            ysum += (double)yv[e];
                            y_s[j] = yv[e];
                            dv_s[j] = dv[e];
                        }
                    }
                }
                ysum = group_sum<32>(ysum);
                asum = group_sum<32>(asum);
                dsum = group_sum<32>(dsum);
                const double inv = NOISE == DMFG_NOISE_ACTIONS ? 1.0 : 1.0 / ysum;
                if (lane == 0) gacc += (double)StreamMath<R>::psi((R)asum) * dsum;
                for (int pp = lane; pp < pd; pp += 32) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = 2 * pp + e;
                        if (j < d) {
                            const double P = (double)y_s[j] * inv;
                            const R Pr = (R)P;
                            gacc += (double)(StreamMath<R>::lnp(Pr) * dv_s[j]);
                            nx_s[j] = fma(pi_i, P, nx_s[j]);
                            if (p.reward_kind == DMFG_REWARD_AC2) racc += pi_i * P * P * (pi_s[j] - pi_i);
                            else if (p.reward_kind == DMFG_REWARD_SYNTHETIC) racc -= 0.5 * pi_i * P * P;
                            if (p.actions) p.actions[row + j] = Pr;
                        }
                    }
                }
                __syncwarp();
            }
            const double rew = group_sum<32>(racc);
            const double grad = group_sum<32>(gacc);
            for (int j = lane; j < d; j += 32) {
                const double v = nx_s[j];
                pi_s[j] = v;
                if (p.states) p.states[(tb + p.B) * d + j] = (R)v;
            }
            if (lane == 0) {
                if (p.rewards) p.rewards[tb] = (R)rew;
                if (p.grads) p.grads[tb] = (R)grad;
            }
            __syncwarp();
        }
        if (p.pi_final)
            for (int j = lane; j < d; j += 32) p.pi_final[b * d + j] = (R)pi_s[j];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// TD errors and accumulators from recorded trajectories (any d)
// ---------------------------------------------------------------------------
template <typename R>
struct TdParams {
    int d, T;
    long long B;
    double gamma;
    int discount_kind;
    const R *states, *rewards, *grads;
    const double* w;
    R* deltas;            // user output (may be null)
    double* delta_buf;    // [T][B] double scratch consumed by td_gw_kernel (may be null)
    double* partials;     // [gridDim.x][2+F]
};

// warp per population: V(pi_t) for t = 0..T, then delta_t (mfg_ac2.py:505 / ac_irl.py:691)
template <typename R>
__global__ void __launch_bounds__(128) td_delta_kernel(const TdParams<R> p) {
    const int d = p.d, lane = threadIdx.x & 31;
    const int Q = d * (d + 1) / 2;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long b = warp; b < p.B; b += nwarp) {
        double v_prev = 0.0, disc = 1.0;
        for (int t = 0; t <= p.T; ++t) {
            const R* s = p.states + ((long long)t * p.B + b) * d;
            double v = 0.0;
            for (int i = lane; i < d; i += 32) {
                const double si = (double)s[i];
                const double* wq = p.w + quad_index(d, i, i);
                double a = p.w[Q + i];
                for (int j = i; j < d; ++j) a = fma(wq[j - i], (double)s[j], a);
                v = fma(a, si, v);
            }
            v = group_sum<32>(v) + p.w[Q + d];
            if (t > 0 && lane == 0) {
                const long long tb = (long long)(t - 1) * p.B + b;
                const double gfac = p.discount_kind == DMFG_DISCOUNT_STEP ? p.gamma : disc;
                const double delta = (double)p.rewards[tb] + gfac * v - v_prev;
                if (p.deltas) p.deltas[tb] = (R)delta;
                if (p.delta_buf) p.delta_buf[tb] = delta;
                disc *= p.gamma;
            }
            v_prev = v;
        }
    }
}

// CTA per chunk of transitions, thread per feature: partial sums of delta*phi, delta*g, r
// skip_quad: the quadratic features are accumulated elsewhere (td_gram_dmma_kernel), leave their slots at 0
template <typename R>
__global__ void __launch_bounds__(256) td_gw_kernel(const TdParams<R> p, int chunk, int skip_quad) {
    extern __shared__ double smem[];
    const int d = p.d;
    const int F = num_features_c(d), Q = d * (d + 1) / 2;
    const long long N = (long long)p.T * p.B;
    double* st = smem;                    // [chunk][d]
    double* dl = smem + (size_t)chunk * d;   // [chunk] delta, then [chunk] delta*g, then [chunk] r
    const int nfeat_thread = (F + 2 + blockDim.x - 1) / blockDim.x;
    double* out = p.partials + (long long)blockIdx.x * (2 + F);
    for (int f = threadIdx.x; f < F + 2; f += blockDim.x) out[f] = 0.0;
    (void)nfeat_thread;
    for (long long n0 = (long long)blockIdx.x * chunk; n0 < N; n0 += (long long)gridDim.x * chunk) {
        const int cnt = (int)min((long long)chunk, N - n0);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * d; e += blockDim.x) st[e] = (double)p.states[n0 * d + e];
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            const double dlt = p.delta_buf[n0 + e];
            dl[e] = dlt;
            dl[chunk + e] = dlt * (double)p.grads[n0 + e];
            dl[2 * chunk + e] = (double)p.rewards[n0 + e];
        }
        __syncthreads();
        for (int f = threadIdx.x; f < F + 2; f += blockDim.x) {
            double s = 0.0;
            if (skip_quad && f >= 1 && f <= Q) continue;
            if (f == 0) {
                for (int e = 0; e < cnt; ++e) s += dl[chunk + e];
            } else if (f == F + 1) {
                for (int e = 0; e < cnt; ++e) s += dl[2 * chunk + e];
            } else {
                const int ff = f - 1;
                if (ff < Q) {
                    int i = 0, rem = ff;
                    while (rem >= d - i) { rem -= d - i; ++i; }
                    const int j = i + rem;
                    for (int e = 0; e < cnt; ++e) s = fma(dl[e], st[e * d + i] * st[e * d + j], s);
                } else if (ff < Q + d) {
                    const int i = ff - Q;
                    for (int e = 0; e < cnt; ++e) s = fma(dl[e], st[e * d + i], s);
                } else {
                    for (int e = 0; e < cnt; ++e) s += dl[e];
                }
            }
            out[f] += s;
        }
    }
}

// calc_features / calc_value for N states: thread per (state, feature)
template <typename R>
__global__ void critic_features_kernel(int d, long long N, const R* __restrict__ states, R* __restrict__ features) {
    const int F = num_features_c(d), Q = d * (d + 1) / 2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * F) return;
    const long long n = idx / F;
    const int f = (int)(idx - n * F);
    const R* s = states + n * d;
    R v;
    if (f < Q) {
        int i = 0, rem = f;
        while (rem >= d - i) { rem -= d - i; ++i; }
        v = s[i] * s[i + rem];
    } else if (f < Q + d) {
        v = s[f - Q];
    } else {
        v = R(1);
    }
    features[idx] = v;
}
// warp per state: V = phi(pi) . w accumulated in double
template <typename R>
__global__ void critic_value_kernel(int d, long long N, const R* __restrict__ states, const double* __restrict__ w,
                                    R* __restrict__ values) {
    const int lane = threadIdx.x & 31, Q = d * (d + 1) / 2;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= N) return;
    const R* s = states + warp * d;
    double v = 0.0;
    for (int i = lane; i < d; i += 32) {
        const double* wq = w + quad_index(d, i, i);
        double a = w[Q + i];
        for (int j = i; j < d; ++j) a = fma(wq[j - i], (double)s[j], a);
        v = fma(a, (double)s[i], v);
    }
    v = group_sum<32>(v) + w[Q + d];
    if (lane == 0) values[warp] = (R)v;
}

// L1 distance and Jensen-Shannon divergence between generated and empirical distributions, one warp per
// (trajectory, hour): mfg_ac2.py:546-563 (JSD: zeros -> 1e-100, M from the unnormalised inputs, entropy()
// normalises each argument) and :627-650.  Element (b,h,j) of X sits at X[b*sb + h*sh + j].
template <typename R>
__global__ void __launch_bounds__(128) traj_metrics_kernel(int d, long long B, int H, const R* __restrict__ gen,
                                                           long long gsb, long long gsh,
                                                           const R* __restrict__ emp, long long esb, long long esh,
                                                           double* __restrict__ l1, double* __restrict__ jsd) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= B * H) return;
    const long long b = warp / H;
    const int h = (int)(warp - b * H);
    const R* g = gen + b * gsb + h * gsh;
    const R* e = emp + b * esb + h * esh;
    double sl1 = 0.0, sp = 0.0, sq = 0.0;
    for (int j = lane; j < d; j += 32) {
        const double p = (double)g[j], q = (double)e[j];
        sl1 += fabs(q - p);
        sp += p == 0.0 ? 1e-100 : p;
        sq += q == 0.0 ? 1e-100 : q;
    }
    sl1 = group_sum<32>(sl1); sp = group_sum<32>(sp); sq = group_sum<32>(sq);
    const double sm = 0.5 * (sp + sq);
    double kp = 0.0, kq = 0.0;
    for (int j = lane; j < d; j += 32) {
        double p = (double)g[j], q = (double)e[j];
        p = p == 0.0 ? 1e-100 : p;
        q = q == 0.0 ? 1e-100 : q;
        const double m = 0.5 * (p + q) / sm;
        p /= sp; q /= sq;
        kp += p * log(p / m);
        kq += q * log(q / m);
    }
    kp = group_sum<32>(kp); kq = group_sum<32>(kq);
    if (lane == 0) {
        if (l1) l1[warp] = sl1;
        if (jsd) jsd[warp] = 0.5 * (kp + kq);
    }
}

// Analytic check of a policy against the MFG backward equation (mfg_synthetic.py:726-899), warp per
// trajectory.  V^T = 0;  V^n = r(P^n) + P^n V^{n+1},  r_i = -1/2 ||P_i||^2;  then per step n the predicted
// matrix A_ij = V_j - V_i (i != j), A_ii = 1 - (sum_j V_j - d V_i):  l1[b][n] = sum_ij |P_ij - A_ij|,
// jsd[b][n] = sum_i JSD(P_i, A_i) with mfg_synthetic.py:529-546's JSD (entries <= 0 -> 1e-100).
// actions: time-major record [T][B][d][d];  shared memory: 2 d doubles per warp.
template <typename R>
__global__ void __launch_bounds__(128) synthetic_check_kernel(int d, long long B, int T, const R* __restrict__ actions,
                                                              double* __restrict__ l1, double* __restrict__ jsd) {
    extern __shared__ double vsm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (b >= B) return;
    double* Vn = vsm + (size_t)wib * 2 * d;
    double* Vo = Vn + d;
    for (int j = lane; j < d; j += 32) Vo[j] = 0.0;
    __syncwarp();
    for (int n = T - 1; n >= 0; --n) {
        const R* P = actions + ((long long)n * B + b) * d * d;
        for (int i = 0; i < d; ++i) {
            double s2 = 0.0, sv = 0.0;
            for (int j = lane; j < d; j += 32) {
                const double p = (double)P[i * d + j];
                s2 = fma(p, p, s2);
                sv = fma(p, Vo[j], sv);
            }
            s2 = group_sum<32>(s2); sv = group_sum<32>(sv);
            if (lane == 0) Vn[i] = -0.5 * s2 + sv;
        }
        __syncwarp();
        double sV = 0.0;
        for (int j = lane; j < d; j += 32) sV += Vn[j];
        sV = group_sum<32>(sV);
        double dl1 = 0.0, djs = 0.0;
        for (int i = 0; i < d; ++i) {
            const double vi = Vn[i];
            double a = 0.0, sp = 0.0, sq = 0.0;
            for (int j = lane; j < d; j += 32) {
                const double p = (double)P[i * d + j];
                const double q = (i == j) ? 1.0 - (sV - d * vi) : Vn[j] - vi;
                a += fabs(p - q);
                sp += p <= 0.0 ? 1e-100 : p;
                sq += q <= 0.0 ? 1e-100 : q;
            }
            a = group_sum<32>(a); sp = group_sum<32>(sp); sq = group_sum<32>(sq);
            dl1 += a;
            if (jsd != nullptr) {
                const double sm = 0.5 * (sp + sq);
                double kp = 0.0, kq = 0.0;
                for (int j = lane; j < d; j += 32) {
                    double p = (double)P[i * d + j];
                    double q = (i == j) ? 1.0 - (sV - d * vi) : Vn[j] - vi;
                    p = p <= 0.0 ? 1e-100 : p;
                    q = q <= 0.0 ? 1e-100 : q;
                    const double m = 0.5 * (p + q) / sm;
                    p /= sp; q /= sq;
                    kp += p * log(p / m);
                    kq += q * log(q / m);
                }
                djs += 0.5 * (group_sum<32>(kp) + group_sum<32>(kq));
            }
        }
        if (lane == 0) {
            if (l1 != nullptr) l1[b * T + n] = dl1;
            if (jsd != nullptr) jsd[b * T + n] = djs;
        }
        __syncwarp();
        double* tmp = Vn; Vn = Vo; Vo = tmp;
    }
}

// theta += lr_a*scale*acc[0];  w[f] += lr_c*scale*acc[1+f]   (mfg_ac2.py:511-522)
static __global__ void ac_apply_update_kernel(int F, double* theta, double* w, const double* __restrict__ acc,
                                       double lr_c, double lr_a, double scale) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < F) w[f] = fma(lr_c * scale, acc[1 + f], w[f]);
    if (f == 0 && theta != nullptr) *theta = fma(lr_a * scale, acc[0], *theta);
}

// the same update with the two effective step sizes read from device memory (lr[0] critic, lr[1] actor): the
// launch arguments do not change between replays of a captured CUDA graph
static __global__ void ac_apply_update_dev_kernel(int F, double* theta, double* w, const double* __restrict__ acc,
                                           const double* __restrict__ lr, double scale) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < F) w[f] = fma(lr[0] * scale, acc[1 + f], w[f]);
    if (f == 0 && theta != nullptr) *theta = fma(lr[1] * scale, acc[0], *theta);
}

// ---------------------------------------------------------------------------
// Independent serial learners: one group = one learner with private (theta, w)
// and per-step online updates -- mfg_ac2.py:448-539 semantics exactly.
// ---------------------------------------------------------------------------
template <typename R>
struct LearnerParams {
    int d, T, E, episode0, S;
    long long L, learner_offset;
    double *theta, *w;
    const double *shift, *alpha_scale;
    double shift_scalar, alpha_scale_scalar, gamma, lr_critic, lr_actor;
    int constant_lr, reward_kind, discount_kind;
    const R* mat_pi0;
    const int* start_rows;
    const R* noise_y;
    unsigned long long seed;
    long long noise_episode_offset;
    double *theta_trace, *delta_trace, *total_reward;
    R* pi_final;
};

template <int D, int G, typename R, int NOISE>
__global__ void __launch_bounds__(kFastThreads)
learners_fast_kernel(const LearnerParams<R> p) {
    constexpr int NT = kFastThreads;
    constexpr int GPB = NT / G;
    constexpr int F = num_features_c(D);
    extern __shared__ double smem[];
    double* wl = smem + threadIdx.x;      // [NSLOT][NT] private critic weights
    const int r = threadIdx.x % G;
    const int grp = threadIdx.x / G;
    const bool row_ok = r < D;
    long long l = (long long)blockIdx.x * GPB + grp;
    const bool live = l < p.L;
    if (!live) l = p.L - 1;
    stage_critic_slots<D>(wl, NT, p.w + l * F, r);
    double theta = p.theta[l];
    const R shift = (R)(p.shift ? p.shift[l] : p.shift_scalar);
    const R scale = (R)(p.alpha_scale ? p.alpha_scale[l] : p.alpha_scale_scalar);
    const NoiseKey nk = make_noise_key(p.seed, (unsigned long long)(p.learner_offset + l));
    double pi_self = 0.0;
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
        pi_self = row_ok ? (double)p.mat_pi0[(long long)start * D + r] : 0.0;
        double pi[G];
        all_gather<G>(pi_self, pi);
        const double lr_c = p.constant_lr ? p.lr_critic : p.lr_critic / (episode + 1.0);
        const double lr_a = p.constant_lr ? p.lr_actor
                                          : p.lr_actor / ((episode + 1.0) * log(log(episode + 20.0)));
        double disc = 1.0, total = 0.0;
        for (int t = 0; t < p.T; ++t) {
            const long long et = ((long long)l * p.E + e) * p.T + t;
            double pi_next[G], next_self, rew, grad;
            StepCore<D, G, R, NOISE>::run(
                pi, pi_self, r, (R)theta, shift, scale, p.reward_kind,
                NOISE == DMFG_NOISE_INJECTED ? p.noise_y + (et * D + r) * D : nullptr, nk,
                (uint32_t)((episode + p.noise_episode_offset) * p.T + t), nullptr, nullptr, nullptr,
                pi_next, next_self, rew, grad);
            const double v_next = critic_value<D, G>(wl, NT, pi_next, next_self);
            const double v_cur = critic_value<D, G>(wl, NT, pi, pi_self);
            const double gfac = p.discount_kind == DMFG_DISCOUNT_STEP ? p.gamma : disc;
            const double delta = rew + gfac * v_next - v_cur;
            // critic first, then actor, both with the same delta (mfg_ac2.py:505-522)
            const double step_w = lr_c * delta;
            const double dp = step_w * pi_self;
#pragma unroll
            for (int k = 0; k < D; ++k)
                if (k >= r) wl[k * NT] = fma(dp, pi[k], wl[k * NT]);
            if (row_ok) wl[D * NT] += dp;
            if (r == 0) wl[(D + 1) * NT] += step_w;
            theta = fma(lr_a * delta, grad, theta);
            if (live && r == 0) {
                if (p.theta_trace) p.theta_trace[et] = theta;
                if (p.delta_trace) p.delta_trace[et] = delta;
            }
            total += rew;
            disc *= p.gamma;
            pi_self = next_self;
#pragma unroll
            for (int j = 0; j < G; ++j) pi[j] = pi_next[j];
        }
        if (live && r == 0 && p.total_reward) p.total_reward[l * p.E + e] = total;
    }
    if (!live) return;
    if (r == 0) p.theta[l] = theta;
    if (p.pi_final && row_ok) p.pi_final[l * D + r] = (R)pi_self;
    // write the private critic weights back in the reference's feature order
    constexpr int Q = D * (D + 1) / 2;
    double* wout = p.w + l * F;
    if (row_ok) {
#pragma unroll
        for (int k = 0; k < D; ++k)
            if (k >= r) wout[quad_index(D, r, k)] = wl[k * NT];
        wout[Q + r] = wl[D * NT];
    }
    if (r == 0) wout[Q + D] = wl[(D + 1) * NT];
}

}  // namespace dmfg
