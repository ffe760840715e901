// rollout_v2_kernel -- the throughput version of the fused population step for d = 15 / 16, float
// streams (sm_100a).  Same arithmetic contract as rollout_fast_kernel (dmfg_rollout.cuh), restructured
// around what ncu showed (profiles/r1_rollout_fast_train_ncu_summary.md for the first kernel,
// profiles/r1_rollout_v2_ncu_summary.md for the two-pass form of this one):
//
//   * a group of 16 lanes owns one population, lane r owns ROW r of P, walked by ONE rolled loop over
//     column pairs (packed f32x2 math):  alpha, alpha', psi(alpha) (alpha_psi_fast2), the Gamma pair (one
//     Philox call, Box-Muller on MUFU, squeeze-accepted Marsaglia-Tsang), and -- because every row-wise
//     sum is linear in the UNNORMALISED variates --
//         sum_j y_ij,   sum_j y_ij^2 (pi_j - pi_i)   [reward],   sum_j alpha'_ij lg2 y_ij   [ln P term]
//     in the same pass; y goes to a [row][column] shared tile as a double.  There is no second pass
//     over the row: 1/s_i enters as a per-row factor afterwards
//         r_i = pi_i s_i^-2 sum_j y_ij^2 (pi_j - pi_i),   sum_j alpha'_ij ln P_ij = ln2 (sum_j alpha'_ij lg2 y_ij - lg2 s_i sum_j alpha'_ij);
//   * pi'_j = sum_i (pi_i / s_i) y_ij is a transposed read of the tile against the 16 published
//     q_i = pi_i / s_i (three independent partial sums: the chain was latency-bound);
//   * the TD error needs ONE 16-lane reduction per step: lanes carry their partial of V(pi) and
//     delta = sum_lanes (r_lane + gamma V'_lane - V_lane); sum delta*g is accumulated per lane (linear);
//   * the critic gradient  sum_n delta_n pi_n pi_n^T  is a real contraction over samples: it runs on the
//     FP64 tensor-core path (DMMA.8x8x4): every two steps the warp's 4 samples (2 populations x 2 steps)
//     are staged as A = delta*pi (16x4), B = pi (4x16) and three m8n8k4 MMAs update the upper-triangular
//     8x8 tiles of the 16x16 Gram matrix held in 6 double registers per thread for the whole kernel
//     (the previous form spent 10 % of the step on 15 LDS+DFMA+STS triples per lane);
//   * shared memory is addressed through explicit 32-bit shared-window addresses (ld/st.shared): the
//     generic-pointer form made the compiler rebuild the window base (S2UR SR_CgaCtaId + 3 uniform
//     ops) in front of every access of the rolled loop;
//   * float everywhere except where the TD error needs it: row sums, reward, flux, state, critic
//     value and all reductions stay double.
#pragma once
#include "dmfg_rollout.cuh"

namespace dmfg {

#ifndef DMFG_V2_UNROLL
#define DMFG_V2_UNROLL 4
#endif
#ifndef DMFG_V2_MINB
#define DMFG_V2_MINB 2
#endif
#ifndef DMFG_V2_THREADS
#define DMFG_V2_THREADS 256
#endif
constexpr int kV2Threads = DMFG_V2_THREADS;
constexpr int kV2Unroll = DMFG_V2_UNROLL;
constexpr int kV2G = 16;
// lanes per population: 16 for d <= 16 (two populations per warp), 32 for d in (16, 32] (one per warp)
constexpr int v2_group(int d) { return d <= 16 ? 16 : 32; }

template <int D>
struct V2Smem {
    static constexpr int G = v2_group(D);
    static constexpr int Slots = G + 1;       // doubles per tile row (G columns + 1 pad => conflict-free both ways)
    static constexpr int StageRow = G + 4;    // doubles per staged sample (G + 4 pad => conflict-free fragment loads)
    static constexpr int PW = 32 / G;         // populations per warp
    static constexpr int SPF = 4 / PW;        // steps per DMMA flush (4 staged samples per warp)
    static constexpr int NB = (D + 7) / 8;    // 8-wide blocks of the Gram matrix
    static constexpr int NBLK = NB * (NB + 1) / 2;
    static constexpr int GPB = kV2Threads / G;
    static constexpr int NW = kV2Threads / 32;
    static constexpr int NSLOT = D + 2;
    // offsets in doubles
    static constexpr int tile = 0;                                  // [GPB][G][G+1] y_ij, double
    static constexpr int pid = tile + GPB * G * Slots;              // [2][GPB][G] state, double
    static constexpr int qv = pid + 2 * GPB * G;                    // [GPB][G] q_i = pi_i / s_i
    static constexpr int wl = qv + GPB * G;                         // [NSLOT][G] critic slots
    static constexpr int pif = wl + NSLOT * G;                      // [2][GPB][G] state, float (GPB*G doubles)
    static constexpr int stage = pif + GPB * G;                     // [NW][2][4][G+4] staged samples (A, B)
    static constexpr int total = stage + NW * 2 * 4 * StageRow;
    // the end-of-kernel reduction reuses the tile ([NW][G][G+1] Gram tiles), the state ([GPB][G] linear) and q ([GPB] bias)
    static constexpr int gram = tile;
    static constexpr int lin = pid;
    static constexpr int bias = qv;
    static_assert(NW * G * Slots <= GPB * G * Slots, "reduction scratch must fit the tile");
    static_assert(8 * NB <= StageRow && D <= G, "staged samples cover the Gram blocks");
};

// ---- explicit shared-window accesses (32-bit addresses from __cvta_generic_to_shared) -----------------
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_f64x2(uint32_t a) {
    double2 v; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t a) {
    float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void warp_fence() { asm volatile("bar.warp.sync 0xffffffff;" ::: "memory"); }

// D += A(8x4, row) * B(4x8, col) on the FP64 tensor-core path: thread (g = lane/4, t = lane%4) holds
// A[g][t], B[t][g] and C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// this lane's partial of V(pi): lane r sums w[r,k] pi_r pi_k over k >= r (zero slots below), + linear + bias
template <int D, int GW = kV2G>
__device__ __forceinline__ double critic_partial_v2(uint32_t a_wl_r, uint32_t a_pid, double pi_self) {
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
#pragma unroll
    for (int k = 0; k + 1 < D; k += 2) {
        const double2 pk = lds_f64x2(a_pid + 8 * k);
        const double w0 = lds_f64(a_wl_r + 8 * GW * k), w1 = lds_f64(a_wl_r + 8 * GW * (k + 1));
        if ((k / 2) % 3 == 0) { v0 = fma(w0, pk.x, v0); v1 = fma(w1, pk.y, v1); }
        else if ((k / 2) % 3 == 1) { v2 = fma(w0, pk.x, v2); v0 = fma(w1, pk.y, v0); }
        else { v1 = fma(w0, pk.x, v1); v2 = fma(w1, pk.y, v2); }
    }
    if (D & 1) v2 = fma(lds_f64(a_wl_r + 8 * GW * (D - 1)), lds_f64(a_pid + 8 * (D - 1)), v2);
    const double v = (v0 + v1) + v2;
    return fma(v, pi_self, lds_f64(a_wl_r + 8 * GW * D) * pi_self) + lds_f64(a_wl_r + 8 * GW * (D + 1));
}

// ---------------------------------------------------------------------------
// The single-pass walk over ONE row of P (lane = row), shared by rollout_v2_kernel and learners_v2_kernel:
// per column pair alpha, alpha', psi(alpha) (GRAD), the Gamma pair, and the row sums that are linear in the
// unnormalised variates; y goes to this lane's row of the shared tile as doubles.
//   ysum = sum_j y_ij,  racc = sum_j y_ij^2 (c1 pi_j + c0)  [reward],  asum = sum_j alpha_ij,  dsum = sum_j alpha'_ij,
//   g1 = -sum_j psi(alpha_ij) alpha'_ij,  g2 = sum_j alpha'_ij lg2 y_ij
// a_pfc / a_pic: shared addresses of the current state (float / double copies); noise_row: injected variates of
// this row (NOISE == INJECTED); alpha_row / deriv_row: optional record of alpha, alpha' (REC).
// ---------------------------------------------------------------------------
struct V2RowSums {
    double ysum, racc;
    float asum, dsum, g1, g2;
};

// The walk is written as two phases per column pair so that several pairs can be in flight at once:
//   phase A (branch-free): alpha, alpha', psi(alpha) and the alpha sums -- long MUFU / FFMA2 chains the compiler can
//                          interleave freely across the pairs of a chunk;
//   phase B: the Gamma pair (one rarely taken out-of-line branch per pair), lg2 y, the row sums, the tile store.
// DMFG_V2_CHUNK pairs run phase A back to back (their alpha / alpha' stay in registers), then phase B; chunk = 1 is the
// plain interleaved loop.  The order of every accumulation is the same for every chunk size: results do not depend on it.
#ifndef DMFG_V2_CHUNK
#define DMFG_V2_CHUNK 1
#endif
template <int D, int NOISE, bool GRAD, bool REC>
struct V2Walk {
    static constexpr int PD = (D + 1) / 2;
    float theta, xi, scale;
    double c0, c1;
    bool has_reward, row_ok;
    uint32_t a_pfc, a_pic, a_row, slot0;
    const float* __restrict__ noise_row;
    float* __restrict__ alpha_row;
    float* __restrict__ deriv_row;
    float2 asum2, dsum2, g12, g22;
    double ysum0, ysum1, racc0, racc1;
    float a_last;

    __device__ __forceinline__ void phase_a(int pp, float2& a, float2& dv) {
        const float2 pj = lds_f32x2(a_pfc + 8 * pp);
        const bool ok1 = (2 * pp + 1) < D;                     // only the last pair of an odd D
        float2 psi;
        if (GRAD) {
            alpha_psi_fast2(theta, __fadd2_rn(pj, splat2(-xi)), a, dv, psi);
        } else {
            policy_alpha_fast2(theta, __fadd2_rn(pj, splat2(-xi)), a, dv);
            psi = make_float2(0.f, 0.f);
        }
        // the phantom column of an odd D (state slot D is 0): alpha' = 0 removes it from every weighted
        // sum, its alpha is taken out of the row sum after the loop, its variate is masked below
        if (!ok1) dv.y = 0.0f;
        a_last = a.y;
        if (GRAD) {
            g12 = __ffma2_rn(psi, neg2(dv), g12);
            asum2 = __fadd2_rn(asum2, a);
            dsum2 = __fadd2_rn(dsum2, dv);
        }
        if (REC && alpha_row != nullptr) {
            alpha_row[2 * pp] = a.x;
            deriv_row[2 * pp] = dv.x;
            if (ok1) { alpha_row[2 * pp + 1] = a.y; deriv_row[2 * pp + 1] = dv.y; }
        }
    }
    __device__ __forceinline__ void phase_b(int pp, const float2 a, const float2 dv, const NoiseKey& nk, const PhiloxKeys& rk) {
        const bool ok1 = (2 * pp + 1) < D;
        float y0, y1;
        if (NOISE == DMFG_NOISE_PHILOX) {
            gamma_pair_fast(nk, rk, slot0 + (uint32_t)pp, a, scale, y0, y1);   // never returns 0
        } else {
            y0 = gamma_floor(row_ok ? noise_row[2 * pp] : 1.0f);                  // mfg_ac2.py:244 (+ denormals)
            y1 = gamma_floor((row_ok && ok1) ? noise_row[2 * pp + 1] : 1.0f);
        }
        if (GRAD) g22 = __ffma2_rn(make_float2(lg2_approx(y0), lg2_approx(y1)), dv, g22);
        const double yd0 = (double)y0, yd1 = (double)(ok1 ? y1 : 0.0f);
        ysum0 += yd0;
        ysum1 += yd1;
        if (has_reward) {
            const double2 pjd = lds_f64x2(a_pic + 16 * pp);
            racc0 = fma(yd0 * yd0, fma(c1, pjd.x, c0), racc0);
            racc1 = fma(yd1 * yd1, fma(c1, pjd.y, c0), racc1);
        }
        sts_f64(a_row + 16 * pp, yd0);
        sts_f64(a_row + 16 * pp + 8, yd1);
    }
    template <int N>
    __device__ __forceinline__ void chunk(int p0, const NoiseKey& nk, const PhiloxKeys& rk) {
        float2 a[N], dv[N];
#pragma unroll
        for (int c = 0; c < N; ++c) phase_a(p0 + c, a[c], dv[c]);
#pragma unroll
        for (int c = 0; c < N; ++c) phase_b(p0 + c, a[c], dv[c], nk, rk);
    }
};

template <int D, int NOISE, bool GRAD, bool REC>
__device__ __forceinline__ V2RowSums v2_row_walk(float theta, float xi, float scale, double c0, double c1, bool has_reward,
                                                 uint32_t a_pfc, uint32_t a_pic, uint32_t a_row, const NoiseKey& nk,
                                                 const PhiloxKeys& rk, uint32_t slot0, const float* __restrict__ noise_row,
                                                 bool row_ok, float* __restrict__ alpha_row, float* __restrict__ deriv_row) {
    constexpr int PD = (D + 1) / 2;
    V2Walk<D, NOISE, GRAD, REC> wk;
    wk.theta = theta; wk.xi = xi; wk.scale = scale; wk.c0 = c0; wk.c1 = c1; wk.has_reward = has_reward; wk.row_ok = row_ok;
    wk.a_pfc = a_pfc; wk.a_pic = a_pic; wk.a_row = a_row; wk.slot0 = slot0;
    wk.noise_row = noise_row; wk.alpha_row = alpha_row; wk.deriv_row = deriv_row;
    wk.asum2 = wk.dsum2 = wk.g12 = wk.g22 = make_float2(0.f, 0.f);
    wk.ysum0 = wk.ysum1 = wk.racc0 = wk.racc1 = 0.0;
    wk.a_last = 0.0f;
    constexpr int CH = DMFG_V2_CHUNK;
    if (CH <= 1) {
#pragma unroll kV2Unroll
        for (int pp = 0; pp < PD; ++pp) wk.template chunk<1>(pp, nk, rk);
    } else {
        constexpr int NFULL = PD / CH, REM = PD % CH;
#pragma unroll kV2Unroll
        for (int k = 0; k < NFULL; ++k) wk.template chunk<CH>(k * CH, nk, rk);
        if (REM > 0) wk.template chunk<(REM > 0 ? REM : 1)>(NFULL * CH, nk, rk);
    }
    V2RowSums o;
    o.ysum = wk.ysum0 + wk.ysum1;
    o.racc = wk.racc0 + wk.racc1;
    o.asum = wk.asum2.x + ((D & 1) ? wk.asum2.y - wk.a_last : wk.asum2.y);
    o.dsum = wk.dsum2.x + wk.dsum2.y;
    o.g1 = wk.g12.x + wk.g12.y;
    o.g2 = wk.g22.x + wk.g22.y;
    return o;
}

// TRAIN = the batched train step: accumulators wanted and NO per-step output stream requested -- every
// output test is compiled out of the step loop.  REC = some per-element stream (actions / alpha) is written.
// GRAD = d log F / d theta is wanted (critic attached or a grads stream): without it -- the IRL sampler and
// evaluate() only want states and actions -- psi(alpha), lg2 y and the alpha' sums are compiled out (-20 %).
// FUSE = dmfg_ac_step: the batch-mean update of (theta, w) is applied by the last CTA of the SAME launch.
template <int D, int NOISE, bool REC, bool TRAIN, bool GRAD, bool FUSE = false>
__global__ void __launch_bounds__(kV2Threads, DMFG_V2_MINB)
rollout_v2_kernel(const RolloutParams<float> p) {
    using S = V2Smem<D>;
    constexpr int G = S::G, NT = kV2Threads, GPB = S::GPB, kV2Slots = S::Slots, kV2StageRow = S::StageRow;
    constexpr int F = num_features_c(D);
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, r = tid & (G - 1), grp = tid / G, lane = tid & 31, warp = tid >> 5;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t a_row = sb + 8u * (S::tile + (grp * G + r) * kV2Slots);      // this lane's row of the tile
    const uint32_t a_col = sb + 8u * (S::tile + grp * G * kV2Slots + r);        // column r of the group's tile
    const uint32_t a_pid = sb + 8u * (S::pid + grp * G);                         // [2] buffers, stride GPB*G doubles
    const uint32_t a_pif = sb + 8u * S::pif + 4u * (grp * G);                    // [2] buffers, stride GPB*G floats
    const uint32_t a_q = sb + 8u * (S::qv + grp * G);
    const uint32_t a_wl_r = sb + 8u * (S::wl + r);
    const uint32_t a_stage = sb + 8u * (S::stage + warp * (2 * 4 * kV2StageRow));
    constexpr uint32_t kBufD = 8u * GPB * G, kBufF = 4u * GPB * G, kStageB = 8u * 4 * kV2StageRow;
    const bool td = TRAIN || p.w != nullptr;
    const bool want_acc = TRAIN || (td && p.partials != nullptr);
    const bool row_ok = r < D;
    if (td) {
        if (tid < G) stage_critic_slots<D>(smem + S::wl + tid, G, p.w, tid);
    }
    for (int k = tid; k < S::NW * 2 * 4 * kV2StageRow; k += NT) smem[S::stage + k] = 0.0;
    __syncthreads();
    const float theta = (float)(p.theta_dev ? *p.theta_dev : p.theta);
    const float shift = p.shift_f, scale = p.scale_f;
    const bool ac2 = p.reward_kind == DMFG_REWARD_AC2;
    const bool has_reward = p.reward_kind != DMFG_REWARD_NONE;
    const double rew_scale = ac2 ? 1.0 : -0.5;
    double sum_dg = 0.0, sum_r = 0.0;            // per-LANE partial sums (reduced once, at the end)
    double gacc[S::NBLK][2];                     // Gram blocks (I <= J) of the warp, C fragments of m8n8k4
#pragma unroll
    for (int k = 0; k < S::NBLK; ++k) gacc[k][0] = gacc[k][1] = 0.0;
    double lin_acc = 0.0, bias_acc = 0.0;
    const int sub = (tid / G) & (S::PW - 1), gid = lane >> 2, tig = lane & 3;   // sub: population of the warp
    const uint32_t a_frag = a_stage + 8u * (tig * kV2StageRow + gid);             // A[gid][tig] of the low tile
    const long long ntiles = (p.B + GPB - 1) / GPB;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        long long b = tile * GPB + grp;
        const bool live = b < p.B;                 // dead groups shadow the last population, writes masked
        if (!live) b = p.B - 1;
        const bool wr = live && row_ok;
        const double livef = live ? 1.0 : 0.0;
        const NoiseKey nk = make_noise_key(p.seed, (unsigned long long)(p.pop_offset + b));
        double pi_self = row_ok ? (double)p.pi0[b * D + r] : 0.0;
        uint32_t cur = 0;
        sts_f64(a_pid + 8 * r, pi_self);
        sts_f32(a_pif + 4 * r, (float)pi_self);
        warp_fence();
        double vc_lane = td ? critic_partial_v2<D, G>(a_wl_r, a_pid, pi_self) : 0.0;
        double v_cur = (td && !TRAIN) ? group_sum<G>(vc_lane) : 0.0;
        double disc = 1.0;
        if (!TRAIN && p.states != nullptr && wr) p.states[b * D + r] = (float)pi_self;
        for (int t = 0; t < p.T; ++t) {
            const long long tb = (long long)t * p.B + b;
            const long long row = (tb * D + r) * D;
            const uint32_t a_pic = a_pid + cur * kBufD, a_pfc = a_pif + cur * kBufF;
            // ------------------------------------------------------------------ the row
            const float xi = (float)pi_self + shift;
            // reward weights: AC2 sum_j y^2 (pi_j - pi_i);  synthetic sum_j y^2
            const double c1 = ac2 ? 1.0 : 0.0, c0 = ac2 ? -pi_self : 1.0;
            const uint32_t slot0 = gamma_slot((uint32_t)(step_base(p) + t), D, r, 0);
            const bool rec_alpha = REC && p.alpha != nullptr && wr;
            const V2RowSums rs = v2_row_walk<D, NOISE, GRAD, REC>(
                theta, xi, scale, c0, c1, has_reward, a_pfc, a_pic, a_row, nk, p.rk, slot0,
                NOISE == DMFG_NOISE_PHILOX ? nullptr : p.noise_y + row, row_ok,
                rec_alpha ? p.alpha + row : nullptr, rec_alpha ? p.alpha_deriv + row : nullptr);
            const float asum = rs.asum, dsum = rs.dsum, g1 = rs.g1, g2 = rs.g2;
            // ------------------------------------------------------------------ row level
            const double ysum = rs.ysum;
            const float ysum_f = (float)ysum;
            double inv = (double)rcp_approx(ysum_f);                   // 1/s: float seed + 2 Newton steps
            inv = inv * (2.0 - ysum * inv);
            inv = inv * (2.0 - ysum * inv);
            const double q = pi_self * inv;
            sts_f64(a_q + 8 * r, q);
            double glane = 0.0;
            if (GRAD) {
                const float psi_row = digamma_fast(asum);
                // sum_j alpha'_ij ln P_ij = ln2 (sum_j alpha'_ij lg2 y_ij - lg2 s_i sum_j alpha'_ij)
                const float lnp_term = DMFG_LN2 * fmaf(-lg2_approx(ysum_f), dsum, g2);
                glane = row_ok ? (double)(g1 + lnp_term + psi_row * dsum) : 0.0;
            }
            double rew_lane = has_reward ? rew_scale * (q * inv) * rs.racc : 0.0;
            if (REC && p.actions != nullptr) {
                const float inv_f = (float)inv;
                if (wr) {
                    float* act_row = p.actions + row;
#pragma unroll 5
                    for (int j = 0; j < D; ++j) act_row[j] = (float)lds_f64(a_row + 8 * j) * inv_f;
                }
            }
            warp_fence();
            // ------------------------------------------------------------------ pi' = sum_i q_i y_ij
            double n0 = 0.0, n1 = 0.0, n2 = 0.0;
#pragma unroll
            for (int i = 0; i + 1 < D; i += 2) {
                const double2 qq = lds_f64x2(a_q + 8 * i);
                const double ya = lds_f64(a_col + 8 * kV2Slots * i), yb = lds_f64(a_col + 8 * kV2Slots * (i + 1));
                if ((i / 2) % 3 == 0) { n0 = fma(qq.x, ya, n0); n1 = fma(qq.y, yb, n1); }
                else if ((i / 2) % 3 == 1) { n2 = fma(qq.x, ya, n2); n0 = fma(qq.y, yb, n0); }
                else { n1 = fma(qq.x, ya, n1); n2 = fma(qq.y, yb, n2); }
            }
            if (D & 1) n2 = fma(lds_f64(a_q + 8 * (D - 1)), lds_f64(a_col + 8 * kV2Slots * (D - 1)), n2);
            double next_self = (n0 + n1) + n2;
            if (!row_ok) next_self = 0.0;
            const uint32_t nxt = cur ^ 1u;
            sts_f64(a_pid + nxt * kBufD + 8 * r, next_self);
            sts_f32(a_pif + nxt * kBufF + 4 * r, (float)next_self);
            warp_fence();
            // ------------------------------------------------------------------ TD error, accumulators
            double rew = 0.0, grad = 0.0;
            if (!TRAIN) {
                rew = group_sum<G>(rew_lane);
                if (GRAD) grad = group_sum<G>(glane);
                if (p.rewards_in != nullptr) rew = (double)p.rewards_in[tb];
            }
            if (td) {
                const double vn_lane = critic_partial_v2<D, G>(a_wl_r, a_pid + nxt * kBufD, next_self);
                const double gfac = p.discount_kind == DMFG_DISCOUNT_STEP ? p.gamma : disc;
                double delta;
                if (TRAIN) {
                    delta = group_sum<G>(rew_lane + fma(gfac, vn_lane, -vc_lane));
                } else {
                    const double v_next = group_sum<G>(vn_lane);
                    delta = rew + gfac * v_next - v_cur;
                    v_cur = v_next;
                    if (p.deltas != nullptr && live && r == 0) p.deltas[tb] = (float)delta;
                }
                vc_lane = vn_lane;
                if (want_acc) {
                    const double dl = delta * livef;
                    const double dp = dl * pi_self;
                    sum_dg = fma(dl, glane, sum_dg);
                    lin_acc += dp;
                    bias_acc += dl;
                    // stage sample k = PW (t mod SPF) + sub:  A[k][r] = delta pi_r,  B[k][r] = pi_r
                    const int ts = t & (S::SPF - 1);
                    const uint32_t a_smp = a_stage + 8u * ((S::PW * ts + sub) * kV2StageRow + r);
                    sts_f64(a_smp, dp);
                    sts_f64(a_smp + kStageB, pi_self);
                    const bool last = t == p.T - 1;
                    if (ts == S::SPF - 1 || last) {
#pragma unroll
                        for (int e = 1; e < S::SPF; ++e)                           // T ends inside a flush group:
                            if (ts + e < S::SPF) sts_f64(a_smp + 8u * S::PW * kV2StageRow * e, 0.0);   // the rest is empty
                        warp_fence();
                        double af[S::NB], bf[S::NB];
#pragma unroll
                        for (int m = 0; m < S::NB; ++m) {
                            af[m] = lds_f64(a_frag + 64 * m);
                            bf[m] = lds_f64(a_frag + kStageB + 64 * m);
                        }
                        // blocks on or above the diagonal (rows 8I.. x cols 8J.., I <= J); below it is not a feature
#pragma unroll
                        for (int I = 0; I < S::NB; ++I)
#pragma unroll
                            for (int J = I; J < S::NB; ++J) {
                                const int k = I * S::NB - (I * (I - 1)) / 2 + (J - I);
                                dmma884(gacc[k][0], gacc[k][1], af[I], bf[J]);
                            }
                        warp_fence();
                    }
                }
            }
            sum_r = fma(livef, TRAIN ? rew_lane : (r == 0 ? rew : 0.0), sum_r);
            if (!TRAIN && live && r == 0) {
                if (p.rewards != nullptr) p.rewards[tb] = (float)rew;
                if (p.grads != nullptr) p.grads[tb] = (float)grad;
            }
            disc *= p.gamma;
            pi_self = next_self;
            cur = nxt;
            if (!TRAIN && p.states != nullptr && wr) p.states[(tb + p.B) * D + r] = (float)pi_self;
        }
        if (!TRAIN && p.pi_final != nullptr && wr) p.pi_final[b * D + r] = (float)pi_self;
        // the next population starts from buffer 0 again: move on only when every lane is done reading
        warp_fence();
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
        if (lane == 0) { red[0][warp] = a; red[1][warp] = c; }
    }
    __syncthreads();                                   // everyone is done with the tile: reuse it
    {
        double* gw = smem + S::gram + warp * (G * kV2Slots);
#pragma unroll
        for (int I = 0; I < S::NB; ++I)
#pragma unroll
            for (int J = I; J < S::NB; ++J) {
                const int k = I * S::NB - (I * (I - 1)) / 2 + (J - I);
                gw[(8 * I + gid) * kV2Slots + 8 * J + 2 * tig] = gacc[k][0];
                gw[(8 * I + gid) * kV2Slots + 8 * J + 2 * tig + 1] = gacc[k][1];
            }
        smem[S::lin + grp * G + r] = lin_acc;
        if (r == 0) smem[S::bias + grp] = bias_acc;
    }
    __syncthreads();
    double* out = p.partials + (long long)blockIdx.x * (2 + F);
    for (int f = tid; f < F; f += NT) {
        constexpr int Q = D * (D + 1) / 2;
        double s = 0.0;
        if (f < Q) {
            // feature f of the reference order (itertools.combinations_with_replacement) -> (i <= j)
            int i = 0, rem = f;
            while (rem >= D - i) { rem -= D - i; ++i; }
            const int j = i + rem;
            for (int wv = 0; wv < S::NW; ++wv) s += smem[S::gram + wv * (G * kV2Slots) + i * kV2Slots + j];
        } else if (f < Q + D) {
            for (int g = 0; g < GPB; ++g) s += smem[S::lin + g * G + (f - Q)];
        } else {
            for (int g = 0; g < GPB; ++g) s += smem[S::bias + g];
        }
        out[1 + f] = s;
    }
    if (tid == 0) {
        double a = 0.0, c = 0.0;
        for (int wv = 0; wv < NT / 32; ++wv) { a += red[0][wv]; c += red[1][wv]; }
        out[0] = a;
        out[1 + F] = c;
    }
    if (FUSE) {
        // ---- one launch per step: whoever finishes last sums the per-CTA partials (CTA order: deterministic) and
        // updates the parameters in place; every other CTA has long finished reading theta and w
        __shared__ unsigned int ticket;
        __threadfence();
        __syncthreads();
        if (tid == 0) ticket = atomicAdd(p.fuse_counter, 1u);
        __syncthreads();
        if (ticket == gridDim.x - 1) {
            __threadfence();
            const double lr_c = p.fuse_lr_dev ? p.fuse_lr_dev[0] : p.fuse_lr_c;
            const double lr_a = p.fuse_lr_dev ? p.fuse_lr_dev[1] : p.fuse_lr_a;
            for (int f = tid; f < F + 2; f += NT) {
                double s = 0.0;
                for (unsigned int c = 0; c < gridDim.x; ++c) s += __ldcg(p.partials + (long long)c * (2 + F) + f);
                if (p.fuse_acc != nullptr) p.fuse_acc[f] = s;
                // (the arithmetic of ac_apply_update_kernel, so the fused step equals the three-launch chain bit for bit)
                if (f == 0) p.fuse_theta[0] = fma(lr_a * p.fuse_scale, s, p.fuse_theta[0]);
                else if (f <= F) p.fuse_w[f - 1] = fma(lr_c * p.fuse_scale, s, p.fuse_w[f - 1]);
            }
            if (tid == 0) *p.fuse_counter = 0u;                          // ready for the next step's launch
        }
    }
}

// ---------------------------------------------------------------------------
// Independent serial learners on the v2 math (float streams; d = 15 / 16 with 16-lane groups, d = 21 -- the
// reference's default, mfg_ac2.py:25 -- with 32-lane groups): one group = one learner with
// private (theta, w) and per-step online updates -- mfg_ac2.py:448-539 semantics exactly, as in
// learners_fast_kernel, but with the single-pass packed row walk of rollout_v2_kernel (v2_row_walk).
// Shared memory per CTA: the y tile, double-buffered state, q, and the learners' critic slots [GPB][D+2][16].
// ---------------------------------------------------------------------------
template <int D, int G>
struct LearnersV2Smem {
    static constexpr int GPB = kV2Threads / G;
    static constexpr int NSLOT = D + 2;
    static constexpr int SL = G + 1;                                 // doubles per tile row (odd => conflict free)
    static constexpr int tile = 0;
    static constexpr int pid = tile + GPB * G * SL;
    static constexpr int qv = pid + 2 * GPB * G;
    static constexpr int pif = qv + GPB * G;                         // [2][GPB][G] floats
    static constexpr int wl = pif + GPB * G;                         // [GPB][NSLOT][G] doubles
    // block of one learner; for G = 16 it is padded so that the two learners of a warp sit 16 (mod 32) words apart: with
    // NSLOT * G doubles = 0 (mod 32) words both groups of a warp updated their slots through the SAME banks (ncu: 2
    // wavefronts per STS.64 instead of 1, 9 % of the kernel's shared-memory traffic)
    static constexpr int WLB = NSLOT * G + ((G == 16 && (NSLOT * G * 2) % 32 == 0) ? 8 : 0);
    static constexpr int total = wl + GPB * WLB;
};

template <int D, int G, int NOISE>
__global__ void __launch_bounds__(kV2Threads, 2)
learners_v2_kernel(const LearnerParams<float> p, const PhiloxKeys rk) {
    using S = LearnersV2Smem<D, G>;
    constexpr int GPB = S::GPB, NSLOT = S::NSLOT, SL = S::SL;
    static_assert(D <= G && (G == 16 || G == 32), "one lane per row of P");
    constexpr int F = num_features_c(D);
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, r = tid & (G - 1), grp = tid / G;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t a_row = sb + 8u * (S::tile + (grp * G + r) * SL);
    const uint32_t a_col = sb + 8u * (S::tile + grp * G * SL + r);
    const uint32_t a_pid = sb + 8u * (S::pid + grp * G);
    const uint32_t a_pif = sb + 8u * S::pif + 4u * (grp * G);
    const uint32_t a_q = sb + 8u * (S::qv + grp * G);
    const uint32_t a_wl_r = sb + 8u * (S::wl + grp * S::WLB + r);
    constexpr uint32_t kBufD = 8u * GPB * G, kBufF = 4u * GPB * G;
    const bool row_ok = r < D;
    long long l = (long long)blockIdx.x * GPB + grp;
    const bool live = l < p.L;
    if (!live) l = p.L - 1;
    stage_critic_slots<D>(smem + S::wl + grp * S::WLB + r, G, p.w + l * F, r);      // lane r only ever reads its own column
    double theta = p.theta[l];
    const float shift = (float)(p.shift ? p.shift[l] : p.shift_scalar);
    const float scale = (float)(p.alpha_scale ? p.alpha_scale[l] : p.alpha_scale_scalar);
    const bool ac2 = p.reward_kind == DMFG_REWARD_AC2;
    const bool has_reward = p.reward_kind != DMFG_REWARD_NONE;
    const double rew_scale = ac2 ? 1.0 : -0.5;
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
        uint32_t cur = 0;
        warp_fence();                                            // everyone is done with the previous episode's buffers
        sts_f64(a_pid + 8 * r, pi_self);
        sts_f32(a_pif + 4 * r, (float)pi_self);
        warp_fence();
        const double lr_c = p.constant_lr ? p.lr_critic : p.lr_critic / (episode + 1.0);
        const double lr_a = p.constant_lr ? p.lr_actor : p.lr_actor / ((episode + 1.0) * log(log(episode + 20.0)));
        double disc = 1.0, total_lane = 0.0;
        for (int t = 0; t < p.T; ++t) {
            const long long et = ((long long)l * p.E + e) * p.T + t;
            const long long row = (et * D + r) * D;
            const uint32_t a_pic = a_pid + cur * kBufD, a_pfc = a_pif + cur * kBufF;
            const float thf = (float)theta;
            // ------------------------------------------------------------------ the row (see rollout_v2_kernel)
            const float xi = (float)pi_self + shift;
            const double c1 = ac2 ? 1.0 : 0.0, c0 = ac2 ? -pi_self : 1.0;
            const uint32_t slot0 = gamma_slot((uint32_t)((episode + p.noise_episode_offset) * p.T + t), D, r, 0);
            const V2RowSums rs = v2_row_walk<D, NOISE, true, false>(
                thf, xi, scale, c0, c1, has_reward, a_pfc, a_pic, a_row, nk, rk, slot0,
                NOISE == DMFG_NOISE_PHILOX ? nullptr : p.noise_y + row, row_ok, nullptr, nullptr);
            const float asum = rs.asum, dsum = rs.dsum, g1 = rs.g1, g2 = rs.g2;
            // ------------------------------------------------------------------ row level
            const double ysum = rs.ysum;
            const float ysum_f = (float)ysum;
            double inv = (double)rcp_approx(ysum_f);
            inv = inv * (2.0 - ysum * inv);
            inv = inv * (2.0 - ysum * inv);
            const double q = pi_self * inv;
            sts_f64(a_q + 8 * r, q);
            const float psi_row = digamma_fast(asum);
            const float lnp_term = DMFG_LN2 * fmaf(-lg2_approx(ysum_f), dsum, g2);
            const double glane = row_ok ? (double)(g1 + lnp_term + psi_row * dsum) : 0.0;
            const double rew_lane = has_reward ? rew_scale * (q * inv) * rs.racc : 0.0;
            warp_fence();
            // ------------------------------------------------------------------ pi' = sum_i q_i y_ij
            double n0 = 0.0, n1 = 0.0, n2 = 0.0;
#pragma unroll
            for (int i = 0; i + 1 < D; i += 2) {
                const double2 qq = lds_f64x2(a_q + 8 * i);
                const double ya = lds_f64(a_col + 8 * SL * i), yb = lds_f64(a_col + 8 * SL * (i + 1));
                if ((i / 2) % 3 == 0) { n0 = fma(qq.x, ya, n0); n1 = fma(qq.y, yb, n1); }
                else if ((i / 2) % 3 == 1) { n2 = fma(qq.x, ya, n2); n0 = fma(qq.y, yb, n0); }
                else { n1 = fma(qq.x, ya, n1); n2 = fma(qq.y, yb, n2); }
            }
            if (D & 1) n2 = fma(lds_f64(a_q + 8 * (D - 1)), lds_f64(a_col + 8 * SL * (D - 1)), n2);
            double next_self = (n0 + n1) + n2;
            if (!row_ok) next_self = 0.0;
            const uint32_t nxt = cur ^ 1u;
            sts_f64(a_pid + nxt * kBufD + 8 * r, next_self);
            sts_f32(a_pif + nxt * kBufF + 4 * r, (float)next_self);
            warp_fence();
            // ------------------------------------------------------------------ TD error with the CURRENT w, updates
            const double vn_lane = critic_partial_v2<D, G>(a_wl_r, a_pid + nxt * kBufD, next_self);
            const double vc_lane = critic_partial_v2<D, G>(a_wl_r, a_pic, pi_self);
            const double gfac = p.discount_kind == DMFG_DISCOUNT_STEP ? p.gamma : disc;
            const double delta = group_sum<G>(rew_lane + fma(gfac, vn_lane, -vc_lane));
            const double grad = group_sum<G>(glane);
            // critic first, then actor, both with the same delta (mfg_ac2.py:505-522)
            const double step_w = lr_c * delta;
            const double dp = step_w * pi_self;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                if (k >= r) {
                    const uint32_t aw = a_wl_r + 8u * G * k;
                    sts_f64(aw, fma(dp, lds_f64(a_pic + 8 * k), lds_f64(aw)));
                }
            }
            if (row_ok) sts_f64(a_wl_r + 8u * G * D, lds_f64(a_wl_r + 8u * G * D) + dp);
            if (r == 0) sts_f64(a_wl_r + 8u * G * (D + 1), lds_f64(a_wl_r + 8u * G * (D + 1)) + step_w);
            theta = fma(lr_a * delta, grad, theta);
            if (live && r == 0) {
                if (p.theta_trace) p.theta_trace[et] = theta;
                if (p.delta_trace) p.delta_trace[et] = delta;
            }
            total_lane += rew_lane;
            disc *= p.gamma;
            pi_self = next_self;
            cur = nxt;
        }
        const double total = group_sum<G>(total_lane);
        if (live && r == 0 && p.total_reward) p.total_reward[l * p.E + e] = total;
    }
    if (!live) return;
    if (r == 0) p.theta[l] = theta;
    if (p.pi_final && row_ok) p.pi_final[l * D + r] = (float)pi_self;
    // write the private critic weights back in the reference's feature order
    constexpr int Q = D * (D + 1) / 2;
    double* wout = p.w + l * F;
    const double* wl = smem + S::wl + grp * S::WLB + r;
    if (row_ok) {
#pragma unroll
        for (int k = 0; k < D; ++k)
            if (k >= r) wout[quad_index(D, r, k)] = wl[k * G];
        wout[Q + r] = wl[D * G];
    }
    if (r == 0) wout[Q + D] = wl[(D + 1) * G];
}

// ---------------------------------------------------------------------------
// WIDE rollout kernel (float streams, any d <= DMFG_MAX_D): warp per population, rows in sequence, lane q*32+l
// owns column pair l + 32 q of every row (NPL pairs per lane).  The v2 single-pass math per pair; per row only
// three warp reductions (sum y in double, sum alpha, sum alpha'), everything else is kept as lane partials for the
// whole step: the reward (pi_i / s_i^2 is row-uniform), the two linear gradient sums, and pi'_j = sum_i q_i y_ij
// in registers (no shared-memory accumulation).  Actions are stored as coalesced runs of a row.
// No critic here: TD errors for these d come from td_delta_kernel / td_gw_kernel on the record.
// ---------------------------------------------------------------------------
constexpr int kWideThreads = 128;
constexpr int kWideBatch = 8;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NPL, int NOISE, bool GRAD>
__global__ void __launch_bounds__(kWideThreads)
rollout_wide_kernel(const RolloutParams<float> p) {
    extern __shared__ __align__(16) double wsm[];
    const int d = p.d, pd = (d + 1) >> 1, dpad = 2 * pd;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int WPB = kWideThreads / 32;
    double* pid = wsm + (size_t)wib * (((dpad + pd + 1) & ~1) + kWideBatch * 32);   // state, double [dpad] (16-byte aligned per warp)
    float* pif = reinterpret_cast<float*>(pid + dpad);             // state, float [dpad]
    // lane partials of (sum alpha, sum alpha') of the last kWideBatch rows: their warp totals are only needed for
    // the row term psi(sum alpha) sum alpha', so they are reduced kWideBatch rows at a time (8 LDS + 2 SHFL per
    // total instead of 10 SHFL per row) and digamma runs once per batch on 8 lane groups in parallel
    float* asb = reinterpret_cast<float*>(pid + ((dpad + pd + 1) & ~1));   // [kWideBatch][32]
    float* dsb = asb + kWideBatch * 32;                                    // [kWideBatch][32]
    const float theta = (float)(p.theta_dev ? *p.theta_dev : p.theta);
    const float shift = p.shift_f, scale = p.scale_f;
    const bool ac2 = p.reward_kind == DMFG_REWARD_AC2;
    const bool has_reward = p.reward_kind != DMFG_REWARD_NONE;
    const double rew_scale = ac2 ? 1.0 : -0.5;
    for (long long b = (long long)blockIdx.x * WPB + wib; b < p.B; b += (long long)gridDim.x * WPB) {
        const NoiseKey nk = make_noise_key(p.seed, (unsigned long long)(p.pop_offset + b));
        for (int j = lane; j < dpad; j += 32) {
            const float v = j < d ? p.pi0[b * d + j] : 0.0f;
            pid[j] = (double)v;
            pif[j] = v;
            if (p.states != nullptr && j < d) p.states[b * d + j] = v;
        }
        __syncwarp();
        for (int t = 0; t < p.T; ++t) {
            const long long tb = (long long)t * p.B + b;
            double nx0[NPL], nx1[NPL];
#pragma unroll
            for (int q = 0; q < NPL; ++q) { nx0[q] = 0.0; nx1[q] = 0.0; }
            double rew_acc = 0.0, grow = 0.0;          // grow: lane partial of the psi(sum alpha) sum alpha' terms
            float glin = 0.0f;                         // lane partial of -ln s_i sum_j alpha'_ij
            float2 g12 = make_float2(0.f, 0.f), g22 = g12;
            for (int i = 0; i < d; ++i) {
                const double pi_i = pid[i];
                const long long row = (tb * d + i) * d;
                const float xi = (float)pi_i + shift;
                const double c1 = ac2 ? 1.0 : 0.0, c0 = ac2 ? -pi_i : 1.0;
                float2 yv[NPL];
                float2 as2 = make_float2(0.f, 0.f), ds2 = as2;
                double ysum = 0.0, racc = 0.0;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int pp = lane + 32 * q;
                    yv[q] = make_float2(0.f, 0.f);
                    if (pp < pd) {
                        const bool ok1 = (2 * pp + 1) < d;
                        const float2 pj = *reinterpret_cast<const float2*>(pif + 2 * pp);
                        float2 a, dv, psi;
                        if (GRAD) {
                            alpha_psi_fast2(theta, __fadd2_rn(pj, splat2(-xi)), a, dv, psi);
                        } else {
                            policy_alpha_fast2(theta, __fadd2_rn(pj, splat2(-xi)), a, dv);
                            psi = make_float2(0.f, 0.f);
                        }
                        if (!ok1) dv.y = 0.0f;
                        if (GRAD) {
                            g12 = __ffma2_rn(psi, neg2(dv), g12);
                            as2 = __fadd2_rn(as2, make_float2(a.x, ok1 ? a.y : 0.0f));
                            ds2 = __fadd2_rn(ds2, dv);
                        }
                        float y0, y1;
                        if (NOISE == DMFG_NOISE_PHILOX) {
                            gamma_pair_fast(nk, p.rk, gamma_slot((uint32_t)(step_base(p) + t), d, i, pp), a, scale, y0, y1);
                        } else {
                            y0 = gamma_floor(p.noise_y[row + 2 * pp]);                       // mfg_ac2.py:244 (+ denormals)
                            y1 = gamma_floor(ok1 ? p.noise_y[row + 2 * pp + 1] : 1.0f);
                        }
                        if (GRAD) g22 = __ffma2_rn(make_float2(lg2_approx(y0), lg2_approx(y1)), dv, g22);
                        if (!ok1) y1 = 0.0f;
                        yv[q] = make_float2(y0, y1);
                        const double yd0 = (double)y0, yd1 = (double)y1;
                        ysum += yd0 + yd1;
                        if (has_reward) {
                            const double2 pjd = *reinterpret_cast<const double2*>(pid + 2 * pp);
                            racc = fma(yd0 * yd0, fma(c1, pjd.x, c0), racc);
                            racc = fma(yd1 * yd1, fma(c1, pjd.y, c0), racc);
                        }
                        if (p.alpha != nullptr) {
                            p.alpha[row + 2 * pp] = a.x;
                            p.alpha_deriv[row + 2 * pp] = dv.x;
                            if (ok1) { p.alpha[row + 2 * pp + 1] = a.y; p.alpha_deriv[row + 2 * pp + 1] = dv.y; }
                        }
                    }
                }
                ysum = group_sum<32>(ysum);
                const float ysum_f = (float)ysum;
                double inv = (double)rcp_approx(ysum_f);               // 1/s: float seed + 2 Newton steps
                inv = inv * (2.0 - ysum * inv);
                inv = inv * (2.0 - ysum * inv);
                const double qi = pi_i * inv;
                const float inv_f = (float)inv;
                if (has_reward) rew_acc = fma(rew_scale * qi * inv, racc, rew_acc);
                if (GRAD) {
                    // the two row-level terms of d log F / d theta: psi(sum_j alpha) sum_j alpha' - ln s sum_j alpha'.
                    // The second is linear in the lane partials of sum alpha' (ln s is row-uniform): no reduction.
                    const float dpart = ds2.x + ds2.y;
                    glin = fmaf(-DMFG_LN2 * lg2_approx(ysum_f), dpart, glin);
                    const int slot = i & (kWideBatch - 1);
                    asb[slot * 32 + lane] = as2.x + as2.y;
                    dsb[slot * 32 + lane] = dpart;
                    if (slot == kWideBatch - 1 || i == d - 1) {
                        __syncwarp();
                        // lane group g = lane / 4 owns row g of the batch; its 4 lanes sum 8 partials each
                        const int g = lane >> 2, sub = (lane & 3) * 8;
                        float av = 0.f, dvv = 0.f;
                        if (g <= slot) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) { av += asb[g * 32 + sub + k]; dvv += dsb[g * 32 + sub + k]; }
                        }
                        av += __shfl_xor_sync(0xffffffffu, av, 1); av += __shfl_xor_sync(0xffffffffu, av, 2);
                        dvv += __shfl_xor_sync(0xffffffffu, dvv, 1); dvv += __shfl_xor_sync(0xffffffffu, dvv, 2);
                        if (g <= slot && (lane & 3) == 0) grow += (double)(digamma_fast(av) * dvv);
                        __syncwarp();
                    }
                }
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int pp = lane + 32 * q;
                    if (pp < pd) {
                        nx0[q] = fma(qi, (double)yv[q].x, nx0[q]);
                        nx1[q] = fma(qi, (double)yv[q].y, nx1[q]);
                        if (p.actions != nullptr) {
                            p.actions[row + 2 * pp] = yv[q].x * inv_f;
                            if ((2 * pp + 1) < d) p.actions[row + 2 * pp + 1] = yv[q].y * inv_f;
                        }
                    }
                }
            }
            __syncwarp();                                              // every lane is done reading this step's state
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                const int pp = lane + 32 * q;
                if (pp < pd) {
                    const bool ok1 = (2 * pp + 1) < d;
                    const double v0 = nx0[q], v1 = ok1 ? nx1[q] : 0.0;
                    *reinterpret_cast<double2*>(pid + 2 * pp) = make_double2(v0, v1);
                    *reinterpret_cast<float2*>(pif + 2 * pp) = make_float2((float)v0, (float)v1);
                    if (p.states != nullptr) {
                        p.states[(tb + p.B) * d + 2 * pp] = (float)v0;
                        if (ok1) p.states[(tb + p.B) * d + 2 * pp + 1] = (float)v1;
                    }
                }
            }
            if (p.rewards != nullptr || p.grads != nullptr) {
                const double rew = group_sum<32>(rew_acc);
                double grad = 0.0;
                if (GRAD) grad = group_sum<32>((double)((g12.x + g12.y) + DMFG_LN2 * (g22.x + g22.y) + glin) + grow);
                if (lane == 0) {
                    if (p.rewards != nullptr) p.rewards[tb] = (float)rew;
                    if (p.grads != nullptr) p.grads[tb] = (float)grad;
                }
            }
            __syncwarp();
        }
        if (p.pi_final != nullptr)
            for (int j = lane; j < d; j += 32) p.pi_final[b * d + j] = (float)pid[j];
        __syncwarp();
    }
}

}  // namespace dmfg
