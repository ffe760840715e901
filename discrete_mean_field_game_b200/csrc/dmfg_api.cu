// extern "C" entry points of libdmfg (see include/dmfg.h): argument checking,
// kernel selection, launch geometry, workspace carving.  No torch, no hidden
// state: every call works on caller-owned device pointers and a caller stream.
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include <cstring>
#include <string>
#include <type_traits>

#include "dmfg_error.h"
#include "dmfg_rollout2.cuh"
#include "dmfg_td_dmma.cuh"
#include "dmfg_td_small.cuh"
#include "dmfg_learner_cta.cuh"

using namespace dmfg;

namespace {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};
}

namespace dmfg {
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int sm_count(int* out) {
    int dev = 0;
    DMFG_CUDA(cudaGetDevice(&dev));
    DMFG_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
    return DMFG_OK;
}
}  // namespace dmfg

namespace {

constexpr int kMaxPartialCtas = 148 * 8;     // upper bound on the persistent grids below
constexpr int kTdChunk = 64;                 // transitions staged per CTA iteration in td_gw_kernel

inline uint64_t align_up(uint64_t x, uint64_t a = 256) { return (x + a - 1) / a * a; }
inline size_t esize(int dtype) { return dtype == DMFG_F64 ? 8 : 4; }
inline bool fast_d(int d) { return d == 4 || d == 15 || d == 16; }

// TD pass on the FP64 tensor-core path (dmfg_td_dmma.cuh): float streams, d a multiple of 16, d >= 32
struct TdDmmaPlan {
    bool use = false;
    int nblk = 0, ksplit = 0;
    long long kchunk = 0;
    uint64_t off_U = 0, off_v = 0, off_part = 0, total = 0;
};
inline TdDmmaPlan td_dmma_plan(int dtype, int d, int T, long long B) {
    TdDmmaPlan pl;
    if (dtype != DMFG_F32 || d < 32 || (d % 16) != 0 || T < 1 || B < 1) return pl;
    pl.use = true;
    const int nb = d / 16;
    pl.nblk = nb * (nb + 1) / 2;
    const long long Nt = (long long)T * B;
    long long ks = (148LL * 32 + pl.nblk - 1) / pl.nblk;              // ~32 warps per SM over all blocks
    if (ks > (Nt + 63) / 64) ks = (Nt + 63) / 64;                        // at least 64 samples per warp
    if (ks < 1) ks = 1;
    pl.kchunk = ((Nt + ks - 1) / ks + 3) / 4 * 4;
    pl.ksplit = (int)((Nt + pl.kchunk - 1) / pl.kchunk);
    uint64_t off = 0;
    pl.off_U = off; off += align_up((uint64_t)d * d * 8);
    pl.off_v = off; off += align_up((uint64_t)(T + 1) * (uint64_t)B * 8);
    pl.off_part = off; off += align_up((uint64_t)pl.ksplit * pl.nblk * 256 * 8);
    pl.total = off;
    return pl;
}


// the throughput kernel: float streams, d = 15 / 16 (16-lane groups) and d = 20 / 21 / 32 (32-lane groups; 21 is the
// default of mfg_ac2.py:25, 20 the size of test2.py:9), sampled or injected Gamma variates
bool v2_d(int d) { return d == 15 || d == 16 || d == 20 || d == 21 || d == 32; }
bool use_v2(const dmfg_rollout_args* a) {
    if (a->variant != DMFG_VARIANT_AUTO && a->variant != DMFG_VARIANT_V2) return false;
    return a->dtype == DMFG_F32 && v2_d(a->d) && a->noise_kind != DMFG_NOISE_ACTIONS;
}
// a fused kernel (TD in the rollout) exists for this call
bool use_fast(const dmfg_rollout_args* a) {
    if (a->variant == DMFG_VARIANT_GENERIC) return false;
    return fast_d(a->d) || use_v2(a);
}

// workspace map of one dmfg_rollout call
struct RolloutWs {
    uint64_t partials = 0, states = 0, rewards = 0, grads = 0, delta_buf = 0, dmma = 0, total = 0;
    bool need_states = false, need_rewards = false, need_grads = false;
};
RolloutWs rollout_ws(const dmfg_rollout_args* a) {
    RolloutWs w;
    const uint64_t F = (uint64_t)num_features_c(a->d);
    const bool td = a->w != nullptr;
    const bool accum = td && a->acc != nullptr;
    uint64_t off = 0;
    if (accum) { w.partials = off; off += align_up((uint64_t)kMaxPartialCtas * (2 + F) * 8); }
    if (!use_fast(a) && td) {
        const uint64_t TB = (uint64_t)a->T * (uint64_t)a->B, es = esize(a->dtype);
        if (!a->states) { w.need_states = true; w.states = off; off += align_up((TB + a->B) * a->d * es); }
        if (!a->rewards && !a->rewards_in) { w.need_rewards = true; w.rewards = off; off += align_up(TB * es); }
        if (!a->grads) { w.need_grads = true; w.grads = off; off += align_up(TB * es); }
        w.delta_buf = off; off += align_up(TB * 8);
        w.dmma = off; off += td_dmma_plan(a->dtype, a->d, a->T, a->B).total;
    }
    w.total = off;
    return w;
}

int check_rollout(const dmfg_rollout_args* a) {
    if (!a) return fail(DMFG_ERR_INVALID, "args is NULL");
    if (a->struct_size != sizeof(dmfg_rollout_args))
        return fail(DMFG_ERR_INVALID, "dmfg_rollout_args.struct_size %u != %zu (header mismatch)",
                    a->struct_size, sizeof(dmfg_rollout_args));
    if (a->dtype != DMFG_F32 && a->dtype != DMFG_F64) return fail(DMFG_ERR_INVALID, "dtype %d", a->dtype);
    if (a->d < 1 || a->d > DMFG_MAX_D) return fail(DMFG_ERR_INVALID, "d=%d outside [1,%d]", a->d, DMFG_MAX_D);
    if (a->T < 0) return fail(DMFG_ERR_INVALID, "T=%d < 0", a->T);
    if (a->B < 0) return fail(DMFG_ERR_INVALID, "B=%lld < 0", (long long)a->B);
    if (a->reward_kind < DMFG_REWARD_NONE || a->reward_kind > DMFG_REWARD_SYNTHETIC)
        return fail(DMFG_ERR_INVALID, "reward_kind %d", a->reward_kind);
    if (a->discount_kind != DMFG_DISCOUNT_STEP && a->discount_kind != DMFG_DISCOUNT_CUMULATIVE)
        return fail(DMFG_ERR_INVALID, "discount_kind %d", a->discount_kind);
    if (a->noise_kind < DMFG_NOISE_INJECTED || a->noise_kind > DMFG_NOISE_ACTIONS)
        return fail(DMFG_ERR_INVALID, "noise_kind %d", a->noise_kind);
    if (a->noise_kind == DMFG_NOISE_ACTIONS && a->T > 1)
        return fail(DMFG_ERR_INVALID, "noise_kind=ACTIONS evaluates single transitions (T must be <= 1)");
    if (a->variant < DMFG_VARIANT_AUTO || a->variant > DMFG_VARIANT_V2)
        return fail(DMFG_ERR_INVALID, "variant %d", a->variant);
    if (a->variant == DMFG_VARIANT_V2 && !use_v2(a))
        return fail(DMFG_ERR_UNSUPPORTED, "the v2 kernel is built for float streams, d in {15,16,20,21,32}, sampled or injected noise");
    if (a->variant == DMFG_VARIANT_FAST && !fast_d(a->d))
        return fail(DMFG_ERR_UNSUPPORTED, "fast variant is built for d in {4,15,16}, not d=%d", a->d);
    if (a->B > 0 && !a->pi0) return fail(DMFG_ERR_INVALID, "pi0 is NULL");
    if (a->B > 0 && a->T > 0 && a->noise_kind != DMFG_NOISE_PHILOX && !a->noise_y)
        return fail(DMFG_ERR_INVALID, "noise_kind=INJECTED/ACTIONS needs noise_y");
    if ((a->deltas || a->acc) && !a->w) return fail(DMFG_ERR_INVALID, "deltas/acc need critic weights w");
    if ((a->alpha == nullptr) != (a->alpha_deriv == nullptr))
        return fail(DMFG_ERR_INVALID, "alpha and alpha_deriv must be requested together");
    if (!(a->alpha_scale > 0.0)) return fail(DMFG_ERR_INVALID, "alpha_scale must be > 0");
    return DMFG_OK;
}

template <typename R>
RolloutParams<R> make_params(const dmfg_rollout_args* a) {
    RolloutParams<R> p;
    p.d = a->d; p.T = a->T; p.B = a->B; p.pop_offset = a->pop_offset;
    p.theta = a->theta; p.theta_dev = a->theta_dev;
    p.shift = a->shift; p.alpha_scale = a->alpha_scale; p.gamma = a->gamma;
    p.reward_kind = a->reward_kind; p.discount_kind = a->discount_kind;
    p.noise_y = (const R*)a->noise_y; p.seed = a->seed; p.step_offset = a->step_offset;
    p.step_offset_dev = (const unsigned long long*)a->step_offset_dev;
    p.pi0 = (const R*)a->pi0; p.w = a->w; p.rewards_in = (const R*)a->rewards_in;
    p.states = (R*)a->states; p.actions = (R*)a->actions; p.alpha = (R*)a->alpha;
    p.alpha_deriv = (R*)a->alpha_deriv; p.rewards = (R*)a->rewards; p.deltas = (R*)a->deltas;
    p.grads = (R*)a->grads; p.pi_final = (R*)a->pi_final; p.partials = nullptr;
    p.rk = make_philox_keys(a->seed);
    p.shift_f = (float)a->shift; p.scale_f = (float)a->alpha_scale;
    p.fuse_counter = nullptr; p.fuse_theta = nullptr; p.fuse_w = nullptr; p.fuse_acc = nullptr; p.fuse_lr_dev = nullptr;
    p.fuse_lr_c = p.fuse_lr_a = p.fuse_scale = 0.0;
    return p;
}

template <int D, int G, typename R, int NOISE>
int launch_fast(RolloutParams<R> p, bool td, int* grid_out, cudaStream_t st) {
    auto kern = rollout_fast_kernel<D, G, R, NOISE>;
    const size_t smem = td ? (size_t)2 * (D + 2) * kFastThreads * sizeof(double) : 0;
    DMFG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0, sms = 0;
    DMFG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kFastThreads, smem));
    if (int rc = sm_count(&sms)) return rc;
    if (occ < 1) return fail(DMFG_ERR_CUDA, "rollout_fast_kernel<%d> does not fit an SM", D);
    const long long gpb = kFastThreads / G;
    const long long ntiles = (p.B + gpb - 1) / gpb;
    long long grid = (long long)sms * occ;
    if (grid > kMaxPartialCtas) grid = kMaxPartialCtas;
    if (grid > ntiles) grid = ntiles;
    *grid_out = (int)grid;
    kern<<<(unsigned)grid, kFastThreads, smem, st>>>(p);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

template <int D, int NOISE, bool REC, bool TRAIN, bool GRAD, bool FUSE = false>
int launch_v2(RolloutParams<float> p, bool td, int* grid_out, cudaStream_t st) {
    auto kern = rollout_v2_kernel<D, NOISE, REC, TRAIN, GRAD, FUSE>;
    const size_t smem = (size_t)V2Smem<D>::total * sizeof(double);
    DMFG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0, sms = 0;
    DMFG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kV2Threads, smem));
    if (int rc = sm_count(&sms)) return rc;
    if (occ < 1) return fail(DMFG_ERR_CUDA, "rollout_v2_kernel<%d> does not fit an SM", D);
    const long long gpb = V2Smem<D>::GPB;
    const long long ntiles = (p.B + gpb - 1) / gpb;
    long long grid = (long long)sms * occ;
    if (grid > kMaxPartialCtas) grid = kMaxPartialCtas;
    if (grid > ntiles) grid = ntiles;
    *grid_out = (int)grid;
    kern<<<(unsigned)grid, kV2Threads, smem, st>>>(p);
    DMFG_LAUNCHED();
    return DMFG_OK;
}
template <int D>
int dispatch_v2_d(const RolloutParams<float>& p, int noise_kind, bool td, int* grid, cudaStream_t st) {
    // REC = any per-element stream (actions / alpha / alpha') is written; the train step compiles them out
    const bool rec = p.actions != nullptr || p.alpha != nullptr;
    const bool train = !rec && td && p.partials != nullptr && !p.states && !p.rewards && !p.deltas && !p.grads &&
                       !p.pi_final && !p.rewards_in;
    // the policy gradient is only evaluated when something consumes it (critic attached or a grads stream)
    const bool grad = td || p.grads != nullptr;
    if (noise_kind == DMFG_NOISE_PHILOX) {
        if (train) return launch_v2<D, DMFG_NOISE_PHILOX, false, true, true>(p, td, grid, st);
        if (grad) return rec ? launch_v2<D, DMFG_NOISE_PHILOX, true, false, true>(p, td, grid, st)
                             : launch_v2<D, DMFG_NOISE_PHILOX, false, false, true>(p, td, grid, st);
        return rec ? launch_v2<D, DMFG_NOISE_PHILOX, true, false, false>(p, td, grid, st)
                   : launch_v2<D, DMFG_NOISE_PHILOX, false, false, false>(p, td, grid, st);
    }
    return launch_v2<D, DMFG_NOISE_INJECTED, true, false, true>(p, td, grid, st);
}
// dmfg_ac_step: one transition + TD sums + in-launch update (no per-element stream)
template <int D>
int dispatch_v2_step_d(const RolloutParams<float>& p, int noise_kind, int* grid, cudaStream_t st) {
    if (noise_kind == DMFG_NOISE_PHILOX) return launch_v2<D, DMFG_NOISE_PHILOX, false, false, true, true>(p, true, grid, st);
    return launch_v2<D, DMFG_NOISE_INJECTED, false, false, true, true>(p, true, grid, st);
}
int dispatch_v2_step(const RolloutParams<float>& p, int noise_kind, int* grid, cudaStream_t st) {
    switch (p.d) {
        case 15: return dispatch_v2_step_d<15>(p, noise_kind, grid, st);
        case 16: return dispatch_v2_step_d<16>(p, noise_kind, grid, st);
        case 21: return dispatch_v2_step_d<21>(p, noise_kind, grid, st);     // mfg_ac2.py:25
    }
    return fail(DMFG_ERR_UNSUPPORTED, "dmfg_ac_step is built for float streams, d in {15,16,21} (got d=%d)", p.d);
}
int dispatch_v2(const RolloutParams<float>& p, int noise_kind, bool td, int* grid, cudaStream_t st) {
    switch (p.d) {
        case 15: return dispatch_v2_d<15>(p, noise_kind, td, grid, st);
        case 16: return dispatch_v2_d<16>(p, noise_kind, td, grid, st);
        case 20: return dispatch_v2_d<20>(p, noise_kind, td, grid, st);     // test2.py:9
        case 21: return dispatch_v2_d<21>(p, noise_kind, td, grid, st);
        case 32: return dispatch_v2_d<32>(p, noise_kind, td, grid, st);
    }
    return fail(DMFG_ERR_UNSUPPORTED, "no v2 kernel for d=%d", p.d);
}

template <typename R, int NOISE>
int dispatch_fast(const RolloutParams<R>& p, bool td, int* grid, cudaStream_t st) {
    switch (p.d) {
        case 4: return launch_fast<4, 4, R, NOISE>(p, td, grid, st);
        case 15: return launch_fast<15, 16, R, NOISE>(p, td, grid, st);
        case 16: return launch_fast<16, 16, R, NOISE>(p, td, grid, st);
    }
    return fail(DMFG_ERR_UNSUPPORTED, "no fast kernel for d=%d", p.d);
}

// float streams, sampled or injected Gamma variates: the wide kernel (v2 math, warp per population)
template <int NPL, int NOISE, bool GRAD>
int launch_wide_n(const RolloutParams<float>& p, cudaStream_t st) {
    auto kern = rollout_wide_kernel<NPL, NOISE, GRAD>;
    constexpr int WPB = kWideThreads / 32;
    const int pd = (p.d + 1) / 2;
    const size_t smem = (size_t)WPB * (((3 * pd + 1) & ~1) + kWideBatch * 32) * sizeof(double);
    DMFG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0, sms = 0;
    DMFG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWideThreads, smem));
    if (int rc = sm_count(&sms)) return rc;
    if (occ < 1) return fail(DMFG_ERR_CUDA, "rollout_wide_kernel does not fit an SM at d=%d", p.d);
    long long grid = (long long)sms * occ;
    const long long need = (p.B + WPB - 1) / WPB;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, kWideThreads, smem, st>>>(p);
    DMFG_LAUNCHED();
    return DMFG_OK;
}
template <int NOISE>
int launch_wide(const RolloutParams<float>& p, cudaStream_t st) {
    const int pd = (p.d + 1) / 2;
    const bool grad = p.grads != nullptr;
    if (pd <= 32) return grad ? launch_wide_n<1, NOISE, true>(p, st) : launch_wide_n<1, NOISE, false>(p, st);
    if (pd <= 64) return grad ? launch_wide_n<2, NOISE, true>(p, st) : launch_wide_n<2, NOISE, false>(p, st);
    return grad ? launch_wide_n<4, NOISE, true>(p, st) : launch_wide_n<4, NOISE, false>(p, st);
}

template <typename R, int NOISE>
int launch_generic(const RolloutParams<R>& p, cudaStream_t st) {
    if constexpr (std::is_same<R, float>::value && NOISE != DMFG_NOISE_ACTIONS) return launch_wide<NOISE>(p, st);
    auto kern = rollout_generic_kernel<R, NOISE>;
    constexpr int WPB = kGenericThreads / 32;
    const size_t smem = (size_t)WPB * 2 * p.d * (sizeof(double) + sizeof(R));
    DMFG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0, sms = 0;
    DMFG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kGenericThreads, smem));
    if (int rc = sm_count(&sms)) return rc;
    if (occ < 1) return fail(DMFG_ERR_CUDA, "rollout_generic_kernel does not fit an SM at d=%d", p.d);
    long long grid = (long long)sms * occ;
    const long long need = (p.B + WPB - 1) / WPB;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, kGenericThreads, smem, st>>>(p);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

template <typename R>
int run_td(const TdParams<R>& p0, double* acc, double* partials, cudaStream_t st, char* dmma_ws = nullptr) {
    TdParams<R> p = p0;
    const int F = num_features_c(p.d);
    int sms = 0;
    if (int rc = sm_count(&sms)) return rc;
    TdDmmaPlan plan;
    if constexpr (std::is_same<R, float>::value) {
        if (dmma_ws != nullptr) plan = td_dmma_plan(DMFG_F32, p.d, p.T, p.B);
    }
    if constexpr (std::is_same<R, float>::value) {
        if (plan.use) {
            double* U = (double*)(dmma_ws + plan.off_U);
            double* vbuf = (double*)(dmma_ws + plan.off_v);
            const long long N = (long long)(p.T + 1) * p.B, Nt = (long long)p.T * p.B;
            td_unpack_w_kernel<<<(p.d * p.d + 255) / 256, 256, 0, st>>>(p.d, p.w, U);
            DMFG_LAUNCHED();
            const size_t smem = (size_t)(kTdDmmaThreads / 32) * 8 * (p.d + 4) * sizeof(double);
            DMFG_CUDA(cudaFuncSetAttribute(td_values_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            long long grid = ((N + 7) / 8 + 3) / 4;
            if (grid > (long long)sms * 8) grid = (long long)sms * 8;
            td_values_dmma_kernel<<<(unsigned)grid, kTdDmmaThreads, smem, st>>>(p.d, N, p.states, U, p.w, vbuf);
            DMFG_LAUNCHED();
            td_delta_from_values_kernel<<<(unsigned)((Nt + 255) / 256), 256, 0, st>>>(p.T, p.B, p.gamma, p.discount_kind,
                                                                                     p.rewards, vbuf, p.deltas, p.delta_buf);
            DMFG_LAUNCHED();
        }
    }
    bool small_done = false;
    if constexpr (std::is_same<R, float>::value) {
        if (!plan.use && (p.d == 15 || p.d == 16)) {
            // small d: a thread per population, critic weights in shared memory (dmfg_td_small.cuh)
            const unsigned grid = (unsigned)((p.B + kTdSmallThreads - 1) / kTdSmallThreads);
            if (p.d == 15) td_delta_small_kernel<15><<<grid, kTdSmallThreads, 0, st>>>(p);
            else td_delta_small_kernel<16><<<grid, kTdSmallThreads, 0, st>>>(p);
            DMFG_LAUNCHED();
            small_done = true;
        }
    }
    if (!plan.use && !small_done) {
        const long long warps_needed = p.B;
        long long grid = (warps_needed + 3) / 4;
        if (grid > (long long)sms * 16) grid = (long long)sms * 16;
        td_delta_kernel<R><<<(unsigned)grid, 128, 0, st>>>(p);
        DMFG_LAUNCHED();
    }
    if (acc) {
        const long long N = (long long)p.T * p.B;
        long long grid = (N + kTdChunk - 1) / kTdChunk;
        if (grid > kMaxPartialCtas) grid = kMaxPartialCtas;
        p.partials = partials;
        const size_t smem = (size_t)kTdChunk * (p.d + 3) * sizeof(double);
        DMFG_CUDA(cudaFuncSetAttribute(td_gw_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        td_gw_kernel<R><<<(unsigned)grid, 256, smem, st>>>(p, kTdChunk, plan.use ? 1 : 0);
        DMFG_LAUNCHED();
        reduce_partials_kernel<<<(F + 2 + 127) / 128, 128, 0, st>>>(partials, (int)grid, F + 2, acc);
        DMFG_LAUNCHED();
        if constexpr (std::is_same<R, float>::value) {
            if (plan.use) {
                // the quadratic features: Gram blocks on the FP64 tensor cores, split-K partials summed in fixed order
                double* gpart = (double*)(dmma_ws + plan.off_part);
                const long long warps = (long long)plan.nblk * plan.ksplit;
                td_gram_dmma_kernel<<<(unsigned)((warps + 3) / 4), kTdDmmaThreads, 0, st>>>(
                    p.d, N, p.states, p.delta_buf, plan.nblk, plan.ksplit, plan.kchunk, gpart);
                DMFG_LAUNCHED();
                const int Q = p.d * (p.d + 1) / 2;
                td_gram_reduce_kernel<<<(Q + 127) / 128, 128, 0, st>>>(p.d, plan.nblk, plan.ksplit, gpart, acc);
                DMFG_LAUNCHED();
            }
        }
    }
    return DMFG_OK;
}

template <typename R>
int rollout_typed(const dmfg_rollout_args* a, cudaStream_t st) {
    const RolloutWs ws = rollout_ws(a);
    if (ws.total > 0 && (!a->workspace || a->workspace_bytes < ws.total))
        return fail(DMFG_ERR_WORKSPACE, "workspace of %llu bytes needed, %llu given",
                    (unsigned long long)ws.total, (unsigned long long)(a->workspace ? a->workspace_bytes : 0));
    char* wsp = (char*)a->workspace;
    RolloutParams<R> p = make_params<R>(a);
    const bool td = a->w != nullptr;
    const bool accum = td && a->acc != nullptr;
    const int F = num_features_c(a->d);
    if (a->B == 0 || (a->T == 0 && !a->states && !a->pi_final)) {
        if (accum) DMFG_CUDA(cudaMemsetAsync(a->acc, 0, (size_t)(2 + F) * 8, st));
        return DMFG_OK;
    }
    if (use_fast(a)) {
        if (accum) p.partials = (double*)(wsp + ws.partials);
        int grid = 0, rc;
        if constexpr (std::is_same<R, float>::value) {
            if (use_v2(a)) {
                rc = dispatch_v2(p, a->noise_kind, td, &grid, st);
                if (rc) return rc;
                if (accum) {
                    reduce_partials_kernel<<<(F + 2 + 127) / 128, 128, 0, st>>>(p.partials, grid, F + 2, a->acc);
                    DMFG_LAUNCHED();
                }
                return DMFG_OK;
            }
        }
        if (a->noise_kind == DMFG_NOISE_PHILOX) rc = dispatch_fast<R, DMFG_NOISE_PHILOX>(p, td, &grid, st);
        else if (a->noise_kind == DMFG_NOISE_ACTIONS) rc = dispatch_fast<R, DMFG_NOISE_ACTIONS>(p, td, &grid, st);
        else rc = dispatch_fast<R, DMFG_NOISE_INJECTED>(p, td, &grid, st);
        if (rc) return rc;
        if (accum) {
            reduce_partials_kernel<<<(F + 2 + 127) / 128, 128, 0, st>>>(p.partials, grid, F + 2, a->acc);
            DMFG_LAUNCHED();
        }
        return DMFG_OK;
    }
    // generic: rollout (recording what the TD pass needs), then TD from the record
    if (td) {
        if (ws.need_states) p.states = (R*)(wsp + ws.states);
        if (ws.need_rewards) p.rewards = (R*)(wsp + ws.rewards);
        if (ws.need_grads) p.grads = (R*)(wsp + ws.grads);
    }
    p.deltas = nullptr;
    int rc;
    if (a->noise_kind == DMFG_NOISE_PHILOX) rc = launch_generic<R, DMFG_NOISE_PHILOX>(p, st);
    else if (a->noise_kind == DMFG_NOISE_ACTIONS) rc = launch_generic<R, DMFG_NOISE_ACTIONS>(p, st);
    else rc = launch_generic<R, DMFG_NOISE_INJECTED>(p, st);
    if (rc) return rc;
    if (td) {
        TdParams<R> t;
        t.d = a->d; t.T = a->T; t.B = a->B; t.gamma = a->gamma; t.discount_kind = a->discount_kind;
        t.states = p.states; t.rewards = a->rewards_in ? (const R*)a->rewards_in : p.rewards;
        t.grads = p.grads; t.w = a->w; t.deltas = (R*)a->deltas;
        t.delta_buf = (double*)(wsp + ws.delta_buf); t.partials = nullptr;
        return run_td<R>(t, accum ? a->acc : nullptr, accum ? (double*)(wsp + ws.partials) : nullptr, st, wsp + ws.dmma);
    }
    return DMFG_OK;
}

// ---- learners --------------------------------------------------------------
template <int D, int G, typename R, int NOISE>
int launch_learners(const LearnerParams<R>& p, cudaStream_t st) {
    auto kern = learners_fast_kernel<D, G, R, NOISE>;
    const size_t smem = (size_t)(D + 2) * kFastThreads * sizeof(double);
    DMFG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long gpb = kFastThreads / G;
    const long long grid = (p.L + gpb - 1) / gpb;
    kern<<<(unsigned)grid, kFastThreads, smem, st>>>(p);
    DMFG_LAUNCHED();
    return DMFG_OK;
}
// float streams, d = 15 / 16: the learners on the v2 math (single-pass packed row walk)
template <int D, int G, int NOISE>
int launch_learners_v2(const LearnerParams<float>& p, cudaStream_t st) {
    auto kern = learners_v2_kernel<D, G, NOISE>;
    const size_t smem = (size_t)LearnersV2Smem<D, G>::total * sizeof(double);
    DMFG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long gpb = kV2Threads / G;
    const long long grid = (p.L + gpb - 1) / gpb;
    kern<<<(unsigned)grid, kV2Threads, smem, st>>>(p, make_philox_keys(p.seed));
    DMFG_LAUNCHED();
    return DMFG_OK;
}
// float streams, d = 15 / 16 / 21, a few learners: one CTA per learner (latency form)
template <int D, int NOISE>
int launch_learner_cta(const LearnerParams<float>& p, cudaStream_t st) {
    learner_cta_kernel<D, NOISE><<<(unsigned)p.L, LearnerCtaGeom<D>::NT, 0, st>>>(p, make_philox_keys(p.seed));
    DMFG_LAUNCHED();
    return DMFG_OK;
}
template <typename R, int NOISE>
int dispatch_learners(const LearnerParams<R>& p, int layout, cudaStream_t st) {
    if constexpr (std::is_same<R, float>::value) {
        if (p.d == 15 || p.d == 16 || p.d == 21) {
            // up to ~8 learners per SM the CTA form wins (4x per step for one learner, 2.3x at 4 per SM); beyond that the
            // 16-lane form does less work per learner-step
            int sms = 0;
            if (int rc = sm_count(&sms)) return rc;
            const bool cta = layout == DMFG_LEARNERS_CTA ||
                             (layout == DMFG_LEARNERS_AUTO && p.L <= (p.d == 21 ? 2LL : 8LL) * sms);
            if (cta) {
                if (p.d == 15) return launch_learner_cta<15, NOISE>(p, st);
                if (p.d == 16) return launch_learner_cta<16, NOISE>(p, st);
                return launch_learner_cta<21, NOISE>(p, st);
            }
        } else if (layout == DMFG_LEARNERS_CTA) {
            return fail(DMFG_ERR_UNSUPPORTED, "the CTA-per-learner kernel is built for float streams, d in {15,16,21}");
        }
        if (p.d == 15) return launch_learners_v2<15, 16, NOISE>(p, st);
        if (p.d == 16) return launch_learners_v2<16, 16, NOISE>(p, st);
        if (p.d == 21) return launch_learners_v2<21, 32, NOISE>(p, st);     // the reference's default d (mfg_ac2.py:25)
    }
    switch (p.d) {
        case 4: return launch_learners<4, 4, R, NOISE>(p, st);
        case 15: return launch_learners<15, 16, R, NOISE>(p, st);
        case 16: return launch_learners<16, 16, R, NOISE>(p, st);
    }
    return fail(DMFG_ERR_UNSUPPORTED, "dmfg_ac_learners is built for d in {4,15,16} (and 21 for float streams), not d=%d", p.d);
}
template <typename R>
int learners_typed(const dmfg_learners_args* a, cudaStream_t st) {
    LearnerParams<R> p;
    p.d = a->d; p.T = a->T; p.E = a->E; p.episode0 = a->episode0; p.S = a->S;
    p.L = a->L; p.learner_offset = a->learner_offset;
    p.theta = a->theta; p.w = a->w; p.shift = a->shift; p.alpha_scale = a->alpha_scale;
    p.shift_scalar = a->shift_scalar; p.alpha_scale_scalar = a->alpha_scale_scalar;
    p.gamma = a->gamma; p.lr_critic = a->lr_critic; p.lr_actor = a->lr_actor;
    p.constant_lr = a->constant_lr; p.reward_kind = a->reward_kind; p.discount_kind = a->discount_kind;
    p.mat_pi0 = (const R*)a->mat_pi0; p.start_rows = a->start_rows; p.noise_y = (const R*)a->noise_y;
    p.seed = a->seed; p.noise_episode_offset = a->noise_episode_offset; p.theta_trace = a->theta_trace; p.delta_trace = a->delta_trace;
    p.total_reward = a->total_reward; p.pi_final = (R*)a->pi_final;
    if (a->layout < DMFG_LEARNERS_AUTO || a->layout > DMFG_LEARNERS_CTA) return fail(DMFG_ERR_INVALID, "layout %d", a->layout);
    if (a->layout == DMFG_LEARNERS_CTA && !std::is_same<R, float>::value)
        return fail(DMFG_ERR_UNSUPPORTED, "the CTA-per-learner kernel is built for float streams, d in {15,16,21}");
    if (a->noise_kind == DMFG_NOISE_PHILOX) return dispatch_learners<R, DMFG_NOISE_PHILOX>(p, a->layout, st);
    return dispatch_learners<R, DMFG_NOISE_INJECTED>(p, a->layout, st);
}

// ---- testing aids ------------------------------------------------------------
__global__ void gamma_sample_kernel(const float* __restrict__ shape, long long n, NoiseKey nk, float* out) {
    const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = 2 * pair;
    if (i >= n) return;
    const float a0 = shape[i], a1 = (i + 1 < n) ? shape[i + 1] : 1.0f;
    float y0, y1;
    gamma_pair(nk, (uint32_t)pair, a0, a1, 1.0f, y0, y1);
    out[i] = y0;
    if (i + 1 < n) out[i + 1] = y1;
}
template <typename R>
__global__ void digamma_kernel(const R* __restrict__ x, long long n, R* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = digamma(x[i]);
}

}  // namespace

extern "C" {

int dmfg_version(void) { return DMFG_VERSION; }
const char* dmfg_last_error(void) { return g_last_error.c_str(); }
uint64_t dmfg_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
int64_t dmfg_num_features(int32_t d) { return num_features_c(d); }
int64_t dmfg_acc_len(int32_t d) { return 2 + (int64_t)num_features_c(d); }

uint64_t dmfg_rollout_workspace_bytes(const dmfg_rollout_args* a) {
    if (!a || a->struct_size != sizeof(dmfg_rollout_args) || a->d < 1 || a->d > DMFG_MAX_D) return 0;
    return rollout_ws(a).total;
}

int dmfg_rollout(const dmfg_rollout_args* a, void* stream) {
    if (int rc = check_rollout(a)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return a->dtype == DMFG_F64 ? rollout_typed<double>(a, st) : rollout_typed<float>(a, st);
}

uint64_t dmfg_td_workspace_bytes(const dmfg_td_args* a) {
    if (!a || a->struct_size != sizeof(dmfg_td_args) || a->d < 1 || a->d > DMFG_MAX_D) return 0;
    const uint64_t F = (uint64_t)num_features_c(a->d);
    uint64_t off = align_up((uint64_t)a->T * (uint64_t)a->B * 8);
    if (a->acc) off += align_up((uint64_t)kMaxPartialCtas * (2 + F) * 8);
    off += td_dmma_plan(a->dtype, a->d, a->T, a->B).total;
    return off;
}

int dmfg_td_accumulate(const dmfg_td_args* a, void* stream) {
    if (!a) return fail(DMFG_ERR_INVALID, "args is NULL");
    if (a->struct_size != sizeof(dmfg_td_args)) return fail(DMFG_ERR_INVALID, "dmfg_td_args.struct_size mismatch");
    if (a->dtype != DMFG_F32 && a->dtype != DMFG_F64) return fail(DMFG_ERR_INVALID, "dtype %d", a->dtype);
    if (a->d < 1 || a->d > DMFG_MAX_D || a->T < 0 || a->B < 0) return fail(DMFG_ERR_INVALID, "bad d/T/B");
    if (!a->states || !a->rewards || !a->w) return fail(DMFG_ERR_INVALID, "states, rewards and w are required");
    if (a->acc && !a->grads) return fail(DMFG_ERR_INVALID, "acc needs grads");
    const uint64_t need = dmfg_td_workspace_bytes(a);
    if (!a->workspace || a->workspace_bytes < need)
        return fail(DMFG_ERR_WORKSPACE, "workspace of %llu bytes needed", (unsigned long long)need);
    cudaStream_t st = (cudaStream_t)stream;
    const int F = num_features_c(a->d);
    if (a->B == 0 || a->T == 0) {
        if (a->acc) DMFG_CUDA(cudaMemsetAsync(a->acc, 0, (size_t)(2 + F) * 8, st));
        return DMFG_OK;
    }
    char* wsp = (char*)a->workspace;
    double* delta_buf = (double*)wsp;
    double* partials = (double*)(wsp + align_up((uint64_t)a->T * (uint64_t)a->B * 8));
    char* dmma_ws = wsp + align_up((uint64_t)a->T * (uint64_t)a->B * 8) +
                    (a->acc ? align_up((uint64_t)kMaxPartialCtas * (2 + (uint64_t)F) * 8) : 0);
    if (a->dtype == DMFG_F64) {
        TdParams<double> t{a->d, a->T, a->B, a->gamma, a->discount_kind, (const double*)a->states,
                           (const double*)a->rewards, (const double*)a->grads, a->w, (double*)a->deltas,
                           delta_buf, nullptr};
        return run_td<double>(t, a->acc, partials, st);
    }
    TdParams<float> t{a->d, a->T, a->B, a->gamma, a->discount_kind, (const float*)a->states,
                      (const float*)a->rewards, (const float*)a->grads, a->w, (float*)a->deltas,
                      delta_buf, nullptr};
    return run_td<float>(t, a->acc, partials, st, dmma_ws);
}

int dmfg_critic_eval(int32_t dtype, int32_t d, int64_t N, const void* states, const double* w, void* features,
                     void* values, void* stream) {
    if (dtype != DMFG_F32 && dtype != DMFG_F64) return fail(DMFG_ERR_INVALID, "dtype %d", dtype);
    if (d < 1 || d > DMFG_MAX_D || N < 0) return fail(DMFG_ERR_INVALID, "dmfg_critic_eval: bad d/N");
    if (N > 0 && !states) return fail(DMFG_ERR_INVALID, "dmfg_critic_eval: states is NULL");
    if (values && !w) return fail(DMFG_ERR_INVALID, "dmfg_critic_eval: values need w");
    if (N == 0) return DMFG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long F = num_features_c(d);
    if (features) {
        const unsigned grid = (unsigned)((N * F + 255) / 256);
        if (dtype == DMFG_F64) critic_features_kernel<double><<<grid, 256, 0, st>>>(d, N, (const double*)states, (double*)features);
        else critic_features_kernel<float><<<grid, 256, 0, st>>>(d, N, (const float*)states, (float*)features);
        DMFG_LAUNCHED();
    }
    if (values) {
        const unsigned grid = (unsigned)((N + 3) / 4);
        if (dtype == DMFG_F64) critic_value_kernel<double><<<grid, 128, 0, st>>>(d, N, (const double*)states, w, (double*)values);
        else critic_value_kernel<float><<<grid, 128, 0, st>>>(d, N, (const float*)states, w, (float*)values);
        DMFG_LAUNCHED();
    }
    return DMFG_OK;
}

int dmfg_traj_metrics(int32_t dtype, int32_t d, int64_t B, int32_t H, const void* generated, int64_t gen_stride_b,
                      int64_t gen_stride_h, const void* empirical, int64_t emp_stride_b, int64_t emp_stride_h,
                      double* l1, double* jsd, void* stream) {
    if (dtype != DMFG_F32 && dtype != DMFG_F64) return fail(DMFG_ERR_INVALID, "dtype %d", dtype);
    if (d < 1 || B < 0 || H < 1) return fail(DMFG_ERR_INVALID, "dmfg_traj_metrics: bad d/B/H");
    if (B > 0 && (!generated || !empirical)) return fail(DMFG_ERR_INVALID, "dmfg_traj_metrics: NULL input");
    if (B == 0) return DMFG_OK;
    const unsigned grid = (unsigned)((B * H + 3) / 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DMFG_F64)
        traj_metrics_kernel<double><<<grid, 128, 0, st>>>(d, B, H, (const double*)generated, gen_stride_b, gen_stride_h,
                                                        (const double*)empirical, emp_stride_b, emp_stride_h, l1, jsd);
    else
        traj_metrics_kernel<float><<<grid, 128, 0, st>>>(d, B, H, (const float*)generated, gen_stride_b, gen_stride_h,
                                                       (const float*)empirical, emp_stride_b, emp_stride_h, l1, jsd);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

int dmfg_synthetic_check(int32_t dtype, int32_t d, int64_t B, int32_t T, const void* actions, double* l1, double* jsd,
                         void* stream) {
    if (dtype != DMFG_F32 && dtype != DMFG_F64) return fail(DMFG_ERR_INVALID, "dtype %d", dtype);
    if (d < 1 || d > DMFG_MAX_D || B < 0 || T < 1) return fail(DMFG_ERR_INVALID, "dmfg_synthetic_check: bad d/B/T");
    if (B > 0 && (!actions || (!l1 && !jsd))) return fail(DMFG_ERR_INVALID, "dmfg_synthetic_check: NULL argument");
    if (B == 0) return DMFG_OK;
    const unsigned grid = (unsigned)((B + 3) / 4);
    const size_t smem = (size_t)4 * 2 * d * sizeof(double);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DMFG_F64)
        synthetic_check_kernel<double><<<grid, 128, smem, st>>>(d, B, T, (const double*)actions, l1, jsd);
    else
        synthetic_check_kernel<float><<<grid, 128, smem, st>>>(d, B, T, (const float*)actions, l1, jsd);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

int dmfg_ac_apply_update(int32_t d, double* theta_dev, double* w, const double* acc, double lr_critic_eff,
                         double lr_actor_eff, double scale, void* stream) {
    if (d < 1 || d > DMFG_MAX_D || !w || !acc) return fail(DMFG_ERR_INVALID, "dmfg_ac_apply_update: bad argument");
    const int F = num_features_c(d);
    ac_apply_update_kernel<<<(F + 127) / 128, 128, 0, (cudaStream_t)stream>>>(F, theta_dev, w, acc, lr_critic_eff,
                                                                            lr_actor_eff, scale);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

int dmfg_ac_apply_update_dev(int32_t d, double* theta_dev, double* w, const double* acc, const double* lr_dev,
                             double scale, void* stream) {
    if (d < 1 || d > DMFG_MAX_D || !w || !acc || !lr_dev) return fail(DMFG_ERR_INVALID, "dmfg_ac_apply_update_dev: bad argument");
    const int F = num_features_c(d);
    ac_apply_update_dev_kernel<<<(F + 127) / 128, 128, 0, (cudaStream_t)stream>>>(F, theta_dev, w, acc, lr_dev, scale);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

uint64_t dmfg_ac_step_workspace_bytes(const dmfg_rollout_args* a) {
    if (!a || a->struct_size != sizeof(dmfg_rollout_args) || a->d < 1 || a->d > DMFG_MAX_D) return 0;
    return align_up((uint64_t)kMaxPartialCtas * (2 + (uint64_t)num_features_c(a->d)) * 8) + 256;
}

int dmfg_ac_step(const dmfg_rollout_args* a, double* theta_dev, double* w, double lr_critic_eff, double lr_actor_eff,
                 double scale, const double* lr_dev, void* stream) {
    if (int rc = check_rollout(a)) return rc;
    if (!theta_dev || !w) return fail(DMFG_ERR_INVALID, "dmfg_ac_step: theta_dev and w are required (updated in place)");
    if (a->T != 1) return fail(DMFG_ERR_INVALID, "dmfg_ac_step runs ONE transition (T = %d)", a->T);
    if (a->dtype != DMFG_F32 || a->noise_kind == DMFG_NOISE_ACTIONS || !(a->d == 15 || a->d == 16 || a->d == 21))
        return fail(DMFG_ERR_UNSUPPORTED, "dmfg_ac_step is built for float streams, d in {15,16,21}, sampled or injected noise");
    if (a->states || a->actions || a->alpha || a->alpha_deriv || a->rewards || a->deltas || a->grads || a->rewards_in)
        return fail(DMFG_ERR_UNSUPPORTED, "dmfg_ac_step writes pi_final and acc only");
    const uint64_t need = dmfg_ac_step_workspace_bytes(a);
    if (!a->workspace || a->workspace_bytes < need)
        return fail(DMFG_ERR_WORKSPACE, "workspace of %llu bytes needed, %llu given", (unsigned long long)need,
                    (unsigned long long)(a->workspace ? a->workspace_bytes : 0));
    cudaStream_t st = (cudaStream_t)stream;
    const int F = num_features_c(a->d);
    if (a->B == 0) {
        if (a->acc) DMFG_CUDA(cudaMemsetAsync(a->acc, 0, (size_t)(2 + F) * 8, st));
        return DMFG_OK;
    }
    RolloutParams<float> p = make_params<float>(a);
    p.theta_dev = theta_dev;
    p.w = w;
    p.partials = (double*)a->workspace;
    p.fuse_counter = (unsigned int*)((char*)a->workspace + need - 256);
    p.fuse_theta = theta_dev; p.fuse_w = w; p.fuse_acc = a->acc; p.fuse_lr_dev = lr_dev;
    p.fuse_lr_c = lr_critic_eff; p.fuse_lr_a = lr_actor_eff; p.fuse_scale = scale;
    // the ticket counter resets itself after every step; the memset covers a workspace that has never been used
    DMFG_CUDA(cudaMemsetAsync(p.fuse_counter, 0, sizeof(unsigned int), st));
    int grid = 0;
    return dispatch_v2_step(p, a->noise_kind, &grid, st);
}

int dmfg_ac_learners(const dmfg_learners_args* a, void* stream) {
    if (!a) return fail(DMFG_ERR_INVALID, "args is NULL");
    if (a->struct_size != sizeof(dmfg_learners_args))
        return fail(DMFG_ERR_INVALID, "dmfg_learners_args.struct_size mismatch");
    if (a->dtype != DMFG_F32 && a->dtype != DMFG_F64) return fail(DMFG_ERR_INVALID, "dtype %d", a->dtype);
    if (a->L < 0 || a->E < 0 || a->T < 0 || a->S < 1) return fail(DMFG_ERR_INVALID, "bad L/E/T/S");
    if (!a->theta || !a->w || !a->mat_pi0) return fail(DMFG_ERR_INVALID, "theta, w and mat_pi0 are required");
    if (a->noise_kind == DMFG_NOISE_INJECTED && (!a->noise_y || !a->start_rows))
        return fail(DMFG_ERR_INVALID, "noise_kind=INJECTED needs noise_y and start_rows");
    if (a->reward_kind < DMFG_REWARD_NONE || a->reward_kind > DMFG_REWARD_SYNTHETIC)
        return fail(DMFG_ERR_INVALID, "reward_kind %d", a->reward_kind);
    if (a->L == 0 || a->E == 0) return DMFG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return a->dtype == DMFG_F64 ? learners_typed<double>(a, st) : learners_typed<float>(a, st);
}

int dmfg_rollout_host(const dmfg_rollout_args* h, void* stream) {
    if (int rc = check_rollout(h)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    dmfg_rollout_args a = *h;
    const size_t es = esize(a.dtype);
    const size_t B = (size_t)a.B, T = (size_t)a.T, d = (size_t)a.d, F = (size_t)num_features_c(a.d);
    struct Buf { void** dev; const void* host_in; void* host_out; size_t bytes; };
    void *pi0 = nullptr, *y = nullptr, *w = nullptr, *rin = nullptr, *states = nullptr, *actions = nullptr,
         *alpha = nullptr, *deriv = nullptr, *rewards = nullptr, *deltas = nullptr, *grads = nullptr,
         *pif = nullptr, *acc = nullptr, *ws = nullptr;
    Buf bufs[] = {
        {&pi0, h->pi0, nullptr, B * d * es},
        {&y, h->noise_kind != DMFG_NOISE_PHILOX ? h->noise_y : nullptr, nullptr, T * B * d * d * es},
        {&w, h->w, nullptr, F * 8},
        {&rin, h->rewards_in, nullptr, T * B * es},
        {&states, nullptr, h->states, (T + 1) * B * d * es},
        {&actions, nullptr, h->actions, T * B * d * d * es},
        {&alpha, nullptr, h->alpha, T * B * d * d * es},
        {&deriv, nullptr, h->alpha_deriv, T * B * d * d * es},
        {&rewards, nullptr, h->rewards, T * B * es},
        {&deltas, nullptr, h->deltas, T * B * es},
        {&grads, nullptr, h->grads, T * B * es},
        {&pif, nullptr, h->pi_final, B * d * es},
        {&acc, nullptr, h->acc, (2 + F) * 8},
    };
    int rc = DMFG_OK;
    auto cleanup = [&]() {
        for (auto& b : bufs) if (*b.dev) cudaFreeAsync(*b.dev, st);
        if (ws) cudaFreeAsync(ws, st);
    };
    for (auto& b : bufs) {
        if ((b.host_in || b.host_out) && b.bytes > 0) {
            cudaError_t e = cudaMallocAsync(b.dev, b.bytes, st);
            if (e == cudaSuccess && b.host_in) e = cudaMemcpyAsync(*b.dev, b.host_in, b.bytes, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) { cleanup(); return fail(DMFG_ERR_CUDA, "dmfg_rollout_host: %s", cudaGetErrorString(e)); }
        }
    }
    a.pi0 = pi0; a.noise_y = y; a.w = (const double*)w; a.rewards_in = rin; a.states = states; a.actions = actions;
    a.alpha = alpha; a.alpha_deriv = deriv; a.rewards = rewards; a.deltas = deltas; a.grads = grads;
    a.pi_final = pif; a.acc = (double*)acc; a.theta_dev = nullptr;
    a.workspace = nullptr; a.workspace_bytes = 0;
    const uint64_t need = rollout_ws(&a).total;
    if (need) {
        cudaError_t e = cudaMallocAsync(&ws, need, st);
        if (e != cudaSuccess) { cleanup(); return fail(DMFG_ERR_CUDA, "dmfg_rollout_host: %s", cudaGetErrorString(e)); }
        a.workspace = ws; a.workspace_bytes = need;
    }
    rc = dmfg_rollout(&a, stream);
    if (rc == DMFG_OK) {
        for (auto& b : bufs) {
            if (b.host_out && b.bytes > 0) {
                cudaError_t e = cudaMemcpyAsync(b.host_out, *b.dev, b.bytes, cudaMemcpyDeviceToHost, st);
                if (e != cudaSuccess) { rc = fail(DMFG_ERR_CUDA, "dmfg_rollout_host: %s", cudaGetErrorString(e)); break; }
            }
        }
    }
    cleanup();
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc == DMFG_OK && e != cudaSuccess) rc = fail(DMFG_ERR_CUDA, "dmfg_rollout_host: %s", cudaGetErrorString(e));
    return rc;
}

void dmfg_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    const uint4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

int32_t dmfg_gamma_philox_rounds(void) { return DMFG_GAMMA_ROUNDS; }

void dmfg_philox4x32_gamma(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    const uint4 r = philox_gamma(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

int dmfg_gamma_sample(const float* shape, int64_t n, uint64_t seed, uint64_t pop, float* out, void* stream) {
    if (n < 0 || (n > 0 && (!shape || !out))) return fail(DMFG_ERR_INVALID, "dmfg_gamma_sample: bad argument");
    if (n == 0) return DMFG_OK;
    const long long pairs = (n + 1) / 2;
    gamma_sample_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        shape, n, make_noise_key(seed, pop), out);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

int dmfg_digamma(int32_t dtype, const void* x, int64_t n, void* out, void* stream) {
    if (n < 0 || (n > 0 && (!x || !out))) return fail(DMFG_ERR_INVALID, "dmfg_digamma: bad argument");
    if (n == 0) return DMFG_OK;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (dtype == DMFG_F64) digamma_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)x, n, (double*)out);
    else if (dtype == DMFG_F32) digamma_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, n, (float*)out);
    else return fail(DMFG_ERR_INVALID, "dtype %d", dtype);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

}  // extern "C"
