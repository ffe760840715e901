// extern "C" entry points of the IRL half of libdmfg (include/dmfg.h): reward network forward /
// backward, IRL loss, TF-style Adam, Dirichlet log-density for calc_z.
#include <cmath>
#include <cstdint>

#include "dmfg_error.h"
#include "dmfg_rnet.cuh"
#include "dmfg_irl_learner.cuh"

using namespace dmfg;

namespace {

constexpr int kG = 16;           // lanes per transition: 16 for d <= 16 (two transitions per warp), 32 for d <= 32
constexpr int kGmax = 32;
constexpr int kNP = 8;           // padded width of fc3 / fc4: n_fc3, n_fc4 <= 8
constexpr int kMaxRnetCtas = 148 * 2;
constexpr int kLossBlocks = 148 * 4;

inline uint64_t align_up(uint64_t x, uint64_t a = 256) { return (x + a - 1) / a * a; }

int check_rnet(const dmfg_rnet_args* a, bool bwd, bool need_drewards = true) {
    if (!a) return fail(DMFG_ERR_INVALID, "args is NULL");
    if (a->struct_size != sizeof(dmfg_rnet_args))
        return fail(DMFG_ERR_INVALID, "dmfg_rnet_args.struct_size %u != %zu (header mismatch)", a->struct_size,
                    sizeof(dmfg_rnet_args));
    if (a->d < 1 || a->n_fc3 < 1 || a->n_fc4 < 1 || a->N < 0) return fail(DMFG_ERR_INVALID, "bad d/n_fc3/n_fc4/N");
    if (a->d > kGmax || a->n_fc3 > kNP || a->n_fc4 > kNP)
        return fail(DMFG_ERR_UNSUPPORTED, "reward-net kernels are built for d <= %d, n_fc3, n_fc4 <= %d (got %d, %d, %d)",
                    kGmax, kNP, a->d, a->n_fc3, a->n_fc4);
    if (bwd && a->d > kG && a->d != 20 && a->d != 21)
        return fail(DMFG_ERR_UNSUPPORTED, "the reward-net backward kernel is built for d <= 16 and d = 20, 21 (got %d): "
                    "its per-transition tiles are compile-time sized", a->d);
    if (!a->params) return fail(DMFG_ERR_INVALID, "params is NULL");
    if (a->N > 0 && (!a->states || !a->actions)) return fail(DMFG_ERR_INVALID, "states/actions are NULL");
    if (a->dropout < DMFG_DROPOUT_NONE || a->dropout > DMFG_DROPOUT_PHILOX) return fail(DMFG_ERR_INVALID, "dropout %d", a->dropout);
    if (a->dropout != DMFG_DROPOUT_NONE && !(a->keep_prob > 0.f && a->keep_prob <= 1.f))
        return fail(DMFG_ERR_INVALID, "keep_prob must be in (0,1]");
    if (a->dropout == DMFG_DROPOUT_MASKS && a->N > 0 && (!a->mask3 || !a->mask4))
        return fail(DMFG_ERR_INVALID, "dropout=MASKS needs mask3 and mask4");
    if (a->gather_T < 0) return fail(DMFG_ERR_INVALID, "gather_T < 0");
    if (a->gather_T > 0) {
        static_assert(DMFG_MAX_GATHER == kRnetMaxGather, "header and kernel disagree on the gather capacity");
        if (a->N > (int64_t)DMFG_MAX_GATHER * a->gather_T)
            return fail(DMFG_ERR_UNSUPPORTED, "a gathered batch holds <= %d trajectories (N = %lld, gather_T = %d)", DMFG_MAX_GATHER,
                        (long long)a->N, a->gather_T);
        for (int64_t j = 0; j * a->gather_T < a->N; ++j)
            if (a->gather_slots[j] < 0) return fail(DMFG_ERR_INVALID, "gather_slots[%lld] < 0", (long long)j);
    }
    if (!bwd && a->N > 0 && !a->rewards) return fail(DMFG_ERR_INVALID, "rewards is NULL");
    if (bwd) {
        if (!a->grad) return fail(DMFG_ERR_INVALID, "grad is NULL");
        if (need_drewards && a->N > 0 && !a->drewards) return fail(DMFG_ERR_INVALID, "drewards is NULL");
    }
    return DMFG_OK;
}

RnetParams make_params(const dmfg_rnet_args* a) {
    RnetParams p;
    p.d = a->d; p.n3 = a->n_fc3; p.n4 = a->n_fc4; p.N = a->N;
    p.params = a->params; p.states = a->states; p.actions = a->actions;
    p.dropout = a->dropout; p.keep_prob = a->keep_prob; p.mask3 = a->mask3; p.mask4 = a->mask4;
    p.seed = a->seed; p.sample_offset = a->sample_offset;
    p.rewards = a->rewards; p.drewards = a->drewards; p.partials = nullptr;
    p.traj_M = 0; p.t_stride = 0; p.j_stride = 0; p.traj_T = 0; p.zpart = nullptr; p.rpart = nullptr;
    p.gather_T = a->gather_T;
    for (int i = 0; i < kRnetMaxGather; ++i) p.gather[i] = a->gather_T > 0 ? a->gather_slots[i] : 0;
    return p;
}

template <bool BWD, int DS, int N3S, int N4S, bool TRAJ = false, int G = kG, bool GATHER = false>
int rnet_grid(const dmfg_rnet_args* a, int* grid, size_t* smem_bytes, long long traj_M = 0) {
    auto kern = rnet_kernel<G, kNP, BWD, DS, N3S, N4S, TRAJ, GATHER>;
    const RnetLayout L = rnet_layout(a->d, a->n_fc3, a->n_fc4);
    const RnetSmem<G, kNP, BWD, DS> S(a->d, L.total);
    const size_t smem = (size_t)S.total * sizeof(float);
    if (smem > 227 * 1024)
        return fail(DMFG_ERR_UNSUPPORTED, "the reward-net kernel needs %zu bytes of shared memory at d = %d: does not fit an SM", smem, a->d);
    DMFG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0, sms = 0;
    DMFG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kRnetThreads, smem));
    if (int rc = sm_count(&sms)) return rc;
    if (occ < 1) return fail(DMFG_ERR_UNSUPPORTED, "rnet_kernel needs %zu bytes of shared memory: does not fit an SM", smem);
    long long g = (long long)sms * occ;
    if (g > kMaxRnetCtas) g = kMaxRnetCtas;
    const long long ntiles = TRAJ ? traj_M : (a->N + kRnetThreads / G - 1) / (kRnetThreads / G);
    if (g > ntiles) g = ntiles;
    if (g < 1) g = 1;
    *grid = (int)g;
    *smem_bytes = smem;
    return DMFG_OK;
}

// grid + launch of one instantiation.  The gather form (dmfg_rnet_args.gather_T > 0: the batch is slot numbers into a
// pool of resident trajectories) is its own instantiation of the 16-lane kernels, so the address arithmetic of the
// plain form -- the one the large batches run -- carries no trace of it.
template <bool BWD, int DS, int N3S, int N4S, bool TRAJ = false, int G = kG>
int rnet_launch(const dmfg_rnet_args* a, const RnetParams& p, cudaStream_t st, int* grid_out, long long traj_M = 0) {
    int grid = 0;
    size_t smem = 0;
    if (a->gather_T > 0) {
        if constexpr (G == kG) {
            if (int rc = rnet_grid<BWD, DS, N3S, N4S, TRAJ, G, true>(a, &grid, &smem, traj_M)) return rc;
            rnet_kernel<G, kNP, BWD, DS, N3S, N4S, TRAJ, true><<<grid, kRnetThreads, smem, st>>>(p);
        } else {
            return fail(DMFG_ERR_UNSUPPORTED, "gathered batches are built for d <= %d (got %d)", kG, a->d);
        }
    } else {
        if (int rc = rnet_grid<BWD, DS, N3S, N4S, TRAJ, G, false>(a, &grid, &smem, traj_M)) return rc;
        rnet_kernel<G, kNP, BWD, DS, N3S, N4S, TRAJ, false><<<grid, kRnetThreads, smem, st>>>(p);
    }
    DMFG_LAUNCHED();
    if (grid_out) *grid_out = grid;
    return DMFG_OK;
}

template <int D>
int launch_irl_learner(const LearnerParams<float>& p, const IrlLearnerNet& net, int noise_kind, cudaStream_t st) {
    const RnetLayout L = rnet_layout(D, net.n3, net.n4);
    const size_t smem = (size_t)(IrlLearnerSmem<D>::w3a + IrlLearnerSmem<D>::w3size + L.total) * sizeof(float);
    if (noise_kind == DMFG_NOISE_PHILOX) {
        DMFG_CUDA(cudaFuncSetAttribute(irl_learner_cta_kernel<D, DMFG_NOISE_PHILOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        irl_learner_cta_kernel<D, DMFG_NOISE_PHILOX><<<(unsigned)p.L, 128, smem, st>>>(p, make_philox_keys(p.seed), net);
    } else {
        DMFG_CUDA(cudaFuncSetAttribute(irl_learner_cta_kernel<D, DMFG_NOISE_INJECTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        irl_learner_cta_kernel<D, DMFG_NOISE_INJECTED><<<(unsigned)p.L, 128, smem, st>>>(p, make_philox_keys(p.seed), net);
    }
    DMFG_LAUNCHED();
    return DMFG_OK;
}

void reg_ranges(int l1l2, int d, int n3, int n4, int* b0, int* e0, int* b1, int* e1) {
    *b0 = *e0 = *b1 = *e1 = 0;
    if (!l1l2) return;
    const RnetLayout L = rnet_layout(d, n3, n4);
    *b0 = L.w3; *e0 = L.b3;
    *b1 = L.w4; *e1 = L.b4;
}

// the backward launch for the shape of `a` (demonstration form: dL/dr from p.drewards)
int launch_rnet_backward(const dmfg_rnet_args* a, const RnetParams& p, cudaStream_t st, int* grid_out) {
    // the reference's default shape (d = 15, n_fc3 = 8, n_fc4 = 4) with every size a compile-time constant (fits the
    // 255 registers without spills since the fc3 weight gradient moved to the tensor cores); d = 15 with other widths;
    // everything else from the arguments
    if (a->d == 15 && a->n_fc3 == 8 && a->n_fc4 == 4) return rnet_launch<true, 15, 8, 4>(a, p, st, grid_out);
    if (a->d == 15) return rnet_launch<true, 15, 0, 0>(a, p, st, grid_out);
    if (a->d <= kG) return rnet_launch<true, 0, 0, 0>(a, p, st, grid_out);
    if (a->d == 20) return rnet_launch<true, 20, 0, 0, false, 32>(a, p, st, grid_out);
    return rnet_launch<true, 21, 0, 0, false, 32>(a, p, st, grid_out);
}

// the trajectory-mode backward launch (generated half of the one-pass update, d <= 16)
int launch_rnet_backward_gen(const dmfg_rnet_args* a, const RnetParams& p, long long M, cudaStream_t st, int* grid_out) {
    if (a->d == 15 && a->n_fc3 == 8 && a->n_fc4 == 4) return rnet_launch<true, 15, 8, 4, true>(a, p, st, grid_out, M);
    if (a->d == 15) return rnet_launch<true, 15, 0, 0, true>(a, p, st, grid_out, M);
    return rnet_launch<true, 0, 0, 0, true>(a, p, st, grid_out, M);
}

int check_gen(const dmfg_rnet_args* a, const dmfg_irl_gen_args* g) {
    if (!g) return fail(DMFG_ERR_INVALID, "gen args is NULL");
    if (g->struct_size != sizeof(dmfg_irl_gen_args)) return fail(DMFG_ERR_INVALID, "dmfg_irl_gen_args.struct_size mismatch");
    if (!a || a->struct_size != sizeof(dmfg_rnet_args)) return fail(DMFG_ERR_INVALID, "dmfg_rnet_args missing / struct_size mismatch");
    if (g->T < 1 || g->T > kRnetThreads / kG)
        return fail(DMFG_ERR_UNSUPPORTED, "the one-pass update holds a trajectory of T <= %d transitions per CTA (T = %d)",
                    kRnetThreads / kG, g->T);
    if (a && a->d > kG)
        return fail(DMFG_ERR_UNSUPPORTED, "the one-pass update is built for d <= %d (16 lanes per transition, a trajectory per CTA): "
                    "use dmfg_rnet_forward -> dmfg_irl_loss_grad -> dmfg_rnet_backward at d = %d", kG, a->d);
    if (g->M < 1 || g->n_demo < 0 || !(g->num_demo_traj > 0)) return fail(DMFG_ERR_INVALID, "bad M/n_demo/num_demo_traj");
    if (a->N != g->M * (int64_t)g->T) return fail(DMFG_ERR_INVALID, "N must be M*T");
    if (!((g->gen_t_stride == g->M && g->gen_j_stride == 1) || (g->gen_t_stride == 1 && g->gen_j_stride == g->T)))
        return fail(DMFG_ERR_INVALID, "strides must be (M,1) time-major or (1,T) trajectory-major");
    if (!g->loss_out || (g->n_demo > 0 && !g->r_demo)) return fail(DMFG_ERR_INVALID, "loss_out / r_demo are required");
    if (a->gather_T > 0 && !(a->gather_T == g->T && g->gen_t_stride == 1))
        return fail(DMFG_ERR_INVALID, "a gathered generated batch is trajectory-major with gather_T = T");
    return check_rnet(a, true, /*need_drewards=*/false);      // dL/dr is formed in the kernel
}

int adam_consts(int64_t n, int64_t step, double lr, double beta1, double beta2, double eps, int32_t l1l2, int32_t d,
                int32_t n_fc3, int32_t n_fc4, bool want_reg_loss, AdamConsts* c, int* rb) {
    if (step < 1) return fail(DMFG_ERR_INVALID, "dmfg_adam_tf: step counts from 1");
    if (n > INT32_MAX) return fail(DMFG_ERR_UNSUPPORTED, "dmfg_adam_tf: n too large");
    if (l1l2 || want_reg_loss) {
        if (d < 1 || n_fc3 < 1 || n_fc4 < 1 || rnet_layout(d, n_fc3, n_fc4).total != n)
            return fail(DMFG_ERR_INVALID, "dmfg_adam_tf: l1l2 needs (d, n_fc3, n_fc4) matching n");
    }
    reg_ranges(l1l2 || want_reg_loss, d, n_fc3, n_fc4, &rb[0], &rb[1], &rb[2], &rb[3]);      // the regulariser value's ranges
    const double lr_t = lr * std::sqrt(1.0 - std::pow(beta2, (double)step)) / (1.0 - std::pow(beta1, (double)step));
    c->grad_scale = 1.f; c->lr_t = (float)lr_t; c->beta1 = (float)beta1; c->beta2 = (float)beta2;
    c->omb1 = (float)(1.0 - beta1); c->omb2 = (float)(1.0 - beta2); c->eps = (float)eps;
    c->reg0_begin = l1l2 ? rb[0] : 0; c->reg0_end = l1l2 ? rb[1] : 0;
    c->reg1_begin = l1l2 ? rb[2] : 0; c->reg1_end = l1l2 ? rb[3] : 0;
    return DMFG_OK;
}

}  // namespace

extern "C" {

int64_t dmfg_rnet_param_count(int32_t d, int32_t n_fc3, int32_t n_fc4) {
    if (d < 1 || n_fc3 < 1 || n_fc4 < 1) return 0;
    return rnet_layout(d, n_fc3, n_fc4).total;
}

int dmfg_rnet_param_offsets(int32_t d, int32_t n_fc3, int32_t n_fc4, int64_t* o) {
    if (d < 1 || n_fc3 < 1 || n_fc4 < 1 || !o) return fail(DMFG_ERR_INVALID, "dmfg_rnet_param_offsets: bad argument");
    const RnetLayout L = rnet_layout(d, n_fc3, n_fc4);
    o[0] = L.k1; o[1] = L.b1; o[2] = L.k2; o[3] = L.b2; o[4] = L.w3; o[5] = L.b3; o[6] = L.w4; o[7] = L.b4;
    o[8] = L.w5; o[9] = L.b5;
    return DMFG_OK;
}

uint64_t dmfg_rnet_workspace_bytes(const dmfg_rnet_args* a) {
    if (!a || a->struct_size != sizeof(dmfg_rnet_args) || a->d < 1 || a->n_fc3 < 1 || a->n_fc4 < 1) return 0;
    if (!a->grad) return 0;
    // per-CTA partial gradients, then (trajectory mode) the per-CTA sums of exp(R_j) and the 1/Z scalar
    return align_up((uint64_t)kMaxRnetCtas * (uint64_t)rnet_layout(a->d, a->n_fc3, a->n_fc4).total * sizeof(float)) +
           align_up((uint64_t)kMaxRnetCtas * 8) + 256;
}

int dmfg_rnet_forward(const dmfg_rnet_args* a, void* stream) {
    if (int rc = check_rnet(a, false)) return rc;
    if (a->N == 0) return DMFG_OK;
    // the reference's default shape (d = 15, n_fc3 = 8, n_fc4 = 4) and d = 15 with other widths are compiled with
    // those sizes as constants; everything else takes them from the arguments
    cudaStream_t st = (cudaStream_t)stream;
    const RnetParams p = make_params(a);
    if (a->d == 15 && a->n_fc3 == 8 && a->n_fc4 == 4) return rnet_launch<false, 15, 8, 4>(a, p, st, nullptr);
    if (a->d == 15) return rnet_launch<false, 15, 0, 0>(a, p, st, nullptr);
    if (a->d <= kG) return rnet_launch<false, 0, 0, 0>(a, p, st, nullptr);
    if (a->d == 20) return rnet_launch<false, 20, 0, 0, false, 32>(a, p, st, nullptr);      // the reference's action files are 20 x 20 (ac_irl.py:164-200)
    if (a->d == 21) return rnet_launch<false, 21, 0, 0, false, 32>(a, p, st, nullptr);      // mfg_ac2.py:25 default d
    return rnet_launch<false, 0, 0, 0, false, 32>(a, p, st, nullptr);                       // 16 < d <= 32: sizes from the arguments
}

int dmfg_rnet_backward(const dmfg_rnet_args* a, void* stream) {
    if (int rc = check_rnet(a, true)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int total = rnet_layout(a->d, a->n_fc3, a->n_fc4).total;
    if (a->N == 0) {
        if (!a->accumulate) DMFG_CUDA(cudaMemsetAsync(a->grad, 0, (size_t)total * sizeof(float), st));
        return DMFG_OK;
    }
    const uint64_t need = dmfg_rnet_workspace_bytes(a);
    if (!a->workspace || a->workspace_bytes < need)
        return fail(DMFG_ERR_WORKSPACE, "workspace of %llu bytes needed, %llu given", (unsigned long long)need,
                    (unsigned long long)(a->workspace ? a->workspace_bytes : 0));
    int grid = 0;
    RnetParams p = make_params(a);
    p.partials = (float*)a->workspace;
    if (int rc = launch_rnet_backward(a, p, st, &grid)) return rc;
    rnet_reduce_partials_kernel<<<(total + 31) / 32, 32 * kReduceSlices, 0, st>>>(p.partials, grid, total, a->accumulate, a->grad);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

int dmfg_rnet_backward_gen(const dmfg_rnet_args* a, const dmfg_irl_gen_args* g, void* stream) {
    if (int rc = check_gen(a, g)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int total = rnet_layout(a->d, a->n_fc3, a->n_fc4).total;
    const uint64_t need = dmfg_rnet_workspace_bytes(a);
    if (!a->workspace || a->workspace_bytes < need)
        return fail(DMFG_ERR_WORKSPACE, "workspace of %llu bytes needed, %llu given", (unsigned long long)need,
                    (unsigned long long)(a->workspace ? a->workspace_bytes : 0));
    RnetParams p = make_params(a);
    p.partials = (float*)a->workspace;
    char* tail = (char*)a->workspace + align_up((uint64_t)kMaxRnetCtas * (uint64_t)total * sizeof(float));
    p.zpart = (double*)tail;
    float* inv_z = (float*)(tail + align_up((uint64_t)kMaxRnetCtas * 8));
    p.traj_M = g->M; p.traj_T = g->T; p.t_stride = g->gen_t_stride; p.j_stride = g->gen_j_stride;
    int grid = 0;
    if (int rc = launch_rnet_backward_gen(a, p, g->M, st, &grid)) return rc;
    irl_gen_finalize_kernel<<<1, 1024, 0, st>>>(p.zpart, grid, g->r_demo, g->n_demo, g->num_demo_traj, g->M, g->loss_out, inv_z,
                                                g->local_sums ? 0 : 1);
    DMFG_LAUNCHED();
    rnet_reduce_partials_kernel<<<(total + 31) / 32, 32 * kReduceSlices, 0, st>>>(p.partials, grid, total, a->accumulate, a->grad, inv_z);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

int dmfg_irl_dp_finalize(int64_t n, const double* reduced, float* grad, double* loss_out, void* stream) {
    if (n < 1 || n > INT32_MAX || !reduced || !grad) return fail(DMFG_ERR_INVALID, "dmfg_irl_dp_finalize: bad argument");
    irl_dp_finalize_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((int)n, reduced, grad, loss_out);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

int dmfg_irl_learners(const dmfg_learners_args* a, const dmfg_irl_net_args* n, void* stream) {
    if (!a || !n) return fail(DMFG_ERR_INVALID, "args is NULL");
    if (a->struct_size != sizeof(dmfg_learners_args)) return fail(DMFG_ERR_INVALID, "dmfg_learners_args.struct_size mismatch");
    if (n->struct_size != sizeof(dmfg_irl_net_args)) return fail(DMFG_ERR_INVALID, "dmfg_irl_net_args.struct_size mismatch");
    if (a->dtype != DMFG_F32) return fail(DMFG_ERR_UNSUPPORTED, "dmfg_irl_learners is built for float streams");
    if (a->d != 15 && a->d != 16) return fail(DMFG_ERR_UNSUPPORTED, "dmfg_irl_learners is built for d in {15,16} (got %d)", a->d);
    if (n->n_fc3 < 1 || n->n_fc4 < 1 || n->n_fc3 > 8 || n->n_fc4 > 8)
        return fail(DMFG_ERR_UNSUPPORTED, "reward-net widths must be in 1..8 (got %d, %d)", n->n_fc3, n->n_fc4);
    if (a->L < 0 || a->E < 0 || a->T < 0 || a->S < 1) return fail(DMFG_ERR_INVALID, "bad L/E/T/S");
    if (!a->theta || !a->w || !a->mat_pi0 || !n->params) return fail(DMFG_ERR_INVALID, "theta, w, mat_pi0 and net params are required");
    if (a->noise_kind == DMFG_NOISE_INJECTED && (!a->noise_y || !a->start_rows))
        return fail(DMFG_ERR_INVALID, "noise_kind=INJECTED needs noise_y and start_rows");
    if (n->dropout != DMFG_DROPOUT_NONE && n->dropout != DMFG_DROPOUT_PHILOX)
        return fail(DMFG_ERR_UNSUPPORTED, "dropout must be NONE or PHILOX");
    if (n->dropout != DMFG_DROPOUT_NONE && !(n->keep_prob > 0.f && n->keep_prob <= 1.f))
        return fail(DMFG_ERR_INVALID, "keep_prob must be in (0,1]");
    if (a->L == 0 || a->E == 0) return DMFG_OK;
    LearnerParams<float> p;
    p.d = a->d; p.T = a->T; p.E = a->E; p.episode0 = a->episode0; p.S = a->S;
    p.L = a->L; p.learner_offset = a->learner_offset;
    p.theta = a->theta; p.w = a->w; p.shift = a->shift; p.alpha_scale = a->alpha_scale;
    p.shift_scalar = a->shift_scalar; p.alpha_scale_scalar = a->alpha_scale_scalar;
    p.gamma = a->gamma; p.lr_critic = a->lr_critic; p.lr_actor = a->lr_actor;
    p.constant_lr = a->constant_lr; p.reward_kind = DMFG_REWARD_NONE; p.discount_kind = a->discount_kind;
    p.mat_pi0 = (const float*)a->mat_pi0; p.start_rows = a->start_rows; p.noise_y = (const float*)a->noise_y;
    p.seed = a->seed; p.noise_episode_offset = a->noise_episode_offset; p.theta_trace = a->theta_trace;
    p.delta_trace = a->delta_trace; p.total_reward = a->total_reward; p.pi_final = (float*)a->pi_final;
    IrlLearnerNet net;
    net.params = n->params; net.n3 = n->n_fc3; net.n4 = n->n_fc4; net.dropout = n->dropout; net.keep_prob = n->keep_prob;
    net.seed = n->seed; net.sample_offset = n->sample_offset; net.reward_trace = n->reward_trace;
    cudaStream_t st = (cudaStream_t)stream;
    return a->d == 15 ? launch_irl_learner<15>(p, net, a->noise_kind, st) : launch_irl_learner<16>(p, net, a->noise_kind, st);
}

int dmfg_umma_probe(const float* A, const float* B, int32_t a_mn_major, uint32_t a_lbo, uint32_t a_sbo, int32_t b_mn_major,
                    uint32_t b_lbo, uint32_t b_sbo, float* out, void* stream, uint32_t a_desc_lbo, uint32_t a_desc_sbo,
                    uint32_t b_desc_lbo, uint32_t b_desc_sbo) {
    if (!A || !B || !out) return fail(DMFG_ERR_INVALID, "dmfg_umma_probe: NULL argument");
    const bool no_mma = a_mn_major < 0;                       // TMEM store / load round trip only
    DMFG_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    umma_probe_kernel<<<1, 128, 65536, (cudaStream_t)stream>>>(A, B, a_mn_major, a_lbo, a_sbo, b_mn_major, b_lbo, b_sbo,
                                                              no_mma ? 0u : umma::idesc_tf32(128, 16, a_mn_major != 0, b_mn_major != 0), out,
                                                              a_desc_lbo, a_desc_sbo, b_desc_lbo, b_desc_sbo);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

int dmfg_umma_selftest(const float* h, const float* z, int32_t passes, float* out, void* stream) {
    if (passes < 0 || !out || (passes > 0 && (!h || !z))) return fail(DMFG_ERR_INVALID, "dmfg_umma_selftest: bad argument");
    const size_t smem = 2 * umma::W3Grad<4, 2>::kBytesA + 2 * umma::W3Grad<4, 2>::kBytesB;
    DMFG_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(h, z, passes, out);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

uint64_t dmfg_irl_loss_workspace_bytes(int64_t M) {
    if (M < 0) return 0;
    return align_up((uint64_t)M * 8) + align_up((uint64_t)kLossBlocks * 2 * 8);
}

int dmfg_irl_loss_grad(const dmfg_irl_loss_args* a, void* stream) {
    if (!a) return fail(DMFG_ERR_INVALID, "args is NULL");
    if (a->struct_size != sizeof(dmfg_irl_loss_args)) return fail(DMFG_ERR_INVALID, "dmfg_irl_loss_args.struct_size mismatch");
    if (a->T < 1 || a->M < 1 || a->n_demo < 0) return fail(DMFG_ERR_INVALID, "bad T/M/n_demo");
    if (!a->r_gen || (a->n_demo > 0 && !a->r_demo) || !a->loss_out) return fail(DMFG_ERR_INVALID, "r_demo, r_gen and loss_out are required");
    if (!(a->num_demo_traj > 0)) return fail(DMFG_ERR_INVALID, "num_demo_traj must be > 0");
    if (!((a->gen_t_stride == a->M && a->gen_j_stride == 1) || (a->gen_t_stride == 1 && a->gen_j_stride == a->T)))
        return fail(DMFG_ERR_INVALID, "r_gen strides must be (M,1) time-major or (1,T) trajectory-major");
    const uint64_t need = dmfg_irl_loss_workspace_bytes(a->M);
    if (!a->workspace || a->workspace_bytes < need)
        return fail(DMFG_ERR_WORKSPACE, "workspace of %llu bytes needed", (unsigned long long)need);
    cudaStream_t st = (cudaStream_t)stream;
    IrlLossParams p;
    p.n_demo = a->n_demo; p.M = a->M; p.T = a->T; p.gen_t_stride = a->gen_t_stride; p.gen_j_stride = a->gen_j_stride;
    p.num_demo_traj = a->num_demo_traj; p.r_demo = a->r_demo; p.r_gen = a->r_gen; p.log_z = a->log_z;
    p.d_demo = a->d_demo; p.d_gen = a->d_gen;
    p.traj_e = (double*)a->workspace;
    p.partials = (double*)((char*)a->workspace + align_up((uint64_t)a->M * 8));
    p.out = a->loss_out;
    p.reg_loss_scale = 0.0;
    const long long work = a->M > a->n_demo ? a->M : a->n_demo;
    int blocks = (int)((work + 255) / 256);
    if (blocks > kLossBlocks) blocks = kLossBlocks;
    if (blocks < 1) blocks = 1;
    irl_loss_stage1_kernel<<<blocks, 256, 0, st>>>(p);
    DMFG_LAUNCHED();
    irl_loss_stage2_kernel<<<1, 256, 0, st>>>(p, blocks);
    DMFG_LAUNCHED();
    if (a->d_gen) {
        long long b3 = (a->M * a->T + 255) / 256;
        if (b3 > kLossBlocks) b3 = kLossBlocks;
        irl_loss_stage3_kernel<<<(int)b3, 256, 0, st>>>(p);
        DMFG_LAUNCHED();
    }
    return DMFG_OK;
}

int dmfg_adam_tf(int64_t n, float* params, float* m, float* v, const float* grad, double grad_scale, int64_t step,
                 double lr, double beta1, double beta2, double eps, int32_t l1l2, int32_t d, int32_t n_fc3,
                 int32_t n_fc4, double* reg_loss_out, void* stream) {
    if (n < 0 || (n > 0 && (!params || !m || !v || !grad))) return fail(DMFG_ERR_INVALID, "dmfg_adam_tf: bad argument");
    AdamConsts c;
    int rb[4];
    if (int rc = adam_consts(n, step, lr, beta1, beta2, eps, l1l2, d, n_fc3, n_fc4, reg_loss_out != nullptr, &c, rb)) return rc;
    c.grad_scale = (float)grad_scale;
    cudaStream_t st = (cudaStream_t)stream;
    if (reg_loss_out) {
        reg_loss_kernel<<<1, 256, 0, st>>>(params, rb[0], rb[1], rb[2], rb[3], reg_loss_out);
        DMFG_LAUNCHED();
    }
    if (n == 0) return DMFG_OK;
    adam_tf_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>((int)n, params, m, v, grad, c);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

uint64_t dmfg_irl_reward_step_workspace_bytes(const dmfg_rnet_args* a) {
    const uint64_t one = dmfg_rnet_workspace_bytes(a);
    if (!one) return 0;
    // two sets of per-CTA partial gradients (demonstrations, generated), per-CTA sums of exp(R_j) and of r_demo
    return one + align_up((uint64_t)kMaxRnetCtas * (uint64_t)rnet_layout(a->d, a->n_fc3, a->n_fc4).total * sizeof(float)) +
           align_up((uint64_t)kMaxRnetCtas * 8);
}

int dmfg_irl_reward_step(const dmfg_rnet_args* demo, const dmfg_rnet_args* gen, const dmfg_irl_gen_args* g,
                         const dmfg_irl_step_args* s, void* stream) {
    if (!demo || !gen || !g || !s) return fail(DMFG_ERR_INVALID, "dmfg_irl_reward_step: NULL argument");
    if (s->struct_size != sizeof(dmfg_irl_step_args)) return fail(DMFG_ERR_INVALID, "dmfg_irl_step_args.struct_size mismatch");
    if (demo->struct_size != sizeof(dmfg_rnet_args) || gen->struct_size != sizeof(dmfg_rnet_args))
        return fail(DMFG_ERR_INVALID, "dmfg_rnet_args.struct_size mismatch");
    if (g->struct_size != sizeof(dmfg_irl_gen_args)) return fail(DMFG_ERR_INVALID, "dmfg_irl_gen_args.struct_size mismatch");
    if (!s->params || s->params != demo->params || s->params != gen->params)
        return fail(DMFG_ERR_INVALID, "dmfg_irl_reward_step: step->params, demo->params and gen->params must be one vector");
    if (demo->d != gen->d || demo->n_fc3 != gen->n_fc3 || demo->n_fc4 != gen->n_fc4)
        return fail(DMFG_ERR_INVALID, "dmfg_irl_reward_step: demo and gen describe different networks");
    if (demo->N > 0 && !demo->rewards) return fail(DMFG_ERR_INVALID, "dmfg_irl_reward_step: demo->rewards (r_demo out) is required");
    if (!demo->grad) return fail(DMFG_ERR_INVALID, "dmfg_irl_reward_step: demo->grad is required");
    const int64_t total = rnet_layout(demo->d, demo->n_fc3, demo->n_fc4).total;
    cudaStream_t st = (cudaStream_t)stream;
    // Fast form (both halves non-empty, a workspace of dmfg_irl_reward_step_workspace_bytes): the two backward launches
    // write their per-CTA partial gradients side by side and ONE launch does what the chain's two reductions, the loss
    // kernel and the Adam kernel do (irl_step_finish_kernel) -- same summation order and roundings, six launches -> three.
    if (demo->N > 0 && demo->workspace && demo->workspace_bytes >= dmfg_irl_reward_step_workspace_bytes(demo) &&
        !g->local_sums) {
        if (int rc = check_rnet(demo, true)) return rc;
        dmfg_irl_gen_args gg = *g;
        gg.r_demo = demo->rewards;
        gg.n_demo = demo->N;
        dmfg_rnet_args gn = *gen;
        gn.grad = demo->grad;
        if (int rc = check_gen(&gn, &gg)) return rc;
        AdamConsts c;
        int rb[4];
        if (int rc = adam_consts(total, s->step, s->lr, s->beta1, s->beta2, s->eps, s->l1l2, demo->d, demo->n_fc3,
                                 demo->n_fc4, s->reg_loss_out != nullptr, &c, rb)) return rc;
        if (!s->m || !s->v) return fail(DMFG_ERR_INVALID, "dmfg_irl_reward_step: m / v are NULL");
        const uint64_t part_bytes = align_up((uint64_t)kMaxRnetCtas * (uint64_t)total * sizeof(float));
        char* ws = (char*)demo->workspace;
        RnetParams pd = make_params(demo);
        pd.partials = (float*)ws;
        pd.rpart = (double*)(ws + 2 * part_bytes + align_up((uint64_t)kMaxRnetCtas * 8));
        RnetParams pg = make_params(&gn);
        pg.partials = (float*)(ws + part_bytes);
        pg.zpart = (double*)(ws + 2 * part_bytes);
        pg.traj_M = gg.M; pg.traj_T = gg.T; pg.t_stride = gg.gen_t_stride; pg.j_stride = gg.gen_j_stride;
        if (s->reg_loss_out) {                                 // the regulariser value BEFORE the step
            reg_loss_kernel<<<1, 256, 0, st>>>(s->params, rb[0], rb[1], rb[2], rb[3], s->reg_loss_out);
            DMFG_LAUNCHED();
        }
        int grid_d = 0, grid_g = 0;
        if (int rc = launch_rnet_backward(demo, pd, st, &grid_d)) return rc;
        if (int rc = launch_rnet_backward_gen(&gn, pg, gg.M, st, &grid_g)) return rc;
        if (grid_d <= 32 * kReduceSlices && grid_g <= 32 * kReduceSlices) {
            irl_step_finish_kernel<<<(int)((total + 31) / 32), 32 * kReduceSlices, 0, st>>>(
                pd.partials, grid_d, pg.partials, grid_g, (int)total, pg.zpart, pd.rpart, gg.num_demo_traj, gg.M, gg.loss_out,
                demo->grad, s->params, s->m, s->v, c);
            DMFG_LAUNCHED();
            return DMFG_OK;
        }
        // (more CTAs than the finishing block has threads: finish with the chain's own kernels)
        float* inv_z = (float*)(ws + 2 * part_bytes + 2 * align_up((uint64_t)kMaxRnetCtas * 8));
        rnet_reduce_partials_kernel<<<(int)((total + 31) / 32), 32 * kReduceSlices, 0, st>>>(pd.partials, grid_d, (int)total, 0, demo->grad);
        DMFG_LAUNCHED();
        irl_gen_finalize_kernel<<<1, 1024, 0, st>>>(pg.zpart, grid_g, gg.r_demo, gg.n_demo, gg.num_demo_traj, gg.M, gg.loss_out, inv_z, 1);
        DMFG_LAUNCHED();
        rnet_reduce_partials_kernel<<<(int)((total + 31) / 32), 32 * kReduceSlices, 0, st>>>(pg.partials, grid_g, (int)total, 1, demo->grad, inv_z);
        DMFG_LAUNCHED();
        adam_tf_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>((int)total, s->params, s->m, s->v, demo->grad, c);
        DMFG_LAUNCHED();
        return DMFG_OK;
    }
    dmfg_rnet_args dm = *demo;
    dm.accumulate = 0;
    if (int rc = dmfg_rnet_backward(&dm, stream)) return rc;
    dmfg_rnet_args gn = *gen;
    gn.grad = demo->grad;
    gn.accumulate = 1;
    gn.workspace = demo->workspace;
    gn.workspace_bytes = demo->workspace_bytes;
    dmfg_irl_gen_args gg = *g;
    gg.r_demo = demo->rewards;
    gg.n_demo = demo->N;
    if (int rc = dmfg_rnet_backward_gen(&gn, &gg, stream)) return rc;
    return dmfg_adam_tf(total, s->params, s->m, s->v, demo->grad, 1.0, s->step, s->lr, s->beta1, s->beta2, s->eps, s->l1l2,
                        demo->d, demo->n_fc3, demo->n_fc4, s->reg_loss_out, stream);
}

int dmfg_dirichlet_logq(int32_t d, int64_t N, int32_t K, const float* states, const float* actions,
                        const double* thetas, double shift, double* logq, void* stream) {
    if (d < 1 || d > DMFG_MAX_D || N < 0 || K < 1) return fail(DMFG_ERR_INVALID, "dmfg_dirichlet_logq: bad d/N/K");
    if (N > 0 && (!states || !actions || !thetas || !logq)) return fail(DMFG_ERR_INVALID, "dmfg_dirichlet_logq: NULL argument");
    if (N == 0) return DMFG_OK;
    const long long warps = N * K;
    dirichlet_logq_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, (cudaStream_t)stream>>>(d, N, K, states, actions, thetas,
                                                                                        shift, logq);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

int dmfg_irl_log_z(int64_t M, int32_t T, int32_t K, int64_t t_stride, int64_t j_stride, const double* logq,
                   double num_start_samples, float* log_z, void* stream) {
    if (M < 0 || T < 1 || K < 1 || !(num_start_samples > 0)) return fail(DMFG_ERR_INVALID, "dmfg_irl_log_z: bad argument");
    if (M > 0 && (!logq || !log_z)) return fail(DMFG_ERR_INVALID, "dmfg_irl_log_z: NULL argument");
    if (M == 0) return DMFG_OK;
    irl_log_z_kernel<<<(unsigned)((M + 127) / 128), 128, 0, (cudaStream_t)stream>>>(M, T, K, t_stride, j_stride, logq,
                                                                                  num_start_samples, log_z);
    DMFG_LAUNCHED();
    return DMFG_OK;
}

}  // extern "C"
