// br2_cuda.cu -- C ABI (include/br2cuda.h) of the B200 rod-integration path and kernel dispatch.
//
// Kernels: br2_fast.cuh (performance kernel for BR2-structured assemblies) and br2_generic.cuh
// (table-interpreting kernel for everything the tables can express).  One CTA integrates one assembly
// instance for the whole launch; see the kernel headers for the phase structure.  Reference semantics:
// SURVEY.md Appendix A; /root/reference/br2/environment.py:278-285 is the loop this replaces.

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "br2cuda.h"

#ifndef BR2_PI
#define BR2_PI 3.141592653589793
#endif

#ifndef BR2_MB128
#define BR2_MB128 3
#endif
#include "br2_fast.cuh"
#include "br2_generic.cuh"
#include "br2_metrics.cuh"
#include "br2_tables.cuh"
#include "br2_tile.cuh"

using br2::DevTables;
using br2::StepParams;

namespace {

// DFMA throughput probe: 8 independent chains per thread, all operands in registers.
__global__ void __launch_bounds__(256) dfma_probe_kernel(double *out, long long iters, double seed) {
    double a0 = seed + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 0.999999, c = 1e-7;
    for (long long i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c);
            a1 = fma(a1, m, c);
            a2 = fma(a2, m, c);
            a3 = fma(a3, m, c);
            a4 = fma(a4, m, c);
            a5 = fma(a5, m, c);
            a6 = fma(a6, m, c);
            a7 = fma(a7, m, c);
        }
    }
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---------------------------------------------------------------------------------------------
// host side of the ABI
// ---------------------------------------------------------------------------------------------
thread_local std::string g_error;

int fail(int code, const std::string &msg) {
    g_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t err__ = (expr);                                                                 \
        if (err__ != cudaSuccess)                                                                   \
            return fail(BR2_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));          \
    } while (0)

typedef void (*kernel_fn)(const StepParams);

struct KernelChoice {
    int threads;
    kernel_fn fn;
    kernel_fn fn_pi;  // tile kernels: the instantiation that reads per-instance material constants (or nullptr)
};

#ifndef BR2_TILE_C
#define BR2_TILE_C 2
#endif
#ifndef BR2_TILE_MB64
#define BR2_TILE_MB64 6
#endif
#ifndef BR2_TILE_MIN_CHUNK_STEPS
#define BR2_TILE_MIN_CHUNK_STEPS 125  // a launch is cut into at most BR2_MAX_CHUNKS chunks of at least this many steps
#endif
// split-phase mbarriers (true) or CTA barriers (false) at the step's three exchange points, per kernel family
// (measured: see DESIGN.md); the tip-to-base joint data is pushed through (distributed) shared memory either way
#ifndef BR2_TILE_SPLIT_SINGLE
#define BR2_TILE_SPLIT_SINGLE false
#endif
#ifndef BR2_TILE_SPLIT_EXTRA
#define BR2_TILE_SPLIT_EXTRA false
#endif
#ifndef BR2_TILE_SPLIT_CLUSTER
#define BR2_TILE_SPLIT_CLUSTER true
#endif
// variant 3: tile kernel (BR2_TILE_C slots per thread, uniform constants); `threads` = CTA size
// template flags: <slots per thread, CTA threads, min CTAs per SM, extra joints, cluster, exact arithmetic, split-phase barriers>
const KernelChoice kTileKernels[] = {  // single-segment assemblies (ring joints only)
    {64, br2::br2_step_tile_kernel<BR2_TILE_C, 64, BR2_TILE_MB64, false, false, false, BR2_TILE_SPLIT_SINGLE>,
     br2::br2_step_tile_kernel<BR2_TILE_C, 64, BR2_TILE_MB64, false, false, false, BR2_TILE_SPLIT_SINGLE, true>},
    {128, br2::br2_step_tile_kernel<BR2_TILE_C, 128, 3, false, false, false, BR2_TILE_SPLIT_EXTRA>,
     br2::br2_step_tile_kernel<BR2_TILE_C, 128, 3, false, false, false, BR2_TILE_SPLIT_EXTRA, true>},
    {320, br2::br2_step_tile_kernel<BR2_TILE_C, 320, 1, false, false, false, BR2_TILE_SPLIT_CLUSTER>,
     br2::br2_step_tile_kernel<BR2_TILE_C, 320, 1, false, false, false, BR2_TILE_SPLIT_CLUSTER, true>},
};
const KernelChoice kTileExtraKernels[] = {  // several segments joined tip to base, one CTA
    {128, br2::br2_step_tile_kernel<BR2_TILE_C, 128, 3, true, false, false, BR2_TILE_SPLIT_EXTRA>,
     br2::br2_step_tile_kernel<BR2_TILE_C, 128, 3, true, false, false, BR2_TILE_SPLIT_EXTRA, true>},
    {192, br2::br2_step_tile_kernel<BR2_TILE_C, 192, 2, true, false, false, BR2_TILE_SPLIT_EXTRA>,
     br2::br2_step_tile_kernel<BR2_TILE_C, 192, 2, true, false, false, BR2_TILE_SPLIT_EXTRA, true>},
    {256, br2::br2_step_tile_kernel<BR2_TILE_C, 256, 1, true, false, false, BR2_TILE_SPLIT_EXTRA>,
     br2::br2_step_tile_kernel<BR2_TILE_C, 256, 1, true, false, false, BR2_TILE_SPLIT_EXTRA, true>},
};
// one CTA per group of ring-connected rods, launched as a thread-block cluster per instance
const KernelChoice kTileClusterKernels[] = {
    {192, br2::br2_step_tile_kernel<BR2_TILE_C, 192, 2, true, true, false, BR2_TILE_SPLIT_CLUSTER>,
     br2::br2_step_tile_kernel<BR2_TILE_C, 192, 2, true, true, false, BR2_TILE_SPLIT_CLUSTER, true>},
    {320, br2::br2_step_tile_kernel<BR2_TILE_C, 320, 1, true, true, false, BR2_TILE_SPLIT_CLUSTER>,
     br2::br2_step_tile_kernel<BR2_TILE_C, 320, 1, true, true, false, BR2_TILE_SPLIT_CLUSTER, true>},
};
const KernelChoice kTileClusterExactKernels[] = {  // same, exact arithmetic: the replay path of cluster-sized assemblies
    {192, br2::br2_step_tile_kernel<BR2_TILE_C, 192, 2, true, true, true, BR2_TILE_SPLIT_CLUSTER>,
     br2::br2_step_tile_kernel<BR2_TILE_C, 192, 2, true, true, true, BR2_TILE_SPLIT_CLUSTER, true>},
    {320, br2::br2_step_tile_kernel<BR2_TILE_C, 320, 1, true, true, true, BR2_TILE_SPLIT_CLUSTER>,
     br2::br2_step_tile_kernel<BR2_TILE_C, 320, 1, true, true, true, BR2_TILE_SPLIT_CLUSTER, true>},
};
const int kTileMaxCtaThreads = 320, kTileOneCtaThreads = 256;

// variant 0: generic, 1: fast with per-slot constants, 2: fast with uniform constants
const KernelChoice kKernels[3][6] = {
    {{128, br2::br2_step_generic_kernel<128, 4>, nullptr}, {256, br2::br2_step_generic_kernel<256, 2>, nullptr},
     {384, br2::br2_step_generic_kernel<384, 1>, nullptr}, {512, br2::br2_step_generic_kernel<512, 1>, nullptr},
     {768, br2::br2_step_generic_kernel<768, 1>, nullptr}, {1024, br2::br2_step_generic_kernel<1024, 1>, nullptr}},
    {{128, br2::br2_step_fast_kernel<128, BR2_MB128, false>, nullptr}, {256, br2::br2_step_fast_kernel<256, 2, false>, nullptr},
     {384, br2::br2_step_fast_kernel<384, 1, false>, nullptr}, {512, br2::br2_step_fast_kernel<512, 1, false>, nullptr},
     {768, br2::br2_step_fast_kernel<768, 1, false>, nullptr}, {1024, br2::br2_step_fast_kernel<1024, 1, false>, nullptr}},
    {{128, br2::br2_step_fast_kernel<128, BR2_MB128, true>, nullptr}, {256, br2::br2_step_fast_kernel<256, 2, true>, nullptr},
     {384, br2::br2_step_fast_kernel<384, 1, true>, nullptr}, {512, br2::br2_step_fast_kernel<512, 1, true>, nullptr},
     {768, br2::br2_step_fast_kernel<768, 1, true>, nullptr}, {1024, br2::br2_step_fast_kernel<1024, 1, true>, nullptr}},
};

}  // namespace

struct br2_handle_s {
    int device = 0;
    DevTables t{};
    std::vector<void *> owned;  // device allocations of the tables
    double *state = nullptr;
    const double *pressures = nullptr;
    long long batch = 0;
    KernelChoice kernel{};
    size_t smem_bytes = 0;
    KernelChoice replay{};  // generic kernel used behind the fast one
    KernelChoice exact{};   // exact-arithmetic instantiation of the cluster tile kernel (assemblies the generic kernel cannot hold)
    size_t replay_smem = 0;
    int variant = 0;        // 0 generic, 1 fast, 2 fast uniform
    int auto_variant = 0;   // what br2_create picked
    bool fast_ok = false, fast_uniform = false, has_extra = false, tile_ok = false, fast_point = false;
    br2::TilePlan tile{};
    std::string tile_why;
    double uni[BR2_FC_COUNT] = {0};
    int *status = nullptr;            // [batch] fast-kernel validity flags (library-owned)
    long long status_batch = 0;
    unsigned long long *work = nullptr;  // tile kernel's queue: [0] item counter, then int progress[batch]
    int tile_capacity = 0;            // CTAs of the tile kernel resident on the device
    unsigned long long *replays = nullptr;  // device counter
    const double *inst_uni = nullptr;          // per-instance material constants (borrowed device buffers), or nullptr
    const double *inst_act_d = nullptr;
    const double *inst_joint_scale = nullptr;
    const double *cls_uni = nullptr;           // [n_cls][BR2_FC_COUNT] constants per class of rods (assemblies with several)
    int n_act = 0;
    double *scratch_state = nullptr;  // br2_step_host
    double *scratch_pressures = nullptr;
    long long scratch_batch = 0;
};

namespace {

template <typename Tp>
int upload(br2_handle h, const Tp *host, size_t count, const Tp **dev) {
    void *d = nullptr;
    size_t bytes = (count ? count : 1) * sizeof(Tp);
    CUDA_TRY(cudaMalloc(&d, bytes));
    h->owned.push_back(d);
    if (count) CUDA_TRY(cudaMemcpy(d, host, count * sizeof(Tp), cudaMemcpyHostToDevice));
    *dev = static_cast<const Tp *>(d);
    return BR2_OK;
}

int validate(const br2_program *p) {
    if (!p) return fail(BR2_EINVAL, "program is NULL");
    if (p->abi_version != BR2_ABI_VERSION) return fail(BR2_EINVAL, "br2_program.abi_version mismatch");
    if (p->n_slots <= 0 || p->n_rods <= 0) return fail(BR2_EINVAL, "empty assembly");
    if (p->n_slots > 8 * 2 * kTileMaxCtaThreads)
        return fail(BR2_ELIMIT, "assembly has " + std::to_string(p->n_slots) + " node slots; one cluster holds at most " +
                                    std::to_string(8 * 2 * kTileMaxCtaThreads));
    if (p->n_actions > 32) return fail(BR2_ELIMIT, "more than 32 actions");
    int slots = 0;
    for (int r = 0; r < p->n_rods; ++r) {
        const int32_t *ri = p->rod_i + r * BR2_ROD_NI;
        if (ri[BR2_ROD_N] < 1) return fail(BR2_EINVAL, "rod with no elements");
        if (ri[BR2_ROD_SLOT0] != slots) return fail(BR2_EINVAL, "rod slots must be contiguous and ordered");
        slots += ri[BR2_ROD_N] + 1;
        if (ri[BR2_ROD_F_BEGIN] < 0 || ri[BR2_ROD_F_END] > p->n_forcing || ri[BR2_ROD_F_BEGIN] > ri[BR2_ROD_F_END])
            return fail(BR2_EINVAL, "forcing range out of bounds");
        if (ri[BR2_ROD_C_BEGIN] < 0 || ri[BR2_ROD_C_END] > p->n_constraints || ri[BR2_ROD_C_BEGIN] > ri[BR2_ROD_C_END])
            return fail(BR2_EINVAL, "constraint range out of bounds");
    }
    if (slots != p->n_slots) return fail(BR2_EINVAL, "n_slots does not match the rod table");
    for (int f = 0; f < p->n_forcing; ++f) {
        const int32_t *fi = p->forcing_i + f * BR2_F_NI;
        if (fi[BR2_F_TYPE] < BR2_FORCE_GRAVITY || fi[BR2_F_TYPE] > BR2_FORCE_POINT)
            return fail(BR2_EINVAL, "unknown forcing type");
        if ((fi[BR2_F_TYPE] == BR2_FORCE_TWIST || fi[BR2_F_TYPE] == BR2_FORCE_BEND) &&
            (fi[BR2_F_ACTION] < 0 || fi[BR2_F_ACTION] >= p->n_actions))
            return fail(BR2_EINVAL, "forcing row references an unknown action");
    }
    for (int j = 0; j < p->n_joints; ++j) {
        const int32_t *ji = p->joint_i + j * BR2_J_NI;
        if (ji[BR2_J_TYPE] != BR2_JOINT_SURFACE && ji[BR2_J_TYPE] != BR2_JOINT_TIP_TO_BASE)
            return fail(BR2_EINVAL, "unknown joint type");
        for (int q = BR2_J_S1; q <= BR2_J_Q2; ++q)
            if (ji[q] < 0 || ji[q] + 1 >= p->n_slots) return fail(BR2_EINVAL, "joint slot out of range");
    }
    return BR2_OK;
}

// Plan of the thread-coarsened kernel (br2_tile.cuh), or the reason the assembly does not qualify.
struct TileExtraTables {
    std::vector<int32_t> xi, xj_i;   // xi: staging rows (BR2_TILE_XI_*), packed into xpack / xitems at the end
    std::vector<int32_t> xpack, xjt, xe_mask;
    std::vector<double> xj_d;
};

bool make_tile_plan(const br2_program *p, int C, br2::TilePlan *plan, TileExtraTables *xt, std::string *why) {
    *plan = br2::TilePlan{};
    plan->C = C;
    if (!p->fast_ok) return *why = "assembly outside the fast path's pattern", false;
    if (p->n_rods > BR2_TILE_MAX_RODS) return *why = "too many rods", false;
    const size_t S = (size_t)p->n_slots;
    // Element constants: one set per CLASS of rods (all elements of a rod equal, rods compared by their first slot; the same
    // 1e-12 the lowering uses for "uniform").  One class: the kernels take the constants from the constant bank; several
    // (segments of different element count or material): the instantiations that read a shared-memory copy, indexed by class.
    plan->n_cls = 0;
    {
        int first_of_cls[BR2_TILE_MAX_CLS];
        auto same = [&](size_t a, size_t b) {
            for (int i = 0; i < BR2_FC_PIN; ++i) {
                const double x = p->fast_c[(size_t)i * S + a], y = p->fast_c[(size_t)i * S + b];
                if (fabs(x - y) > 1e-12 * std::max(fabs(x), 1e-300)) return false;
            }
            return true;
        };
        for (int r = 0; r < p->n_rods; ++r) {
            const int32_t *ri = p->rod_i + r * BR2_ROD_NI;
            const size_t s0 = (size_t)ri[BR2_ROD_SLOT0];
            for (int k = 1; k < ri[BR2_ROD_N]; ++k)  // (element slots; the node-only last slot carries padding)
                if (!same(s0, s0 + k)) return *why = "element constants vary along a rod", false;
            int cls = -1;
            for (int c = 0; c < plan->n_cls && cls < 0; ++c)
                if (same((size_t)first_of_cls[c], s0)) cls = c;
            if (cls < 0) {
                if (plan->n_cls == BR2_TILE_MAX_CLS) return *why = "more than " + std::to_string(BR2_TILE_MAX_CLS) + " classes of rod constants", false;
                cls = plan->n_cls++;
                first_of_cls[cls] = (int)s0;
            }
            plan->rod_cls[r] = cls;
            plan->cls_slot[cls] = (int)s0;
        }
        if (plan->n_cls == 1 && !p->fast_uniform) return *why = "assembly constants are not uniform", false;  // (cannot happen)
    }
    // ring-connected rods (a segment) must share a CTA: union-find over the owned surface joints
    int group[BR2_TILE_MAX_RODS];
    for (int r = 0; r < p->n_rods; ++r) group[r] = r;
    auto find = [&](int r) {
        while (group[r] != r) r = group[r] = group[group[r]];
        return r;
    };
    for (size_t sl = 0; sl < S; ++sl) {
        const int j = p->fast_i[sl * BR2_FAST_NI + BR2_FAST_OWN_JOINT];
        if (j < 0) continue;
        const int a = find(p->slot_rod[p->joint_i[(size_t)j * BR2_J_NI + BR2_J_S1]]), b = find(p->slot_rod[sl]);
        if (a != b) group[a > b ? a : b] = a > b ? b : a;
    }
    int total_threads = 0;
    for (int r = 0; r < p->n_rods; ++r) total_threads += (p->rod_i[r * BR2_ROD_NI + BR2_ROD_N] + 1 + C - 1) / C;
    int one_cta_threads = kTileOneCtaThreads;
    if (const char *env = getenv("BR2_TILE_ONE_CTA_THREADS")) one_cta_threads = atoi(env);  // tests force the cluster path
    const bool split = total_threads > one_cta_threads;
    int rank_of_group[BR2_TILE_MAX_RODS];
    for (int r = 0; r < p->n_rods; ++r) rank_of_group[r] = -1;
    plan->n_ctas = 0;
    for (int r = 0; r < p->n_rods; ++r) {
        const int g = find(r);
        if (rank_of_group[g] < 0) rank_of_group[g] = split ? plan->n_ctas++ : 0;
        if (plan->n_ctas > 8) return *why = "more than 8 groups of ring-connected rods", false;
        const int rank = rank_of_group[g];
        plan->rod_cta[r] = rank;
        plan->rod_t0[r] = plan->cta_threads[rank];
        plan->cta_threads[rank] += (p->rod_i[r * BR2_ROD_NI + BR2_ROD_N] + 1 + C - 1) / C;
        plan->partner[r] = plan->owner[r] = -1;
    }
    if (!split) plan->n_ctas = 1;
    int threads = 0;  // of the largest CTA
    for (int c = 0; c < plan->n_ctas; ++c) threads = plan->cta_threads[c] > threads ? plan->cta_threads[c] : threads;
    if (threads > kTileMaxCtaThreads)
        return *why = "a group of ring-connected rods needs " + std::to_string(threads) + " threads (max " +
                      std::to_string(kTileMaxCtaThreads) + ")", false;
    plan->rod_t0[BR2_TILE_MAX_RODS] = threads;
    plan->xi_stride = ((threads + 63) / 64) * 64 + 64;
    auto close = [](double a, double b, double tol) { return fabs(a - b) <= tol; };
    for (int r = 0; r < p->n_rods; ++r) {
        const int32_t *ri = p->rod_i + r * BR2_ROD_NI;
        const int n = ri[BR2_ROD_N], slot0 = ri[BR2_ROD_SLOT0];
        int owned = 0;
        for (int e = 0; e < n; ++e) {
            const int j = p->fast_i[(size_t)(slot0 + e) * BR2_FAST_NI + BR2_FAST_OWN_JOINT];
            if (j < 0) continue;
            ++owned;
            const int32_t *ji = p->joint_i + (size_t)j * BR2_J_NI;
            const double *jd = p->joint_d + (size_t)j * BR2_J_ND;
            const int r1 = p->slot_rod[ji[BR2_J_S1]];
            const int e1 = ji[BR2_J_S1] - p->rod_i[r1 * BR2_ROD_NI + BR2_ROD_SLOT0];
            if (ji[BR2_J_S2] != slot0 + e || e1 != e) return *why = "a surface joint connects different element indices", false;
            if (p->rod_i[r1 * BR2_ROD_NI + BR2_ROD_N] != n) return *why = "joined rods differ in element count", false;
            if (owned == 1) {
                if (plan->owner[r1] >= 0) return *why = "a rod is rod one of two rods", false;
                plan->partner[r] = r1;
                plan->owner[r1] = r;
                for (int i = 0; i < 3; ++i) {
                    plan->jd[r][i] = jd[6 + i];       // dv2: this rod's side
                    plan->jd[r1][4 + i] = jd[3 + i];  // dv1: the partner's side
                }
                plan->jd[r][3] = plan->jd[r1][7] = 0.5 * jd[9];
            } else {
                if (plan->partner[r] != r1) return *why = "a rod owns joints with two rods", false;
                bool same = close(0.5 * jd[9], plan->jd[r][3], 1e-16);
                for (int i = 0; i < 3; ++i)
                    same = same && close(jd[6 + i], plan->jd[r][i], 1e-15) && close(jd[3 + i], plan->jd[r1][4 + i], 1e-15);
                if (!same) return *why = "joint descriptors vary along a rod", false;
            }
            // the per-joint parameters must be the uniform ones the kernel reads from the constant bank
            const double *fc = p->fast_c + slot0;  // (the owning rod's constants)
            if (jd[0] != fc[(size_t)BR2_FC_JK * S] || jd[1] != fc[(size_t)BR2_FC_JNU * S] || jd[2] != fc[(size_t)BR2_FC_JKREP * S])
                return *why = "joint parameters are not uniform", false;
        }
        if (owned != 0 && owned != n) return *why = "only some elements of a rod own a joint", false;
    }
    plan->round_section = 1;
    for (int c = 0; c < plan->n_cls; ++c) {
        const double *fc = p->fast_c + plan->cls_slot[c];
        const double l0 = fc[(size_t)BR2_FC_LN_CR * S], l1 = fc[(size_t)(BR2_FC_LN_CR + 1) * S], l2 = fc[(size_t)(BR2_FC_LN_CR + 2) * S];
        if (!(l0 == l1 && fabs(l0 - 2.0 * l2) <= 1e-13 * fabs(l0))) plan->round_section = 0;
    }
    for (int r = 0; r < p->n_rods; ++r)
        if (plan->jd[r][2] != 0.0 || plan->jd[r][6] != 0.0)
            return *why = "a surface joint's direction vector has a component along the rod axis", false;

    // ---- extra joints (everything that is not an owned ring joint): tip-to-base joints only ----------------
    const int XN = BR2_TILE_XI_N(C);
    const int xrows = plan->n_ctas * plan->xi_stride;  // rows for the idle threads of the CTAs too
    xt->xi.assign((size_t)xrows * XN, -1);
    for (int t = 0; t < xrows; ++t)
        for (int j = 0; j < C; ++j) xt->xi[(size_t)t * XN + BR2_TILE_XI_NCNT(C, j)] = 0;
    int slot_rank = 0, slot_row = 0;  // set by thread_of: cluster rank and thread (= shared-memory row) of the slot
    auto thread_of = [&](int slot, int *j) {  // row of the slot's thread in xi
        const int r = p->slot_rod[slot];
        const int k = slot - p->rod_i[r * BR2_ROD_NI + BR2_ROD_SLOT0];
        *j = k % C;
        slot_rank = plan->rod_cta[r];
        slot_row = plan->rod_t0[r] + k / C;
        return slot_rank * plan->xi_stride + slot_row;
    };
    std::vector<int> xe_of_slot(S, -1);
    std::vector<int> xe_rank;  // the CTA that holds (and publishes) the element
    auto xe_id = [&](int slot) {
        if (xe_of_slot[slot] < 0) {
            xe_of_slot[slot] = plan->n_xe++;
            int j;
            const int t = thread_of(slot, &j);
            xt->xi[(size_t)t * XN + BR2_TILE_XI_PUB(C, j)] = xe_of_slot[slot];
            xe_rank.push_back(slot_rank);
        }
        return xe_of_slot[slot];
    };
    std::vector<char> owned_joint(p->n_joints, 0);
    for (size_t sl = 0; sl < S; ++sl) {
        const int j = p->fast_i[sl * BR2_FAST_NI + BR2_FAST_OWN_JOINT];
        if (j >= 0) owned_joint[j] = 1;
    }
    // Every CTA evaluates the extra joints that act on its own nodes -- a joint between two CTAs of a cluster is evaluated
    // by both, with bit-identical arithmetic on the same pushed data, so no force ever travels between CTAs -- and an
    // element's data is pushed to exactly the CTAs that read it (joint evaluation, director / omega overwrite).
    std::vector<int32_t> xe_mask;      // per published element: consumer ranks
    std::vector<int> xj_ranks;         // per extra joint: bit mask of the evaluating ranks
    auto consume = [&](int xe, int rank) {
        if ((int)xe_mask.size() <= xe) xe_mask.resize(xe + 1, 0);
        xe_mask[xe] |= 1 << rank;
    };
    for (int j = 0; j < p->n_joints; ++j) {
        if (owned_joint[j]) continue;
        const int32_t *ji = p->joint_i + (size_t)j * BR2_J_NI;
        const double *jd = p->joint_d + (size_t)j * BR2_J_ND;
        if (ji[BR2_J_TYPE] != BR2_JOINT_TIP_TO_BASE) return *why = "a surface joint outside the ring pattern", false;
        const int xj = plan->n_xj++;
        int j1, j2;
        thread_of(ji[BR2_J_S1], &j1);
        const int rank1 = slot_rank, row1 = slot_row;
        thread_of(ji[BR2_J_S2], &j2);
        const int rank2 = slot_rank, row2 = slot_row;
        const int32_t row[12] = {rank1, row1, j1, rank2, row2, j2, xe_id(ji[BR2_J_Q1]), xe_id(ji[BR2_J_Q2]),
                                 xe_id(ji[BR2_J_S1]), xe_id(ji[BR2_J_S2]), 0, 0};
        xt->xj_i.insert(xt->xj_i.end(), row, row + 12);
        const double drow[8] = {jd[0], jd[1], jd[3], jd[4], jd[5], jd[6], jd[7], jd[8]};
        xt->xj_d.insert(xt->xj_d.end(), drow, drow + 8);
        // the four nodes of the two elements: rod one +F/2, rod two -F/2
        const int node_slot[4] = {ji[BR2_J_S1], ji[BR2_J_S1] + 1, ji[BR2_J_S2], ji[BR2_J_S2] + 1};
        int ranks = 0;
        for (int q = 0; q < 4; ++q) {
            int jn;
            const int tn = thread_of(node_slot[q], &jn);
            ranks |= 1 << slot_rank;
            int32_t &cnt = xt->xi[(size_t)tn * XN + BR2_TILE_XI_NCNT(C, jn)];
            if (cnt >= BR2_FAST_NODE_ITEMS) return *why = "a node receives too many extra joint forces", false;
            xt->xi[(size_t)tn * XN + BR2_TILE_XI_NITEM(C, jn, cnt)] = (xj << 1) | (q >= 2 ? 1 : 0);
            ++cnt;
        }
        xj_ranks.push_back(ranks);
        for (int r = 0; r < plan->n_ctas; ++r)
            if (ranks & (1 << r))
                for (int q = 6; q < 10; ++q) consume(row[q], r);
    }
    // point forces (PointForces, /root/reference/br2/custom_activation.py:46-81): an "extra joint" of type 1 that puts
    // F min(1, t / ramp_up_time) on one node
    for (int r = 0; r < p->n_rods; ++r) {
        const int32_t *ri = p->rod_i + r * BR2_ROD_NI;
        for (int f = ri[BR2_ROD_F_BEGIN]; f < ri[BR2_ROD_F_END]; ++f) {
            const int32_t *fi = p->forcing_i + (size_t)f * BR2_F_NI;
            const double *fd = p->forcing_d + (size_t)f * BR2_F_ND;
            if (fi[BR2_F_TYPE] != BR2_FORCE_POINT) continue;
            if (fi[BR2_F_NODE] < 0 || fi[BR2_F_NODE] > ri[BR2_ROD_N]) return *why = "point force on a node outside its rod", false;
            const int xj = plan->n_xj++;
            int jn;
            const int tn = thread_of(ri[BR2_ROD_SLOT0] + fi[BR2_F_NODE], &jn);
            const int32_t row[12] = {slot_rank, slot_row, jn, slot_rank, slot_row, jn, 0, 0, 0, 0, 1 /* type: point force */, 0};
            xt->xj_i.insert(xt->xj_i.end(), row, row + 12);
            const double drow[8] = {fd[0], fd[1], fd[2], fd[3], 0, 0, 0, 0};
            xt->xj_d.insert(xt->xj_d.end(), drow, drow + 8);
            xj_ranks.push_back(1 << slot_rank);
            int32_t &cnt = xt->xi[(size_t)tn * XN + BR2_TILE_XI_NCNT(C, jn)];
            if (cnt >= BR2_FAST_NODE_ITEMS) return *why = "a node receives too many extra forces", false;
            xt->xi[(size_t)tn * XN + BR2_TILE_XI_NITEM(C, jn, cnt)] = (xj << 1) | 0;
            ++cnt;
        }
    }
    for (size_t sl = 0; sl < S; ++sl) {
        if (p->overwrite_src[sl] < 0) continue;
        int j;
        const int t = thread_of((int)sl, &j);
        const int reader_rank = slot_rank;  // (xe_id may move slot_rank)
        const int xe = xe_id(p->overwrite_src[sl]);
        xt->xi[(size_t)t * XN + BR2_TILE_XI_OW(C, j)] = xe;
        consume(xe, reader_rank);
    }
    xe_mask.resize(plan->n_xe, 0);
    xt->xe_mask = xe_mask;
    for (int r = 0; r < 8; ++r) plan->xb_count[r] = 0;
    for (int xe = 0; xe < plan->n_xe; ++xe)
        for (int r = 0; r < plan->n_ctas; ++r)
            if ((xe_mask[xe] & (1 << r)) && xe_rank[xe] != r) ++plan->xb_count[r];  // (a CTA's own elements reach it by plain stores)
    // The extra joints are evaluated one COMPONENT per thread, by consecutive lanes of as few warps as possible: a warp
    // executes the joint code once whichever of its lanes take part.  One CTA: the last warps' threads.  Cluster: ALL 32 lanes
    // of the last warp, idle ones included (its rod-end threads put it on a lightly loaded sub-partition anyway, below), then
    // the lanes of warp 0 (ditto) -- not a warp of a fully loaded sub-partition, where 800 more cycles per step would make
    // it the slowest of the CTA.
    for (int rank = 0; rank < plan->n_ctas; ++rank) {
        const int ct = plan->cta_threads[rank], warps = (ct + 31) / 32;
        const bool cluster = plan->n_ctas > 1 && !(getenv("BR2_TILE_WARP_MAP") && atoi(getenv("BR2_TILE_WARP_MAP")) == 0);
        const int last = cluster ? 32 : ct - (warps - 1) * 32;  // lanes available in the last warp
        int used = 0;
        for (int xj = 0; xj < plan->n_xj; ++xj) {
            if (!(xj_ranks[xj] & (1 << rank))) continue;
            for (int comp = 0; comp < 3; ++comp) {
                const int k = used++;  // k-th (joint, component) of this CTA
                int t;
                if (k < last)
                    t = (warps - 1) * 32 + k;
                else if (cluster)
                    t = (k - last < 32 && warps > 1) ? k - last : (warps - 2 - (k - last - 32) / 32) * 32 + (k - last) % 32;  // warp 0, then from the back
                else
                    t = (warps - 2 - (k - last) / 32) * 32 + (k - last) % 32;
                if (t < 0 || t >= plan->xi_stride) return *why = "more extra joint components than threads", false;
                xt->xi[(size_t)(rank * plan->xi_stride + t) * XN + BR2_TILE_XI_JOINT] = (xj << 2) | comp;
            }
        }
    }
    // Which hardware warp runs which logical warp (cluster kernels; see the kernel): hardware warp w is scheduled by
    // sub-partition w % 4, so with W warps in the CTA some sub-partitions hold one warp less; the logical warps that carry
    // rod ends (publish / overwrite / gather) -- then those that evaluate joint components -- go there.
    for (int w = 0; w < 16; ++w) plan->warp_map[w] = w;
    if (plan->n_ctas > 1 && !(getenv("BR2_TILE_WARP_MAP") && atoi(getenv("BR2_TILE_WARP_MAP")) == 0)) {
        const int W = (threads + 31) / 32;
        std::vector<int> score(W, 0);
        for (int rank = 0; rank < plan->n_ctas; ++rank)
            for (int t = 0; t < 32 * ((plan->cta_threads[rank] + 31) / 32); ++t) {
                const int32_t *row = &xt->xi[(size_t)(rank * plan->xi_stride + t) * XN];
                int s = row[BR2_TILE_XI_JOINT] >= 0 ? 1 : 0;
                for (int j = 0; j < C; ++j)
                    if (row[BR2_TILE_XI_PUB(C, j)] >= 0 || row[BR2_TILE_XI_OW(C, j)] >= 0 || row[BR2_TILE_XI_NCNT(C, j)] > 0) s |= 2;
                score[t / 32] |= s;
            }
        std::vector<int> logical(W), hardware(W);
        for (int w = 0; w < W; ++w) logical[w] = hardware[w] = w;
        std::stable_sort(logical.begin(), logical.end(), [&](int a, int b) { return score[a] > score[b]; });
        auto load = [&](int w) { return (W - (w % 4) + 3) / 4; };  // warps on hardware warp w's sub-partition
        std::stable_sort(hardware.begin(), hardware.end(), [&](int a, int b) { return load(a) < load(b); });
        for (int i = 0; i < W; ++i) plan->warp_map[hardware[i]] = logical[i];
    }
    // pack: one int per slot + the joint component per thread.  A node that takes extra-joint forces gets a NODE ROW: the
    // threads that evaluate the joints scatter their (signed) half forces into the rows of the joint's four nodes, item by
    // item in the order the reference applies them, and the node's thread adds its row up -- one level of loads at fixed
    // offsets instead of item list -> joint id -> force.  xjt: per rank and joint, the four nodes' row positions (or -1 when
    // the node lives in another CTA of the cluster).
    if (plan->n_xe > 254) return *why = "too many elements in extra joints", false;
    int max_cnt = 1;
    for (int t = 0; t < xrows; ++t)
        for (int j = 0; j < C; ++j) max_cnt = std::max(max_cnt, (int)xt->xi[(size_t)t * XN + BR2_TILE_XI_NCNT(C, j)]);
    plan->nf_items = max_cnt;
    xt->xjt.assign((size_t)plan->n_ctas * plan->n_xj * 4, -1);
    xt->xpack.assign((size_t)xrows * (C + 1), 0);
    std::vector<int> rows_of_rank(plan->n_ctas, 0);
    plan->n_nr = 0;
    for (int t = 0; t < xrows; ++t) {
        const int32_t *row = &xt->xi[(size_t)t * XN];
        const int rank = t / plan->xi_stride;
        for (int j = 0; j < C; ++j) {
            const int cnt = row[BR2_TILE_XI_NCNT(C, j)];
            int nr = 0;
            if (cnt) {
                nr = rows_of_rank[rank]++;  // (row ids are per CTA: every CTA has its own array)
                if (nr > 4095) return *why = "too many nodes with extra joint forces", false;
                for (int q = 0; q < cnt; ++q) {
                    const int item = row[BR2_TILE_XI_NITEM(C, j, q)], xj = item >> 1;
                    int32_t *tg = &xt->xjt[((size_t)rank * plan->n_xj + xj) * 4];
                    // which of the joint's four nodes is this?  rod one's two nodes carry sign 0, rod two's sign 1; the first
                    // free place of that sign (a node appears once per joint)
                    const int m0 = (item & 1) ? 2 : 0;
                    const int m = tg[m0] < 0 ? m0 : m0 + 1;
                    tg[m] = (nr * max_cnt + q) * 4;
                }
            }
            xt->xpack[(size_t)t * (C + 1) + j] = (row[BR2_TILE_XI_PUB(C, j)] + 1) | ((row[BR2_TILE_XI_OW(C, j)] + 1) << 8) |
                                                 (cnt << 16) | (nr << 20);
        }
        xt->xpack[(size_t)t * (C + 1) + C] = row[BR2_TILE_XI_JOINT];
    }
    for (int r = 0; r < plan->n_ctas; ++r) plan->n_nr = std::max(plan->n_nr, rows_of_rank[r]);
    return true;
}

}  // namespace

extern "C" {

const char *br2_last_error(void) { return g_error.c_str(); }
int br2_abi_version(void) { return BR2_ABI_VERSION; }

int br2_create(const br2_program *p, int device, br2_handle *out) {
    if (!out) return fail(BR2_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = validate(p);
    if (rc != BR2_OK) return rc;
    CUDA_TRY(cudaSetDevice(device));
    br2_handle h = new br2_handle_s();
    h->device = device;
    DevTables &t = h->t;
    t.n_slots = p->n_slots;
    t.n_rods = p->n_rods;
    t.n_forcing = p->n_forcing;
    t.n_constraints = p->n_constraints;
    t.n_joints = p->n_joints;
    t.n_actions = p->n_actions;
    t.words = p->words_per_instance;
    const size_t S = (size_t)p->n_slots;
#define UP(field, count)                                                          \
    do {                                                                          \
        rc = upload(h, p->field, (size_t)(count), &t.field);                      \
        if (rc != BR2_OK) {                                                       \
            br2_destroy(h);                                                       \
            return rc;                                                            \
        }                                                                         \
    } while (0)
    UP(slot_rod, S);
    UP(rod_i, (size_t)p->n_rods * BR2_ROD_NI);
    UP(rod_d, (size_t)p->n_rods * BR2_ROD_ND);
    UP(slot_c, S * BR2_SC_COUNT);
    UP(forcing_i, (size_t)p->n_forcing * BR2_F_NI);
    UP(forcing_d, (size_t)p->n_forcing * BR2_F_ND);
    UP(cons_i, (size_t)p->n_constraints * BR2_C_NI);
    UP(cons_d, (size_t)p->n_constraints * BR2_C_ND);
    UP(joint_i, (size_t)p->n_joints * BR2_J_NI);
    UP(joint_d, (size_t)p->n_joints * BR2_J_ND);
    UP(node_ptr, S + 1);
    UP(node_items, (size_t)p->node_ptr[S]);
    UP(elem_ptr, S + 1);
    UP(elem_items, (size_t)p->elem_ptr[S]);
    UP(overwrite_src, S);
    h->fast_ok = p->fast_ok != 0 && p->fast_i && p->fast_c;
    h->fast_point = p->fast_point_forces != 0;
    h->fast_uniform = h->fast_ok && p->fast_uniform != 0;
    if (h->fast_ok) {
        UP(fast_i, S * BR2_FAST_NI);
        UP(fast_c, S * BR2_FC_COUNT);
        UP(act_i, (size_t)p->n_act);
        UP(act_d, (size_t)p->n_act * 3);
        h->n_act = p->n_act;
        t.act_inv_ramp = 1.0 / p->act_ramp_up_time;
        for (int i = 0; i < BR2_FC_COUNT; ++i) h->uni[i] = p->fast_c[(size_t)i * S];  // slot 0's column
        for (size_t sl = 0; sl < S; ++sl) {
            const int32_t *fi = p->fast_i + sl * BR2_FAST_NI;
            if (fi[BR2_FAST_EXTRA_JOINT] >= 0 || fi[BR2_FAST_NODE_ITEM0] >= 0 || fi[BR2_FAST_ELEM_ITEM0] >= 0 || p->overwrite_src[sl] >= 0)
                h->has_extra = true;
        }
    }
#undef UP
    {
        cudaError_t e = cudaMalloc((void **)&h->replays, 4 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemset(h->replays, 0, 4 * sizeof(unsigned long long));
        if (e != cudaSuccess) {
            br2_destroy(h);
            return fail(BR2_ECUDA, std::string("replay counter: ") + cudaGetErrorString(e));
        }
    }
    {
        TileExtraTables xt;
        h->tile_ok = h->fast_ok && make_tile_plan(p, BR2_TILE_C, &h->tile, &xt, &h->tile_why);
        if (h->tile_ok && h->tile.n_cls > 1) {  // one column of constants per class of rods, read by the kPerInst kernels
            std::vector<double> cls((size_t)h->tile.n_cls * BR2_FC_COUNT);
            for (int c = 0; c < h->tile.n_cls; ++c)
                for (int i = 0; i < BR2_FC_COUNT; ++i) cls[(size_t)c * BR2_FC_COUNT + i] = p->fast_c[(size_t)i * S + h->tile.cls_slot[c]];
            rc = upload(h, cls.data(), cls.size(), &h->cls_uni);
            if (rc != BR2_OK) {
                br2_destroy(h);
                return rc;
            }
        }
        if (h->tile_ok && h->tile.n_xj > 0) {
            rc = upload(h, xt.xpack.data(), xt.xpack.size(), &h->tile.xi);
            if (rc == BR2_OK) rc = upload(h, xt.xjt.data(), xt.xjt.size(), &h->tile.xjt);
            if (rc == BR2_OK) rc = upload(h, xt.xj_i.data(), xt.xj_i.size(), &h->tile.xj_i);
            if (rc == BR2_OK) rc = upload(h, xt.xj_d.data(), xt.xj_d.size(), &h->tile.xj_d);
            if (rc == BR2_OK) rc = upload(h, xt.xe_mask.data(), xt.xe_mask.size(), &h->tile.xe_mask);
            if (rc != BR2_OK) {
                br2_destroy(h);
                return rc;
            }
        }
    }
    h->auto_variant = h->tile_ok ? 3 : (h->fast_ok && !h->fast_point) ? (h->fast_uniform ? 2 : 1) : 0;
    rc = br2_select_kernel(h, -1);
    if (rc != BR2_OK) {
        br2_destroy(h);
        return rc;
    }
    *out = h;
    return BR2_OK;
}

int br2_select_kernel(br2_handle h, int32_t variant) {
    if (!h) return fail(BR2_EINVAL, "br2_select_kernel: NULL handle");
    if (variant < 0) variant = h->auto_variant;
    if (variant > 3) return fail(BR2_EINVAL, "br2_select_kernel: unknown variant");
    if (variant == 3 && !h->tile_ok)
        return fail(BR2_EINVAL, "br2_select_kernel: assembly does not qualify for the tile kernel: " + h->tile_why);
    if (variant >= 1 && !h->fast_ok) return fail(BR2_EINVAL, "br2_select_kernel: assembly does not qualify for the fast kernel");
    if (variant == 2 && !h->fast_uniform) return fail(BR2_EINVAL, "br2_select_kernel: assembly constants are not uniform");
    if ((variant == 1 || variant == 2) && h->fast_point)
        return fail(BR2_EINVAL, "br2_select_kernel: the one-slot-per-thread fast kernel does not serve point forces");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t S = (size_t)h->t.n_slots;
    KernelChoice pick{0, nullptr, nullptr};
    if (variant == 3) {
        const int need = h->tile.rod_t0[BR2_TILE_MAX_RODS];  // threads of the largest CTA
        h->exact = KernelChoice{0, nullptr, nullptr};
        if (h->tile.n_ctas > 1) {
            for (size_t i = 0; i < sizeof(kTileClusterKernels) / sizeof(kTileClusterKernels[0]); ++i)
                if (kTileClusterKernels[i].threads >= need) {
                    pick = kTileClusterKernels[i];
                    h->exact = kTileClusterExactKernels[i];
                    break;
                }
        } else if (h->tile.n_xj > 0) {
            for (const KernelChoice &kc : kTileExtraKernels)
                if (kc.threads >= need) {
                    pick = kc;
                    break;
                }
        } else {
            for (const KernelChoice &kc : kTileKernels)
                if (kc.threads >= need) {
                    pick = kc;
                    break;
                }
        }
    } else {
        for (const KernelChoice &kc : kKernels[variant])
            if (kc.threads >= h->t.n_slots) {
                pick = kc;
                break;
            }
    }
    if (!pick.fn) return fail(BR2_ELIMIT, "no kernel variant holds this many slots");
    size_t smem;
    if (variant == 0)
        smem = sizeof(double) * (30 * S + 9 * (size_t)h->t.n_joints + 32);
    else if (variant == 3)  // rows, then the extra-joint area (layout: TileSmem)
        smem = sizeof(double) * ((size_t)br2::TileSmem<BR2_TILE_C, 0>::ROWS_END * (size_t)(pick.threads + 1) +
                                 2 * br2::TileSmem<BR2_TILE_C, 0>::EXW * (size_t)h->tile.n_xe +
                                 4 * (size_t)h->tile.n_nr * (size_t)h->tile.nf_items + (8 + 4 + 2) * (size_t)h->tile.n_xj +
                                 ((size_t)h->tile.n_xe + 1) / 2);
    else
    {
        // fixed part: 45 T doubles + two [3][T+8] accumulators + O1 (T ints); tail (omega, extra-joint rows,
        // 14 T ints of bookkeeping) only for assemblies with extra joints / director overwrites
        const size_t Tn = (size_t)pick.threads;
        smem = sizeof(double) * (45 * Tn + 6 * (Tn + 8) + Tn / 2);
        if (h->has_extra) smem += sizeof(double) * (3 * Tn + 9 * (size_t)BR2_FAST_MAX_EXTRA + 7 * Tn);
    }
    if (smem > 227 * 1024)
        return fail(BR2_ELIMIT, "assembly needs " + std::to_string(smem) + " B of shared memory per CTA (max 232448)");
    CUDA_TRY(cudaFuncSetAttribute((const void *)pick.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (pick.fn_pi)
        CUDA_TRY(cudaFuncSetAttribute((const void *)pick.fn_pi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (variant == 3 && h->exact.fn) {
        CUDA_TRY(cudaFuncSetAttribute((const void *)h->exact.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (h->exact.fn_pi)
            CUDA_TRY(cudaFuncSetAttribute((const void *)h->exact.fn_pi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (variant != 3) h->exact = KernelChoice{0, nullptr, nullptr};
    h->kernel = pick;
    h->smem_bytes = smem;
    h->variant = variant;
    if (variant == 3 && h->tile.n_ctas > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)h->tile.n_ctas * 64);
        cfg.blockDim = dim3((unsigned)pick.threads);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)h->tile.n_ctas;
        attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int clusters = 0;
        CUDA_TRY(cudaOccupancyMaxActiveClusters(&clusters, (const void *)pick.fn, &cfg));
        if (clusters < 1) return fail(BR2_ELIMIT, "the tile kernel's cluster does not fit on the device");
        h->tile_capacity = clusters;  // in clusters
    } else if (variant == 3) {
        int blocks = 0, sms = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, (const void *)pick.fn, pick.threads, smem));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
        if (blocks < 1) return fail(BR2_ELIMIT, "the tile kernel does not fit on an SM");
        h->tile_capacity = blocks * sms;
    }
    if (variant != 0) {
        // exact-path replay kernel; an assembly that only fits a cluster has none (one CTA cannot hold it): instances
        // that leave the optimistic kernel's range then stay flagged and frozen -- br2_unresolved_count reports them
        h->replay = KernelChoice{0, nullptr, nullptr};
        for (const KernelChoice &kc : kKernels[0])
            if (kc.threads >= h->t.n_slots) {
                h->replay = kc;
                break;
            }
        h->replay_smem = sizeof(double) * (30 * S + 9 * (size_t)h->t.n_joints + 32);
        if (h->replay_smem > 227 * 1024) h->replay = KernelChoice{0, nullptr, nullptr};
        if (!h->replay.fn && variant != 3) return fail(BR2_ELIMIT, "assembly too large for the replay kernel");
        if (h->replay.fn)
            CUDA_TRY(cudaFuncSetAttribute((const void *)h->replay.fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)h->replay_smem));
    }
    return BR2_OK;
}

int br2_destroy(br2_handle h) {
    if (!h) return BR2_OK;
    cudaSetDevice(h->device);
    for (void *d : h->owned) cudaFree(d);
    if (h->status) cudaFree(h->status);
    if (h->work) cudaFree(h->work);
    if (h->replays) cudaFree(h->replays);
    if (h->scratch_state) cudaFree(h->scratch_state);
    if (h->scratch_pressures) cudaFree(h->scratch_pressures);
    delete h;
    return BR2_OK;
}

int64_t br2_words_per_instance(br2_handle h) { return h ? h->t.words : 0; }

int br2_bind_state(br2_handle h, double *dev_state, int64_t batch) {
    if (!h || !dev_state || batch <= 0) return fail(BR2_EINVAL, "br2_bind_state: bad argument");
    h->state = dev_state;
    h->batch = batch;
    if (h->status_batch < batch) {
        CUDA_TRY(cudaSetDevice(h->device));
        if (h->status) cudaFree(h->status);
        h->status = nullptr;
        CUDA_TRY(cudaMalloc((void **)&h->status, sizeof(int) * (size_t)batch));
        if (h->work) cudaFree(h->work);
        h->work = nullptr;
        CUDA_TRY(cudaMalloc((void **)&h->work, 2 * sizeof(unsigned long long) + sizeof(int) * (size_t)batch));
        h->status_batch = batch;
    }
    CUDA_TRY(cudaMemset(h->status, 0, sizeof(int) * (size_t)batch));
    return BR2_OK;
}

int br2_set_actuation(br2_handle h, const double *dev_pressures) {
    if (!h) return fail(BR2_EINVAL, "br2_set_actuation: NULL handle");
    if (!dev_pressures && h->t.n_actions > 0) return fail(BR2_EINVAL, "br2_set_actuation: NULL pressures");
    h->pressures = dev_pressures;
    return BR2_OK;
}

int br2_set_instance_params(br2_handle h, const double *dev_uni, const double *dev_act_d, const double *dev_joint_scale) {
    if (!h) return fail(BR2_EINVAL, "br2_set_instance_params: NULL handle");
    if (!dev_uni) {  // back to the constants every instance shares
        h->inst_uni = h->inst_act_d = h->inst_joint_scale = nullptr;
        return BR2_OK;
    }
    if (h->variant != 3 || !h->kernel.fn_pi)
        return fail(BR2_ESTATE, "br2_set_instance_params: only the tile kernel reads per-instance constants (" +
                                    (h->tile_ok ? std::string("select variant 3") : h->tile_why) + ")");
    if ((h->n_act > 0 && !dev_act_d) || !dev_joint_scale) return fail(BR2_EINVAL, "br2_set_instance_params: NULL table");
    if (h->tile.n_cls > 1)
        return fail(BR2_ESTATE, "br2_set_instance_params: the assembly's rods differ in their constants; per-instance tables "
                                "are defined for assemblies with one set of element constants");
    h->inst_uni = dev_uni;
    h->inst_act_d = dev_act_d;
    h->inst_joint_scale = dev_joint_scale;
    return BR2_OK;
}

int br2_count_steps(double t0, double duration, double dt, int64_t *n_steps, double *t_end) {
    if (!(dt > 0.0)) return fail(BR2_EINVAL, "br2_count_steps: dt must be positive");
    volatile double time = t0;
    const double end = t0 + duration, half = 0.5 * dt;
    int64_t n = 0;
    while (time < end) {
        time += half;
        time += half;
        ++n;
    }
    if (n_steps) *n_steps = n;
    if (t_end) *t_end = time;
    return BR2_OK;
}

int br2_capture_plan(double t0, int64_t n_steps, double dt, const int64_t *every, int32_t n_every, int64_t *ordinals,
                     double *times, int64_t capacity, int64_t *count) {
    if (!(dt > 0.0) || n_steps < 0 || !count || (n_every > 0 && !every))
        return fail(BR2_EINVAL, "br2_capture_plan: bad argument");
    volatile double time = t0;
    const double half = 0.5 * dt;
    int64_t found = 0;
    for (int64_t ordinal = 1; ordinal <= n_steps; ++ordinal) {
        time += half;
        time += half;
        const double q = time / dt;
        const long long step = llrint(q);  // python's round(): to nearest, ties to even (default rounding mode)
        bool due = false;
        for (int32_t i = 0; i < n_every && !due; ++i) due = every[i] <= 1 || step % every[i] == 0;
        if (!due) continue;
        if (found < capacity) {
            if (ordinals) ordinals[found] = ordinal;
            if (times) times[found] = time;
        }
        ++found;
    }
    *count = found;
    return BR2_OK;
}

int br2_step(br2_handle h, int64_t n_steps, double t0, double dt, void *stream, double *t_end) {
    if (!h) return fail(BR2_EINVAL, "br2_step: NULL handle");
    if (!h->state) return fail(BR2_ESTATE, "br2_step: no state bound (call br2_bind_state)");
    if (!h->pressures && h->t.n_actions > 0) return fail(BR2_ESTATE, "br2_step: no actuation set (call br2_set_actuation)");
    if (n_steps < 0 || !(dt > 0.0)) return fail(BR2_EINVAL, "br2_step: bad n_steps / dt");
    if (n_steps > (1LL << 30)) return fail(BR2_EINVAL, "br2_step: more than 2^30 steps in one launch; split the call");
    volatile double time = t0;
    const double half = 0.5 * dt;
    for (int64_t i = 0; i < n_steps; ++i) {
        time += half;
        time += half;
    }
    if (t_end) *t_end = time;
    if (n_steps == 0) return BR2_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    StepParams sp;
    sp.t = h->t;
    sp.state = h->state;
    sp.pressures = h->pressures;
    sp.batch = h->batch;
    sp.n_steps = n_steps;
    sp.t0 = t0;
    sp.dt = dt;
    memcpy(sp.uni, h->uni, sizeof(sp.uni));
    {
        const double *u = h->uni;
        for (int i = 0; i < 3; ++i) {
            sp.kd[i] = dt * u[BR2_FC_CT] * u[BR2_FC_GRAVITY + i];
            sp.kd[8 + i] = dt * u[BR2_FC_JINV + i];
        }
        sp.kd[3] = 0.5 * u[BR2_FC_INV_REST_VLEN];
        sp.kd[4] = -0.5 * u[BR2_FC_INV_REST_VLEN];
        sp.kd[5] = -0.25 * u[BR2_FC_JNU];
        sp.kd[6] = 0.5 * u[BR2_FC_JK];
        sp.kd[7] = -0.5 * u[BR2_FC_JKREP];
        sp.kd[11] = 0.5 * dt;
        sp.kd[12] = u[BR2_FC_SHEAR] * u[BR2_FC_INV_REST_LEN];
        sp.kd[13] = u[BR2_FC_SHEAR + 1] * u[BR2_FC_INV_REST_LEN];
    }
    sp.tile = h->tile;
    sp.inst_uni = h->inst_uni;
    sp.inst_act_d = h->inst_act_d;
    sp.inst_joint_scale = h->inst_joint_scale;
    sp.inst_stride = 1;
    sp.n_act = h->n_act;
    const bool per_instance = h->inst_uni != nullptr;
    // rods that differ in their constants: the same kernels, every instance reading the one table of classes
    const bool per_class = !per_instance && h->variant == 3 && h->tile.n_cls > 1;
    if (per_class) {
        sp.inst_uni = h->cls_uni;
        sp.inst_stride = 0;
    }
    if (per_instance && (h->variant != 3 || !h->kernel.fn_pi))
        return fail(BR2_ESTATE, "br2_step: per-instance material parameters need the tile kernel (br2_select_kernel 3)");
    const void *step_fn = (per_instance || per_class) ? (const void *)h->kernel.fn_pi : (const void *)h->kernel.fn;
    const void *exact_fn = (per_instance || per_class) ? (const void *)h->exact.fn_pi : (const void *)h->exact.fn;
    sp.status = h->status;
    sp.replays = h->replays;
    sp.only_flagged = 0;
    sp.fail_counter = h->work;          // work buffer: [flagged instances][queue head][progress per instance]
    sp.work_counter = h->work + 1;
    sp.progress = (int *)(h->work + 2);
    // chunking: the tile kernel's persistent CTAs pull (chunk, instance) items; everything else runs one chunk
    sp.n_chunks = 1;
    if (h->variant == 3) {
        long long c = n_steps / BR2_TILE_MIN_CHUNK_STEPS;
        if (const char *env = getenv("BR2_TILE_CHUNKS")) {  // tuning knob: upper bound on the chunks per launch
            const long long cap = atoll(env);
            if (cap >= 1 && c > cap) c = cap;
        }
        sp.n_chunks = (int)(c < 1 ? 1 : (c > BR2_MAX_CHUNKS ? BR2_MAX_CHUNKS : c));
    }
    sp.chunk_steps = (int)((n_steps + sp.n_chunks - 1) / sp.n_chunks);
    sp.n_chunks = (int)((n_steps + sp.chunk_steps - 1) / sp.chunk_steps);
    // queue order: groups of about two waves of resident CTAs, chunk-major inside a group (the state a chunk writes back is
    // read again while it is still in the L2); BR2_TILE_QUEUE_GROUP overrides (0 = the whole batch is one group)
    sp.queue_group = h->batch;
    if (h->variant == 3 && h->tile_capacity > 0 && 2LL * h->tile_capacity < h->batch) {
        // equal groups of at least two waves: a short last group would hand out an instance's next chunk before the previous
        // one is done and run at partial occupancy (measured: -4 % with 1776 + 1776 + 544 instead of 2048 + 2048)
        const long long groups = h->batch / (2LL * h->tile_capacity);
        sp.queue_group = (h->batch + groups - 1) / groups;
    }
    if (const char *env = getenv("BR2_TILE_QUEUE_GROUP")) {
        const long long g = atoll(env);
        sp.queue_group = (g >= 1 && g < h->batch) ? g : h->batch;
    }
    if (sp.queue_group < 1) sp.queue_group = 1;
    {
        volatile double tc = t0;
        for (int c = 0; c < sp.n_chunks; ++c) {
            sp.chunk_t0[c] = tc;
            for (int i = 0; i < sp.chunk_steps; ++i) {
                tc += half;
                tc += half;
            }
        }
    }
    void *args[] = {&sp};
    unsigned grid = (unsigned)h->batch;
    if (h->variant == 3) {
        CUDA_TRY(cudaMemsetAsync(h->work, 0, 2 * sizeof(unsigned long long) + sizeof(int) * (size_t)h->batch, (cudaStream_t)stream));
        const long long items = h->batch * (long long)sp.n_chunks;
        grid = (unsigned)(items < h->tile_capacity ? items : h->tile_capacity);
    }
    if (h->variant == 3 && h->tile.n_ctas > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid * (unsigned)h->tile.n_ctas);  // `grid` clusters
        cfg.blockDim = dim3((unsigned)h->kernel.threads);
        cfg.dynamicSmemBytes = h->smem_bytes;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)h->tile.n_ctas;
        attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelExC(&cfg, step_fn, args));
        if (exact_fn) {
            // exact-path pass of the same kernel (IEEE division / libm, no validity ranges) for the instances the
            // optimistic pass flagged; every CTA returns at once when there are none.  The queue restarts, the count
            // of flagged instances stays.
            CUDA_TRY(cudaMemsetAsync(h->work + 1, 0, sizeof(unsigned long long) + sizeof(int) * (size_t)h->batch, (cudaStream_t)stream));
            CUDA_TRY(cudaLaunchKernelExC(&cfg, exact_fn, args));
        }
    } else {
        CUDA_TRY(cudaLaunchKernel(step_fn, dim3(grid), dim3((unsigned)h->kernel.threads), args,
                                  h->smem_bytes, (cudaStream_t)stream));
    }
    // (per-instance material parameters: the generic kernel reads the shared tables, so it cannot redo such an instance;
    // one that leaves the optimistic range stays flagged and br2_unresolved_count reports it)
    if (h->variant != 0 && h->replay.fn && !per_instance) {
        // exact-path replay, stream-ordered behind the fast kernel: CTAs of unflagged instances exit at once
        sp.only_flagged = 1;
        CUDA_TRY(cudaLaunchKernel((const void *)h->replay.fn, dim3((unsigned)h->batch), dim3((unsigned)h->replay.threads),
                                  args, h->replay_smem, (cudaStream_t)stream));
    }
    return BR2_OK;
}

int br2_step_host(br2_handle h, double *host_state, const double *host_pressures, int64_t batch, int64_t n_steps,
                  double t0, double dt, double *t_end) {
    if (!h || !host_state || batch <= 0) return fail(BR2_EINVAL, "br2_step_host: bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->scratch_batch < batch) {
        if (h->scratch_state) cudaFree(h->scratch_state);
        if (h->scratch_pressures) cudaFree(h->scratch_pressures);
        h->scratch_state = nullptr;
        h->scratch_pressures = nullptr;
        CUDA_TRY(cudaMalloc((void **)&h->scratch_state, sizeof(double) * (size_t)batch * (size_t)h->t.words));
        CUDA_TRY(cudaMalloc((void **)&h->scratch_pressures, sizeof(double) * (size_t)batch * (size_t)(h->t.n_actions + 1)));
        h->scratch_batch = batch;
    }
    const size_t state_bytes = sizeof(double) * (size_t)batch * (size_t)h->t.words;
    CUDA_TRY(cudaMemcpy(h->scratch_state, host_state, state_bytes, cudaMemcpyHostToDevice));
    if (h->t.n_actions > 0) {
        if (!host_pressures) return fail(BR2_EINVAL, "br2_step_host: NULL pressures");
        CUDA_TRY(cudaMemcpy(h->scratch_pressures, host_pressures, sizeof(double) * (size_t)batch * (size_t)h->t.n_actions,
                            cudaMemcpyHostToDevice));
    }
    if (h->status_batch < batch) {
        if (h->status) cudaFree(h->status);
        h->status = nullptr;
        CUDA_TRY(cudaMalloc((void **)&h->status, sizeof(int) * (size_t)batch));
        if (h->work) cudaFree(h->work);
        h->work = nullptr;
        CUDA_TRY(cudaMalloc((void **)&h->work, 2 * sizeof(unsigned long long) + sizeof(int) * (size_t)batch));
        h->status_batch = batch;
        CUDA_TRY(cudaMemset(h->status, 0, sizeof(int) * (size_t)batch));
    }
    double *saved_state = h->state;
    const double *saved_p = h->pressures;
    long long saved_batch = h->batch;
    h->state = h->scratch_state;
    h->pressures = h->scratch_pressures;
    h->batch = batch;
    int rc = br2_step(h, n_steps, t0, dt, nullptr, t_end);
    h->state = saved_state;
    h->pressures = saved_p;
    h->batch = saved_batch;
    if (rc != BR2_OK) return rc;
    CUDA_TRY(cudaMemcpy(host_state, h->scratch_state, state_bytes, cudaMemcpyDeviceToHost));
    return BR2_OK;
}

int br2_kernel_info(br2_handle h, int32_t *regs, int32_t *smem_bytes, int32_t *threads, int32_t *blocks_per_sm,
                    int32_t *sm_count, int32_t *variant) {
    if (!h) return fail(BR2_EINVAL, "br2_kernel_info: NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaFuncAttributes attr;
    CUDA_TRY(cudaFuncGetAttributes(&attr, (const void *)h->kernel.fn));
    int blocks = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, (const void *)h->kernel.fn, h->kernel.threads,
                                                           h->smem_bytes));
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    if (h->variant == 3 && h->tile.n_ctas <= 1 && sms > 0) blocks = h->tile_capacity / sms;  // as launched (see br2_select_kernel)
    if (regs) *regs = attr.numRegs;
    if (smem_bytes) *smem_bytes = (int32_t)h->smem_bytes;
    if (threads) *threads = h->kernel.threads;
    if (blocks_per_sm) *blocks_per_sm = blocks;
    if (sm_count) *sm_count = sms;
    if (variant) *variant = h->variant;
    return BR2_OK;
}

int br2_replay_count(br2_handle h, int64_t *count) {
    if (!h || !count) return fail(BR2_EINVAL, "br2_replay_count: bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    unsigned long long v = 0;
    CUDA_TRY(cudaMemcpy(&v, h->replays, sizeof(v), cudaMemcpyDeviceToHost));  // synchronises with prior work
    *count = (int64_t)v;
    return BR2_OK;
}

int br2_describe_plan(const br2_program *p, char *buf, int64_t buf_len) {
    if (!buf || buf_len <= 0) return fail(BR2_EINVAL, "br2_describe_plan: no buffer");
    int rc = validate(p);
    if (rc != BR2_OK) return rc;
    std::string text;
    br2::TilePlan plan{};
    TileExtraTables xt;
    std::string why;
    if (p->fast_ok && p->fast_i && p->fast_c && make_tile_plan(p, BR2_TILE_C, &plan, &xt, &why)) {
        text = "tile: " + std::to_string(plan.n_ctas) + " CTA(s) per instance";
        for (int c = 0; c < plan.n_ctas; ++c) text += (c ? ", " : " of ") + std::to_string(plan.cta_threads[c]);
        text += " threads (" + std::to_string(BR2_TILE_C) + " slots per thread), " + std::to_string(plan.n_xj) +
                " tip-to-base joint(s) / point force(s)";
        if (plan.n_cls > 1) text += ", " + std::to_string(plan.n_cls) + " classes of rod constants (shared-memory copy)";
    } else if (p->fast_ok && p->fast_i && p->fast_c) {
        text = p->fast_point_forces ? "generic: table-interpreting kernel (tile kernel refused: " + why + ")"
                                    : std::string(p->fast_uniform ? "fast-uniform" : "fast") + ": one slot per thread (tile kernel refused: " + why + ")";
    } else {
        text = "generic: table-interpreting kernel (the assembly is outside the BR2 fast-path pattern)";
    }
    if (p->n_slots > 1024 && text.compare(0, 5, "tile:") != 0)
        return fail(BR2_ELIMIT, "assembly has " + std::to_string(p->n_slots) + " node slots and does not qualify for the cluster kernel: " + why);
    snprintf(buf, (size_t)buf_len, "%s", text.c_str());
    return BR2_OK;
}

int br2_unresolved_count(br2_handle h, int64_t *count) {
    if (!h || !count) return fail(BR2_EINVAL, "br2_unresolved_count: bad argument");
    *count = 0;
    if (!h->status || h->batch <= 0) return BR2_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    std::vector<int> host((size_t)h->batch);
    CUDA_TRY(cudaMemcpy(host.data(), h->status, sizeof(int) * (size_t)h->batch, cudaMemcpyDeviceToHost));
    for (int v : host) *count += (v != 0);
    return BR2_OK;
}

int br2_replay_reasons(br2_handle h, int64_t *rotation, int64_t *curvature, int64_t *strain) {
    if (!h) return fail(BR2_EINVAL, "br2_replay_reasons: NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    unsigned long long v[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpy(v, h->replays, sizeof(v), cudaMemcpyDeviceToHost));
    if (rotation) *rotation = (int64_t)v[1];
    if (curvature) *curvature = (int64_t)v[2];
    if (strain) *strain = (int64_t)v[3];
    return BR2_OK;
}

int br2_metrics(br2_handle h, const double *dev_previous, double *dev_out, void *stream) {
    if (!h || !dev_out) return fail(BR2_EINVAL, "br2_metrics: bad argument");
    if (!h->state) return fail(BR2_ESTATE, "br2_metrics: no state bound (call br2_bind_state)");
    CUDA_TRY(cudaSetDevice(h->device));
    br2::MetricsParams mp;
    mp.t = h->t;
    mp.state = h->state;
    if (h->t.n_rods > BR2_METRICS_MAX_RODS) return fail(BR2_ELIMIT, "br2_metrics: more rods than the metrics kernel's table holds");
    mp.previous = dev_previous;
    mp.out = dev_out;
    mp.batch = h->batch;
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    const long long want = h->batch < 16LL * sms ? h->batch : 16LL * sms;  // grid-stride over the instances
    br2::br2_metrics_kernel<<<(unsigned)want, br2::kMetricsThreads, 0, (cudaStream_t)stream>>>(mp);
    CUDA_TRY(cudaGetLastError());
    return BR2_OK;
}

int br2_selftest_math(int32_t which, const double *dev_in, double *dev_out, int64_t n) {
    if (!dev_in || !dev_out || n <= 0) return fail(BR2_EINVAL, "br2_selftest_math: bad argument");
    br2::selftest_math_kernel<<<(unsigned)((n + 255) / 256), 256>>>(which, dev_in, dev_out, (long long)n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return BR2_OK;
}

int br2_fp64_peak_tflops(int device, int32_t repeats, double *tflops) {
    if (!tflops) return fail(BR2_EINVAL, "br2_fp64_peak_tflops: NULL out");
    CUDA_TRY(cudaSetDevice(device));
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sms * 8, threads = 256;
    const long long iters = 20000;
    double *out = nullptr;
    CUDA_TRY(cudaMalloc((void **)&out, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int r = 0; r < repeats + 1; ++r) {
        CUDA_TRY(cudaEventRecord(e0));
        dfma_probe_kernel<<<blocks, threads>>>(out, iters, 1.0 + r);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (r > 0 && tf > best) best = tf;  // r == 0 is the warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return BR2_OK;
}

}  // extern "C"
