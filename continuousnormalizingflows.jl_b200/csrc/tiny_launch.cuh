// tiny_launch.cuh -- host launchers for one instantiation of the tiny family and
// the macro that registers it with the dispatcher in api.cu.
#pragma once
#include <algorithm>

#include "family.h"
#include "tiny.cuh"
#include "tiny_sp.cuh"
#include "tiny_vcabm.cuh"

namespace icnf {
namespace tiny {

template <class N>
struct Launch {
    static constexpr size_t smem_rhs = 0;
    static constexpr size_t smem_solve = sizeof(float) * (6 * N::D * NT);
    static constexpr size_t smem_bwd = sizeof(float) * (6 * N::D * NT > (NT / 32) * N::NP ? 6 * N::D * NT : (NT / 32) * N::NP);
    static_assert(2 * sizeof(WBlock<N>) + sizeof(SolveArgs) + sizeof(SPPlan) + 16 <= 32764, "weights must fit the kernel parameter space");

    // theta (host, native ComponentArray order) -> padded parameter block
    static void pack(const float* theta, WBlock<N>& w) {
        for (int i = 0; i < N::WPAD; ++i) w.v[i] = 0.0f;
        for (int l = 0; l < N::NL; ++l) {
            const int nin = N::n(l), nout = N::n(l + 1);
            for (int k = 0; k < nin; ++k)
                for (int j = 0; j < nout; ++j) w.v[N::woff(l) + k * N::ld(l) + j] = theta[N::toff(l) + k * nout + j];
            for (int j = 0; j < nout; ++j) w.v[N::boff(l) + j] = theta[N::toff(l) + nin * nout + j];
            // layout B: rows of W_l over the inputs that carry a z-derivative
            for (int j = 0; j < nout; ++j)
                for (int k = 0; k < N::kz(l); ++k) w.v[N::wtoff(l) + j * N::ldt(l) + k] = theta[N::toff(l) + k * nout + j];
        }
        // exact-trace block
        auto W = [&](int l, int j, int k) { return theta[N::toff(l) + k * N::n(l + 1) + j]; };
        if (N::NL == 3) {
            for (int k = 0; k < N::n(1); ++k)
                for (int j = 0; j < N::n(2); ++j) {
                    float s = 0.f;
                    for (int i = 0; i < N::D; ++i) s += W(0, k, i) * W(2, i, j);
                    w.v[N::troff + k * N::ld(1) + j] = W(1, j, k) * s;
                }
        } else if (N::NL == 2) {
            for (int k = 0; k < N::n(1); ++k) {
                float s = 0.f;
                for (int i = 0; i < N::D; ++i) s += W(0, k, i) * W(1, i, k);
                w.v[N::troff + k] = s;
            }
        } else if (N::NL == 1) {
            float s = 0.f;
            for (int i = 0; i < N::D; ++i) s += W(0, i, i);
            w.v[N::troff] = s;
        }
    }

    template <class K>
    static cudaError_t prep(K kernel, size_t smem) {
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    template <class K>
    static int occupancy(K kernel, size_t smem, int threads = NT) {
        int nb = 0;
        if (prep(kernel, smem) != cudaSuccess) return 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess) return 0;
        return nb;
    }
    static int grid_for(long long B, int per_sm, int sm_count) {
        long long need = (B + NT - 1) / NT;
        long long cap = (long long)std::max(per_sm, 1) * sm_count;
        return (int)std::max(1LL, std::min(need, cap));
    }

    static cudaError_t rhs(void*, const float* theta, const RhsArgs& a, bool exact, int sm_count, cudaStream_t st) {
        auto k = exact ? rhs_kernel<N, true> : rhs_kernel<N, false>;
        int grid = grid_for(a.B, occupancy(k, smem_rhs), sm_count);
        WBlock<N> w;
        pack(theta, w);
        k<<<grid, NT, smem_rhs, st>>>(w, a);
        return cudaGetLastError();
    }
    static cudaError_t solve_fixed(void*, const float* theta, const SolveArgs& a, int nvars, bool exact, int sm_count, cudaStream_t st) {
        auto k = exact ? solve_fixed_kernel<N, true> : solve_fixed_kernel<N, false>;
        int grid = grid_for(a.B, occupancy(k, smem_solve), sm_count);
        WBlock<N> w;
        pack(theta, w);
        k<<<grid, NT, smem_solve, st>>>(w, a, nvars);
        return cudaGetLastError();
    }
    static int adaptive_max_grid(bool exact, int sm_count) {
        auto k = exact ? solve_adaptive_kernel<N, true> : solve_adaptive_kernel<N, false>;
        return occupancy(k, sizeof(float) * 6 * N::D * 64, 64) * sm_count;
    }
    // CTA size: spread the batch over all SMs, one CTA each once it is large enough
    static int adaptive_block(long long B, int sm_count) {
        long long per_sm = (B + sm_count - 1) / sm_count;
        long long bs = ((per_sm + 31) / 32) * 32;
        return (int)std::max(64LL, std::min<long long>(NTA, bs));
    }
    static cudaError_t solve_adaptive(void*, const float* theta, const SolveArgs& a, int nvars, bool exact, int sm_count, cudaStream_t st) {
        auto k = exact ? solve_adaptive_kernel<N, true> : solve_adaptive_kernel<N, false>;
        const int bs = adaptive_block(a.B, sm_count);
        const size_t smem = sizeof(float) * 6 * N::D * bs;
        const int per_sm = occupancy(k, smem, bs);
        if (per_sm <= 0) return cudaErrorLaunchOutOfResources;
        const long long need = (a.B + bs - 1) / bs;
        const int grid = (int)std::max(1LL, std::min<long long>(need, (long long)per_sm * sm_count));
        SolveArgs aa = a;
        int nv = nvars;
        WBlock<N> w;
        pack(theta, w);
        void* args[] = {(void*)&w, (void*)&aa, (void*)&nv};
        return cudaLaunchCooperativeKernel((const void*)k, dim3(grid), dim3(bs), args, smem, st);
    }
    static cudaError_t solve_vcabm(void*, const float* theta, const SolveArgs& a, int nvars, bool exact, int sm_count, cudaStream_t st) {
        auto k = exact ? solve_vcabm_kernel<N, true> : solve_vcabm_kernel<N, false>;
        const int bs = adaptive_block(a.B, sm_count);
        const int per_sm = occupancy(k, 0, bs);
        if (per_sm <= 0) return cudaErrorLaunchOutOfResources;
        const long long need = (a.B + bs - 1) / bs;
        const int grid = (int)std::max(1LL, std::min<long long>(need, (long long)per_sm * sm_count));
        SolveArgs aa = a;
        int nv = nvars;
        WBlock<N> w;
        pack(theta, w);
        void* args[] = {(void*)&w, (void*)&aa, (void*)&nv};
        return cudaLaunchCooperativeKernel((const void*)k, dim3(grid), dim3(bs), args, 0, st);
    }
    // networks with at least one hidden layer use the sample-parallel backward (tiny_sp.cuh); the thread-per-sample
    // backward_kernel of tiny.cuh serves networks without a hidden layer (a linear field)
    static constexpr bool USE_SP = (N::NL >= 2);
    // sample-parallel backward: CTA size (= samples per tile) and grid for a batch.
    // The CTA size is the smallest multiple of 32 that lets the batch finish in the fewest
    // rounds of ICNF_SP_MINB CTAs per SM, capped by ICNF_SP_MAXT and by shared memory.
    template <bool EXACT>
    static void sp_plan(long long B, int sm_count, int& ns, int& grid) {
        using C = SPCfg<N, EXACT>;
        const int cap = C::PITCH;
        const int lo = ((C::NBLK + 31) / 32) * 32;
        const long long slots = (long long)ICNF_SP_MINB * sm_count;
        const long long rounds = std::max(1LL, (B + slots * cap - 1) / (slots * cap));
        long long per = (B + slots * rounds - 1) / (slots * rounds);
        ns = (int)std::min<long long>(cap, std::max<long long>(lo, ((per + 31) / 32) * 32));
        grid = (int)std::max(1LL, std::min(slots, (B + ns - 1) / ns));
    }
    // SM count of the CURRENT device (a process may hold handles on several devices)
    static int current_sm_count() {
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n;
    }
    static int backward_grid(bool exact, int sm_count, long long B) {
        if constexpr (USE_SP) {
            int ns, grid;
            if (exact) sp_plan<true>(B, sm_count, ns, grid);
            else sp_plan<false>(B, sm_count, ns, grid);
            return grid;
        } else {
            auto k = exact ? backward_kernel<N, true> : backward_kernel<N, false>;
            return grid_for(B, occupancy(k, smem_bwd), sm_count);
        }
    }
    // hand the CTA's threads to the dW blocks.  A warp must not mix blocks (their control flow and
    // shared-memory rows differ: measured 1.6x slower when it does), so with at least one warp per
    // block the unit is a warp and spare warps go to the block whose threads have the most work
    // (sample pairs per thread x cost per pair); smaller CTAs split their threads evenly.
    template <bool EXACT>
    static SPPlan sp_threads(int ns) {
        using C = SPCfg<N, EXACT>;
        int ng[C::NBLK];
        const int nduo = ns / 2, nwarp = ns / 32;
        // shared-memory bound of the final reduction: groups * NP floats must fit the record area
        const int gmax = std::max(1, (int)((size_t)C::NPR * C::PITCH * 8 / ((size_t)N::NP * 4)));
        if (nwarp >= C::NBLK) {
            for (int b = 0; b < C::NBLK; ++b) ng[b] = 32;
            auto load = [&](int b) { return (long long)((nduo + ng[b] - 1) / ng[b]) * C::cost(b); };
            for (int left = nwarp - C::NBLK; left > 0; --left) {
                int best = -1;
                for (int b = 0; b < C::NBLK; ++b)
                    if (ng[b] + 32 <= std::min(nduo, gmax) && (best < 0 || load(b) > load(best))) best = b;
                if (best < 0) break;
                ng[best] += 32;
            }
        } else {
            for (int b = 0; b < C::NBLK; ++b) ng[b] = std::max(1, std::min(ns / C::NBLK, gmax));
        }
        SPPlan p{};
        int o = 0, mg = 1;
        for (int b = 0; b < C::NBLK; ++b) { p.first[b] = (unsigned short)o; o += ng[b]; mg = std::max(mg, ng[b]); }
        for (int b = C::NBLK; b < 34; ++b) p.first[b] = (unsigned short)o;
        p.max_groups = (unsigned short)mg;
        return p;
    }
    template <bool EXACT>
    static void plan_sp(int sm_count, long long B, int* threads, int* grid, int* first, int* n_blocks) {
        int ns, g;
        sp_plan<EXACT>(B, sm_count, ns, g);
        const SPPlan p = sp_threads<EXACT>(ns);
        *threads = ns; *grid = g; *n_blocks = SPCfg<N, EXACT>::NBLK;
        for (int b = 0; b <= SPCfg<N, EXACT>::NBLK; ++b) first[b] = p.first[b];
    }
    static void backward_plan(bool exact, int sm_count, long long B, int* threads, int* grid, int* first, int* n_blocks) {
        if constexpr (USE_SP) {
            if (exact) plan_sp<true>(sm_count, B, threads, grid, first, n_blocks);
            else plan_sp<false>(sm_count, B, threads, grid, first, n_blocks);
        } else {
            *threads = NT; *grid = backward_grid(exact, sm_count, B); *n_blocks = 0; first[0] = 0;
        }
    }
    template <bool EXACT>
    static cudaError_t backward_sp(const float* theta, const BackwardArgs& a, int grid, cudaStream_t st) {
        using C = SPCfg<N, EXACT>;
        int ns, g2;
        sp_plan<EXACT>(a.B, current_sm_count(), ns, g2);
        auto k = backward_sp_kernel<N, EXACT>;
        const SPPlan plan = sp_threads<EXACT>(ns);
        const size_t smem = C::smem_bytes(plan.max_groups);
        cudaError_t e = prep(k, smem);
        if (e != cudaSuccess) return e;
        WBlock<N> w;
        pack(theta, w);
        k<<<grid, ns, smem, st>>>(w, w, plan, a);   // two copies: see the note on common-subexpression elimination in tiny_sp.cuh
        return cudaGetLastError();
    }
    static cudaError_t backward(void*, const float* theta, const BackwardArgs& a, bool exact, int grid, cudaStream_t st) {
        if constexpr (USE_SP) {
            return exact ? backward_sp<true>(theta, a, grid, st) : backward_sp<false>(theta, a, grid, st);
        } else {
            auto k = exact ? backward_kernel<N, true> : backward_kernel<N, false>;
            cudaError_t e = prep(k, smem_bwd);
            if (e != cudaSuccess) return e;
            WBlock<N> w;
            pack(theta, w);
            k<<<grid, NT, smem_bwd, st>>>(w, a);
            return cudaGetLastError();
        }
    }
    static Family make() {
        Family f{};
        f.name = "tiny";
        f.shape.act = N::ACT; f.shape.D = N::D; f.shape.C = N::C; f.shape.NL = N::NL;
        for (int l = 0; l <= N::NL; ++l) f.shape.n[l] = N::n(l);
        f.n_params = N::NP;
        f.rhs = &rhs;
        f.solve_fixed = &solve_fixed;
        f.solve_adaptive = &solve_adaptive;
        f.solve_vcabm = &solve_vcabm;
        f.global_norm_capable = [](void*, const SolveArgs&, bool) { return true; };
        f.adaptive_max_grid = &adaptive_max_grid;
        f.backward = &backward;
        f.backward_grid = &backward_grid;
        f.backward_partials_per_block = 1;
        f.supports_backward = 1;
        f.fuses_loss_sum_adaptive = 1;
        f.adaptive_threads = 64;
        f.ckpt_stages = 6;
        f.backward_plan = USE_SP ? &backward_plan : nullptr;
        return f;
    }
};

}  // namespace tiny
}  // namespace icnf

#define ICNF_TINY_CAT2(a, b) a##b
#define ICNF_TINY_CAT(a, b) ICNF_TINY_CAT2(a, b)
// ICNF_REGISTER_TINY(act, D', ncond, n_layers, n0, n1, ...)
#define ICNF_REGISTER_TINY(...)                                                                       \
    static const icnf::Family ICNF_TINY_CAT(s_family_, __LINE__) =                                    \
        icnf::tiny::Launch<icnf::tiny::Net<__VA_ARGS__>>::make();                                     \
    static icnf::TinyRegistrar ICNF_TINY_CAT(s_registrar_, __LINE__)(&ICNF_TINY_CAT(s_family_, __LINE__));
