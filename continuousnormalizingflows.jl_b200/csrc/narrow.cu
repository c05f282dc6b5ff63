// narrow.cu -- host side of the narrow fast path (kernel: narrow_kernel.cuh; instantiations: narrow_inst_*.cu)
#include <algorithm>
#include <cstring>

#include "narrow_kernel.cuh"

namespace icnf {
namespace narrow {

#define ICNF_NARROW_DECL(N) cudaError_t N(const Params& P, int JP, int grid, size_t smem, cudaStream_t st);
ICNF_NARROW_DECL(launch_o2_softplus_exact) ICNF_NARROW_DECL(launch_o2_any_exact) ICNF_NARROW_DECL(launch_o4_softplus_exact) ICNF_NARROW_DECL(launch_o4_any_exact)
ICNF_NARROW_DECL(launch_o2_softplus_hutch) ICNF_NARROW_DECL(launch_o2_any_hutch) ICNF_NARROW_DECL(launch_o4_softplus_hutch) ICNF_NARROW_DECL(launch_o4_any_hutch)
ICNF_NARROW_DECL(launch_vcabm_o2_exact) ICNF_NARROW_DECL(launch_vcabm_o2_hutch) ICNF_NARROW_DECL(launch_vcabm_o4_exact) ICNF_NARROW_DECL(launch_vcabm_o4_hutch)
#undef ICNF_NARROW_DECL

// ------------------------------------------------------------------ host side
static int round_up(int x, int m) { return (x + m - 1) / m * m; }

bool supported(const icnf_config& cfg, bool exact, const SolveArgs& a) {
    if (cfg.n_layers != 3 || cfg.precision != ICNF_FP32) return false;
    if (a.out_loss) return false;
    if (!exact && a.mode == ICNF_TEST) return false;
    const int D = cfg.nvars + cfg.naug;
    if (D > 32 || cfg.sizes[1] > 128 || cfg.sizes[2] > 128 || cfg.sizes[0] > 160) return false;
    static const bool off = [] { const char* e = getenv("ICNF_NARROW"); return e && atoi(e) == 0; }();
    return !off;
}

static Layout make_layout(const icnf_config& cfg, bool exact, int& JP, int& JP3) {
    Layout L;
    memset(&L, 0, sizeof L);
    L.D = cfg.nvars + cfg.naug; L.C = cfg.ncond; L.tin = cfg.autonomous ? 0 : 1; L.act = cfg.activation;
    L.n0 = cfg.sizes[0]; L.n1 = cfg.sizes[1]; L.n2 = cfg.sizes[2];
    L.jt1 = (L.n1 + NW - 1) / NW; L.jt2 = (L.n2 + NW - 1) / NW; L.jt3 = (L.D + NW - 1) / NW;
    {
        const int need = std::max(L.jt1, L.jt2);
        JP = need <= 4 ? 4 : need <= 8 ? 8 : need <= 10 ? 10 : 16;
    }
    JP3 = L.jt3 <= 2 ? 2 : 4;
    L.ld12 = NW * JP; L.ld3 = NW * JP3;
    int o = 0;
    auto take = [&](int n) { const int at = o; o += round_up(n, 4); return at; };
    L.exact = exact ? 1 : 0;
    L.w1 = take(L.n0 * L.ld12); L.b1 = take(L.ld12);
    L.w2 = take(L.n1 * L.ld12); L.at = exact ? take(L.n1 * L.ld12) : 0; L.b2 = take(L.ld12);
    L.w3 = take(L.n2 * L.ld3); L.b3 = take(L.ld3);
    if (!exact) { L.w3b = take(L.D * L.ld12); L.w2b = take(L.n2 * L.ld12); L.w1b = take(L.n1 * L.ld3); }
    L.x = take(L.n0 * NS);                          // everything before x is zero-filled by the kernel (weight padding)
    L.h1 = take(L.n1 * NS); L.d1 = exact ? take(L.n1 * NS) : 0; L.h2 = take(L.n2 * NS);
    if (!exact) { L.eps = take(L.D * NS); L.zdt = take(L.D * NS); }
    L.red = take((NW + 1) * NS);
    L.total = o;
    return L;
}

size_t smem_bytes(const icnf_config& cfg, bool exact) {
    int JP, JP3;
    return sizeof(float) * (size_t)make_layout(cfg, exact, JP, JP3).total;
}

cudaError_t solve(const icnf_config& cfg, const float* amat, const SolveArgs& a, int nvars, bool exact, bool adaptive, int sm_count,
                  cudaStream_t st) {
    Params P;
    memset(&P, 0, sizeof P);
    P.a = a; P.amat = amat; P.nvars = nvars; P.adaptive = adaptive ? 1 : 0;
    int JP, JP3;
    P.L = make_layout(cfg, exact, JP, JP3);
    long long off = 0;
    for (int l = 0; l < 3; ++l) {
        P.woff[l] = off; off += (long long)cfg.sizes[l] * cfg.sizes[l + 1];
        P.boff[l] = off; off += cfg.sizes[l + 1];
    }
    const size_t smem = sizeof(float) * (size_t)P.L.total;
    if (smem > 227 * 1024 - 1024) return cudaErrorInvalidConfiguration;   // static shared memory of the kernel: < 1 KB
    const long long ntiles = (a.B + NS - 1) / NS;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, sm_count));
    if (a.alg == ICNF_ALG_VCABM) {
        if (!adaptive || !a.vc_hist) return cudaErrorInvalidValue;
        if (exact) return JP3 == 2 ? launch_vcabm_o2_exact(P, JP, grid, smem, st) : launch_vcabm_o4_exact(P, JP, grid, smem, st);
        return JP3 == 2 ? launch_vcabm_o2_hutch(P, JP, grid, smem, st) : launch_vcabm_o4_hutch(P, JP, grid, smem, st);
    }
    const bool sp = cfg.activation == ICNF_ACT_SOFTPLUS;
    if (exact) {
        if (JP3 == 2) return sp ? launch_o2_softplus_exact(P, JP, grid, smem, st) : launch_o2_any_exact(P, JP, grid, smem, st);
        return sp ? launch_o4_softplus_exact(P, JP, grid, smem, st) : launch_o4_any_exact(P, JP, grid, smem, st);
    }
    if (JP3 == 2) return sp ? launch_o2_softplus_hutch(P, JP, grid, smem, st) : launch_o2_any_hutch(P, JP, grid, smem, st);
    return sp ? launch_o4_softplus_hutch(P, JP, grid, smem, st) : launch_o4_any_hutch(P, JP, grid, smem, st);
}

}  // namespace narrow
}  // namespace icnf
