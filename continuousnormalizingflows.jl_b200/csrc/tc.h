// tc.h -- host interface of the tensor-core GEMM (tc.cu) used by the generic family when
// precision = ICNF_BF16_TC / ICNF_BF16X3_TC: the forward RHS and the whole reverse sweep.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>

namespace icnf {
namespace tc {

// CTA tile: 128 rows of the A operand (the TMEM lanes) x BN rows of the B operand (TMEM columns), K blocks of 64.
// BN = 256 halves the operand bytes per MAC that a CTA pulls through L2 (the limiter of these GEMMs: the SM's
// share of L2 bandwidth, not the tensor pipe) and turns a 512-unit layer at 8192 samples into ONE wave of
// 128 CTAs; BN = 128 serves narrow outputs.  Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2.. = epilogue, WQ per TMEM lane quarter (the
// epilogue -- activation, hi/lo split, uncoalesced row stores -- is latency-bound: 8 warps left it at twice the main loop).
#ifndef ICNF_TC_WQ
#define ICNF_TC_WQ 3
#endif
constexpr int WQ = ICNF_TC_WQ;                      // epilogue warps per TMEM lane quarter
constexpr int EPI_THREADS = 4 * WQ * 32;
// Warp roles: warps 0 .. 4 WQ - 1 = epilogue, then the MMA issuer, then the TMA producer.  The two single-thread
// roles sit at the HIGHEST warp ids on purpose: the warp scheduler favours high ids, and with the roles the other way
// round the MMA thread needed 3000 cycles instead of 900 to issue a K block while the epilogue warps were busy.
constexpr int TBM = 128, TBK = 64, TTHREADS = EPI_THREADS + 64;
constexpr int MMA_WARP = 4 * WQ, TMA_WARP = 4 * WQ + 1;
constexpr int EPI_STAGE_BYTES = 2048;      // per epilogue warp: one 32-row x 16-column bf16 box, hi and lo halves
constexpr int A_TILE_BYTES = TBM * TBK * 2;
__host__ __device__ constexpr int b_tile_bytes(int bn) { return bn * TBK * 2; }
// shared-memory ring depth: as many stages as fit next to the barriers and the bias slice
__host__ __device__ constexpr int stages(bool split, int bn) { return split ? (bn == 256 ? 2 : 3) : (bn == 256 ? 4 : 6); }
__host__ __device__ constexpr int stage_bytes(bool split, int bn) { return (split ? 2 : 1) * (A_TILE_BYTES + b_tile_bytes(bn)); }
__host__ __device__ constexpr int smem_bytes(bool split, int bn) {
    return stages(split, bn) * stage_bytes(split, bn) + 1024 /*align*/ + 256 /*barriers*/ + 2 * 256 * 4 /*bias, double-buffered*/ +
           4 * WQ * EPI_STAGE_BYTES /*store staging*/;
}

// Tile mode of the GEMM: 1 = one CTA per 128 x 128 tile (the default: measured fastest at the batch sizes of
// BASELINE.json, where a layer is one or two tiles per CTA and the epilogue's write traffic, not the operand loads,
// sets the pace), 2 = CTA pairs (cta_group::2, 256 x 256 tiles: half the operand bytes per MAC), 3 = one CTA per
// 128 x 256 tile.  ICNF_TC_MODE=1|2|3 in the environment forces one (a tuning knob for measurements; every mode
// computes the same products in the same order per output element).
inline int tile_mode(int N) {
    static const int forced = [] { const char* e = getenv("ICNF_TC_MODE"); return e ? atoi(e) : 0; }();
    if (forced >= 1 && forced <= 3) return (forced != 1 && N <= 128) ? 1 : forced;
    return 1;
}
// width of the unit (B-operand row) tiles for a GEMM with N output units
inline int unit_tile_width(int N) { return tile_mode(N) == 1 ? 128 : 256; }
inline int unit_tiles(int N) { const int bn = unit_tile_width(N); return (N + bn - 1) / bn; }

enum TcEpilogue {
    TEP_ACT = 0,        // H[m][n] = act(acc + bias[n]), Dv[m][n] = act' (only when out1 != null)   (bf16, row pitch ldo)  [+ H transposed]
    TEP_LIN_SOA = 1,    // out_f32[n * M + m] = acc + bias[n]              (fp32, [unit][sample])
    TEP_MULD = 2,       // G[m][n] = acc * aux[m][n]                        (bf16)                 [+ G transposed]
    TEP_PLAIN_SOA = 3,  // out_f32[n * M + m] = acc
    TEP_TRACE = 4,      // out_f32[part * M + m] = sum_{n in part} acc * aux[m][n],  part = WQ * unit tile + epilogue warp of the lane quarter
    // reverse sweep (derivation: tiny.cuh rhs_reverse / DESIGN.md):
    TEP_TANGENT = 5,    // out0 = acc * aux (sigma');  out1 = acc * aux1 (chain g) * phi(aux2 (h), aux),  phi = sigma''/sigma'
    TEP_MULADD = 6,     // out0 = acc * aux (sigma') + aux1                                         [+ out0 transposed]
    TEP_WGRAD = 7       // out_f32[slice * slice_stride + n * ldw + m] += acc      (weight gradient, column-major n_out x n_in)
};

struct TcArgs {
    int M, N, K;               // rows of A (TMEM lanes), rows of B (TMEM columns), reduction length of segment 0
    int ep, act;
    const float* bias;         // N
    __nv_bfloat16* out0;       // H / G / Wv / AB
    __nv_bfloat16* out1;       // Dv / AEX
    const __nv_bfloat16* aux;  // sigma' (same pitch as out0), or h when aux_is_h (sigma' is then derived from it)
    int aux_is_h;              // sigma' = f(h) for every supported activation (softplus: 1 - exp(-h), tanh: 1 - h^2,
                               // sigmoid: h (1 - h)): the Hutchinson paths never write or read a sigma' array
    const __nv_bfloat16* aux1; // TANGENT: chain g;  MULADD: AEX
    const __nv_bfloat16* aux2; // TANGENT: h
    int ldo;                   // row pitch (elements) of out0/out1/aux*
    float* out_f32;            // SoA output / row sums / weight-gradient slices
    int n_limit;               // SoA: only units < n_limit are written
    int atomic_rowsum;         // TEP_TRACE with several unit tiles
    const int* done;
    // split precision: every bf16 matrix row is [hi (pitch) | lo (pitch)]; lo_* = column offset of the
    // lo half in A, B and in out0/out1/aux (ldo is then the full row pitch, 2 x lo_o)
    int split, lo_a, lo_b, lo_o;
    // transposed copy of out0 for the weight-gradient GEMMs: outT[n * ldT + m] (hi) and outT[n * ldT + lo_T + m] (lo)
    __nv_bfloat16* outT;
    long long ldT;
    int lo_T;
    // second operand pair accumulated into the same tile (weight gradient: abar h' + g w'): K2 = its reduction
    // length (0: none), N2 = valid rows of its B operand (tiles at n0 >= N2 skip the segment)
    int K2, N2, lo_a2, lo_b2;
    // split-K: the K blocks of every segment are cut into nslices ranges, one work item per (tile, slice)
    int nslices;
    long long slice_stride;
    int ldw;
    // optional timeline of CTA 0 (development aid, icnf_tc_gemm_timeline): trace[0] = event count, then (tag, clock64) pairs.
    // tags: 1000 + it = TMA issued for K block it; 2000 + it = operands of K block it landed (MMA thread);
    //       3000 + it = MMAs of K block it issued; 4000 + i = accumulator i complete (epilogue warp 2); 5000 + i = epilogue of item i done
    long long* trace;
};

// ---- GEMM chains: several dependent GEMMs in ONE persistent launch ------------------------------------------------
// The GEMMs of an RHS evaluation (forward layers, VJP chain) or of a reverse-sweep stage (tangent pass, weight
// gradients, backprop) depend on each other only through the row tile (128 samples) they share, or -- the weight
// gradients, whose reduction runs over the samples -- through a whole earlier GEMM.  gemm_chain() runs such a list as
// one kernel: every CTA walks the concatenated work list in order, the TMA producer of a work item waits on device-side
// counters until the items it reads have been published, and the epilogue warps publish theirs once their stores have
// completed.  The epilogue of one GEMM's last tiles then overlaps the main loop of the next GEMM's first ones, and the
// launch gaps and per-launch pipeline fills of a layer-by-layer sequence disappear.
constexpr int CHAIN_MAXG = 12;
struct ChainStep {
    const __nv_bfloat16 *A, *B, *A2, *B2;
    long long lda, ldb, lda2, ldb2;
    TcArgs g;
    int dep_row[2];   // earlier steps of the chain whose outputs this step reads by row tile (A operand, aux arrays); -1 = none
    int dep_all[2];   // earlier steps that must be complete before any tile of this one starts (operands with K = samples); -1 = none
};
struct ChainState;   // device counters + cached launch parameters (tensor maps) of the chains of one workspace
ChainState* chain_state_create();
void chain_state_destroy(ChainState*);
// slot: which cached parameter block to compare against / refill (one per call site, < 8)
cudaError_t gemm_chain(ChainState* cs, int slot, const ChainStep* steps, int n, const int* done, cudaStream_t st, long long* trace = nullptr);
// tuning knobs (environment at first use; icnf_tc_knob_set changes them at run time for A/B measurements in one process):
// 0 = chains on/off (ICNF_TC_CHAIN), 1 = direct row stores: -1 auto / 0 / 1 (ICNF_CHAIN_DIRECT), 2 = row-tile super-groups:
// 0 off (default) / 1 = three waves per group / n = n waves (ICNF_CHAIN_SG), 3 = cluster size (ICNF_CHAIN_CLUSTER),
// 4 = TMEM accumulators in rotation, 2 (default) or 4 (ICNF_CHAIN_NACC)
int& knob(int which);
inline bool chain_enabled() { return knob(0) != 0 && tile_mode(512) == 1; }   // chains run 128 x 128 tiles

// A (M x K) and B (N x K), both K contiguous; A2/B2: the optional second segment
cudaError_t gemm(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* B, long long ldb, TcArgs g, cudaStream_t st,
                 const __nv_bfloat16* A2 = nullptr, long long lda2 = 0, const __nv_bfloat16* B2 = nullptr, long long ldb2 = 0);
cudaError_t pack_matrix(const float* src, long long rs, long long cs, __nv_bfloat16* dst, int rows, int cols, int pitch,
                        int split, cudaStream_t st);
// SoA fp32 rows [zi (D); t; ys (C)] -> bf16 rows X[b][pitch] and, when XT != null, the transposed copy XT[k][ldT]
cudaError_t pack_input(const float* zi, const float* ys, __nv_bfloat16* X, long long B, int D, int tin, int C, int pitch,
                       float t_fixed, const float* ctrl_f, float c_i, const int* done, int split, cudaStream_t st,
                       __nv_bfloat16* XT = nullptr, long long ldT = 0, int lo_T = 0);
cudaError_t trace_dot(const float* gvec, const __nv_bfloat16* D1, float* TR, int n1, int pitch, long long B, const int* done,
                      int split, cudaStream_t st);
cudaError_t pack_soa(const float* src, __nv_bfloat16* dst, long long B, int rows, int pitch, const int* done, int split,
                     cudaStream_t st, __nv_bfloat16* XT = nullptr, long long ldT = 0, int lo_T = 0);
// db[j] += sum_b XT[j][b] (hi + lo), one CTA per row, fixed summation order
cudaError_t row_sums(const __nv_bfloat16* XT, long long ldT, int lo_T, int split, int rows, long long B, float* db, cudaStream_t st);
}  // namespace tc
}  // namespace icnf
