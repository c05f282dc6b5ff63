// tc.h -- host interface of the tensor-core GEMM (tc.cu) used by the generic family when
// precision = ICNF_BF16_TC.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace icnf {
namespace tc {

// 3 stages = 96 KB of operand ring per CTA, so two CTAs share an SM and one CTA's epilogue
// overlaps the other's main loop
constexpr int TBM = 128, TBN = 128, TBK = 64, TSTAGES = 3, TTHREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int A_TILE_BYTES = TBM * TBK * 2, B_TILE_BYTES = TBN * TBK * 2;
constexpr int SMEM_BYTES = TSTAGES * (A_TILE_BYTES + B_TILE_BYTES) + 1024 /*align*/ + 256 /*barriers*/ + 2 * TBN * 4 /*bias, double-buffered*/;
// split precision (bf16 x 3: every operand is hi + lo, three MMAs per K step): four tiles per stage
constexpr int SMEM_BYTES_SPLIT = TSTAGES * 2 * (A_TILE_BYTES + B_TILE_BYTES) + 1024 + 256 + 2 * TBN * 4;

enum TcEpilogue {
    TEP_ACT = 0,        // H[m][n] = act(acc + bias[n]), Dv[m][n] = act'   (bf16, row pitch ldo)
    TEP_LIN_SOA = 1,    // out_f32[n * M + m] = acc + bias[n]              (fp32, [unit][sample])
    TEP_MULD = 2,       // G[m][n] = acc * aux[m][n]                        (bf16)
    TEP_PLAIN_SOA = 3,  // out_f32[n * M + m] = acc
    TEP_TRACE = 4       // rowsum[m] (+)= sum_n acc * aux[m][n]
};

struct TcArgs {
    int M, N, K;               // samples, units, reduction
    int ep, act;
    const float* bias;         // N
    __nv_bfloat16* out0;       // H or G
    __nv_bfloat16* out1;       // Dv
    const __nv_bfloat16* aux;  // D (same pitch as out0)
    int ldo;                   // row pitch (elements) of out0/out1/aux
    float* out_f32;            // SoA output / row sums
    int n_limit;               // SoA: only units < n_limit are written
    int atomic_rowsum;         // TEP_TRACE with several unit tiles
    const int* done;
    // split precision: every bf16 matrix row is [hi (pitch) | lo (pitch)]; lo_* = column offset of the
    // lo half in A, B and in out0/out1/aux (ldo is then the full row pitch, 2 x lo_o)
    int split, lo_a, lo_b, lo_o;
};

cudaError_t gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, TcArgs g, cudaStream_t st);
cudaError_t pack_matrix(const float* src, long long rs, long long cs, __nv_bfloat16* dst, int rows, int cols, int pitch,
                        int split, cudaStream_t st);
cudaError_t pack_input(const float* zi, const float* ys, __nv_bfloat16* X, long long B, int D, int tin, int C, int pitch,
                       float t_fixed, const float* ctrl_f, float c_i, const int* done, int split, cudaStream_t st);
cudaError_t trace_dot(const float* gvec, const __nv_bfloat16* D1, float* TR, int n1, int pitch, long long B, const int* done,
                      int split, cudaStream_t st);
cudaError_t pack_soa(const float* src, __nv_bfloat16* dst, long long B, int rows, int pitch, const int* done, int split,
                     cudaStream_t st);
}  // namespace tc
}  // namespace icnf
