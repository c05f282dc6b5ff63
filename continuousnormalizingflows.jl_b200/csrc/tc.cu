// tc.cu -- host side of the tensor-core GEMM (tc_gemm.cuh): TMA tensor maps, launch helper,
// bf16 packing kernels, and a self-test entry point.
#include <algorithm>
#include "tc_gemm.cuh"

#include <cstring>

namespace icnf {
namespace tc {

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
    static EncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return (EncodeFn) nullptr;
        return (EncodeFn)p;
    }();
    return fn;
}

// rows x K bf16 matrix, K contiguous, row pitch `pitch` elements (multiple of 8); box = 64 (K) x box_rows
static bool make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t K, uint64_t pitch, uint32_t box_rows) {
    EncodeFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {K, rows};
    cuuint64_t strides[1] = {pitch * 2};
    cuuint32_t box[2] = {(cuuint32_t)TBK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// lda / ldb are the full row pitches; with g.split every row is [hi | lo] and the tensor map's inner
// extent covers both halves (the zero padding between K and the pitch keeps the hi tiles clean)
cudaError_t gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, TcArgs g, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_SPLIT);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    CUtensorMap mapA, mapB;
    const uint64_t ka = g.split ? (uint64_t)g.lo_a + g.K : (uint64_t)g.K;
    const uint64_t kb = g.split ? (uint64_t)g.lo_b + g.K : (uint64_t)g.K;
    if (!make_map(&mapA, A, (uint64_t)g.M, ka, (uint64_t)lda, TBM) || !make_map(&mapB, B, (uint64_t)g.N, kb, (uint64_t)ldb, TBN))
        return cudaErrorInvalidValue;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // persistent grid: one (split) or two CTAs per SM, each walking tiles blockIdx.x, + gridDim.x, ...
    const long long ntiles = (long long)((g.M + TBM - 1) / TBM) * ((g.N + TBN - 1) / TBN);
    const dim3 grid((unsigned)std::min<long long>(ntiles, (long long)sms * (g.split ? 1 : 2)));
    if (g.split) tc_gemm_kernel<true><<<grid, TTHREADS, SMEM_BYTES_SPLIT, st>>>(mapA, mapB, g);
    else tc_gemm_kernel<false><<<grid, TTHREADS, SMEM_BYTES, st>>>(mapA, mapB, g);
    return cudaGetLastError();
}

// ---- packing kernels ------------------------------------------------------------------------
// fp32 matrix element (r, c) at src[r * rs + c * cs] -> bf16 dst[r * pitch + c], zero padding up to pitch
// `pitch` = columns of one half; with split the row is [hi (pitch) | lo (pitch)]
__global__ void pack_matrix_kernel(const float* src, long long rs, long long cs, __nv_bfloat16* dst, int rows, int cols,
                                   int pitch, int split) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * pitch) return;
    const int r = (int)(idx / pitch), c = (int)(idx - (long long)r * pitch);
    const float x = c < cols ? src[r * rs + c * cs] : 0.f;
    const long long rowp = (long long)r * (split ? 2 * pitch : pitch);
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    dst[rowp + c] = hi;
    if (split) dst[rowp + pitch + c] = __float2bfloat16_rn(x - __bfloat162float(hi));
}
cudaError_t pack_matrix(const float* src, long long rs, long long cs, __nv_bfloat16* dst, int rows, int cols, int pitch,
                        int split, cudaStream_t st) {
    const long long n = (long long)rows * pitch;
    pack_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, rs, cs, dst, rows, cols, pitch, split);
    return cudaGetLastError();
}

// SoA fp32 sources ([row][B], coalesced along samples) -> bf16 rows dst[b][pitch] (coalesced along k),
// through a 32 x 32 shared-memory tile.  Rows k < D come from `zi`, row D is the time (if tin),
// then `ys`; pack_soa is the special case tin = 0, C = 0.
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* zi, const float* ys, __nv_bfloat16* X, long long B, int D,
                                                        int tin, int C, int pitch, float t_fixed, const float* ctrl_f, float c_i,
                                                        const int* done, int split) {
    if (done && *done) return;
    __shared__ float tile[32][33];
    const long long b0 = (long long)blockIdx.x * 32;
    const int k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const float tnow = ctrl_f ? fmaf(c_i, ctrl_f[3] * ctrl_f[1], ctrl_f[0]) : t_fixed;
    for (int kk = ty; kk < 32; kk += 8) {
        const int k = k0 + kk;
        const long long b = b0 + tx;
        float v = 0.f;
        if (b < B) {
            if (k < D) v = zi[(long long)k * B + b];
            else if (tin && k == D) v = tnow;
            else if (k < D + tin + C) v = ys[(long long)(k - D - tin) * B + b];
        }
        tile[kk][tx] = v;
    }
    __syncthreads();
    for (int bb = ty; bb < 32; bb += 8) {
        const long long b = b0 + bb;
        const int k = k0 + tx;
        if (b < B && k < pitch) {
            const float x = tile[tx][bb];
            const long long rowp = b * (split ? 2 * pitch : pitch);
            const __nv_bfloat16 hi = __float2bfloat16_rn(x);
            X[rowp + k] = hi;
            if (split) X[rowp + pitch + k] = __float2bfloat16_rn(x - __bfloat162float(hi));
        }
    }
}
cudaError_t pack_input(const float* zi, const float* ys, __nv_bfloat16* X, long long B, int D, int tin, int C, int pitch,
                       float t_fixed, const float* ctrl_f, float c_i, const int* done, int split, cudaStream_t st) {
    dim3 grid((unsigned)((B + 31) / 32), (unsigned)((pitch + 31) / 32));
    pack_rows_kernel<<<grid, 256, 0, st>>>(zi, ys, X, B, D, tin, C, pitch, t_fixed, ctrl_f, c_i, done, split);
    return cudaGetLastError();
}
cudaError_t pack_soa(const float* src, __nv_bfloat16* dst, long long B, int rows, int pitch, const int* done, int split,
                     cudaStream_t st) {
    dim3 grid((unsigned)((B + 31) / 32), (unsigned)((pitch + 31) / 32));
    pack_rows_kernel<<<grid, 256, 0, st>>>(src, nullptr, dst, B, rows, 0, 0, pitch, 0.f, nullptr, 0.f, done, split);
    return cudaGetLastError();
}

// one hidden layer exact trace: TR[b] = sum_k gvec[k] * D1[b][k]
__global__ void trace_dot_kernel(const float* gvec, const __nv_bfloat16* D1, float* TR, int n1, int pitch, long long B,
                                 const int* done, int split) {
    if (done && *done) return;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long rowp = b * (split ? 2 * pitch : pitch);
    float s = 0.f;
    for (int k = 0; k < n1; ++k) {
        float d = __bfloat162float(D1[rowp + k]);
        if (split) d += __bfloat162float(D1[rowp + pitch + k]);
        s = fmaf(gvec[k], d, s);
    }
    TR[b] = s;
}
cudaError_t trace_dot(const float* gvec, const __nv_bfloat16* D1, float* TR, int n1, int pitch, long long B, const int* done,
                      int split, cudaStream_t st) {
    trace_dot_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(gvec, D1, TR, n1, pitch, B, done, split);
    return cudaGetLastError();
}

}  // namespace tc
}  // namespace icnf

// Self-test of the tensor-core GEMM: D (N x M, SoA: D[n * M + m]) = A (M x K) * B (N x K)' with
// bf16-rounded inputs; host fp32 buffers in and out.
extern "C" __attribute__((visibility("default"))) int icnf_tc_gemm_selftest(int M, int N, int K, const float* A,
                                                                           const float* B, float* D, int split) {
    using namespace icnf::tc;
    const int kp = (K + 7) & ~7;
    const int rp = split ? 2 * kp : kp;
    float *dA = nullptr, *dB = nullptr, *dD = nullptr;
    __nv_bfloat16 *a16 = nullptr, *b16 = nullptr;
    int rc = 0;
    auto ok = [&](cudaError_t e) { if (e != cudaSuccess && rc == 0) rc = 2; return e == cudaSuccess; };
    if (ok(cudaMalloc(&dA, sizeof(float) * M * K)) && ok(cudaMalloc(&dB, sizeof(float) * N * K)) &&
        ok(cudaMalloc(&dD, sizeof(float) * M * N)) && ok(cudaMalloc(&a16, 2 * (size_t)M * rp)) &&
        ok(cudaMalloc(&b16, 2 * (size_t)N * rp))) {
        ok(cudaMemcpy(dA, A, sizeof(float) * M * K, cudaMemcpyHostToDevice));
        ok(cudaMemcpy(dB, B, sizeof(float) * N * K, cudaMemcpyHostToDevice));
        ok(cudaMemset(dD, 0, sizeof(float) * M * N));
        ok(pack_matrix(dA, K, 1, a16, M, K, kp, split, 0));
        ok(pack_matrix(dB, K, 1, b16, N, K, kp, split, 0));
        TcArgs g;
        memset(&g, 0, sizeof g);
        g.M = M; g.N = N; g.K = K; g.ep = TEP_PLAIN_SOA; g.out_f32 = dD; g.n_limit = N;
        g.split = split; g.lo_a = kp; g.lo_b = kp;
        ok(gemm(a16, rp, b16, rp, g, 0));
        ok(cudaDeviceSynchronize());
        ok(cudaMemcpy(D, dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(a16); cudaFree(b16);
    return rc;
}
