// scan_f32.cu -- K1: coalesced, 128-bit vectorised streaming scan for small query batches.
//
// Replaces faiss's exhaustive_L2sqr_seq / exhaustive_inner_product_seq (the nq < 20 path behind
// index.search at faiss_store.py:64 and rag_datastore_manager.py:218).  Arithmetic is the reference's:
// fp32 exact-difference sum((x - q)^2) for L2, fp32 dot product for IP -- no expanded form, no
// reduced precision, so its distances are final (no re-rank).
//
// HBM-bound by design: every database row is read exactly once per call with 16-byte loads
// (ld.global.nc.L1::no_allocate), R=4 rows x NQ<=8 queries are register-tiled so each query chunk
// read from shared memory is reused 4x, warps keep 8 x 16 B loads in flight, and the per-(query,row)
// totals come out of a transposing butterfly (31 shuffles per 32 results).  Top-k is a per-warp
// sorted list in shared memory guarded by a register threshold; after the first few hundred rows
// almost nothing passes the threshold, so selection costs ~nothing and the distance matrix never
// exists anywhere.  Algorithmic bytes per launch: n * d * sizeof(row element).
#include "common.cuh"

namespace b2f {

constexpr int kScanWarps = 8;
constexpr int kScanThreads = kScanWarps * kWarp;
constexpr int kRowsPerGroup = 4;

template <typename RowT>
struct RowVec;

template <>
struct RowVec<float> {  // 4 fp32 per 16-byte load
    static constexpr int N = 4;
    float v[4];
    __device__ __forceinline__ void load(const float* p) {
        float4 t = ldg_stream(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void zero() { v[0] = v[1] = v[2] = v[3] = 0.f; }
};

template <>
struct RowVec<__nv_bfloat16> {  // 8 bf16 per 16-byte load, widened to fp32 exactly
    static constexpr int N = 8;
    float v[8];
    __device__ __forceinline__ void load(const __nv_bfloat16* p) {
        float4 t = ldg_stream(reinterpret_cast<const float4*>(p));
        const uint32_t w[4] = {__float_as_uint(t.x), __float_as_uint(t.y), __float_as_uint(t.z), __float_as_uint(t.w)};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = 0.f;
    }
};

struct ScalarRow {  // d % 4 != 0: rows are not 16-byte aligned, one element per lane per step
    static constexpr int N = 1;
    float v[1];
    __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
    __device__ __forceinline__ void zero() { v[0] = 0.f; }
};

template <int NQ, bool L2, typename RowT, typename Vec>
__global__ void __launch_bounds__(kScanThreads)
scan_kernel(const RowT* __restrict__ rows, int64_t pitch, int64_t n, int d, int dq, const float* __restrict__ q,
            const int32_t* __restrict__ qsel, int nq_valid, int k, float* __restrict__ pk, int32_t* __restrict__ pi) {
    constexpr int R = kRowsPerGroup;
    constexpr int V = R * NQ;
    constexpr int VN = Vec::N;
    extern __shared__ __align__(16) float smem[];
    float* sq = smem;                                                   // [NQ][dq]
    float* lk = sq + NQ * dq;                                           // [warps][NQ][k]
    int32_t* li = reinterpret_cast<int32_t*>(lk + kScanWarps * NQ * k);  // [warps][NQ][k]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NQ * dq; i += kScanThreads) {
        const int qi = i / dq, c = i - qi * dq;
        sq[i] = (qi < nq_valid && c < d) ? q[(int64_t)(qsel ? qsel[qi] : qi) * d + c] : 0.f;
    }
    for (int i = threadIdx.x; i < kScanWarps * NQ * k; i += kScanThreads) {
        lk[i] = FLT_MAX;
        li[i] = -1;
    }
    __syncthreads();

    float* wlk = lk + warp * NQ * k;
    int32_t* wli = li + warp * NQ * k;
    const int myq = lane & (NQ - 1);
    float thr_k = FLT_MAX;
    int32_t thr_i = -1;

    const int nchunk = dq / VN;
    const int64_t ngroups = (n + R - 1) / R;
    const int64_t W = (int64_t)gridDim.x * kScanWarps;
    for (int64_t g = (int64_t)blockIdx.x * kScanWarps + warp; g < ngroups; g += W) {
        const int64_t row0 = g * R;
        float acc[V];
#pragma unroll
        for (int i = 0; i < V; i++) acc[i] = 0.f;

        Vec xn[R];
        if (lane < nchunk) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (row0 + r < n) xn[r].load(rows + (row0 + r) * pitch + (int64_t)lane * VN);
                else xn[r].zero();
            }
        }
        for (int c = lane; c < nchunk; c += kWarp) {
            Vec x[R];
#pragma unroll
            for (int r = 0; r < R; r++) x[r] = xn[r];
            if (c + kWarp < nchunk) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if (row0 + r < n) xn[r].load(rows + (row0 + r) * pitch + (int64_t)(c + kWarp) * VN);
                    else xn[r].zero();
                }
            }
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) {
                float qv[VN];
                if constexpr (VN == 1) {
                    qv[0] = sq[qi * dq + c];
                } else {
#pragma unroll
                    for (int h = 0; h < VN / 4; h++) {
                        const float4 t = *reinterpret_cast<const float4*>(sq + qi * dq + c * VN + 4 * h);
                        qv[4 * h] = t.x; qv[4 * h + 1] = t.y; qv[4 * h + 2] = t.z; qv[4 * h + 3] = t.w;
                    }
                }
#pragma unroll
                for (int r = 0; r < R; r++) {
#pragma unroll
                    for (int e = 0; e < VN; e++) {
                        if constexpr (L2) {
                            const float t = x[r].v[e] - qv[e];
                            acc[r * NQ + qi] = fmaf(t, t, acc[r * NQ + qi]);
                        } else {
                            acc[r * NQ + qi] = fmaf(x[r].v[e], qv[e], acc[r * NQ + qi]);
                        }
                    }
                }
            }
        }
        warp_multi_reduce<V>(acc, lane);
        const float key = L2 ? acc[0] : -acc[0];
        const int v = lane & (V - 1);
        const int64_t row = row0 + v / NQ;
        const bool pass = lane < V && row < n && (v % NQ) < nq_valid && cand_less(key, (int32_t)row, thr_k, thr_i);
        unsigned m = __ballot_sync(kFull, pass);
        if (m) {
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float ck = __shfl_sync(kFull, key, src);
                const int32_t ci = __shfl_sync(kFull, (int32_t)row, src);
                const int cq = src % NQ;
                float* qk = wlk + cq * k;
                int32_t* qi_ = wli + cq * k;
                if (cand_less(ck, ci, qk[k - 1], qi_[k - 1])) warp_sorted_insert(qk, qi_, k, ck, ci, lane);
            }
            __syncwarp();
            thr_k = wlk[myq * k + k - 1];
            thr_i = wli[myq * k + k - 1];
        }
    }
    __syncthreads();
    // in-CTA merge: warp w < NQ merges the 8 per-warp lists of query w into this CTA's partial list
    if (warp < NQ && warp < nq_valid) {
        const int64_t o = ((int64_t)warp * gridDim.x + blockIdx.x) * k;
        warp_merge_lists(lk + warp * k, li + warp * k, kScanWarps, k, (int64_t)NQ * k, k, pk + o, pi + o, lane);
    }
}

static int g_scan_blocks_per_sm = 2;

int scan_max_parts() { return kNumSMs * 4; }

template <int NQ, bool L2, typename RowT, typename Vec>
static int launch_one(const RowT* rows, int64_t pitch, int64_t n, int d, const float* q, const int32_t* qsel, int nq, int k, float* pk,
                      int32_t* pi, int* nparts_out, cudaStream_t st) {
    const int dq = ((d + Vec::N - 1) / Vec::N) * Vec::N;
    const size_t smem = (size_t)NQ * dq * 4 + (size_t)kScanWarps * NQ * k * 8;
    auto kern = scan_kernel<NQ, L2, RowT, Vec>;
    static bool configured = false;
    static int occ = 1;
    static size_t configured_smem = 0;
    if (!configured || smem > configured_smem) {
        B2F_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 65536 ? smem : 65536)));
        configured_smem = smem > 65536 ? smem : 65536;
        configured = true;
    }
    B2F_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kScanThreads, smem));
    if (occ < 1) {
        set_error("scan kernel does not fit: smem %zu", smem);
        return B2F_EINVAL;
    }
    if (occ > 4) occ = 4;
    g_scan_blocks_per_sm = occ;
    const int64_t ngroups = (n + kRowsPerGroup - 1) / kRowsPerGroup;
    int64_t want = (ngroups + kScanWarps - 1) / kScanWarps;
    int blocks = (int)(want < (int64_t)kNumSMs * occ ? want : (int64_t)kNumSMs * occ);
    if (blocks < 1) blocks = 1;
    *nparts_out = blocks;
    kern<<<blocks, kScanThreads, smem, st>>>(rows, pitch, n, d, dq, q, qsel, nq, k, pk, pi);
    B2F_CUDA(cudaGetLastError());
    return B2F_OK;
}

template <typename RowT, typename Vec>
static int dispatch(const RowT* rows, int64_t pitch, int64_t n, int d, int metric, const float* q, const int32_t* qsel, int nq, int k,
                    float* pk, int32_t* pi, int* nparts_out, cudaStream_t st) {
    const bool l2 = metric == B2F_METRIC_L2;
#define B2F_SCAN_CASE(NQ)                                                                                     \
    return l2 ? launch_one<NQ, true, RowT, Vec>(rows, pitch, n, d, q, qsel, nq, k, pk, pi, nparts_out, st)          \
              : launch_one<NQ, false, RowT, Vec>(rows, pitch, n, d, q, qsel, nq, k, pk, pi, nparts_out, st)
    if (nq <= 1) { B2F_SCAN_CASE(1); }
    if (nq <= 2) { B2F_SCAN_CASE(2); }
    if (nq <= 4) { B2F_SCAN_CASE(4); }
    if (nq <= 8) { B2F_SCAN_CASE(8); }
#undef B2F_SCAN_CASE
    set_error("scan: nq %d > 8 per launch", nq);
    return B2F_EINVAL;
}

int launch_scan_f32(const float* rows, int64_t n, int d, int metric, const float* q, const int32_t* qsel, int nq, int k,
                    float* pk, int32_t* pi, int* nparts_out, cudaStream_t st) {
    if (d % 4 == 0) return dispatch<float, RowVec<float>>(rows, d, n, d, metric, q, qsel, nq, k, pk, pi, nparts_out, st);
    return dispatch<float, ScalarRow>(rows, d, n, d, metric, q, qsel, nq, k, pk, pi, nparts_out, st);
}

int launch_scan_bf16(const __nv_bfloat16* rows, int64_t pitch, int64_t n, int d, int metric, const float* q,
                     const int32_t* qsel, int nq, int k, float* pk, int32_t* pi, int* nparts_out, cudaStream_t st) {
    return dispatch<__nv_bfloat16, RowVec<__nv_bfloat16>>(rows, pitch, n, d, metric, q, qsel, nq, k, pk, pi, nparts_out, st);
}

}  // namespace b2f
