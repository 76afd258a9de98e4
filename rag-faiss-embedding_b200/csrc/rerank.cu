// rerank.cu -- K4: exact fp32 re-rank of the tensor path's coarse candidates + certification.
//
// The tcgen05 pass ranks rows by a bf16-rounded, expanded-form key.  Here every candidate's distance
// is recomputed the way the reference computes it (fp32, exact-difference form for L2 -- faiss
// fvec_L2sqr; fp32 dot product for IP), the k' candidates are re-sorted by (distance, id) and the
// best k are emitted.  A query is *certified* when the rounding-error bound proves that no row
// outside its candidate list can beat its k-th result; uncertified queries are listed for the exact
// streaming scan (K1), so the final answer never depends on bf16.
//
// Bound (L2).  x~, q~ = bf16-rounded row / query, e_x = |x - x~| <= E, e_q = |q - q~|.
//   coarse c(x) = |q~|^2 + |x~|^2 - 2 <q~,x~> + nu,  |nu| <= NU = 4 (d+16) 2^-24 (2|q~||x~| + |q~|^2 + |x~|^2)
//   every non-candidate has c(x) >= c_k'  (the k'-th smallest coarse value)
//   sqrt(dist(q,x)) = |q - x| >= |q~ - x~| - e_q - e_x >= sqrt(max(c_k' - NU, 0)) - e_q - E =: L
//   certified  <=>  L > 0 and L^2 > tau_k  (tau_k = exact k-th best distance among the candidates)
// Bound (IP).  <q,x> <= <q~,x~> + |q| e_x + |x~| e_q  ->  U = s_k' + NU + |q| E + Xmax e_q;
//   certified  <=>  U < tau_k (k-th largest exact inner product).
#include "rerank.cuh"

namespace b2f {

__global__ void __launch_bounds__(kRerankThreads) rerank_kernel(RerankArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* ek = smem;                                        // [kp] exact keys
    int32_t* ei = reinterpret_cast<int32_t*>(smem + a.kp);  // [kp]
    const int q = blockIdx.x;
    const float* ck = a.cand_key + (int64_t)q * a.kp;
    const int32_t* ci = a.cand_id + (int64_t)q * a.kp;
    int valid = 0;
    for (int t = threadIdx.x; t < a.kp; t += kRerankThreads) valid += ci[t] >= 0 ? 1 : 0;
    valid = __syncthreads_count(valid);  // kp <= 256 = block size: one candidate per thread
    __shared__ float s_tau_k;   // exact k-th key of the candidates: what a failed query hands to the range pass
    const bool cert = rerank_block(a, q, ck, ci, a.kp, ck[a.kp - 1], valid < a.kp, ek, ei, &s_tau_k);
    if (threadIdx.x == 0 && a.certify && !cert) rerank_record_failure(a, q, false, s_tau_k);
}

int launch_rerank(const RerankArgs& a, cudaStream_t st) {
    if (a.nq <= 0) return B2F_OK;
    if (a.kp < a.k || a.kp > kRerankThreads) {
        set_error("rerank: kp %d not in [k = %d, %d]", a.kp, a.k, kRerankThreads);
        return B2F_EINVAL;
    }
    rerank_kernel<<<a.nq, kRerankThreads, (size_t)a.kp * 8, st>>>(a);
    B2F_CUDA(cudaGetLastError());
    return B2F_OK;
}

}  // namespace b2f
