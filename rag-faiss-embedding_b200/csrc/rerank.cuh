// rerank.cuh -- K4 as a block-level device function (128 threads), shared by the standalone re-rank kernel
// (HEAP mode) and the fused list-merge + re-rank kernel (LIST mode).  See rerank.cu for the bound.
#pragma once
#include "common.cuh"

namespace b2f {

constexpr int kRerankThreads = 128;

// ck/ci: the query's kp coarse candidates, ascending by coarse key (shared or global memory).
// ek/ei: kp floats / ints of shared scratch.  All 128 threads of the block must call.
__device__ __forceinline__ void rerank_block(const RerankArgs& a, int q, const float* ck, const int32_t* ci, float* ek,
                                             int32_t* ei) {
    __shared__ float s_tau;
    __shared__ int s_nvalid;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* qv = a.q + (int64_t)q * a.d;
    const bool l2 = a.metric == B2F_METRIC_L2;
    if (threadIdx.x == 0) {
        s_tau = FLT_MAX;
        s_nvalid = 0;
    }
    __syncthreads();
    // 8 lanes per candidate, 4 candidates per warp, 16 per pass: every lane has independent 16-byte
    // loads in flight, so a pass costs about one DRAM round trip instead of eight.
    const int sub = lane >> 3, sl = lane & 7;
    const bool vec_f32 = a.rows_f32 && (a.d & 3) == 0;
    const bool vec_b16 = !a.rows_f32 && (a.d & 7) == 0;
    for (int c0 = warp * 4; c0 < a.kp; c0 += (kRerankThreads / 32) * 4) {
        const int c = c0 + sub;
        int32_t id = c < a.kp ? ci[c] : -1;
        if ((int64_t)id >= a.ntotal) id = -1;  // never dereference a label outside the index
        float acc = 0.f;
        if (id >= 0) {
            if (vec_f32) {
                const float* x = a.rows_f32 + (int64_t)id * a.d;
                for (int j = sl * 4; j < a.d; j += 32) {
                    const float4 xv = *reinterpret_cast<const float4*>(x + j);
                    const float4 qq = *reinterpret_cast<const float4*>(qv + j);
                    if (l2) {
                        const float t0 = xv.x - qq.x, t1 = xv.y - qq.y, t2 = xv.z - qq.z, t3 = xv.w - qq.w;
                        acc = fmaf(t0, t0, acc); acc = fmaf(t1, t1, acc); acc = fmaf(t2, t2, acc); acc = fmaf(t3, t3, acc);
                    } else {
                        acc = fmaf(xv.x, qq.x, acc); acc = fmaf(xv.y, qq.y, acc); acc = fmaf(xv.z, qq.z, acc); acc = fmaf(xv.w, qq.w, acc);
                    }
                }
            } else if (vec_b16) {
                const __nv_bfloat16* x = a.rows_bf16 + (int64_t)id * a.pitch_bf16;
                for (int j = sl * 8; j < a.d; j += 64) {
                    const uint4 w = *reinterpret_cast<const uint4*>(x + j);
                    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const float x0 = __uint_as_float(ww[h] << 16), x1 = __uint_as_float(ww[h] & 0xffff0000u);
                        const float q0 = qv[j + 2 * h], q1 = qv[j + 2 * h + 1];
                        if (l2) {
                            const float t0 = x0 - q0, t1 = x1 - q1;
                            acc = fmaf(t0, t0, acc); acc = fmaf(t1, t1, acc);
                        } else {
                            acc = fmaf(x0, q0, acc); acc = fmaf(x1, q1, acc);
                        }
                    }
                }
            } else {
                for (int j = sl; j < a.d; j += 8) {
                    const float xv = a.rows_f32 ? a.rows_f32[(int64_t)id * a.d + j]
                                                : __bfloat162float(a.rows_bf16[(int64_t)id * a.pitch_bf16 + j]);
                    if (l2) {
                        const float t = xv - qv[j];
                        acc = fmaf(t, t, acc);
                    } else {
                        acc = fmaf(xv, qv[j], acc);
                    }
                }
            }
        }
        acc += __shfl_xor_sync(kFull, acc, 4);
        acc += __shfl_xor_sync(kFull, acc, 2);
        acc += __shfl_xor_sync(kFull, acc, 1);
        if (sl == 0 && c < a.kp) {
            const bool ok = id >= 0 && !(acc != acc);  // NaN never enters (faiss heap semantics)
            ek[c] = ok ? (l2 ? acc : -acc) : FLT_MAX;
            ei[c] = ok ? id : -1;
            if (id >= 0) atomicAdd(&s_nvalid, 1);
        }
    }
    __syncthreads();
    // rank sort by (key, id, slot); ranks are a permutation of [0, kp)
    for (int t = threadIdx.x; t < a.kp; t += kRerankThreads) {
        const float mk = ek[t];
        const int32_t mi = ei[t];
        int rank = 0;
        for (int j = 0; j < a.kp; j++) {
            const float ok_ = ek[j];
            const int32_t oi_ = ei[j];
            rank += (cand_less(ok_, oi_, mk, mi) || (ok_ == mk && oi_ == mi && j < t)) ? 1 : 0;
        }
        if (rank < a.k) {
            if (a.D) {  // fused finalize: faiss conventions straight to the caller's buffers
                const int64_t o = (int64_t)q * a.k + rank;
                a.D[o] = mi < 0 ? (l2 ? FLT_MAX : -FLT_MAX) : (l2 ? mk : -mk);
                a.I[o] = mi < 0 ? -1 : (int64_t)mi + a.id_offset;
            } else {
                a.out_key[(int64_t)q * a.k + rank] = mk;
                a.out_id[(int64_t)q * a.k + rank] = mi;
            }
        }
        if (rank == a.k - 1) s_tau = mk;
    }
    // kp < k cannot happen (host guarantees kp >= k)
    __syncthreads();
    const bool overflowed = a.overflow && a.overflow[q];
    if (threadIdx.x == 0 && (a.certify || overflowed)) {
        bool certified;
        if (overflowed) {
            certified = false;  // some candidates were dropped: only the exact scan can answer
        } else if (s_nvalid < a.kp) {
            certified = true;  // every row of the index is already a candidate
        } else {
            const float tau = s_tau;                               // exact k-th best key
            const float ckp = ck[a.kp - 1];  // k'-th coarse key (without |q~|^2)
            const float qn2 = a.qnorm[q];
            const float qn = sqrtf(qn2);
            const float eq = a.qerr[q];
            // fp32 accumulation slack: gamma_n * sum|q_i x_i| <= n u |q||x| per inner product (n = d + 16 guards
            // the tensor core's chunked accumulation order, u = 2^-24), doubled for non-IEEE accumulator rounding
            const float gam = 4.f * (float)(a.d + 16) * 5.9604645e-8f;
            if (l2) {
                // c = |q~|^2 + |x~|^2 - 2<q~,x~>: the two norms are fp32 sums as well
                const float nu = gam * (2.f * qn * a.max_row_norm + qn2 + a.max_row_norm * a.max_row_norm);
                const float c = ckp + qn2 - nu;
                const float L = sqrtf(fmaxf(c, 0.f)) - eq - a.max_row_err;
                certified = L > 0.f && L * L * (1.f - 4e-7f) > tau;
            } else {
                // keys are negated inner products: non-candidates have <q,x> <= -ckp + slack
                const float nu = gam * qn * a.max_row_norm;
                const float U = -ckp + nu + (qn + eq) * a.max_row_err + a.max_row_norm * eq;
                certified = U < -tau;
            }
        }
        if (!certified) {
            const int slot = atomicAdd(a.fail_count, 1);
            a.fail_list[slot] = q;
            if (overflowed) atomicAdd(a.fail_count + 1, 1);
        }
    }
    // Completion flag: the last block to finish publishes the counters to mapped host memory, so the host
    // learns "done, n failures" by polling one cache line instead of a D2H copy + stream synchronize.
    if (threadIdx.x == 0 && a.host_flag) {
        __threadfence();
        const int prev = atomicAdd(a.fail_count + 4, 1);
        if (prev == a.nblocks - 1) {
            __threadfence();
            volatile int32_t* fc = a.fail_count;
            volatile int32_t* hf = a.host_flag;
            hf[0] = fc[0];
            hf[1] = fc[1];
            hf[2] = fc[2];
            hf[3] = fc[3];
            __threadfence_system();
            hf[4] = a.seq;
        }
    }
}

}  // namespace b2f
