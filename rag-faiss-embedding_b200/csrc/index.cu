// index.cu -- host side of the C ABI (include/b200flat.h): device storage, ingest, the search
// pipelines (K1 scan | K2 tensor scan -> K3 merge -> K4 re-rank -> finalize), FAISS-compatible
// file I/O.  No CPU compute path exists here: without a usable sm_100 device every compute entry
// point returns B2F_ENOGPU.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"

namespace b2f {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

std::mutex& launch_cache_mutex() {
    static std::mutex m;
    return m;
}

}  // namespace b2f

using namespace b2f;

struct b2f_index {
    int d = 0, metric = 1, storage = 0, device = 0;
    int64_t ntotal = 0, cap = 0, dpad = 0;
    float* rows_f32 = nullptr;        // [cap, d]      authoritative rows (B2F_STORE_F32)
    __nv_bfloat16* scan = nullptr;    // [cap, dpad]   bf16 scan copy (authoritative for B2F_STORE_BF16)
    float* norms = nullptr;           // [cap]         |bf16 row|^2 in fp32
    float* stats = nullptr;           // [2] device    max |x~|^2, max |x - x~|^2
    cudaStream_t stream = nullptr;
    char* ws = nullptr;               // device workspace (grow only)
    size_t ws_bytes = 0;
    char* pinned = nullptr;           // host staging
    size_t pinned_bytes = 0;
    std::vector<cudaEvent_t> ev;      // profiling events
    std::mutex mu;
    b2f_stats st{};
    float host_stats[2] = {0.f, 0.f};
    bool stats_dirty = true;
    float* qpool = nullptr;           // pooled query batch of search_pooled (grow only)
    size_t qpool_bytes = 0;
    float* centre = nullptr;          // [d] device: centre of the scan copy (fp32 storage; see ingest_kernel), fixed at the first add
    bool mu_set = false;
    float mu_norm = 0.f;              // |mu| (host copy, refreshed with host_stats)
    int32_t* host_flag = nullptr;     // mapped pinned memory [16]: counters of the latest finished tensor-path search + its seq
    int32_t seq = 0;
    uint32_t scan_launches = 0;       // launch parity of the scan's cross-CTA bounds
    int32_t harvested_seq = 0;        // last seq whose counters were folded into the host-side statistics
    int slack_boost = 0;              // extra candidates per query, raised when too many queries fail certification
    bool range_mode = false;          // earlier batches left many queries uncertified: run the range pass (one host sync per search)
    int32_t range_ran_seq = 0;        // the search (sequence number) whose range pass served range_ran_nfail queries
    int range_ran_nfail = 0;
    int range_futile = 0;             // consecutive range passes that left most of their queries to the exact scan anyway
    int range_ban = 0;                // searches to wait before the range pass may be switched on again
    int range_quiet = 0;              // consecutive range-mode searches whose first pass certified everything
    // Order-robust mode.  A batch whose candidate lists overflowed en masse (rows stored so that a query's best rows sit in
    // one or two lists: the shared thresholds never tighten) switches the index to per-thread heaps for the first pass
    // (plan_force_heap) plus the range pass for what the heaps cannot certify; LIST mode is tried again after heap_span
    // searches, and the span doubles every time it fails again.
    int heap_left = 0;                // searches still to run in this mode
    int heap_span = 64;
    unsigned long long* totals = nullptr;  // device [4]: fallback queries, overflowed queries, rescued queries (running totals)
    cudaEvent_t ev_done = nullptr;    // recorded at the end of every search (searches return without synchronising)
    cudaStream_t last_stream = nullptr;
    bool search_recorded = false;
    cudaEvent_t ev_ingest = nullptr;  // recorded behind every add on the stream it was enqueued on
    cudaStream_t last_add_stream = nullptr;
    bool add_recorded = false;
    // profiling: a ring of event sets so that timing never forces a host synchronisation inside a search
    struct ProfSlot {
        cudaEvent_t t0 = nullptr, t1 = nullptr;
        std::vector<cudaEvent_t> m;   // 2 per launch of the dominant kernel
        int n_main = 0;
        bool pending = false;
    };
    static constexpr int kProfSlots = 16;
    ProfSlot prof[kProfSlots];
    int prof_head = 0;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int check_device(int device) {
    // cached: cudaGetDeviceProperties costs milliseconds and this runs on per-search entry points
    static std::mutex cache_mu;   // indexes on several devices may be created from several threads
    static int cached_count = -1;
    static int cached_major[64];
    std::lock_guard<std::mutex> cache_lk(cache_mu);
    if (cached_count < 0) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n <= 0) {
            cudaGetLastError();
            set_error("no CUDA device available (%s); this engine has no CPU path", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
            return B2F_ENOGPU;
        }
        if (n > 64) n = 64;
        for (int i = 0; i < n; i++) {
            int major = 0;
            if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) != cudaSuccess) {
                cudaGetLastError();
                major = 0;
            }
            cached_major[i] = major;
        }
        cached_count = n;
    }
    if (device < 0 || device >= cached_count) {
        set_error("device %d out of range [0,%d)", device, cached_count);
        return B2F_EINVAL;
    }
    if (cached_major[device] != 10) {
        set_error("device %d is not sm_100 (Blackwell B200); kernels are built for sm_100a only", device);
        return B2F_ENOGPU;
    }
    return B2F_OK;
}

// simple bump allocator over the index workspace
struct Bump {
    char* base;
    size_t off = 0;
    explicit Bump(char* b) : base(b) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* p = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return p;
    }
};

int ensure_ws(b2f_index* ix, size_t bytes) {
    if (bytes <= ix->ws_bytes) return B2F_OK;
    if (ix->ws) {
        B2F_CUDA(cudaStreamSynchronize(ix->stream));
        B2F_CUDA(cudaDeviceSynchronize());
        B2F_CUDA(cudaFree(ix->ws));
        ix->ws = nullptr;
        ix->ws_bytes = 0;
    }
    bytes = align_up(bytes + (bytes >> 2), 1 << 20);
    cudaError_t e = cudaMalloc(&ix->ws, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("workspace cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return B2F_ENOMEM;
    }
    ix->ws_bytes = bytes;
    return B2F_OK;
}

int ensure_pinned(b2f_index* ix, size_t bytes) {
    if (bytes <= ix->pinned_bytes) return B2F_OK;
    if (ix->pinned) {
        B2F_CUDA(cudaStreamSynchronize(ix->stream));
        B2F_CUDA(cudaFreeHost(ix->pinned));
        ix->pinned = nullptr;
        ix->pinned_bytes = 0;
    }
    cudaError_t e = cudaMallocHost(&ix->pinned, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return B2F_ENOMEM;
    }
    ix->pinned_bytes = bytes;
    return B2F_OK;
}

int ensure_capacity(b2f_index* ix, int64_t need) {
    if (need <= ix->cap) return B2F_OK;
    if (need > (int64_t)INT32_MAX - 64) {
        set_error("a single index holds at most 2^31-65 rows (row-shard across GPUs beyond that)");
        return B2F_EINVAL;
    }
    int64_t ncap = ix->cap + ix->cap / 2;
    if (ncap < need) ncap = need;
    if (ncap < 1024) ncap = 1024;
    float* nrows = nullptr;
    __nv_bfloat16* nscan = nullptr;
    float* nnorm = nullptr;
    cudaError_t e = cudaSuccess;
    if (ix->storage == B2F_STORE_F32) e = cudaMalloc(&nrows, (size_t)ncap * ix->d * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&nscan, (size_t)ncap * ix->dpad * sizeof(__nv_bfloat16));
    if (e == cudaSuccess) e = cudaMalloc(&nnorm, (size_t)ncap * sizeof(float));
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(nrows);
        cudaFree(nscan);
        cudaFree(nnorm);
        set_error("device allocation for %lld rows failed: %s", (long long)ncap, cudaGetErrorString(e));
        return B2F_ENOMEM;
    }
    if (ix->ntotal > 0) {
        // an ingest may still be pending on a caller stream (e.g. behind an encoder forward): its rows must have landed
        // before they are copied into the new buffers
        if (ix->add_recorded && ix->last_add_stream != ix->stream) B2F_CUDA(cudaStreamWaitEvent(ix->stream, ix->ev_ingest, 0));
        if (nrows) B2F_CUDA(cudaMemcpyAsync(nrows, ix->rows_f32, (size_t)ix->ntotal * ix->d * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
        B2F_CUDA(cudaMemcpyAsync(nscan, ix->scan, (size_t)ix->ntotal * ix->dpad * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice, ix->stream));
        B2F_CUDA(cudaMemcpyAsync(nnorm, ix->norms, (size_t)ix->ntotal * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
        ix->st.launches += 0;
    }
    B2F_CUDA(cudaStreamSynchronize(ix->stream));
    B2F_CUDA(cudaDeviceSynchronize());  // other streams may still be searching the old buffers
    cudaFree(ix->rows_f32);
    cudaFree(ix->scan);
    cudaFree(ix->norms);
    ix->rows_f32 = nrows;
    ix->scan = nscan;
    ix->norms = nnorm;
    ix->cap = ncap;
    ix->st.bytes_rows = nrows ? (int64_t)ncap * ix->d * 4 : (int64_t)ncap * ix->dpad * 2;
    ix->st.bytes_scan = (nrows ? (int64_t)ncap * ix->dpad * 2 : 0) + (int64_t)ncap * 4;
    return B2F_OK;
}

cudaEvent_t get_event(b2f_index* ix, size_t i) {
    while (ix->ev.size() <= i) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        ix->ev.push_back(e);
    }
    return ix->ev[i];
}

int refresh_host_stats(b2f_index* ix, cudaStream_t st) {
    if (!ix->stats_dirty) return B2F_OK;
    B2F_CUDA(cudaMemcpyAsync(ix->host_stats, ix->stats, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    std::vector<float> m((size_t)ix->d, 0.f);
    if (ix->mu_set) B2F_CUDA(cudaMemcpyAsync(m.data(), ix->centre, (size_t)ix->d * sizeof(float), cudaMemcpyDeviceToHost, st));
    B2F_CUDA(cudaStreamSynchronize(st));
    double s2 = 0.0;
    for (float v : m) s2 += (double)v * v;
    ix->mu_norm = (float)(sqrt(s2) * (1.0 + 1e-6));
    ix->stats_dirty = false;
    return B2F_OK;
}

// make `st` wait for everything queued so far on the index's own stream (ingest) and vice versa
int order_after(cudaStream_t waiter, cudaStream_t producer, b2f_index* ix) {
    if (waiter == producer) return B2F_OK;
    cudaEvent_t e = get_event(ix, 0);
    if (!e) {
        set_error("cudaEventCreate failed");
        return B2F_ECUDA;
    }
    B2F_CUDA(cudaEventRecord(e, producer));
    B2F_CUDA(cudaStreamWaitEvent(waiter, e, 0));
    return B2F_OK;
}

// Adds are enqueued on the caller's stream (a device-tensor add / add_pooled runs behind the encoder on torch's
// stream) and return without synchronising.  Everything that reads rows, norms, stats or the centre on ANOTHER
// stream -- searches, write, reconstruct, the stats refresh, the growth copies -- first waits for the latest add.
int wait_ingest(b2f_index* ix, cudaStream_t st) {
    if (ix->add_recorded && ix->last_add_stream != st) B2F_CUDA(cudaStreamWaitEvent(st, ix->ev_ingest, 0));
    return B2F_OK;
}
int record_ingest(b2f_index* ix, cudaStream_t st) {
    if (!ix->ev_ingest) B2F_CUDA(cudaEventCreateWithFlags(&ix->ev_ingest, cudaEventDisableTiming));
    // an add on a new stream is ordered behind the previous add (both write stats / may share the workspace)
    B2F_CUDA(cudaEventRecord(ix->ev_ingest, st));
    ix->last_add_stream = st;
    ix->add_recorded = true;
    return B2F_OK;
}

// ---- K1: one cooperative launch (scan + merge + faiss formatting) for any number of queries ----------
// qsel / nsel_dev (optional, device): the list of query rows to process and its length -- the tensor path's
// uncertified queries, which the host never counts.  counters != null marks the closing kernel of a
// tensor-path search: it also publishes the search's counters.
int enqueue_scan(b2f_index* ix, const float* qd, const int32_t* qsel, const int32_t* nsel_dev, int nsel, int k, float* Dd,
                 int64_t* Id, int64_t id_offset, void* scratch, int32_t* counters, int32_t seq, int nq_batch, int certify,
                 cudaStream_t st) {
    ScanArgs a{};
    a.rows_f32 = ix->storage == B2F_STORE_F32 ? ix->rows_f32 : nullptr;
    a.rows_bf16 = ix->scan;
    a.pitch_bf16 = ix->dpad;
    a.mu = (ix->storage == B2F_STORE_BF16 && ix->mu_set) ? ix->centre : nullptr;
    a.n = ix->ntotal;
    a.d = ix->d;
    a.metric = ix->metric;
    a.k = k;
    a.q = qd;
    a.qsel = qsel;
    a.nsel_dev = nsel_dev;
    a.nsel = nsel;
    a.D = Dd;
    a.I = Id;
    a.id_offset = id_offset;
    a.scratch = scratch;
    a.counters = counters;
    a.totals = counters ? ix->totals : nullptr;
    a.host_flag = counters ? ix->host_flag : nullptr;
    a.seq = seq;
    a.nq_batch = nq_batch;
    a.certify = certify;
    a.tub = reinterpret_cast<uint32_t*>(ix->totals + 4);
    a.parity = (int32_t)(ix->scan_launches++ & 1);
    B2F_TRY(launch_scan(a, st));
    ix->st.launches += 1;
    ix->st.last_launches += 1;
    return B2F_OK;
}

// Folds the counters the latest finished tensor-path search left in mapped host memory into the host-side
// view (never blocks; a search that is still running is simply picked up later).
void harvest_flag(b2f_index* ix) {
    volatile int32_t* hf = ix->host_flag;
    const int32_t s = hf[4];
    if (s == 0 || s == ix->harvested_seq) return;
    std::atomic_thread_fence(std::memory_order_acquire);
    const int32_t c0 = hf[0], c1 = hf[1], c2 = hf[2], c3 = hf[3], nqb = hf[6], certify = hf[7];
    std::atomic_thread_fence(std::memory_order_acquire);
    if (hf[4] != s) return;  // a later search is publishing right now: take that one next time
    ix->harvested_seq = s;
    ix->st.last_list_entries = (int64_t)(((uint64_t)(uint32_t)c3 << 32) | (uint32_t)c2);
    // The slack that certification needs grows with the neighbour density at rank k (i.e. with the database
    // size and the data distribution): when too many queries of a batch had to fall back, keep more candidates
    // per query from now on.  "Too many" is where the exact scans cost more than the wider tensor pass would: the
    // fallback walks the database once per four queries (F / 4 x n d 4 B at ~5.5 TB/s) while the batch's tensor
    // pass costs 2 nq n d flop at ~1.3 PFLOP/s and grows by < 10 % with 32 more candidates -- break-even at
    // F ~ nq / 1200; the boost waits for 0.5 % because the range pass below is cheaper still.  (Round 1 waited for 2 % of
    // the batch: BASELINE config 5 on 2 GPUs spent 9 of its 34 ms per search in exact scans for 0.5 % of the queries.)
    if (certify && c0 - c1 > (nqb / 200 > 2 ? nqb / 200 : 2) && ix->slack_boost < 224) ix->slack_boost += 32;
    // Well before that (F > nq / 1000) the range pass takes over: a second TENSOR pass over the uncertified queries with a
    // fixed threshold per query, one database pass for all of them instead of one exact scan per four (search_locked).
    // ... unless it does not help: neighbourhoods so dense that the fixed-threshold lists overflow too (thousands of
    // near-duplicates) leave its queries to the exact scan anyway; two such searches in a row switch it off for a while.
    if (ix->range_ban > 0) ix->range_ban--;
    if (ix->range_mode && ix->range_ran_seq == s) {
        if (2 * c0 > ix->range_ran_nfail) {
            if (++ix->range_futile >= 2) {
                ix->range_mode = false;
                ix->range_futile = 0;
                ix->range_ban = 256;
            }
        } else {
            ix->range_futile = 0;
        }
    }
    if (certify && c1 > (nqb / 8 > 8 ? nqb / 8 : 8) && ix->heap_left == 0) {
        ix->heap_left = ix->heap_span;
        if (ix->heap_span < 4096) ix->heap_span *= 2;
        if (ix->range_ban == 0) {
            ix->range_mode = true;   // the heaps find the k' best whatever the row order; dense neighbourhoods still need the range pass
            ix->range_quiet = 0;
        }
    }
    if (certify && !ix->range_mode && ix->range_ban == 0 && c0 - c1 > (nqb / 1000 > 2 ? nqb / 1000 : 2)) {
        ix->range_mode = true;
        ix->range_quiet = 0;
    }
}

void harvest_prof_slot(b2f_index* ix, b2f_index::ProfSlot& sl) {
    if (!sl.pending) return;
    sl.pending = false;
    if (cudaEventSynchronize(sl.t1) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    float ms = 0.f, tot = 0.f;
    for (int i = 0; i < sl.n_main; i++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, sl.m[2 * (size_t)i], sl.m[2 * (size_t)i + 1]) == cudaSuccess) ms += t;
    }
    if (cudaEventElapsedTime(&tot, sl.t0, sl.t1) != cudaSuccess) tot = 0.f;
    cudaGetLastError();
    ix->st.last_main_ms = ms;
    ix->st.last_total_ms = tot;
    ix->st.last_main_launches = sl.n_main;
    ix->st.prof_main_ms_sum += ms;
    ix->st.prof_total_ms_sum += tot;
    ix->st.prof_main_launches += sl.n_main;
    ix->st.prof_searches += 1;
}

cudaEvent_t prof_main_event(b2f_index::ProfSlot& sl, size_t i) {
    while (sl.m.size() <= i) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        sl.m.push_back(e);
    }
    return sl.m[i];
}

// Candidates kept per query by the tensor pass.  The slack above k must cover the rows whose coarse
// (bf16) key can fall inside the rounding band around the k-th result; that count grows with k (the
// neighbour density at rank k), hence ~0.9 k with a floor of 22.  32 / 64 are also the heap sizes of HEAP
// mode; larger values (multiples of 8, up to 256) exist in LIST mode only.  0 = k too large for K2.
int tensor_kprime(int k, int slack) {
    int extra = slack > 0 ? slack : ((9 * k + 9) / 10 > 22 ? (9 * k + 9) / 10 : 22);
    int kp = k + extra;
    if (kp <= 32) return 32;
    if (kp <= 64) return 64;
    kp = (kp + 7) / 8 * 8;
    return kp <= 256 ? kp : 0;
}

// Query chunking of the tensor path.  A batch is processed in several passes for feasibility only: more query tile
// units than units (one wave), or a big k' with so many tiles that a tile is left with too few voucher lists for the
// shared threshold (j <= 16).  Balance is no longer a reason: a pass's work is shared equally among the units whatever
// the tile count is (k2::Seg), so 16 pair tiles over 74 pairs cost 16/74 of the database per unit in ONE pass (before:
// tiles got 4 or 5 whole units, and two passes of 8 tiles were cheaper than one of 16).
// The planner still evaluates 1..8 equal passes and a few fixed pass sizes with a small cost model (tensor time per
// unit, floored by the HBM time of one database pass, plus a fixed per-pass overhead) and returns the cheapest.
int plan_tensor_chunked(int nq, int64_t n, int d, int kp, TensorScanPlan* plan) {
    int feasible = nq;
    while (true) {
        if (plan_tensor_scan(feasible, n, d, kp, plan) == B2F_OK) break;
        if (feasible <= 128) return 0;
        feasible = ((feasible / 2 + 127) / 128) * 128;
    }
    const double dpad = (double)((d + 63) / 64 * 64);
    const double kSmRate = 10.0e12, kHbm = 6.0e12, kPassOverhead = 20e-6;
    const double t_hbm = (double)n * dpad * 2.0 / kHbm;
    int best_chunk = feasible;
    double best_cost = 1e30;
    const char* nb = getenv("B200FLAT_NO_BALANCE");   // diagnostics: "1" = feasibility chunking only
    const bool balance = !(nb && nb[0] == '1');
    // candidates: 1..8 equal passes, and the fixed pass sizes that fill one wave of pairs (74 / 37 pair tiles) or
    // give every tile many splits -- so that very large batches are cut into LIST-mode passes instead of falling
    // into the multi-wave HEAP selection (the first version of K2, ~6x slower per query)
    int cands[12];
    int nc = 0;
    for (int m = 1; m <= (balance ? 8 : 1); m++) {
        int chunk = (nq + m - 1) / m;
        chunk = (chunk + 255) / 256 * 256;   // whole pair tiles
        if (chunk > feasible) chunk = m == 1 ? feasible : 0;
        if (chunk >= 256 || m == 1) cands[nc++] = chunk;
    }
    if (balance) {
        const int fixed[4] = {74 * 256, 37 * 256, 4096, 2048};
        for (int i = 0; i < 4; i++)
            if (fixed[i] < nq && fixed[i] <= feasible) cands[nc++] = fixed[i];
    }
    for (int i = 0; i < nc; i++) {
        const int chunk = cands[i];
        if (chunk <= 0) continue;
        TensorScanPlan p{};
        if (plan_tensor_scan(chunk, n, d, kp, &p) != B2F_OK) continue;
        const int passes = (nq + chunk - 1) / chunk;
        const double waves = p.list_mode ? 1.0 : (double)((p.units + kNumSMs - 1) / kNumSMs);
        // rows every unit streams: LIST mode shares the pass equally (tile_units / units of the database per unit,
        // whatever the two numbers are), HEAP mode splits every tile nsplits ways
        const double rows_per_unit = p.list_mode ? (double)n * p.tile_units / p.units : (double)n / p.nsplits;
        double t_mma = rows_per_unit * 128.0 * dpad * 2.0 / kSmRate * waves;  // per SM
        if (!p.list_mode) t_mma *= 6.0;   // per-thread heaps instead of shared-threshold lists
        const double cost = passes * ((t_mma > t_hbm ? t_mma : t_hbm) + kPassOverhead);
        if (cost < best_cost * 0.97) {  // a new candidate has to buy at least 3%
            best_cost = cost;
            best_chunk = chunk;
        }
    }
    if (plan_tensor_scan(best_chunk, n, d, kp, plan) != B2F_OK) return 0;
    return best_chunk;
}

}  // namespace

// =================================================================================================
extern "C" {

int b2f_version(void) { return B2F_VERSION; }

const char* b2f_last_error(void) { return g_err; }

int b2f_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
    }
    return ok;
}

int b2f_index_create(int32_t d, int32_t metric, int32_t storage, int32_t device, b2f_index** out) {
    if (!out) {
        set_error("out is NULL");
        return B2F_EINVAL;
    }
    *out = nullptr;
    if (d <= 0 || d > 16384) {
        set_error("d=%d out of range [1,16384]", d);
        return B2F_EINVAL;
    }
    if (metric != B2F_METRIC_L2 && metric != B2F_METRIC_INNER_PRODUCT) {
        set_error("metric %d not supported (0 = inner product, 1 = L2)", metric);
        return B2F_EINVAL;
    }
    if (storage != B2F_STORE_F32 && storage != B2F_STORE_BF16) {
        set_error("storage %d not supported", storage);
        return B2F_EINVAL;
    }
    B2F_TRY(check_device(device));
    DeviceGuard g(device);
    b2f_index* ix = new (std::nothrow) b2f_index();
    if (!ix) return B2F_ENOMEM;
    ix->d = d;
    ix->metric = metric;
    ix->storage = storage;
    ix->device = device;
    ix->dpad = (int64_t)align_up((size_t)d, 64);
    cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&ix->stats, 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(ix->stats, 0, 2 * sizeof(float));
    // centre [d] floats, then (8-byte aligned) d 64-bit words: the fixed-point accumulators its mean is summed in
    const size_t centre_bytes = (((size_t)d * sizeof(float) + 7) & ~(size_t)7) + (size_t)d * 8;
    if (e == cudaSuccess) e = cudaMalloc(&ix->centre, centre_bytes);
    if (e == cudaSuccess) e = cudaMemset(ix->centre, 0, centre_bytes);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&ix->host_flag), 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) memset(ix->host_flag, 0, 64);
    if (e == cudaSuccess) e = cudaMalloc(&ix->totals, 4 * sizeof(unsigned long long) + kScanTubWords * 4);
    if (e == cudaSuccess) e = cudaMemset(ix->totals, 0, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ix->totals + 4, 0xff, kScanTubWords * 4);  // the scan's cross-CTA bounds, idle = all ones
    if (e != cudaSuccess) {
        set_error("index init failed: %s", cudaGetErrorString(e));
        delete ix;
        return B2F_ECUDA;
    }
    *out = ix;
    return B2F_OK;
}

int b2f_index_destroy(b2f_index* ix) {
    if (!ix) return B2F_OK;
    DeviceGuard g(ix->device);
    cudaDeviceSynchronize();
    cudaFree(ix->rows_f32);
    cudaFree(ix->scan);
    cudaFree(ix->norms);
    cudaFree(ix->stats);
    cudaFree(ix->centre);
    cudaFree(ix->qpool);
    cudaFree(ix->ws);
    if (ix->pinned) cudaFreeHost(ix->pinned);
    if (ix->host_flag) cudaFreeHost(ix->host_flag);
    cudaFree(ix->totals);
    for (cudaEvent_t e : ix->ev) cudaEventDestroy(e);
    if (ix->ev_done) cudaEventDestroy(ix->ev_done);
    if (ix->ev_ingest) cudaEventDestroy(ix->ev_ingest);
    for (auto& sl : ix->prof) {
        if (sl.t0) cudaEventDestroy(sl.t0);
        if (sl.t1) cudaEventDestroy(sl.t1);
        for (cudaEvent_t e : sl.m) cudaEventDestroy(e);
    }
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
    return B2F_OK;
}

int b2f_index_reset(b2f_index* ix) {
    if (!ix) {
        set_error("index is NULL");
        return B2F_EINVAL;
    }
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    ix->ntotal = 0;
    ix->mu_set = false;
    // what earlier searches taught the index about its data goes with the data
    ix->slack_boost = 0;
    ix->range_mode = false;
    ix->range_quiet = ix->range_futile = ix->range_ban = 0;
    ix->heap_left = 0;
    ix->heap_span = 64;
    if (ix->search_recorded) B2F_CUDA(cudaEventSynchronize(ix->ev_done));
    if (ix->add_recorded) B2F_CUDA(cudaEventSynchronize(ix->ev_ingest));
    B2F_CUDA(cudaMemsetAsync(ix->stats, 0, 2 * sizeof(float), ix->stream));
    B2F_CUDA(cudaStreamSynchronize(ix->stream));
    ix->stats_dirty = true;
    return B2F_OK;
}

int b2f_index_reserve(b2f_index* ix, int64_t nrows) {
    if (!ix) {
        set_error("index is NULL");
        return B2F_EINVAL;
    }
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    return ensure_capacity(ix, nrows);
}

int64_t b2f_index_ntotal(const b2f_index* ix) { return ix ? ix->ntotal : -1; }
int32_t b2f_index_d(const b2f_index* ix) { return ix ? ix->d : -1; }
int32_t b2f_index_metric(const b2f_index* ix) { return ix ? ix->metric : -1; }
int32_t b2f_index_storage(const b2f_index* ix) { return ix ? ix->storage : -1; }
int32_t b2f_index_device(const b2f_index* ix) { return ix ? ix->device : -1; }

int b2f_plan_describe(int64_t nq, int64_t n, int32_t d, int64_t k, int32_t slack, int32_t out[12]) {
    if (!out || nq <= 0 || n <= 0 || d <= 0 || k <= 0 || nq > (1 << 24)) {
        set_error("plan_describe: bad arguments");
        return B2F_EINVAL;
    }
    for (int i = 0; i < 12; i++) out[i] = 0;
    const int kp = k <= 1024 ? tensor_kprime((int)k, slack) : 0;
    if (kp <= 0) return B2F_OK;
    TensorScanPlan plan{};
    const int chunk = plan_tensor_chunked((int)nq, n, d, kp, &plan);
    if (chunk <= 0) return B2F_OK;
    out[0] = kp;
    out[1] = chunk;
    out[2] = (int32_t)((nq + chunk - 1) / chunk);
    out[3] = plan.list_mode;
    out[4] = plan.pair_mode;
    out[5] = plan.units;
    out[6] = plan.nsplits;
    out[7] = plan.nlists;
    out[8] = plan.list_j;
    out[9] = plan.list_cap;
    out[10] = plan.nq_tiles;
    out[11] = plan.round_tiles;
    return B2F_OK;
}

int b2f_plan_unit_work(int32_t tile_units, int32_t units, int32_t round_tiles, int64_t db_tiles, int32_t kprime, int32_t unit,
                       int32_t seg_info[14], int64_t* tiles, int64_t cap, int32_t counts[2]) {
    const int rc = plan_unit_work(tile_units, units, round_tiles, db_tiles, kprime, unit, seg_info, tiles, cap, counts);
    if (rc < 0) set_error("plan_unit_work: bad arguments");
    return rc < 0 ? B2F_EINVAL : rc;
}

int b2f_index_stats(const b2f_index* cix, b2f_stats* out) {
    if (!cix || !out) {
        set_error("NULL argument");
        return B2F_EINVAL;
    }
    // searches return without synchronising and keep their counters on the device: settle them first
    b2f_index* ix = const_cast<b2f_index*>(cix);
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    if (g.ok && ix->totals) {
        if (ix->search_recorded) B2F_CUDA(cudaEventSynchronize(ix->ev_done));
        unsigned long long t[4] = {0, 0, 0, 0};
        B2F_CUDA(cudaMemcpy(t, ix->totals, sizeof(t), cudaMemcpyDeviceToHost));
        ix->st.fallback_queries = (int64_t)t[0];
        ix->st.overflow_queries = (int64_t)t[1];
        ix->st.rescued_queries = (int64_t)t[2];
        harvest_flag(ix);
        for (int i = 0; i < b2f_index::kProfSlots; i++)
            harvest_prof_slot(ix, ix->prof[(ix->prof_head + i) % b2f_index::kProfSlots]);  // oldest first
    }
    *out = ix->st;
    return B2F_OK;
}

// ---- add -------------------------------------------------------------------------------------------
static bool centring_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200FLAT_NO_CENTER");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

static int add_device_rows(b2f_index* ix, const float* src_dev, int64_t n, cudaStream_t st) {
    // src_dev: [n, d] fp32 on the device (may alias the tail of rows_f32)
    float* rows_out = nullptr;
    if (ix->storage == B2F_STORE_F32) {
        float* dst = ix->rows_f32 + ix->ntotal * ix->d;
        if (src_dev != dst) rows_out = dst;
    }
    // The scan copy is taken around the mean of the first rows the index sees (at most 65536), fixed from then on: any
    // centre is correct, a representative one makes the bf16 pass far more decisive on embeddings that share a large
    // common component.  bf16 storage keeps the same centred copy as its ONLY copy: the authoritative row is then
    // fl32(mu + bf16(x - mu)) -- closer to x than bf16(x) on such data, and the tensor pass certifies as on fp32 storage.
    if (!ix->mu_set && ix->ntotal == 0 && centring_enabled()) {
        unsigned long long* acc = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ix->centre) + (((size_t)ix->d * sizeof(float) + 7) & ~(size_t)7));
        B2F_TRY(launch_mean_rows(src_dev, n < 65536 ? n : 65536, ix->d, ix->centre, acc, st));
        ix->mu_set = true;
        ix->st.launches += 2;
    }
    B2F_TRY(launch_ingest(src_dev, n, ix->d, rows_out, ix->scan + ix->ntotal * ix->dpad, ix->dpad,
                          ix->norms + ix->ntotal, ix->stats, ix->mu_set ? ix->centre : nullptr,
                          ix->metric == B2F_METRIC_INNER_PRODUCT ? 1 : 0, ix->storage == B2F_STORE_BF16 ? 1 : 0, st));
    ix->st.launches++;
    return B2F_OK;
}

int b2f_index_add(b2f_index* ix, int64_t n, const float* x, int32_t mem, void* stream) {
    if (!ix || n < 0 || (n > 0 && !x)) {
        set_error("add: bad arguments");
        return B2F_EINVAL;
    }
    if (n == 0) return B2F_OK;
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    if (!g.ok) {
        set_error("cudaSetDevice(%d) failed", ix->device);
        return B2F_ENOGPU;
    }
    B2F_TRY(ensure_capacity(ix, ix->ntotal + n));
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    if (ix->search_recorded && ix->last_stream != st) B2F_CUDA(cudaStreamWaitEvent(st, ix->ev_done, 0));
    B2F_TRY(wait_ingest(ix, st));
    if (mem == B2F_MEM_DEVICE) {
        B2F_TRY(add_device_rows(ix, x, n, st));
        ix->ntotal += n;
    } else if (ix->storage == B2F_STORE_F32) {
        // host -> final location, then derive the scan copy in place
        float* dst = ix->rows_f32 + ix->ntotal * ix->d;
        B2F_CUDA(cudaMemcpyAsync(dst, x, (size_t)n * ix->d * sizeof(float), cudaMemcpyHostToDevice, st));
        B2F_TRY(add_device_rows(ix, dst, n, st));
        ix->ntotal += n;
    } else {
        // bf16 storage: stage fp32 chunks through the workspace
        const int64_t chunk = (64LL << 20) / ((int64_t)ix->d * 4) > 0 ? (64LL << 20) / ((int64_t)ix->d * 4) : 1;
        B2F_TRY(ensure_ws(ix, (size_t)chunk * ix->d * 4 + 4096));
        for (int64_t r0 = 0; r0 < n; r0 += chunk) {
            const int64_t m = n - r0 < chunk ? n - r0 : chunk;
            B2F_CUDA(cudaMemcpyAsync(ix->ws, x + r0 * ix->d, (size_t)m * ix->d * 4, cudaMemcpyHostToDevice, st));
            B2F_TRY(add_device_rows(ix, reinterpret_cast<const float*>(ix->ws), m, st));
            ix->ntotal += m;
        }
    }
    ix->stats_dirty = true;
    B2F_TRY(record_ingest(ix, st));
    if (mem == B2F_MEM_HOST) B2F_CUDA(cudaStreamSynchronize(st));
    return B2F_OK;
}

int b2f_index_add_synth(b2f_index* ix, uint64_t seed, int64_t row0, int64_t nrows, int32_t normalize) {
    if (!ix || nrows < 0) {
        set_error("add_synth: bad arguments");
        return B2F_EINVAL;
    }
    if (nrows == 0) return B2F_OK;
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    B2F_TRY(ensure_capacity(ix, ix->ntotal + nrows));
    cudaStream_t st = ix->stream;
    if (ix->search_recorded && ix->last_stream != st) B2F_CUDA(cudaStreamWaitEvent(st, ix->ev_done, 0));
    B2F_TRY(wait_ingest(ix, st));
    if (ix->storage == B2F_STORE_F32) {
        float* dst = ix->rows_f32 + ix->ntotal * ix->d;
        B2F_TRY(launch_synth(seed, row0, nrows, ix->d, normalize, dst, st));
        B2F_TRY(add_device_rows(ix, dst, nrows, st));
        ix->ntotal += nrows;
        ix->st.launches++;
    } else {
        const int64_t chunk = (256LL << 20) / ((int64_t)ix->d * 4) > 0 ? (256LL << 20) / ((int64_t)ix->d * 4) : 1;
        B2F_TRY(ensure_ws(ix, (size_t)chunk * ix->d * 4 + 4096));
        for (int64_t r0 = 0; r0 < nrows; r0 += chunk) {
            const int64_t m = nrows - r0 < chunk ? nrows - r0 : chunk;
            B2F_TRY(launch_synth(seed, row0 + r0, m, ix->d, normalize, reinterpret_cast<float*>(ix->ws), st));
            B2F_TRY(add_device_rows(ix, reinterpret_cast<const float*>(ix->ws), m, st));
            ix->ntotal += m;
            ix->st.launches++;
        }
    }
    ix->stats_dirty = true;
    B2F_TRY(record_ingest(ix, st));
    B2F_CUDA(cudaStreamSynchronize(st));
    return B2F_OK;
}

int b2f_index_add_pooled(b2f_index* ix, const float* hidden, const int64_t* mask, int64_t B, int64_t T, int32_t pool,
                         int32_t normalize, void* stream) {
    if (!ix || B < 0 || (B > 0 && !hidden)) {
        set_error("add_pooled: bad arguments");
        return B2F_EINVAL;
    }
    if (B == 0) return B2F_OK;
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    B2F_TRY(ensure_capacity(ix, ix->ntotal + B));
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    if (ix->search_recorded && ix->last_stream != st) B2F_CUDA(cudaStreamWaitEvent(st, ix->ev_done, 0));
    B2F_TRY(wait_ingest(ix, st));
    // The pooled rows go through the SAME ingest kernel every other add uses: it derives the scan copy, the per-row
    // bias of the tensor pass (L2: |x~'|^2, IP: -mu.x -- NOT a norm) and the stats for either storage mode and metric.
    if (ix->storage == B2F_STORE_F32) {
        // pooled rows land in their final place
        float* dst = ix->rows_f32 + ix->ntotal * ix->d;
        B2F_TRY(launch_pool(hidden, mask, B, T, ix->d, pool, normalize, dst, nullptr, 0, nullptr, nullptr, st));
        ix->st.launches++;
        B2F_TRY(add_device_rows(ix, dst, B, st));
        ix->ntotal += B;
    } else {
        // bf16 storage: pool fp32 chunks into the workspace, ingest from there
        const int64_t chunk = (64LL << 20) / ((int64_t)ix->d * 4) > 0 ? (64LL << 20) / ((int64_t)ix->d * 4) : 1;
        B2F_TRY(ensure_ws(ix, (size_t)(B < chunk ? B : chunk) * ix->d * 4 + 4096));
        for (int64_t b0 = 0; b0 < B; b0 += chunk) {
            const int64_t m = B - b0 < chunk ? B - b0 : chunk;
            float* tmp = reinterpret_cast<float*>(ix->ws);
            B2F_TRY(launch_pool(hidden + b0 * T * ix->d, mask ? mask + b0 * T : nullptr, m, T, ix->d, pool, normalize, tmp, nullptr, 0,
                                nullptr, nullptr, st));
            ix->st.launches++;
            B2F_TRY(add_device_rows(ix, tmp, m, st));
            ix->ntotal += m;
        }
    }
    ix->stats_dirty = true;
    B2F_TRY(record_ingest(ix, st));
    return B2F_OK;
}

int b2f_pool_normalize(const float* hidden, const int64_t* mask, int64_t B, int64_t T, int32_t d, int32_t pool,
                       int32_t normalize, float* out, int32_t device, void* stream) {
    if (B < 0 || d <= 0 || (B > 0 && (!hidden || !out))) {
        set_error("pool_normalize: bad arguments");
        return B2F_EINVAL;
    }
    B2F_TRY(check_device(device));
    DeviceGuard g(device);
    return launch_pool(hidden, mask, B, T, d, pool, normalize, out, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int b2f_synth_rows(uint64_t seed, int64_t row0, int64_t nrows, int32_t d, int32_t normalize, float* out, int32_t device,
                   void* stream) {
    if (nrows < 0 || d <= 0 || (nrows > 0 && !out)) {
        set_error("synth_rows: bad arguments");
        return B2F_EINVAL;
    }
    B2F_TRY(check_device(device));
    DeviceGuard g(device);
    return launch_synth(seed, row0, nrows, d, normalize, out, (cudaStream_t)stream);
}

int b2f_merge_topk_strided(int32_t metric, int64_t nq, int64_t k, int32_t nparts, const float* D_parts, const int64_t* I_parts,
                           int64_t part_stride_bytes, float* D, int64_t* I, int32_t device, void* stream) {
    if (nq < 0 || k <= 0 || !D_parts || !I_parts || !D || !I || part_stride_bytes <= 0 || (part_stride_bytes & 7)) {
        set_error("merge_topk: bad arguments");
        return B2F_EINVAL;
    }
    B2F_TRY(check_device(device));
    DeviceGuard g(device);
    return launch_merge_faiss(metric, nq, k, nparts, D_parts, I_parts, part_stride_bytes, part_stride_bytes, D, I, (cudaStream_t)stream);
}

int b2f_merge_topk(int32_t metric, int64_t nq, int64_t k, int32_t nparts, const float* D_parts, const int64_t* I_parts,
                   float* D, int64_t* I, int32_t device, void* stream) {
    if (nq < 0 || k <= 0 || !D_parts || !I_parts || !D || !I) {
        set_error("merge_topk: bad arguments");
        return B2F_EINVAL;
    }
    B2F_TRY(check_device(device));
    DeviceGuard g(device);
    return launch_merge_faiss(metric, nq, k, nparts, D_parts, I_parts, nq * k * 4, nq * k * 8, D, I, (cudaStream_t)stream);
}

}  // extern "C"

// ---- range pass (second tensor pass over uncertified queries) -----------------------------------------
constexpr int kRangeCap = 1024;   // failed queries one range pass holds; more go to the exact scan
constexpr int kRangeListCap = 256;

static int plan_range_pass(int nfail, int64_t n, int d, int kp, TensorScanPlan* plan) {
    if (plan_tensor_scan(nfail, n, d, kp, plan) != B2F_OK || !plan->list_mode) return B2F_EINVAL;
    if (plan->list_cap < kRangeListCap) plan->list_cap = kRangeListCap;   // lists are not pruned: every row at or below the threshold
    return B2F_OK;
}
static size_t range_lists_bytes(const TensorScanPlan& p) {
    const size_t nq_pad = (size_t)p.nq_tiles * 128;
    return 3 * align_up(nq_pad * p.nlists * 4, 256) + align_up(nq_pad * p.nlists * (size_t)p.list_cap * 8, 256) + align_up(nq_pad, 256);
}
static size_t range_pass_bytes(const b2f_index* ix, int nq, int kp) {
    size_t lists = 0;
    for (int f = 128; f <= kRangeCap; f *= 2) {   // the list geometry depends on the number of failed queries: take the largest
        TensorScanPlan p{};
        if (plan_range_pass(f, ix->ntotal, ix->d, kp, &p) == B2F_OK && range_lists_bytes(p) > lists) lists = range_lists_bytes(p);
    }
    const size_t cap = kRangeCap;
    return align_up(cap * ix->d * 4, 256) + align_up(cap * ix->dpad * 2, 256) + 7 * align_up(cap * 4, 256) +
           2 * align_up((size_t)nq * 4, 256) + 8192 + lists;
}

// ---- search ----------------------------------------------------------------------------------------
// The caller holds ix->mu (search_pooled pools the queries into the index's buffer and searches them under ONE
// lock, so a second thread cannot overwrite or re-allocate that buffer in between).
static int search_locked(b2f_index* ix, int64_t nq64, const float* q, int64_t k64, float* D, int64_t* I, int32_t mem,
                         void* stream, const b2f_search_params* params) {
    if (k64 <= 0) {
        set_error("k must be > 0 (faiss asserts k > 0)");
        return B2F_EINVAL;
    }
    if (k64 > 1024) {
        set_error("k=%lld > 1024 is not supported by the GPU selection kernels", (long long)k64);
        return B2F_EINVAL;
    }
    if (nq64 < 0 || nq64 > (1 << 24) || (nq64 > 0 && (!q || !D || !I))) {
        set_error("search: bad arguments");
        return B2F_EINVAL;
    }
    if (nq64 == 0) return B2F_OK;
    const int nq = (int)nq64, k = (int)k64;
    b2f_search_params P{};
    if (params) P = *params;
    DeviceGuard g(ix->device);
    if (!g.ok) {
        set_error("cudaSetDevice(%d) failed", ix->device);
        return B2F_ENOGPU;
    }
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    B2F_TRY(order_after(st, ix->stream, ix));
    B2F_TRY(wait_ingest(ix, st));   // the latest add may sit on another stream (torch's, behind an encoder forward)
    // searches return without synchronising: a search on another stream must not reuse the workspace early
    if (ix->search_recorded && ix->last_stream != st) B2F_CUDA(cudaStreamWaitEvent(st, ix->ev_done, 0));
    const bool host = mem == B2F_MEM_HOST;
    const bool profile = P.profile != 0;
    harvest_flag(ix);

    int algo = P.algo;
    // AUTO: fp32 storage serves the reference's own batch size (nq = 1) with the fp32 streaming scan -- the
    // reference's arithmetic, 90% of HBM on the fp32 rows.  bf16 storage has nothing to gain from it (both paths
    // read the same bf16 rows; the tensor path streams them at the full HBM rate, the CUDA-core scan at half
    // of it because of the unpacking), so every batch size goes to the tensor path there.
    const int scan_max = P.scan_max_nq > 0 ? P.scan_max_nq : (ix->storage == B2F_STORE_BF16 ? 0 : 1);
    int kp = tensor_kprime(k, P.slack);
    {
        // diagnostics: B200FLAT_KPRIME=<multiple of 8> forces k' (LIST-mode shapes; same-box A/B of the candidate count)
        static int kp_env = -1;
        if (kp_env < 0) {
            const char* e = getenv("B200FLAT_KPRIME");
            kp_env = e ? atoi(e) : 0;
            if (kp_env < 0 || kp_env > 256 || (kp_env & 7)) kp_env = 0;
        }
        if (kp_env >= k + 2 && kp > 0) kp = kp_env;
    }
    if (kp > 0 && P.slack <= 0 && ix->slack_boost > 0) {  // adaptive slack learned from earlier searches on this index
        int boosted = ((kp + ix->slack_boost + 7) / 8) * 8;
        if (boosted > 64) kp = boosted <= 256 ? boosted : 256;
    }
    TensorScanPlan plan{};
    const bool heap_first = ix->heap_left > 0 && P.certify >= 0;
    plan_force_heap(heap_first);
    const int chunk_nq = (kp > 0 && ix->ntotal > 0) ? plan_tensor_chunked(nq, ix->ntotal, ix->d, kp, &plan) : 0;
    plan_force_heap(false);
    const bool tensor_ok = chunk_nq > 0;
    if (algo == B2F_ALGO_AUTO) algo = (nq <= scan_max || !tensor_ok) ? B2F_ALGO_SCAN : B2F_ALGO_TENSOR;
    // an explicit TENSOR request on a shape the tensor path cannot plan (k' > 256, or a database of a few rows with
    // k' other than 32 / 64) is served by the exact scan, like AUTO would; stats().last_algo tells
    if (algo == B2F_ALGO_TENSOR && !tensor_ok) algo = B2F_ALGO_SCAN;
    const int certify = P.certify >= 0 ? 1 : 0;
    if (algo == B2F_ALGO_TENSOR && heap_first) ix->heap_left--;

    // ---- workspace -----------------------------------------------------------------------------
    size_t need = 8192;
    if (host) need += align_up((size_t)nq * ix->d * 4, 256) + align_up((size_t)nq * k * 4, 256) + align_up((size_t)nq * k * 8, 256);
    need += align_up(scan_scratch_bytes(k), 256) + 256;
    if (algo == B2F_ALGO_TENSOR) {
        const size_t nq_pad = (size_t)plan.nq_tiles * 128;
        need += align_up(nq_pad * ix->dpad * 2, 256) + 3 * align_up(nq_pad * 4, 256);
        if (plan.list_mode) {
            need += align_up(nq_pad * plan.nlists * 4, 256) * 3;                       // shared thresholds + counts + final thresholds
            need += align_up(nq_pad * plan.nlists * (size_t)plan.list_cap * 8, 256);   // candidate lists
            need += align_up(nq_pad, 256);                                              // warp-merge pass-on flags
        } else {
            need += 2 * align_up(nq_pad * plan.nsplits * kp * 4, 256);  // partial lists
            need += 2 * align_up((size_t)chunk_nq * kp * 4, 256);       // merged coarse
        }
        need += align_up((size_t)nq * 4, 256) + 256;                    // fail list + counters
        if (ix->range_mode) need += range_pass_bytes(ix, nq, kp);
    }
    B2F_TRY(ensure_ws(ix, need));
    Bump bump(ix->ws);

    const float* qd = q;
    float* Dd = D;
    int64_t* Id = I;
    if (host) {
        float* qbuf = bump.take<float>((size_t)nq * ix->d);
        Dd = bump.take<float>((size_t)nq * k);
        Id = bump.take<int64_t>((size_t)nq * k);
        B2F_CUDA(cudaMemcpyAsync(qbuf, q, (size_t)nq * ix->d * 4, cudaMemcpyHostToDevice, st));
        qd = qbuf;
    }
    void* scan_scratch = bump.take<char>(scan_scratch_bytes(k));
    ix->st.searches++;
    ix->st.last_launches = 0;
    ix->st.last_algo = algo;
    ix->st.last_kprime = 0;
    int n_main = 0;
    b2f_index::ProfSlot* slot = nullptr;
    if (profile) {
        slot = &ix->prof[ix->prof_head % b2f_index::kProfSlots];
        ix->prof_head++;
        harvest_prof_slot(ix, *slot);  // blocks only when the host is a whole ring of searches ahead
        if (!slot->t0) B2F_CUDA(cudaEventCreate(&slot->t0));
        if (!slot->t1) B2F_CUDA(cudaEventCreate(&slot->t1));
        B2F_CUDA(cudaEventRecord(slot->t0, st));
    }
    auto main_event = [&](int which) -> int {  // which: 0 = before, 1 = after the dominant kernel's launch
        if (!slot) return B2F_OK;
        cudaEvent_t e = prof_main_event(*slot, 2 * (size_t)n_main + which);
        if (!e) {
            set_error("cudaEventCreate failed");
            return B2F_ECUDA;
        }
        B2F_CUDA(cudaEventRecord(e, st));
        return B2F_OK;
    };

    if (ix->ntotal == 0) {
        B2F_TRY(launch_finalize(nullptr, nullptr, nq, 0, k, ix->metric, P.id_offset, nullptr, Dd, Id, st));
        ix->st.launches++;
        ix->st.last_launches++;
    } else if (algo == B2F_ALGO_SCAN) {
        B2F_TRY(main_event(0));
        B2F_TRY(enqueue_scan(ix, qd, nullptr, nullptr, nq, k, Dd, Id, P.id_offset, scan_scratch, nullptr, 0, nq, 0, st));
        B2F_TRY(main_event(1));
        n_main++;
    } else {
        const int nq_pad = plan.nq_tiles * 128;
        ix->st.last_kprime = kp;
        __nv_bfloat16* qb = bump.take<__nv_bfloat16>((size_t)nq_pad * ix->dpad);
        float* qnorm = bump.take<float>(nq_pad);
        float* qerr = bump.take<float>(nq_pad);
        float* qconst = bump.take<float>(nq_pad);
        float* pk = nullptr;
        int32_t* pi = nullptr;
        float* ck = nullptr;
        int32_t* ci = nullptr;
        TensorScanLists lists{};
        if (plan.list_mode) {
            lists.shared_thr = bump.take<float>((size_t)nq_pad * plan.nlists);
            lists.counts = bump.take<int32_t>((size_t)nq_pad * plan.nlists);
            lists.final_thr = bump.take<float>((size_t)nq_pad * plan.nlists);
            lists.cand = bump.take<uint2>((size_t)nq_pad * plan.nlists * plan.list_cap);
            lists.big_flag = bump.take<uint8_t>((size_t)nq_pad);
        } else {
            pk = bump.take<float>((size_t)nq_pad * plan.nsplits * kp);
            pi = bump.take<int32_t>((size_t)nq_pad * plan.nsplits * kp);
            ck = bump.take<float>((size_t)chunk_nq * kp);
            ci = bump.take<int32_t>((size_t)chunk_nq * kp);
        }
        int32_t* fail_list = bump.take<int32_t>(nq);
        float* fail_tau = ix->range_mode ? bump.take<float>(nq) : nullptr;
        // [0] uncertified queries, [1] of which list overflows, [2..3] u64 list entries, [4] rescued by the extended pass
        int32_t* counters = bump.take<int32_t>(16);   // [8..11]: ticket counters of K2's die-aware unit assignment
        lists.die_ctr = counters + 8;
        B2F_TRY(refresh_host_stats(ix, st));
        for (int c0 = 0; c0 < nq; c0 += chunk_nq) {
            const int cn = nq - c0 < chunk_nq ? nq - c0 : chunk_nq;  // queries in this pass (the plan covers chunk_nq)
            const float* qc = qd + (int64_t)c0 * ix->d;
            // one launch: bf16 copy / norms of the queries, reset the shared thresholds, clear the counters (first pass)
            B2F_TRY(launch_prep_queries(qc, cn, nq_pad, ix->d, qb, ix->dpad, qnorm, qerr, qconst, ix->mu_set ? ix->centre : nullptr,
                                        reinterpret_cast<uint32_t*>(counters),
                                        c0 == 0 ? 16 : 0, reinterpret_cast<uint32_t*>(lists.shared_thr),
                                        plan.list_mode ? (int64_t)nq_pad * plan.nlists : 0, st));
            B2F_TRY(main_event(0));
            B2F_TRY(launch_tensor_scan(ix->scan, ix->dpad, ix->norms, ix->ntotal, ix->metric, qb, cn, nq_pad, plan, pk, pi, lists, st));
            B2F_TRY(main_event(1));
            n_main++;
            RerankArgs ra{};
            ra.rows_f32 = ix->storage == B2F_STORE_F32 ? ix->rows_f32 : nullptr;
            ra.rows_bf16 = ix->scan;
            ra.pitch_bf16 = ix->dpad;
            ra.centre = (ix->storage == B2F_STORE_BF16 && ix->mu_set) ? ix->centre : nullptr;
            ra.q = qc;
            ra.qnorm = qnorm;
            ra.qerr = qerr;
            ra.qconst = qconst;
            ra.mu_norm = ix->mu_norm;
            ra.cand_key = ck;
            ra.cand_id = ci;
            ra.nq = cn;
            ra.kp = kp;
            ra.k = k;
            ra.d = ix->d;
            ra.metric = ix->metric;
            ra.ntotal = ix->ntotal;
            ra.max_row_norm = sqrtf(ix->host_stats[0]);
            // fp32 storage: the bf16 rounding of the centred row; bf16 storage: what fl32(mu + x~') loses (0 uncentred)
            ra.max_row_err = sqrtf(ix->host_stats[1]);
            ra.certify = certify;
            ra.D = Dd + (int64_t)c0 * k;  // the re-rank writes faiss-formatted results directly (no finalize launch)
            ra.I = Id + (int64_t)c0 * k;
            ra.id_offset = P.id_offset;
            ra.fail_list = fail_list;
            ra.fail_tau = fail_tau;
            ra.fail_count = counters;
            ra.q_base = c0;
            {   // First-stage re-rank (the best n candidates alone before the k' best): OFF.  Same-box A/B (r02n): no gain at
                // 4096 / 8192 queries with k = 10 (the merge is not bound by the candidate row reads after all), and a loss
                // at k = 100, k' = 192 (0.69 -> 0.93 ms per C3 search).  B200FLAT_STAGE1=<n> enables it for experiments.
                static int s1_env = -2;
                if (s1_env == -2) {
                    const char* e = getenv("B200FLAT_STAGE1");
                    s1_env = e ? atoi(e) : 0;
                }
                ra.stage1 = (s1_env >= k && s1_env < kp) ? s1_env : 0;
            }
            if (plan.list_mode) {
                // K3b + K4 + finalize fused: per query, merge the lists, re-rank exactly, certify, write (D, I)
                B2F_TRY(launch_merge_lists(lists, cn, plan, reinterpret_cast<unsigned long long*>(counters + 2), ra, st));
                ix->st.launches += 3;
                ix->st.last_launches += 3;
            } else {
                B2F_TRY(launch_merge_parts(pk, pi, cn, plan.nsplits, kp, kp, ck, ci, st));
                B2F_TRY(launch_rerank(ra, st));
                ix->st.launches += 4;
                ix->st.last_launches += 4;
            }
        }  // query chunks
        int32_t* scan_list = fail_list;
        int32_t* scan_counters = counters;
        int range_served = 0;
        if (ix->range_mode && certify && fail_tau) {
            // Range pass.  Earlier batches on this index left many queries uncertified (data denser than the bf16 band even
            // after centring): instead of one exact scan per four of them, ONE more tensor pass serves them all.  For a
            // failed query the first pass found an exact k-th key tau, an upper bound of the true one; every true top-k row
            // therefore has a coarse key <= thr(tau) (the certification bound, inverted), so a pass that lists EVERY row at
            // or below that fixed threshold and re-ranks all of them is exact.  The host must know how many queries failed
            // to plan the pass: the one synchronisation inside a search, paid only by indexes in this mode.
            volatile int32_t* hf = ix->host_flag + 12;
            B2F_CUDA(cudaMemcpyAsync(const_cast<int32_t*>(hf), counters, 4, cudaMemcpyDeviceToHost, st));
            B2F_CUDA(cudaStreamSynchronize(st));
            const int nfail = hf[0];
            if (nfail == 0) {
                if (++ix->range_quiet >= 8) ix->range_mode = false;   // the data (or the adaptive slack) no longer needs it
            } else {
                ix->range_quiet = 0;
            }
            TensorScanPlan rp{};
            const int nr = nfail < kRangeCap ? nfail : kRangeCap;
            if (nfail >= 8 && plan_range_pass(nr, ix->ntotal, ix->d, kp, &rp) == B2F_OK) {
                const int nq_pad2 = rp.nq_tiles * 128;
                float* qf = bump.take<float>((size_t)kRangeCap * ix->d);
                __nv_bfloat16* qb2 = bump.take<__nv_bfloat16>((size_t)kRangeCap * ix->dpad);
                float* qnorm2 = bump.take<float>(kRangeCap);
                float* qerr2 = bump.take<float>(kRangeCap);
                float* qconst2 = bump.take<float>(kRangeCap);
                float* tau2 = bump.take<float>(kRangeCap);
                float* thr2 = bump.take<float>(kRangeCap);
                int32_t* out_map = bump.take<int32_t>(kRangeCap);
                int32_t* fail_list2 = bump.take<int32_t>(nq);
                int32_t* counters_b = bump.take<int32_t>(32);   // [0..7] as `counters`, [8..11] K2's tickets, [16] compacted queries
                TensorScanLists l2{};
                l2.shared_thr = bump.take<float>((size_t)nq_pad2 * rp.nlists);
                l2.counts = bump.take<int32_t>((size_t)nq_pad2 * rp.nlists);
                l2.final_thr = bump.take<float>((size_t)nq_pad2 * rp.nlists);
                l2.cand = bump.take<uint2>((size_t)nq_pad2 * rp.nlists * rp.list_cap);
                l2.big_flag = bump.take<uint8_t>((size_t)nq_pad2);
                l2.die_ctr = counters_b + 8;
                l2.range_thr = thr2;
                B2F_CUDA(cudaMemsetAsync(counters_b, 0, 32 * 4, st));
                B2F_CUDA(cudaMemsetAsync(out_map, 0xff, (size_t)kRangeCap * 4, st));
                B2F_CUDA(cudaMemsetAsync(qf, 0, (size_t)nq_pad2 * ix->d * 4, st));
                B2F_TRY(launch_gather_failed(qd, ix->d, fail_list, fail_tau, counters, nfail, nr, qf, tau2, out_map, counters_b + 16,
                                             fail_list2, counters_b, st));
                B2F_TRY(launch_prep_queries(qf, nr, nq_pad2, ix->d, qb2, ix->dpad, qnorm2, qerr2, qconst2, ix->mu_set ? ix->centre : nullptr,
                                            nullptr, 0, reinterpret_cast<uint32_t*>(l2.shared_thr), (int64_t)nq_pad2 * rp.nlists, st));
                B2F_TRY(launch_range_thresholds(tau2, out_map, nr, ix->d, ix->metric, qnorm2, qerr2, qconst2, sqrtf(ix->host_stats[0]),
                                                sqrtf(ix->host_stats[1]), ix->mu_norm, thr2, st));
                B2F_TRY(launch_tensor_scan(ix->scan, ix->dpad, ix->norms, ix->ntotal, ix->metric, qb2, nr, nq_pad2, rp, nullptr, nullptr, l2, st));
                RerankArgs r2{};
                r2.rows_f32 = ix->storage == B2F_STORE_F32 ? ix->rows_f32 : nullptr;
                r2.rows_bf16 = ix->scan;
                r2.pitch_bf16 = ix->dpad;
                r2.centre = (ix->storage == B2F_STORE_BF16 && ix->mu_set) ? ix->centre : nullptr;
                r2.q = qf;
                r2.qnorm = qnorm2;
                r2.qerr = qerr2;
                r2.qconst = qconst2;
                r2.mu_norm = ix->mu_norm;
                r2.nq = nr;
                r2.kp = kp;
                r2.k = k;
                r2.d = ix->d;
                r2.metric = ix->metric;
                r2.ntotal = ix->ntotal;
                r2.max_row_norm = sqrtf(ix->host_stats[0]);
                r2.max_row_err = sqrtf(ix->host_stats[1]);
                r2.certify = 1;
                r2.D = Dd;
                r2.I = Id;
                r2.id_offset = P.id_offset;
                r2.fail_list = fail_list2;
                r2.fail_count = counters_b;
                r2.out_map = out_map;
                r2.range = 1;
                B2F_TRY(launch_merge_lists(l2, nr, rp, nullptr, r2, st));
                // the statistics the closing kernel publishes: what the first pass counted, with the final fallback count
                B2F_CUDA(cudaMemcpyAsync(counters_b + 1, counters + 1, 7 * 4, cudaMemcpyDeviceToDevice, st));
                ix->st.launches += 5;
                ix->st.last_launches += 5;
                ix->st.range_queries += nr;
                range_served = nr;
                scan_list = fail_list2;
                scan_counters = counters_b;
            }
        }
        // Closing kernel: the exact scan over the queries that could not be certified.  The count lives on the
        // device -- with none (the usual case) the kernel publishes the counters and exits -- so the host never
        // waits inside a search and consecutive searches run back to back on the GPU.
        B2F_TRY(enqueue_scan(ix, qd, scan_list, scan_counters, 0, k, Dd, Id, P.id_offset, scan_scratch, scan_counters, ++ix->seq, nq,
                             certify, st));
        if (range_served) {
            ix->range_ran_seq = ix->seq;
            ix->range_ran_nfail = range_served;
        }
    }
    if (slot) {
        B2F_CUDA(cudaEventRecord(slot->t1, st));
        slot->n_main = n_main;
        slot->pending = true;
    }
    if (host) {
        B2F_CUDA(cudaMemcpyAsync(D, Dd, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
        B2F_CUDA(cudaMemcpyAsync(I, Id, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
    }
    if (!ix->ev_done) B2F_CUDA(cudaEventCreateWithFlags(&ix->ev_done, cudaEventDisableTiming));
    B2F_CUDA(cudaEventRecord(ix->ev_done, st));
    ix->last_stream = st;
    ix->search_recorded = true;
    if (host) {
        B2F_CUDA(cudaStreamSynchronize(st));  // faiss semantics: results are in the caller's arrays on return
        harvest_flag(ix);
    }
    return B2F_OK;
}

extern "C" {

int b2f_index_search(b2f_index* ix, int64_t nq, const float* q, int64_t k, float* D, int64_t* I, int32_t mem, void* stream,
                     const b2f_search_params* params) {
    if (!ix) {
        set_error("index is NULL");
        return B2F_EINVAL;
    }
    std::lock_guard<std::mutex> lk(ix->mu);
    return search_locked(ix, nq, q, k, D, I, mem, stream, params);
}

int b2f_index_search_pooled(b2f_index* ix, const float* hidden, const int64_t* mask, int64_t B, int64_t T, int32_t pool,
                            int32_t normalize, int64_t k, float* D, int64_t* I, void* stream, const b2f_search_params* params) {
    if (!ix || B < 0 || (B > 0 && (!hidden || !D || !I))) {
        set_error("search_pooled: bad arguments");
        return B2F_EINVAL;
    }
    if (B == 0) return B2F_OK;
    float* qbuf = nullptr;
    std::lock_guard<std::mutex> lk(ix->mu);   // held across pooling AND the search: the pooled queries live in ix->qpool
    {
        DeviceGuard g(ix->device);
        if (!g.ok) {
            set_error("cudaSetDevice(%d) failed", ix->device);
            return B2F_ENOGPU;
        }
        // pooled queries live in their own grow-only buffer (the search below carves the workspace itself)
        const size_t need = (size_t)B * ix->d * sizeof(float);
        if (need > ix->qpool_bytes) {
            if (ix->search_recorded) B2F_CUDA(cudaEventSynchronize(ix->ev_done));
            cudaFree(ix->qpool);
            ix->qpool = nullptr;
            ix->qpool_bytes = 0;
            if (cudaMalloc(&ix->qpool, need + (need >> 2)) != cudaSuccess) {
                cudaGetLastError();
                set_error("search_pooled: cudaMalloc(%zu) failed", need);
                return B2F_ENOMEM;
            }
            ix->qpool_bytes = need + (need >> 2);
        }
        qbuf = ix->qpool;
        cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
        if (ix->search_recorded && ix->last_stream != st) B2F_CUDA(cudaStreamWaitEvent(st, ix->ev_done, 0));
        B2F_TRY(launch_pool(hidden, mask, B, T, ix->d, pool, normalize, qbuf, nullptr, 0, nullptr, nullptr, st));
        ix->st.launches++;
    }
    return search_locked(ix, B, qbuf, k, D, I, B2F_MEM_DEVICE, stream, params);
}

int b2f_index_reconstruct(b2f_index* ix, int64_t i0, int64_t n, float* out, int32_t mem, void* stream) {
    if (!ix || !out || n < 0) {
        set_error("reconstruct: bad arguments");
        return B2F_EINVAL;
    }
    if (i0 < 0 || i0 + n > ix->ntotal) {
        set_error("reconstruct: rows [%lld, %lld) out of range [0, %lld)", (long long)i0, (long long)(i0 + n), (long long)ix->ntotal);
        return B2F_ERANGE;
    }
    if (n == 0) return B2F_OK;
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    B2F_TRY(order_after(st, ix->stream, ix));
    B2F_TRY(wait_ingest(ix, st));
    if (ix->search_recorded && ix->last_stream != st) B2F_CUDA(cudaStreamWaitEvent(st, ix->ev_done, 0));  // shares the workspace
    const size_t bytes = (size_t)n * ix->d * 4;
    if (ix->storage == B2F_STORE_F32) {
        B2F_CUDA(cudaMemcpyAsync(out, ix->rows_f32 + i0 * ix->d, bytes, mem == B2F_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
    } else if (mem == B2F_MEM_DEVICE) {
        B2F_TRY(launch_bf16_to_f32(ix->scan + i0 * ix->dpad, ix->dpad, n, ix->d, ix->mu_set ? ix->centre : nullptr, out, st));
        ix->st.launches++;
    } else {
        B2F_TRY(ensure_ws(ix, bytes + 4096));
        B2F_TRY(launch_bf16_to_f32(ix->scan + i0 * ix->dpad, ix->dpad, n, ix->d, ix->mu_set ? ix->centre : nullptr, reinterpret_cast<float*>(ix->ws), st));
        ix->st.launches++;
        B2F_CUDA(cudaMemcpyAsync(out, ix->ws, bytes, cudaMemcpyDeviceToHost, st));
    }
    if (mem == B2F_MEM_HOST) B2F_CUDA(cudaStreamSynchronize(st));
    return B2F_OK;
}

// ---- file I/O (FAISS IndexFlat layout; see include/b200flat.h) --------------------------------------
static const size_t kIoChunk = 32u << 20;

int b2f_index_write(b2f_index* ix, const char* path) {
    if (!ix || !path) {
        set_error("write: bad arguments");
        return B2F_EINVAL;
    }
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    B2F_TRY(wait_ingest(ix, ix->stream));
    if (ix->search_recorded && ix->last_stream != ix->stream) B2F_CUDA(cudaStreamWaitEvent(ix->stream, ix->ev_done, 0));
    FILE* f = fopen(path, "wb");
    if (!f) {
        set_error("could not open %s for writing", path);
        return B2F_EIO;
    }
    const char* cc = ix->metric == B2F_METRIC_L2 ? "IxF2" : "IxFI";
    const int32_t d32 = ix->d, mt = ix->metric;
    const int64_t nt = ix->ntotal, dummy = 1 << 20;
    const uint8_t tr = 1;
    const uint64_t sz = (uint64_t)ix->ntotal * ix->d;
    bool ok = fwrite(cc, 1, 4, f) == 4 && fwrite(&d32, 4, 1, f) == 1 && fwrite(&nt, 8, 1, f) == 1 &&
              fwrite(&dummy, 8, 1, f) == 1 && fwrite(&dummy, 8, 1, f) == 1 && fwrite(&tr, 1, 1, f) == 1 &&
              fwrite(&mt, 4, 1, f) == 1 && fwrite(&sz, 8, 1, f) == 1;
    int rc = B2F_OK;
    if (ok && ix->ntotal > 0) {
        const int64_t rows_per = (int64_t)(kIoChunk / ((size_t)ix->d * 4)) > 0 ? (int64_t)(kIoChunk / ((size_t)ix->d * 4)) : 1;
        const size_t cbytes = (size_t)rows_per * ix->d * 4;
        rc = ensure_pinned(ix, 2 * cbytes);
        if (rc == B2F_OK && ix->storage == B2F_STORE_BF16) rc = ensure_ws(ix, 2 * cbytes + 4096);
        cudaEvent_t evs[2] = {get_event(ix, 1), get_event(ix, 2)};
        int64_t pending_rows[2] = {0, 0};
        int slot = 0;
        // double-buffered: D2H of chunk i+1 overlaps fwrite of chunk i
        for (int64_t r0 = 0; rc == B2F_OK && ok && r0 < ix->ntotal + rows_per; r0 += rows_per, slot ^= 1) {
            if (r0 < ix->ntotal) {
                const int64_t m = ix->ntotal - r0 < rows_per ? ix->ntotal - r0 : rows_per;
                char* hbuf = ix->pinned + (size_t)slot * cbytes;
                cudaError_t e;
                if (ix->storage == B2F_STORE_F32) {
                    e = cudaMemcpyAsync(hbuf, ix->rows_f32 + r0 * ix->d, (size_t)m * ix->d * 4, cudaMemcpyDeviceToHost, ix->stream);
                } else {
                    float* dbuf = reinterpret_cast<float*>(ix->ws + (size_t)slot * cbytes);
                    rc = launch_bf16_to_f32(ix->scan + r0 * ix->dpad, ix->dpad, m, ix->d, ix->mu_set ? ix->centre : nullptr, dbuf, ix->stream);
                    ix->st.launches++;
                    e = cudaMemcpyAsync(hbuf, dbuf, (size_t)m * ix->d * 4, cudaMemcpyDeviceToHost, ix->stream);
                }
                if (e == cudaSuccess) e = cudaEventRecord(evs[slot], ix->stream);
                if (e != cudaSuccess) {
                    set_error("write: device read failed: %s", cudaGetErrorString(e));
                    rc = B2F_ECUDA;
                }
                pending_rows[slot] = m;
            } else {
                pending_rows[slot] = 0;
            }
            const int prev = slot ^ 1;
            if (rc == B2F_OK && r0 > 0 && pending_rows[prev] > 0) {
                if (cudaEventSynchronize(evs[prev]) != cudaSuccess) {
                    set_error("write: event sync failed");
                    rc = B2F_ECUDA;
                } else {
                    const size_t cnt = (size_t)pending_rows[prev] * ix->d;
                    ok = fwrite(ix->pinned + (size_t)prev * cbytes, 4, cnt, f) == cnt;
                    pending_rows[prev] = 0;
                }
            }
        }
    }
    if (fclose(f) != 0) ok = false;
    if (rc != B2F_OK) return rc;
    if (!ok) {
        set_error("short write to %s", path);
        return B2F_EIO;
    }
    return B2F_OK;
}

int b2f_index_read(const char* path, int32_t storage, int32_t device, b2f_index** out) {
    if (!path || !out) {
        set_error("read: bad arguments");
        return B2F_EINVAL;
    }
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) {
        set_error("could not open %s for reading", path);
        return B2F_EIO;
    }
    char cc[4];
    int32_t d32 = 0, mt = 0;
    int64_t nt = 0, dm[2];
    uint8_t tr = 0;
    uint64_t sz = 0;
    bool ok = fread(cc, 1, 4, f) == 4;
    if (ok && memcmp(cc, "IxF2", 4) != 0 && memcmp(cc, "IxFI", 4) != 0 && memcmp(cc, "IxFl", 4) != 0) {
        fclose(f);
        set_error("%s is not an IndexFlat file (fourcc %.4s)", path, cc);
        return B2F_EFORMAT;
    }
    ok = ok && fread(&d32, 4, 1, f) == 1 && fread(&nt, 8, 1, f) == 1 && fread(dm, 8, 2, f) == 2 &&
         fread(&tr, 1, 1, f) == 1 && fread(&mt, 4, 1, f) == 1;
    if (ok && mt > 1) {
        float arg;
        ok = fread(&arg, 4, 1, f) == 1;
    }
    ok = ok && fread(&sz, 8, 1, f) == 1;
    if (!ok || d32 <= 0 || nt < 0 || sz != (uint64_t)nt * (uint64_t)d32) {
        fclose(f);
        set_error("%s: truncated or inconsistent IndexFlat header", path);
        return B2F_EFORMAT;
    }
    if (mt != B2F_METRIC_L2 && mt != B2F_METRIC_INNER_PRODUCT) {
        fclose(f);
        set_error("%s: metric_type %d is not supported (only L2 and inner product)", path, mt);
        return B2F_EFORMAT;
    }
    b2f_index* ix = nullptr;
    int rc = b2f_index_create(d32, mt, storage, device, &ix);
    if (rc != B2F_OK) {
        fclose(f);
        return rc;
    }
    {
        DeviceGuard g(device);
        rc = ensure_capacity(ix, nt);
        const int64_t rows_per = (int64_t)(kIoChunk / ((size_t)d32 * 4)) > 0 ? (int64_t)(kIoChunk / ((size_t)d32 * 4)) : 1;
        const size_t cbytes = (size_t)rows_per * d32 * 4;
        if (rc == B2F_OK && nt > 0) rc = ensure_pinned(ix, 2 * cbytes);
        if (rc == B2F_OK && nt > 0 && storage == B2F_STORE_BF16) rc = ensure_ws(ix, 2 * cbytes + 4096);
        cudaEvent_t evs[2] = {get_event(ix, 1), get_event(ix, 2)};
        bool used[2] = {false, false};
        int slot = 0;
        // double-buffered: fread of chunk i+1 overlaps H2D + ingest of chunk i
        for (int64_t r0 = 0; rc == B2F_OK && r0 < nt; r0 += rows_per, slot ^= 1) {
            const int64_t m = nt - r0 < rows_per ? nt - r0 : rows_per;
            char* hbuf = ix->pinned + (size_t)slot * cbytes;
            if (used[slot] && cudaEventSynchronize(evs[slot]) != cudaSuccess) {
                set_error("read: event sync failed");
                rc = B2F_ECUDA;
                break;
            }
            const size_t cnt = (size_t)m * d32;
            if (fread(hbuf, 4, cnt, f) != cnt) {
                set_error("%s: truncated payload", path);
                rc = B2F_EFORMAT;
                break;
            }
            cudaError_t e;
            const float* src;
            if (storage == B2F_STORE_F32) {
                float* dst = ix->rows_f32 + ix->ntotal * d32;
                e = cudaMemcpyAsync(dst, hbuf, cnt * 4, cudaMemcpyHostToDevice, ix->stream);
                src = dst;
            } else {
                float* dbuf = reinterpret_cast<float*>(ix->ws + (size_t)slot * cbytes);
                e = cudaMemcpyAsync(dbuf, hbuf, cnt * 4, cudaMemcpyHostToDevice, ix->stream);
                src = dbuf;
            }
            if (e != cudaSuccess) {
                set_error("read: H2D failed: %s", cudaGetErrorString(e));
                rc = B2F_ECUDA;
                break;
            }
            rc = add_device_rows(ix, src, m, ix->stream);
            if (rc != B2F_OK) break;
            cudaEventRecord(evs[slot], ix->stream);
            used[slot] = true;
            ix->ntotal += m;
        }
        if (rc == B2F_OK && cudaStreamSynchronize(ix->stream) != cudaSuccess) {
            set_error("read: stream sync failed");
            rc = B2F_ECUDA;
        }
        ix->stats_dirty = true;
    }
    fclose(f);
    if (rc != B2F_OK) {
        b2f_index_destroy(ix);
        return rc;
    }
    *out = ix;
    return B2F_OK;
}

}  // extern "C"
