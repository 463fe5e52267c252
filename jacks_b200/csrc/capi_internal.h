// Internals shared by the C-ABI translation units (not part of the public header).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>

namespace jacks {

// Records the message jacks_last_error() returns and passes `code` through.
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

// Grow-only device buffer (avoids cudaMalloc/cudaFree on every call).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
};

}  // namespace jacks

#include <vector>

// Device buffers of one batched-EM plan.  Plans come and go with every host call; their buffers are recycled
// through JacksCtx::plan_pool instead of cudaMalloc/cudaFree (both synchronise the device).
struct PlanBufs {
    jacks::DevBuf d_offsets, d_rows, d_ids, d_slot_off, d_slot_rows, d_deferred, d_counters, d_scratch, d_tpd;
    jacks::DevBuf out[9];                   // device-side outputs of a host call that used this plan ([8]: p-values)
    long long sig_genes = -1, sig_guides = -1;   // shape of the last plan served (pool lookups prefer a match)
    // pinned staging of the index arrays: the GPU pulls them itself (copy kernel on the plan stream), so an index never
    // queues behind a screen-sized upload in the copy engine
    void* h_stage = nullptr; size_t h_cap = 0;
    cudaEvent_t ready = nullptr;          // the index arrays of the plan now served are on the device (plan stream)
    void release() {
        if (h_stage) cudaFreeHost(h_stage);
        h_stage = nullptr; h_cap = 0;
        if (ready) cudaEventDestroy(ready);
        ready = nullptr;
        for (auto& o : out) o.release();
        d_offsets.release(); d_rows.release(); d_ids.release(); d_slot_off.release(); d_slot_rows.release();
        d_deferred.release(); d_counters.release(); d_scratch.release(); d_tpd.release();
    }
};

struct JacksEmPlan;

// Staged copies between PAGEABLE host memory and the device.  cudaMemcpy on pageable memory moves 11 GB/s up and only
// 2 GB/s down into freshly allocated pages on this box (tools/pcie_probe.py; the page faults of a first touch are
// taken by one driver thread); here chunks pass through two pinned buffers while several host threads copy the
// previous chunk to / from the caller's memory (and take the first-touch faults in parallel).  Pinned host memory
// is recognised and copied directly.  Both calls return when the host side is complete (the device side of an
// upload may still be in flight on `st`).
struct HostStager {
    static constexpr size_t kChunk = 32u << 20;
    void* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int threads = 0;
    int init();
    void release();
    int h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st);
    int d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t st);
};

struct JacksCtx {
    int device;
    int n_sm;
    int64_t launches;
    cudaStream_t stream;        // compute + default copies of the host entry points
    cudaStream_t copy_stream;   // host -> device copies of the pipelined host entry points
    cudaStream_t out_stream;    // device -> host copies of the pipelined host entry points
    cudaStream_t out2_stream = nullptr;   // second download stream: the p-values of a screen job (they become ready long after
                                          // the other outputs of their range and must not hold those up in an in-order stream)
    cudaStream_t plan_stream = nullptr;   // uploads of gene indexes (jacks_em_plan_create)
    cudaStream_t cmp_stream = nullptr;    // row gathers of the compact cubes (kernels reading pinned host memory)
    cudaStream_t side_stream = nullptr;   // second kernel stream of jacks_run_screen_device: the pseudo-gene EM runs beside the
    cudaEvent_t side_fork = nullptr;      // real-gene EM, so that its CTAs fill the SMs the other kernel's tail leaves idle
    cudaEvent_t side_join = nullptr;
    cudaEvent_t res_busy = nullptr;       // the kernels that read the resident cubes (res_test / res_ctrl) have run
    // data cubes kept resident between host calls (jacks_em_batched_host keep_resident)
    jacks::DevBuf res_test, res_ctrl, res_fx;
    int64_t res_N = 0;
    int res_L = 0, res_ctrl_lines = 0;
    bool res_valid = false;
    jacks::DevBuf out[8];       // device-side outputs of the host entry points
    jacks::DevBuf tmp;          // scratch of the p-value / preprocessing entry points
    jacks::DevBuf pre;          // scratch of the preprocessing cores (sort keys, prefix sums)
    jacks::DevBuf kde_tail;     // queue of far-lower-tail evaluation points of the p-value kernels
    // compact cubes of JACKS_UPLOAD_USED_ROWS calls: pinned staging, device copy, "upload enqueued" event
    void* cmp_stage = nullptr; size_t cmp_stage_cap = 0;
    jacks::DevBuf cmp_dev, cmp_rows;
    cudaEvent_t cmp_ready = nullptr;     // the compact upload has read its pinned staging
    cudaEvent_t cmp_done = nullptr;      // the kernels of the last compact run have read the device cube
    bool pre_attr_set = false;  // the radix sort's shared-memory attribute has been set on this device
    bool kde_attr_set = false;  // the KDE's derivative table and kernel attributes have been set up on this device
    HostStager stager;          // pinned ring of the synchronous host entry points (pageable caller buffers)
    std::vector<PlanBufs*> plan_pool;          // recycled plan buffers
    std::vector<JacksEmPlan*> pending_plans;   // plans of asynchronous host calls, destroyed at the next synchronise
    std::vector<cudaEvent_t> pending_events;
};

// ------------------------------------------------------------------------------------------------
// Batched-EM plan (built by jacks_em_plan_create, capi.cu)
// ------------------------------------------------------------------------------------------------
#include "../../include/jacks_b200.h"
#include "em_common.cuh"

constexpr int kCountersPerRange = 16;      // [0] deferred count, [1..] one work counter per launch of a gene range
constexpr int kMaxRanges = 64;
constexpr int kCounterInts = kCountersPerRange * kMaxRanges;

struct EmBucket {
    int G;            // guide count (fast buckets); 0 for the generic bucket
    int W;            // warps per gene of the fast kernel
    int n;            // genes in the bucket
    int ids_off;      // offset into JacksEmPlan::d_ids / d_slot_off
    long long rows_off;   // offset into JacksEmPlan::d_slot_rows (fast buckets)
};

struct JacksEmPlan {
    JacksCtx* ctx;
    int64_t n_genes, n_rows, total_guides;
    int L;
    PlanBufs* b;               // device buffers (from the context's pool)
    std::vector<EmBucket> fast;
    std::vector<int> h_ids;    // host copy of d_ids (ascending gene ids inside each bucket)
    EmBucket generic;          // n == 0 if empty
    int generic_maxG;
    EmBucket big;              // genes of the generic kind large enough for a CTA each (em_big_kernel)
    int big_maxG;
    int big_grid;
    long long big_slab;        // doubles of global scratch per CTA
    int fast_maxG;
    int n_fast_genes;
    jacks::EmGenericConfig cfg_generic, cfg_deferred;
    // per-slot gene offsets and guide rows of the fast buckets: d_slot_off / d_slot_rows, or -- when ONE fast bucket holds
    // every gene in caller order (a library with a fixed guide count; every pseudo-gene index) -- d_offsets / d_rows themselves
    const long long* slot_off = nullptr;
    const int* slot_rows = nullptr;
    jacks::EmGenericConfig cfg_deferred_small;   // the deferred queue's kernel with CTAs small enough to run BESIDE a resident fast kernel
};

// Enqueue the EM of the genes with id in [g0, g1) (caller order) on `st`.  `range` selects the counter block
// (0 .. kMaxRanges-1); the caller zeroes d_counters before the first range of a run.
// grid_limit: 0 none; > 0 at most that many CTAs; < 0 leave |grid_limit| thousandths of the resident CTA slots free.
// beside: another (persistent) EM launch is resident while this one runs: the deferred queue's kernel takes its small
// configuration, or its CTAs would wait for a whole SM and stall the stream.
int em_run_range(JacksCtx* ctx, const JacksEmPlan* plan, const JacksEmParams* params, const JacksEmDeviceIO* io,
                 cudaStream_t st, int64_t g0, int64_t g1, int range, int grid_limit = 0, bool beside = false);
int em_check_io(const JacksEmPlan* plan, const JacksEmParams* params, const JacksEmDeviceIO* io);
bool host_pointer_is_pinned(const void* p);

// ------------------------------------------------------------------------------------------------
// KDE p-values (kde_pvalues.cu): the null density of every line, prepared once, evaluated for any block of genes
// ------------------------------------------------------------------------------------------------
namespace jacks {
struct KdeNull {
    double* sT; double* srt; double* stats; int* cell_first; double* loc;     // views into JacksCtx::tmp
    double* win; double* outv;    // per line: window of the cell table + the null values outside it (kde_pvalues.cu)
    long long ld; int n_lines; int positive; int brute;
};
int kde_null_prepare(JacksCtx* ctx, const double* w1_pseudo, int64_t n_pseudo, int32_t n_lines, int32_t positive,
                     cudaStream_t st, KdeNull* out);
int kde_eval(JacksCtx* ctx, const KdeNull& nl, const double* w1_genes, int64_t n_genes, double* pvals, cudaStream_t st);
}  // namespace jacks
