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
    jacks::DevBuf out[8];                   // device-side outputs of a host call that used this plan
    long long sig_genes = -1, sig_guides = -1;   // shape of the last plan served (pool lookups prefer a match)
    void release() {
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
    // data cubes kept resident between host calls (jacks_em_batched_host keep_resident)
    jacks::DevBuf res_test, res_ctrl, res_fx;
    int64_t res_N = 0;
    int res_L = 0, res_ctrl_lines = 0;
    bool res_valid = false;
    jacks::DevBuf out[8];       // device-side outputs of the host entry points
    jacks::DevBuf tmp;          // scratch of the p-value / preprocessing entry points
    // compact cubes of JACKS_UPLOAD_USED_ROWS calls: pinned staging, device copy, "upload enqueued" event
    void* cmp_stage = nullptr; size_t cmp_stage_cap = 0;
    jacks::DevBuf cmp_dev, cmp_rows;
    cudaEvent_t cmp_ready = nullptr;
    HostStager stager;          // pinned ring of the synchronous host entry points (pageable caller buffers)
    std::vector<PlanBufs*> plan_pool;          // recycled plan buffers
    std::vector<JacksEmPlan*> pending_plans;   // plans of asynchronous host calls, destroyed at the next synchronise
    std::vector<cudaEvent_t> pending_events;
};
