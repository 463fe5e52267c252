// C-ABI of jacks_b200 (include/jacks_b200.h): context, batched-EM plan, host/device entry points.
//
// Host-side work here mirrors the reference's gene loop (jacks/jacks/infer.py:18-30): the gene ->
// guide-row index arrives in CSR form, is bucketed by guide count (each bucket runs a kernel
// compiled for that G), and results are written in CSR order.  No arithmetic of the EM happens on
// the host; without a CUDA device every compute call fails with JACKS_E_CUDA.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/jacks_b200.h"
#include "capi_internal.h"
#include "em_common.cuh"

namespace jacks {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail(e == cudaErrorMemoryAllocation ? JACKS_E_NOMEM : JACKS_E_CUDA, "%s: %s (%s)", what,
                cudaGetErrorString(e), cudaGetErrorName(e));
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return JACKS_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMalloc"); }
    cap = want;
    return JACKS_OK;
}
void DevBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

}  // namespace jacks

using namespace jacks;

// JACKS_TRACE: host-side stamps relative to the start of the running screen job
static std::chrono::steady_clock::time_point g_trace_t0;
static bool g_trace_on = false;
static void trace_stamp(const char* what) {
    if (g_trace_on) fprintf(stderr, "[jacks trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_trace_t0).count());
}


namespace {

// Warps per gene of the fast kernel for (G, L); 0 = use the generic kernel.
int choose_team(int G, int L) {
    if (G < 1 || G > kMaxFastG) return 0;
    const char* env = getenv("JACKS_EM_W");          // tuning override
    int W = 0;
    if (env) W = atoi(env);
    if (!(W == 1 || W == 2 || W == 4)) {
        // One warp per gene is the most efficient split (no barriers, one reduction per pass) as long as a few genes fit
        // an SM -- in shared memory or, with tau in tensor memory, four per CTA (em_fast_single_warp_genes).  Long line
        // axes that leave room for one or two genes get 4 or 2 warps each instead (measured at G = 4: 1,100 lines
        // W = 1 / 2 / 4 -> 3.8 / 4.1 / 4.1 ms per 16,000 genes; 1,800 lines 23.4 / 14.7 / 9.0 ms).
        const int g1 = em_fast_single_warp_genes(G, L);
        W = g1 >= 3 ? 1 : (g1 == 2 ? 2 : 4);
        while (W > 1 && em_fast_smem_bytes(G, W, L) > kMaxDynSmem) W >>= 1;       // (tile per warp: cannot happen before W = 1 fails)
    }
    if (!em_fast_available(G, W) || em_fast_smem_bytes(G, W, L) > kMaxDynSmem) return 0;
    return W;
}

}  // namespace

// ---- HostStager -----------------------------------------------------------------------------------------------------

static void parallel_memcpy(void* dst, const void* src, size_t bytes, int threads) {
    if (threads <= 1 || bytes < (4u << 20)) { memcpy(dst, src, bytes); return; }
    const size_t per = ((bytes / threads) + 4095) & ~(size_t)4095;
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) {
        const size_t lo = per * t;
        if (lo >= bytes) break;
        const size_t n = std::min(per, bytes - lo);
        pool.emplace_back([=]() { memcpy((char*)dst + lo, (const char*)src + lo, n); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (auto& th : pool) th.join();
}

bool host_pointer_is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

int HostStager::init() {
    if (buf[0]) return JACKS_OK;
    cudaError_t e;
    for (int i = 0; i < 2; ++i) {
        if ((e = cudaHostAlloc(&buf[i], kChunk, cudaHostAllocDefault)) != cudaSuccess) { release(); return jacks::cuda_fail(e, "cudaHostAlloc(staging)"); }
        if ((e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)) != cudaSuccess) { release(); return jacks::cuda_fail(e, "cudaEventCreate"); }
    }
    const unsigned hw = std::thread::hardware_concurrency();
    threads = (int)std::max(1u, std::min(8u, hw ? hw / 2 : 1u));
    if (const char* env = getenv("JACKS_COPY_THREADS")) threads = std::max(1, atoi(env));
    return JACKS_OK;
}

void HostStager::release() {
    for (int i = 0; i < 2; ++i) {
        if (buf[i]) cudaFreeHost(buf[i]);
        if (ev[i]) cudaEventDestroy(ev[i]);
        buf[i] = nullptr; ev[i] = nullptr;
    }
}

int HostStager::h2d(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    cudaError_t e;
    if (bytes == 0) return JACKS_OK;
    if (bytes < (8u << 20) || host_pointer_is_pinned(src)) {
        if ((e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) return jacks::cuda_fail(e, "H2D");
        return JACKS_OK;
    }
    int rc = init();
    if (rc != JACKS_OK) return rc;
    int k = 0;
    for (size_t off = 0; off < bytes; off += kChunk, k ^= 1) {
        const size_t n = std::min(kChunk, bytes - off);
        if (off >= 2 * kChunk && (e = cudaEventSynchronize(ev[k])) != cudaSuccess) return jacks::cuda_fail(e, "staging event");   // buffer free again?
        parallel_memcpy(buf[k], (const char*)src + off, n, threads);
        if ((e = cudaMemcpyAsync((char*)dst + off, buf[k], n, cudaMemcpyHostToDevice, st)) != cudaSuccess) return jacks::cuda_fail(e, "H2D (staged)");
        cudaEventRecord(ev[k], st);
    }
    // the pinned buffers must not be refilled by a later call before these uploads have read them
    for (int i = 0; i < 2; ++i)
        if (bytes > (size_t)i * kChunk && (e = cudaEventSynchronize(ev[i])) != cudaSuccess) return jacks::cuda_fail(e, "staging event");
    return JACKS_OK;
}

int HostStager::d2h(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    cudaError_t e;
    if (bytes == 0) return JACKS_OK;
    if (bytes < (8u << 20) || host_pointer_is_pinned(dst)) {
        // small or pinned: plain asynchronous copy -- the caller synchronises `st` before it reads dst
        if ((e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return jacks::cuda_fail(e, "D2H");
        return JACKS_OK;
    }
    int rc = init();
    if (rc != JACKS_OK) return rc;
    const size_t nchunk = (bytes + kChunk - 1) / kChunk;
    for (size_t i = 0; i <= nchunk; ++i) {
        if (i < nchunk) {
            const size_t off = i * kChunk, n = std::min(kChunk, bytes - off);
            if ((e = cudaMemcpyAsync(buf[i & 1], (const char*)src + off, n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return jacks::cuda_fail(e, "D2H (staged)");
            cudaEventRecord(ev[i & 1], st);
        }
        if (i > 0) {          // drain the previous chunk while the one just issued is in flight
            const size_t j = i - 1, off = j * kChunk, n = std::min(kChunk, bytes - off);
            if ((e = cudaEventSynchronize(ev[j & 1])) != cudaSuccess) return jacks::cuda_fail(e, "staging event");
            parallel_memcpy((char*)dst + off, buf[j & 1], n, threads);
        }
    }
    return JACKS_OK;
}

extern "C" {

const char* jacks_last_error(void) { return g_last_error.c_str(); }
int jacks_abi_version(void) { return JACKS_B200_ABI_VERSION; }

int jacks_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int jacks_ctx_create(int device, JacksCtx** out) {
    if (!out) return fail(JACKS_E_INVALID, "jacks_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(JACKS_E_CUDA, "jacks_b200 needs a CUDA device and has no CPU fallback: %s",
                    e != cudaSuccess ? cudaGetErrorString(e) : "no device visible");
    }
    if (device < 0 || device >= n) return fail(JACKS_E_INVALID, "device %d out of range (0..%d)", device, n - 1);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    if (prop.major < 10)
        return fail(JACKS_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    JacksCtx* c = new JacksCtx();
    c->device = device;
    c->n_sm = prop.multiProcessorCount;
    c->launches = 0;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
    if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
    if ((e = cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking)) != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
    if ((e = cudaStreamCreateWithFlags(&c->plan_stream, cudaStreamNonBlocking)) != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
    if ((e = cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking)) != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
    if ((e = cudaStreamCreateWithFlags(&c->cmp_stream, cudaStreamNonBlocking)) != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
    if ((e = cudaStreamCreateWithFlags(&c->out2_stream, cudaStreamNonBlocking)) != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
    if ((e = cudaEventCreateWithFlags(&c->side_fork, cudaEventDisableTiming)) != cudaSuccess) { delete c; return cuda_fail(e, "cudaEventCreate"); }
    if ((e = cudaEventCreateWithFlags(&c->side_join, cudaEventDisableTiming)) != cudaSuccess) { delete c; return cuda_fail(e, "cudaEventCreate"); }
    *out = c;
    return JACKS_OK;
}

void jacks_ctx_destroy(JacksCtx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (JacksEmPlan* pl : c->pending_plans) jacks_em_plan_destroy(pl);
    for (cudaEvent_t ev : c->pending_events) cudaEventDestroy(ev);
    for (PlanBufs* pb : c->plan_pool) { pb->release(); delete pb; }
    c->res_test.release(); c->res_ctrl.release(); c->res_fx.release();
    for (auto& b : c->out) b.release();
    c->tmp.release();
    c->kde_tail.release();
    c->pre.release();
    c->stager.release();
    if (c->cmp_stage) cudaFreeHost(c->cmp_stage);
    if (c->cmp_ready) cudaEventDestroy(c->cmp_ready);
    if (c->cmp_done) cudaEventDestroy(c->cmp_done);
    c->cmp_dev.release(); c->cmp_rows.release();
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->out_stream) cudaStreamDestroy(c->out_stream);
    if (c->plan_stream) cudaStreamDestroy(c->plan_stream);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    if (c->cmp_stream) cudaStreamDestroy(c->cmp_stream);
    if (c->out2_stream) cudaStreamDestroy(c->out2_stream);
    if (c->side_fork) cudaEventDestroy(c->side_fork);
    if (c->side_join) cudaEventDestroy(c->side_join);
    if (c->res_busy) cudaEventDestroy(c->res_busy);
    delete c;
}

int64_t jacks_ctx_launch_count(const JacksCtx* c) { return c ? c->launches : 0; }

int jacks_ctx_synchronize(JacksCtx* c) {
    if (!c) return fail(JACKS_E_INVALID, "jacks_ctx_synchronize: NULL ctx");
    cudaSetDevice(c->device);
    const cudaError_t e0 = cudaStreamSynchronize(c->side_stream), e00a = cudaStreamSynchronize(c->cmp_stream);
    const cudaError_t e00b = cudaStreamSynchronize(c->plan_stream), e00c = cudaStreamSynchronize(c->out2_stream);
    const cudaError_t e00 = e00a != cudaSuccess ? e00a : (e00b != cudaSuccess ? e00b : e00c);
    const cudaError_t e1 = cudaStreamSynchronize(c->copy_stream), e2a = cudaStreamSynchronize(c->stream),
                      e3 = cudaStreamSynchronize(c->out_stream);
    const cudaError_t e2 = e2a != cudaSuccess ? e2a : (e0 != cudaSuccess ? e0 : e00);
    for (JacksEmPlan* pl : c->pending_plans) jacks_em_plan_destroy(pl);
    c->pending_plans.clear();
    for (cudaEvent_t ev : c->pending_events) cudaEventDestroy(ev);
    c->pending_events.clear();
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return cuda_fail(e2 != cudaSuccess ? e2 : (e1 != cudaSuccess ? e1 : e3), "EM pipeline");
    return JACKS_OK;
}

void jacks_em_default_params(JacksEmParams* p) {
    if (!p) return;
    p->n_iter = 50; p->apply_w_hp = 0; p->tol = 0.1; p->mu0_x = 1.0; p->var0_x = 1.0;
    p->mu0_w = 0.0; p->var0_w = 1e4; p->tau_prior_strength = 0.5;     // infer.py:18, :55
}

int jacks_host_alloc(void** out, uint64_t bytes) {
    if (!out) return fail(JACKS_E_INVALID, "jacks_host_alloc: out is NULL");
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { *out = nullptr; return cuda_fail(e, "cudaHostAlloc"); }
    return JACKS_OK;
}
void jacks_host_free(void* p) { if (p) cudaFreeHost(p); }

int jacks_host_register(void* p, uint64_t bytes) {
    if (!p || bytes == 0) return fail(JACKS_E_INVALID, "jacks_host_register: empty range");
    const cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable);
    return e == cudaSuccess ? JACKS_OK : cuda_fail(e, "cudaHostRegister");
}

int jacks_host_unregister(void* p) {
    if (!p) return JACKS_OK;
    const cudaError_t e = cudaHostUnregister(p);
    return e == cudaSuccess ? JACKS_OK : cuda_fail(e, "cudaHostUnregister");
}

int jacks_copy_to_host_async(void* dst_host, const void* src_device, uint64_t bytes, void* stream) {
    if (bytes == 0) return JACKS_OK;
    if (!dst_host || !src_device) return fail(JACKS_E_INVALID, "jacks_copy_to_host_async: NULL pointer");
    const cudaError_t e = cudaMemcpyAsync(dst_host, src_device, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    return e == cudaSuccess ? JACKS_OK : cuda_fail(e, "cudaMemcpyAsync (device -> host)");
}

}  // extern "C"

// dst[i] = src[i] for n 16-byte words; src may be pinned HOST memory (zero-copy reads over PCIe)
__global__ void __launch_bounds__(256) pull_words_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

extern "C" {

int jacks_em_plan_create(JacksCtx* ctx, const int64_t* gene_offsets, const int32_t* guide_rows, int64_t n_genes,
                         int64_t n_rows, int32_t n_lines, JacksEmPlan** out) {
    if (!ctx || !out) return fail(JACKS_E_INVALID, "jacks_em_plan_create: NULL ctx/out");
    *out = nullptr;
    if (n_genes < 0 || n_rows < 0 || n_lines < 1) return fail(JACKS_E_INVALID, "bad sizes: n_genes=%lld n_rows=%lld n_lines=%d", (long long)n_genes, (long long)n_rows, n_lines);
    if (n_genes > 0 && (!gene_offsets || !guide_rows)) return fail(JACKS_E_INVALID, "NULL gene index");
    if (n_genes >= (1LL << 31)) return fail(JACKS_E_INVALID, "too many genes");
    const int64_t total = n_genes ? gene_offsets[n_genes] : 0;
    if (n_genes && gene_offsets[0] != 0) return fail(JACKS_E_INVALID, "gene_offsets[0] must be 0");
    // pass 1 over the genes: guide counts -> bucket sizes (a gene's kind depends only on its guide count)
    int teamW[kMaxFastG + 1] = {0};
    for (int G = 1; G <= kMaxFastG; ++G) teamW[G] = choose_team(G, n_lines);
    long long cnt[kMaxFastG + 1] = {0};
    long long n_gen = 0, n_big = 0;
    int gen_maxG = 0, big_maxG = 0;
    for (int64_t g = 0; g < n_genes; ++g) {
        const int64_t G = gene_offsets[g + 1] - gene_offsets[g];
        if (G < 1) return fail(JACKS_E_INVALID, "gene %lld has %lld guides; every gene needs at least one", (long long)g, (long long)G);
        if (G > (1 << 20)) return fail(JACKS_E_INVALID, "gene %lld has too many guides", (long long)g);
        if (G <= kMaxFastG && teamW[G] > 0) cnt[G]++;
        else if ((long long)G * n_lines >= kBigGeneElements) { n_big++; big_maxG = std::max(big_maxG, (int)G); }
        else { n_gen++; gen_maxG = std::max(gen_maxG, (int)G); }
    }
    {
        unsigned bad = 0;                               // branch-free sweep (vectorises); the offender is looked up only on failure
        const unsigned lim = (unsigned)std::min<int64_t>(n_rows, 0x7fffffff);
        for (int64_t i = 0; i < total; ++i) bad |= ((unsigned)guide_rows[i] >= lim) ? 1u : 0u;
        if (bad)
            for (int64_t i = 0; i < total; ++i)
                if (guide_rows[i] < 0 || guide_rows[i] >= n_rows)
                    return fail(JACKS_E_INVALID, "guide row %d at position %lld outside [0, %lld)", guide_rows[i], (long long)i, (long long)n_rows);
    }
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");

    JacksEmPlan* p = new JacksEmPlan();
    if (!ctx->plan_pool.empty()) {
        size_t pick = ctx->plan_pool.size() - 1;          // prefer buffers that served the same shape: no regrowth
        for (size_t i = 0; i < ctx->plan_pool.size(); ++i)
            if (ctx->plan_pool[i]->sig_genes == n_genes && ctx->plan_pool[i]->sig_guides == total) { pick = i; break; }
        p->b = ctx->plan_pool[pick];
        ctx->plan_pool.erase(ctx->plan_pool.begin() + pick);
    } else p->b = new PlanBufs();
    p->b->sig_genes = n_genes; p->b->sig_guides = total;
    // recycled buffers: the copy kernel of the plan they served last must have read the pinned staging (it has, unless
    // that plan was created and destroyed without anything waiting for it)
    if (p->b->ready && (e = cudaEventSynchronize(p->b->ready)) != cudaSuccess) { cudaGetLastError(); }
    p->ctx = ctx; p->n_genes = n_genes; p->n_rows = n_rows; p->total_guides = total; p->L = n_lines;
    p->generic = EmBucket{0, 0, 0, 0, 0};
    p->big = EmBucket{0, 0, 0, 0, 0};
    p->generic_maxG = gen_maxG; p->fast_maxG = 0; p->n_fast_genes = 0; p->big_maxG = big_maxG; p->big_grid = 0; p->big_slab = 0;

    // bucket layout: fast buckets by descending guide count (the heavier buckets start while the GPU is empty), then the
    // generic bucket, then the CTA-per-gene bucket; inside a bucket the caller's order
    long long ids_cur[kMaxFastG + 3] = {0}, rows_cur[kMaxFastG + 1] = {0};       // write cursors; [kMaxFastG+1]: generic, [+2]: big
    long long n_ids = 0, n_slot_rows = 0;
    for (int G = kMaxFastG; G >= 1; --G) {
        if (!cnt[G]) continue;
        p->fast.push_back(EmBucket{G, teamW[G], (int)cnt[G], (int)n_ids, n_slot_rows});
        ids_cur[G] = n_ids; rows_cur[G] = n_slot_rows;
        n_ids += cnt[G]; n_slot_rows += cnt[G] * G;
        p->fast_maxG = std::max(p->fast_maxG, G);
        p->n_fast_genes += (int)cnt[G];
    }
    const long long n_slots = n_ids;                       // fast genes: slot_off entries
    if (n_gen) { p->generic = EmBucket{0, 0, (int)n_gen, (int)n_ids, 0}; ids_cur[kMaxFastG + 1] = n_ids; n_ids += n_gen; }
    if (n_big) {
        p->big = EmBucket{0, 0, (int)n_big, (int)n_ids, 0}; ids_cur[kMaxFastG + 2] = n_ids; n_ids += n_big;
        p->big_grid = (int)std::min<long long>(n_big, 2LL * ctx->n_sm);
        p->big_slab = em_generic_work_doubles(p->big_maxG, n_lines);
    }

    int rc = JACKS_OK;
    // The index arrays are built IN the plan's pinned staging and pulled by a copy kernel on the plan stream: a
    // cudaMemcpyAsync would queue in the copy engine behind whatever upload is in flight (the screen of a screen job).
    // one bucket holds every gene: slot order is gene order, and the per-slot arrays ARE the index arrays (no second copy)
    const bool one_bucket = p->fast.size() == 1 && !n_gen && !n_big;
    const size_t lens[5] = {n_genes ? (size_t)(n_genes + 1) * sizeof(int64_t) : 0, (size_t)total * sizeof(int32_t), (size_t)n_ids * sizeof(int),
                            one_bucket ? 0 : (size_t)n_slots * sizeof(long long), one_bucket ? 0 : (size_t)n_slot_rows * sizeof(int)};
    DevBuf* dsts[5] = {&p->b->d_offsets, &p->b->d_rows, &p->b->d_ids, &p->b->d_slot_off, &p->b->d_slot_rows};
    size_t offs5[5], stage_bytes = 0;
    for (int i = 0; i < 5; ++i) { offs5[i] = stage_bytes; stage_bytes += (lens[i] + 255) & ~(size_t)255; }
    if (stage_bytes > p->b->h_cap) {
        if (p->b->h_stage) cudaFreeHost(p->b->h_stage);
        p->b->h_stage = nullptr; p->b->h_cap = 0;
        const size_t want = stage_bytes + stage_bytes / 8 + 4096;
        if ((e = cudaHostAlloc(&p->b->h_stage, want, cudaHostAllocDefault)) != cudaSuccess) rc = cuda_fail(e, "cudaHostAlloc(plan staging)");
        else p->b->h_cap = want;
    }
    if (rc == JACKS_OK && n_genes > 0) {
        char* hs = (char*)p->b->h_stage;
        memcpy(hs + offs5[0], gene_offsets, lens[0]);
        memcpy(hs + offs5[1], guide_rows, lens[1]);
        int* s_ids = (int*)(hs + offs5[2]);
        long long* s_off = (long long*)(hs + offs5[3]);
        int* s_rows = (int*)(hs + offs5[4]);
        if (one_bucket) {
            for (int64_t g = 0; g < n_genes; ++g) s_ids[g] = (int)g;
        } else {
            for (int64_t g = 0; g < n_genes; ++g) {
                const long long o = gene_offsets[g];
                const int G = (int)(gene_offsets[g + 1] - o);
                if (G <= kMaxFastG && teamW[G] > 0) {
                    const long long slot = ids_cur[G]++;
                    s_ids[slot] = (int)g;
                    s_off[slot] = o;
                    int* dst = s_rows + rows_cur[G];
                    for (int k = 0; k < G; ++k) dst[k] = guide_rows[o + k];
                    rows_cur[G] += G;
                } else if ((long long)G * n_lines >= kBigGeneElements) s_ids[ids_cur[kMaxFastG + 2]++] = (int)g;
                else s_ids[ids_cur[kMaxFastG + 1]++] = (int)g;
            }
        }
        p->h_ids.assign(s_ids, s_ids + n_ids);
    }
    trace_stamp("    plan: buckets built");
    void* stage_dev = nullptr;
    if (rc == JACKS_OK && stage_bytes && (e = cudaHostGetDevicePointer(&stage_dev, p->b->h_stage, 0)) != cudaSuccess) rc = cuda_fail(e, "cudaHostGetDevicePointer");
    for (int i = 0; i < 5 && rc == JACKS_OK; ++i) {
        rc = dsts[i]->reserve(lens[i] ? ((lens[i] + 15) & ~(size_t)15) : 16);
        if (rc != JACKS_OK || !lens[i]) continue;
        const size_t n16 = (lens[i] + 15) / 16;               // staging slots and device buffers are padded to 16 bytes
        pull_words_kernel<<<(unsigned)std::min<size_t>((n16 + 255) / 256, 592), 256, 0, ctx->plan_stream>>>(
            (const uint4*)((const char*)stage_dev + offs5[i]), (uint4*)dsts[i]->p, n16);
        if ((e = cudaGetLastError()) != cudaSuccess) rc = cuda_fail(e, "pull_words_kernel launch");
        ctx->launches++;
    }
    p->slot_off = (const long long*)(one_bucket ? p->b->d_offsets.p : p->b->d_slot_off.p);
    p->slot_rows = (const int*)(one_bucket ? p->b->d_rows.p : p->b->d_slot_rows.p);
    if (rc == JACKS_OK) rc = p->b->d_deferred.reserve((size_t)std::max<int64_t>(n_genes, 1) * sizeof(int));
    if (rc == JACKS_OK) rc = p->b->d_counters.reserve(kCounterInts * sizeof(int));
    if (rc == JACKS_OK && p->fast_maxG > 0)
        rc = p->b->d_tpd.reserve((size_t)ctx->n_sm * kMaxFastCtasPerSm * p->fast_maxG * em_fast_tpd_stride(n_lines) * sizeof(double));
    // generic kernel configurations (and global scratch if a gene's working set exceeds shared memory)
    memset(&p->cfg_generic, 0, sizeof p->cfg_generic);
    memset(&p->cfg_deferred, 0, sizeof p->cfg_deferred);
    memset(&p->cfg_deferred_small, 0, sizeof p->cfg_deferred_small);
    long long scratch_doubles = (long long)p->big_grid * p->big_slab;
    if (rc == JACKS_OK && p->generic.n > 0) {
        e = em_generic_configure(p->generic_maxG, n_lines, p->generic.n, ctx->n_sm, &p->cfg_generic);
        if (e != cudaSuccess) rc = cuda_fail(e, "em_generic_configure");
        else scratch_doubles = std::max(scratch_doubles, p->cfg_generic.scratch_doubles_per_warp * p->cfg_generic.grid * p->cfg_generic.warps_per_cta);
    }
    if (rc == JACKS_OK && p->n_fast_genes > 0) {
        e = em_generic_configure(p->fast_maxG, n_lines, p->n_fast_genes, ctx->n_sm, &p->cfg_deferred);
        if (e != cudaSuccess) rc = cuda_fail(e, "em_generic_configure");
        else scratch_doubles = std::max(scratch_doubles, p->cfg_deferred.scratch_doubles_per_warp * p->cfg_deferred.grid * p->cfg_deferred.warps_per_cta);
        // beside a resident fast kernel an SM has ~100 KB of shared memory to spare (two TMEM-variant CTAs leave 54 KB, one 140 KB)
        if (rc == JACKS_OK && (e = em_generic_configure(p->fast_maxG, n_lines, p->n_fast_genes, ctx->n_sm, &p->cfg_deferred_small, 100 * 1024)) != cudaSuccess) rc = cuda_fail(e, "em_generic_configure");
        else scratch_doubles = std::max(scratch_doubles, p->cfg_deferred_small.scratch_doubles_per_warp * p->cfg_deferred_small.grid * p->cfg_deferred_small.warps_per_cta);
    }
    if (rc == JACKS_OK && scratch_doubles > 0) rc = p->b->d_scratch.reserve((size_t)scratch_doubles * sizeof(double));
    trace_stamp("    plan: pulls enqueued");
    if (rc == JACKS_OK) {
        // The index arrives on a stream of its own and the plan's consumers wait for it ON THE DEVICE (em_plan_wait): the
        // host does not (the staging belongs to the plan's buffers and lives until they are recycled), and neither do
        // kernels of an earlier asynchronous call on the compute stream.
        if (!p->b->ready && (e = cudaEventCreateWithFlags(&p->b->ready, cudaEventDisableTiming)) != cudaSuccess) rc = cuda_fail(e, "cudaEventCreate");
        else if ((e = cudaEventRecord(p->b->ready, ctx->plan_stream)) != cudaSuccess) rc = cuda_fail(e, "cudaEventRecord");
    }
    if (rc != JACKS_OK) { jacks_em_plan_destroy(p); return rc; }
    *out = p;
    return JACKS_OK;
}

void jacks_em_plan_destroy(JacksEmPlan* p) {
    if (!p) return;
    if (p->b) p->ctx->plan_pool.push_back(p->b);      // buffers are recycled, freed with the context
    delete p;
}

}  // extern "C"

// Enqueue the EM of the genes with id in [g0, g1) (caller order) on `st`.  `range` selects the counter block;
// the caller zeroes d_counters before the first range of a run.
int em_run_range(JacksCtx* ctx, const JacksEmPlan* plan, const JacksEmParams* params, const JacksEmDeviceIO* io,
                        cudaStream_t st, int64_t g0, int64_t g1, int range, int grid_limit, bool beside) {
    cudaError_t e;
    if (plan->b->ready && (e = cudaStreamWaitEvent(st, plan->b->ready, 0)) != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent(plan)");
    int* counters = (int*)plan->b->d_counters.p + range * kCountersPerRange;
    EmLaunch a;
    memset(&a, 0, sizeof a);
    a.test = (const double2*)io->test; a.ctrl = (const double2*)io->ctrl;
    a.L = plan->L; a.ctrl_lines = io->ctrl_lines; a.ctrl_rowmean_fill = io->ctrl_rowmean_fill;
    a.gene_offsets = (const long long*)plan->b->d_offsets.p; a.guide_rows = (const int*)plan->b->d_rows.p;
    a.fixed_x1 = io->fixed_x1; a.fixed_x2 = io->fixed_x2;
    a.x1 = io->x1; a.x2 = io->x2; a.w1 = io->w1; a.w2 = io->w2; a.y = io->y; a.tau = io->tau;
    a.n_iters = io->n_iters; a.bound = io->bound;
    a.deferred_ids = (int*)plan->b->d_deferred.p + g0; a.deferred_count = counters;     // a range defers at most its own genes
    a.n_iter = params->n_iter; a.apply_w_hp = params->apply_w_hp; a.tol = params->tol;
    a.mu0_x = params->mu0_x; a.var0_x = params->var0_x; a.mu0_w = params->mu0_w; a.var0_w = params->var0_w;
    a.tps = params->tau_prior_strength;
    a.scratch = (double*)plan->b->d_scratch.p;
    a.tpd_scratch = (double*)plan->b->d_tpd.p;
    a.tpd_stride = em_fast_tpd_stride(plan->L);

    auto slots_of = [&](const EmBucket& b, int* s0, int* s1) {     // ids ascend inside a bucket
        const int* first = plan->h_ids.data() + b.ids_off;
        *s0 = (int)(std::lower_bound(first, first + b.n, (int)g0) - first);
        *s1 = (int)(std::lower_bound(first, first + b.n, (int)g1) - first);
    };
    int slot = 1, n_fast = 0;
    for (const EmBucket& b : plan->fast) {
        int s0, s1;
        slots_of(b, &s0, &s1);
        if (s1 <= s0) continue;
        a.gene_ids = (const int*)plan->b->d_ids.p + b.ids_off + s0;
        a.slot_off = plan->slot_off + b.ids_off + s0;
        a.slot_rows = plan->slot_rows + b.rows_off + (long long)s0 * b.G;
        a.n_work = s1 - s0; a.n_work_dev = nullptr;
        a.work_counter = counters + slot++;
        int grid = 0;
        e = launch_em_fast(b.G, b.W, a, ctx->n_sm, grid_limit, st, &grid);
        if (e != cudaSuccess) return cuda_fail(e, "em_fast_kernel launch");
        ctx->launches++;
        n_fast += s1 - s0;
    }
    if (plan->generic.n > 0) {
        int s0, s1;
        slots_of(plan->generic, &s0, &s1);
        if (s1 > s0) {
            a.gene_ids = (const int*)plan->b->d_ids.p + plan->generic.ids_off + s0;
            a.n_work = s1 - s0; a.n_work_dev = nullptr;
            a.work_counter = counters + slot++;
            a.scratch_stride = plan->cfg_generic.scratch_doubles_per_warp;
            e = launch_em_generic(plan->cfg_generic, a, st);
            if (e != cudaSuccess) return cuda_fail(e, "em_generic_kernel launch");
            ctx->launches++;
        }
    }
    if (plan->big.n > 0) {
        int s0, s1;
        slots_of(plan->big, &s0, &s1);
        if (s1 > s0) {
            a.gene_ids = (const int*)plan->b->d_ids.p + plan->big.ids_off + s0;
            a.n_work = s1 - s0; a.n_work_dev = nullptr;
            a.work_counter = counters + slot++;
            e = launch_em_big(a, std::min(plan->big_grid, s1 - s0), plan->big_slab, st);
            if (e != cudaSuccess) return cuda_fail(e, "em_big_kernel launch");
            ctx->launches++;
        }
    }
    if (n_fast > 0) {
        // genes the fast kernels handed back because their data holds NaN/inf
        a.gene_ids = (const int*)plan->b->d_deferred.p + g0;
        a.n_work = n_fast; a.n_work_dev = counters;
        a.work_counter = counters + slot++;
        const EmGenericConfig& dcfg = beside ? plan->cfg_deferred_small : plan->cfg_deferred;
        a.scratch_stride = dcfg.scratch_doubles_per_warp;
        e = launch_em_generic(dcfg, a, st);
        if (e != cudaSuccess) return cuda_fail(e, "em_generic_kernel (deferred) launch");
        ctx->launches++;
    }
    return JACKS_OK;
}

int em_check_io(const JacksEmPlan* plan, const JacksEmParams* params, const JacksEmDeviceIO* io) {
    if (!io->test || !io->ctrl || !io->w1) return fail(JACKS_E_INVALID, "test, ctrl and w1 are required");
    if (io->ctrl_lines != plan->L && io->ctrl_lines != 1)
        return fail(JACKS_E_SHAPE, "Expecting matched size between control and test data");   // infer.py:26
    if ((io->fixed_x1 == nullptr) != (io->fixed_x2 == nullptr)) return fail(JACKS_E_INVALID, "fixed_x1 and fixed_x2 go together");
    if (params->n_iter < 0) return fail(JACKS_E_INVALID, "n_iter < 0");
    return JACKS_OK;
}

extern "C" {

int jacks_em_run_device(JacksCtx* ctx, const JacksEmPlan* plan, const JacksEmParams* params,
                        const JacksEmDeviceIO* io, void* stream) {
    if (!ctx || !plan || !params || !io) return fail(JACKS_E_INVALID, "jacks_em_run_device: NULL argument");
    if (plan->n_genes == 0) return JACKS_OK;
    int rc = em_check_io(plan, params, io);
    if (rc != JACKS_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    if ((e = cudaMemsetAsync(plan->b->d_counters.p, 0, kCountersPerRange * sizeof(int), st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    return em_run_range(ctx, plan, params, io, st, 0, plan->n_genes, 0);
}

}  // extern "C"

// Device cubes that hold only the rows an index references (JACKS_UPLOAD_USED_ROWS), already enqueued for upload.
struct CompactCubes {
    const double* d_test; const double* d_ctrl;
    const double* d_fx1; const double* d_fx2;
    int64_t n_rows;
    cudaEvent_t ready;
};

// JACKS_TRACE: device-side timeline of a screen job (timing events recorded on the streams, printed after the synchronise)
struct TraceMarks {
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    cudaEvent_t t0 = nullptr;
    void start(cudaStream_t st) { cudaEventCreate(&t0); cudaEventRecord(t0, st); }
    void mark(const std::string& what, cudaStream_t st) {
        cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, st); marks.emplace_back(what, ev);
    }
    void print() {
        for (auto& m : marks) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, t0, m.second) == cudaSuccess) fprintf(stderr, "[jacks trace]   device %-22s %8.3f ms\n", m.first.c_str(), ms);
            cudaEventDestroy(m.second);
        }
        if (t0) cudaEventDestroy(t0);
        marks.clear(); t0 = nullptr;
    }
};

// What the pipelined host run does besides the plain inferJACKS call.
struct HostRunOpts {
    TraceMarks* trace = nullptr;
    int32_t flags = 0;                    // JACKS_KEEP_RESIDENT | JACKS_ASYNC
    const CompactCubes* cc = nullptr;     // run on compact cubes (already enqueued for upload) instead of the resident screen
    bool w1_stays = false;                // E[w] is produced on the device even when io->w1 is NULL (nothing is downloaded for it)
    const jacks::KdeNull* kde = nullptr;  // p-values of every gene range right after its EM (jacks_run_screen_host) ...
    double* pvals_host = nullptr;         // ... downloaded here, [n_genes, L]
    JacksEmPlan** plan_out = nullptr;     // the (pending) plan: its device outputs stay valid until the next synchronise
    const struct PreUpload* pre = nullptr; // the screen is already on its way up (enqueue_screen_upload): ranges + events
    cudaStream_t em_stream = nullptr;     // the EM kernels of the ranges run here instead of the compute stream (the p-values
                                          // stay on the compute stream, behind whatever produces the null there)
    int grid_limit = 0;                   // em_run_range's grid limit (< 0: leave a share of the SM slots to a parallel run)
    int ranges = 0;                       // > 0: this many gene ranges instead of host_run_ranges() (1: one launch per bucket --
                                          // a run that neither waits for an upload nor downloads anything gains nothing from ranges
                                          // and pays a kernel tail for each)
};

// Gene ranges of the pipelined host run: contiguous blocks; with p-values on, boundaries sit on multiples of 1,024 genes.
static int host_run_ranges(int64_t n_genes) {
    static const int cap = []() { const char* v = getenv("JACKS_HOST_RANGES"); return v ? atoi(v) : 16; }();      // tuning
    int R = 1;
    if (n_genes >= 4096) R = (int)std::min<int64_t>(std::max(1, cap), n_genes / 2048);
    return std::min(R, kMaxRanges);
}
static int64_t host_run_bound(int64_t n_genes, int R, int r, bool kde) {
    if (r >= R) return n_genes;
    int64_t g = n_genes * r / R;
    if (kde && R > 1) g = std::min<int64_t>(n_genes, (g + 512) / 1024 * 1024);
    return g;
}

// Upload of a screen enqueued AHEAD of everything else of a screen job (pinned host cubes only: the copies are
// asynchronous): ev_in[r] fires when the rows of gene range r are on the device.
struct PreUpload {
    TraceMarks* trace = nullptr;
    int R = 0;
    std::vector<cudaEvent_t> ev_in;
};

static int enqueue_screen_upload(JacksCtx* ctx, const JacksEmHostIO* io, const int64_t* gene_offsets, const int32_t* guide_rows,
                                 int64_t n_genes, bool kde, PreUpload* pre) {
    const int64_t N = io->n_rows;
    const int L = io->n_lines;
    const size_t row_t = (size_t)L * 2 * sizeof(double), row_c = (size_t)io->ctrl_lines * 2 * sizeof(double);
    int rc;
    cudaError_t e;
    if ((rc = ctx->res_test.reserve(N ? N * row_t : 8)) != JACKS_OK) return rc;
    if ((rc = ctx->res_ctrl.reserve(N ? N * row_c : 8)) != JACKS_OK) return rc;
    ctx->res_N = N; ctx->res_L = L; ctx->res_ctrl_lines = io->ctrl_lines; ctx->res_valid = true;
    cudaStream_t s_in = ctx->copy_stream;
    if (ctx->res_busy) cudaStreamWaitEvent(s_in, ctx->res_busy, 0);
    const int R = host_run_ranges(n_genes);
    const int64_t T = n_genes ? gene_offsets[n_genes] : 0;
    bool ascending = true;
    for (int64_t i = 1; i < T && ascending; ++i) ascending = guide_rows[i] >= guide_rows[i - 1];
    pre->R = R;
    pre->ev_in.assign(R, nullptr);
    int64_t row_done = 0;
    for (int r = 0; r < R; ++r) {
        const int64_t g0 = host_run_bound(n_genes, R, r, kde), g1 = host_run_bound(n_genes, R, r + 1, kde);
        int64_t row_hi = N;
        if (ascending && r + 1 < R && g1 > g0) row_hi = std::min<int64_t>(N, (int64_t)guide_rows[gene_offsets[g1] - 1] + 1);
        else if (!ascending && r > 0) row_hi = row_done;
        if (row_hi > row_done) {
            if ((e = cudaMemcpyAsync((char*)ctx->res_test.p + row_done * row_t, (const char*)io->test + row_done * row_t, (row_hi - row_done) * row_t, cudaMemcpyHostToDevice, s_in)) != cudaSuccess) return cuda_fail(e, "H2D screen");
            if ((e = cudaMemcpyAsync((char*)ctx->res_ctrl.p + row_done * row_c, (const char*)io->ctrl + row_done * row_c, (row_hi - row_done) * row_c, cudaMemcpyHostToDevice, s_in)) != cudaSuccess) return cuda_fail(e, "H2D screen");
            row_done = row_hi;
        }
        if ((e = cudaEventCreateWithFlags(&pre->ev_in[r], cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
        cudaEventRecord(pre->ev_in[r], s_in);
        if (pre->trace && (r == 0 || r == R - 1 || r == R / 2)) pre->trace->mark("upload range " + std::to_string(r), s_in);
        ctx->pending_events.push_back(pre->ev_in[r]);
    }
    return JACKS_OK;
}

static int em_batched_host_core(JacksCtx* ctx, const JacksEmParams* params, const int64_t* gene_offsets,
                                const int32_t* guide_rows, int64_t n_genes, const JacksEmHostIO* io, const HostRunOpts& opt);

// Row gather straight out of PINNED host memory (zero-copy reads over PCIe): one CTA per (row, cube), 16-byte loads.
// The SMs pull the rows themselves, so neither a host-side gather nor the H2D copy engine's queue (FIFO across
// streams: a small copy enqueued behind a screen-sized upload waits for all of it) stands in the way.
__global__ void __launch_bounds__(128) gather_rows_kernel(const double2* __restrict__ src_test, const double2* __restrict__ src_ctrl,
                                                          const long long* __restrict__ used, long long row_t2, long long row_c2,
                                                          double2* __restrict__ dst_test, double2* __restrict__ dst_ctrl) {
    const long long u = blockIdx.x;
    const long long r = used[u];
    const bool is_ctrl = blockIdx.y != 0;
    const long long n = is_ctrl ? row_c2 : row_t2;
    const double2* s = (is_ctrl ? src_ctrl : src_test) + r * n;
    double2* d = (is_ctrl ? dst_ctrl : dst_test) + u * n;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

// JACKS_UPLOAD_USED_ROWS: gather the rows the index references (ascending) into a compact cube -- by a kernel reading
// the caller's pinned memory directly, or on the host (all cores, pinned staging) when the memory is pageable --
// and run the EM on a re-based index.  For a pseudo-gene index over a few thousand control guides
// this moves tens of megabytes instead of the whole screen, so its kernels and downloads need not wait for it.
static int em_batched_host_compact(JacksCtx* ctx, const JacksEmParams* params, const int64_t* gene_offsets,
                                   const int32_t* guide_rows, int64_t n_genes, const JacksEmHostIO* io, HostRunOpts opt) {
    const int64_t N = io->n_rows;
    const int L = io->n_lines;
    const int64_t T = n_genes ? gene_offsets[n_genes] : 0;
    std::vector<int32_t> remap((size_t)N, -1);
    for (int64_t i = 0; i < T; ++i) {
        const int32_t r = guide_rows[i];
        if (r < 0 || r >= N) return fail(JACKS_E_INVALID, "guide row %d at position %lld outside [0, %lld)", r, (long long)i, (long long)N);
        remap[(size_t)r] = 0;
    }
    std::vector<int64_t> used;
    for (int64_t r = 0; r < N; ++r)
        if (remap[(size_t)r] == 0) { remap[(size_t)r] = (int32_t)used.size(); used.push_back(r); }
    const int64_t U = (int64_t)used.size();
    trace_stamp("  compact: rows marked");
    const size_t row_t = (size_t)L * 2 * sizeof(double), row_c = (size_t)io->ctrl_lines * 2 * sizeof(double);
    const bool fx = io->fixed_x1 && io->fixed_x2;
    const size_t bytes = (size_t)U * (row_t + row_c + (fx ? 2 * sizeof(double) : 0));
    cudaError_t e;
    int rc;
    // staging is reused from call to call: wait until the previous compact upload has read it
    if (ctx->cmp_ready && (e = cudaEventSynchronize(ctx->cmp_ready)) != cudaSuccess) return cuda_fail(e, "compact staging event");
    if (bytes > ctx->cmp_stage_cap) {
        if (ctx->cmp_stage) cudaFreeHost(ctx->cmp_stage);
        ctx->cmp_stage = nullptr; ctx->cmp_stage_cap = 0;
        if ((e = cudaHostAlloc(&ctx->cmp_stage, bytes + (bytes >> 2) + 4096, cudaHostAllocDefault)) != cudaSuccess) return cuda_fail(e, "cudaHostAlloc(compact staging)");
        ctx->cmp_stage_cap = bytes + (bytes >> 2) + 4096;
    }
    if (!ctx->cmp_ready && (e = cudaEventCreateWithFlags(&ctx->cmp_ready, cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
    if ((rc = ctx->cmp_dev.reserve(bytes ? bytes : 8)) != JACKS_OK) return rc;
    char* st_test = (char*)ctx->cmp_stage;
    char* st_ctrl = st_test + (size_t)U * row_t;
    double* st_fx = (double*)(st_ctrl + (size_t)U * row_c);
    const bool zero_copy = U > 0 && host_pointer_is_pinned(io->test) && host_pointer_is_pinned(io->ctrl);
    // pinned cubes: the GPU pulls the rows itself on a stream of its own (nothing of this touches the copy engine, whose
    // queue may hold a screen-sized upload); pageable cubes: host gather into the staging, one copy on the copy stream
    cudaStream_t s_in = zero_copy ? ctx->cmp_stream : ctx->copy_stream;
    // the device cube is reused from call to call too: the kernels of the previous compact run (compute stream) may
    // still be reading it, so the refill below is ordered behind them
    if (ctx->cmp_done) cudaStreamWaitEvent(s_in, ctx->cmp_done, 0);
    if (zero_copy) {
        // the list of rows stays in the pinned staging (the gather kernel reads it from there) and the (small) fixed-x
        // gather goes through it too
        long long* st_used = (long long*)ctx->cmp_stage;            // staging is at least U*(row_t+row_c) >= 32*U bytes
        for (int64_t u = 0; u < U; ++u) st_used[u] = used[(size_t)u];
        void* dt = nullptr; void* dc = nullptr; void* du = nullptr;
        if ((e = cudaHostGetDevicePointer(&dt, (void*)io->test, 0)) != cudaSuccess) return cuda_fail(e, "cudaHostGetDevicePointer");
        if ((e = cudaHostGetDevicePointer(&dc, (void*)io->ctrl, 0)) != cudaSuccess) return cuda_fail(e, "cudaHostGetDevicePointer");
        if ((e = cudaHostGetDevicePointer(&du, ctx->cmp_stage, 0)) != cudaSuccess) return cuda_fail(e, "cudaHostGetDevicePointer");
        char* dev = (char*)ctx->cmp_dev.p;
        gather_rows_kernel<<<dim3((unsigned)U, 2), 128, 0, s_in>>>((const double2*)dt, (const double2*)dc, (const long long*)du,
                                                                   (long long)L, (long long)io->ctrl_lines, (double2*)dev,
                                                                   (double2*)(dev + (size_t)U * row_t));
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "gather_rows_kernel launch");
        ctx->launches++;
        if (fx) {
            std::vector<double> hfx((size_t)(2 * U));
            for (int64_t u = 0; u < U; ++u) { hfx[(size_t)u] = io->fixed_x1[used[(size_t)u]]; hfx[(size_t)(U + u)] = io->fixed_x2[used[(size_t)u]]; }
            if ((e = cudaMemcpyAsync(dev + (size_t)U * (row_t + row_c), hfx.data(), hfx.size() * sizeof(double), cudaMemcpyHostToDevice, s_in)) != cudaSuccess) return cuda_fail(e, "H2D fixed_x");
            if ((e = cudaStreamSynchronize(s_in)) != cudaSuccess) return cuda_fail(e, "H2D fixed_x");      // hfx goes out of scope
        }
    } else {
        const unsigned hw = std::thread::hardware_concurrency();
        int n_thr = (int)std::max(1u, std::min(8u, hw ? hw / 2 : 1u));
        if (bytes < (4u << 20)) n_thr = 1;
        auto gather = [&](int64_t u0, int64_t u1) {
            for (int64_t u = u0; u < u1; ++u) {
                const int64_t r = used[(size_t)u];
                memcpy(st_test + (size_t)u * row_t, (const char*)io->test + (size_t)r * row_t, row_t);
                memcpy(st_ctrl + (size_t)u * row_c, (const char*)io->ctrl + (size_t)r * row_c, row_c);
                if (fx) { st_fx[u] = io->fixed_x1[r]; st_fx[U + u] = io->fixed_x2[r]; }
            }
        };
        std::vector<std::thread> pool;
        const int64_t per = (U + n_thr - 1) / n_thr;
        for (int t = 1; t < n_thr; ++t) {
            const int64_t u0 = per * t, u1 = std::min<int64_t>(U, u0 + per);
            if (u0 < u1) pool.emplace_back(gather, u0, u1);
        }
        gather(0, std::min<int64_t>(U, per));
        for (auto& th : pool) th.join();
        if (bytes && (e = cudaMemcpyAsync(ctx->cmp_dev.p, ctx->cmp_stage, bytes, cudaMemcpyHostToDevice, s_in)) != cudaSuccess) return cuda_fail(e, "H2D compact cube");
    }
    cudaEventRecord(ctx->cmp_ready, s_in);
    trace_stamp("  compact: gather enqueued");
    // the index re-based onto the compact cube -- after the gather is on its way: the pseudo-gene EM of a screen job waits
    // for those rows (25 MB pulled over a PCIe link the screen upload is using too), not for this loop
    std::vector<int32_t> rows2((size_t)T);
    for (int64_t i = 0; i < T; ++i) rows2[(size_t)i] = remap[(size_t)guide_rows[i]];
    trace_stamp("  compact: rows re-based");
    CompactCubes cc;
    cc.d_test = (const double*)ctx->cmp_dev.p;
    cc.d_ctrl = (const double*)((const char*)ctx->cmp_dev.p + (size_t)U * row_t);
    cc.d_fx1 = fx ? (const double*)((const char*)ctx->cmp_dev.p + (size_t)U * (row_t + row_c)) : nullptr;
    cc.d_fx2 = fx ? cc.d_fx1 + U : nullptr;
    cc.n_rows = U;
    cc.ready = ctx->cmp_ready;
    opt.cc = &cc;
    const int rc2 = em_batched_host_core(ctx, params, gene_offsets, rows2.data(), n_genes, io, opt);
    if (!ctx->cmp_done && (e = cudaEventCreateWithFlags(&ctx->cmp_done, cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
    cudaEventRecord(ctx->cmp_done, ctx->stream);        // the last kernel that reads the compact cube
    return rc2;
}

extern "C" {

int jacks_em_batched_host(JacksCtx* ctx, const JacksEmParams* params, const int64_t* gene_offsets,
                          const int32_t* guide_rows, int64_t n_genes, const JacksEmHostIO* io, int32_t flags) {
    if (!ctx || !params || !io) return fail(JACKS_E_INVALID, "jacks_em_batched_host: NULL argument");
    if ((flags & JACKS_UPLOAD_USED_ROWS) && io->test) {
        if (!io->ctrl) return fail(JACKS_E_INVALID, "ctrl is NULL");
        if (io->n_rows < 0 || io->n_lines < 1) return fail(JACKS_E_INVALID, "bad data shape N=%lld L=%d", (long long)io->n_rows, io->n_lines);
        if (io->ctrl_lines != io->n_lines && io->ctrl_lines != 1)
            return fail(JACKS_E_SHAPE, "Expecting matched size between control and test data");
        if ((io->fixed_x1 == nullptr) != (io->fixed_x2 == nullptr)) return fail(JACKS_E_INVALID, "fixed_x1 and fixed_x2 go together");
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
        HostRunOpts opt;
        opt.flags = flags;
        return em_batched_host_compact(ctx, params, gene_offsets, guide_rows, n_genes, io, opt);
    }
    HostRunOpts opt;
    opt.flags = flags;
    return em_batched_host_core(ctx, params, gene_offsets, guide_rows, n_genes, io, opt);
}

}  // extern "C"

static int em_batched_host_core(JacksCtx* ctx, const JacksEmParams* params, const int64_t* gene_offsets,
                                const int32_t* guide_rows, int64_t n_genes, const JacksEmHostIO* io, const HostRunOpts& opt) {
    const CompactCubes* cc = opt.cc;
    const int32_t keep_resident = opt.flags;
    const int64_t N = cc ? cc->n_rows : io->n_rows;
    const int L = io->n_lines;
    if (N < 0 || L < 1) return fail(JACKS_E_INVALID, "bad data shape N=%lld L=%d", (long long)N, L);
    if (io->ctrl_lines != L && io->ctrl_lines != 1)
        return fail(JACKS_E_SHAPE, "Expecting matched size between control and test data");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    cudaStream_t s_k = ctx->stream, s_in = ctx->copy_stream;
    cudaStream_t s_em = opt.em_stream ? opt.em_stream : s_k;      // EM kernels; s_k keeps the p-values
    const bool split = s_em != s_k;
    int rc;

    // ---- inputs: upload, or reuse the cubes a previous call left resident -------------------------
    const size_t row_t = (size_t)L * 2 * sizeof(double), row_c = (size_t)io->ctrl_lines * 2 * sizeof(double);
    const bool upload = (io->test != nullptr) && !cc && !opt.pre;
    if (cc) {
        // compact cubes: uploaded by the caller, the resident screen (if any) stays as it is
    } else if (upload) {
        if (!io->ctrl) return fail(JACKS_E_INVALID, "ctrl is NULL");
        if ((rc = ctx->res_test.reserve(N ? N * row_t : 8)) != JACKS_OK) return rc;
        if ((rc = ctx->res_ctrl.reserve(N ? N * row_c : 8)) != JACKS_OK) return rc;
        ctx->res_N = N; ctx->res_L = L; ctx->res_ctrl_lines = io->ctrl_lines; ctx->res_valid = true;
    } else {
        if (!ctx->res_valid || ctx->res_N != N || ctx->res_L != L || ctx->res_ctrl_lines != io->ctrl_lines)
            return fail(JACKS_E_INVALID, "test is NULL but no matching resident data cube (call with keep_resident first)");
    }
    const double* d_fx1 = cc ? cc->d_fx1 : nullptr; const double* d_fx2 = cc ? cc->d_fx2 : nullptr;
    if (!cc && (io->fixed_x1 || io->fixed_x2)) {
        if (!io->fixed_x1 || !io->fixed_x2) return fail(JACKS_E_INVALID, "fixed_x1 and fixed_x2 go together");
        if ((rc = ctx->res_fx.reserve((size_t)N * 2 * sizeof(double) + 8)) != JACKS_OK) return rc;
        if (N) {
            if ((e = cudaMemcpyAsync(ctx->res_fx.p, io->fixed_x1, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, s_em)) != cudaSuccess) return cuda_fail(e, "H2D fixed_x1");
            if ((e = cudaMemcpyAsync((double*)ctx->res_fx.p + N, io->fixed_x2, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, s_em)) != cudaSuccess) return cuda_fail(e, "H2D fixed_x2");
        }
        d_fx1 = (const double*)ctx->res_fx.p; d_fx2 = d_fx1 + N;
    }

    JacksEmPlan* plan = nullptr;
    if ((rc = jacks_em_plan_create(ctx, gene_offsets, guide_rows, n_genes, N, L, &plan)) != JACKS_OK) return rc;
    trace_stamp("  core: plan created");
    const int64_t T = plan->total_guides;

    // ---- outputs ------------------------------------------------------------------------------------
    // per-unit sizes: [0,1] per guide, [2,3] per gene row, [4,5] per guide row, [6,7] per gene
    constexpr int NO = 9;         // [8]: p-values of the screen job
    struct Out { char* host; size_t unit; bool per_guide; } outs[NO] = {
        {(char*)io->x1, sizeof(double), true},         {(char*)io->x2, sizeof(double), true},
        {(char*)io->w1, (size_t)L * sizeof(double), false}, {(char*)io->w2, (size_t)L * sizeof(double), false},
        {(char*)io->y, (size_t)L * sizeof(double), true},   {(char*)io->tau, (size_t)L * sizeof(double), true},
        {(char*)io->n_iters, sizeof(int32_t), false},  {(char*)io->bound, sizeof(double), false},
        {(char*)(opt.kde ? opt.pvals_host : nullptr), (size_t)L * sizeof(double), false}};
    if (!io->w1 && !opt.w1_stays && n_genes) { jacks_em_plan_destroy(plan); return fail(JACKS_E_INVALID, "w1 is required"); }
    char* dptr[NO] = {nullptr};
    for (int i = 0; i < NO; ++i) {
        const size_t bytes = outs[i].unit * (size_t)(outs[i].per_guide ? T : n_genes);
        if ((!outs[i].host && !(i == 2 && opt.w1_stays)) || !bytes) continue;
        if ((rc = plan->b->out[i].reserve(bytes)) != JACKS_OK) { jacks_em_plan_destroy(plan); return rc; }
        dptr[i] = (char*)plan->b->out[i].p;
    }
    if (opt.plan_out) *opt.plan_out = plan;
    JacksEmDeviceIO dio;
    memset(&dio, 0, sizeof dio);
    dio.test = cc ? cc->d_test : (const double*)ctx->res_test.p; dio.ctrl = cc ? cc->d_ctrl : (const double*)ctx->res_ctrl.p;
    dio.ctrl_lines = io->ctrl_lines; dio.ctrl_rowmean_fill = io->ctrl_rowmean_fill;
    dio.fixed_x1 = d_fx1; dio.fixed_x2 = d_fx2;
    dio.x1 = (double*)dptr[0]; dio.x2 = (double*)dptr[1]; dio.w1 = (double*)dptr[2]; dio.w2 = (double*)dptr[3];
    dio.y = (double*)dptr[4]; dio.tau = (double*)dptr[5]; dio.n_iters = (int32_t*)dptr[6]; dio.bound = (double*)dptr[7];
    rc = (n_genes > 0) ? em_check_io(plan, params, &dio) : JACKS_OK;
    if (cc) cudaStreamWaitEvent(s_em, cc->ready, 0);         // the compact cubes are on their way (copy or gather stream)

    // ---- pipeline: gene ranges flow through upload -> kernels -> download on three streams ------------
    // Ranges are contiguous gene blocks.  When the guide rows ascend along the index (a real screen: genes own
    // consecutive rows) each range uploads just the rows up to its last one, so the first kernels start after
    // 1/R of the upload; otherwise (pseudo-genes, arbitrary gathers) everything is uploaded before the kernels.
    const int R = opt.pre ? opt.pre->R : (opt.ranges > 0 ? std::min(opt.ranges, kMaxRanges) : host_run_ranges(n_genes));
    // range r covers genes [gb(r), gb(r+1))
    auto gb = [&](int r) -> int64_t { return host_run_bound(n_genes, R, r, opt.kde != nullptr); };
    bool ascending = upload;
    if (upload && R > 1) {
        for (int64_t i = 1; i < T && ascending; ++i) ascending = guide_rows[i] >= guide_rows[i - 1];
    }
    // ev_in[r]: rows of range r are up; ev_em[r]: its EM has run (only when the EM has a stream of its own); ev_k[r]: all
    // its kernels have run (EM and p-values)
    std::vector<cudaEvent_t> ev_in(R, nullptr), ev_k(R, nullptr), ev_em(R, nullptr);
    for (int r = 0; r < R && rc == JACKS_OK; ++r) {
        if ((e = cudaEventCreateWithFlags(&ev_in[r], cudaEventDisableTiming)) != cudaSuccess) rc = cuda_fail(e, "cudaEventCreate");
        else if ((e = cudaEventCreateWithFlags(&ev_k[r], cudaEventDisableTiming)) != cudaSuccess) rc = cuda_fail(e, "cudaEventCreate");
        else if (split && (e = cudaEventCreateWithFlags(&ev_em[r], cudaEventDisableTiming)) != cudaSuccess) rc = cuda_fail(e, "cudaEventCreate");
    }
    if (rc == JACKS_OK && (e = cudaMemsetAsync(plan->b->d_counters.p, 0, kCounterInts * sizeof(int), s_em)) != cudaSuccess) rc = cuda_fail(e, "cudaMemsetAsync");
    cudaEvent_t ev_start = nullptr;
    if (rc == JACKS_OK && upload) {
        // the upload overwrites the resident cubes: it waits for the kernels of earlier calls that READ them -- and for
        // nothing else on the compute stream (a screen job's pseudo-gene EM and null densities, which read the compact
        // cube, run while the screen goes up)
        if (ctx->res_busy) cudaStreamWaitEvent(s_in, ctx->res_busy, 0);
    }
    int64_t row_done = 0;        // rows [0, row_done) are already enqueued for upload
    for (int r = 0; r < R && rc == JACKS_OK; ++r) {
        const int64_t g0 = gb(r), g1 = gb(r + 1);
        if (upload) {
            int64_t row_hi = N;
            if (ascending && r + 1 < R && g1 > g0) row_hi = std::min<int64_t>(N, (int64_t)guide_rows[gene_offsets[g1] - 1] + 1);
            else if (!ascending && r > 0) row_hi = row_done;
            if (row_hi > row_done) {
                // pinned caller memory goes up directly; pageable memory through the pinned ring with threaded host copies
                if ((rc = ctx->stager.h2d((char*)ctx->res_test.p + row_done * row_t, (const char*)io->test + row_done * row_t,
                                          (row_hi - row_done) * row_t, s_in)) != JACKS_OK) break;
                if ((rc = ctx->stager.h2d((char*)ctx->res_ctrl.p + row_done * row_c, (const char*)io->ctrl + row_done * row_c,
                                          (row_hi - row_done) * row_c, s_in)) != JACKS_OK) break;
                row_done = row_hi;
            }
            cudaEventRecord(ev_in[r], s_in);
            cudaStreamWaitEvent(s_em, ev_in[r], 0);
        } else if (opt.pre) {
            cudaStreamWaitEvent(s_em, opt.pre->ev_in[r], 0);
        }
        if (g1 > g0) rc = em_run_range(ctx, plan, params, &dio, s_em, g0, g1, r, opt.grid_limit, split);
        if (opt.trace && (r == 0 || r == R - 1 || r == R / 2)) opt.trace->mark((cc ? "pseudo EM range " : "real EM range ") + std::to_string(r), s_em);
        if (rc == JACKS_OK && split) {
            cudaEventRecord(ev_em[r], s_em);
            cudaStreamWaitEvent(s_k, ev_em[r], 0);
        }
        if (rc == JACKS_OK && opt.kde && g1 > g0 && dptr[8])
            rc = jacks::kde_eval(ctx, *opt.kde, (const double*)dptr[2] + g0 * L, g1 - g0, (double*)dptr[8] + g0 * L, s_k);
        if (rc != JACKS_OK) break;
        cudaEventRecord(ev_k[r], s_k);
        if (opt.trace && opt.kde && (r == 0 || r == R - 1)) opt.trace->mark("p-values range " + std::to_string(r), s_k);
    }
    // ---- downloads: everything but the p-values as soon as a range's EM has run, the p-values in a second sweep (with
    // the EM on its own stream they come much later: one in-order download stream must not park the other outputs
    // behind them).  Synchronous call: the host thread drains range r (through the pinned ring when the caller's arrays
    // are pageable) while the kernels of the later ranges run.
    const bool staged_out = !(keep_resident & JACKS_ASYNC);
    bool any_out[2] = {false, false};     // is there anything to download in sweep 0 / sweep 1?  (if not, the download stream
    for (int i = 0; i < NO; ++i)          //  must not be made to wait for this run's kernels)
        if (dptr[i] && outs[i].host) any_out[(split && i == 8) ? 1 : 0] = true;
    for (int sweep = 0; sweep < (split ? 2 : 1) && rc == JACKS_OK; ++sweep) {
        if (!any_out[sweep]) continue;
        cudaStream_t s_out = (sweep == 1) ? ctx->out2_stream : ctx->out_stream;      // the late p-values travel on a stream of their own
        for (int r = 0; r < R && rc == JACKS_OK; ++r) {
            const int64_t g0 = gb(r), g1 = gb(r + 1);
            const int64_t t0 = gene_offsets[g0], t1 = gene_offsets[g1];
            cudaStreamWaitEvent(s_out, (split && sweep == 0) ? ev_em[r] : ev_k[r], 0);
            for (int i = 0; i < NO && rc == JACKS_OK; ++i) {
                if (!dptr[i] || !outs[i].host) continue;
                if (split && (sweep == 0) == (i == 8)) continue;          // sweep 0: all but the p-values; sweep 1: the p-values
                const size_t lo = outs[i].unit * (size_t)(outs[i].per_guide ? t0 : g0), hi = outs[i].unit * (size_t)(outs[i].per_guide ? t1 : g1);
                if (hi <= lo) continue;
                if (staged_out) rc = ctx->stager.d2h(outs[i].host + lo, dptr[i] + lo, hi - lo, s_out);
                else if ((e = cudaMemcpyAsync(outs[i].host + lo, dptr[i] + lo, hi - lo, cudaMemcpyDeviceToHost, s_out)) != cudaSuccess) rc = cuda_fail(e, "D2H result");
            }
        }
        if (opt.trace) opt.trace->mark(std::string(cc ? "pseudo" : "real") + " downloads sweep " + std::to_string(sweep), s_out);
    }
    if (split && R > 0 && ev_em[R - 1]) cudaStreamWaitEvent(s_k, ev_em[R - 1], 0);      // res_busy below covers the EM stream
    for (int r = 0; r < R; ++r) if (ev_em[r]) ctx->pending_events.push_back(ev_em[r]);
    for (int r = 0; r < R; ++r) { if (ev_in[r]) ctx->pending_events.push_back(ev_in[r]); if (ev_k[r]) ctx->pending_events.push_back(ev_k[r]); }
    if (ev_start) ctx->pending_events.push_back(ev_start);
    if (!cc) {                    // the last kernel of this call that reads the resident cubes
        if (!ctx->res_busy && (e = cudaEventCreateWithFlags(&ctx->res_busy, cudaEventDisableTiming)) != cudaSuccess && rc == JACKS_OK) rc = cuda_fail(e, "cudaEventCreate");
        if (ctx->res_busy) cudaEventRecord(ctx->res_busy, s_k);
    }
    ctx->pending_plans.push_back(plan);
    if (!cc && !(keep_resident & JACKS_KEEP_RESIDENT)) ctx->res_valid = false;
    if (rc != JACKS_OK || !(keep_resident & JACKS_ASYNC)) {
        const int rs = jacks_ctx_synchronize(ctx);
        if (rc == JACKS_OK) rc = rs;
    }
    return rc;
}


// ---- the whole inference job of one screen (jacks_io.py:425-441) ------------------------------------------------------

extern "C" int jacks_run_screen_device(JacksCtx* ctx, const JacksEmPlan* plan_real, const JacksEmPlan* plan_pseudo,
                                       const JacksEmParams* params, const JacksScreenDeviceIO* io, void* stream) {
    if (!ctx || !plan_real || !params || !io) return fail(JACKS_E_INVALID, "jacks_run_screen_device: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    int rc;
    const int L = plan_real->L;
    const int64_t n_pseudo = plan_pseudo ? plan_pseudo->n_genes : 0;
    // The two EM launches are independent.  The real genes go first on the caller's stream; the pseudo-genes run beside
    // them on the context's side stream: both kernels are persistent (every SM full), so the pseudo-gene CTAs move in as
    // the real-gene kernel's CTAs drain and its tail costs nothing.  The caller's stream joins before the p-values.
    const bool both = n_pseudo > 0 && plan_real->n_genes > 0;
    cudaStream_t pst = both ? ctx->side_stream : st;
    if (both) {
        if ((e = cudaEventRecord(ctx->side_fork, st)) != cudaSuccess) return cuda_fail(e, "cudaEventRecord");
        if ((e = cudaStreamWaitEvent(ctx->side_stream, ctx->side_fork, 0)) != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent");
    }
    if (plan_real->n_genes > 0) {
        if ((rc = em_check_io(plan_real, params, &io->real)) != JACKS_OK) return rc;
        if ((e = cudaMemsetAsync(plan_real->b->d_counters.p, 0, kCountersPerRange * sizeof(int), st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
        if ((rc = em_run_range(ctx, plan_real, params, &io->real, st, 0, plan_real->n_genes, 0)) != JACKS_OK) return rc;
    }
    if (n_pseudo > 0) {
        if (plan_pseudo->L != L || plan_pseudo->n_rows != plan_real->n_rows) return fail(JACKS_E_INVALID, "the two plans describe different cubes");
        if (!io->pseudo_w1) return fail(JACKS_E_INVALID, "pseudo_w1 is required when there are pseudo-genes");
        // pseudo-genes: same cubes, never fixed_x (jacks_io.py:430); only E[w] (and what the caller asks for) is kept
        JacksEmDeviceIO pio;
        memset(&pio, 0, sizeof pio);
        pio.test = io->real.test; pio.ctrl = io->real.ctrl; pio.ctrl_lines = io->real.ctrl_lines;
        pio.ctrl_rowmean_fill = io->real.ctrl_rowmean_fill;
        pio.w1 = io->pseudo_w1; pio.w2 = io->pseudo_w2; pio.n_iters = io->pseudo_n_iters;
        if ((rc = em_check_io(plan_pseudo, params, &pio)) != JACKS_OK) return rc;
        if ((e = cudaMemsetAsync(plan_pseudo->b->d_counters.p, 0, kCountersPerRange * sizeof(int), pst)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
        rc = em_run_range(ctx, plan_pseudo, params, &pio, pst, 0, n_pseudo, 0);
        if (both) {        // join even after a failed launch: the side stream must not be left dangling behind the caller's
            cudaEventRecord(ctx->side_join, ctx->side_stream);
            cudaStreamWaitEvent(st, ctx->side_join, 0);
        }
        if (rc != JACKS_OK) return rc;
    }
    if (io->pvals || io->pseudo_pvals) {
        if (n_pseudo < 2) return fail(JACKS_E_INVALID, "p-values need at least two pseudo-genes (got %lld)", (long long)n_pseudo);
        jacks::KdeNull nl;
        if ((rc = jacks::kde_null_prepare(ctx, io->pseudo_w1, n_pseudo, L, io->positive, st, &nl)) != JACKS_OK) return rc;
        if (io->pvals && (rc = jacks::kde_eval(ctx, nl, io->real.w1, plan_real->n_genes, io->pvals, st)) != JACKS_OK) return rc;
        if (io->pseudo_pvals && (rc = jacks::kde_eval(ctx, nl, io->pseudo_w1, n_pseudo, io->pseudo_pvals, st)) != JACKS_OK) return rc;
    }
    return JACKS_OK;
}

extern "C" int jacks_run_screen_host(JacksCtx* ctx, const JacksEmParams* params, const int64_t* gene_offsets,
                                     const int32_t* guide_rows, int64_t n_genes, const int64_t* pseudo_offsets,
                                     const int32_t* pseudo_rows, int64_t n_pseudo, const JacksScreenHostIO* io, int32_t flags) {
    if (!ctx || !params || !io) return fail(JACKS_E_INVALID, "jacks_run_screen_host: NULL argument");
    const JacksEmHostIO& rio = io->real;
    if (!rio.test || !rio.ctrl) return fail(JACKS_E_INVALID, "test and ctrl are required");
    if (rio.n_rows < 0 || rio.n_lines < 1) return fail(JACKS_E_INVALID, "bad data shape N=%lld L=%d", (long long)rio.n_rows, rio.n_lines);
    if (rio.ctrl_lines != rio.n_lines && rio.ctrl_lines != 1)
        return fail(JACKS_E_SHAPE, "Expecting matched size between control and test data");
    if (n_pseudo < 0 || (n_pseudo > 0 && (!pseudo_offsets || !pseudo_rows))) return fail(JACKS_E_INVALID, "NULL pseudo-gene index");
    const bool want_p = io->pvals || io->pseudo_pvals;
    if (want_p && n_pseudo < 2) return fail(JACKS_E_INVALID, "p-values need at least two pseudo-genes (got %lld)", (long long)n_pseudo);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    const int L = rio.n_lines;
    int rc;
    jacks::KdeNull nl;
    JacksEmPlan* pplan = nullptr;
    PreUpload pre;
    // Tuning / A-B knobs.  Defaults: the screen starts going up before anything else; the real genes' EM runs on the side
    // stream BESIDE the pseudo-gene EM, which leaves it 13.5 % of the SM slots (the real genes are 15 % of the EM work and
    // arrive over the whole upload; run behind the pseudo-genes, their 16 small launches cost 3 ms after the upload's end).
    static const int upload_first_env = []() { const char* v = getenv("JACKS_SCREEN_UPLOAD_FIRST"); return v ? atoi(v) : 1; }();
    static const int split_env = []() { const char* v = getenv("JACKS_SCREEN_SPLIT"); return v ? atoi(v) : 1; }();
    static const int reserve_env = []() { const char* v = getenv("JACKS_SCREEN_RESERVE"); return v ? atoi(v) : 135; }();
    static const bool trace = getenv("JACKS_TRACE") != nullptr;
    g_trace_t0 = std::chrono::steady_clock::now();
    g_trace_on = trace;
    auto stamp = [&](const char* what) { trace_stamp(what); };
    const bool upload_first = upload_first_env && n_pseudo > 0 && host_pointer_is_pinned(rio.test) && host_pointer_is_pinned(rio.ctrl);
    TraceMarks tm;
    if (trace) { tm.start(ctx->stream); pre.trace = &tm; }
    if (upload_first) {
        // 0. pinned cubes: the screen starts going up (copy stream, range by range) before any host-side preparation of
        //    the pseudo-gene run; the control rows the pseudo-genes need are pulled by a kernel, not by the copy engine
        if ((rc = enqueue_screen_upload(ctx, &rio, gene_offsets, guide_rows, n_genes, io->pvals != nullptr, &pre)) != JACKS_OK) { jacks_ctx_synchronize(ctx); return rc; }
        stamp("screen upload enqueued");
    }
    if (n_pseudo > 0) {
        // 1. pseudo-genes on a compact cube of the rows they reference, uploaded ahead of the screen
        JacksEmHostIO pio;
        memset(&pio, 0, sizeof pio);
        pio.test = rio.test; pio.ctrl = rio.ctrl; pio.n_rows = rio.n_rows; pio.n_lines = L; pio.ctrl_lines = rio.ctrl_lines;
        pio.ctrl_rowmean_fill = rio.ctrl_rowmean_fill;
        pio.w1 = io->pseudo_w1; pio.w2 = io->pseudo_w2; pio.n_iters = io->pseudo_n_iters;
        HostRunOpts popt;
        popt.flags = JACKS_ASYNC;
        popt.w1_stays = true;
        popt.plan_out = &pplan;
        if (trace) popt.trace = &tm;
        if (split_env && n_genes > 0 && reserve_env > 0 && reserve_env < 900) popt.grid_limit = -reserve_env;
        if (!io->pseudo_w1 && !io->pseudo_w2 && !io->pseudo_n_iters) popt.ranges = 1;     // nothing comes down: one launch
        if ((rc = em_batched_host_compact(ctx, params, pseudo_offsets, pseudo_rows, n_pseudo, &pio, popt)) != JACKS_OK) { jacks_ctx_synchronize(ctx); return rc; }
        stamp("pseudo-gene EM enqueued");
        if (trace) { tm.mark("compact cube gathered", ctx->cmp_stream); tm.mark("pseudo EM done", ctx->stream); }
        // 2. null densities from the pseudo-gene E[w], still on the device
        if (want_p && (rc = jacks::kde_null_prepare(ctx, (const double*)pplan->b->out[2].p, n_pseudo, L, io->positive, ctx->stream, &nl)) != JACKS_OK) { jacks_ctx_synchronize(ctx); return rc; }
        if (io->pseudo_pvals) {
            if ((rc = pplan->b->out[8].reserve((size_t)n_pseudo * L * sizeof(double))) == JACKS_OK)
                rc = jacks::kde_eval(ctx, nl, (const double*)pplan->b->out[2].p, n_pseudo, (double*)pplan->b->out[8].p, ctx->stream);
            if (rc == JACKS_OK) {
                cudaEvent_t ev = nullptr;
                if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) rc = cuda_fail(e, "cudaEventCreate");
                else {
                    cudaEventRecord(ev, ctx->stream); cudaStreamWaitEvent(ctx->out_stream, ev, 0); ctx->pending_events.push_back(ev);
                    if ((e = cudaMemcpyAsync(io->pseudo_pvals, pplan->b->out[8].p, (size_t)n_pseudo * L * sizeof(double), cudaMemcpyDeviceToHost, ctx->out_stream)) != cudaSuccess) rc = cuda_fail(e, "D2H pseudo p-values");
                }
            }
            if (rc != JACKS_OK) { jacks_ctx_synchronize(ctx); return rc; }
        }
    }
    // 3. the screen itself: upload once, real genes range by range (EM, p-values, download)
    HostRunOpts ropt;
    ropt.flags = flags;
    if (upload_first) ropt.pre = &pre;
    if (trace) { ropt.trace = &tm; tm.mark("null densities done", ctx->stream); }
    if (split_env && n_pseudo > 0) ropt.em_stream = ctx->side_stream;
    if (io->pvals) { ropt.kde = &nl; ropt.pvals_host = io->pvals; }
    stamp("null densities enqueued");
    rc = em_batched_host_core(ctx, params, gene_offsets, guide_rows, n_genes, &rio, ropt);
    stamp("real genes done / enqueued");
    if (rc != JACKS_OK || !(flags & JACKS_ASYNC)) {
        const int rs = jacks_ctx_synchronize(ctx);
        if (rc == JACKS_OK) rc = rs;
    }
    stamp("synchronised");
    if (trace && !(flags & JACKS_ASYNC)) tm.print();
    return rc;
}

extern "C" {

// ---- per-line ("single JACKS") inference ----------------------------------------------------------------------------

// Enqueue the per-line EM of every (gene, line) pair of the plan on `st`.
static int em_lines_run(JacksCtx* ctx, const JacksEmPlan* plan, const JacksEmParams* params, const JacksEmDeviceIO* io,
                        cudaStream_t st) {
    const int L = plan->L;
    if ((long long)plan->n_genes * L >= (1LL << 31)) return fail(JACKS_E_INVALID, "per-line run: n_genes * L must stay below 2^31");
    cudaError_t e;
    int rc;
    // the deferred queue holds (gene, line) pairs here: one slot per problem of the fast buckets
    if ((rc = plan->b->d_deferred.reserve((size_t)std::max<int64_t>(plan->n_genes * L, 1) * sizeof(int))) != JACKS_OK) return rc;
    if ((e = cudaMemsetAsync(plan->b->d_counters.p, 0, kCountersPerRange * sizeof(int), st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    if (plan->b->ready && (e = cudaStreamWaitEvent(st, plan->b->ready, 0)) != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent(plan)");
    int* counters = (int*)plan->b->d_counters.p;
    EmLaunch a;
    memset(&a, 0, sizeof a);
    a.test = (const double2*)io->test; a.ctrl = (const double2*)io->ctrl;
    a.L = L; a.ctrl_lines = io->ctrl_lines; a.ctrl_rowmean_fill = 0;
    a.gene_offsets = (const long long*)plan->b->d_offsets.p; a.guide_rows = (const int*)plan->b->d_rows.p;
    a.fixed_x1 = io->fixed_x1; a.fixed_x2 = io->fixed_x2;
    a.x1 = io->x1; a.x2 = io->x2; a.w1 = io->w1; a.w2 = io->w2; a.y = io->y; a.tau = io->tau;
    a.n_iters = io->n_iters; a.bound = io->bound;
    a.deferred_ids = (int*)plan->b->d_deferred.p; a.deferred_count = counters;
    a.n_iter = params->n_iter; a.apply_w_hp = 0; a.tol = params->tol;      // infer.py:92 needs L > 1
    a.mu0_x = params->mu0_x; a.var0_x = params->var0_x; a.mu0_w = params->mu0_w; a.var0_w = params->var0_w;
    a.tps = params->tau_prior_strength;
    a.L_full = L;
    int slot = 1;
    int64_t n_fast = 0;
    for (const EmBucket& b : plan->fast) {
        a.gene_ids = (const int*)plan->b->d_ids.p + b.ids_off;
        a.slot_off = plan->slot_off + b.ids_off;
        a.slot_rows = plan->slot_rows + b.rows_off;
        a.n_work = b.n; a.n_work_dev = nullptr;
        e = launch_em_lines(b.G, a, ctx->n_sm, st);
        if (e != cudaSuccess) return cuda_fail(e, "em_lines_kernel launch");
        ctx->launches++;
        n_fast += b.n;
    }
    // NaN-safe warp-per-problem kernel on one-line views: genes without a fast kernel, then the problems handed back
    a.L = 1;
    a.scratch = (double*)plan->b->d_scratch.p;
    EmGenericConfig cfg;
    // (with one line a gene's working set is small whatever its guide count: the big-gene bucket runs here too)
    const EmBucket* rest[2] = {&plan->generic, &plan->big};
    const int rest_maxG[2] = {plan->generic_maxG, plan->big_maxG};
    for (int k = 0; k < 2; ++k) {
        if (rest[k]->n == 0) continue;
        if ((e = em_generic_configure(rest_maxG[k], 1, (long long)rest[k]->n * L, ctx->n_sm, &cfg)) != cudaSuccess) return cuda_fail(e, "em_generic_configure");
        if (cfg.scratch_doubles_per_warp > 0 &&
            (rc = plan->b->d_scratch.reserve((size_t)cfg.scratch_doubles_per_warp * cfg.grid * cfg.warps_per_cta * sizeof(double))) != JACKS_OK) return rc;
        a.scratch = (double*)plan->b->d_scratch.p;
        a.lines_mode = 1;
        a.gene_ids = (const int*)plan->b->d_ids.p + rest[k]->ids_off;
        a.n_work = rest[k]->n; a.n_work_dev = nullptr;
        a.work_counter = counters + slot++;
        a.scratch_stride = cfg.scratch_doubles_per_warp;
        if ((e = launch_em_generic(cfg, a, st)) != cudaSuccess) return cuda_fail(e, "em_generic_kernel (lines) launch");
        ctx->launches++;
    }
    if (n_fast > 0) {
        if ((e = em_generic_configure(plan->fast_maxG, 1, n_fast * L, ctx->n_sm, &cfg)) != cudaSuccess) return cuda_fail(e, "em_generic_configure");
        a.lines_mode = 2;
        a.gene_ids = (const int*)plan->b->d_deferred.p;
        a.n_work = (int)(n_fast * L); a.n_work_dev = counters;
        a.work_counter = counters + slot++;
        a.scratch_stride = cfg.scratch_doubles_per_warp;      // 0: a fast-kernel gene with one line always fits shared memory
        if ((e = launch_em_generic(cfg, a, st)) != cudaSuccess) return cuda_fail(e, "em_generic_kernel (lines, deferred) launch");
        ctx->launches++;
    }
    return JACKS_OK;
}

int jacks_em_lines_device(JacksCtx* ctx, const JacksEmPlan* plan, const JacksEmParams* params,
                          const JacksEmDeviceIO* io, void* stream) {
    if (!ctx || !plan || !params || !io) return fail(JACKS_E_INVALID, "jacks_em_lines_device: NULL argument");
    if (plan->n_genes == 0) return JACKS_OK;
    int rc = em_check_io(plan, params, io);
    if (rc != JACKS_OK) return rc;
    if (io->ctrl_lines != plan->L) return fail(JACKS_E_SHAPE, "Expecting matched size between control and test data");   // infer.py:26
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return em_lines_run(ctx, plan, params, io, (cudaStream_t)stream);
}

int jacks_em_lines_host(JacksCtx* ctx, const JacksEmParams* params, const int64_t* gene_offsets,
                        const int32_t* guide_rows, int64_t n_genes, const JacksEmHostIO* io, int32_t flags) {
    if (!ctx || !params || !io) return fail(JACKS_E_INVALID, "jacks_em_lines_host: NULL argument");
    const int64_t N = io->n_rows;
    const int L = io->n_lines;
    if (N < 0 || L < 1) return fail(JACKS_E_INVALID, "bad data shape N=%lld L=%d", (long long)N, L);
    if (io->ctrl_lines != L) return fail(JACKS_E_SHAPE, "Expecting matched size between control and test data");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    int rc = jacks_ctx_synchronize(ctx);          // this entry point is synchronous: drain earlier async calls first
    if (rc != JACKS_OK) return rc;
    cudaStream_t st = ctx->stream;
    const size_t cube = (size_t)N * L * 2 * sizeof(double);
    if (io->test) {
        if (!io->ctrl) return fail(JACKS_E_INVALID, "ctrl is NULL");
        if ((rc = ctx->res_test.reserve(cube ? cube : 8)) != JACKS_OK) return rc;
        if ((rc = ctx->res_ctrl.reserve(cube ? cube : 8)) != JACKS_OK) return rc;
        ctx->res_N = N; ctx->res_L = L; ctx->res_ctrl_lines = L; ctx->res_valid = true;
        if (cube) {
            if ((rc = ctx->stager.h2d(ctx->res_test.p, io->test, cube, st)) != JACKS_OK) return rc;
            if ((rc = ctx->stager.h2d(ctx->res_ctrl.p, io->ctrl, cube, st)) != JACKS_OK) return rc;
        }
    } else if (!ctx->res_valid || ctx->res_N != N || ctx->res_L != L || ctx->res_ctrl_lines != L) {
        return fail(JACKS_E_INVALID, "test is NULL but no matching resident data cube (call with keep_resident first)");
    }
    const double* d_fx1 = nullptr; const double* d_fx2 = nullptr;
    if (io->fixed_x1 || io->fixed_x2) {
        if (!io->fixed_x1 || !io->fixed_x2) return fail(JACKS_E_INVALID, "fixed_x1 and fixed_x2 go together");
        if ((rc = ctx->res_fx.reserve((size_t)N * 2 * sizeof(double) + 8)) != JACKS_OK) return rc;
        if (N) {
            if ((e = cudaMemcpyAsync(ctx->res_fx.p, io->fixed_x1, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D fixed_x1");
            if ((e = cudaMemcpyAsync((double*)ctx->res_fx.p + N, io->fixed_x2, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D fixed_x2");
        }
        d_fx1 = (const double*)ctx->res_fx.p; d_fx2 = d_fx1 + N;
    }
    if (!io->w1 && n_genes) return fail(JACKS_E_INVALID, "w1 is required");
    JacksEmPlan* plan = nullptr;
    if ((rc = jacks_em_plan_create(ctx, gene_offsets, guide_rows, n_genes, N, L, &plan)) != JACKS_OK) return rc;
    trace_stamp("  core: plan created");
    const int64_t T = plan->total_guides;
    // every output has L entries per gene or per guide
    struct Out { char* host; size_t unit; bool per_guide; } outs[8] = {
        {(char*)io->x1, sizeof(double), true},  {(char*)io->x2, sizeof(double), true},
        {(char*)io->w1, sizeof(double), false}, {(char*)io->w2, sizeof(double), false},
        {(char*)io->y, sizeof(double), true},   {(char*)io->tau, sizeof(double), true},
        {(char*)io->n_iters, sizeof(int32_t), false}, {(char*)io->bound, sizeof(double), false}};
    char* dptr[8] = {nullptr};
    size_t nbytes[8] = {0};
    for (int i = 0; i < 8 && rc == JACKS_OK; ++i) {
        nbytes[i] = outs[i].unit * (size_t)L * (size_t)(outs[i].per_guide ? T : n_genes);
        if (!outs[i].host || !nbytes[i]) continue;
        if ((rc = plan->b->out[i].reserve(nbytes[i])) == JACKS_OK) dptr[i] = (char*)plan->b->out[i].p;
    }
    if (rc == JACKS_OK && n_genes > 0) {
        JacksEmDeviceIO dio;
        memset(&dio, 0, sizeof dio);
        dio.test = (const double*)ctx->res_test.p; dio.ctrl = (const double*)ctx->res_ctrl.p;
        dio.ctrl_lines = L;
        dio.fixed_x1 = d_fx1; dio.fixed_x2 = d_fx2;
        dio.x1 = (double*)dptr[0]; dio.x2 = (double*)dptr[1]; dio.w1 = (double*)dptr[2]; dio.w2 = (double*)dptr[3];
        dio.y = (double*)dptr[4]; dio.tau = (double*)dptr[5]; dio.n_iters = (int32_t*)dptr[6]; dio.bound = (double*)dptr[7];
        if ((rc = em_check_io(plan, params, &dio)) == JACKS_OK) rc = em_lines_run(ctx, plan, params, &dio, st);
        for (int i = 0; i < 8 && rc == JACKS_OK; ++i)
            if (dptr[i]) rc = ctx->stager.d2h(outs[i].host, dptr[i], nbytes[i], st);
    }
    ctx->pending_plans.push_back(plan);
    if (!(flags & JACKS_KEEP_RESIDENT)) ctx->res_valid = false;
    const int rs = jacks_ctx_synchronize(ctx);
    return rc != JACKS_OK ? rc : rs;
}

}  // extern "C"
