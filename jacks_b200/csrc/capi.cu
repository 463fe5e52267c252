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
#include <string>
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

// ------------------------------------------------------------------------------------------------
// Plan
// ------------------------------------------------------------------------------------------------
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
    DevBuf d_offsets, d_rows, d_ids, d_slot_off, d_slot_rows, d_deferred, d_counters, d_scratch, d_tpd;
    std::vector<EmBucket> fast;
    EmBucket generic;          // n == 0 if empty
    int generic_maxG;
    int fast_maxG;
    int n_fast_genes;
    EmGenericConfig cfg_generic, cfg_deferred;
};

namespace {

// Warps per gene of the fast kernel for (G, L); 0 = use the generic kernel.
int choose_team(int G, int L) {
    if (G < 1 || G > 8) return 0;
    const char* env = getenv("JACKS_EM_W");          // tuning override
    int W = 0;
    if (env) W = atoi(env);
    if (!(W == 1 || W == 2 || W == 4)) {
        // One warp per gene is the most efficient split (no barriers, one reduction per pass) and, at
        // the shared-memory-bound occupancy of a few genes per SM, also the fastest measured at L = 400
        // (W=1 9.2 ms, W=2 10.1 ms, W=4 15.3 ms per Avana pass).  Very long line axes get more warps.
        const int strips = (L + 31) / 32;
        W = strips > 64 ? 4 : (strips > 32 ? 2 : 1);
    }
    if (!em_fast_available(G, W) || em_fast_smem_bytes(G, W, L) > kMaxDynSmem) return 0;
    return W;
}

}  // namespace

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
    *out = c;
    return JACKS_OK;
}

void jacks_ctx_destroy(JacksCtx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->res_test.release(); c->res_ctrl.release(); c->res_fx.release();
    for (auto& b : c->out) b.release();
    c->tmp.release();
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

int64_t jacks_ctx_launch_count(const JacksCtx* c) { return c ? c->launches : 0; }

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

int jacks_em_plan_create(JacksCtx* ctx, const int64_t* gene_offsets, const int32_t* guide_rows, int64_t n_genes,
                         int64_t n_rows, int32_t n_lines, JacksEmPlan** out) {
    if (!ctx || !out) return fail(JACKS_E_INVALID, "jacks_em_plan_create: NULL ctx/out");
    *out = nullptr;
    if (n_genes < 0 || n_rows < 0 || n_lines < 1) return fail(JACKS_E_INVALID, "bad sizes: n_genes=%lld n_rows=%lld n_lines=%d", (long long)n_genes, (long long)n_rows, n_lines);
    if (n_genes > 0 && (!gene_offsets || !guide_rows)) return fail(JACKS_E_INVALID, "NULL gene index");
    if (n_genes >= (1LL << 31)) return fail(JACKS_E_INVALID, "too many genes");
    const int64_t total = n_genes ? gene_offsets[n_genes] : 0;
    if (n_genes && gene_offsets[0] != 0) return fail(JACKS_E_INVALID, "gene_offsets[0] must be 0");
    int maxG_all = 0;
    for (int64_t g = 0; g < n_genes; ++g) {
        const int64_t G = gene_offsets[g + 1] - gene_offsets[g];
        if (G < 1) return fail(JACKS_E_INVALID, "gene %lld has %lld guides; every gene needs at least one", (long long)g, (long long)G);
        if (G > (1 << 20)) return fail(JACKS_E_INVALID, "gene %lld has too many guides", (long long)g);
        maxG_all = std::max<int>(maxG_all, (int)G);
    }
    for (int64_t i = 0; i < total; ++i)
        if (guide_rows[i] < 0 || guide_rows[i] >= n_rows)
            return fail(JACKS_E_INVALID, "guide row %d at position %lld outside [0, %lld)", guide_rows[i], (long long)i, (long long)n_rows);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");

    JacksEmPlan* p = new JacksEmPlan();
    p->ctx = ctx; p->n_genes = n_genes; p->n_rows = n_rows; p->total_guides = total; p->L = n_lines;
    p->generic = EmBucket{0, 0, 0, 0, 0};
    p->generic_maxG = 0; p->fast_maxG = 0; p->n_fast_genes = 0;

    // bucket genes by guide count, keeping the caller's order inside each bucket
    std::vector<std::vector<int>> byG(9);
    std::vector<int> gen;
    int teamW[9] = {0};
    for (int G = 1; G <= 8; ++G) teamW[G] = choose_team(G, n_lines);
    for (int64_t g = 0; g < n_genes; ++g) {
        const int G = (int)(gene_offsets[g + 1] - gene_offsets[g]);
        if (G <= 8 && teamW[G] > 0) byG[G].push_back((int)g);
        else { gen.push_back((int)g); p->generic_maxG = std::max(p->generic_maxG, G); }
    }
    std::vector<int> ids;
    ids.reserve((size_t)n_genes);
    std::vector<long long> slot_off;       // gene_offsets[] in bucket-slot order
    std::vector<int> slot_rows;            // guide rows in bucket-slot order (fast buckets)
    slot_off.reserve((size_t)n_genes);
    slot_rows.reserve((size_t)total);
    // larger G first: the heavier buckets start while the GPU is empty
    for (int G = 8; G >= 1; --G) {
        if (byG[G].empty()) continue;
        EmBucket b{G, teamW[G], (int)byG[G].size(), (int)ids.size(), (long long)slot_rows.size()};
        ids.insert(ids.end(), byG[G].begin(), byG[G].end());
        for (int g : byG[G]) {
            slot_off.push_back(gene_offsets[g]);
            for (int k = 0; k < G; ++k) slot_rows.push_back(guide_rows[gene_offsets[g] + k]);
        }
        p->fast.push_back(b);
        p->fast_maxG = std::max(p->fast_maxG, G);
        p->n_fast_genes += b.n;
    }
    if (!gen.empty()) {
        p->generic = EmBucket{0, 0, (int)gen.size(), (int)ids.size(), 0};
        ids.insert(ids.end(), gen.begin(), gen.end());
    }

    int rc = JACKS_OK;
    auto up = [&](DevBuf& b, const void* src, size_t bytes) {
        if (rc != JACKS_OK) return;
        rc = b.reserve(bytes ? bytes : 8);
        if (rc == JACKS_OK && bytes) {
            cudaError_t ee = cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
            if (ee != cudaSuccess) rc = cuda_fail(ee, "cudaMemcpyAsync(plan)");
        }
    };
    up(p->d_offsets, gene_offsets, (size_t)(n_genes + 1) * sizeof(int64_t));
    up(p->d_rows, guide_rows, (size_t)total * sizeof(int32_t));
    up(p->d_ids, ids.data(), ids.size() * sizeof(int));
    up(p->d_slot_off, slot_off.data(), slot_off.size() * sizeof(long long));
    up(p->d_slot_rows, slot_rows.data(), slot_rows.size() * sizeof(int));
    if (rc == JACKS_OK) rc = p->d_deferred.reserve((size_t)std::max<int64_t>(n_genes, 1) * sizeof(int));
    if (rc == JACKS_OK) rc = p->d_counters.reserve(64 * sizeof(int));
    if (rc == JACKS_OK && p->fast_maxG > 0)
        rc = p->d_tpd.reserve((size_t)ctx->n_sm * kMaxFastCtasPerSm * p->fast_maxG * n_lines * sizeof(double));
    // generic kernel configurations (and global scratch if a gene's working set exceeds shared memory)
    memset(&p->cfg_generic, 0, sizeof p->cfg_generic);
    memset(&p->cfg_deferred, 0, sizeof p->cfg_deferred);
    long long scratch_doubles = 0;
    if (rc == JACKS_OK && p->generic.n > 0) {
        e = em_generic_configure(p->generic_maxG, n_lines, p->generic.n, ctx->n_sm, &p->cfg_generic);
        if (e != cudaSuccess) rc = cuda_fail(e, "em_generic_configure");
        else scratch_doubles = std::max(scratch_doubles, p->cfg_generic.scratch_doubles_per_warp * p->cfg_generic.grid * p->cfg_generic.warps_per_cta);
    }
    if (rc == JACKS_OK && p->n_fast_genes > 0) {
        e = em_generic_configure(p->fast_maxG, n_lines, p->n_fast_genes, ctx->n_sm, &p->cfg_deferred);
        if (e != cudaSuccess) rc = cuda_fail(e, "em_generic_configure");
        else scratch_doubles = std::max(scratch_doubles, p->cfg_deferred.scratch_doubles_per_warp * p->cfg_deferred.grid * p->cfg_deferred.warps_per_cta);
    }
    if (rc == JACKS_OK && scratch_doubles > 0) rc = p->d_scratch.reserve((size_t)scratch_doubles * sizeof(double));
    if (rc == JACKS_OK) {
        e = cudaStreamSynchronize(ctx->stream);      // host vectors go out of scope
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize(plan)");
    }
    if (rc != JACKS_OK) { jacks_em_plan_destroy(p); return rc; }
    *out = p;
    return JACKS_OK;
}

void jacks_em_plan_destroy(JacksEmPlan* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    p->d_offsets.release(); p->d_rows.release(); p->d_ids.release(); p->d_slot_off.release();
    p->d_slot_rows.release(); p->d_deferred.release(); p->d_tpd.release();
    p->d_counters.release(); p->d_scratch.release();
    delete p;
}

int jacks_em_run_device(JacksCtx* ctx, const JacksEmPlan* plan, const JacksEmParams* params,
                        const JacksEmDeviceIO* io, void* stream) {
    if (!ctx || !plan || !params || !io) return fail(JACKS_E_INVALID, "jacks_em_run_device: NULL argument");
    if (plan->n_genes == 0) return JACKS_OK;
    if (!io->test || !io->ctrl || !io->w1) return fail(JACKS_E_INVALID, "test, ctrl and w1 are required");
    if (io->ctrl_lines != plan->L && io->ctrl_lines != 1)
        return fail(JACKS_E_SHAPE, "Expecting matched size between control and test data");   // infer.py:26
    if ((io->fixed_x1 == nullptr) != (io->fixed_x2 == nullptr)) return fail(JACKS_E_INVALID, "fixed_x1 and fixed_x2 go together");
    if (params->n_iter < 0) return fail(JACKS_E_INVALID, "n_iter < 0");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");

    int* counters = (int*)plan->d_counters.p;     // [0] deferred count, [1..] one work counter per launch
    if ((e = cudaMemsetAsync(counters, 0, 64 * sizeof(int), st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");

    EmLaunch a;
    memset(&a, 0, sizeof a);
    a.test = (const double2*)io->test; a.ctrl = (const double2*)io->ctrl;
    a.L = plan->L; a.ctrl_lines = io->ctrl_lines; a.ctrl_rowmean_fill = io->ctrl_rowmean_fill;
    a.gene_offsets = (const long long*)plan->d_offsets.p; a.guide_rows = (const int*)plan->d_rows.p;
    a.fixed_x1 = io->fixed_x1; a.fixed_x2 = io->fixed_x2;
    a.x1 = io->x1; a.x2 = io->x2; a.w1 = io->w1; a.w2 = io->w2; a.y = io->y; a.tau = io->tau;
    a.n_iters = io->n_iters; a.bound = io->bound;
    a.deferred_ids = (int*)plan->d_deferred.p; a.deferred_count = counters;
    a.n_iter = params->n_iter; a.apply_w_hp = params->apply_w_hp; a.tol = params->tol;
    a.mu0_x = params->mu0_x; a.var0_x = params->var0_x; a.mu0_w = params->mu0_w; a.var0_w = params->var0_w;
    a.tps = params->tau_prior_strength;
    a.scratch = (double*)plan->d_scratch.p;
    a.tpd_scratch = (double*)plan->d_tpd.p;

    int slot = 1;
    for (const EmBucket& b : plan->fast) {
        a.gene_ids = (const int*)plan->d_ids.p + b.ids_off;
        a.slot_off = (const long long*)plan->d_slot_off.p + b.ids_off;
        a.slot_rows = (const int*)plan->d_slot_rows.p + b.rows_off;
        a.n_work = b.n; a.n_work_dev = nullptr;
        a.work_counter = counters + slot++;
        int grid = 0;
        e = launch_em_fast(b.G, b.W, a, ctx->n_sm, 0, st, &grid);
        if (e != cudaSuccess) return cuda_fail(e, "em_fast_kernel launch");
        ctx->launches++;
    }
    if (plan->generic.n > 0) {
        a.gene_ids = (const int*)plan->d_ids.p + plan->generic.ids_off;
        a.n_work = plan->generic.n; a.n_work_dev = nullptr;
        a.work_counter = counters + slot++;
        a.scratch_stride = plan->cfg_generic.scratch_doubles_per_warp;
        e = launch_em_generic(plan->cfg_generic, a, st);
        if (e != cudaSuccess) return cuda_fail(e, "em_generic_kernel launch");
        ctx->launches++;
    }
    if (plan->n_fast_genes > 0) {
        // genes the fast kernels handed back because their data holds NaN/inf
        a.gene_ids = (const int*)plan->d_deferred.p;
        a.n_work = plan->n_fast_genes; a.n_work_dev = counters;
        a.work_counter = counters + slot++;
        a.scratch_stride = plan->cfg_deferred.scratch_doubles_per_warp;
        e = launch_em_generic(plan->cfg_deferred, a, st);
        if (e != cudaSuccess) return cuda_fail(e, "em_generic_kernel (deferred) launch");
        ctx->launches++;
    }
    return JACKS_OK;
}

int jacks_em_batched_host(JacksCtx* ctx, const JacksEmParams* params, const int64_t* gene_offsets,
                          const int32_t* guide_rows, int64_t n_genes, const JacksEmHostIO* io, int32_t keep_resident) {
    if (!ctx || !params || !io) return fail(JACKS_E_INVALID, "jacks_em_batched_host: NULL argument");
    const int64_t N = io->n_rows;
    const int L = io->n_lines;
    if (N < 0 || L < 1) return fail(JACKS_E_INVALID, "bad data shape N=%lld L=%d", (long long)N, L);
    if (io->ctrl_lines != L && io->ctrl_lines != 1)
        return fail(JACKS_E_SHAPE, "Expecting matched size between control and test data");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    cudaStream_t st = ctx->stream;
    int rc;

    // ---- inputs: upload, or reuse the cubes a previous call left resident -------------------------
    const size_t test_bytes = (size_t)N * L * 2 * sizeof(double);
    const size_t ctrl_bytes = (size_t)N * io->ctrl_lines * 2 * sizeof(double);
    if (io->test) {
        if (!io->ctrl) return fail(JACKS_E_INVALID, "ctrl is NULL");
        if ((rc = ctx->res_test.reserve(test_bytes ? test_bytes : 8)) != JACKS_OK) return rc;
        if ((rc = ctx->res_ctrl.reserve(ctrl_bytes ? ctrl_bytes : 8)) != JACKS_OK) return rc;
        if (test_bytes && (e = cudaMemcpyAsync(ctx->res_test.p, io->test, test_bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D test");
        if (ctrl_bytes && (e = cudaMemcpyAsync(ctx->res_ctrl.p, io->ctrl, ctrl_bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D ctrl");
        ctx->res_N = N; ctx->res_L = L; ctx->res_ctrl_lines = io->ctrl_lines; ctx->res_valid = true;
    } else {
        if (!ctx->res_valid || ctx->res_N != N || ctx->res_L != L || ctx->res_ctrl_lines != io->ctrl_lines)
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

    JacksEmPlan* plan = nullptr;
    if ((rc = jacks_em_plan_create(ctx, gene_offsets, guide_rows, n_genes, N, L, &plan)) != JACKS_OK) return rc;
    const int64_t T = plan->total_guides;

    // ---- outputs ------------------------------------------------------------------------------------
    struct Out { void* host; size_t bytes; } outs[8] = {
        {io->x1, (size_t)T * sizeof(double)},        {io->x2, (size_t)T * sizeof(double)},
        {io->w1, (size_t)n_genes * L * sizeof(double)}, {io->w2, (size_t)n_genes * L * sizeof(double)},
        {io->y, (size_t)T * L * sizeof(double)},     {io->tau, (size_t)T * L * sizeof(double)},
        {io->n_iters, (size_t)n_genes * sizeof(int32_t)}, {io->bound, (size_t)n_genes * sizeof(double)}};
    if (!io->w1 && n_genes) { jacks_em_plan_destroy(plan); return fail(JACKS_E_INVALID, "w1 is required"); }
    void* dptr[8] = {nullptr};
    for (int i = 0; i < 8; ++i) {
        if (!outs[i].host || !outs[i].bytes) continue;
        if ((rc = ctx->out[i].reserve(outs[i].bytes)) != JACKS_OK) { jacks_em_plan_destroy(plan); return rc; }
        dptr[i] = ctx->out[i].p;
    }
    JacksEmDeviceIO dio;
    memset(&dio, 0, sizeof dio);
    dio.test = (const double*)ctx->res_test.p; dio.ctrl = (const double*)ctx->res_ctrl.p;
    dio.ctrl_lines = io->ctrl_lines; dio.ctrl_rowmean_fill = io->ctrl_rowmean_fill;
    dio.fixed_x1 = d_fx1; dio.fixed_x2 = d_fx2;
    dio.x1 = (double*)dptr[0]; dio.x2 = (double*)dptr[1]; dio.w1 = (double*)dptr[2]; dio.w2 = (double*)dptr[3];
    dio.y = (double*)dptr[4]; dio.tau = (double*)dptr[5]; dio.n_iters = (int32_t*)dptr[6]; dio.bound = (double*)dptr[7];
    rc = jacks_em_run_device(ctx, plan, params, &dio, st);
    if (rc == JACKS_OK) {
        for (int i = 0; i < 8 && rc == JACKS_OK; ++i) {
            if (!dptr[i]) continue;
            if ((e = cudaMemcpyAsync(outs[i].host, dptr[i], outs[i].bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) rc = cuda_fail(e, "D2H result");
        }
    }
    e = cudaStreamSynchronize(st);
    if (rc == JACKS_OK && e != cudaSuccess) rc = cuda_fail(e, "EM kernels");
    jacks_em_plan_destroy(plan);
    if (!keep_resident) ctx->res_valid = false;
    return rc;
}

}  // extern "C"
