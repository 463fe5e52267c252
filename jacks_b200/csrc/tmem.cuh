// Tensor memory (TMEM, 256 KB per SM on sm_100a) as lane-private FP64 storage.
//
// TMEM is 128 lanes x 512 columns of 32 bits; warp w of a CTA (w mod 4) owns lanes 32w .. 32w+31 and thread t of the
// warp its lane 32w + t.  With the 32x32b shape one tcgen05.ld / tcgen05.st moves N consecutive columns of every
// thread's own lane to / from N registers -- a dynamically indexable extension of the register file that does
// not go through the L1 / shared-memory data pipe.  A double takes two columns (lo, hi).
// Measured on B200 (tools/probes/tmem_probe.cu): exact round trip; 12 warps per SM doing ld -> FMA -> st over
// 128 columns sustain >= 175 B/clk/SM in each direction (shared memory: 128 B/clk/SM for both).
//
// All of these are .sync.aligned: every thread of the warp must execute them, converged.
#pragma once
#include <cstdint>

namespace jacks {

// Allocate `ncols` (power of two, 32..512) columns for the CTA; executed by ONE full warp.  The base address lands
// in *smem_out (shared memory).  Blocks until the columns are free on this SM.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out, uint32_t ncols) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_out);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int N> struct TmemCols;      // N = 2, 4, 8, 16, 32 columns per instruction

template <> struct TmemCols<2> {
    static __device__ __forceinline__ void ld(uint32_t addr, uint32_t* r) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];"
                     : "=r"(r[0]), "=r"(r[1])
                     : "r"(addr));
    }
    static __device__ __forceinline__ void st(uint32_t addr, const uint32_t* r) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};"
                     :
                     : "r"(addr), "r"(r[0]), "r"(r[1]));
    }
};

template <> struct TmemCols<4> {
    static __device__ __forceinline__ void ld(uint32_t addr, uint32_t* r) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(addr));
    }
    static __device__ __forceinline__ void st(uint32_t addr, const uint32_t* r) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                     :
                     : "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]));
    }
};

template <> struct TmemCols<8> {
    static __device__ __forceinline__ void ld(uint32_t addr, uint32_t* r) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(addr));
    }
    static __device__ __forceinline__ void st(uint32_t addr, const uint32_t* r) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     :
                     : "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
    }
};

template <> struct TmemCols<16> {
    static __device__ __forceinline__ void ld(uint32_t addr, uint32_t* r) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(addr));
    }
    static __device__ __forceinline__ void st(uint32_t addr, const uint32_t* r) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                     :
                     : "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
    }
};

template <> struct TmemCols<32> {
    static __device__ __forceinline__ void ld(uint32_t addr, uint32_t* r) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(addr));
    }
    static __device__ __forceinline__ void st(uint32_t addr, const uint32_t* r) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                     :
                     : "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]));
    }
};


__host__ __device__ constexpr int tmem_chunk(int cols) { return cols >= 32 ? 32 : cols >= 16 ? 16 : cols >= 8 ? 8 : cols >= 4 ? 4 : 2; }

// Load ND doubles (2*ND columns starting at `addr`) into v; the caller issues tmem_wait_ld() before using them.
template <int ND>
__device__ __forceinline__ void tmem_ld_issue(uint32_t addr, uint32_t (&r)[2 * ND]) {
    constexpr int C = tmem_chunk(2 * ND);
    TmemCols<C>::ld(addr, r);
    if constexpr (2 * ND > C) {
        uint32_t(&rest)[2 * ND - C] = reinterpret_cast<uint32_t(&)[2 * ND - C]>(r[C]);
        tmem_ld_issue<(2 * ND - C) / 2>(addr + C, rest);
    }
}
template <int ND>
__device__ __forceinline__ void tmem_st_issue(uint32_t addr, const uint32_t (&r)[2 * ND]) {
    constexpr int C = tmem_chunk(2 * ND);
    TmemCols<C>::st(addr, r);
    if constexpr (2 * ND > C) {
        const uint32_t(&rest)[2 * ND - C] = reinterpret_cast<const uint32_t(&)[2 * ND - C]>(r[C]);
        tmem_st_issue<(2 * ND - C) / 2>(addr + C, rest);
    }
}

// Convenience forms on doubles.
template <int ND>
__device__ __forceinline__ void tmem_load(uint32_t addr, double* v) {
    uint32_t r[2 * ND];
    tmem_ld_issue<ND>(addr, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < ND; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
template <int ND>
__device__ __forceinline__ void tmem_store(uint32_t addr, const double* v) {
    uint32_t r[2 * ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) { r[2 * i] = (uint32_t)__double2loint(v[i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[i]); }
    tmem_st_issue<ND>(addr, r);
}

// Columns of one [lines per lane, rounded up to 4][G] array of doubles in a thread's lane.
__host__ __device__ constexpr int tmem_array_cols(int G, int L) { return 2 * G * (((L + 31) / 32 + 3) / 4 * 4); }
// Columns a CTA allocates for `arrays` such arrays, rounded to what tcgen05.alloc accepts; 0 if they cannot fit.
__host__ __device__ constexpr int tmem_cols_for(int G, int L, int arrays = 1) {
    const int need = arrays * tmem_array_cols(G, L);
    int c = 32;
    while (c < need) c <<= 1;
    return c <= 512 ? c : 0;
}

}  // namespace jacks
