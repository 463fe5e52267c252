// Switch from a runtime guide count to the per-G translation units (em_fast_inst.cu).
#include "em_common.cuh"

namespace jacks {

#define JACKS_DECL(g)                                                                                         \
    cudaError_t launch_em_fast_g##g(int KL, int W, const EmLaunch& a, int n_sm, int grid_limit,               \
                                    cudaStream_t st, int* grid_out);                                          \
    int em_fast_occupancy_g##g(int KL, int W);                                                                \
    bool em_fast_available_g##g(int KL, int W);
JACKS_DECL(1) JACKS_DECL(2) JACKS_DECL(3) JACKS_DECL(4) JACKS_DECL(5) JACKS_DECL(6) JACKS_DECL(7) JACKS_DECL(8)
#undef JACKS_DECL

#define JACKS_SWITCH(call)            \
    switch (G) {                      \
        case 1: return call(1);       \
        case 2: return call(2);       \
        case 3: return call(3);       \
        case 4: return call(4);       \
        case 5: return call(5);       \
        case 6: return call(6);       \
        case 7: return call(7);       \
        case 8: return call(8);       \
        default: break;               \
    }

cudaError_t launch_em_fast(int G, int KL, int W, const EmLaunch& a, int n_sm, int grid_limit, cudaStream_t st,
                           int* grid_out) {
#define CALL(g) launch_em_fast_g##g(KL, W, a, n_sm, grid_limit, st, grid_out)
    JACKS_SWITCH(CALL)
#undef CALL
    return cudaErrorInvalidValue;
}

bool em_fast_available(int G, int KL, int W) {
#define CALL(g) em_fast_available_g##g(KL, W)
    JACKS_SWITCH(CALL)
#undef CALL
    return false;
}

int em_fast_occupancy(int G, int KL, int W) {
#define CALL(g) em_fast_occupancy_g##g(KL, W)
    JACKS_SWITCH(CALL)
#undef CALL
    return 0;
}

}  // namespace jacks
