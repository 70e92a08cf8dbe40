// dp_jump.cu -- instantiates the DP kernels of one aligner kind (see dp_kernels.cuh)
#include "dp_kernels.cuh"

template int mb200_dp_run_kind<1>(mb200_ctx*, const mb200_scores*, int, std::vector<DpTask>&, void*, uint32_t*, uint64_t,
                                  unsigned long long*, int*, uint32_t*, uint64_t*, uint64_t*);
