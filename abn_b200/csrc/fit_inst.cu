// fit_inst.cu - one translation unit per (row op, design width): compiled many times by the
// Makefile with -DABN_INST_KIND=<RowKind> -DABN_INST_K=<k> so the instantiations of the streaming
// fit kernel (fit_kernel.cuh) build in parallel.
#include "fit_kernel.cuh"

#ifndef ABN_INST_KIND
#error "compile with -DABN_INST_KIND=... -DABN_INST_K=..."
#endif

namespace abn {

#define ABN_CAT_(a, b, c) a##b##_##c
#define ABN_CAT(a, b, c) ABN_CAT_(a, b, c)

// abn_fit_launch_<kind>_<k>(icpt)
FitLaunchFn ABN_CAT(abn_fit_launch_, ABN_INST_KIND, ABN_INST_K)(int icpt)
{
#if ABN_INST_KIND == 2 /* RK_BIN_BAYES: the Bayes design always has an intercept */
    return icpt ? launch_fit<ABN_INST_KIND, ABN_INST_K, true> : nullptr;
#else
    return icpt ? launch_fit<ABN_INST_KIND, ABN_INST_K, true> : launch_fit<ABN_INST_KIND, ABN_INST_K, false>;
#endif
}

} // namespace abn
