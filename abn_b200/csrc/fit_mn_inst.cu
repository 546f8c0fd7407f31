// fit_mn_inst.cu - one translation unit per design width: -DABN_INST_K=<k> instantiates the
// multinomial-child kernel (fit_mn_kernel.cuh) for L = 2..5 levels, with and without intercept.
#include "fit_mn_kernel.cuh"

#ifndef ABN_INST_K
#error "compile with -DABN_INST_K=..."
#endif

namespace abn {

#define ABN_CAT2_(a, b) a##b
#define ABN_CAT2(a, b) ABN_CAT2_(a, b)

MnLaunchFn ABN_CAT2(abn_mn_launch_, ABN_INST_K)(int L, int icpt)
{
    switch (L) {
    case 2: return icpt ? launch_mn<2, ABN_INST_K, true> : launch_mn<2, ABN_INST_K, false>;
    case 3: return icpt ? launch_mn<3, ABN_INST_K, true> : launch_mn<3, ABN_INST_K, false>;
    case 4: return icpt ? launch_mn<4, ABN_INST_K, true> : launch_mn<4, ABN_INST_K, false>;
    case 5: return icpt ? launch_mn<5, ABN_INST_K, true> : launch_mn<5, ABN_INST_K, false>;
    }
    return nullptr;
}

} // namespace abn
