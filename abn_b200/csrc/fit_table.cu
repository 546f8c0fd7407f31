// fit_table.cu - lookup of the streaming fit kernel launchers compiled from fit_inst.cu.
#include "fit_kernel.cuh"
#include "fit_mn_kernel.cuh"

namespace abn {

#define DECL(KIND) \
    FitLaunchFn abn_fit_launch_##KIND##_1(int); FitLaunchFn abn_fit_launch_##KIND##_2(int); \
    FitLaunchFn abn_fit_launch_##KIND##_3(int); FitLaunchFn abn_fit_launch_##KIND##_4(int); \
    FitLaunchFn abn_fit_launch_##KIND##_5(int); FitLaunchFn abn_fit_launch_##KIND##_6(int); \
    FitLaunchFn abn_fit_launch_##KIND##_7(int); FitLaunchFn abn_fit_launch_##KIND##_8(int);
DECL(0)
DECL(1)
DECL(2)

typedef FitLaunchFn (*Getter)(int);
#define ROW(KIND) { nullptr, abn_fit_launch_##KIND##_1, abn_fit_launch_##KIND##_2, abn_fit_launch_##KIND##_3, \
                    abn_fit_launch_##KIND##_4, abn_fit_launch_##KIND##_5, abn_fit_launch_##KIND##_6, \
                    abn_fit_launch_##KIND##_7, abn_fit_launch_##KIND##_8 }
static const Getter getters[3][MAXK + 1] = { ROW(0), ROW(1), ROW(2) };
static_assert(MAXK == 8, "extend the launcher table");
static_assert(RK_BIN_MLE == 0 && RK_POIS == 1 && RK_BIN_BAYES == 2, "row-kind numbering changed");

static FitLaunchFn lookup(int kind, int k, int icpt)
{
    if (k < 1 || k > MAXK) return nullptr;
    return getters[kind][k](icpt);
}

FitLaunchFn get_launch_bin_mle(int k, int icpt) { return lookup(RK_BIN_MLE, k, icpt); }
FitLaunchFn get_launch_bin_bayes(int k, int icpt) { return lookup(RK_BIN_BAYES, k, icpt); }
FitLaunchFn get_launch_pois(int k, int icpt) { return lookup(RK_POIS, k, icpt); }

MnLaunchFn abn_mn_launch_1(int, int); MnLaunchFn abn_mn_launch_2(int, int); MnLaunchFn abn_mn_launch_3(int, int);
MnLaunchFn abn_mn_launch_4(int, int); MnLaunchFn abn_mn_launch_5(int, int); MnLaunchFn abn_mn_launch_6(int, int);
MnLaunchFn abn_mn_launch_7(int, int); MnLaunchFn abn_mn_launch_8(int, int);

MnLaunchFn get_launch_mn(int L, int k, int icpt)
{
    typedef MnLaunchFn (*G)(int, int);
    static const G g[MAXK + 1] = { nullptr, abn_mn_launch_1, abn_mn_launch_2, abn_mn_launch_3, abn_mn_launch_4,
                                   abn_mn_launch_5, abn_mn_launch_6, abn_mn_launch_7, abn_mn_launch_8 };
    if (k < 1 || k > MAXK || L < 2 || L > MN_MAXL) return nullptr;
    return g[k](L, icpt);
}

} // namespace abn
