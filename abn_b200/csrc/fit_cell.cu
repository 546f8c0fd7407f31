// fit_cell.cu - the contingency-table kernels for all-discrete designs (fit_cell.cuh) and their launcher.
#include <algorithm>

#include "fit_cell.cuh"

using namespace abn;

template <int KIND, int CAPBITS>
static cudaError_t launch_cell(const DataMeta &m, const Control &c, const FitDesc *descs, const int *child, const int *list,
                               int count, int *counter, ResultPtrs out, int sm_count, cudaStream_t stream)
{
    constexpr size_t smem = (size_t)CELL_WARPS * ((1 << CAPBITS) + 32) * sizeof(unsigned);
    static bool configured[64] = {false};
    if (first_launch_on_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(fit_cell_kernel<KIND, CAPBITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int per_sm = CAPBITS == CELL_PLANES_CAPBITS ? CELL_PLANES_PER_SM : (CAPBITS == CELL_SMALL_CAPBITS ? 4 : CELL_LARGE_PER_SM);
    const int grid = std::min((count + CELL_WARPS - 1) / CELL_WARPS, sm_count * per_sm);
    // overflow-list slots (DataMeta::cell_ovf): [binomial small | binomial large | Poisson small | Poisson large], each
    // sized for its largest grid, so that the four kernels can run side by side
    const int small_slots = sm_count * 4 * CELL_WARPS, large_slots = sm_count * CELL_LARGE_PER_SM * CELL_WARPS;
    const int fam = (KIND == RK_POIS) ? 1 : 0;
    // (the bit-plane kernel has no capped columns and never touches the lists)
    const int slot0 = fam * (small_slots + large_slots) + (CAPBITS == CELL_LARGE_CAPBITS ? small_slots : 0);
    fit_cell_kernel<KIND, CAPBITS><<<grid, CELL_WARPS * 32, smem, stream>>>(m, c, descs, child, list, count, counter, out, slot0);
    return cudaGetLastError();
}

template <int CAPBITS>
static cudaError_t launch_cell_kind(int kind, const DataMeta &m, const Control &c, const FitDesc *descs, const int *child,
                                    const int *list, int count, int *counter, ResultPtrs out, int sm_count, cudaStream_t stream)
{
    if (kind == RK_POIS) return launch_cell<RK_POIS, CAPBITS>(m, c, descs, child, list, count, counter, out, sm_count, stream);
    if (kind == RK_BIN_BAYES) return launch_cell<RK_BIN_BAYES, CAPBITS>(m, c, descs, child, list, count, counter, out, sm_count, stream);
    return launch_cell<RK_BIN_MLE, CAPBITS>(m, c, descs, child, list, count, counter, out, sm_count, stream);
}

int abn_cell_ovf_slots(int sm_count) { return 2 * (sm_count * 4 * CELL_WARPS + sm_count * CELL_LARGE_PER_SM * CELL_WARPS); }

cudaError_t abn_cell_launch(int kind, int cls, const abnc::DataMeta &m, const abnc::Control &c, const abn::FitDesc *descs,
                            const int *child, const int *list, int count, int *counter, abnc::ResultPtrs out,
                            int sm_count, cudaStream_t stream)
{
    // cls: 0 small table, 1 large table, 2 all-binary design on bit planes (plan.cuh CELL_BUCKET0)
    if (cls == 2) return launch_cell_kind<CELL_PLANES_CAPBITS>(kind, m, c, descs, child, list, count, counter, out, sm_count, stream);
    return cls == 1 ? launch_cell_kind<CELL_LARGE_CAPBITS>(kind, m, c, descs, child, list, count, counter, out, sm_count, stream)
                    : launch_cell_kind<CELL_SMALL_CAPBITS>(kind, m, c, descs, child, list, count, counter, out, sm_count, stream);
}
