// abn_b200.cu - C ABI (include/abn_b200.h), device data set, global Gram, design resolution,
// Gram-only Gaussian kernels and the dispatch of the streaming fit kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/abn_b200.h"
#include "fit_kernel.cuh"
#include "fit_mn_kernel.cuh"
#include "fit_cell.cuh"

using namespace abn;

// wide.cu: generic kernel for designs of 9..32 columns
cudaError_t abn_wide_launch(const abnc::DataMeta &m, const abnc::Control &c, const int *child, const int *parents,
                            int width, const int *list, int count, int *counter, abnc::ResultPtrs out, int sm_count,
                            cudaStream_t stream);
int abn_wide_max_k();
int abn_cell_ovf_slots(int sm_count); // fit_cell.cu: warps of all contingency-table kernels that may run together
// fit_cell.cu: all-discrete designs on contingency tables (kind: RK_BIN_MLE, RK_POIS, RK_BIN_BAYES)
cudaError_t abn_cell_launch(int kind, int cls, const abnc::DataMeta &m, const abnc::Control &c, const abn::FitDesc *descs,
                            const int *child, const int *list, int count, int *counter, abnc::ResultPtrs out,
                            int sm_count, cudaStream_t stream);

// per-bucket slot of abn_data::dCounters: [0] work-queue head, [2..15] 64-bit pass tallies by PassMode (fit_kernel.cuh)
constexpr int BUCKET_SLOT_INTS = 16;

// FP64 instructions one observation row costs in a pass of the streaming kernels, by row kind, design width and pass
// mode - counted from the SASS of the lean passes (profiles/sass_r2_*; e.g. Poisson K = 4: 3 eta + 8 exp + 13 H = 24,
// binomial K = 4 HG: 3 + 8 + 1 + 2 rcp + 1 w + 13 H + 4 S = 32). The rare careful / no-intercept passes are
// counted with the same formula.
static int fp64_per_row(int fam /*0 binomial, 1 Poisson*/, int k, int mode)
{
    const bool need_h = (mode == PASS_FULL || mode == PASS_HG || mode == PASS_HESS || mode == PASS_PINIT);
    const bool need_g = (mode == PASS_FULL || mode == PASS_HG || mode == PASS_GRAD || mode == PASS_PINIT);
    const bool need_l = (mode == PASS_FULL || mode == PASS_LL);
    const int kd = k > 0 ? k - 1 : 0;
    const int hsum = 1 + 2 * kd + k * kd / 2; // w; w x_q; (w x_p); (w x_p) x_q
    int c = kd + (ABN_EXP_REPLICATED ? 10 : (ABN_LEAN_EXP1 ? 8 : 9)); // eta, exp
    if (fam == 1) return c + (need_h ? hsum : (need_g ? 1 + kd : 1));
    c += 1;                                   // d = 1 + e
    if (need_l) c += 13;                      // log(1 + e) and its sum
    if (need_h || need_g) c += ABN_LEAN_RCP1 ? 2 : 3; // reciprocal
    if (need_h) c += 1 + hsum;
    if (need_g) c += 1 + kd;
    return c;
}

static_assert(ABN_FAST_K == MAXK, "include/abn_b200.h and solver.cuh disagree on the fast-path design width");
static_assert(MN_MAXL == MN_MAXL_, "plan.cuh and solver_mn.cuh disagree on the multinomial level limit");

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
// Device memory comes from the device's default stream-ordered pool with the release threshold
// lifted, so the buffers of a finished call stay cached in the process and the next one-shot
// call (abn_b200_score_cache: create + score + destroy) pays neither cudaMalloc nor the page
// unmapping of cudaFree - measured at 5..30 ms and 3..380 ms per call for a 0.8 GB data set.
static cudaError_t dev_alloc(void **p, size_t bytes, cudaStream_t s)
{
    cudaError_t e = cudaMallocAsync(p, bytes, s);
    // ABN_B200_FILL=<byte>: fill every new allocation (debugging reads of memory never written)
    static int fill = -2;
    if (fill == -2) { const char *v = std::getenv("ABN_B200_FILL"); fill = v ? std::atoi(v) : -1; }
    if (e == cudaSuccess && fill >= 0) e = cudaMemsetAsync(*p, fill, bytes, s);
    return e;
}
template <typename T>
static cudaError_t dev_alloc(T **p, size_t bytes, cudaStream_t s) { return dev_alloc(reinterpret_cast<void **>(p), bytes, s); }
static void dev_free(void *p, cudaStream_t s) { if (p) cudaFreeAsync(p, s); }
static void keep_pool(int device)
{
    static bool done[64] = {false};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
    done[device] = true;
}

// ABN_B200_TRACE=1: wall-clock phases of the host-side calls on stderr (profiles/, bench.py e2e)
struct PhaseTrace {
    bool on;
    std::chrono::steady_clock::time_point t;
    explicit PhaseTrace() : on(std::getenv("ABN_B200_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
    void mark(const char *what)
    {
        if (!on) return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[abn_b200] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

static int fail(char *err, size_t errlen, int code, const char *fmt, ...)
{
    if (err && errlen) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, errlen, fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            int code_ = (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ||            \
                         e_ == cudaErrorNoKernelImageForDevice) ? ABN_ERR_NO_GPU : ABN_ERR_CUDA;    \
            if (e_ == cudaErrorMemoryAllocation) code_ = ABN_ERR_ALLOC;                             \
            cudaGetLastError();                                                                     \
            return fail(err, errlen, code_, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_),       \
                        __FILE__, __LINE__);                                                        \
        }                                                                                           \
    } while (0)

// ABN_B200_DEBUG_SYNC=1: synchronise and check after every launch, naming the step that failed
static bool debug_sync()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("ABN_B200_DEBUG_SYNC"); v = (e && *e && *e != '0') ? 1 : 0; }
    return v == 1;
}
#define KCHECK(what, stream)                                                                        \
    do {                                                                                            \
        cudaError_t e_ = cudaGetLastError();                                                        \
        if (e_ == cudaSuccess && debug_sync()) { e_ = cudaStreamSynchronize(stream); if (e_ == cudaSuccess) e_ = cudaGetLastError(); } \
        if (e_ != cudaSuccess)                                                                      \
            return fail(err, errlen, ABN_ERR_CUDA, "%s: %s (%s:%d)", what, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// K1: global Gram  G = Z^T Z,  Z = [1 | X | log(y + 0.1) of the Poisson nodes], FP64, deterministic
// ---------------------------------------------------------------------------------------------
constexpr int GT = 6;        // register tile edge
constexpr int GRAM_THREADS = 256;

// value of Gram column a at row i; `shift` (optional, [1 + n_cols], shift[0] = 0) is subtracted: the centred Gram
__device__ __forceinline__ double z_value(const DataMeta &m, int a, long long i, const double *__restrict__ shift)
{
    if (a == 0) return 1.0;
    if (a <= m.n_cols) return __ldg(m.X + (long long)(a - 1) * m.ld + i) - (shift ? shift[a] : 0.0);
    const int node = m.pois_node[a - 1 - m.n_cols]; // virtual column: log(y + 0.1) of a Poisson node (Bayes start values)
    return log(__ldg(m.X + (long long)m.node_col0[node] * m.ld + i) + 0.1);
}

// grid.x = row chunks, grid.y = groups of GRAM_THREADS tile pairs. Each thread owns one GT x GT
// tile (I <= J) of the upper block triangle; the CTA stages RC rows of all D columns in shared
// memory (row-major, padded) and every thread sweeps them. Partial tiles go to
// part[chunk][pair][GT*GT]; gram_reduce sums the chunks in a fixed order.
__global__ void __launch_bounds__(GRAM_THREADS)
gram_partial(const DataMeta m, int Dz, int T, int NP, int RC, int Dp, long long rows_per_chunk,
             double *__restrict__ part, const double *__restrict__ shift)
{
    extern __shared__ double zs[]; // RC x Dp
    const int pair = blockIdx.y * GRAM_THREADS + threadIdx.x;
    int I = 0, J = 0;
    if (pair < NP) { // unrank (I, J), I <= J, row-major over the upper triangle
        int p = pair, rowlen = T;
        while (p >= rowlen) { p -= rowlen; rowlen--; I++; }
        J = I + p;
    }
    double acc[GT][GT];
#pragma unroll
    for (int a = 0; a < GT; a++)
#pragma unroll
        for (int b = 0; b < GT; b++) acc[a][b] = 0.0;

    const long long r0 = (long long)blockIdx.x * rows_per_chunk;
    const long long r1 = min(m.n, r0 + rows_per_chunk);
    for (long long base = r0; base < r1; base += RC) {
        const int nr = (int)min((long long)RC, r1 - base);
        __syncthreads();
        // coalesced along rows within a column; transposed store. Each GT-column tile is padded to GT + 1
        // slots: consecutive threads read consecutive J tiles of the same row, and a stride of GT = 6
        // doubles (12 banks) is a 4-way bank conflict where 7 doubles (14 banks) is conflict-free for
        // 64-bit loads. Columns past Dz in the last tile are staged as zeros.
        for (int e = threadIdx.x; e < RC * T * GT; e += GRAM_THREADS) {
            const int col = e / RC, r = e - col * RC;
            zs[r * Dp + (col / GT) * (GT + 1) + col % GT] = (col < Dz && r < nr) ? z_value(m, col, base + r, shift) : 0.0;
        }
        __syncthreads();
        if (pair < NP) {
            for (int r = 0; r < nr; r++) {
                double a[GT], b[GT];
#pragma unroll
                for (int t = 0; t < GT; t++) { a[t] = zs[r * Dp + I * (GT + 1) + t]; b[t] = zs[r * Dp + J * (GT + 1) + t]; }
#pragma unroll
                for (int s = 0; s < GT; s++)
#pragma unroll
                    for (int t = 0; t < GT; t++) acc[s][t] = fma(a[s], b[t], acc[s][t]);
            }
        }
    }
    if (pair < NP) {
        double *o = part + ((long long)blockIdx.x * NP + pair) * (GT * GT);
#pragma unroll
        for (int s = 0; s < GT; s++)
#pragma unroll
            for (int t = 0; t < GT; t++) o[s * GT + t] = acc[s][t];
    }
}

__global__ void gram_reduce(int D, int T, int NP, int chunks, const double *__restrict__ part, double *__restrict__ G)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x; // element of a tile pair
    if (e >= NP * GT * GT) return;
    const int pair = e / (GT * GT), st = e - pair * (GT * GT);
    int I = 0, p = pair, rowlen = T;
    while (p >= rowlen) { p -= rowlen; rowlen--; I++; }
    const int J = I + p;
    const int i = I * GT + st / GT, j = J * GT + st % GT;
    if (i >= D || j >= D) return;
    // diagonal tiles hold (i, j) and (j, i) both: only the upper element is used and mirrored, so that no
    // two threads write the same entry (with row weights the two sums would even round differently)
    if (I == J && i > j) return;
    double s = 0.0;
    for (int c = 0; c < chunks; c++) s += part[((long long)c * NP + pair) * (GT * GT) + st];
    G[(long long)i * D + j] = s;
    G[(long long)j * D + i] = s;
}

// shift[0] = 0, shift[1 + j] = mean of stored column j = G[0][1 + j] / n
__global__ void col_mean_kernel(int n_cols, int D, double n, const double *__restrict__ G, double *__restrict__ shift)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a <= n_cols) shift[a] = (a == 0) ? 0.0 : G[a] / n;
}

// sum lgamma(y + 1) per Poisson node (src/irls_poisson_fast.cpp:32-34, gsl_sf_lnfact in
// src/node_poisson.c:561): one CTA per node, fixed-order tree
__global__ void __launch_bounds__(256) lfact_kernel(const DataMeta m, double *__restrict__ out)
{
    __shared__ double red[256];
    const int node = blockIdx.x;
    double s = 0.0;
    if (m.node_family[node] == POISSON) {
        const double *y = m.X + (long long)m.node_col0[node] * m.ld;
        for (long long i = threadIdx.x; i < m.n; i += 256) s += lgamma(__ldg(y + i) + 1.0);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[node] = red[0];
}

// ---------------------------------------------------------------------------------------------
// byte codes of the discrete columns (fit_cell.cuh): one pass over X at data_create. Every thread takes 16
// consecutive rows of one column (128 bytes in, one 16-byte store out); a column counts as discrete when all
// its values are integers in 0..255, and its cardinality is the largest code + 1.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
discretize_kernel(const double *__restrict__ X, long long ld, long long n, unsigned char *__restrict__ Xq, long long ldq,
                  int *__restrict__ col_bad, int *__restrict__ col_max)
{
    const int col = blockIdx.y;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; // group of 16 rows
    const long long r0 = g * 16;
    int bad = 0, mx = 0;
    if (r0 < n) {
        const double *x = X + (long long)col * ld + r0;
        unsigned w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const double v = (r0 + i < n) ? __ldg(x + i) : 0.0;
            const int c = (v >= 0.0 && v <= 255.0) ? (int)v : 0;
            bad |= ((double)c != v);
            mx = max(mx, c);
            w[i >> 2] |= (unsigned)c << (8 * (i & 3));
        }
        *reinterpret_cast<uint4 *>(Xq + (long long)col * ldq + r0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    bad = __any_sync(0xffffffffu, bad);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicOr(&col_bad[col], 1);
        atomicMax(&col_max[col], mx);
    }
}

// col_tail[c * 8 + b] = rows of column c whose byte code is >= 2^b: what capping the column at b bits would send
// to the overflow lists (fit_cell.cuh cell_bits_of)
__global__ void __launch_bounds__(256)
col_tail_kernel(const unsigned char *__restrict__ Xq, long long ldq, long long n, unsigned *__restrict__ col_tail)
{
    __shared__ unsigned acc[8];
    if (threadIdx.x < 8) acc[threadIdx.x] = 0u;
    __syncthreads();
    const int col = blockIdx.y;
    const long long r0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    unsigned cnt[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (r0 < n) {
        const uint4 v = *reinterpret_cast<const uint4 *>(Xq + (long long)col * ldq + r0);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (r0 + i >= n) break;
            const unsigned code = (w[i >> 2] >> (8 * (i & 3))) & 0xffu;
#pragma unroll
            for (int b = 0; b < 8; b++) cnt[b] += (code >= (1u << b)) ? 1u : 0u;
        }
    }
#pragma unroll
    for (int b = 0; b < 8; b++) {
        unsigned s = cnt[b];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&acc[b], s);
    }
    __syncthreads();
    if (threadIdx.x < 8 && acc[threadIdx.x]) atomicAdd(&col_tail[col * 8 + threadIdx.x], acc[threadIdx.x]);
}

// bit planes of the binary columns (fit_cell.cuh cell_count_planes): one word per warp, all ldb words of a column
// are written (zero beyond n and for columns that are not binary)
__global__ void bitplane_kernel(const unsigned char *__restrict__ Xq, long long ldq, long long n, const int *__restrict__ col_card,
                                unsigned *__restrict__ Xb, long long ldb)
{
    const int col = blockIdx.y;
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int card = col_card[col];
    const bool bit = (card == 1 || card == 2) && r < n && (Xq[(long long)col * ldq + r] & 1u);
    const unsigned w = __ballot_sync(0xffffffffu, bit);
    if ((threadIdx.x & 31) == 0 && (r >> 5) < ldb) Xb[(long long)col * ldb + (r >> 5)] = w;
}

__global__ void col_card_kernel(int n_cols, const int *__restrict__ col_bad, const int *__restrict__ col_max, int *__restrict__ col_card)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cols) col_card[c] = col_bad[c] ? 0 : col_max[c] + 1;
}

// ---------------------------------------------------------------------------------------------
// design resolution: one thread per candidate
// ---------------------------------------------------------------------------------------------
__global__ void resolve_kernel(const DataMeta m, int method, int flags, long long rows, int width, const int *__restrict__ child,
                               const int *__restrict__ parents, FitDesc *__restrict__ descs,
                               unsigned char *__restrict__ keys, int *__restrict__ idx)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    FitDesc d;
    resolve_design(m, method, child[r], parents + r * width, width, d);
    if (d.k == 0 || (d.family == MULTINOMIAL && (d.nlev < 2 || d.nlev > MN_MAXL || method != MLE))) d.err = d.err ? d.err : 2;
    descs[r] = d;
    int key = bucket_of(d);
    const int cell = cell_eligible(m, d, method, flags);
    if (cell) key = CELL_BUCKET0 + (d.family == POISSON ? 3 : 0) + (cell - 1);
    keys[r] = (unsigned char)key;
    idx[r] = (int)r;
}

__global__ void bucket_bounds(const unsigned char *__restrict__ keys, long long rows, int *__restrict__ start)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int k = keys[r];
    if (r == 0 || keys[r - 1] != k) start[k] = (int)r;
}

// rows the streaming kernels cannot take (design too wide / multinomial child / empty design)
__global__ void unsupported_kernel(const int *__restrict__ list, int count, ResultPtrs out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int row = list[t];
    if (out.mlik) out.mlik[row] = NAN;
    if (out.aic) out.aic[row] = NAN;
    if (out.bic) out.bic[row] = NAN;
    if (out.mdl) out.mdl[row] = NAN;
    if (out.status) out.status[row] = FIT_NON_FINITE;
    if (out.iters) out.iters[row] = 0;
}

// ---------------------------------------------------------------------------------------------
// Gaussian children: no data pass, everything from Gram blocks. One thread per fit.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gauss_kernel(const DataMeta m, const Control c, const FitDesc *__restrict__ descs, const int *__restrict__ child,
             const int *__restrict__ list, int count, ResultPtrs out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int row = list[t];
    const FitDesc d = descs[row];
    if (c.method == MLE) {
        double ll, aic, bic, x[MAXM], se[MAXM], sse;
        int status;
        gauss_mle_fit(m, d, &ll, &aic, &bic, &status, x, out.se ? se : nullptr, &sse);
        write_coef(out, row, d, x, d.k);
        if (out.se) { ResultPtrs o = out; o.coef = out.se; write_coef(o, row, d, se, d.k); }
        if (out.sse) out.sse[row] = sse;
        if (out.mlik) out.mlik[row] = ll;
        if (out.aic) out.aic[row] = aic;
        if (out.bic) out.bic[row] = bic;
        if (out.mdl) out.mdl[row] = mdl_of(bic, d, c.nvars);
        if (out.status) out.status[row] = status;
        if (out.iters) out.iters[row] = 1;
        return;
    }
    FitState s;
    PassAcc a;
    GaussAux aux;
    state_setup(m, d, child[row], c, s);
    if (gauss_centered_block(m, d, aux)) s.gaux = &aux;
    bayes_inits(m, d, child[row], s);
    gauss_gram_acc(m, d, a);
    int mode = fit_begin(s, c);
    while (mode != PASS_DONE) mode = fit_consume(s, c, a);
    if (out.mlik) out.mlik[row] = s.loglik;
    if (out.status) out.status[row] = s.status;
    write_coef(out, row, d, s.x, s.M);
    if (out.iters) out.iters[row] = 1;
}

// ---------------------------------------------------------------------------------------------
// device data set
// ---------------------------------------------------------------------------------------------
struct abn_data {
    int device = 0;
    DataMeta m{};
    double *dX = nullptr, *dG = nullptr, *dLfact = nullptr;
    int *dFamily = nullptr, *dCol0 = nullptr, *dGwIndex = nullptr, *dPoisNode = nullptr;
    unsigned char *dXq = nullptr; // byte codes of the discrete columns (fit_cell.cuh)
    unsigned *dXb = nullptr;      // bit planes of the binary ones
    long long ldb = 0;
    int *dColCard = nullptr;
    unsigned *dColTail = nullptr, *dCellOvf = nullptr; // capped codes: tail counts, overflow lists
    double *dGc = nullptr, *dColMean = nullptr; // centred Gram and the shifts it was built with ([0] = 0)
    std::vector<int> family, col0, gw_index;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    static constexpr int NSIDE = 3;
    cudaStream_t side[NSIDE] = {nullptr, nullptr, nullptr}; // small buckets run beside the large ones
    abn_timing timing{};
    // scratch that grows with the largest candidate list seen
    long long cap_rows = 0;
    int cap_width = 0;
    int *dChild = nullptr, *dParents = nullptr, *dIdx = nullptr, *dIdxSorted = nullptr, *dBucketStart = nullptr,
        *dCounters = nullptr, *dStatus = nullptr, *dIters = nullptr;
    unsigned char *dKeys = nullptr, *dKeysSorted = nullptr;
    FitDesc *dDescs = nullptr;
    double *dRes = nullptr; // 4 x rows
    void *dSortTmp = nullptr;
    size_t sortTmpBytes = 0;
};

static void free_scratch(abn_data *d)
{
    dev_free(d->dChild, d->stream); dev_free(d->dParents, d->stream); dev_free(d->dIdx, d->stream); dev_free(d->dIdxSorted, d->stream);
    dev_free(d->dStatus, d->stream); dev_free(d->dIters, d->stream); dev_free(d->dKeys, d->stream); dev_free(d->dKeysSorted, d->stream);
    dev_free(d->dDescs, d->stream); dev_free(d->dRes, d->stream); dev_free(d->dSortTmp, d->stream);
    d->dChild = d->dParents = d->dIdx = d->dIdxSorted = d->dStatus = d->dIters = nullptr;
    d->dKeys = d->dKeysSorted = nullptr; d->dDescs = nullptr; d->dRes = nullptr; d->dSortTmp = nullptr;
    d->cap_rows = 0; d->cap_width = 0; d->sortTmpBytes = 0;
}

extern "C" void abn_b200_data_destroy(abn_data *d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    PhaseTrace tr;
    free_scratch(d);
    dev_free(d->dX, d->stream); dev_free(d->dG, d->stream); dev_free(d->dLfact, d->stream); dev_free(d->dFamily, d->stream); dev_free(d->dCol0, d->stream);
    dev_free(d->dGwIndex, d->stream); dev_free(d->dPoisNode, d->stream);
    dev_free(d->dXq, d->stream); dev_free(d->dXb, d->stream); dev_free(d->dColCard, d->stream);
    dev_free(d->dColTail, d->stream); dev_free(d->dCellOvf, d->stream);
    dev_free(d->dGc, d->stream); dev_free(d->dColMean, d->stream);
    dev_free(d->dBucketStart, d->stream); dev_free(d->dCounters, d->stream);
    cudaStreamSynchronize(d->stream); // the frees above are stream-ordered
    if (d->stream) cudaStreamDestroy(d->stream);
    for (auto &ss : d->side) if (ss) cudaStreamDestroy(ss);
    delete d;
    tr.mark("destroy");
}

// G = Z'Z over the Dz columns of Z, result Dz x Dz at dst
static int launch_gram(abn_data *d, int Dz, double *dst, const double *shift, char *err, size_t errlen)
{
    const DataMeta &m = d->m;
    const int T = (Dz + GT - 1) / GT;
    const int NP = T * (T + 1) / 2;
    const int Dp = (T * (GT + 1)) | 1; // padded smem row: GT + 1 slots per tile, odd stride (conflict-free transposed stores)
    int RC = 64;
    while (RC > 8 && (size_t)RC * Dp * sizeof(double) > 160 * 1024) RC >>= 1;
    if ((size_t)RC * Dp * sizeof(double) > 200 * 1024)
        return fail(err, errlen, ABN_ERR_LIMIT, "data set has too many columns for the Gram kernel (D=%d)", Dz);
    const int gy = (NP + GRAM_THREADS - 1) / GRAM_THREADS;
    int chunks = std::max(1, (2 * d->sm_count) / gy);
    long long rpc = (m.n + chunks - 1) / chunks;
    rpc = ((rpc + RC - 1) / RC) * RC;
    if (rpc < RC) rpc = RC;
    chunks = (int)((m.n + rpc - 1) / rpc);
    if (chunks < 1) chunks = 1;
    double *part = nullptr;
    CUDA_TRY(dev_alloc(&part, (size_t)chunks * NP * GT * GT * sizeof(double), d->stream));
    const size_t smem = (size_t)RC * Dp * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(gram_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024));
    if (e == cudaSuccess) {
        gram_partial<<<dim3(chunks, gy), GRAM_THREADS, smem, d->stream>>>(m, Dz, T, NP, RC, Dp, rpc, part, shift);
        gram_reduce<<<(NP * GT * GT + 255) / 256, 256, 0, d->stream>>>(Dz, T, NP, chunks, part, dst);
        e = cudaStreamSynchronize(d->stream);
    }
    dev_free(part, d->stream);
    CUDA_TRY(e);
    CUDA_TRY(cudaGetLastError());
    d->timing.launches += 2;
    return ABN_OK;
}

static int compute_gram(abn_data *d, char *err, size_t errlen)
{
    const DataMeta &m = d->m;
    int rc = launch_gram(d, m.D, d->dG, nullptr, err, errlen);
    if (rc != ABN_OK) return rc;
    // centred second moments for the Gaussian children: column means from the raw Gram, then the Gram of
    // [1 | X - mean] (types.cuh DataMeta::Gc)
    if (d->dGc) {
        col_mean_kernel<<<(m.n_cols + 1 + 127) / 128, 128, 0, d->stream>>>(m.n_cols, m.D, (double)m.n, d->dG, d->dColMean);
        rc = launch_gram(d, 1 + m.n_cols, d->dGc, d->dColMean, err, errlen);
        if (rc != ABN_OK) return rc;
        d->timing.launches += 1;
        d->m.Gc = d->dGc;
        d->m.col_mean = d->dColMean + 1;
    }
    lfact_kernel<<<m.n_nodes, 256, 0, d->stream>>>(m, d->dLfact);
    d->timing.launches += 1;
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    CUDA_TRY(cudaGetLastError());
    return ABN_OK;
}

extern "C" int abn_b200_data_create(const abn_dataset *ds, int32_t device, abn_data **out, char *err, size_t errlen)
{
    if (!ds || !out) return fail(err, errlen, ABN_ERR_ARG, "null argument");
    *out = nullptr;
    if (ds->n <= 0 || ds->n_nodes <= 0 || ds->n_cols < ds->n_nodes || !ds->X || !ds->node_family || !ds->node_col0)
        return fail(err, errlen, ABN_ERR_ARG, "empty or inconsistent data set (n=%lld, nodes=%d, cols=%d)",
                    (long long)ds->n, ds->n_nodes, ds->n_cols);
    if (ds->node_col0[0] != 0 || ds->node_col0[ds->n_nodes] != ds->n_cols)
        return fail(err, errlen, ABN_ERR_ARG, "node_col0 must start at 0 and end at n_cols");
    for (int i = 0; i < ds->n_nodes; i++) {
        const int f = ds->node_family[i], w = ds->node_col0[i + 1] - ds->node_col0[i];
        if (f < ABN_BINOMIAL || f > ABN_MULTINOMIAL) return fail(err, errlen, ABN_ERR_ARG, "node %d: unknown family %d", i, f);
        if (w < 1 || (f != ABN_MULTINOMIAL && w != 1))
            return fail(err, errlen, ABN_ERR_ARG, "node %d: family %d cannot own %d columns", i, f, w);
    }
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(err, errlen, ABN_ERR_NO_GPU, "no CUDA device available (%s); this library has no CPU path",
                    e0 == cudaSuccess ? "device count is 0" : cudaGetErrorString(e0));
    }
    if (device < 0 || device >= ndev) return fail(err, errlen, ABN_ERR_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    CUDA_TRY(cudaSetDevice(device));
    keep_pool(device);
    PhaseTrace tr;
    abn_data *d = new (std::nothrow) abn_data();
    if (!d) return fail(err, errlen, ABN_ERR_ALLOC, "out of host memory");
    d->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) d->sm_count = prop.multiProcessorCount;
    d->family.assign(ds->node_family, ds->node_family + ds->n_nodes);
    d->col0.assign(ds->node_col0, ds->node_col0 + ds->n_nodes + 1);

    const long long ld = ((ds->n + 15) / 16) * 16; // 128-byte aligned columns
    int n_pois = 0;
    d->gw_index.assign(ds->n_nodes, -1);
    std::vector<int> pois_node;
    for (int i = 0; i < ds->n_nodes; i++) if (ds->node_family[i] == ABN_POISSON) { d->gw_index[i] = n_pois++; pois_node.push_back(i); }
    const int D = 1 + ds->n_cols + n_pois; // [1 | X | log(y + 0.1) of the Poisson nodes]
    int rc = ABN_OK;
#define TRY_OR_FREE(expr)                                                       \
    do {                                                                        \
        cudaError_t e_ = (expr);                                                \
        if (e_ != cudaSuccess) {                                                \
            rc = fail(err, errlen, e_ == cudaErrorMemoryAllocation ? ABN_ERR_ALLOC : ABN_ERR_CUDA, \
                      "%s: %s", #expr, cudaGetErrorString(e_));                 \
            cudaGetLastError();                                                 \
            abn_b200_data_destroy(d);                                           \
            return rc;                                                          \
        }                                                                       \
    } while (0)
    TRY_OR_FREE(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    for (auto &ss : d->side) TRY_OR_FREE(cudaStreamCreateWithFlags(&ss, cudaStreamNonBlocking));
    TRY_OR_FREE(dev_alloc(&d->dX, (size_t)ld * ds->n_cols * sizeof(double), d->stream));
    TRY_OR_FREE(dev_alloc(&d->dG, (size_t)D * D * sizeof(double), d->stream));
    TRY_OR_FREE(dev_alloc(&d->dLfact, (size_t)ds->n_nodes * sizeof(double), d->stream));
    bool any_gaussian = false;
    for (int i = 0; i < ds->n_nodes; i++) any_gaussian |= ds->node_family[i] == ABN_GAUSSIAN;
    if (any_gaussian) {
        TRY_OR_FREE(dev_alloc(&d->dGc, (size_t)(1 + ds->n_cols) * (1 + ds->n_cols) * sizeof(double), d->stream));
        TRY_OR_FREE(dev_alloc(&d->dColMean, (size_t)(1 + ds->n_cols) * sizeof(double), d->stream));
    }
    TRY_OR_FREE(dev_alloc(&d->dFamily, (size_t)ds->n_nodes * sizeof(int), d->stream));
    TRY_OR_FREE(dev_alloc(&d->dCol0, (size_t)(ds->n_nodes + 1) * sizeof(int), d->stream));
    TRY_OR_FREE(dev_alloc(&d->dGwIndex, (size_t)ds->n_nodes * sizeof(int), d->stream));
    TRY_OR_FREE(dev_alloc(&d->dPoisNode, std::max<size_t>(1, pois_node.size()) * sizeof(int), d->stream));
    TRY_OR_FREE(dev_alloc(&d->dBucketStart, (size_t)(NBUCKETS_ALL + 1) * sizeof(int), d->stream));
    TRY_OR_FREE(dev_alloc(&d->dCounters, (size_t)(NBUCKETS_ALL + 1) * BUCKET_SLOT_INTS * sizeof(int), d->stream));
    // ABN_B200_NO_CELLS=1: every fit takes the streaming kernels (A/B runs, debugging)
    const char *no_cells = std::getenv("ABN_B200_NO_CELLS");
    const bool use_cells = !(no_cells && *no_cells && *no_cells != '0');
    if (use_cells) {
        TRY_OR_FREE(dev_alloc(&d->dXq, (size_t)ld * ds->n_cols, d->stream));
        d->ldb = ((ds->n + 31) / 32 + 3) / 4 * 4;
        const char *no_planes = std::getenv("ABN_B200_NO_BITPLANES");
        if (!(no_planes && *no_planes && *no_planes != '0'))
            TRY_OR_FREE(dev_alloc(&d->dXb, (size_t)d->ldb * ds->n_cols * sizeof(unsigned), d->stream));
        TRY_OR_FREE(dev_alloc(&d->dColCard, (size_t)ds->n_cols * 3 * sizeof(int), d->stream)); // card | bad | max
        TRY_OR_FREE(dev_alloc(&d->dColTail, (size_t)ds->n_cols * 8 * sizeof(unsigned), d->stream));
        // overflow lists of the capped contingency tables: n / 8 rows per resident warp (3.5 GB for n = 1e6 on 148 SMs);
        // skipped - and with it the capping - when that would exceed 8 GB or no column can need it
        bool any_count = false;
        for (int i = 0; i < ds->n_nodes; i++) any_count |= ds->node_family[i] == ABN_POISSON;
        const size_t stride = (size_t)ds->n / 8 + 64;
        const size_t ovf_bytes = (size_t)abn_cell_ovf_slots(d->sm_count) * stride * sizeof(unsigned);
        const char *no_cap = std::getenv("ABN_B200_NO_CELL_CAPS");
        if (any_count && ovf_bytes <= (8ull << 30) && !(no_cap && *no_cap && *no_cap != '0'))
            TRY_OR_FREE(dev_alloc(&d->dCellOvf, ovf_bytes, d->stream));
    }
    tr.mark("create: streams + cudaMalloc");
    TRY_OR_FREE(cudaMemsetAsync(d->dX, 0, (size_t)ld * ds->n_cols * sizeof(double), d->stream));
    TRY_OR_FREE(cudaMemcpy2DAsync(d->dX, (size_t)ld * sizeof(double), ds->X, (size_t)ds->n * sizeof(double),
                                  (size_t)ds->n * sizeof(double), (size_t)ds->n_cols, cudaMemcpyHostToDevice, d->stream));
    TRY_OR_FREE(cudaMemcpyAsync(d->dFamily, ds->node_family, (size_t)ds->n_nodes * sizeof(int), cudaMemcpyHostToDevice, d->stream));
    TRY_OR_FREE(cudaMemcpyAsync(d->dCol0, ds->node_col0, (size_t)(ds->n_nodes + 1) * sizeof(int), cudaMemcpyHostToDevice, d->stream));
    TRY_OR_FREE(cudaMemcpyAsync(d->dGwIndex, d->gw_index.data(), (size_t)ds->n_nodes * sizeof(int), cudaMemcpyHostToDevice, d->stream));
    if (n_pois) TRY_OR_FREE(cudaMemcpyAsync(d->dPoisNode, pois_node.data(), pois_node.size() * sizeof(int), cudaMemcpyHostToDevice, d->stream));
    TRY_OR_FREE(cudaStreamSynchronize(d->stream));
    tr.mark("create: upload");
#undef TRY_OR_FREE
    d->m.n = ds->n;
    d->m.ld = ld;
    d->m.n_nodes = ds->n_nodes;
    d->m.n_cols = ds->n_cols;
    d->m.D = D;
    d->m.X = d->dX;
    d->m.G = d->dG;
    d->m.node_family = d->dFamily;
    d->m.node_col0 = d->dCol0;
    d->m.node_lfact = d->dLfact;
    d->m.node_gw = d->dGwIndex;
    d->m.n_pois = n_pois;
    d->m.pois_node = d->dPoisNode;
    d->m.Xq = nullptr;
    d->m.ldq = ld;
    d->m.col_card = nullptr;
    d->m.Xb = nullptr;
    d->m.ldb = 0;
    d->m.col_tail = nullptr;
    d->m.cell_ovf = nullptr;
    d->m.cell_ovf_stride = 0;
    d->m.Gc = nullptr;
    d->m.col_mean = nullptr;
    if (d->dXq) {
        int *bad = d->dColCard + ds->n_cols, *mx = d->dColCard + 2 * ds->n_cols;
        cudaError_t e = cudaMemsetAsync(d->dColCard, 0, (size_t)ds->n_cols * 3 * sizeof(int), d->stream);
        if (e == cudaSuccess) {
            const long long groups = (ds->n + 15) / 16;
            discretize_kernel<<<dim3((unsigned)((groups + 255) / 256), ds->n_cols), 256, 0, d->stream>>>(d->dX, ld, ds->n, d->dXq, ld, bad, mx);
            col_card_kernel<<<(ds->n_cols + 127) / 128, 128, 0, d->stream>>>(ds->n_cols, bad, mx, d->dColCard);
            if (d->dXb)
                bitplane_kernel<<<dim3((unsigned)((d->ldb * 32 + 255) / 256), ds->n_cols), 256, 0, d->stream>>>(d->dXq, ld, ds->n, d->dColCard,
                                                                                                         d->dXb, d->ldb);
            e = cudaMemsetAsync(d->dColTail, 0, (size_t)ds->n_cols * 8 * sizeof(unsigned), d->stream);
            col_tail_kernel<<<dim3((unsigned)((groups + 255) / 256), ds->n_cols), 256, 0, d->stream>>>(d->dXq, ld, ds->n, d->dColTail);
            if (e == cudaSuccess) e = cudaGetLastError();
        }
        if (e != cudaSuccess) {
            rc = fail(err, errlen, ABN_ERR_CUDA, "discretize: %s", cudaGetErrorString(e));
            abn_b200_data_destroy(d);
            return rc;
        }
        d->timing.launches += d->dXb ? 4 : 3;
        d->m.Xb = d->dXb;
        d->m.ldb = d->ldb;
        d->m.Xq = d->dXq;
        d->m.col_card = d->dColCard;
        d->m.col_tail = d->dColTail;
        d->m.cell_ovf = d->dCellOvf;
        d->m.cell_ovf_stride = (long long)(ds->n / 8 + 64);
    }
    rc = compute_gram(d, err, errlen);
    if (rc != ABN_OK) { abn_b200_data_destroy(d); return rc; }
    tr.mark("create: Gram + lfact");
    *out = d;
    return ABN_OK;
}

static int ensure_scratch(abn_data *d, long long rows, int width, char *err, size_t errlen)
{
    if (rows <= d->cap_rows && width <= d->cap_width) return ABN_OK;
    free_scratch(d);
    const long long cap = std::max<long long>(rows, 1024);
    const int w = std::max(width, 1);
    CUDA_TRY(dev_alloc(&d->dChild, (size_t)cap * sizeof(int), d->stream));
    CUDA_TRY(dev_alloc(&d->dParents, (size_t)cap * w * sizeof(int), d->stream));
    CUDA_TRY(dev_alloc(&d->dIdx, (size_t)cap * sizeof(int), d->stream));
    CUDA_TRY(dev_alloc(&d->dIdxSorted, (size_t)cap * sizeof(int), d->stream));
    CUDA_TRY(dev_alloc(&d->dStatus, (size_t)cap * sizeof(int), d->stream));
    CUDA_TRY(dev_alloc(&d->dIters, (size_t)cap * sizeof(int), d->stream));
    CUDA_TRY(dev_alloc(&d->dKeys, (size_t)cap, d->stream));
    CUDA_TRY(dev_alloc(&d->dKeysSorted, (size_t)cap, d->stream));
    CUDA_TRY(dev_alloc(&d->dDescs, (size_t)cap * sizeof(FitDesc), d->stream));
    CUDA_TRY(dev_alloc(&d->dRes, (size_t)cap * 4 * sizeof(double), d->stream));
    size_t tmp = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp, d->dKeys, d->dKeysSorted, d->dIdx, d->dIdxSorted, (int)cap, 0, 8, d->stream));
    CUDA_TRY(dev_alloc(&d->dSortTmp, tmp, d->stream));
    d->sortTmpBytes = tmp;
    d->cap_rows = cap;
    d->cap_width = w;
    return ABN_OK;
}

// ---------------------------------------------------------------------------------------------
// per-call bookkeeping on the device: candidate validation and the roofline account (no host sweeps)
// ---------------------------------------------------------------------------------------------
// first invalid row (child out of range, parent out of range / below -1 / equal to the child), or INT_MAX
__global__ void validate_kernel(long long rows, int width, int n_nodes, const int *__restrict__ child,
                                const int *__restrict__ parents, int *__restrict__ first_bad)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int c = child[r];
    bool bad = c < 0 || c >= n_nodes;
    for (int j = 0; j < width; j++) {
        const int p = parents[r * width + j];
        bad |= p < -1 || p >= n_nodes || p == c;
    }
    if (bad) atomicMin(first_bad, (int)r);
}

// per bucket: sum over its fits of passes x (k_data + response columns) - times n x 8 that is SURVEY 8d's
// algorithmic bytes - and, for the contingency-table buckets, of the byte columns read once
__global__ void __launch_bounds__(256)
account_kernel(long long rows, const unsigned char *__restrict__ keys_sorted, const int *__restrict__ idx_sorted,
               const FitDesc *__restrict__ descs, const int *__restrict__ iters, unsigned long long *__restrict__ units)
{
    __shared__ unsigned long long acc[NBUCKETS_ALL];
    for (int i = threadIdx.x; i < NBUCKETS_ALL; i += blockDim.x) acc[i] = 0ull;
    __syncthreads();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < rows) {
        const int key = keys_sorted[t], row = idx_sorted[t];
        const FitDesc d = descs[row];
        unsigned long long u;
        if (key >= CELL_BUCKET0) u = (unsigned long long)(d.kdata + (d.family == BINOMIAL ? 1 : 0));
        else u = (unsigned long long)iters[row] * (unsigned long long)(d.kdata + d.nlev);
        atomicAdd(&acc[key], u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NBUCKETS_ALL; i += blockDim.x)
        if (acc[i]) atomicAdd(&units[i], acc[i]);
}

// Everything one abn_b200_score call creates; released on every exit path. The destructor first drains the
// main and the side streams, so no kernel is still writing into a buffer that is being freed.
struct ScoreCall {
    abn_data *d;
    std::vector<cudaEvent_t> events;
    std::vector<void *> buffers;
    explicit ScoreCall(abn_data *dd) : d(dd) {}
    cudaError_t event(cudaEvent_t *e, unsigned flags = cudaEventDefault)
    {
        cudaError_t rc = cudaEventCreateWithFlags(e, flags);
        if (rc == cudaSuccess) events.push_back(*e);
        return rc;
    }
    template <typename T>
    cudaError_t alloc(T **p, size_t bytes)
    {
        cudaError_t rc = dev_alloc(p, bytes, d->stream);
        if (rc == cudaSuccess) buffers.push_back(*p);
        return rc;
    }
    ~ScoreCall()
    {
        for (auto &ss : d->side) if (ss) cudaStreamSynchronize(ss);
        for (void *p : buffers) dev_free(p, d->stream);
        cudaStreamSynchronize(d->stream);
        for (auto &e : events) cudaEventDestroy(e);
        cudaGetLastError();
    }
};

// The scoring pipeline. cand->child == NULL with resident = true: the candidates are already in d->dChild /
// d->dParents (abn_b200_score_enumerated generated them there) - nothing is uploaded or validated.
static int score_impl(abn_data *d, const abn_candidates *cand, bool resident, const abn_control *ctrl, abn_result *out,
                      char *err, size_t errlen)
{
    if (!d || !cand || !ctrl || !out) return fail(err, errlen, ABN_ERR_ARG, "null argument");
    if (cand->rows < 0 || cand->max_parents < 0) return fail(err, errlen, ABN_ERR_ARG, "negative candidate count / width");
    if (cand->rows > 2000000000LL) return fail(err, errlen, ABN_ERR_LIMIT, "more than 2e9 candidates in one call; split the list");
    if (ctrl->method != ABN_METHOD_MLE && ctrl->method != ABN_METHOD_BAYES) return fail(err, errlen, ABN_ERR_ARG, "unknown method %d", ctrl->method);
    if (ctrl->method == ABN_METHOD_BAYES)
        for (int f : d->family)
            if (f == ABN_MULTINOMIAL)
                return fail(err, errlen, ABN_ERR_ARG, "Multinomial nodes are not yet implemented for method 'bayes'. Try with method='mle'.");
    const bool on_device = (ctrl->flags & ABN_FLAG_RESULT_ON_DEVICE) != 0;
    const int coef_stride = out->coef ? out->coef_stride : 0;
    if (out->coef && coef_stride < 1) return fail(err, errlen, ABN_ERR_ARG, "coef_stride must be positive when coef is given");
    if (out->se && !out->coef) return fail(err, errlen, ABN_ERR_ARG, "se needs coef (they share coef_stride)");
    if (on_device && (out->coef || out->se || out->sse))
        return fail(err, errlen, ABN_ERR_ARG, "ABN_FLAG_RESULT_ON_DEVICE covers mlik / aic / bic / mdl / status / iters only");
    if (on_device && !out->mlik) return fail(err, errlen, ABN_ERR_ARG, "ABN_FLAG_RESULT_ON_DEVICE needs a device pointer for mlik");
    const long long rows = cand->rows;
    const int width = cand->max_parents;
    d->timing = abn_timing{};
    if (rows == 0) return ABN_OK;
    if (!resident && (!cand->child || (width > 0 && !cand->parents))) return fail(err, errlen, ABN_ERR_ARG, "null candidate arrays");
    CUDA_TRY(cudaSetDevice(d->device));
    PhaseTrace tr;
    int rc = ensure_scratch(d, rows, width, err, errlen);
    if (rc != ABN_OK) return rc;
    tr.mark("score: scratch");

    ScoreCall call(d);
    cudaStream_t s = d->stream;
    cudaEvent_t ev[5];
    for (auto &e : ev) CUDA_TRY(call.event(&e));
    long long launches = 0;

    CUDA_TRY(cudaEventRecord(ev[0], s));
    if (!resident) {
        CUDA_TRY(cudaMemcpyAsync(d->dChild, cand->child, (size_t)rows * sizeof(int), cudaMemcpyHostToDevice, s));
        if (width > 0) CUDA_TRY(cudaMemcpyAsync(d->dParents, cand->parents, (size_t)rows * width * sizeof(int), cudaMemcpyHostToDevice, s));
    }
    CUDA_TRY(cudaEventRecord(ev[1], s));

    Control c{ctrl->method, ctrl->max_iters, ctrl->solver, ctrl->nvars, ctrl->tol, ctrl->epsabs,
              ctrl->prior_mean, ctrl->prior_sd, ctrl->gam_shape, ctrl->gam_scale, ctrl->flags};
    double *dCoef = nullptr, *dSe = nullptr, *dSse = nullptr;
    if (out->coef) {
        CUDA_TRY(call.alloc(&dCoef, (size_t)rows * coef_stride * sizeof(double)));
        CUDA_TRY(cudaMemsetAsync(dCoef, 0xff, (size_t)rows * coef_stride * sizeof(double), s)); // all-ones = NaN
    }
    if (out->se) {
        CUDA_TRY(call.alloc(&dSe, (size_t)rows * coef_stride * sizeof(double)));
        CUDA_TRY(cudaMemsetAsync(dSe, 0xff, (size_t)rows * coef_stride * sizeof(double), s));
    }
    if (out->sse) {
        CUDA_TRY(call.alloc(&dSse, (size_t)rows * sizeof(double)));
        CUDA_TRY(cudaMemsetAsync(dSse, 0xff, (size_t)rows * sizeof(double), s));
    }
    // where the kernels write: the handle's scratch slabs, or straight into the caller's device buffers
    ResultPtrs rp{d->dRes, d->dRes + d->cap_rows, d->dRes + 2 * d->cap_rows, d->dRes + 3 * d->cap_rows, d->dStatus, d->dIters,
                  dCoef, coef_stride, dSe, dSse};
    if (on_device) {
        rp.mlik = out->mlik; rp.aic = out->aic; rp.bic = out->bic; rp.mdl = out->mdl;
        rp.status = out->status ? out->status : d->dStatus;
        rp.iters = out->iters ? out->iters : d->dIters;
    }

    // per-call device words: [0] first invalid row, then per bucket the roofline units (account_kernel)
    unsigned long long *dAcct = nullptr;
    CUDA_TRY(call.alloc(&dAcct, (size_t)(1 + NBUCKETS_ALL) * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(dAcct, 0, (size_t)(1 + NBUCKETS_ALL) * sizeof(unsigned long long), s));
    int *dFirstBad = reinterpret_cast<int *>(dAcct);
    const int int_max = 2147483647;
    CUDA_TRY(cudaMemcpyAsync(dFirstBad, &int_max, sizeof(int), cudaMemcpyHostToDevice, s));

    const int tb = 128;
    int first_bad = int_max;
    if (!resident) {
        validate_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(rows, width, d->m.n_nodes, d->dChild, d->dParents, dFirstBad);
        KCHECK("validate_kernel", s);
        CUDA_TRY(cudaMemcpyAsync(&first_bad, dFirstBad, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        launches++;
    }
    if (first_bad != int_max) {
        const long long r = first_bad;
        const int ch = cand->child[r];
        if (ch < 0 || ch >= d->m.n_nodes) return fail(err, errlen, ABN_ERR_ARG, "row %lld: child %d out of range", r, ch);
        for (int j = 0; j < width; j++) {
            const int p = cand->parents[r * width + j];
            if (p < -1 || p >= d->m.n_nodes || p == ch) return fail(err, errlen, ABN_ERR_ARG, "row %lld: parent %d invalid", r, p);
        }
        return fail(err, errlen, ABN_ERR_ARG, "row %lld: invalid candidate", r);
    }
    resolve_kernel<<<(unsigned)((rows + tb - 1) / tb), tb, 0, s>>>(d->m, c.method, c.flags, rows, width, d->dChild, d->dParents,
                                                                  d->dDescs, d->dKeys, d->dIdx);
    KCHECK("resolve_kernel", s);
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(d->dSortTmp, d->sortTmpBytes, d->dKeys, d->dKeysSorted, d->dIdx, d->dIdxSorted,
                                             (int)rows, 0, 8, s));
    CUDA_TRY(cudaMemsetAsync(d->dBucketStart, 0xff, (size_t)(NBUCKETS_ALL + 1) * sizeof(int), s));
    CUDA_TRY(cudaMemsetAsync(d->dCounters, 0, (size_t)(NBUCKETS_ALL + 1) * BUCKET_SLOT_INTS * sizeof(int), s));
    KCHECK("radix sort", s);
    bucket_bounds<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(d->dKeysSorted, rows, d->dBucketStart);
    KCHECK("bucket_bounds", s);
    launches += 3;
    int hstart[NBUCKETS_ALL + 1];
    CUDA_TRY(cudaMemcpyAsync(hstart, d->dBucketStart, sizeof(hstart), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    CUDA_TRY(cudaGetLastError());
    hstart[NBUCKETS_ALL] = (int)rows;
    for (int b = NBUCKETS_ALL - 1; b >= 0; b--) if (hstart[b] < 0) hstart[b] = hstart[b + 1];
    CUDA_TRY(cudaEventRecord(ev[2], s));

    int n_unsupported = 0;
    // Launch order: buckets too small to fill the GPU go first, on side streams, so that they run in
    // the gaps of the large ones (each large kernel is a persistent grid that drains its own queue);
    // the large buckets follow on the main stream in cache order. Per-kernel CUDA events are recorded
    // on the stream the kernel runs on.
    std::vector<cudaEvent_t> kev0, kev1;
    std::vector<int> kid, kcount;
    const int small_limit = d->sm_count * ABN_MIN_BLOCKS * FIT_WARPS * 2;
    cudaEvent_t planned;
    CUDA_TRY(call.event(&planned, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(planned, s));
    for (auto &ss : d->side) CUDA_TRY(cudaStreamWaitEvent(ss, planned, 0));
    int n_side = 0;
    for (int pass = 0; pass < 2; pass++) {
        for (int b = 0; b < NBUCKETS_ALL; b++) {
            const int count = hstart[b + 1] - hstart[b];
            if (count <= 0) continue;
            // pass 0: the Gram-only Gaussian kernel (largest per-thread stack: launching it first
            // keeps the one-off local-memory pool resize from synchronising with running kernels) and
            // the small buckets on side streams; pass 1: the large buckets on the main stream
            const bool small = (b >= 2) && count < small_limit;
            const bool early = small || b <= 1;
            if ((pass == 0) != early) continue;
            cudaStream_t ks = small ? d->side[n_side++ % abn_data::NSIDE] : s;
            const bool timed = ((int)kid.size() < ABN_MAX_KERNELS);
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (timed) {
                CUDA_TRY(call.event(&e0));
                CUDA_TRY(call.event(&e1));
                CUDA_TRY(cudaEventRecord(e0, ks));
            }
            const int *list = d->dIdxSorted + hstart[b];
            int *slot = d->dCounters + b * BUCKET_SLOT_INTS;
            if (b == 0) {
                // not for the register-resident kernels: wider than 8 columns (-> generic wide kernel,
                // which re-resolves the design with its own limit and flags what it cannot take)
                CUDA_TRY(abn_wide_launch(d->m, c, d->dChild, d->dParents, width, list, count, slot, rp, d->sm_count, ks));
            } else if (b == 1) {
                gauss_kernel<<<(count + 127) / 128, 128, 0, ks>>>(d->m, c, d->dDescs, d->dChild, list, count, rp);
            } else if (b >= CELL_BUCKET0) {
                const int kind = (b >= CELL_BUCKET0 + 3) ? RK_POIS : (c.method == MLE ? RK_BIN_MLE : RK_BIN_BAYES);
                CUDA_TRY(abn_cell_launch(kind, (b - CELL_BUCKET0) % 3, d->m, c, d->dDescs, d->dChild, list, count, slot, rp, d->sm_count, ks));
            } else {
                const int code = b - 2, icpt = code & 1, fk = code >> 1, fam = fk / (MAXK + 1), k = fk % (MAXK + 1);
                FitLaunchFn fn = nullptr;
                MnLaunchFn mfn = nullptr;
                if (fam == 0) fn = (c.method == MLE) ? get_launch_bin_mle(k, icpt) : get_launch_bin_bayes(k, icpt);
                else if (fam == 1) fn = get_launch_pois(k, icpt);
                else mfn = get_launch_mn(fam, k, icpt);
                if (mfn) {
                    int grid = std::min((count + MN_WARPS - 1) / MN_WARPS, d->sm_count * 2);
                    CUDA_TRY(mfn(d->m, c, d->dDescs, list, count, slot, rp, grid, ks));
                } else if (!fn) {
                    n_unsupported += count;
                    unsupported_kernel<<<(count + 255) / 256, 256, 0, ks>>>(list, count, rp);
                } else {
                    int grid = std::min((count + FIT_WARPS - 1) / FIT_WARPS, d->sm_count * ABN_MIN_BLOCKS);
                    CUDA_TRY(fn(d->m, c, d->dDescs, d->dChild, list, count, slot, rp, grid, ks));
                }
            }
            {
                char what[64];
                snprintf(what, sizeof(what), "fit kernel bucket %d (%d fits)", b, count);
                KCHECK(what, ks);
            }
            if (timed) {
                CUDA_TRY(cudaEventRecord(e1, ks));
                kev0.push_back(e0);
                kev1.push_back(e1);
                kid.push_back(b);
                kcount.push_back(count);
            }
            launches++;
        }
    }
    for (int i = 0; i < std::min(n_side, (int)abn_data::NSIDE); i++) {
        cudaEvent_t done;
        CUDA_TRY(call.event(&done, cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(done, d->side[i]));
        CUDA_TRY(cudaStreamWaitEvent(s, done, 0));
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ev[3], s));
    // roofline account on the device (a few hundred bytes come back instead of two O(rows) host sweeps)
    account_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(rows, d->dKeysSorted, d->dIdxSorted, d->dDescs, rp.iters, dAcct + 1);
    KCHECK("account_kernel", s);
    launches++;
    unsigned long long units[NBUCKETS_ALL];
    std::vector<int> slots((size_t)NBUCKETS_ALL * BUCKET_SLOT_INTS);
    CUDA_TRY(cudaMemcpyAsync(units, dAcct + 1, sizeof(units), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(slots.data(), d->dCounters, slots.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (!on_device) {
        if (out->mlik) CUDA_TRY(cudaMemcpyAsync(out->mlik, rp.mlik, (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (ctrl->method == ABN_METHOD_MLE) {
            if (out->aic) CUDA_TRY(cudaMemcpyAsync(out->aic, rp.aic, (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost, s));
            if (out->bic) CUDA_TRY(cudaMemcpyAsync(out->bic, rp.bic, (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost, s));
            if (out->mdl) CUDA_TRY(cudaMemcpyAsync(out->mdl, rp.mdl, (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        if (out->status) CUDA_TRY(cudaMemcpyAsync(out->status, rp.status, (size_t)rows * sizeof(int), cudaMemcpyDeviceToHost, s));
        if (out->iters) CUDA_TRY(cudaMemcpyAsync(out->iters, rp.iters, (size_t)rows * sizeof(int), cudaMemcpyDeviceToHost, s));
        if (dCoef) CUDA_TRY(cudaMemcpyAsync(out->coef, dCoef, (size_t)rows * coef_stride * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (dSe) CUDA_TRY(cudaMemcpyAsync(out->se, dSe, (size_t)rows * coef_stride * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (dSse) CUDA_TRY(cudaMemcpyAsync(out->sse, dSse, (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaEventRecord(ev[4], s));
    CUDA_TRY(cudaStreamSynchronize(s));
    CUDA_TRY(cudaGetLastError());
    tr.mark("score: upload + plan + fits + download");

    float ms[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
    d->timing.h2d_ms = ms[0];
    d->timing.plan_ms = ms[1];
    d->timing.fit_ms = ms[2];
    d->timing.d2h_ms = ms[3];
    d->timing.launches = launches;
    // algorithmic bytes: passes * n * (k_data + 1) * 8 per streamed fit; a fit on a contingency table
    // (fit_cell.cuh) reads its byte-coded columns once: n * (k_data + 1 for a binomial response)
    long long bytes = 0;
    d->timing.n_kernels = (int)kid.size();
    for (size_t i = 0; i < kid.size(); i++) {
        const int b = kid[i];
        float kms = 0;
        cudaEventElapsedTime(&kms, kev0[i], kev1[i]);
        const long long kb = (long long)units[b] * d->m.n * (b >= CELL_BUCKET0 ? 1 : 8);
        bytes += kb;
        long long fp64 = 0;
        if (b >= 2 && b < CELL_BUCKET0) {
            const int fk = (b - 2) >> 1, fam = fk / (MAXK + 1), k = fk % (MAXK + 1);
            if (fam <= 1) {
                const unsigned long long *tally = reinterpret_cast<const unsigned long long *>(slots.data() + (size_t)b * BUCKET_SLOT_INTS + 2);
                for (int mode = 1; mode <= 6; mode++) fp64 += (long long)tally[mode] * d->m.n * fp64_per_row(fam, k, mode);
            }
        }
        d->timing.kernel_id[i] = b;
        d->timing.kernel_fits[i] = kcount[i];
        d->timing.kernel_bytes[i] = kb;
        d->timing.kernel_ms[i] = kms;
        d->timing.kernel_fp64[i] = fp64;
    }
    d->timing.algo_bytes = bytes;
    tr.mark("score: timing bookkeeping");
    if (n_unsupported > 0 && n_unsupported == rows)
        return fail(err, errlen, ABN_ERR_LIMIT, "no candidate could be scored by the fast kernels");
    return ABN_OK;
}

extern "C" int abn_b200_score(abn_data *d, const abn_candidates *cand, const abn_control *ctrl, abn_result *out,
                              char *err, size_t errlen)
{
    return score_impl(d, cand, false, ctrl, out, err, errlen);
}

// ---------------------------------------------------------------------------------------------
// candidate enumeration on the device: the list of abn_b200_enumerate, row r computed from r
// ---------------------------------------------------------------------------------------------
struct EnumNode {
    int child, n_free, n_must, jmax;
    long long row0; // first row of this node in the list
};

// Row r of the list: node by binary search over row0, parent count j by the node's per-size offsets, then
// the (r - offset)-th (j - n_must)-subset of the node's free parents in lexicographic order (combinatorial
// unranking with a binomial table), merged with the retained parents. A subset that contains the retained
// parents precedes another one exactly when its free part does (the smallest element of the symmetric
// difference decides either comparison), so this is the order buildcachematrix visits the survivors in.
__global__ void enumerate_kernel(long long rows, int n_list, int d, int W, const EnumNode *__restrict__ nodes,
                                 const long long *__restrict__ size_row0 /* [n_list][W + 2] */,
                                 const int *__restrict__ free_list, const int *__restrict__ must_list,
                                 const unsigned long long *__restrict__ binom /* [d + 1][W + 1] */,
                                 int *__restrict__ child, int *__restrict__ parents)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    int lo = 0, hi = n_list - 1;
    while (lo < hi) { // last node with row0 <= r
        const int mid = (lo + hi + 1) >> 1;
        if (nodes[mid].row0 <= r) lo = mid; else hi = mid - 1;
    }
    const EnumNode nd = nodes[lo];
    const long long *off = size_row0 + (long long)lo * (W + 2);
    int j = 0;
    while (j < nd.jmax && off[j + 1] <= r) j++;
    unsigned long long q = (unsigned long long)(r - off[j]);
    const int t = j - nd.n_must; // free parents to choose
    const int *fl = free_list + (long long)lo * d, *ml = must_list + (long long)lo * d;
    int *row = parents + r * W;
    int x = 0, mi = 0, o = 0;
    for (int i = 0; i < t; i++) {
        for (;;) { // subsets that start with free[x] at position i: C(n_free - x - 1, t - i - 1)
            const unsigned long long c = binom[(long long)(nd.n_free - x - 1) * (W + 1) + (t - i - 1)];
            if (q < c) break;
            q -= c;
            x++;
        }
        const int p = fl[x++];
        while (mi < nd.n_must && ml[mi] < p) row[o++] = ml[mi++];
        row[o++] = p;
    }
    while (mi < nd.n_must) row[o++] = ml[mi++];
    while (o < W) row[o++] = -1;
    child[r] = nd.child;
}

extern "C" int abn_b200_score_enumerated(abn_data *d, const int32_t *banned, const int32_t *retained,
                                         const int32_t *max_parents, const int32_t *which_nodes, int32_t n_which,
                                         const abn_control *ctrl, abn_result *out, int64_t rows_expected, int32_t width,
                                         int32_t *children, int32_t *parents, char *err, size_t errlen)
{
    if (!d || !max_parents || !ctrl || !out) return fail(err, errlen, ABN_ERR_ARG, "null argument");
    if ((children == nullptr) != (parents == nullptr)) return fail(err, errlen, ABN_ERR_ARG, "children and parents must be given together");
    const int dn = d->m.n_nodes;
    const int n_list = which_nodes ? n_which : dn;
    // per listed node: free and retained parents (ascending), rows per parent count - O(n_list x d) host work
    int W = 0;
    for (int w = 0; w < n_list; w++) {
        const int node = which_nodes ? which_nodes[w] : w;
        if (node < 0 || node >= dn) return fail(err, errlen, ABN_ERR_ARG, "which_nodes[%d] = %d out of range", w, node);
        if (max_parents[node] < 0) return fail(err, errlen, ABN_ERR_ARG, "max_parents[%d] < 0", node);
        W = std::max(W, std::min<int>(max_parents[node], dn));
    }
    if (width < W) return fail(err, errlen, ABN_ERR_ARG, "width %d too small, need %d", width, W);
    const int rowW = std::max(width, 1);
    std::vector<unsigned long long> binom((size_t)(dn + 1) * (rowW + 1), 0ull);
    for (int m = 0; m <= dn; m++)
        for (int t = 0; t <= rowW; t++) {
            unsigned long long v = (t == 0) ? 1ull : ((m == 0) ? 0ull : binom[(size_t)(m - 1) * (rowW + 1) + t - 1] + binom[(size_t)(m - 1) * (rowW + 1) + t]);
            binom[(size_t)m * (rowW + 1) + t] = std::min(v, 1ull << 62); // saturate: such lists are refused below anyway
        }
    std::vector<EnumNode> nodes;
    std::vector<long long> size_row0;
    std::vector<int> free_list, must_list;
    long long total = 0;
    for (int w = 0; w < n_list; w++) {
        const int node = which_nodes ? which_nodes[w] : w;
        std::vector<int> fr, mu;
        bool impossible = false;
        for (int p = 0; p < dn; p++) {
            const bool ban = banned && banned[node + (long long)p * dn] == 1;
            const bool ret = retained && retained[node + (long long)p * dn] == 1;
            if (ret && (p == node || ban)) impossible = true; // a retained arc that can never be present
            if (ret) mu.push_back(p);
            else if (p != node && !ban) fr.push_back(p);
        }
        if (impossible) continue;
        EnumNode e{node, (int)fr.size(), (int)mu.size(), std::min<int>(max_parents[node], dn), total};
        std::vector<long long> off(rowW + 2, 0);
        for (int j = 0; j <= rowW; j++) {
            off[j] = total;
            if (j <= e.jmax && j >= e.n_must && j - e.n_must <= e.n_free)
                total = std::min<long long>(total + (long long)binom[(size_t)e.n_free * (rowW + 1) + (j - e.n_must)], 1ll << 62);
        }
        off[rowW + 1] = total;
        if (total == e.row0) continue; // no candidate for this node
        nodes.push_back(e);
        size_row0.insert(size_row0.end(), off.begin(), off.end());
        fr.resize(dn, -1);
        mu.resize(dn, -1);
        free_list.insert(free_list.end(), fr.begin(), fr.end());
        must_list.insert(must_list.end(), mu.begin(), mu.end());
    }
    if (total != rows_expected)
        return fail(err, errlen, ABN_ERR_ARG, "the enumeration has %lld rows, the caller expects %lld (abn_b200_enumerate reports the count)",
                    total, (long long)rows_expected);
    d->timing = abn_timing{};
    if (total == 0) return ABN_OK;
    if (total > 2000000000LL) return fail(err, errlen, ABN_ERR_LIMIT, "more than 2e9 candidates in one call; split the list");
    CUDA_TRY(cudaSetDevice(d->device));
    int rc = ensure_scratch(d, total, rowW, err, errlen);
    if (rc != ABN_OK) return rc;
    {
        ScoreCall call(d);
        cudaStream_t s = d->stream;
        EnumNode *dNodes = nullptr;
        long long *dOff = nullptr;
        int *dFree = nullptr, *dMust = nullptr;
        unsigned long long *dBinom = nullptr;
        CUDA_TRY(call.alloc(&dNodes, nodes.size() * sizeof(EnumNode)));
        CUDA_TRY(call.alloc(&dOff, size_row0.size() * sizeof(long long)));
        CUDA_TRY(call.alloc(&dFree, free_list.size() * sizeof(int)));
        CUDA_TRY(call.alloc(&dMust, must_list.size() * sizeof(int)));
        CUDA_TRY(call.alloc(&dBinom, binom.size() * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemcpyAsync(dNodes, nodes.data(), nodes.size() * sizeof(EnumNode), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(dOff, size_row0.data(), size_row0.size() * sizeof(long long), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(dFree, free_list.data(), free_list.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(dMust, must_list.data(), must_list.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(dBinom, binom.data(), binom.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
        enumerate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(total, (int)nodes.size(), dn, rowW, dNodes, dOff, dFree, dMust,
                                                                      dBinom, d->dChild, d->dParents);
        KCHECK("enumerate_kernel", s);
        if (children) {
            CUDA_TRY(cudaMemcpyAsync(children, d->dChild, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaMemcpyAsync(parents, d->dParents, (size_t)total * rowW * sizeof(int), cudaMemcpyDeviceToHost, s));
        }
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    abn_candidates cand{total, rowW, nullptr, nullptr};
    rc = score_impl(d, &cand, true, ctrl, out, err, errlen);
    d->timing.launches += 1;
    return rc;
}

static int score_cache_on(const abn_dataset *ds, const abn_candidates *cand, const abn_control *ctrl, int device,
                          abn_result *out, char *err, size_t errlen)
{
    abn_data *d = nullptr;
    int rc = abn_b200_data_create(ds, device, &d, err, errlen);
    if (rc != ABN_OK) return rc;
    rc = abn_b200_score(d, cand, ctrl, out, err, errlen);
    abn_b200_data_destroy(d);
    return rc;
}

// rows device `rank` of `world` owns under the block-cyclic split: blocks rank, rank + world, ...
static long long shard_rows(long long rows, int block, int world, int rank)
{
    const long long nblocks = (rows + block - 1) / block;
    long long mine = 0;
    for (long long b = rank; b < nblocks; b += world) mine += std::min<long long>(block, rows - b * block);
    return mine;
}

extern "C" int abn_b200_score_cache(const abn_dataset *ds, const abn_candidates *cand, const abn_control *ctrl,
                                    abn_result *out, char *err, size_t errlen)
{
    if (!ctrl || !cand || !out) return fail(err, errlen, ABN_ERR_ARG, "null argument");
    if (ctrl->flags & ABN_FLAG_RESULT_ON_DEVICE)
        return fail(err, errlen, ABN_ERR_ARG, "ABN_FLAG_RESULT_ON_DEVICE needs a resident data set (abn_b200_score)");
    const int world = ctrl->ngpus;
    if (world <= 1) return score_cache_on(ds, cand, ctrl, world == 1 ? ctrl->devices[0] : ctrl->device, out, err, errlen);

    // ---- one host thread per device; every device holds the data set and scores its blocks of the list ----------
    if (world > ABN_MAX_DEVICES) return fail(err, errlen, ABN_ERR_ARG, "ngpus = %d exceeds ABN_MAX_DEVICES = %d", world, ABN_MAX_DEVICES);
    const int ndev = abn_b200_device_count();
    if (ndev == 0) return fail(err, errlen, ABN_ERR_NO_GPU, "no CUDA device available; this library has no CPU path");
    for (int i = 0; i < world; i++) {
        if (ctrl->devices[i] < 0 || ctrl->devices[i] >= ndev)
            return fail(err, errlen, ABN_ERR_ARG, "devices[%d] = %d out of range (0..%d)", i, ctrl->devices[i], ndev - 1);
        for (int j = 0; j < i; j++)
            if (ctrl->devices[j] == ctrl->devices[i]) return fail(err, errlen, ABN_ERR_ARG, "device %d listed twice", ctrl->devices[i]);
    }
    if (out->coef || out->se || out->sse)
        return fail(err, errlen, ABN_ERR_ARG, "coef / se / sse are single-device outputs (ngpus <= 1)");
    if (cand->rows < 0 || cand->max_parents < 0) return fail(err, errlen, ABN_ERR_ARG, "negative candidate count / width");
    if (cand->rows > 0 && (!cand->child || (cand->max_parents > 0 && !cand->parents)))
        return fail(err, errlen, ABN_ERR_ARG, "null candidate arrays");
    const long long rows = cand->rows;
    const int width = cand->max_parents;
    const int block = ctrl->block > 0 ? ctrl->block : 1024;
    const long long nblocks = (rows + block - 1) / block;

    std::vector<int> rcs(world, ABN_OK);
    std::vector<std::string> msgs(world);
    auto work = [&](int rank) {
        char lerr[512] = {0};
        const long long mine = shard_rows(rows, block, world, rank);
        std::vector<int32_t> child(mine), parents((size_t)mine * std::max(width, 1)), status(mine), iters(mine);
        std::vector<double> val[4];
        long long o = 0;
        for (long long b = rank; b < nblocks; b += world) { // gather this device's blocks
            const long long r0 = b * block, len = std::min<long long>(block, rows - r0);
            std::memcpy(child.data() + o, cand->child + r0, (size_t)len * sizeof(int32_t));
            if (width > 0) std::memcpy(parents.data() + o * width, cand->parents + r0 * width, (size_t)len * width * sizeof(int32_t));
            o += len;
        }
        abn_candidates sub{mine, width, child.data(), parents.data()};
        double *dst[4] = {out->mlik, out->aic, out->bic, out->mdl};
        abn_result res{};
        for (int f = 0; f < 4; f++)
            if (dst[f]) { val[f].assign(mine, NAN); (&res.mlik)[f] = val[f].data(); }
        res.status = status.data();
        res.iters = iters.data();
        abn_control lc = *ctrl;
        lc.ngpus = 0;
        lc.device = ctrl->devices[rank];
        rcs[rank] = (mine > 0) ? score_cache_on(ds, &sub, &lc, lc.device, &res, lerr, sizeof(lerr)) : ABN_OK;
        msgs[rank] = lerr;
        if (rcs[rank] != ABN_OK) return;
        o = 0;
        for (long long b = rank; b < nblocks; b += world) { // scatter back into cache order
            const long long r0 = b * block, len = std::min<long long>(block, rows - r0);
            for (int f = 0; f < 4; f++)
                if (dst[f]) std::memcpy(dst[f] + r0, val[f].data() + o, (size_t)len * sizeof(double));
            if (out->status) std::memcpy(out->status + r0, status.data() + o, (size_t)len * sizeof(int32_t));
            if (out->iters) std::memcpy(out->iters + r0, iters.data() + o, (size_t)len * sizeof(int32_t));
            o += len;
        }
    };
    std::vector<std::thread> threads;
    for (int r = 1; r < world; r++) threads.emplace_back(work, r);
    work(0);
    for (auto &t : threads) t.join();
    for (int r = 0; r < world; r++)
        if (rcs[r] != ABN_OK) return fail(err, errlen, rcs[r], "device %d: %s", ctrl->devices[r], msgs[r].c_str());
    return ABN_OK;
}

extern "C" int abn_b200_last_timing(const abn_data *d, abn_timing *t)
{
    if (!d || !t) return ABN_ERR_ARG;
    *t = d->timing;
    return ABN_OK;
}

extern "C" int32_t abn_b200_version(void) { return ABN_B200_VERSION; }

extern "C" int32_t abn_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void abn_b200_default_control(int32_t method, abn_control *c)
{
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->method = method;
    c->max_iters = 100;            // R/build_score_cache.R:97
    c->tol = 1e-8;                 // :113
    c->epsabs = 1e-7;              // :98
    c->prior_mean = 0.0;           // :93
    c->prior_sd = 1.0 / sqrt(0.001); // 1/sqrt(prec), :94 and _bayes.R:314
    c->gam_shape = 1.0;            // :95
    c->gam_scale = 1.0 / 5e-05;    // 1/loggam.inv.scale, :96 and _bayes.R:315
    c->solver = ABN_SOLVER_HYBRID;
    c->device = 0;
    c->nvars = 0;
    c->ngpus = 0;                  // one device: `device`
    c->block = 1024;
    for (int i = 0; i < ABN_MAX_DEVICES; i++) c->devices[i] = i;
}
