// fit_cell.cuh - sufficient-statistic collapse for all-discrete designs: one warp per fit.
//
// When every design column of a binomial or Poisson child is discrete (integer codes 0..255: binomial
// nodes, the one-hot columns of multinomial nodes, small counts), the linear predictor takes as many
// values as there are distinct parent patterns ("cells"), and every sum a data pass produces,
//     H = sum_i w(eta_i) x_i x_i',  g = sum_i r(eta_i, y_i) x_i,  ll = sum_i l(eta_i, y_i),
// is the same sum over cells weighted by the cell counts. This is to discrete designs what the global
// Gram matrix is to Gaussian children (plan.cuh): the warp builds the contingency table of the fit ONCE -
// an integer pass over the byte-coded columns (DataMeta::Xq, 1 byte per value instead of 8, ~4 integer
// instructions per row instead of ~40 mixed ones per row and pass) - and then runs the unchanged solver
// state machine (solver.cuh) with passes over <= 4096 weighted cells instead of n rows.
//
//   binomial child (either method): cells over (y, parent pattern). Designs with an intercept use the lean row
//       op and the Gram identities of lean_pass_bin (the solver's shortcuts apply); a fit that meets an eta
//       outside that row op's range switches to the careful one - the reference's arithmetic per distinct
//       (y, x), rho correction always on - for the rest of its passes
//   Poisson child: cells over the parent pattern only; X'y comes from the global Gram matrix
//       (g = X'y - sum N E x, ll = beta'X'y - sum N E, see warp_pass_pois)
//
// Table: the codes of the columns are packed into the cell index bit field by bit field (ceil(log2 card)
// bits each, <= 10 bits in all); the warp's shared-memory table holds `copies` interleaved replicas of every
// counter (copy = lane mod copies, 4096 / cells of them, at most 32): see cell_count_rows.
// Counting is exact integer work, so the table - and with it every result - is independent of how the rows
// are split over lanes: bit-identical under any partition of the candidate list.
#pragma once

#include <cuda_runtime.h>

#include <type_traits>

#include "fit_kernel.cuh"

namespace ABN_NS {

// Two table classes, two kernel instantiations, both at 16 warps per SM (the whole kernel is latency-bound: dependent
// shared-memory updates, a serial solver on lane 0 - warps are what hides it)
constexpr int CELL_SMALL_CAPBITS = 11;    // cell index <= 6 bits: 2 048 counters per warp (8 KB), 16 warps per SM
#ifndef ABN_CELL_LARGE_CAPBITS
#define ABN_CELL_LARGE_CAPBITS 10
#endif
// cell index 7..10 bits. 12: 4 096 counters per warp (16 KB, 4..32 replicas per counter), 8 warps per SM.
// 10: 1 024 counters (4 KB), at most 8 replicas - every update is a shared-memory atomic - and 16 warps per SM
constexpr int CELL_LARGE_CAPBITS = ABN_CELL_LARGE_CAPBITS;
constexpr int CELL_LARGE_PER_SM = CELL_LARGE_CAPBITS <= 10 ? 4 : 2; // CTAs per SM the large-table kernel is built for
constexpr int CELL_SMALL_BITS = 6;
// all-binary designs (every byte column has codes 0 / 1, bit planes built): the table comes out of registers
// (cell_count_planes), shared memory holds just the 64 cells, and the kernel is built for 24 warps per SM - the
// solver on lane 0 is latency-bound and warps are what hides it (-18 % on config 3 against 16 warps per SM)
constexpr int CELL_PLANES_CAPBITS = 7;
constexpr int CELL_PLANES_PER_SM = 6;
constexpr int CELL_MAX_BITS = 10;         // widest cell index taken
constexpr int CELL_WARPS = 4;             // warps per CTA
#ifndef ABN_CELL_ATOMIC_PHASES
#define ABN_CELL_ATOMIC_PHASES 4
#endif
constexpr int CELL_ATOMIC_PHASES = ABN_CELL_ATOMIC_PHASES; // lanes per counter replica from which the counters are updated atomically

// Bits the codes of a column get in the cell index. A count column whose codes would need more than 3 bits is
// CAPPED where the data allow it: codes below 2^b are binned, the rows at or beyond it - at most 1/32 of the
// data per column, known from the column's tail counts - are kept as an explicit overflow list and summed row by
// row in every pass. A Poisson parent with counts up to 40 then costs 3..4 bits instead of 6, which brings
// designs with two or three count parents into the table sizes this kernel is good at.
ABN_HD int cell_bits_of(const DataMeta &m, int col, bool *capped)
{
    const int card = m.col_card[col];
    int full = 1;
    while ((1 << full) < card) full++;
    *capped = false;
    if (full <= 3 || m.col_tail == nullptr || m.cell_ovf == nullptr) return full;
    int b = 3;
    while (b < full && (long long)m.col_tail[col * 8 + b] * 32 > m.n) b++;
    *capped = b < full;
    return b;
}

// Is this resolved design one the cell kernel takes? 0 = no, 1 = small table, 2 = large table, 3 = bit planes (called by
// resolve_kernel; host builds have no byte codes and never say yes)
ABN_HD int cell_eligible(const DataMeta &m, const FitDesc &d, int method, int flags)
{
    if (m.col_card == nullptr || d.err || d.k < 1) return 0;
    if (d.family != BINOMIAL && d.family != POISSON) return 0;
    if (d.family == POISSON && method == MLE && (flags & 1)) return 0; // the reference's start-value pass needs y per row
    int bits = (d.family == BINOMIAL) ? 1 : 0;                             // the response is part of a binomial cell
    int ncapped = 0;
    bool binary = true; // every design column has codes 0 / 1 (the response of a binomial child always has)
    for (int j = 0; j < d.k; j++) {
        if (d.cols[j] < 0) continue;
        if (m.col_card[d.cols[j]] <= 0) return 0;
        binary &= m.col_card[d.cols[j]] <= 2;
        bool capped;
        bits += cell_bits_of(m, d.cols[j], &capped);
        ncapped += capped;
    }
    if (d.family == BINOMIAL && m.col_card[d.ycol] != 2) return 0;
    if (ncapped > 4) return 0; // the overflow list holds n / 8 rows: 4 columns x n / 32
    if (binary && m.Xb != nullptr && bits >= 1 && bits <= CELL_SMALL_BITS) return 3;
    return bits <= CELL_SMALL_BITS ? 1 : (bits <= CELL_MAX_BITS ? 2 : 0);
}

struct CellPlan {
    int ncol;            // byte columns that enter the cell index (binomial: the response first)
    int col[MAXK + 1];   // stored column of each
    int shift[MAXK + 1]; // bit position of its code in the cell index
    int bits[MAXK + 1];
    int total_bits;
    int xslot[MAXK];     // design column j takes its value from byte column xslot[j] (-1: intercept)
    int cap[MAXK + 1];   // codes >= cap go to the overflow list (256: the column is not capped)
    int any_cap;
    int all_binary;      // every byte column is binary and the bit planes exist: cell_count_planes
    unsigned n_ovf;      // rows in the overflow list of the fit in flight
};

// Contingency table of the fit. Counter of cell c: tab[(c << lc) | (set << lcl) | (lane mod copies)], copies = 2^lcl
// replicas across lanes and 2^(lc - lcl) "sets" across consecutive rows of a lane.
// Shared-memory atomics cost ~2 cycles per LANE on this part (the load-store unit serialises them), so the
// counters are plain read-modify-writes instead:
//   * with 32 lane replicas every lane owns its counters outright; with fewer, the lanes that share a replica take
//     turns (32 / copies phases separated by __syncwarp, each phase conflict-free by construction);
//   * a read-modify-write is a ~35-cycle dependent chain (LDS, add, STS) and two rows of a lane may meet the same
//     counter, so the chain cannot be overlapped - unless consecutive rows go to different sets: with 4 sets the
//     4 counters of a row group are distinct by construction and their loads, adds and stores are issued together.
// Rows with a code at or beyond its column's cap are not counted: their indices go to `ovf` (ascending row order
// within a lane step, lane by lane - a fixed order) and `*n_ovf` returns how many.
template <int NC>
__device__ __forceinline__ void cell_count_rows(const DataMeta &m, const CellPlan &cp, unsigned *tab, int cap, int lcl, int lc,
                                                unsigned *ovf, unsigned *n_ovf)
{
    const int lane = threadIdx.x & 31;
    const unsigned char *colp[NC];
    int sh[NC];
    unsigned capv[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) { colp[j] = m.Xq + (long long)cp.col[j] * m.ldq; sh[j] = cp.shift[j]; capv[j] = (unsigned)cp.cap[j]; }
    const bool any_cap = cp.any_cap != 0;
    unsigned novf = 0;
    const unsigned lane_lt = (1u << lane) - 1u;
    unsigned *const mine = tab + (lane & ((1 << lcl) - 1));
    const int nphase = 32 >> lcl, my_phase = lane >> lcl;
    const int nsets = 1 << (lc - lcl); // 1, 2 or 4 (only with 32 lane replicas)
    unsigned *const dead = tab + cap + lane; // scratch counter of a lane without a row
    const long long n16 = m.n / 16;
    const long long steps = (n16 + 31) / 32; // every lane runs the same number of steps (the phases synchronise)
    // 16 rows per lane and step: one 16-byte load per column; the codes of 4 rows are combined at once when
    // the whole index fits a byte (<= 256 cells), otherwise row by row
    const bool bytewise = cp.total_bits <= 8 && !any_cap;
    // the loads of step st + 1 are issued before the counters of step st are updated: the ~600 cycles of an L2 hit
    // overlap the dependent shared-memory updates instead of preceding them
    uint4 vn[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) vn[j] = (lane < n16) ? __ldg(reinterpret_cast<const uint4 *>(colp[j]) + lane) : make_uint4(0u, 0u, 0u, 0u);
    for (long long st = 0; st < steps; st++) {
        const long long c = st * 32 + lane;
        const bool live = c < n16;
        uint4 v[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) v[j] = vn[j];
        {
            const long long cn = c + 32;
#pragma unroll
            for (int j = 0; j < NC; j++) vn[j] = (cn < n16) ? __ldg(reinterpret_cast<const uint4 *>(colp[j]) + cn) : make_uint4(0u, 0u, 0u, 0u);
        }
        unsigned cells[16];
        if (bytewise) {
            unsigned w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < NC; j++) {
                w[0] |= v[j].x << sh[j]; w[1] |= v[j].y << sh[j]; w[2] |= v[j].z << sh[j]; w[3] |= v[j].w << sh[j];
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int b = 0; b < 4; b++) cells[4 * q + b] = (w[q] >> (8 * b)) & 0xffu;
        } else {
            unsigned over = 0; // bit r: row r of this lane's 16 goes to the overflow list
#pragma unroll
            for (int q = 0; q < 4; q++) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    unsigned cell = 0, o = 0;
#pragma unroll
                    for (int j = 0; j < NC; j++) {
                        const unsigned word = q == 0 ? v[j].x : (q == 1 ? v[j].y : (q == 2 ? v[j].z : v[j].w));
                        const unsigned code = (word >> (8 * b)) & 0xffu;
                        o |= (code >= capv[j]) ? 1u : 0u;
                        cell |= code << sh[j];
                    }
                    cells[4 * q + b] = cell;
                    over |= o << (4 * q + b);
                }
            }
            if (any_cap) {
                // a row beyond a cap is sent to the scratch counter and appended to the list: per row of the step one
                // ballot gives every lane its slot (lanes in order, rows in order)
                if (!live) over = 0;
#pragma unroll
                for (int r = 0; r < 16; r++) {
                    const bool mine_over = (over >> r) & 1u;
                    const unsigned bal = __ballot_sync(0xffffffffu, mine_over);
                    if (mine_over) {
                        ovf[novf + __popc(bal & lane_lt)] = (unsigned)(c * 16 + r);
                        cells[r] = 0xffffffffu;
                    }
                    novf += __popc(bal);
                }
            }
        }
        if (nphase >= CELL_ATOMIC_PHASES) {
            // few replicas (wide tables): shared-memory atomics - ~2 cycles per lane, 64 per warp instruction - beat
            // nphase turns of a ~35-cycle read-modify-write from 4 phases on; integer adds, same table
            if (live) {
#pragma unroll
                for (int r = 0; r < 16; r++)
                    if (cells[r] != 0xffffffffu) atomicAdd(mine + (cells[r] << lc), 1u);
            }
        } else if (any_cap) {
            // counted rows only (an overflow row carries the marker); same phases as below
            for (int ph = 0; ph < nphase; ph++) {
                if (my_phase == ph && live) {
#pragma unroll
                    for (int r = 0; r < 16; r++)
                        if (cells[r] != 0xffffffffu) { unsigned *p = mine + (cells[r] << lc); *p = *p + 1u; }
                }
                if (nphase > 1) __syncwarp();
            }
        } else if (nphase == 1) {
            if (live) {
                if (nsets == 4) {
#pragma unroll
                    for (int g = 0; g < 16; g += 4) {
                        unsigned *p[4], cnt[4];
#pragma unroll
                        for (int t = 0; t < 4; t++) { p[t] = mine + ((cells[g + t] << lc) | (unsigned)(t << 5)); cnt[t] = *p[t]; }
#pragma unroll
                        for (int t = 0; t < 4; t++) *p[t] = cnt[t] + 1u;
                    }
                } else if (nsets == 2) {
#pragma unroll
                    for (int g = 0; g < 16; g += 2) {
                        unsigned *p0 = mine + (cells[g] << lc), *p1 = mine + ((cells[g + 1] << lc) | 32u);
                        const unsigned c0 = *p0, c1 = *p1;
                        *p0 = c0 + 1u;
                        *p1 = c1 + 1u;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 16; r++) { unsigned *p = mine + (cells[r] << lc); *p = *p + 1u; }
                }
            }
        } else {
            for (int ph = 0; ph < nphase; ph++) {
                if (my_phase == ph && live) {
#pragma unroll
                    for (int r = 0; r < 16; r++) { unsigned *p = mine + (cells[r] << lc); *p = *p + 1u; }
                }
                __syncwarp();
            }
        }
    }
    for (long long i0 = n16 * 16; i0 < m.n; i0 += 32) { // fewer than 16 rows left
        const long long i = i0 + lane;
        const bool live = i < m.n;
        unsigned cell = 0;
        bool o = false;
        if (live) {
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const unsigned code = __ldg(colp[j] + i);
                o |= code >= capv[j];
                cell |= code << sh[j];
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, o);
        if (o) ovf[novf + __popc(bal & lane_lt)] = (unsigned)i;
        novf += __popc(bal);
        unsigned *p = (live && !o) ? mine + (cell << lc) : dead;
        if (nphase >= CELL_ATOMIC_PHASES) {
            atomicAdd(p, 1u);
        } else {
            for (int ph = 0; ph < nphase; ph++) {
                if (my_phase == ph) *p = *p + 1u;
                __syncwarp();
            }
        }
    }
    *n_ovf = novf;
}

// Contingency table of a fit whose byte columns are all BINARY, from the bit planes (DataMeta::Xb): the number of
// rows on which every plane of a subset S is one is popcount(AND of the planes of S), and the table - rows on which
// EXACTLY the planes of T are one - follows by inclusion-exclusion. Subsets of fewer than three planes need no data
// at all (n, column sums and pair counts are entries of the global Gram matrix, exact integers in FP64), so a
// binomial child with four binary parents costs 16 AND + 16 POPC + 16 adds per 32 ROWS and lane, in registers, with
// no dependent shared-memory update anywhere (cell_count_rows: ~8 instructions and a ~35-cycle chain per row).
// All integer work: the table is the one cell_count_rows builds, bit for bit.
__host__ __device__ constexpr int cell_popc(int x) { return x == 0 ? 0 : (x & 1) + cell_popc(x >> 1); }
__host__ __device__ constexpr int cell_ctz(int x) { return (x & 1) ? 0 : 1 + cell_ctz(x >> 1); }

// f(integral_constant<I>) for I = I0 .. N-1: the subset loops below index register arrays, so every index has to be
// a constant expression (a "#pragma unroll" over 63 subsets x 4 words is not honoured and the arrays go to local memory)
template <int I, int N, typename F>
__device__ __forceinline__ void cell_static_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        cell_static_for<I + 1, N>(f);
    }
}

template <int NP>
__device__ __forceinline__ void cell_count_planes(const DataMeta &m, const CellPlan &cp, unsigned *tab)
{
    constexpr int NS = 1 << NP;
    const int lane = threadIdx.x & 31;
    const uint4 *pp[NP];
#pragma unroll
    for (int j = 0; j < NP; j++) pp[j] = reinterpret_cast<const uint4 *>(m.Xb + (long long)cp.col[j] * m.ldb);
    unsigned cnt[NS];
#pragma unroll
    for (int S = 0; S < NS; S++) cnt[S] = 0u;
    const long long nq = m.ldb / 4; // 128 rows per 16-byte load; words beyond n are zero
    uint4 vn[NP]; // loads of the next step in flight while this one is counted
#pragma unroll
    for (int j = 0; j < NP; j++) vn[j] = (lane < nq) ? __ldg(pp[j] + lane) : make_uint4(0u, 0u, 0u, 0u);
    for (long long q = lane; q < nq; q += 32) {
        uint4 v[NP];
#pragma unroll
        for (int j = 0; j < NP; j++) v[j] = vn[j];
#pragma unroll
        for (int j = 0; j < NP; j++) vn[j] = (q + 32 < nq) ? __ldg(pp[j] + q + 32) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int w = 0; w < 4; w++) {
            unsigned pl[NP], msk[NS];
#pragma unroll
            for (int j = 0; j < NP; j++) pl[j] = w == 0 ? v[j].x : (w == 1 ? v[j].y : (w == 2 ? v[j].z : v[j].w));
            // subsets in increasing order: the mask of S is the mask of (S without its lowest plane) AND that plane
            cell_static_for<1, NS>([&](auto Sc) {
                constexpr int S = decltype(Sc)::value;
                constexpr int j = cell_ctz(S), rest = S & (S - 1);
                if constexpr (rest == 0) msk[S] = pl[j];
                else msk[S] = msk[rest] & pl[j];
                if constexpr (cell_popc(S) >= 3) cnt[S] += __popc(msk[S]);
            });
        }
    }
    cell_static_for<1, NS>([&](auto Sc) {
        constexpr int S = decltype(Sc)::value;
        if constexpr (cell_popc(S) >= 3) {
            unsigned c = cnt[S];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) tab[S] = c;
        }
    });
    for (int S = lane; S < NS; S += 32) {
        const int pc = __popc(S);
        if (pc == 0) tab[S] = (unsigned)m.n;
        else if (pc <= 2) {
            const int a = gram_col(cp.col[__ffs(S) - 1]), b = pc == 1 ? 0 : gram_col(cp.col[31 - __clz(S)]);
            tab[S] = (unsigned)m.G[(long long)a * m.D + b];
        }
    }
    __syncwarp();
    // f(S) = rows with all planes of S set  ->  g(T) = rows with exactly the planes of T set: plane by plane
#pragma unroll
    for (int i = 0; i < NP; i++) {
        for (int S = lane; S < NS; S += 32)
            if (!((S >> i) & 1)) tab[S] -= tab[S | (1 << i)];
        __syncwarp();
    }
}

// One "data pass" over the cells (weight = count) and the overflow rows (weight 1). acc gets the sums a row pass
// would produce. K = design columns (compile time: the accumulators live in registers).
//   LEAN (binomial, designs with an intercept): the row op and the Gram identities of lean_pass_bin - sigma, w and
//     log(1 + e) do not depend on the response, X'y and X'1 come from the Gram matrix - so the solver's shortcuts
//     apply (has_gl) and the fit needs ~2 passes; returns false when some eta leaves the range the fast row op is
//     exact for (eta >= 12 for mle: the reference's rho noise; |eta| >= 700 for bayes), and the careful variant
//     redoes the pass: the reference's arithmetic per distinct (y, x) with the rho correction always on.
//   Poisson: H = sum N E x x', score and log-likelihood from X'y as in warp_pass_pois.
template <int KIND, int K, bool LEAN>
__device__ __forceinline__ bool cell_pass(const DataMeta &m, const CellPlan &cp, const unsigned *cnt, const unsigned *ovf,
                                          const double *beta_sm, const double *xty, const double *xt1, int mode,
                                          const RowConst rc, PassAcc *acc_sm)
{
    constexpr int NH = K * (K + 1) / 2;
    constexpr bool MLE_ = (KIND == RK_BIN_MLE);
    const int lane = threadIdx.x & 31;
    const bool need_h = (mode == PASS_FULL || mode == PASS_HG || mode == PASS_HESS);
    const bool need_g = (mode == PASS_FULL || mode == PASS_HG || mode == PASS_GRAD);
    const bool need_l = (mode == PASS_FULL || mode == PASS_LL);
    double beta[K], H[NH], g[K], ll = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) { beta[j] = beta_sm[j]; g[j] = 0.0; }
#pragma unroll
    for (int j = 0; j < NH; j++) H[j] = 0.0;
    RowTrack tk = row_track_init();
    const int ncell = 1 << cp.total_bits;
    const int nunit = ncell + (int)cp.n_ovf;
    for (int u = lane; u < nunit; u += 32) {
        double dn, x[K], y = 0.0;
        if (u < ncell) {
            const unsigned N = cnt[u];
            if (N == 0) continue;
            dn = (double)N;
#pragma unroll
            for (int j = 0; j < K; j++) {
                const int sl = cp.xslot[j];
                x[j] = sl < 0 ? 1.0 : (double)((u >> cp.shift[sl]) & ((1 << cp.bits[sl]) - 1));
            }
            y = (double)(u & 1); // the response is the first bit field of a binomial cell
        } else {
            const long long row = ovf[u - ncell];
            dn = 1.0;
#pragma unroll
            for (int j = 0; j < K; j++) {
                const int sl = cp.xslot[j];
                x[j] = sl < 0 ? 1.0 : (double)m.Xq[(long long)cp.col[sl] * m.ldq + row];
            }
            if (KIND != RK_POIS) y = (double)m.Xq[(long long)cp.col[0] * m.ldq + row];
        }
        double eta = 0.0;
#pragma unroll
        for (int j = 0; j < K; j++) eta = fma(x[j], beta[j], eta);
        double w = 0.0, r = 0.0, l = 0.0;
        if (KIND == RK_POIS) {
            row_track(tk, eta);
            const double E = fast_exp(clamp_lo700(eta), rc.tab);
            w = E; r = E; l = E; // r, l: the parts that are summed per cell; X'y is added below
        } else if (LEAN) {
            row_track(tk, eta);
            const double e = fast_exp(-clamp700(eta), rc.tab);
            const double d = 1.0 + e;
            const double gg = fast_rcp_approx(d);
            w = fma(-gg, gg, gg);
            r = gg;                                  // S = sum g x
            if (need_l) l = fast_log_ge1(d, rc.logtab); // L = sum log(1 + e)
        } else {
            RowOp<KIND>::template apply<true, true, true>(eta, y, rc, w, r, l, tk, true);
        }
        w *= dn; r *= dn; l *= dn;
        if (need_h) {
            int t = 0;
#pragma unroll
            for (int p = 0; p < K; p++) {
                const double wx = w * x[p];
#pragma unroll
                for (int q = p; q < K; q++) { H[t] = fma(wx, x[q], H[t]); t++; }
            }
        }
        if (need_g) {
#pragma unroll
            for (int p = 0; p < K; p++) g[p] = fma(r, x[p], g[p]);
        }
        if (need_l) ll += l;
    }
    if (LEAN) {
        const int out_of_range = MLE_ ? ((tk.smax >= 0x40280000) | (tk.umax >= 0xC0762E42u))
                                      : ((tk.smax >= 0x4085E000) | (tk.umax >= 0xC085E000u));
        if (__any_sync(0xffffffffu, out_of_range)) return false;
    }
    if (need_h) {
#pragma unroll
        for (int j = 0; j < NH; j++) H[j] = warp_sum(H[j]);
    }
    if (need_g) {
#pragma unroll
        for (int j = 0; j < K; j++) g[j] = warp_sum(g[j]);
    }
    if (need_l) ll = warp_sum(ll);
    const int bad = LEAN ? 0 : __any_sync(0xffffffffu, RowOp<KIND>::bad_of(tk));
    if (lane == 0) {
        if (need_h) {
#pragma unroll
            for (int j = 0; j < NH; j++) acc_sm->H[j] = H[j];
        }
        if (need_g) {
#pragma unroll
            for (int j = 0; j < K; j++) {
                if (KIND == RK_POIS) acc_sm->g[j] = xty[j] - g[j];
                else if (LEAN) {
                    const double dbp = rc.bp1 - rc.bp0;
                    acc_sm->g[j] = MLE_ ? fma(dbp, xty[j], rc.bp0 * xt1[j]) - g[j] : xty[j] - g[j];
                    if (MLE_) acc_sm->gl[j] = xty[j] - g[j];
                } else acc_sm->g[j] = g[j];
            }
        }
        double lsum = 0.0;
        if (need_l) {
            if (KIND == RK_POIS) {
                lsum = -ll;
#pragma unroll
                for (int j = 0; j < K; j++) lsum = fma(beta[j], xty[j], lsum);
            } else if (LEAN) {
                lsum = -ll;
#pragma unroll
                for (int j = 0; j < K; j++) lsum = fma(beta[j], xty[j] - xt1[j], lsum);
            } else {
                lsum = ll;
            }
        }
        acc_sm->ll = lsum;
        acc_sm->bad = bad;
        acc_sm->has_gl = (LEAN && MLE_ && need_g) ? 1 : 0;
    }
    return true;
}

// run-time design width -> compile-time K; binomial designs with an intercept try the lean variant first
template <int KIND>
__device__ __forceinline__ void cell_pass_any(const DataMeta &m, const CellPlan &cp, const unsigned *cnt, const unsigned *ovf,
                                              int k, bool icpt, const double *beta_sm, const double *xty, const double *xt1,
                                              int mode, const RowConst rc, PassAcc *acc_sm, int &careful)
{
#define ABN_CELL_CASE(KK)                                                                                              \
    case KK:                                                                                                           \
        if (KIND != RK_POIS && icpt && !careful) {                                                                    \
            if (cell_pass<KIND, KK, true>(m, cp, cnt, ovf, beta_sm, xty, xt1, mode, rc, acc_sm)) return;              \
            careful = 1;                                                                                               \
        }                                                                                                              \
        cell_pass<KIND, KK, false>(m, cp, cnt, ovf, beta_sm, xty, xt1, mode, rc, acc_sm);                              \
        return;
    switch (k) {
        ABN_CELL_CASE(1) ABN_CELL_CASE(2) ABN_CELL_CASE(3) ABN_CELL_CASE(4)
        ABN_CELL_CASE(5) ABN_CELL_CASE(6) ABN_CELL_CASE(7) ABN_CELL_CASE(8)
    }
#undef ABN_CELL_CASE
}

template <int KIND, int CAPBITS>
__global__ void __launch_bounds__(CELL_WARPS * 32, CAPBITS == CELL_PLANES_CAPBITS ? CELL_PLANES_PER_SM : (CAPBITS == CELL_SMALL_CAPBITS ? 4 : CELL_LARGE_PER_SM))
fit_cell_kernel(const DataMeta m, const Control c, const FitDesc *__restrict__ descs, const int *__restrict__ child,
                const int *__restrict__ list, int count, int *counter, ResultPtrs out, int ovf_slot0)
{
    __shared__ FitState st_all[CELL_WARPS];
    __shared__ PassAcc acc_all[CELL_WARPS];
    __shared__ CellPlan plan_all[CELL_WARPS];
    __shared__ double xty_all[CELL_WARPS][2 * MAXK]; // (X'y)_j and (X'1)_j of the fit in flight
    __shared__ double exp_tab[ABN_EXP2_N];
    __shared__ double log_tab[128];
    constexpr int CAP = 1 << CAPBITS;
    extern __shared__ __align__(16) unsigned cell_tabs[]; // CELL_WARPS x (CAP counters + 32 scratch)
    if (threadIdx.x < 64) fill_log_table(log_tab, threadIdx.x);
    if (threadIdx.x < ABN_EXP2_N) exp_tab[threadIdx.x] = c_exp2_table[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FitState &st = st_all[warp];
    PassAcc *acc = &acc_all[warp];
    CellPlan &cp = plan_all[warp];
    double *xty = xty_all[warp];
    unsigned *tab = cell_tabs + (size_t)warp * (CAP + 32);
    // this warp's overflow list (row indices beyond a cap); `ovf_slot0` separates the kernels that may run together
    unsigned *ovf = (m.cell_ovf && CAPBITS != CELL_PLANES_CAPBITS) // binary columns are never capped: the bit-plane kernel owns no list
                        ? m.cell_ovf + ((long long)ovf_slot0 + (long long)blockIdx.x * CELL_WARPS + warp) * m.cell_ovf_stride : nullptr;

    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(counter, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= count) break;
        const int row = list[idx];
        const FitDesc d = descs[row];
        const int ch = child[row];
        const int k = d.k;
        if (lane == 0) {
            int nc = 0, sh = 0;
            if (d.family == BINOMIAL) { cp.col[0] = d.ycol; cp.shift[0] = 0; cp.bits[0] = 1; nc = 1; sh = 1; }
            cp.cap[0] = 256;
            cp.any_cap = 0;
            int binary = m.Xb != nullptr;
            for (int j = 0; j < k; j++) {
                if (d.cols[j] < 0) { cp.xslot[j] = -1; continue; }
                bool capped;
                const int b = cell_bits_of(m, d.cols[j], &capped);
                cp.col[nc] = d.cols[j]; cp.shift[nc] = sh; cp.bits[nc] = b;
                cp.cap[nc] = capped ? (1 << b) : 256;
                cp.any_cap |= capped ? 1 : 0;
                binary &= m.col_card[d.cols[j]] <= 2;
                cp.xslot[j] = nc;
                nc++; sh += b;
            }
            cp.ncol = nc;
            cp.all_binary = binary && nc <= CELL_SMALL_BITS;
            cp.total_bits = sh;
            cp.n_ovf = 0;
            const long long gy = gram_col(d.ycol);
            for (int j = 0; j < k; j++) {
                xty[j] = m.G[(long long)gidx(d, j) * m.D + gy];
                xty[MAXK + j] = m.G[(long long)gidx(d, j) * m.D];
            }
        }
        __syncwarp();
        const int tb = cp.total_bits;
        if ((CAPBITS == CELL_PLANES_CAPBITS || CAPBITS == CELL_SMALL_CAPBITS) && cp.all_binary && cp.ncol >= 1) {
            switch (cp.ncol) {
            case 1: cell_count_planes<1>(m, cp, tab); break;
            case 2: cell_count_planes<2>(m, cp, tab); break;
            case 3: cell_count_planes<3>(m, cp, tab); break;
            case 4: cell_count_planes<4>(m, cp, tab); break;
            case 5: cell_count_planes<5>(m, cp, tab); break;
            default: cell_count_planes<6>(m, cp, tab); break;
            }
        } else {
            const int lcl = min(5, CAPBITS - tb);       // log2 of the replicas across lanes
            const int lc = min(7, CAPBITS - tb);        // ... times the sets across rows (cell_count_rows)
            for (int i = lane; i < (1 << (tb + lc)); i += 32) tab[i] = 0u;
            __syncwarp();
            unsigned novf = 0;
            switch (cp.ncol) {
            case 0: break;
            case 1: cell_count_rows<1>(m, cp, tab, CAP, lcl, lc, ovf, &novf); break;
            case 2: cell_count_rows<2>(m, cp, tab, CAP, lcl, lc, ovf, &novf); break;
            case 3: cell_count_rows<3>(m, cp, tab, CAP, lcl, lc, ovf, &novf); break;
            case 4: cell_count_rows<4>(m, cp, tab, CAP, lcl, lc, ovf, &novf); break;
            case 5: cell_count_rows<5>(m, cp, tab, CAP, lcl, lc, ovf, &novf); break;
            case 6: cell_count_rows<6>(m, cp, tab, CAP, lcl, lc, ovf, &novf); break;
            case 7: cell_count_rows<7>(m, cp, tab, CAP, lcl, lc, ovf, &novf); break;
            case 8: cell_count_rows<8>(m, cp, tab, CAP, lcl, lc, ovf, &novf); break;
            default: cell_count_rows<MAXK + 1>(m, cp, tab, CAP, lcl, lc, ovf, &novf); break;
            }
            __syncwarp();
            if (lane == 0) cp.n_ovf = novf;
            // fold the replicas: counter of cell c ends up at tab[c]. Round r handles cells 32 r .. 32 r + 31: all
            // reads of a round precede its writes, and later rounds read only beyond what has been written
            if (cp.ncol == 0) {
                if (lane == 0) tab[0] = (unsigned)m.n; // intercept-only Poisson fit: one cell
            } else if (lc > 0) {
                const int copies = 1 << lc;
                for (int c0 = 0; c0 < (1 << tb); c0 += 32) {
                    const int cc = c0 + lane;
                    unsigned s = 0;
                    if (cc < (1 << tb))
                        for (int r = 0; r < copies; r++) s += tab[(cc << lc) + r];
                    __syncwarp();
                    if (cc < (1 << tb)) tab[cc] = s;
                    __syncwarp();
                }
            }
            __syncwarp();

        }
        __syncwarp();

        const RowConst rc = make_row_const((c.flags & 2) ? 0 : k, (double)m.n, exp_tab, log_tab, nullptr);
        int mode = 0;
        if (lane == 0) {
            state_setup(m, d, ch, c, st);
            if (c.method == BAYES) bayes_inits(m, d, ch, st);
            mode = fit_begin(st, c);
            if (KIND == RK_BIN_MLE) {
                binomial_first_acc(m, d, *acc, !(c.flags & 2));
                st.have_ll = 1;
                mode = fit_consume(st, c, *acc);
                st.passes--;
            }
            if (KIND == RK_POIS && mode == PASS_PINIT) {
                poisson_init_acc(m, d, ch, *acc);
                mode = fit_consume(st, c, *acc);
                st.passes--;
            }
        }
        mode = __shfl_sync(0xffffffffu, mode, 0);
        __syncwarp();
        int careful = 0;
        const bool icpt = d.cols[0] < 0;
        while (mode != PASS_DONE) {
            cell_pass_any<KIND>(m, cp, tab, ovf, k, icpt, st.x, xty, xty + MAXK, mode, rc, acc, careful);
            __syncwarp();
            if (lane == 0) mode = fit_consume(st, c, *acc);
            mode = __shfl_sync(0xffffffffu, mode, 0);
            __syncwarp();
        }
        if (lane == 0) {
            if (out.mlik) out.mlik[row] = st.loglik;
            if (c.method == MLE) {
                if (out.aic) out.aic[row] = st.aic;
                if (out.bic) out.bic[row] = st.bic;
                if (out.mdl) out.mdl[row] = mdl_of(st.bic, d, c.nvars);
            }
            if (out.status) out.status[row] = st.status;
            if (out.iters) out.iters[row] = st.passes;
            write_coef(out, row, d, st.x, st.M);
            write_se(out, row, d, st, c.method);
            if (out.sse && c.method == MLE) out.sse[row] = sse_of(m, d, st.x);
        }
        __syncwarp();
    }
}

} // namespace ABN_NS
