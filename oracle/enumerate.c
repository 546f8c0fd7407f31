/* TEST INFRASTRUCTURE - part of the CPU oracle (see abn_oracle.h). Candidate enumeration. */
#include <stdlib.h>
#include <string.h>

#include "abn_oracle.h"

/* gsl_combination_next: advance c[0..k) to the next k-subset of {0..n-1} in lexicographic
 * order; returns 0 when there is none. */
static int comb_next(int *c, int n, int k)
{
    if (k == 0) return 0;
    int i = k - 1;
    while (i > 0 && c[i] == n - k + i) i--;
    if (i == 0 && c[i] == n - k) return 0;
    c[i]++;
    for (; i < k - 1; i++) c[i + 1] = c[i] + 1;
    return 1;
}

/* src/buildcachematrix.c:150-162 */
static int isretained(int node, const int *row, const int *retained, int d)
{
    for (int j = 0; j < d; j++)
        if (retained[node + j * d] == 1 && row[j] != 1) return 0;
    return 1;
}

int64_t or_enumerate_bayes(int d, const int *banned, const int *retained, const int *max_parents,
                           const int *which_nodes, int n_which, int *children, int *node_defn)
{
    /* src/buildcachematrix.c:41-74 (count) and :94-130 (fill): identical loops */
    int64_t total = 0;
    int *c = (int *)malloc(sizeof(int) * (d + 1));
    int *row = (int *)malloc(sizeof(int) * d);
    for (int w = 0; w < n_which; w++) {
        int node = which_nodes[w];
        for (int j = 0; j <= max_parents[node]; j++) {
            if (j > d) break;
            for (int t = 0; t < j; t++) c[t] = t; /* gsl_combination_calloc */
            do {
                int drop = 0;
                memset(row, 0, sizeof(int) * d);
                for (int t = 0; t < j; t++) {
                    int p = c[t];
                    if (p == node || banned[node + p * d] == 1) { drop = 1; break; } /* :52-56 */
                    row[p] = 1;
                }
                if (!drop && !isretained(node, row, retained, d)) drop = 1;             /* :61-62 */
                if (!drop) {
                    if (children) {
                        children[total] = node;
                        memcpy(node_defn + total * d, row, sizeof(int) * d);
                    }
                    total++;
                }
            } while (comb_next(c, d, j));
        }
    }
    free(c);
    free(row);
    return total;
}

int64_t or_enumerate_mle(int d, const int *banned, const int *retained, int max_parents,
                         int *children, int *node_defn)
{
    /* R/build_score_cache_mle.R:444-479: per child j an all-zero row, then for i = 1..mp the
     * rows of combn(d-1, i) (lexicographic) as indicator vectors of length d-1 with a 0
     * spliced in at position j; :492-516 then deletes rows violating retain / ban. The
     * deletions keep the order, so filtering while generating gives the same list. */
    int64_t total = 0;
    int *c = (int *)malloc(sizeof(int) * (d + 1));
    int *row = (int *)malloc(sizeof(int) * d);
    if (max_parents >= d) max_parents = d - 1; /* :465-468 */
    for (int node = 0; node < d; node++) {
        for (int i = 0; i <= max_parents; i++) {
            for (int t = 0; t < i; t++) c[t] = t;
            do {
                memset(row, 0, sizeof(int) * d);
                for (int t = 0; t < i; t++) {
                    int p = c[t];            /* index among the d-1 non-child variables */
                    row[p >= node ? p + 1 : p] = 1; /* append(x, 0, after = j-1), :471 */
                }
                int drop = 0;
                for (int p = 0; p < d && !drop; p++) {
                    if (retained[node + p * d] != 0 && row[p] == 0) drop = 1; /* :496-503 */
                    if (banned[node + p * d] != 0 && row[p] == 1) drop = 1;   /* :506-513 */
                }
                if (!drop) {
                    if (children) {
                        children[total] = node;
                        memcpy(node_defn + total * d, row, sizeof(int) * d);
                    }
                    total++;
                }
            } while (comb_next(c, d - 1, i));
        }
    }
    free(c);
    free(row);
    return total;
}
