/*
 * assign_oracle.c -- exact CPU solver for the balanced n x n assignment problem.
 *
 * TEST INFRASTRUCTURE ONLY (checker for the CUDA path).
 *
 * The reference solves  min sum c[i][j] x[n*i+j]  s.t. row sums = column sums = 1, x binary
 * by handing a dense 2n x n^2 equality system to cvxopt.glpk.ilp (solver.py:11-27,
 * python.py:6-25, split.py:139-155, heuristic.py:7-17,37).  cvxopt/GLPK is a THIRD-PARTY
 * dependency that is absent from /root/reference and from this image, and the reference pins
 * no version (hint only: solver.py:1-4 "conda ... python=3.8 ... cvxopt").  The optimal
 * OBJECTIVE of that model is mathematically unique, so any exact method is a valid oracle for
 * it; the chosen x among tied optima is NOT pinned (KAT A1 has two
 * optimal permutations).  This file restates the published Jonker-Volgenant / Kuhn-Munkres
 * shortest-augmenting-path method with integer potentials (O(n^3), int64 arithmetic).
 * Parity status: objective PINNED by KAT A1 (python.py:7 -> 101), A2 (glpk.mod:27-31 -> 101),
 * A3 (procedure.py:32-51 -> 17) and cross-checked against scipy.optimize.linear_sum_assignment
 * and the LP relaxation in the reference's own constraint layout (oracle/assign_ref.py).
 */
#include <stdint.h>
#include <stdlib.h>

/* col_of_row_out[i] = column assigned to row i; returns the optimal objective via *objective_out */
int assign_oracle(const int32_t *cost, int n, int32_t *col_of_row_out, int64_t *objective_out) {
    if (n < 0) return -2;
    if (n == 0) { if (objective_out) *objective_out = 0; return 0; }
    const int64_t INF = INT64_MAX / 4;
    int64_t *u = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    int64_t *v = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    int64_t *minv = (int64_t *)malloc(((size_t)n + 1) * sizeof(int64_t));
    int32_t *match = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t)); /* row (1-based) matched to column j */
    int32_t *way = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t));
    uint8_t *used = (uint8_t *)malloc((size_t)n + 1);
    if (!u || !v || !minv || !match || !way || !used) return -1;
    for (int i = 1; i <= n; i++) {
        match[0] = i;
        int j0 = 0;
        for (int j = 0; j <= n; j++) { minv[j] = INF; used[j] = 0; }
        do {
            used[j0] = 1;
            int i0 = match[j0], j1 = 0;
            int64_t delta = INF;
            const int32_t *row = cost + (int64_t)(i0 - 1) * n;
            for (int j = 1; j <= n; j++) {
                if (used[j]) continue;
                int64_t cur = (int64_t)row[j - 1] - u[i0] - v[j];
                if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
                if (minv[j] < delta) { delta = minv[j]; j1 = j; }
            }
            for (int j = 0; j <= n; j++) {
                if (used[j]) { u[match[j]] += delta; v[j] -= delta; }
                else minv[j] -= delta;
            }
            j0 = j1;
        } while (match[j0] != 0);
        do {
            int j1 = way[j0];
            match[j0] = match[j1];
            j0 = j1;
        } while (j0);
    }
    int64_t obj = 0;
    for (int j = 1; j <= n; j++) {
        int i = match[j] - 1;
        if (col_of_row_out) col_of_row_out[i] = j - 1;
        obj += cost[(int64_t)i * n + (j - 1)];
    }
    if (objective_out) *objective_out = obj;
    free(u); free(v); free(minv); free(match); free(way); free(used);
    return 0;
}
