/*
 * TEST INFRASTRUCTURE - CPU oracle of the neighbour-table build, plain C.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this; nothing under lantern_b200/ does.
 *
 * Restates oracle/lantern_oracle.py:neighbor_table (which is checked against it bit for bit in
 * tests/test_oracle_c.py) so that the full BASELINE sizes (16384 x 8, 8192 x 256) can be verified in seconds:
 *   reference: entrypoints/generate_codebook.py:53-60  (cdist -> fill_diagonal_(inf) -> topk(N-1, largest=False))
 *   pinned definition: squared L2 distance by direct differences accumulated in fp64 in dimension
 *   order, neighbours ordered by (distance, id), self excluded.
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (no FMA contraction: every product and sum is rounded
 * separately, exactly like the NumPy statement `d += diff * diff`).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double d; int32_t j; } pair_t;

static int cmp_pair(const void* a, const void* b) {
  const pair_t* x = (const pair_t*)a;
  const pair_t* y = (const pair_t*)b;
  if (x->d < y->d) return -1;
  if (x->d > y->d) return 1;
  return (x->j > y->j) - (x->j < y->j);
}

/* E: [N, d] fp32 row-major; out: [N, K] int32.  Returns 0, or -1 on a bad argument / allocation failure. */
int lantern_oracle_neighbor_table(const float* E, int32_t N, int32_t d, int32_t K, int32_t* out) {
  if (!E || !out || N < 2 || d < 1 || K < 1 || K > N - 1) return -1;
  int fail = 0;
#pragma omp parallel
  {
    pair_t* row = (pair_t*)malloc((size_t)N * sizeof(pair_t));
    double* ei = (double*)malloc((size_t)d * sizeof(double));
    if (!row || !ei) {
#pragma omp atomic write
      fail = 1;
    } else {
#pragma omp for schedule(dynamic, 16)
      for (int32_t i = 0; i < N; ++i) {
        for (int32_t c = 0; c < d; ++c) ei[c] = (double)E[(size_t)i * d + c];
        int32_t m = 0;
        for (int32_t j = 0; j < N; ++j) {
          if (j == i) continue;
          const float* ej = E + (size_t)j * d;
          double acc = 0.0;
          for (int32_t c = 0; c < d; ++c) {
            const double df = ei[c] - (double)ej[c];
            acc += df * df;
          }
          row[m].d = acc;
          row[m].j = j;
          ++m;
        }
        qsort(row, (size_t)m, sizeof(pair_t), cmp_pair);
        for (int32_t t = 0; t < K; ++t) out[(size_t)i * K + t] = row[t].j;
      }
    }
    free(row);
    free(ei);
  }
  return fail ? -1 : 0;
}

/* The reference's own arithmetic for information (mismatch count vs fp32 cdist + topk is reported, not
 * gated): fp32 squared distances through the expanded form |a|^2 + |b|^2 - 2ab accumulated in fp32, the way
 * torch.cdist's matmul path forms them (generate_codebook.py:55), ties by id.  Counts the positions of the first K
 * columns where that order differs from `exact` ([N, K], from the function above). */
int64_t lantern_oracle_fp32_order_mismatches(const float* E, int32_t N, int32_t d, int32_t K, const int32_t* exact) {
  if (!E || !exact || N < 2 || d < 1 || K < 1 || K > N - 1) return -1;
  float* nrm = (float*)malloc((size_t)N * sizeof(float));
  if (!nrm) return -1;
  for (int32_t i = 0; i < N; ++i) {
    float s = 0.f;
    for (int32_t c = 0; c < d; ++c) s += E[(size_t)i * d + c] * E[(size_t)i * d + c];
    nrm[i] = s;
  }
  int64_t total = 0;
#pragma omp parallel reduction(+ : total)
  {
    pair_t* row = (pair_t*)malloc((size_t)N * sizeof(pair_t));
    if (row) {
#pragma omp for schedule(dynamic, 16)
      for (int32_t i = 0; i < N; ++i) {
        int32_t m = 0;
        for (int32_t j = 0; j < N; ++j) {
          if (j == i) continue;
          float dot = 0.f;
          for (int32_t c = 0; c < d; ++c) dot += E[(size_t)i * d + c] * E[(size_t)j * d + c];
          float v = nrm[i] + nrm[j] - 2.0f * dot;
          if (v < 0.f) v = 0.f;
          row[m].d = (double)v;
          row[m].j = j;
          ++m;
        }
        qsort(row, (size_t)m, sizeof(pair_t), cmp_pair);
        for (int32_t t = 0; t < K; ++t) total += row[t].j != exact[(size_t)i * K + t];
      }
      free(row);
    }
  }
  free(nrm);
  return total;
}
