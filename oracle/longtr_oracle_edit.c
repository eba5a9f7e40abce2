/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the candidate-haplotype clustering arithmetic.
 * Follows the reference line by line; the product never links or loads this file.
 *
 *   ltr_oracle_edit_score      HaplotypeGenerator::needleman_wunsch   src/SeqAlignment/HaplotypeGenerator.cpp:201-235
 *   ltr_oracle_greedy_cluster  HaplotypeGenerator::greedy_clustering  src/SeqAlignment/HaplotypeGenerator.cpp:238-271
 *
 * Pinned by oracle/_ref/libltr_ref.so (oracle/edit_driver.cpp calls the reference's own member functions, compiled in
 * place): tests/test_oracle_edit.py holds the two to each other on seeded inputs and on tests/golden/edit.json.
 */
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>

/* :201-235.  cent = rows (n), read = columns (m). */
int32_t ltr_oracle_edit_score(const uint8_t* cent, int32_t n, const uint8_t* read, int32_t m, int32_t T) {
  if (abs(n - m) > T) return T + 1; /* :203-206 */
  const int32_t gap = 1, match = 0, mismatch = 1;
  int32_t* dp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1) * (size_t)(m + 1));
  const size_t w = (size_t)m + 1;
  for (int32_t i = 0; i < n + 1; i++) dp[(size_t)i * w] = i * gap;      /* :212-214 */
  for (int32_t j = 0; j < m + 1; j++) dp[j] = j * gap;                   /* :216-218 */
  for (int32_t i = 1; i < n + 1; i++) {
    int32_t min_row = 1000;                                             /* :221 */
    for (int32_t j = 1; j < m + 1; j++) {
      const int32_t S = (cent[i - 1] == read[j - 1]) ? match : mismatch;
      int32_t a = dp[(size_t)(i - 1) * w + j] + gap, b = dp[(size_t)i * w + j - 1] + gap, c = dp[(size_t)(i - 1) * w + j - 1] + S;
      int32_t v = b < c ? b : c;
      v = a < v ? a : v;
      dp[(size_t)i * w + j] = v;                                         /* :224-225 */
      const int32_t t = v + abs((n - m) - (i - j));
      if (t < min_row) min_row = t;                                      /* :226 */
    }
    if (min_row > T) {                                                   /* :228-231 */
      free(dp);
      return T + 1;
    }
  }
  const int32_t score = dp[(size_t)n * w + m];
  free(dp);
  return score;
}

/* :238-271.  items[0..n_items) name sequences; centroid_of[i] = position (in items) of the centroid item i joins.
 * Returns 1, or 0 when a 16th centroid would be needed (:262-264; the assignments made so far stay in place). */
int32_t ltr_oracle_greedy_cluster(const uint8_t* seq_bytes, const uint32_t* seq_off, const uint32_t* items,
                                  int32_t n_items, int32_t threshold, int32_t* centroid_of, int32_t* n_centroids_out) {
  int32_t centroids[16];
  int32_t n_centroids = 0;
  *n_centroids_out = 0;
  if (n_items <= 0) return 1;
  centroids[n_centroids++] = 0; /* :240-241 */
  centroid_of[0] = 0;
  for (int32_t i = 1; i < n_items; i++) {
    int32_t min_score = INT_MAX, min_cntr = -1;
    const uint32_t si = items[i];
    for (int32_t j = 0; j < n_centroids; j++) {
      const int32_t T = threshold; /* :247 */
      const uint32_t sc = items[centroids[j]];
      const int32_t score = ltr_oracle_edit_score(seq_bytes + seq_off[si], (int32_t)(seq_off[si + 1] - seq_off[si]),
                                                  seq_bytes + seq_off[sc], (int32_t)(seq_off[sc + 1] - seq_off[sc]), T);
      if ((score < T) & (score < min_score)) { /* :252-255 */
        min_cntr = j;
        min_score = score;
      }
    }
    if (min_cntr != -1) centroid_of[i] = centroids[min_cntr]; /* :257-259 */
    else {
      if (n_centroids + 1 > 15) { /* :261-264: the push happens first, then the size test */
        *n_centroids_out = n_centroids + 1;
        return 0;
      }
      centroids[n_centroids++] = i;
      centroid_of[i] = i;
    }
  }
  *n_centroids_out = n_centroids;
  return 1;
}
