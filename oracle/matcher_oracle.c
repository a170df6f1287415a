/*
 * ORACLE (test infrastructure only -- never linked into or called from the product path).
 *
 * Plain-C restatement of the reference's Hungarian matcher:
 *   - cost blocks: utils/matcher.py:53-72 + utils/box_ops.py:9-13 (box_cxcywh_to_xyxy), 24-37 (box_iou),
 *     40-59 (generalized_box_iou), torchvision.ops.boxes.box_area, torch.cdist(p=1), softmax(-1);
 *   - assignment: scipy.optimize.linear_sum_assignment (matcher.py:76).  scipy is a third-party dependency
 *     that is not vendored in /root/reference (unpinned; pulled in by scikit-image 0.17.2, setup_conda_env.sh:5;
 *     this container has scipy 1.18.1).  The algorithm restated here is the published one scipy implements:
 *     D. F. Crouse, "On implementing 2D rectangular assignment algorithms", IEEE TAES 52(4), 2016 -- shortest
 *     augmenting paths with dual variables, rows > cols solved on the transpose, output rows ascending.
 *
 * Pinning: oracle/pin_matcher.py checks both functions against the reference's own HungarianMatcher.forward
 * (imported from /root/reference) and against scipy on random, ragged, empty and tie-heavy inputs, and writes
 * the fixtures under tests/golden/matcher_*.npz.  Build:  gcc -O2 -ffp-contract=off -shared -fPIC.
 * -ffp-contract=off matters: every a*b+c below must round twice, like the reference's separate tensor ops.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* cost[b][q][t] for ragged targets; dense blocks with row stride Tmax (entries t >= T_b untouched). */
void oracle_matcher_cost(const float* logits, const float* boxes, const float* tboxes, const int64_t* tlabels,
                         const int32_t* toff, int B, int Q, int C, int Tmax, float w_class, float w_bbox,
                         float w_giou, float* cost) {
  for (int b = 0; b < B; ++b) {
    const int t0 = toff[b], T = toff[b + 1] - t0;
    for (int q = 0; q < Q; ++q) {
      const float* lg = logits + ((size_t)b * Q + q) * C;
      /* softmax over the class dim (matcher.py:56): ATen CPU kernel = exp(x - max) * (1 / sum) */
      float mx = lg[0];
      for (int c = 1; c < C; ++c) mx = lg[c] > mx ? lg[c] : mx;
      float sum = 0.0f;
      for (int c = 0; c < C; ++c) sum = sum + expf(lg[c] - mx);
      const float inv = 1.0f / sum;
      const float* ob = boxes + ((size_t)b * Q + q) * 4;
      /* box_cxcywh_to_xyxy (box_ops.py:9-13) */
      const float ax0 = ob[0] - 0.5f * ob[2], ay0 = ob[1] - 0.5f * ob[3];
      const float ax1 = ob[0] + 0.5f * ob[2], ay1 = ob[1] + 0.5f * ob[3];
      const float area1 = (ax1 - ax0) * (ay1 - ay0);
      for (int t = 0; t < T; ++t) {
        const float* tb = tboxes + (size_t)(t0 + t) * 4;
        const float prob = expf(lg[tlabels[t0 + t]] - mx) * inv;
        const float cost_class = -prob; /* matcher.py:63 */
        /* cdist p=1 (matcher.py:66): ((|d0|+|d1|)+|d2|)+|d3| */
        float l1 = fabsf(ob[0] - tb[0]);
        l1 = l1 + fabsf(ob[1] - tb[1]);
        l1 = l1 + fabsf(ob[2] - tb[2]);
        l1 = l1 + fabsf(ob[3] - tb[3]);
        const float bx0 = tb[0] - 0.5f * tb[2], by0 = tb[1] - 0.5f * tb[3];
        const float bx1 = tb[0] + 0.5f * tb[2], by1 = tb[1] + 0.5f * tb[3];
        const float area2 = (bx1 - bx0) * (by1 - by0);
        /* box_iou (box_ops.py:24-37) */
        float iw = fminf(ax1, bx1) - fmaxf(ax0, bx0);
        float ih = fminf(ay1, by1) - fmaxf(ay0, by0);
        iw = iw > 0.0f ? iw : 0.0f;
        ih = ih > 0.0f ? ih : 0.0f;
        const float inter = iw * ih;
        const float uni = (area1 + area2) - inter;
        const float iou = inter / uni;
        /* generalized_box_iou (box_ops.py:53-59) */
        float ew = fmaxf(ax1, bx1) - fminf(ax0, bx0);
        float eh = fmaxf(ay1, by1) - fminf(ay0, by0);
        ew = ew > 0.0f ? ew : 0.0f;
        eh = eh > 0.0f ? eh : 0.0f;
        const float earea = ew * eh;
        const float giou = iou - (earea - uni) / earea;
        const float cost_giou = -giou;
        /* matcher.py:72, evaluated left to right */
        float c = w_bbox * l1 + w_class * cost_class;
        c = c + w_giou * cost_giou;
        cost[((size_t)b * Q + q) * Tmax + t] = c;
      }
    }
  }
}

/* Shortest augmenting path from row i (Crouse 2016, Alg. 1 inner loop; scipy's scan order and tie rule). */
static int augmenting_path(int nc, const double* cost, const double* u, const double* v, int* path,
                           const int* row4col, double* spc, int i, unsigned char* SR, unsigned char* SC,
                           int* remaining, double* p_minVal) {
  double minVal = 0.0;
  int num_remaining = nc;
  for (int it = 0; it < nc; ++it) remaining[it] = nc - it - 1;
  memset(SC, 0, (size_t)nc);
  for (int j = 0; j < nc; ++j) spc[j] = INFINITY;
  int sink = -1;
  while (sink == -1) {
    int index = -1;
    double lowest = INFINITY;
    SR[i] = 1;
    for (int it = 0; it < num_remaining; ++it) {
      const int j = remaining[it];
      const double r = minVal + cost[(size_t)i * nc + j] - u[i] - v[j];
      if (r < spc[j]) {
        path[j] = i;
        spc[j] = r;
      }
      /* among equal minima prefer a column that is still unassigned (it ends the search) */
      if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) {
        lowest = spc[j];
        index = it;
      }
    }
    minVal = lowest;
    if (minVal == INFINITY) return -1;
    const int j = remaining[index];
    if (row4col[j] == -1) sink = j; else i = row4col[j];
    SC[j] = 1;
    remaining[index] = remaining[--num_remaining];
  }
  *p_minVal = minVal;
  return sink;
}

/* Solve one nr x nc problem (row-major fp64 cost). Writes min(nr,nc) pairs (a[k], b[k]) with a ascending.
 * Returns the number of pairs, or -1 if infeasible. */
int oracle_lsap_f64(int nr_in, int nc_in, const double* cost_in, int64_t* a, int64_t* b) {
  if (nr_in == 0 || nc_in == 0) return 0;
  const int transpose = nc_in < nr_in;
  int nr = nr_in, nc = nc_in;
  double* cost = (double*)malloc(sizeof(double) * (size_t)nr_in * nc_in);
  if (transpose) {
    for (int i = 0; i < nr_in; ++i)
      for (int j = 0; j < nc_in; ++j) cost[(size_t)j * nr_in + i] = cost_in[(size_t)i * nc_in + j];
    nr = nc_in;
    nc = nr_in;
  } else {
    memcpy(cost, cost_in, sizeof(double) * (size_t)nr * nc);
  }
  double* u = (double*)calloc((size_t)nr, sizeof(double));
  double* v = (double*)calloc((size_t)nc, sizeof(double));
  double* spc = (double*)malloc(sizeof(double) * (size_t)nc);
  int* path = (int*)malloc(sizeof(int) * (size_t)nc);
  int* col4row = (int*)malloc(sizeof(int) * (size_t)nr);
  int* row4col = (int*)malloc(sizeof(int) * (size_t)nc);
  int* remaining = (int*)malloc(sizeof(int) * (size_t)nc);
  unsigned char* SR = (unsigned char*)malloc((size_t)nr);
  unsigned char* SC = (unsigned char*)malloc((size_t)nc);
  for (int i = 0; i < nr; ++i) col4row[i] = -1;
  for (int j = 0; j < nc; ++j) { row4col[j] = -1; path[j] = -1; }
  int ok = 1;
  for (int cur = 0; cur < nr && ok; ++cur) {
    double minVal = 0.0;
    memset(SR, 0, (size_t)nr);
    const int sink = augmenting_path(nc, cost, u, v, path, row4col, spc, cur, SR, SC, remaining, &minVal);
    if (sink < 0) { ok = 0; break; }
    u[cur] += minVal;
    for (int i = 0; i < nr; ++i)
      if (SR[i] && i != cur) u[i] += minVal - spc[col4row[i]];
    for (int j = 0; j < nc; ++j)
      if (SC[j]) v[j] -= minVal - spc[j];
    int j = sink;
    while (1) {
      const int i = path[j];
      row4col[j] = i;
      const int tmp = col4row[i];
      col4row[i] = j;
      j = tmp;
      if (i == cur) break;
    }
  }
  int n = -1;
  if (ok) {
    n = nr;
    if (transpose) {
      /* rows of the original problem are col4row values: emit them ascending */
      int k = 0;
      for (int c = 0; c < nc; ++c)
        if (row4col[c] != -1) { a[k] = c; b[k] = row4col[c]; ++k; }
    } else {
      for (int i = 0; i < nr; ++i) { a[i] = i; b[i] = col4row[i]; }
    }
  }
  free(cost); free(u); free(v); free(spc); free(path); free(col4row); free(row4col); free(remaining); free(SR); free(SC);
  return n;
}

/* Batched front end with the product's calling convention: fp32 cost blocks [B][Q][Tmax], ragged T_b. */
void oracle_lsap(const float* cost, const int32_t* toff, int B, int Q, int Tmax, int64_t* out_q, int64_t* out_t) {
  const int Kmax = Q < Tmax ? Q : Tmax;
  for (int b = 0; b < B; ++b) {
    const int T = toff[b + 1] - toff[b];
    for (int k = 0; k < Kmax; ++k) { out_q[(size_t)b * Kmax + k] = -1; out_t[(size_t)b * Kmax + k] = -1; }
    if (T == 0) continue;
    double* c = (double*)malloc(sizeof(double) * (size_t)Q * T);
    for (int q = 0; q < Q; ++q)
      for (int t = 0; t < T; ++t) c[(size_t)q * T + t] = (double)cost[((size_t)b * Q + q) * Tmax + t];
    oracle_lsap_f64(Q, T, c, out_q + (size_t)b * Kmax, out_t + (size_t)b * Kmax);
    free(c);
  }
}
