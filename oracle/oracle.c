/* TEST INFRASTRUCTURE ONLY - CPU restatement (plain C) of the integer/byte parts of the reference hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call this.
 * The product path never links or loads it.
 *
 * PARITY PINNING: the reference has no tests or golden vectors. The functions that are
 * pure Python in the reference (pred_to_string, TopKCERSampler/CerRangeSampler.query, padder/get_text_stack,
 * AddGaussianNoice arithmetic) are pinned by tests/golden/ fixtures generated from the real reference imported
 * in the build container (oracle/gen_golden.py). Levenshtein.distance lives in python-Levenshtein==0.12.0
 * (requirements.txt:70), which is NOT in /root/reference and not installable offline: "parity unpinned" for
 * that one function - it is restated from its published definition (unit-cost insert/delete/substitute edit
 * distance over code points) and pinned only by textbook known answers and by the fixture consistency check
 * cer * max(1, len(label)) in Z over cer_data_utils/pos_dataset_cers.json.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* Levenshtein.distance(a, b) as called at utils.py:106: full (la+1) x (lb+1) Wagner-Fischer matrix. */
int oracle_levenshtein(const int32_t* a, int la, const int32_t* b, int lb) {
  int* d = (int*)malloc(sizeof(int) * (size_t)(la + 1) * (size_t)(lb + 1));
  const int w = lb + 1;
  for (int i = 0; i <= la; ++i) d[i * w] = i;
  for (int j = 0; j <= lb; ++j) d[j] = j;
  for (int i = 1; i <= la; ++i) {
    for (int j = 1; j <= lb; ++j) {
      int sub = d[(i - 1) * w + (j - 1)] + (a[i - 1] != b[j - 1] ? 1 : 0);
      int del = d[(i - 1) * w + j] + 1;
      int ins = d[i * w + (j - 1)] + 1;
      int m = sub < del ? sub : del;
      d[i * w + j] = m < ins ? m : ins;
    }
  }
  int r = d[la * w + lb];
  free(d);
  return r;
}

/* compare_labels(preds, labels) utils.py:95-110 over CSR-encoded strings (int32 code points).
 * lab = labels (denominator), prd = preds. Writes per-pair distance and CER; returns the exact-match count and
 * accumulates total_cer in list order in double, as the Python loop does. */
int oracle_compare_labels(const int32_t* prd, const int32_t* prd_off, const int32_t* lab, const int32_t* lab_off, int n,
                          int32_t* dist, double* cer, double* total_cer) {
  int correct = 0;
  double total = 0.0;
  for (int i = 0; i < n; ++i) {
    const int lp = prd_off[i + 1] - prd_off[i], ll = lab_off[i + 1] - lab_off[i];
    const int32_t* p = prd + prd_off[i];
    const int32_t* l = lab + lab_off[i];
    if (lp == ll && memcmp(p, l, sizeof(int32_t) * (size_t)lp) == 0) correct++;
    const int d = oracle_levenshtein(l, ll, p, lp);
    const double c = (double)d / (double)(ll > 1 ? ll : 1);
    if (dist) dist[i] = d;
    if (cer) cer[i] = c;
    total += c;
  }
  if (total_cer) *total_cer = total;
  return correct;
}

/* pred_to_string utils.py:74-92: per sample, argmax over classes per timestep (first maximum), append when
 * non-blank(0) and (output empty or different from previous timestep's class). scores (T,B,V) row-major. */
void oracle_greedy_decode(const float* scores, int T, int B, int V, int32_t* out, int32_t* out_len) {
  for (int b = 0; b < B; ++b) {
    int n = 0, prev = -1;
    for (int t = 0; t < T; ++t) {
      const float* r = scores + ((size_t)t * B + b) * V;
      int bi = 0;
      float best = r[0];
      for (int c = 1; c < V; ++c) {
        /* torch.argmax: NaN is treated as the maximum, first occurrence wins */
        if ((r[c] > best) || (r[c] != r[c] && best == best)) { best = r[c]; bi = c; }
      }
      if (n == 0) { if (bi != 0) out[(size_t)b * T + n++] = bi; }
      else if (bi != 0 && prev != bi) out[(size_t)b * T + n++] = bi;
      prev = bi;
    }
    out_len[b] = n;
    for (int i = n; i < T; ++i) out[(size_t)b * T + i] = -1;
  }
}

/* TopKCERSampler.query selection_utils.py:144-151 with the stable tie order:
 * indices of the k largest values, descending, equal values lowest-index-first. Insertion sort (stable). */
void oracle_topk_stable(const float* v, int n, int k, int64_t* out) {
  int* idx = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) {
    int j = i;
    while (j > 0 && v[idx[j - 1]] < v[i]) { idx[j] = idx[j - 1]; --j; }
    idx[j] = i;
  }
  for (int i = 0; i < k && i < n; ++i) out[i] = idx[i];
  free(idx);
}

/* CerRangeSampler.query selection_utils.py:118-134: points = (max-min)*rand + min in float32 (separate
 * multiply and add), then per point the first argmin of |point - copy| and copy[index] = 100. */
void oracle_range_select(const float* v, int n, const float* rands, int k, int64_t* out, float* points) {
  if (n == 0) return;
  float* c = (float*)malloc(sizeof(float) * (size_t)n);
  float mx = v[0], mn = v[0];
  for (int i = 0; i < n; ++i) { c[i] = v[i]; if (v[i] > mx) mx = v[i]; if (v[i] < mn) mn = v[i]; }
  volatile float span = mx - mn;
  for (int j = 0; j < k; ++j) {
    volatile float prod = span * rands[j];
    volatile float point = prod + mn;
    if (points) points[j] = point;
    int bi = 0;
    float best = fabsf(point - c[0]);
    for (int i = 1; i < n; ++i) {
      volatile float df = point - c[i];
      float d = fabsf(df);
      if (d < best) { best = d; bi = i; }
    }
    out[j] = bi;
    c[bi] = 100.0f;
  }
  free(c);
}

/* padder + get_text_stack utils.py:118-141 for one (H,W) image: crop [y0:y1, x0:x1] (python slice clipping),
 * centre with floor-division pads, fill 1.0. boxes (n,4) = x_min,y_min,x_max,y_max. out (n,oh,ow). */
static int floordiv2(int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
void oracle_crop_pad(const float* img, int H, int W, const int32_t* boxes, int n, int oh, int ow, float* out) {
  for (int b = 0; b < n; ++b) {
    const int x0 = clampi(boxes[4 * b + 0], 0, W), y0 = clampi(boxes[4 * b + 1], 0, H);
    const int x1 = clampi(boxes[4 * b + 2], x0, W), y1 = clampi(boxes[4 * b + 3], y0, H);
    const int cw = x1 - x0, ch = y1 - y0;
    const int pl = floordiv2(ow - cw), pt = floordiv2(oh - ch);
    for (int y = 0; y < oh; ++y)
      for (int x = 0; x < ow; ++x) {
        const int cy = y - pt, cx = x - pl;
        float v = 1.0f;
        if (cy >= 0 && cy < ch && cx >= 0 && cx < cw) v = img[(size_t)(y0 + cy) * W + (x0 + cx)];
        out[((size_t)b * oh + y) * ow + x] = v;
      }
  }
}

/* AddGaussianNoice.__call__ transform_helper.py:40-41 on a given noise tensor: clamp(img - coef*noise, 0, 1),
 * multiply and subtract rounded separately in float32. */
void oracle_jitter_apply(const float* img, const float* noise, float coef, size_t n, float* out) {
  for (size_t i = 0; i < n; ++i) {
    volatile float p = coef * noise[i];
    volatile float v = img[i] - p;
    float r = v;
    if (r < 0.0f) r = 0.0f;
    if (r > 1.0f) r = 1.0f;
    out[i] = r;
  }
}

/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator the jitter kernel uses. */
void oracle_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
