/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * CPU restatement of the detection losses' arithmetic.  The reference holds no fixtures; all three are pinned bit
 * for bit against the reference's own operator sources compiled through oracle/shim (tests/test_oracle_ref_cxx.py,
 * vectors in tests/golden/reference_cxx_ops.npz).  mshadow expression templates evaluate element-wise in float,
 * left to right.
 *   oracle_focal_loss_backward   operator_cxx/contrib/focal_loss-inl.h:180-230
 *   oracle_bbox_norm_backward    operator_cxx/contrib/bbox_norm-inl.h:99-129
 *   oracle_sigmoid_ce_forward/backward   operator_cxx/contrib/sigmoid_cross_entropy.cu:45-129
 */
#include <math.h>
#include <stddef.h>

/* out = sigmoid(data) is the op's forward (:113); the backward reads `out`, not `data`. */
void oracle_sigmoid(const float* x, long n, float* y) {
  for (long i = 0; i < n; ++i) y[i] = 1.0f / (1.0f + expf(-x[i])); /* mshadow_op::sigmoid */
}

/* out (B,N,K) = sigmoid probabilities, label (B,N) in {-1, 0, 1..K}.  normalization: 0 null,
 * 1 batch, 2 valid.  ograd may be NULL (out_grad = false). */
void oracle_focal_loss_backward(const float* out, const float* label, const float* ograd, int B, int N,
                                int K, float alpha, float gamma, float grad_scale, int normalization,
                                float* gdata) {
  const long rows = (long)B * N;
  float temp = 0.f;
  for (long r = 0; r < rows; ++r) temp += (1.f <= label[r]) ? 1.f : 0.f; /* :218-219 */
  temp = temp + 1.f;                                                      /* :220 */
  for (long r = 0; r < rows; ++r) {
    const float l = label[r];
    const int hot = (int)(l - 1.f); /* one_hot index = label - 1 (:198-201) */
    for (int k = 0; k < K; ++k) {
      const float p = out[r * K + k];
      float g;
      if (l - 1.f >= 0.f && hot == k && hot < K) { /* positive (:192-193) */
        g = alpha * powf(1.f - p, gamma) * (gamma * p * logf(p + 1e-14f) + p - 1.f);
      } else { /* negative (:194-197) */
        g = -((1.f - alpha) * powf(p, gamma) * (gamma * (1.f - p) * logf(1.f - p + 1e-14f) - p));
      }
      if (l == -1.f) g = 0.f; /* ignore (:205-209) */
      if (ograd) g *= ograd[r * K + k];
      if (normalization == 2) g = g * grad_scale / temp;
      else if (normalization == 1) g = g * (grad_scale / B);
      else g = g * grad_scale;
      gdata[r * K + k] = g;
    }
  }
}

/* gout (B, M) flattened to 2D by the leading dim; label (B, L).  gdata = gout / max(sum(label>=1)+1, 1) */
void oracle_bbox_norm_backward(const float* gout, long n, const float* label, long nl, float* gdata) {
  float temp = 0.f;
  for (long i = 0; i < nl; ++i) temp += (1.f <= label[i]) ? 1.f : 0.f;
  temp = temp + 1.f;
  temp = 1.f > temp ? 1.f : temp;
  for (long i = 0; i < n; ++i) gdata[i] = gout[i] / temp;
}

/* data, label (R, D): per-row mean BCE-with-logits over non-ignored (label != -1) entries.
 * The `-1.` / `1.` literals make parts of the expressions double (sigmoid_cross_entropy.cu:57-61,82). */
void oracle_sigmoid_ce_forward(const float* x, const float* t, int R, long D, float* out) {
  for (int r = 0; r < R; ++r) {
    float loss_sum = 0.f, count_sum = 0.f;
    for (long i = 0; i < D; ++i) {
      const float xi = x[r * D + i], ti = t[r * D + i];
      float l = 0.f, c = 0.f;
      if (ti != -1.f) {
        const int ge = xi >= 0;
        l = (float)(-1. * xi * (ti - ge) + logf(1 + expf(xi - 2 * xi * ge)));
        c = 1.f;
      }
      loss_sum += l;
      count_sum += c;
    }
    count_sum += 1e-5f;
    out[r] = loss_sum / count_sum;
  }
}

void oracle_sigmoid_ce_backward(const float* x, const float* t, int R, long D, float scale, float* dx) {
  for (int r = 0; r < R; ++r) {
    float count_sum = 0.f;
    for (long i = 0; i < D; ++i) count_sum += (t[r * D + i] != -1.f) ? 1.f : 0.f;
    count_sum += 1e-5f;
    for (long i = 0; i < D; ++i) {
      const float xi = x[r * D + i], ti = t[r * D + i];
      float d = 0.f;
      if (ti != -1.f) d = (float)(1. / (1. + expf(-xi)) - ti);
      d /= count_sum;
      d *= scale;
      dx[r * D + i] = d;
    }
  }
}
