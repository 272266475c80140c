/* ORACLE (test infrastructure only) — plain-C restatement of the scalar GRPO arithmetic of Time-R1.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 *
 *   oracle_iou            /root/reference/main.py:163-168 (iou), :205-218 (iou_v2)
 *   oracle_group_adv      src/time_r1/rl/timer1_trainer.py:700-712 (sum of reward fns, mean, unbiased std, +1e-4)
 *   oracle_grpo_loss      src/time_r1/rl/timer1_trainer.py:635-639 (k3 KL), :713-737 (loss), :763-777 (metrics)
 * Pinned by tests/golden/iou_train_2k5.json (2500 known-answer rows shipped with the reference) and
 * tests/golden/rewards_golden.json (outputs of the unmodified reference functions).
 */
#include <math.h>
#include <stddef.h>

static double dmax(double a, double b) { return a > b ? a : b; }
static double dmin(double a, double b) { return a < b ? a : b; }

/* returns NaN where the reference leaves `iou` unbound (union <= 0) */
double oracle_iou(int mode, double ps, double pe, double gs, double ge, double dur) {
  double inter = dmax(0.0, dmin(pe, ge) - dmax(ps, gs));
  double uni = dmax(pe, ge) - dmin(ps, gs);
  if (!(uni > 0.0)) return NAN;
  double iou = inter / uni;
  if (mode == 0) return iou;
  double gsn = 1.0 * gs / dur, gen = 1.0 * ge / dur, psn = 1.0 * ps / dur, pen = 1.0 * pe / dur;
  return iou * (1 - fabs(gsn - psn)) * (1 - fabs(gen - pen));
}

void oracle_group_adv(const float* rpf, int n, int nf, int G, float* rewards, float* adv, float* mean, float* std_) {
  for (int g = 0; g < n / G; ++g) {
    float m = 0.f;
    for (int j = 0; j < G; ++j) {
      float r = 0.f;
      for (int f = 0; f < nf; ++f) r += rpf[(size_t)(g * G + j) * nf + f];
      rewards[g * G + j] = r;
      m += r;
    }
    m /= (float)G;
    float v = 0.f;
    for (int j = 0; j < G; ++j) v += (rewards[g * G + j] - m) * (rewards[g * G + j] - m);
    float sd = sqrtf(v / (float)(G - 1));
    for (int j = 0; j < G; ++j) adv[g * G + j] = (rewards[g * G + j] - m) / (sd + 1e-4f);
    mean[g] = m;
    std_[g] = sd;
  }
}

/* out[0]=loss out[1]=mean_kl out[2]=mean completion length; dlogp[rows*C] = dL/dlogp (on-policy: ratio == 1) */
void oracle_grpo_loss(const float* logp, const float* ref_logp, const float* adv, const int* mask, int rows, int C,
                      float beta, int use_grpo, double* out, double* dlogp) {
  double total = 0.0;
  for (int i = 0; i < rows * C; ++i) total += mask[i];
  double loss = 0.0, mkl = 0.0;
  for (int r = 0; r < rows; ++r) {
    double cnt = 0.0, seq_loss = 0.0, seq_kl = 0.0;
    for (int t = 0; t < C; ++t) cnt += mask[r * C + t];
    for (int t = 0; t < C; ++t) {
      int i = r * C + t;
      double kl = 0.0, dkl = 0.0;
      if (beta != 0.f && ref_logp) {
        double d = (double)ref_logp[i] - (double)logp[i];
        kl = exp(d) - d - 1.0;
        dkl = 1.0 - exp(d);
      }
      double tok = -(double)adv[r] + beta * kl; /* -min(1*A, clip(1)*A) = -A */
      double w = use_grpo ? (mask[i] / cnt) / rows : mask[i] / total;
      seq_loss += tok * mask[i];
      seq_kl += kl * mask[i];
      if (dlogp) dlogp[i] = w * (-(double)adv[r] + beta * dkl);
    }
    loss += use_grpo ? (seq_loss / cnt) / rows : seq_loss / total;
    mkl += (seq_kl / cnt) / rows;
  }
  out[0] = loss;
  out[1] = mkl;
  out[2] = total / rows;
}
