/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the
 * product path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 *
 * srcnn_oracle: a CPU restatement, in plain C, of the SRCNN hot path of
 * Scthe/cnn-Super-Resolution.  Every function names the reference file:line it
 * follows (paths relative to /root/reference).  PARITY IS PINNED: this file is
 * checked (tests/test_oracle.py) against
 *   - the reference's golden vectors (test/data/test_cases.json,
 *     test/specs/LayerDeltasTest.cpp:33-126, test/specs/BackpropagationTest.cpp:31-90),
 *   - the formulas of its formula-pinned specs (LastLayerDeltaTest.cpp:55-67,
 *     SquaredErrorTest.cpp:53-60, UpdateParametersTest.cpp:19-27),
 *   - outputs of the reference's OWN kernels compiled by g++ (oracle/_ref, built by
 *     oracle/Makefile from /root/reference/src/kernel/ *.cl), on seeded random inputs,
 *     and fixtures generated from those kernels committed under tests/golden/.
 *
 * Layouts (reference: src/kernel/layer_uber_kernel.cl:1-34,51-56):
 *   activations  [S][H][W][C]        C fastest
 *   weights      [f][f][C_in][C_out] C_out fastest
 *   ground truth [S][H][W]
 * All arithmetic is float32 with the reference's loop order, so that results agree
 * with the reference kernels to rounding (FMA contraction may differ).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* forward: valid convolution + bias (+ ReLU).
 * reference: src/kernel/layer_uber_kernel.cl:36-96
 *   out[s][y][x][n] = act(B[n] + sum_{dy,dx,k} W[dy][dx][k][n] * in[s][y+dy][x+dx][k])
 * accumulation order dy -> dx -> k (lines 70-85), bias added last (line 89),
 * ReLU = max(v,0) unless skip_relu (lines 90-94). */
void oracle_forward(const float* in, float* out, const float* W, const float* B, int k,
                    int n, int f, int skip_relu, int in_w, int in_h, int S) {
  const int ow = in_w - f + 1, oh = in_h - f + 1;
  if (ow <= 0 || oh <= 0) return;
#pragma omp parallel for collapse(2) schedule(static)
  for (int s = 0; s < S; s++) {
    for (int y = 0; y < oh; y++) {
      float* acc = (float*)malloc(sizeof(float) * (size_t)n);
      const float* img = in + (size_t)s * k * in_w * in_h;
      float* dst = out + (size_t)s * n * ow * oh;
      for (int x = 0; x < ow; x++) {
        for (int c = 0; c < n; c++) acc[c] = 0.0f;
        for (int dy = 0; dy < f; dy++)
          for (int dx = 0; dx < f; dx++) {
            const float* px = img + ((size_t)(y + dy) * in_w + (x + dx)) * k;
            const float* w2 = W + (size_t)(dy * f + dx) * n * k;
            for (int kk = 0; kk < k; kk++) {
              const float v = px[kk];
              const float* w3 = w2 + (size_t)kk * n;
              for (int c = 0; c < n; c++) acc[c] += w3[c] * v;
            }
          }
        float* o = dst + ((size_t)y * ow + x) * n;
        for (int c = 0; c < n; c++) {
          float r = acc[c] + B[c];
          o[c] = skip_relu ? r : (r > 0.0f ? r : 0.0f);
        }
      }
      free(acc);
    }
  }
}

/* squared error against the centre crop of the ground truth.
 * reference: src/kernel/squared_error.cl:36-92 (padding = (gt_w - algo_w)/2, line 48;
 * d = y - t, d*d, lines 66-71).  The reference reduces in float32 with a local tree
 * and a CAS atomic in arbitrary order; the deterministic definition used here is the
 * double-precision sum of the float32 squares, like its test does
 * (test/specs/SquaredErrorTest.cpp:53-60). */
double oracle_squared_error(const float* gt, const float* algo, int gt_w, int gt_h,
                            int algo_w, int algo_h, int S) {
  const size_t pad = (size_t)(gt_w - algo_w) / 2;
  double acc = 0.0;
  for (int s = 0; s < S; s++)
    for (int y = 0; y < algo_h; y++)
      for (int x = 0; x < algo_w; x++) {
        const float t = gt[(size_t)s * gt_w * gt_h + (y + pad) * gt_w + pad + x];
        const float v = algo[(size_t)s * algo_w * algo_h + (size_t)y * algo_w + x];
        const float d = v - t;
        acc += (double)(d * d);
      }
  return acc;
}

/* last-layer delta, including quirk Q2: the ReLU derivative [y > 0] is applied even
 * though layer 3 is linear.  reference: src/kernel/last_layer_delta.cl:14-50
 * (d = y - t line 44; relu_deriv line 45; product line 48). */
void oracle_last_layer_delta(const float* gt, const float* algo, float* target, int gt_w,
                             int gt_h, int algo_w, int algo_h, int S) {
  const size_t pad = (size_t)(gt_w - algo_w) / 2;
  for (int s = 0; s < S; s++)
    for (int y = 0; y < algo_h; y++)
      for (int x = 0; x < algo_w; x++) {
        const float t = gt[(size_t)s * gt_w * gt_h + (y + pad) * gt_w + pad + x];
        const size_t i = (size_t)s * algo_w * algo_h + (size_t)y * algo_w + x;
        const float v = algo[i];
        target[i] = (v - t) * (v > 0.0f ? 1.0f : 0.0f);
      }
}

/* deltas of layer (l-1) from deltas of layer (l).
 * reference: src/kernel/layer_deltas.cl:42-127
 *   d_prev[s][j][i][n] = sum_{dy,dx,k} d_next[s][j-dy][i-dx][k] * W_l[dy][dx][n][k] * [out_prev[s][j][i][n] > 0]
 * with d_next taken as 0 outside its extent (lines 92-99), weight index
 * ((dy*f+dx)*n_next*n_curr) + n*n_next + k (lines 82-83,105), the activation
 * derivative multiplied inside the sum (line 112). */
void oracle_deltas(const float* deltas_next, const float* layer_output, float* target,
                   const float* W, int n_curr, int f_next, int n_next, int out_w, int out_h,
                   int S) {
  const int nw = out_w - f_next + 1, nh = out_h - f_next + 1;
#pragma omp parallel for collapse(2) schedule(static)
  for (int s = 0; s < S; s++) {
    for (int j = 0; j < out_h; j++) {
      float* acc = (float*)malloc(sizeof(float) * (size_t)n_curr * 2);
      float* der = acc + n_curr;
      const float* dn = deltas_next + (size_t)s * n_next * nw * nh;
      for (int i = 0; i < out_w; i++) {
        const size_t idx = (size_t)s * n_curr * out_w * out_h + ((size_t)j * out_w + i) * n_curr;
        for (int n = 0; n < n_curr; n++) {
          acc[n] = 0.0f;
          der[n] = layer_output[idx + n] > 0.0f ? 1.0f : 0.0f;
        }
        for (int dy = 0; dy < f_next; dy++)
          for (int dx = 0; dx < f_next; dx++) {
            const int ni = i - dx, nj = j - dy;
            const int in_range = ni >= 0 && ni < nw && nj >= 0 && nj < nh;
            const float* w2 = W + (size_t)(dy * f_next + dx) * n_next * n_curr;
            for (int k = 0; k < n_next; k++) {
              const float d = in_range ? dn[((size_t)nj * nw + ni) * n_next + k] : 0.0f;
              for (int n = 0; n < n_curr; n++) acc[n] += d * w2[(size_t)n * n_next + k] * der[n];
            }
          }
        for (int n = 0; n < n_curr; n++) target[idx + n] = acc[n];
      }
      free(acc);
    }
  }
}

/* weight and bias gradients, ACCUMULATING into grad_w / grad_b.
 * reference: src/kernel/backpropagate.cl:56-114
 *   gW[dy][dx][k][n] += sum_{row,col} d[s][row][col][n] * in[s][row+dy][col+dx][k]   (lines 89-110)
 *   gB[n]            += sum_{row,col} d[s][row][col][n]                              (lines 94,111-112)
 * One partial sum per (weight, sample) in row-major pixel order, then `+=` into the
 * accumulator; samples are visited in order 0..S-1 (the race-free reading of the
 * reference, whose `+=` on gW races across the sample dimension, line 110). */
void oracle_backpropagate(const float* deltas, const float* layer_input, float* grad_w,
                          float* grad_b, int n, int k, int f, int out_w, int out_h, int S) {
  const int in_w = out_w + f - 1, in_h = out_h + f - 1;
  const int wsize = f * f * k * n;
  for (int s = 0; s < S; s++) {
    const float* d = deltas + (size_t)s * n * out_w * out_h;
    const float* in = layer_input + (size_t)s * k * in_w * in_h;
#pragma omp parallel for schedule(static)
    for (int id = 0; id < wsize; id++) {
      const int dy = id / (f * k * n), r1 = id - dy * (f * k * n);
      const int dx = r1 / (k * n), r2 = r1 - dx * (k * n);
      const int kk = r2 / n, nn = r2 - kk * n;
      float gw = 0.0f, gb = 0.0f;
      for (int row = 0; row < out_h; row++)
        for (int col = 0; col < out_w; col++) {
          const float dv = d[((size_t)row * out_w + col) * n + nn];
          gb += dv;
          gw += in[((size_t)(row + dy) * in_w + (col + dx)) * k + kk] * dv;
        }
      grad_w[id] += gw;
      if (kk == 0 && dx == 0 && dy == 0) grad_b[nn] += gb;
    }
  }
}

/* momentum + weight-decay update, quirk Q3 included.
 * reference: src/kernel/update_parameters.cl:1-33
 *   dw = momentum*prev_dw + lr*gW + decay*w ; w -= dw / batch ; prev_dw = dw   (lines 17-24)
 *   db = momentum*prev_db + lr*gB           ; b -= db / batch ; prev_db = db   (lines 27-32)
 * decay is NOT scaled by lr, the whole delta (momentum included) is divided by the
 * batch size, the stored delta is undivided, the bias has no decay. */
void oracle_update_params(float* w, float* b, const float* gw, const float* gb, float* pdw,
                          float* pdb, float momentum, float decay, float lr, unsigned batch,
                          unsigned wsize, unsigned bsize) {
  for (unsigned i = 0; i < wsize; i++) {
    const float wv = w[i];
    const float dw = momentum * pdw[i] + lr * gw[i] + decay * wv;
    w[i] = wv - dw / batch;
    pdw[i] = dw;
  }
  for (unsigned i = 0; i < bsize; i++) {
    const float db = momentum * pdb[i] + lr * gb[i];
    b[i] -= db / batch;
    pdb[i] = db;
  }
}

/* sum / sum of squares.  reference: src/kernel/sum.cl:35-68 (float32 tree + CAS
 * atomic, arbitrary order); deterministic definition: double sum of float32 terms. */
double oracle_sum(const float* data, unsigned len, int squared) {
  double acc = 0.0;
  for (unsigned i = 0; i < len; i++) {
    float v = data[i];
    if (squared) v = v * v;
    acc += (double)v;
  }
  return acc;
}

/* reference: src/kernel/subtract_from_all.cl:1-8 */
void oracle_sub_from_all(float* data, float value, unsigned len) {
  for (unsigned i = 0; i < len; i++) data[i] = data[i] - value;
}

/* DataPipeline::subtract_mean INCLUDING quirk Q1: the cl_event* argument lands in
 * sum()'s `bool squared` parameter, so when an event pointer is passed (both CLI call
 * sites do, src/Main_cl.cpp:141,227) the value subtracted is the mean of SQUARES.
 * reference: src/DataPipeline.cpp:268-280 vs src/DataPipeline.hpp:171.
 * `with_event` != 0 reproduces the CLI behaviour; 0 gives the true mean. */
float oracle_subtract_mean(float* data, unsigned len, int with_event) {
  const float s = (float)oracle_sum(data, len, with_event ? 1 : 0);
  const float mean = s / len;
  oracle_sub_from_all(data, mean, len);
  return mean;
}

/* luma extraction.  reference: src/kernel/extract_luma.cl:7-23
 *   Y = dot((R,G,B,A), (0.299, 0.587, 0.114, 0)) [/ 255 when NORMALIZE] */
void oracle_extract_luma(const unsigned char* rgba, float* target, int w, int h,
                         int normalize) {
  for (int i = 0; i < w * h; i++) {
    const float r = (float)rgba[4 * i + 0], g = (float)rgba[4 * i + 1],
                bl = (float)rgba[4 * i + 2], a = (float)rgba[4 * i + 3];
    (void)a; /* alpha weight is 0 */
    const float y = (r * 0.299f + g * 0.587f) + bl * 0.114f;
    target[i] = normalize ? y / 255.0f : y;
  }
}

static float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* luma swap: new luma (x255) + chroma of the original, border copied.
 * reference: src/kernel/swap_luma.cl:18-69 (coefficients lines 7-15; padding line 24;
 * border copy lines 38-43; YCbCr->RGB, clamp, float->uint truncation lines 50-63). */
void oracle_swap_luma(const unsigned char* rgba, const float* new_luma, unsigned char* target,
                      int gt_w, int gt_h, int luma_w, int luma_h) {
  const int pad = (gt_w - luma_w) / 2;
  for (int y = 0; y < gt_h; y++)
    for (int x = 0; x < gt_w; x++) {
      const size_t i = (size_t)y * gt_w + x;
      const float r = (float)rgba[4 * i + 0], g = (float)rgba[4 * i + 1],
                  b = (float)rgba[4 * i + 2];
      const int lx = x - pad, ly = y - pad;
      unsigned char o0, o1, o2;
      if (lx < 0 || lx >= luma_w || ly < 0 || ly >= luma_h) {
        o0 = rgba[4 * i + 0];
        o1 = rgba[4 * i + 1];
        o2 = rgba[4 * i + 2];
      } else {
        const float Y = new_luma[(size_t)ly * luma_w + lx] * 255.0f;
        const float Cb = r * -0.1687f + g * -0.3312f + b * 0.5f;
        const float Cr = r * 0.5f + g * -0.4186f + b * -0.0813f;
        /* dot(YCbCr, row): Y*1 + Cb*c1 + Cr*c2; the zero terms add exactly nothing */
        const float R = clampf(Y + Cr * 1.4f, 0.0f, 255.0f);
        const float G = clampf((Y + Cb * -0.343f) + Cr * -0.711f, 0.0f, 255.0f);
        const float Bv = clampf(Y + Cb * 1.765f, 0.0f, 255.0f);
        o0 = (unsigned char)(unsigned int)R;
        o1 = (unsigned char)(unsigned int)G;
        o2 = (unsigned char)(unsigned int)Bv;
      }
      target[3 * i + 0] = o0;
      target[3 * i + 1] = o1;
      target[3 * i + 2] = o2;
    }
}

/* ------------------------------------------------------------------------------
 * Sequencing of the three layers
 * ---------------------------------------------------------------------------- */
typedef struct {
  int n1, n2, f1, f2, f3;
  float *w[3], *b[3];   /* parameters                          */
  float *gw[3], *gb[3]; /* accumulating gradients              */
  float *pw[3], *pb[3]; /* previous deltas (momentum state)    */
} OracleNet;

static void layer_shape(const OracleNet* net, int l, int* k, int* n, int* f) {
  /* reference: src/ConfigBasedDataPipeline.cpp:24-30 */
  if (l == 0) { *k = 1; *n = net->n1; *f = net->f1; }
  if (l == 1) { *k = net->n1; *n = net->n2; *f = net->f2; }
  if (l == 2) { *k = net->n2; *n = 1; *f = net->f3; }
}

/* reference: ConfigBasedDataPipeline::forward, src/ConfigBasedDataPipeline.cpp:200-241 */
void oracle_net_forward(const OracleNet* net, const float* in, int w, int h, int S,
                        float* out1, float* out2, float* out3) {
  const int w1 = w - net->f1 + 1, h1 = h - net->f1 + 1;
  const int w2 = w1 - net->f2 + 1, h2 = h1 - net->f2 + 1;
  oracle_forward(in, out1, net->w[0], net->b[0], 1, net->n1, net->f1, 0, w, h, S);
  oracle_forward(out1, out2, net->w[1], net->b[1], net->n1, net->n2, net->f2, 0, w1, h1, S);
  oracle_forward(out2, out3, net->w[2], net->b[2], net->n2, 1, net->f3, 1, w2, h2, S);
}

/* reference: ConfigBasedDataPipeline::backpropagate,
 * src/ConfigBasedDataPipeline.cpp:243-323: last_layer_delta -> deltas(2<-3) ->
 * deltas(1<-2) -> grads layer 3, 2, 1. */
void oracle_net_backward(OracleNet* net, const float* in, const float* gt, int w, int h,
                         int S, const float* out1, const float* out2, const float* out3,
                         float* d1, float* d2, float* d3) {
  const int w1 = w - net->f1 + 1, h1 = h - net->f1 + 1;
  const int w2 = w1 - net->f2 + 1, h2 = h1 - net->f2 + 1;
  const int w3 = w2 - net->f3 + 1, h3 = h2 - net->f3 + 1;
  oracle_last_layer_delta(gt, out3, d3, w, h, w3, h3, S);
  oracle_deltas(d3, out2, d2, net->w[2], net->n2, net->f3, 1, w2, h2, S);
  oracle_deltas(d2, out1, d1, net->w[1], net->n1, net->f2, net->n2, w1, h1, S);
  oracle_backpropagate(d3, out2, net->gw[2], net->gb[2], 1, net->n2, net->f3, w3, h3, S);
  oracle_backpropagate(d2, out1, net->gw[1], net->gb[1], net->n2, net->n1, net->f2, w2, h2, S);
  oracle_backpropagate(d1, in, net->gw[0], net->gb[0], net->n1, 1, net->f1, w1, h1, S);
}

/* reference: ConfigBasedDataPipeline::update_parameters,
 * src/ConfigBasedDataPipeline.cpp:325-361 (layer 3,2,1; lr[2],lr[1],lr[0]; then the six
 * gradient accumulators are zeroed, lines 353-358). */
void oracle_net_update(OracleNet* net, unsigned batch_size, float momentum, float decay,
                       const float* lr3) {
  for (int l = 2; l >= 0; l--) {
    int k, n, f;
    layer_shape(net, l, &k, &n, &f);
    const unsigned ws = (unsigned)(f * f * k * n), bs = (unsigned)n;
    oracle_update_params(net->w[l], net->b[l], net->gw[l], net->gb[l], net->pw[l],
                         net->pb[l], momentum, decay, lr3[l], batch_size, ws, bs);
    memset(net->gw[l], 0, sizeof(float) * ws);
    memset(net->gb[l], 0, sizeof(float) * bs);
  }
}

/* One epoch as main() drives it (src/Main_cl.cpp:161-170 with execute_batch,
 * src/ConfigBasedDataPipeline.cpp:128-195): chunks of `mini_batch` samples,
 * gradients accumulate over all chunks, ONE update with batch_size = n_samples
 * (quirk Q4).  Returns 0, or -1 when out of memory. */
int oracle_net_train_epoch(OracleNet* net, const float* inputs, const float* gts, int w,
                           int h, int n_samples, int mini_batch, float momentum,
                           float decay, const float* lr3, int do_update) {
  const int w1 = w - net->f1 + 1, h1 = h - net->f1 + 1;
  const int w2 = w1 - net->f2 + 1, h2 = h1 - net->f2 + 1;
  const int w3 = w2 - net->f3 + 1, h3 = h2 - net->f3 + 1;
  const size_t e1 = (size_t)w1 * h1 * net->n1, e2 = (size_t)w2 * h2 * net->n2,
               e3 = (size_t)w3 * h3;
  float* buf = (float*)malloc(sizeof(float) * 2 * (e1 + e2 + e3) * (size_t)mini_batch);
  if (!buf) return -1;
  float* o1 = buf;
  float* o2 = o1 + e1 * mini_batch;
  float* o3 = o2 + e2 * mini_batch;
  float* d1 = o3 + e3 * mini_batch;
  float* d2 = d1 + e1 * mini_batch;
  float* d3 = d2 + e2 * mini_batch;
  for (int i = 0; i < n_samples; i += mini_batch) {
    const int S = n_samples - i < mini_batch ? n_samples - i : mini_batch;
    const float* in = inputs + (size_t)i * w * h;
    const float* gt = gts + (size_t)i * w * h;
    oracle_net_forward(net, in, w, h, S, o1, o2, o3);
    oracle_net_backward(net, in, gt, w, h, S, o1, o2, o3, d1, d2, d3);
  }
  free(buf);
  if (do_update) oracle_net_update(net, (unsigned)n_samples, momentum, decay, lr3);
  return 0;
}

/* Validation pass: sum of squared errors over a sample set
 * (src/ConfigBasedDataPipeline.cpp:177-187). */
double oracle_net_validate(const OracleNet* net, const float* inputs, const float* gts, int w,
                           int h, int n_samples) {
  const int w1 = w - net->f1 + 1, h1 = h - net->f1 + 1;
  const int w2 = w1 - net->f2 + 1, h2 = h1 - net->f2 + 1;
  const int w3 = w2 - net->f3 + 1, h3 = h2 - net->f3 + 1;
  const size_t e1 = (size_t)w1 * h1 * net->n1, e2 = (size_t)w2 * h2 * net->n2,
               e3 = (size_t)w3 * h3;
  float* buf = (float*)malloc(sizeof(float) * (e1 + e2 + e3));
  if (!buf) return NAN;
  double acc = 0.0;
  for (int i = 0; i < n_samples; i++) {
    oracle_net_forward(net, inputs + (size_t)i * w * h, w, h, 1, buf, buf + e1, buf + e1 + e2);
    acc += oracle_squared_error(gts + (size_t)i * w * h, buf + e1 + e2, w, h, w3, h3, 1);
  }
  free(buf);
  return acc;
}
