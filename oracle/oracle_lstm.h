// TEST INFRASTRUCTURE ONLY — see gmix_oracle.h.
// Restatement of the reference's online-trained LSTM byte model: one layer, 50 cells, horizon 100,
// lr 0.03, clip 10 (models/lstm-model.cpp:7). Cites are relative to /root/reference/src/models.
// Every sum keeps the reference's order: libstdc++ valarray::sum() runs ascending seeded with
// element 0, while _Expr::sum() (sum of an expression such as (a*b).sum()) runs DESCENDING seeded
// with the last element (bits/valarray_after.h) — both are mirrored below.
#ifndef ORACLE_LSTM_H_
#define ORACLE_LSTM_H_
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace oracle_lstm {

enum { CELLS = 50, HORIZON = 100, NIN = 307, ROW = 563, NOUT = 256, HID = 51, UPDATE_LIMIT = 3000 };

inline float Logistic(float x) { return 1 / (1 + expf(-x)); }

struct Gate {  // NeuronLayer, lstm-layer.h:14-27
  float error[CELLS], ivar[HORIZON], gamma[CELLS], gamma_u[CELLS], gamma_m[CELLS], gamma_v[CELLS];
  float beta[CELLS], beta_u[CELLS], beta_m[CELLS], beta_v[CELLS];
  float state[HORIZON][CELLS], norm[HORIZON][CELLS];
  float transpose[HID][CELLS];
  std::vector<float> w, update, m, v;  // [CELLS][ROW]
};

struct Lstm {
  Gate g[3];  // forget, input node, output (lstm-layer.cpp:168-173)
  std::vector<float> wout;  // [HORIZON][NOUT][HID]   (LongTermMemory::lstm_output_layer)
  float state[CELLS], state_error[CELLS], stored_error[CELLS];
  float tanh_state[HORIZON][CELLS], input_gate_state[HORIZON][CELLS], last_state[HORIZON][CELLS];
  unsigned layer_epoch;
  unsigned long long update_steps;
  unsigned input_history[HORIZON];
  float hidden[HID], hidden_error[CELLS];
  float layer_input[HORIZON][NIN];
  float output[HORIZON][NOUT];
  unsigned epoch;

  static float Rand() { return static_cast<float>(rand()) / static_cast<float>(RAND_MAX); }

  void Init() {  // lstm.cpp:8-43, lstm-layer.cpp:36-54,156-196
    for (int k = 0; k < 3; ++k) {
      Gate& G = g[k];
      memset(G.error, 0, sizeof(G.error)); memset(G.ivar, 0, sizeof(G.ivar));
      for (int i = 0; i < CELLS; ++i) G.gamma[i] = 1.0f;
      memset(G.gamma_u, 0, sizeof(G.gamma_u)); memset(G.gamma_m, 0, sizeof(G.gamma_m));
      memset(G.gamma_v, 0, sizeof(G.gamma_v)); memset(G.beta, 0, sizeof(G.beta));
      memset(G.beta_u, 0, sizeof(G.beta_u)); memset(G.beta_m, 0, sizeof(G.beta_m));
      memset(G.beta_v, 0, sizeof(G.beta_v)); memset(G.state, 0, sizeof(G.state));
      memset(G.norm, 0, sizeof(G.norm)); memset(G.transpose, 0, sizeof(G.transpose));
      G.w.assign(CELLS * ROW, 0.0f); G.update.assign(CELLS * ROW, 0.0f);
      G.m.assign(CELLS * ROW, 0.0f); G.v.assign(CELLS * ROW, 0.0f);
    }
    wout.assign((size_t)HORIZON * NOUT * HID, 0.0f);
    memset(state, 0, sizeof(state)); memset(state_error, 0, sizeof(state_error));
    memset(stored_error, 0, sizeof(stored_error)); memset(tanh_state, 0, sizeof(tanh_state));
    memset(input_gate_state, 0, sizeof(input_gate_state)); memset(last_state, 0, sizeof(last_state));
    layer_epoch = 0; update_steps = 0; epoch = 0;
    memset(input_history, 0, sizeof(input_history));
    memset(hidden, 0, sizeof(hidden)); hidden[HID - 1] = 1;
    memset(hidden_error, 0, sizeof(hidden_error));
    memset(layer_input, 0, sizeof(layer_input));
    for (int e = 0; e < HORIZON; ++e) layer_input[e][NIN - 1] = 1;
    for (int e = 0; e < HORIZON; ++e) for (int i = 0; i < NOUT; ++i) output[e][i] = (float)(1.0 / NOUT);
    // weight init, lstm-layer.cpp:176-195: gates interleaved per element, then forget bias = 1
    float val = sqrtf(6.0f / float(256 + 256));
    float low = -val;
    float range = 2 * val;
    for (int i = 0; i < CELLS; ++i) {
      for (int j = 0; j < ROW; ++j) {
        g[0].w[i * ROW + j] = low + Rand() * range;
        g[1].w[i * ROW + j] = low + Rand() * range;
        g[2].w[i * ROW + j] = low + Rand() * range;
      }
      g[0].w[i * ROW + ROW - 1] = 1;
    }
  }

  void SetInput(const float* ppm) { memcpy(layer_input[epoch], ppm, 256 * sizeof(float)); }  // lstm.cpp:45-50

  void GateForward(Gate& G, const float* in, int sym) {  // lstm-layer.cpp:222-241
    const unsigned e = layer_epoch;
    for (int i = 0; i < CELLS; ++i) {
      const float* w = &G.w[i * ROW];
      float f = w[sym];
      for (int j = 0; j < NIN; ++j) f += in[j] * w[NOUT + j];
      G.norm[e][i] = f;
    }
    float s = G.norm[e][CELLS - 1] * G.norm[e][CELLS - 1];  // _Expr::sum(): descending
    for (int i = CELLS - 2; i >= 0; --i) s += G.norm[e][i] * G.norm[e][i];
    G.ivar[e] = 1.0f / sqrtf((s / CELLS) + 1e-5f);
    for (int i = 0; i < CELLS; ++i) G.norm[e][i] *= G.ivar[e];
    for (int i = 0; i < CELLS; ++i) G.state[e][i] = G.norm[e][i] * G.gamma[i] + G.beta[i];
  }

  const float* Predict(unsigned sym) {  // lstm.cpp:91-122 + lstm-layer.cpp:198-220
    float* in = layer_input[epoch];
    memcpy(in + 256, hidden, CELLS * sizeof(float));
    const unsigned e = layer_epoch;
    memcpy(last_state[e], state, sizeof(state));
    for (int k = 0; k < 3; ++k) GateForward(g[k], in, sym);
    for (int i = 0; i < CELLS; ++i) {
      g[0].state[e][i] = Logistic(g[0].state[e][i]);
      g[1].state[e][i] = tanhf(g[1].state[e][i]);
      g[2].state[e][i] = Logistic(g[2].state[e][i]);
    }
    for (int i = 0; i < CELLS; ++i) input_gate_state[e][i] = 1.0f - g[0].state[e][i];
    for (int i = 0; i < CELLS; ++i) state[i] *= g[0].state[e][i];
    for (int i = 0; i < CELLS; ++i) state[i] += g[1].state[e][i] * input_gate_state[e][i];
    for (int i = 0; i < CELLS; ++i) tanh_state[e][i] = tanhf(state[i]);
    for (int i = 0; i < CELLS; ++i) hidden[i] = g[2].state[e][i] * tanh_state[e][i];
    if (++layer_epoch == HORIZON) layer_epoch = 0;

    const float* W = &wout[(size_t)epoch * NOUT * HID];
    float max_out = 0;
    for (int i = 0; i < NOUT; ++i) {
      float sum = 0;
      for (int j = 0; j < HID; ++j) sum += hidden[j] * W[i * HID + j];
      output[epoch][i] = sum;
      max_out = sum > max_out ? sum : max_out;  // std::max(sum, max_out)
    }
    for (int i = 0; i < NOUT; ++i) output[epoch][i] = expf(output[epoch][i] - max_out);
    float s = output[epoch][0];  // valarray::sum(): ascending
    for (int i = 1; i < NOUT; ++i) s += output[epoch][i];
    for (int i = 0; i < NOUT; ++i) output[epoch][i] /= s;
    unsigned ret = epoch;
    if (++epoch == HORIZON) epoch = 0;
    return output[ret];
  }

  static void Clip(float* a, int n) {  // lstm-layer.cpp:243-250
    for (int i = 0; i < n; ++i) { if (a[i] < -10.0f) a[i] = -10.0f; else if (a[i] > 10.0f) a[i] = 10.0f; }
  }

  // Adam, lstm-layer.cpp:12-34. The scalars depend only on t.
  struct AdamK { float alpha, d1, d2; };
  static AdamK AdamScalars(float t, float lr) {
    const float beta1 = 0.025, beta2 = 0.9999;
    const unsigned long long update_limit = UPDATE_LIMIT;
    AdamK k;
    if (t < update_limit) {
      k.alpha = lr * 0.1f / sqrt(5e-5f * t + 1.0f);
      k.d1 = (float)(1.0f - pow(beta1, t));
      k.d2 = (float)(1.0f - pow(beta2, t));
    } else {
      k.alpha = lr * 0.1f / sqrt(5e-5f * update_limit + 1.0f);
      k.d1 = (float)(1.0f - pow(beta1, update_limit));
      k.d2 = (float)(1.0f - pow(beta2, update_limit));
    }
    return k;
  }
  static void Adam(float* gr, float* m, float* v, float* w, int n, const AdamK& k) {
    const float beta1 = 0.025, beta2 = 0.9999, eps = 1e-6f;
    for (int i = 0; i < n; ++i) m[i] *= beta1;
    for (int i = 0; i < n; ++i) m[i] += (1.0f - beta1) * gr[i];
    for (int i = 0; i < n; ++i) v[i] *= beta2;
    for (int i = 0; i < n; ++i) v[i] += (1.0f - beta2) * gr[i] * gr[i];
    for (int i = 0; i < n; ++i) w[i] -= k.alpha * ((m[i] / k.d1) / (sqrtf(v[i] / k.d2 + eps)));
  }

  void GateBackward(Gate& G, const float* in, int e, int sym, const AdamK& k) {  // lstm-layer.cpp:297-354
    if (e == HORIZON - 1) {
      memset(G.gamma_u, 0, sizeof(G.gamma_u));
      memset(G.beta_u, 0, sizeof(G.beta_u));
      for (int i = 0; i < CELLS; ++i) {
        memset(&G.update[i * ROW], 0, ROW * sizeof(float));
        for (int j = 0; j < HID; ++j) G.transpose[j][i] = G.w[i * ROW + j + 512];
      }
    }
    for (int i = 0; i < CELLS; ++i) G.beta_u[i] += G.error[i];
    for (int i = 0; i < CELLS; ++i) G.gamma_u[i] += G.error[i] * G.norm[e][i];
    for (int i = 0; i < CELLS; ++i) G.error[i] *= G.gamma[i] * G.ivar[e];
    float s = G.error[CELLS - 1] * G.norm[e][CELLS - 1];  // _Expr::sum(): descending
    for (int i = CELLS - 2; i >= 0; --i) s += G.error[i] * G.norm[e][i];
    s = s / CELLS;
    for (int i = 0; i < CELLS; ++i) G.error[i] -= s * G.norm[e][i];
    if (e > 0) {
      for (int i = 0; i < CELLS; ++i) {
        float f = 0;
        for (int j = 0; j < CELLS; ++j) f += G.error[j] * G.transpose[i][j];
        stored_error[i] += f;
      }
    }
    for (int i = 0; i < CELLS; ++i) {
      float* u = &G.update[i * ROW];
      for (int j = 0; j < NIN; ++j) u[NOUT + j] += G.error[i] * in[j];
      u[sym] += G.error[i];
    }
    if (e == 0) {
      for (int i = 0; i < CELLS; ++i) Adam(&G.update[i * ROW], &G.m[i * ROW], &G.v[i * ROW], &G.w[i * ROW], ROW, k);
      Adam(G.gamma_u, G.gamma_m, G.gamma_v, G.gamma, CELLS, k);
      Adam(G.beta_u, G.beta_m, G.beta_v, G.beta, CELLS, k);
    }
  }

  void LayerBackward(const float* in, int e, int sym) {  // lstm-layer.cpp:252-295
    if (e == HORIZON - 1) {
      memcpy(stored_error, hidden_error, sizeof(stored_error));
      memset(state_error, 0, sizeof(state_error));
    } else {
      for (int i = 0; i < CELLS; ++i) stored_error[i] += hidden_error[i];
    }
    const float* F = g[0].state[e]; const float* I = g[1].state[e]; const float* O = g[2].state[e];
    for (int i = 0; i < CELLS; ++i)
      g[2].error[i] = tanh_state[e][i] * stored_error[i] * O[i] * (1.0f - O[i]);
    for (int i = 0; i < CELLS; ++i)
      state_error[i] += stored_error[i] * O[i] * (1.0f - (tanh_state[e][i] * tanh_state[e][i]));
    for (int i = 0; i < CELLS; ++i)
      g[1].error[i] = state_error[i] * input_gate_state[e][i] * (1.0f - (I[i] * I[i]));
    for (int i = 0; i < CELLS; ++i)
      g[0].error[i] = (last_state[e][i] - I[i]) * state_error[i] * F[i] * input_gate_state[e][i];
    memset(hidden_error, 0, sizeof(hidden_error));
    if (e > 0) {
      for (int i = 0; i < CELLS; ++i) state_error[i] *= F[i];
      memset(stored_error, 0, sizeof(stored_error));
    } else {
      if (update_steps < UPDATE_LIMIT) ++update_steps;
    }
    AdamK k = AdamScalars((float)update_steps, 0.03f);
    for (int q = 0; q < 3; ++q) GateBackward(g[q], in, e, sym, k);
    Clip(state_error, CELLS); Clip(stored_error, CELLS); Clip(hidden_error, CELLS);
  }

  void Perceive(unsigned input) {  // lstm.cpp:52-89
    int last = (int)epoch - 1;
    if (last == -1) last = HORIZON - 1;
    int old_input = input_history[last];
    input_history[last] = input;
    if (epoch == 0) {
      for (int e = HORIZON - 1; e >= 0; --e) {
        const float* W = &wout[(size_t)e * NOUT * HID];
        for (unsigned i = 0; i < NOUT; ++i) {
          float error = (i == input_history[e]) ? (output[e][i] - 1) : output[e][i];
          for (int j = 0; j < CELLS; ++j) hidden_error[j] += W[i * HID + j] * error;
        }
        int prev = e - 1;
        if (prev == -1) prev = HORIZON - 1;
        int sym = input_history[prev];
        if (e == 0) sym = old_input;
        LayerBackward(layer_input[e], e, sym);
      }
    }
    const float lr = 0.03f;
    const float* Wl = &wout[(size_t)last * NOUT * HID];
    float* Wc = &wout[(size_t)epoch * NOUT * HID];
    for (unsigned i = 0; i < NOUT; ++i) {
      float error = (i == input) ? (output[last][i] - 1) : output[last][i];
      float le = lr * error;
      for (int j = 0; j < HID; ++j) Wc[i * HID + j] = Wl[i * HID + j];
      for (int j = 0; j < HID; ++j) Wc[i * HID + j] -= le * hidden[j];
    }
  }
};

}  // namespace oracle_lstm
#endif
