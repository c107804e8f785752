// Plan-level C ABI (ou_plan_create / ou_condition_forward / ou_score_step): a recorded list of
// the launches of one network evaluation (ScoreNetwork.forward, or ConditionerNetwork.forward with its mel front-end) with every device pointer resolved, replayed by ONE call per
// evaluation.  The per-evaluation inputs (signal, FiLM row, EDM input scale, update coefficients, noise,
// outputs) arrive in ou_step_args.  The host lowers the network once (fold weights, pack, plan buffers:
// open_universe_b200/engine/program.py), records the plan, and from then on a step costs one foreign call
// instead of ~34 -- from Python, or from a C / C++ host that has no Python at all.
#include <new>
#include <vector>

#include "common.cuh"

namespace ou {

enum PlanOpKind { OP_CONV, OP_TRUNK, OP_INPUT, OP_OUTPUT, OP_GRU, OP_MEL };

struct PlanOp {
  PlanOpKind kind;
  int32_t film_off;          // conv / trunk: column of gamma in the FiLM row (beta follows), or -1
  ou_conv_params conv;
  ou_trunk_params trunk;
  // input conv / output conv / GRU arguments
  const float *w, *bias_v, *b_hh;
  const void *src, *add;
  void* out;
  float bias_s, scale;
  int32_t batch, t, c, k, t_out, hidden, use_in_scale;
  // mel front-end (ou_mel_power + ou_mel_finalize): tables and scratch
  const float *window, *fb, *dft;
  float *power, *mel, *energy;
  int32_t n_fft, hop, n_mels, pad_left, frames;
};

}  // namespace ou

struct ou_plan {
  std::vector<ou::PlanOp> ops;
};

extern "C" int ou_plan_create(ou_plan** plan) {
  OU_REQUIRE(plan != nullptr, "ou_plan_create: null output");
  *plan = new (std::nothrow) ou_plan();
  if (*plan == nullptr) {
    ou::set_error("ou_plan_create: out of host memory");
    return OU_ERR_CUDA;
  }
  return OU_OK;
}

extern "C" int ou_plan_destroy(ou_plan* plan) {
  delete plan;
  return OU_OK;
}

extern "C" int ou_plan_size(const ou_plan* plan) { return plan ? (int)plan->ops.size() : 0; }

extern "C" int ou_plan_add_conv(ou_plan* plan, const ou_conv_params* p, int32_t film_off) {
  OU_REQUIRE(plan && p, "ou_plan_add_conv: null argument");
  ou::PlanOp op{};
  op.kind = ou::OP_CONV, op.film_off = film_off, op.conv = *p;
  plan->ops.push_back(op);
  return OU_OK;
}

extern "C" int ou_plan_add_trunk(ou_plan* plan, const ou_trunk_params* p, int32_t film_off) {
  OU_REQUIRE(plan && p, "ou_plan_add_trunk: null argument");
  ou::PlanOp op{};
  op.kind = ou::OP_TRUNK, op.film_off = film_off, op.trunk = *p;
  plan->ops.push_back(op);
  return OU_OK;
}

extern "C" int ou_plan_add_input_conv(ou_plan* plan, const float* w, const float* bias, void* out, int batch,
                                      int t, int cout, int k, int use_in_scale) {
  OU_REQUIRE(plan && w && out, "ou_plan_add_input_conv: null argument");
  ou::PlanOp op{};
  op.kind = ou::OP_INPUT, op.film_off = -1, op.w = w, op.bias_v = bias, op.out = out;
  op.batch = batch, op.t = t, op.c = cout, op.k = k, op.use_in_scale = use_in_scale;
  plan->ops.push_back(op);
  return OU_OK;
}

extern "C" int ou_plan_add_output_sde(ou_plan* plan, const void* src, const float* w, float bias, int batch,
                                      int cin, int k, int t_src, int t_sig) {
  OU_REQUIRE(plan && src && w, "ou_plan_add_output_sde: null argument");
  ou::PlanOp op{};
  op.kind = ou::OP_OUTPUT, op.film_off = -1, op.src = src, op.w = w, op.bias_s = bias;
  op.batch = batch, op.c = cin, op.k = k, op.t = t_src, op.t_out = t_sig;
  plan->ops.push_back(op);
  return OU_OK;
}

extern "C" int ou_plan_add_gru(ou_plan* plan, const float* gx, const float* w_hh, const float* b_hh,
                               const void* add, float scale, void* out, int batch, int t, int hidden,
                               int cluster_ctas) {
  OU_REQUIRE(plan && gx && w_hh && b_hh && out, "ou_plan_add_gru: null argument");
  ou::PlanOp op{};
  op.kind = ou::OP_GRU, op.film_off = -1, op.src = gx, op.w = w_hh, op.b_hh = b_hh, op.add = add;
  op.scale = scale, op.out = out, op.batch = batch, op.t = t, op.hidden = hidden, op.k = cluster_ctas;
  plan->ops.push_back(op);
  return OU_OK;
}

extern "C" int ou_plan_add_mel(ou_plan* plan, const float* window, const float* fb, const float* dft, float* power,
                               float* mel, float* energy, void* mel_blocked, int batch, int t, int n_fft, int hop,
                               int n_mels, int pad_left, int frames) {
  OU_REQUIRE(plan && window && fb && dft && power && mel && energy && mel_blocked, "ou_plan_add_mel: null argument");
  ou::PlanOp op{};
  op.kind = ou::OP_MEL, op.film_off = -1;
  op.window = window, op.fb = fb, op.dft = dft, op.power = power, op.mel = mel, op.energy = energy, op.out = mel_blocked;
  op.batch = batch, op.t = t, op.n_fft = n_fft, op.hop = hop, op.n_mels = n_mels, op.pad_left = pad_left;
  op.frames = frames;
  plan->ops.push_back(op);
  return OU_OK;
}

extern "C" int ou_plan_run(const ou_plan* plan, const ou_step_args* args, int first, int count, void* stream) {
  OU_REQUIRE(plan && args, "ou_plan_run: null argument");
  const int n = (int)plan->ops.size();
  OU_REQUIRE(first >= 0 && first <= n, "ou_plan_run: first op out of range");
  const int last = count < 0 ? n : (first + count < n ? first + count : n);
  for (int i = first; i < last; i++) {
    const ou::PlanOp& op = plan->ops[i];
    int rc = OU_OK;
    switch (op.kind) {
      case ou::OP_CONV: {
        ou_conv_params p = op.conv;
        if (op.film_off >= 0) {
          OU_REQUIRE(args->film != nullptr, "ou_plan_run: op %d needs a FiLM row", i);
          p.gamma = args->film + op.film_off;
          p.beta = p.gamma + p.cout;
          p.film_bstride = args->film_bstride;
        }
        rc = ou_conv1d(&p, stream);
        break;
      }
      case ou::OP_TRUNK: {
        ou_trunk_params p = op.trunk;
        if (op.film_off >= 0) {
          OU_REQUIRE(args->film != nullptr, "ou_plan_run: op %d needs a FiLM row", i);
          p.gamma = args->film + op.film_off;
          p.beta = p.gamma + p.channels;
          p.film_bstride = args->film_bstride;
        }
        if (p.out_w != nullptr) {   // fused output conv + EDM / SDE update: per-evaluation arguments
          p.out_coef = args->coef, p.out_x = args->x, p.out_noise = args->noise;
          p.out_xout = args->xout, p.out_net = args->net_out;
        }
        rc = ou_conv_trunk(&p, stream);
        break;
      }
      case ou::OP_INPUT:
        OU_REQUIRE(args->x != nullptr, "ou_plan_run: op %d needs the signal x", i);
        rc = ou_input_conv(args->x, op.w, op.bias_v, op.use_in_scale ? args->in_scale : nullptr, op.out, op.batch,
                           op.t, op.c, op.k, stream);
        break;
      case ou::OP_OUTPUT:
        rc = ou_output_sde(op.src, op.w, op.bias_s, args->coef, args->x, args->noise, args->xout, args->net_out,
                           op.batch, op.c, op.k, op.t, op.t_out, stream);
        break;
      case ou::OP_MEL: {
        const float* wav = args->x_wav ? args->x_wav : args->x;
        OU_REQUIRE(wav != nullptr, "ou_plan_run: op %d needs the waveform (x_wav or x)", i);
        rc = ou_mel_power(wav, op.window, op.fb, op.dft, op.power, op.mel, op.energy, op.batch, op.t, op.n_fft,
                          op.hop, op.n_mels, op.pad_left, op.frames, stream);
        if (rc == OU_OK)
          rc = ou_mel_finalize(op.mel, op.energy, op.mel, op.out, op.batch, op.n_mels, op.frames, stream);
        break;
      }
      case ou::OP_GRU:
        rc = ou_gru_bidir_ex((const float*)op.src, op.w, op.b_hh, op.add, op.scale, op.out, op.batch, op.t,
                             op.hidden, op.k, stream);
        break;
    }
    if (rc != OU_OK) return rc;
  }
  return OU_OK;
}
