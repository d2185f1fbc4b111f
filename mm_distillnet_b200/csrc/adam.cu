// Optimizer step of the distillation recipe on device (SURVEY.md 8d cfg 3: "... focal loss on pseudo-labels, Adam"):
// torch.optim.Adam as the reference constructs it (src/optimization/train_methods.py:825-833: lr, betas from the config, eps
// 1e-8, no weight decay; optimizer.step() at src/optimization/traditional.py:190) — or AdamW (:834-842) — for ALL student
// parameters in ONE launch over the flat fp32 gradient buffer that the backward leaves behind (DistillStep.flat_grad):
// the reference's optimizer walks ~360 small tensors with several ATen kernels each.
//
// The parameters stay PyTorch's own fp32 tensors; the moments live in two flat buffers with the gradient buffer's layout.
// A chunk table (tensor index, first element, length <= 1024) built once by the host maps CTAs to tensors, so every access
// is a contiguous run.  The step count lives on the device and is advanced by a second 1-thread launch, so the pair can be
// captured into a CUDA graph and replayed.  HBM-bound: 4 reads + 3 writes of 4 bytes per parameter.
#include "common.cuh"

namespace mmd {
namespace adam {

constexpr int kThreads = 256, kChunk = 1024;

struct P {
  int n_chunks;
  double lr_d, beta1_d, beta2_d;
  float lr, beta1, beta2, omb1, omb2, eps, weight_decay, decay_mul;   // omb = 1 - beta, formed in double, rounded once
  int decoupled;
  const int64_t* chunks;   // [n_chunks][3]: tensor, first element inside the tensor, length
  const int64_t* offsets;  // [n_tensors]: first element of the tensor in the flat buffers
  float* const* params;
  const float* grad;
  float* m;
  float* v;
  const int64_t* step;
};

__global__ void __launch_bounds__(kThreads) adam_kernel(const __grid_constant__ P p) {
  __shared__ float s_c[2];
  if (threadIdx.x == 0) {
    const double t = (double)(*p.step + 1);                      // state['step'] += 1 before the update
    const double bc1 = 1.0 - pow(p.beta1_d, t), bc2 = 1.0 - pow(p.beta2_d, t);
    s_c[0] = (float)(p.lr_d / bc1);                        // step_size = lr / bias_correction1
    s_c[1] = (float)sqrt(bc2);                                   // denom = sqrt(v) / sqrt(bias_correction2) + eps
  }
  __syncthreads();
  const float step_size = s_c[0], bc2s = s_c[1];
  const int64_t* c = p.chunks + 3 * (int64_t)blockIdx.x;
  const int64_t tensor = c[0], first = c[1], len = c[2];
  float* w = p.params[tensor] + first;
  const int64_t fo = p.offsets[tensor] + first;
  for (int i = threadIdx.x; i < len; i += kThreads) {
    float g = p.grad[fo + i];
    float x = w[i];
    if (p.weight_decay != 0.f) {
      if (p.decoupled) x = x * p.decay_mul;                      // AdamW: p.mul_(1 - lr * weight_decay)
      else g = g + p.weight_decay * x;                           // Adam: grad.add(p, alpha=weight_decay)
    }
    const float m = p.beta1 * p.m[fo + i] + p.omb1 * g;                      // exp_avg.mul_(b1).add_(1 - b1, grad)
    const float v = p.beta2 * p.v[fo + i] + p.omb2 * g * g;                  // exp_avg_sq.mul_(b2).addcmul_(1 - b2, grad, grad)
    p.m[fo + i] = m;
    p.v[fo + i] = v;
    const float denom = sqrtf(v) / bc2s + p.eps;
    w[i] = x - step_size * (m / denom);                                     // p.addcdiv_(-step_size, exp_avg, denom)
  }
}

__global__ void adam_advance_kernel(int64_t* step) { *step += 1; }

}  // namespace adam
}  // namespace mmd

using namespace mmd;

extern "C" size_t mmd_sizeof_adam_args(void) { return sizeof(MmdAdamArgs); }

extern "C" int mmd_adam_step(const MmdAdamArgs* a, mmd_stream_t stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  MMD_CHECK_ARG(a != nullptr, "adam: null arguments");
  MMD_CHECK_ARG(a->n_chunks >= 1 && a->chunks && a->offsets && a->params && a->grad && a->exp_avg && a->exp_avg_sq && a->step,
                "adam: null table / buffer (n_chunks=%d)", a->n_chunks);
  MMD_CHECK_ARG(a->lr >= 0.0 && a->beta1 >= 0.0 && a->beta1 < 1.0 && a->beta2 >= 0.0 && a->beta2 < 1.0 && a->eps >= 0.0,
                "adam: lr=%g betas=(%g, %g) eps=%g", a->lr, a->beta1, a->beta2, a->eps);
  adam::P p;
  p.n_chunks = a->n_chunks;
  p.lr_d = a->lr; p.beta1_d = a->beta1; p.beta2_d = a->beta2;
  p.lr = (float)a->lr; p.beta1 = (float)a->beta1; p.beta2 = (float)a->beta2;
  p.omb1 = (float)(1.0 - a->beta1); p.omb2 = (float)(1.0 - a->beta2);
  p.eps = (float)a->eps; p.weight_decay = (float)a->weight_decay;
  p.decay_mul = (float)(1.0 - a->lr * a->weight_decay);
  p.decoupled = a->decoupled_weight_decay ? 1 : 0;
  p.chunks = a->chunks; p.offsets = a->offsets; p.params = a->params; p.grad = a->grad;
  p.m = a->exp_avg; p.v = a->exp_avg_sq; p.step = a->step;
  {
    ProfScope prof(PK_ADAM, 28.0 * a->n_elements, s);
    adam::adam_kernel<<<a->n_chunks, adam::kThreads, 0, s>>>(p);
    MMD_LAUNCH_CHECK();
  }
  adam::adam_advance_kernel<<<1, 1, 0, s>>>(a->step);
  MMD_LAUNCH_CHECK();
  return 0;
}
