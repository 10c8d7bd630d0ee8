// Fused multi-tensor AdamW (the optimizer the reference builds, model.py:289: torch.optim.AdamW over the trainable parameters): one
// launch updates every parameter of a group from its .grad and the two moment buffers, the step counter and the bias corrections live on
// the device, so the whole update is capturable in the training step's CUDA graph (contextflow_b200/graphed.py).  Update rule = torch's
// _single_tensor_adamw:  p *= 1 - lr wd;  m += (g - m)(1 - b1);  v = v b2 + g g (1 - b2);  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps).
#include "common.cuh"

namespace cfpp {
namespace aw {

constexpr int kChunk = 2048;                       // elements per block (256 threads x 8)

// state[0] = step count (float, as torch keeps it), state[1] = 1 - b1^t, state[2] = sqrt(1 - b2^t)
__global__ void adamw_tick_kernel(float* __restrict__ state, double beta1, double beta2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float t = state[0] + 1.f;
    state[0] = t;
    state[1] = (float)(1.0 - pow(beta1, (double)t));          // torch evaluates the bias corrections in double (python floats)
    state[2] = (float)sqrt(1.0 - pow(beta2, (double)t));
  }
}

// tensors: 4 pointers per tensor {param, grad, exp_avg, exp_avg_sq}; chunks: {tensor index, first element, element count} per block
__global__ void __launch_bounds__(256) adamw_kernel(const unsigned long long* __restrict__ tensors, const int* __restrict__ chunks,
                                                    const float* __restrict__ state, const float* __restrict__ lr_dev,
                                                    float beta1, float beta2, float w1, float w2, float eps, float weight_decay) {
  const int t = chunks[3 * blockIdx.x], first = chunks[3 * blockIdx.x + 1], count = chunks[3 * blockIdx.x + 2];
  float* __restrict__ p = reinterpret_cast<float*>(tensors[4 * t]) + first;
  const float* __restrict__ g = reinterpret_cast<const float*>(tensors[4 * t + 1]) + first;
  float* __restrict__ m = reinterpret_cast<float*>(tensors[4 * t + 2]) + first;
  float* __restrict__ v = reinterpret_cast<float*>(tensors[4 * t + 3]) + first;
  const float lr = lr_dev[0];
  const float decay = 1.f - lr * weight_decay, step_size = lr / state[1], bc2 = state[2];
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const float gi = g[i];
    const float pi = p[i] * decay;
    const float mi = m[i] + (gi - m[i]) * w1;
    const float vi = v[i] * beta2 + gi * gi * w2;
    m[i] = mi; v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) / bc2 + eps));
  }
}

}  // namespace aw
}  // namespace cfpp
using namespace cfpp;

extern "C" int cfpp_adamw_chunk(void) { return aw::kChunk; }

extern "C" int cfpp_adamw_step(const void* tensors, const int* chunks, int n_chunks, float* state, const float* lr, double beta1, double beta2,
                               double eps, double weight_decay, void* stream) {
  CFPP_REQUIRE(state && lr && (n_chunks == 0 || (tensors && chunks)), "adamw_step: null table");
  aw::adamw_tick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(state, beta1, beta2);
  int rc = check_launch("adamw_tick");
  if (rc != CFPP_OK || n_chunks <= 0) return rc;
  aw::adamw_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>((const unsigned long long*)tensors, chunks, state, lr, (float)beta1, (float)beta2,
                                                                     (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (float)weight_decay);
  return check_launch("adamw_step");
}
