// gp_adam.cu -- fused unpack + descale + Adam-family update (see also gp_sgd.cu).
//
// Reference being replaced (chainer v7.8.1):
//   K3 div_by_size              chainermn/communicators/pure_nccl_communicator.py:183-189
//   K2 batched unpack           chainermn/communicators/_memory_utility.py:361-429
//   K5 momentum_sgd             chainer/optimizers/momentum_sgd.py:75-88
//        v = momentum * v - lr * grad;  param += v;            (T = param dtype)
//   K6 adam / amsgrad / adabound / amsbound    chainer/optimizers/adam.py:237-332
//        T grad_ = grad; T m_ = m; T v_ = v;
//        m_ += one_minus_beta1 * (grad_ - m_);
//        v_ += one_minus_beta2 * (grad_ * grad_ - v_);
//        [vhat_ = max(vhat_, v_); vhat = vhat_;]
//        m = m_; v = v_;
//        param -= eta * (alpha_t * m_ / (sqrt(v_|vhat_) + eps) + weight_decay_rate * param);
//        [adabound: max(min(alpha_t / (sqrt(.) + eps), upper), lower) * m_ instead]
//   (one launch per parameter in the reference; K3 and K2 are two more full
//    passes over the gradient.)
//
// Here one launch covers the whole parameter list and each mean-gradient
// element is read from HBM once (from the allreduced packed buffer), never
// materialised unless write_grad asks for param.grad to stay observable, as
// reference callers may expect (tests/chainermn_tests/optimizer_tests/
// test_multi_node_optimizer.py:57-110).
//
// Algorithmic HBM bytes per element (fp32 params, buffer itemsize b):
//   MomentumSGD  b + 8 (param r/w) + 8 (v r/w)  [+4 write_grad]
//   Adam         b + 8 + 8 (m) + 8 (v)          [+4 write_grad] [+8 vhat]
#include "gp_adam_op.cuh"

extern "C" int gp_unpack_adam(const void* buffer, int buf_dtype, const int64_t* d_csum,
                              const gp_seg_t* d_segs, int n_segs, int64_t elem_begin,
                              int64_t elem_end, double scale, double alpha_t,
                              double one_minus_beta1, double one_minus_beta2, double eps,
                              double eta, double weight_decay_rate, double lower, double upper,
                              int adam_flags, int write_grad, int layout_hint, void* stream) {
  const HookArgs none = {nullptr, 0.0, 0.0};
  return launch_adam<false>(buffer, buf_dtype, d_csum, d_segs, n_segs, elem_begin, elem_end, scale,
                            alpha_t, one_minus_beta1, one_minus_beta2, eps, eta, weight_decay_rate,
                            lower, upper, adam_flags, write_grad, layout_hint, none, stream,
                            "gp_unpack_adam");
}
