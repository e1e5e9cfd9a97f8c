// Analytic inverted double pendulum step, both precisions.
//
// The reference's I2P accelerations come from MuJoCo's mj_step (third-party).  What is implemented is the
// reference's own Lagrangian model of classic_control/auxiliary/lagrange_eqs.py:12-60 cartpole(2) (cart + two
// thin rods, relative hinge angles, inertia 1/3 m l^2 about the COM) with the COMPLETE potential energy -- the
// script omits the height of pole 1's hinge for n >= 2 (lagrange_eqs.py:45; see oracle/gen_golden_i2p.py) --
// the XML's constants (inverted_double_pendulum.xml:25,31,32,35,38,45), the forward-Euler rule of
// mujoco_env.py:91-97, the observation of inverted_double_pendulum.py:56-60 (precedence quirk replicated) and the
// reward / terminal of :84-90,114-122,150-157,185-196 through reward_terminal_kernel on that observation.
//
// A(q) a = b(q, qd, F), symmetric 3x3, solved by LDL^T in the operation order of oracle/emei_oracle.py:i2p_accel
// (float64 is compiled with -fmad=false: same bits as the oracle up to the 1-ulp difference of CUDA's sin/cos).
// One env per thread, persistent grid-stride; state [n,6] = [x, th0, th1, v, w0, w1] (qpos||qvel).
#pragma once
#include "common.cuh"

namespace emei {

template <typename R>
struct I2PConsts {
  R k_a, k_b, k_c, k_d, k_e, k_g1, g, l0, a00, gear, ctrl_low, ctrl_high, dt, pi;
  int freq_rate, swingup, action_kind;
};

template <typename R>
inline I2PConsts<R> make_i2p_consts(const emei_i2p_params& p) {
  I2PConsts<R> k;
  const R M = static_cast<R>(p.mass_cart), m0 = static_cast<R>(p.mass_pole0), m1 = static_cast<R>(p.mass_pole1);
  const R l0 = static_cast<R>(p.length0), l1 = static_cast<R>(p.length1), g = static_cast<R>(p.gravity);
  // the same expressions, in the same order, as the oracle (evaluated in R on the host: IEEE, no contraction)
  k.k_a = (m0 + R(2) * m1) * l0;
  k.k_b = m1 * l1;
  k.k_c = R(2) * m1 * l0 * l1;
  k.k_d = static_cast<R>(4.0 / 3.0) * m1 * l1 * l1;
  k.k_e = static_cast<R>(4.0 / 3.0) * m0 * l0 * l0 + R(4) * m1 * l0 * l0 + k.k_d;
  k.k_g1 = g * l0 * (m0 + R(2) * m1);
  k.g = g;
  k.l0 = l0;
  k.a00 = M + m0 + m1;
  k.gear = static_cast<R>(p.gear);
  k.ctrl_low = static_cast<R>(p.ctrl_low);
  k.ctrl_high = static_cast<R>(p.ctrl_high);
  k.dt = static_cast<R>(p.dt);
  k.pi = static_cast<R>(3.141592653589793238462643383279502884);
  k.freq_rate = p.freq_rate;
  k.swingup = (p.variant == EMEI_I2P_REBOUND_SWINGUP || p.variant == EMEI_I2P_BOUNDARY_SWINGUP) ? 1 : 0;
  k.action_kind = p.action_kind;
  return k;
}

template <typename R>
__global__ void __launch_bounds__(kBlock)
    i2p_step_kernel(const R* __restrict__ state_in, R* __restrict__ state_out, R* __restrict__ obs_out,
                    const void* __restrict__ action, int64_t n, const I2PConsts<R> k, const NoiseConsts z) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kBlock;
  const R sign = k.swingup ? R(-1) : R(1);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x; i < n; i += stride) {
    R y[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) y[j] = state_in[6 * i + j];
    R ctrl = load_ctrl<R>(action, i, k.action_kind);
    ctrl = ctrl < k.ctrl_low ? k.ctrl_low : (ctrl > k.ctrl_high ? k.ctrl_high : ctrl);  // mj_step clamps ctrl
    const R F = k.gear * ctrl;
    for (int sub = 0; sub < k.freq_rate; ++sub) {
      const R th0 = y[1], th1 = y[2], w0 = y[4], w1 = y[5];
      R s0, c0, s1, c1, s01, c01;
      sincos_r(th0, &s0, &c0);
      sincos_r(th1, &s1, &c1);
      sincos_r(th0 + th1, &s01, &c01);
      s0 = sign * s0, c0 = sign * c0, s01 = sign * s01, c01 = sign * c01;
      const R a01 = k.k_a * c0 + k.k_b * c01;
      const R a02 = k.k_b * c01;
      const R a11 = k.k_e + R(2) * k.k_c * c1;
      const R a12 = k.k_c * c1 + k.k_d;
      const R a22 = k.k_d;
      const R ws = w0 + w1;
      const R b0 = F + k.k_a * s0 * (w0 * w0) + k.k_b * s01 * (ws * ws);
      const R b1 = k.k_g1 * s0 + k.g * k.k_b * s01 + k.k_c * s1 * (w1 * (R(2) * w0 + w1));
      const R b2 = k.k_b * (k.g * s01 - R(2) * k.l0 * s1 * (w0 * w0));
      const R l10 = a01 / k.a00;
      const R l20 = a02 / k.a00;
      const R d1 = a11 - l10 * a01;
      const R t12 = a12 - l10 * a02;
      const R l21 = t12 / d1;
      const R d2 = a22 - l20 * a02 - l21 * t12;
      const R y1 = b1 - l10 * b0;
      const R y2 = b2 - l20 * b0 - l21 * y1;
      const R z2 = y2 / d2;
      const R z1 = y1 / d1 - l21 * z2;
      const R z0 = b0 / k.a00 - l10 * z1 - l20 * z2;
      // mujoco_env.py:91-97: (q, v) <- (q + v h, v + a h)
      const R q0 = y[0] + y[3] * k.dt, q1 = y[1] + y[4] * k.dt, q2 = y[2] + y[5] * k.dt;
      const R v0 = y[3] + z0 * k.dt, v1 = y[4] + z1 * k.dt, v2 = y[5] + z2 * k.dt;
      y[0] = q0, y[1] = q1, y[2] = q2, y[3] = v0, y[4] = v1, y[5] = v2;
      if (z.on)  // mujoco_env.py:98-104: Gaussian state noise after every sub-step
        add_state_noise<R, 6>(y, z, z.env_offset + static_cast<unsigned long long>(i), z.substep0 + static_cast<unsigned long long>(sub));
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) state_out[6 * i + j] = y[j];
    // inverted_double_pendulum.py:56-60: (theta + pi) % 2 * pi - pi   (sic)
    R o[6] = {y[0], py_mod(y[1] + k.pi, R(2)) * k.pi - k.pi, py_mod(y[2] + k.pi, R(2)) * k.pi - k.pi, y[3], y[4], y[5]};
#pragma unroll
    for (int j = 0; j < 6; ++j) obs_out[6 * i + j] = o[j];
  }
}

// (included at the end of kernels.cuh: reward_terminal<R>() is defined above)
template <typename R>
int i2p_step(const R* state_in, R* state_out, R* obs_out, const void* action, R* reward, uint8_t* done, double* stats,
             int64_t n, const emei_i2p_params* p, emei_stream_t stream, const emei_noise_params* noise = nullptr) {
  if (n < 0) return EMEI_ERR_BAD_SIZE;
  EMEI_CHECK_PTR(p);
  if (p->variant < EMEI_I2P_REBOUND_BALANCING || p->variant > EMEI_I2P_BOUNDARY_SWINGUP) return EMEI_ERR_BAD_VARIANT;
  if (p->action_kind < EMEI_ACTION_DISCRETE_U8 || p->action_kind > EMEI_ACTION_CONTINUOUS_F64) return EMEI_ERR_BAD_ACTION_KIND;
  if (p->freq_rate < 1 || !(p->dt > 0.0) || !(p->mass_cart > 0.0) || !(p->mass_pole0 > 0.0) || !(p->mass_pole1 > 0.0) ||
      !(p->length0 > 0.0) || !(p->length1 > 0.0))
    return EMEI_ERR_BAD_PARAM;
  if (noise != nullptr)
    for (int j = 0; j < 6; ++j)
      if (!(noise->sigma[j] >= 0.0)) return EMEI_ERR_BAD_PARAM;
  if (n == 0) return EMEI_OK;
  EMEI_CHECK_PTR(state_in);
  EMEI_CHECK_PTR(state_out);
  EMEI_CHECK_PTR(obs_out);
  EMEI_CHECK_PTR(action);
  EMEI_CHECK_PTR(reward);
  EMEI_CHECK_PTR(done);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const I2PConsts<R> k = make_i2p_consts<R>(*p);
  i2p_step_kernel<R><<<resident_grid(i2p_step_kernel<R>, n), kBlock, 0, s>>>(state_in, state_out, obs_out, action, n, k,
                                                                             make_noise_consts(noise, p->freq_rate));
  const int rc = launch_status();
  if (rc != EMEI_OK) return rc;
  emei_scoring_params sp = {};
  sp.family = p->variant;
  sp.x_left = p->x_left;
  sp.x_right = p->x_right;
  sp.dt = 1.0;  // unused by the pendulum families; validated > 0
  return reward_terminal<R>(obs_out, nullptr, reward, done, stats, nullptr, n, &sp, stream);
}

}  // namespace emei
