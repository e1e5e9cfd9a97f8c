// Numerics study (development tool; NOT part of the product or of the tests' oracle).
// Compiles emei_b200/csrc/f32math.cuh for the HOST and measures, on random states, how far the
// lean float32 cart-pole step is from the reference's arithmetic (float64 derivative, float32
// increment, float64 accumulate: cartpole.py:48-60 + base_control.py:160-164) in units of the
// BASELINE.json envelope 1e-6 + 1e-5*|ref|.  Same float32 inputs on both sides.
//   nvcc -O2 -std=c++17 -o /tmp/cp_env tools/f32_study/cartpole_envelope.cu && /tmp/cp_env
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "../../emei_b200/csrc/f32math.cuh"

static void ref_step(double y[4], double force, int fr, double dt) {
  const double g = 9.8, mp = 0.1, mt = 1.1, l = 0.5, pml = mp * l;
  const float dt32 = (float)dt;
  for (int i = 0; i < fr; ++i) {
    double c = cos(y[2]), s = sin(y[2]);
    double temp = (force + pml * (y[3] * y[3]) * s) / mt;
    double tha = (g * s - c * temp) / (l * (4.0 / 3.0 - mp * (c * c) / mt));
    double xa = temp - pml * tha * c / mt;
    float d[4] = {(float)y[1], (float)xa, (float)y[3], (float)tha};
    for (int j = 0; j < 4; ++j) y[j] += (double)(d[j] * dt32);
  }
}

// the plain all-float32 restatement in the reference's evaluation order (libm sinf/cosf, IEEE divides):
// what the first version of the kernel computed; the yardstick for "what float32 can do at all".
static void plain_f32_step(float y[4], float force, int fr, float dt) {
  const float g = 9.8f, mp = 0.1f, mt = 1.1f, l = 0.5f, pml = (float)(0.1 * 0.5), ft = (float)(4.0 / 3.0);
  for (int i = 0; i < fr; ++i) {
    float c = cosf(y[2]), s = sinf(y[2]);
    float temp = (force + pml * (y[3] * y[3]) * s) / mt;
    float tha = (g * s - c * temp) / (l * (ft - mp * (c * c) / mt));
    float xa = temp - pml * tha * c / mt;
    float n0 = y[0] + y[1] * dt, n1 = y[1] + xa * dt, n2 = y[2] + y[3] * dt, n3 = y[3] + tha * dt;
    y[0] = n0; y[1] = n1; y[2] = n2; y[3] = n3;
  }
}

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 20000000;
  const int fr = argc > 2 ? atoi(argv[2]) : 4;
  const double th_scale = argc > 3 ? atof(argv[3]) : M_PI;
  const double w_scale = argc > 4 ? atof(argv[4]) : 8.0;
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> U(-1, 1);
  const double g = 9.8, mp = 0.1, mt = 1.1, l = 0.5, pml = mp * l, dt = 0.02;
  emei::f32::CartPoleK k;
  k.g = (float)g; k.kpm = (float)(pml / mt); k.inv_mt = (float)(1.0 / mt);
  k.den0 = (float)(l * 4.0 / 3.0); k.den1 = (float)(l * mp / mt); k.dt = (float)dt;
  double worst[4] = {0, 0, 0, 0}, worst_sc = 0;
  double sum2[4] = {0,0,0,0};
  double worst_plain[4] = {0, 0, 0, 0};
  double worst_sub[4] = {0, 0, 0, 0}, worst_cos = 0, worst_cos_sub = 0;
  long n_fallback = 0;
  for (long i = 0; i < n; ++i) {
    float st[4] = {(float)(U(rng) * 4), (float)(U(rng) * 5), (float)(U(rng) * th_scale), (float)(U(rng) * w_scale)};
    float a = (float)U(rng);
    float force = 10.0f * a;
    double y[4] = {st[0], st[1], st[2], st[3]};
    ref_step(y, (double)force, fr, dt);
    float x = st[0], xd = st[1], th = st[2], w = st[3];
    const float f_mt = force * k.inv_mt;
    // shipped path: one full sincos, angle-addition between sub-steps (libm per sub-step when the guard trips)
    emei::f32::LaneMax<float> dmax;
    const float cfin = emei::f32::cartpole_integrate<float, 0>(x, xd, th, w, -f_mt, 0u, k, fr, dmax);
    float cos_rew = cfin;
    if (!(fabsf(st[2]) <= emei::f32::kSinCosSaneMax && dmax.m <= emei::f32::kDeltaMax)) {
      ++n_fallback;
      x = st[0]; xd = st[1]; th = st[2]; w = st[3];
      for (int j = 0; j < fr; ++j) emei::f32::cartpole_substep<true>(x, xd, th, w, f_mt, 0u, k);
      cos_rew = cosf(th);
    }
    float o[4] = {x, xd, th, w};
    {
      const double e = fabs((double)cos_rew - cos(y[2]));
      if (e > worst_cos) worst_cos = e;
    }
    {  // previous generation: full sincos at every sub-step
      float q[4] = {st[0], st[1], st[2], st[3]};
      for (int j = 0; j < fr; ++j) emei::f32::cartpole_substep<false>(q[0], q[1], q[2], q[3], f_mt, 0u, k);
      for (int j = 0; j < 4; ++j) {
        double e = fabs((double)q[j] - y[j]) / (1e-6 + 1e-5 * fabs(y[j]));
        if (e > worst_sub[j]) worst_sub[j] = e;
      }
      const double e = fabs((double)emei::f32::cos_fast(q[2]) - cos(y[2]));
      if (e > worst_cos_sub) worst_cos_sub = e;
    }
    for (int j = 0; j < 4; ++j) {
      double e = fabs((double)o[j] - y[j]) / (1e-6 + 1e-5 * fabs(y[j]));
      if (e > worst[j]) worst[j] = e;
      sum2[j] += e*e;
    }
    float pl[4] = {st[0], st[1], st[2], st[3]};
    plain_f32_step(pl, force, fr, (float)dt);
    for (int j = 0; j < 4; ++j) {
      double e = fabs((double)pl[j] - y[j]) / (1e-6 + 1e-5 * fabs(y[j]));
      if (e > worst_plain[j]) worst_plain[j] = e;
    }
    float s, c;
    emei::f32::sincos_fast(st[2] * 30.0f, &s, &c);
    double es = fabs((double)s - sin((double)(st[2] * 30.0f))), ec = fabs((double)c - cos((double)(st[2] * 30.0f)));
    if (es > worst_sc) worst_sc = es;
    if (ec > worst_sc) worst_sc = ec;
    double ec2 = fabs((double)emei::f32::cos_fast(st[2] * 30.0f) - cos((double)(st[2] * 30.0f)));
    if (ec2 > worst_sc) worst_sc = ec2;
  }
  printf("n=%ld fr=%d th_scale=%g w_scale=%g\n", n, fr, th_scale, w_scale);
  printf("worst envelope fraction  x=%.4f xd=%.4f th=%.4f w=%.4f\n", worst[0], worst[1], worst[2], worst[3]);
  printf("sincos-per-sub-step      x=%.4f xd=%.4f th=%.4f w=%.4f\n", worst_sub[0], worst_sub[1], worst_sub[2], worst_sub[3]);
  printf("worst |cos(final theta) - ref|: %.3e (angle addition)  %.3e (cos of the stored float32 theta)   libm fallbacks: %ld\n", worst_cos, worst_cos_sub, n_fallback);
  printf("plain-f32 worst fraction x=%.4f xd=%.4f th=%.4f w=%.4f\n", worst_plain[0], worst_plain[1], worst_plain[2], worst_plain[3]);
  printf("rms envelope fraction    x=%.4f xd=%.4f th=%.4f w=%.4f\n", sqrt(sum2[0]/n), sqrt(sum2[1]/n), sqrt(sum2[2]/n), sqrt(sum2[3]/n));
  printf("worst |sincos_fast - libm| over |x|<=%g: %.3e (float32 ulp(1)=1.19e-7)\n", 30 * th_scale, worst_sc);
  return 0;
}
