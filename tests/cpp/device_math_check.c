/* Host replay of the device-only math routines of csrc/shc_math.cuh (bounded-argument sincos_, refined-seed rsqrt_):
 * the same operation sequence with fma() from libm (exact fused multiply-add), checked against long double.
 * Prints: max ulp error of sin, cos, rsqrt.  Keep the constants in sync with shc_math.cuh (the test greps for them). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void sincos_dev(double a, double* s, double* c) {
  double j = fma(a, 0.6366197723675814, 6755399441055744.0);
  int64_t bits;
  memcpy(&bits, &j, 8);
  const int q = (int)(uint32_t)bits; /* __double2loint */
  j -= 6755399441055744.0;
  double r = fma(-j, 1.5707963267948966, a);
  r = fma(-j, 6.123233995736766e-17, r);
  r = fma(-j, -1.4973849048591698e-33, r);
  const double z = r * r;
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double sr = fma(r * z, ps, r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
  const double ss = (q & 1) ? cr : sr;
  const double cc = (q & 1) ? sr : cr;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
}

static double ulp_err(double got, long double want) {
  double w = (double)want;
  if (w == 0.0) return fabs(got) > 0 ? 1e9 : 0;
  return fabs((double)((long double)got - want)) / ldexp(1.0, ilogb(w) - 52);
}

int main(void) {
  double es = 0, ec = 0, er = 0;
  srand(12345);
  for (int i = 0; i < 4000000; ++i) {
    double a = ((double)rand() / RAND_MAX - 0.5) * (i % 5 == 0 ? 2000.0 : 14.0);
    double s, c;
    sincos_dev(a, &s, &c);
    double e1 = ulp_err(s, sinl((long double)a)), e2 = ulp_err(c, cosl((long double)a));
    if (e1 > es) es = e1;
    if (e2 > ec) ec = e2;
    /* rsqrt: seed with ~2^-22 relative error (what rsqrt.approx.ftz.f64 guarantees), two FMA Newton steps */
    double x = exp(((double)rand() / RAND_MAX - 0.5) * 80.0);
    double y = (double)(float)(1.0 / sqrt(x)) * (1.0 + ((rand() & 1) ? 2.0e-7 : -2.0e-7));
    double h = 0.5 * x;
    double e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    double e3 = ulp_err(y, 1.0L / sqrtl((long double)x));
    if (e3 > er) er = e3;
  }
  printf("%.4f %.4f %.4f\n", es, ec, er);
  return 0;
}
