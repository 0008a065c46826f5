// oracle/shim/boost/numeric/odeint.hpp — TEST INFRASTRUCTURE ONLY: the two Boost.Odeint names the reference's admittance
// controller uses (admittance_controller.cpp:42-46), written from the published algorithm: the classic fourth-order
// Runge-Kutta stepper and integrate_const, which takes steps of dt while t + dt <= t_end (up to rounding) and recomputes
// t = t_start + n * dt after every step.
#ifndef SHC_SHIM_BOOST_ODEINT_HPP
#define SHC_SHIM_BOOST_ODEINT_HPP
#include <cmath>
#include <cstddef>
#include <limits>
namespace boost { namespace numeric { namespace odeint {
template <class State>
class runge_kutta4 {
 public:
  template <class System>
  void do_step(System system, State& x, double t, double dt) {
    State k1(x), k2(x), k3(x), k4(x), xt(x);
    const std::size_t n = x.size();
    const double dh = 0.5 * dt, th = t + dh;
    system(x, k1, t);
    for (std::size_t i = 0; i < n; ++i) xt[i] = x[i] + dh * k1[i];
    system(xt, k2, th);
    for (std::size_t i = 0; i < n; ++i) xt[i] = x[i] + dh * k2[i];
    system(xt, k3, th);
    for (std::size_t i = 0; i < n; ++i) xt[i] = x[i] + dt * k3[i];
    system(xt, k4, t + dt);
    const double dt6 = dt / 6.0, dt3 = dt / 3.0;
    for (std::size_t i = 0; i < n; ++i) x[i] = x[i] + dt6 * k1[i] + dt3 * k2[i] + dt3 * k3[i] + dt6 * k4[i];
  }
};
template <class Stepper, class System, class State>
std::size_t integrate_const(Stepper stepper, System system, State& x, double t0, double t1, double dt) {
  double t = t0;
  std::size_t step = 0;
  const double eps = std::numeric_limits<double>::epsilon();
  while ((t + dt) - t1 <= eps) {  // less_eq_with_sign for dt > 0
    stepper.do_step(system, x, t, dt);
    ++step;
    t = t0 + double(step) * dt;
  }
  return step;
}
}}}  // namespace boost::numeric::odeint
#endif
