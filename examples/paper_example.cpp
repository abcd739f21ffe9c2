// The Goldfarb-Idnani paper example of the reference's test suite
// (tests/GoldfarbIdnaniSolverTest.cpp:51-73) through the C++ mirror of the reference interface.
//   g++ -std=c++17 -Iinclude examples/paper_example.cpp -Ljrl-qp_b200/_build -ljrlqp_b200 -Wl,-rpath,$PWD/jrl-qp_b200/_build -o paper_example
#include "jrlqp_b200.hpp"

#include <cmath>
#include <cstdio>

int main()
{
  using namespace jrlqp_b200;
  double G[4] = {4, -2, -2, 4}; // column-major 2x2
  double a[2] = {6, 0};
  double C[2] = {1, 1}; // 2 x 1: one constraint per column
  double bl[1] = {2}, bu[1] = {10};
  double xl[2] = {0, 0}, xu[2] = {10, 10};
  try
  {
    GoldfarbIdnaniSolver qp(2, 1, true);
    TerminationStatus ret = qp.solve({G, 2, 2, 2}, {a, 2}, {C, 2, 1, 2}, {bl, 1}, {bu, 1}, {xl, 2}, {xu, 2});
    std::printf("status %d iterations %d f %.17g x (%.17g, %.17g) u (%.17g, %.17g, %.17g) L (%.6f, %.6f, %.6f)\n", static_cast<int>(ret),
                qp.iterations(), qp.objectiveValue(), qp.solution()[0], qp.solution()[1], qp.multipliers()[0], qp.multipliers()[1],
                qp.multipliers()[2], G[0], G[1], G[3]);
    bool ok = ret == TerminationStatus::SUCCESS && qp.iterations() == 1 && std::fabs(qp.solution()[0] - 0.5) < 1e-14
              && std::fabs(qp.solution()[1] - 1.5) < 1e-14 && std::fabs(qp.multipliers()[0] + 5.0) < 1e-13
              && std::fabs(qp.objectiveValue() - 6.5) < 1e-13 && qp.activeSet()[0] == ActivationStatus::LOWER;
    std::puts(ok ? "OK" : "MISMATCH");
    return ok ? 0 : 1;
  }
  catch(const std::exception & e)
  {
    std::printf("error: %s\n", e.what());
    return 2;
  }
}
