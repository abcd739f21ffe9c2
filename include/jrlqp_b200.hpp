// jrlqp_b200.hpp — header-only C++ mirror of jrl::qp::GoldfarbIdnaniSolver / DualSolver over the C-ABI
// (jrlqp_b200.h). Same method names, argument meaning and error behaviour as the reference classes
// (include/jrl-qp/GoldfarbIdnaniSolver.h:15-33, include/jrl-qp/DualSolver.h:26-60), without the Eigen
// dependency: matrices are (pointer, leading dimension) views, exactly what an Eigen::Ref carries.
// INTEGRATION.md shows the Eigen::Ref adaptor a jrl-qp maintainer would put in front of it.
#pragma once

#include "jrlqp_b200.h"

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace jrlqp_b200
{

enum class ActivationStatus : std::int8_t
{
  INACTIVE = JRLQP_INACTIVE,
  LOWER = JRLQP_LOWER,
  UPPER = JRLQP_UPPER,
  EQUALITY = JRLQP_EQUALITY,
  LOWER_BOUND = JRLQP_LOWER_BOUND,
  UPPER_BOUND = JRLQP_UPPER_BOUND,
  FIXED = JRLQP_FIXED
};

enum class TerminationStatus : int
{
  SUCCESS = JRLQP_SUCCESS,
  INCONSISTENT_INPUT = JRLQP_INCONSISTENT_INPUT,
  NON_POS_HESSIAN = JRLQP_NON_POS_HESSIAN,
  INFEASIBLE = JRLQP_INFEASIBLE,
  MAX_ITER_REACHED = JRLQP_MAX_ITER_REACHED,
  LINEAR_DEPENDENCY_DETECTED = JRLQP_LINEAR_DEPENDENCY_DETECTED,
  OVERCONSTRAINED_PROBLEM = JRLQP_OVERCONSTRAINED_PROBLEM,
  UNKNOWN = JRLQP_UNKNOWN
};

/** include/jrl-qp/SolverOptions.h:14-88 (the log stream is not carried). */
struct SolverOptions
{
  int maxIter_ = 500;
  double bigBnd_ = 1e100;
  bool warmStart_ = false;
  std::uint32_t logFlags_ = 0;
  SolverOptions & maxIter(int m)
  {
    maxIter_ = m;
    return *this;
  }
  SolverOptions & bigBnd(double b)
  {
    bigBnd_ = b;
    return *this;
  }
  SolverOptions & warmStart(bool w)
  {
    warmStart_ = w;
    return *this;
  }
  SolverOptions & logFlags(std::uint32_t f)
  {
    logFlags_ = f;
    return *this;
  }
};

/** Column-major matrix view: what Eigen::Ref<(const) MatrixXd> carries (include/jrl-qp/defs.h:11-14). */
struct MatrixView
{
  double * data;
  int rows, cols, ld;
};
struct ConstMatrixView
{
  const double * data;
  int rows, cols, ld;
};
struct ConstVectorView
{
  const double * data;
  int size;
};

/** Batched solver: `batch` independent QPs per call, host pointers (see jrlqp_problem for the layout). */
class BatchedGoldfarbIdnaniSolver
{
public:
  BatchedGoldfarbIdnaniSolver(int nbVar, int nbCstr, bool useBounds, std::int64_t batchCapacity, int device = 0)
  : n_(nbVar), mc_(nbCstr), nb_(useBounds ? nbVar : 0)
  {
    int rc = jrlqp_create(&h_, nbVar, nbCstr, useBounds ? 1 : 0, batchCapacity, device);
    if(rc != JRLQP_OK)
    {
      std::string msg = h_ ? jrlqp_last_error(h_) : "invalid arguments";
      if(h_) jrlqp_destroy(h_);
      h_ = nullptr;
      throw std::runtime_error("jrlqp_create failed (no CPU fallback): " + msg);
    }
  }
  ~BatchedGoldfarbIdnaniSolver()
  {
    if(h_) jrlqp_destroy(h_);
  }
  BatchedGoldfarbIdnaniSolver(const BatchedGoldfarbIdnaniSolver &) = delete;
  BatchedGoldfarbIdnaniSolver & operator=(const BatchedGoldfarbIdnaniSolver &) = delete;

  void options(const SolverOptions & o)
  {
    jrlqp_options c{o.maxIter_, o.bigBnd_, o.warmStart_ ? 1 : 0, o.logFlags_};
    jrlqp_set_options(h_, &c);
  }
  /** Returns the worst TerminationStatus of the batch; throws on a CUDA/argument error. */
  TerminationStatus solve(const jrlqp_problem & pb, const jrlqp_result & res)
  {
    int rc = jrlqp_solve_batch_host(h_, &pb, &res);
    if(rc < 0) throw std::runtime_error(std::string("jrlqp_solve_batch_host: ") + jrlqp_last_error(h_));
    return static_cast<TerminationStatus>(rc);
  }
  jrlqp_solver * handle() { return h_; }
  int nbVar() const { return n_; }
  int nbCstr() const { return mc_; }
  int nbBnd() const { return nb_; }

private:
  jrlqp_solver * h_ = nullptr;
  int n_, mc_, nb_;
};

/** One QP per call: the reference's interface, a batch of one through the same kernels. */
class GoldfarbIdnaniSolver
{
public:
  GoldfarbIdnaniSolver() = default;
  GoldfarbIdnaniSolver(int nbVar, int nbCstr, bool useBounds, int device = 0) : device_(device) { resize(nbVar, nbCstr, useBounds); }
  ~GoldfarbIdnaniSolver()
  {
    if(h_) jrlqp_destroy(h_);
  }
  GoldfarbIdnaniSolver(const GoldfarbIdnaniSolver &) = delete;
  GoldfarbIdnaniSolver & operator=(const GoldfarbIdnaniSolver &) = delete;

  /** DualSolver::resize (src/DualSolver.cpp:18-24) */
  void resize(int nbVar, int nbCstr, bool useBounds)
  {
    int nb = useBounds ? nbVar : 0;
    if(h_ && nbVar == n_ && nbCstr == mc_ && nb == nb_) return;
    if(h_) jrlqp_destroy(h_);
    h_ = nullptr;
    int rc = jrlqp_create(&h_, nbVar, nbCstr, useBounds ? 1 : 0, 1, device_);
    if(rc != JRLQP_OK)
    {
      std::string msg = h_ ? jrlqp_last_error(h_) : "invalid arguments";
      if(h_) jrlqp_destroy(h_);
      h_ = nullptr;
      throw std::runtime_error("jrlqp_create failed (no CPU fallback): " + msg);
    }
    n_ = nbVar;
    mc_ = nbCstr;
    nb_ = nb;
    x_.assign(static_cast<size_t>(n_), 0.0);
    u_.assign(static_cast<size_t>(mc_ + nb_), 0.0);
    as_.assign(static_cast<size_t>(mc_ + nb_), ActivationStatus::INACTIVE);
    L_.assign(static_cast<size_t>(n_) * n_, 0.0);
    options(options_);
  }

  /** DualSolver::options (src/DualSolver.cpp:26-31) */
  void options(const SolverOptions & o)
  {
    options_ = o;
    if(h_)
    {
      jrlqp_options c{o.maxIter_, o.bigBnd_, o.warmStart_ ? 1 : 0, o.logFlags_};
      jrlqp_set_options(h_, &c);
    }
  }

  /** GoldfarbIdnaniSolver::solve (src/GoldfarbIdnaniSolver.cpp:18-54). xl.size == 0 <=> no bounds.
   * As in the reference, G is an in/out argument: its lower triangle holds the Cholesky factor afterwards. */
  TerminationStatus solve(MatrixView G,
                          ConstVectorView a,
                          ConstMatrixView C,
                          ConstVectorView bl,
                          ConstVectorView bu,
                          ConstVectorView xl,
                          ConstVectorView xu)
  {
    (void)bu;
    (void)xu;
    resize(G.rows, C.cols, xl.size > 0);
    jrlqp_problem pb{};
    pb.batch = 1;
    pb.G = G.data;
    pb.ldg = G.ld;
    pb.a = a.data;
    pb.C = mc_ ? C.data : nullptr;
    pb.ldc = mc_ ? C.ld : n_;
    pb.bl = mc_ ? bl.data : nullptr;
    pb.bu = mc_ ? bu.data : nullptr;
    pb.xl = nb_ ? xl.data : nullptr;
    pb.xu = nb_ ? xu.data : nullptr;
    jrlqp_result res{};
    int status = 0;
    res.x = x_.data();
    res.u = u_.data();
    res.f = &f_;
    res.iterations = &it_;
    res.status = &status;
    res.active_set = reinterpret_cast<std::int8_t *>(as_.data());
    res.L = L_.data();
    int rc = jrlqp_solve_batch_host(h_, &pb, &res);
    if(rc < 0) throw std::runtime_error(std::string("jrlqp_solve_batch_host: ") + jrlqp_last_error(h_));
    if(rc != JRLQP_NON_POS_HESSIAN)
      for(int j = 0; j < n_; ++j)
        for(int i = j; i < n_; ++i) G.data[i + static_cast<size_t>(j) * G.ld] = L_[static_cast<size_t>(i) + static_cast<size_t>(j) * n_];
    return static_cast<TerminationStatus>(rc);
  }

  const std::vector<double> & solution() const { return x_; }
  const std::vector<double> & multipliers() const { return u_; }
  double objectiveValue() const { return f_; }
  int iterations() const { return it_; }
  const std::vector<ActivationStatus> & activeSet() const { return as_; }
  void resetActiveSet() {} // the stable solver resets at every solve (src/GoldfarbIdnaniSolver.cpp:75)

private:
  jrlqp_solver * h_ = nullptr;
  int device_ = 0;
  int n_ = 0, mc_ = 0, nb_ = 0;
  SolverOptions options_;
  std::vector<double> x_, u_, L_;
  std::vector<ActivationStatus> as_;
  double f_ = 0;
  int it_ = 0;
};

} // namespace jrlqp_b200
