// TEST INFRASTRUCTURE — NOT PRODUCT CODE. C entry points over the CPU oracle so
// that tests/ and bench.py's cpu_baseline / --impl reference legs can drive it
// through ctypes. See gi_oracle.hpp for scope and provenance.
#include "gi_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

using namespace gi_oracle;

extern "C"
{

/** Solve `batch` independent QPs with the oracle, `nthreads` host threads (one
 * solver object per thread, re-used across its QPs, as benchmarks/Solvers.cpp:513-518).
 * Every array has a per-instance element stride (0 = shared by all instances).
 * G is copied to a per-thread scratch before each solve (the reference overwrites it,
 * benchmarks restore it: benchmarks/Solvers.cpp:447-453); L_out (nullable) receives it.
 * nb is 0 (no bounds) or n. Outputs are dense: x[batch][n], u[batch][mc+nb], ...
 * Returns the worst status.
 */
static int solve_batch_impl(const signed char * as,
                            long sas,
                            int experimental,
                            int warm_start,
                            int n,
                            int mc,
                          int nb,
                          long batch,
                          const double * G,
                          long sG,
                          int ldg,
                          const double * a,
                          long sa,
                          const double * C,
                          long sC,
                          int ldc,
                          const double * bl,
                          long sbl,
                          const double * bu,
                          long sbu,
                          const double * xl,
                          long sxl,
                          const double * xu,
                          long sxu,
                          int max_iter,
                          double big_bnd,
                          double * x,
                          double * u,
                          double * f,
                          int * iters,
                          int * status,
                          signed char * act,
                          int * active_list,
                          int * nactive,
                          double * L_out,
                          double * flops,
                          double * margin,
                          int nthreads)
{
  const int m = mc + nb;
  if(nthreads < 1) nthreads = 1;
  if(nthreads > batch) nthreads = static_cast<int>(std::max<long>(1, batch));
  std::atomic<long> next(0);
  std::atomic<int> worst(0);
  const bool instr = flops != nullptr || margin != nullptr;

  auto worker = [&]()
  {
    GIOracle solver(n, mc, nb > 0);
    SolverOptions opt;
    opt.maxIter = max_iter;
    opt.bigBnd = big_bnd;
    opt.warmStart = warm_start != 0;
    solver.options(opt);
    solver.instrument(instr);
    std::vector<double> Gs(static_cast<size_t>(n) * n);
    const long chunk = 16;
    int localWorst = 0;
    for(;;)
    {
      long b0 = next.fetch_add(chunk);
      if(b0 >= batch) break;
      long b1 = std::min(batch, b0 + chunk);
      for(long b = b0; b < b1; ++b)
      {
        const double * Gb = G + b * sG;
        for(int j = 0; j < n; ++j) std::memcpy(Gs.data() + static_cast<size_t>(j) * n, Gb + static_cast<size_t>(j) * ldg, sizeof(double) * n);
        int st;
        if(experimental)
        {
          // one solver object per thread is re-used across instances: "reuse the previous active set"
          // (as == nullptr with warm start) is therefore not meaningful here and is replaced by an empty guess
          solver.resetActiveSet();
          st = solver.solveExperimental(Gs.data(), n, a + b * sa, C + b * sC, ldc, bl + b * sbl, bu + b * sbu, nb ? xl + b * sxl : nullptr,
                                        nb ? xu + b * sxu : nullptr, as ? reinterpret_cast<const int8_t *>(as + b * sas) : nullptr);
        }
        else
          st = solver.solve(Gs.data(), n, a + b * sa, C + b * sC, ldc, bl + b * sbl, bu + b * sbu, nb ? xl + b * sxl : nullptr,
                            nb ? xu + b * sxu : nullptr);
        localWorst = std::max(localWorst, st);
        if(st == NON_POS_HESSIAN || st == OVERCONSTRAINED_PROBLEM)
        {
          // The reference leaves x/u/f/active set unspecified (stale) when the factorisation fails
          // (src/DualSolver.cpp:93-94); both the oracle and the CUDA path report zeros / empty set.
          if(x) std::memset(x + b * n, 0, sizeof(double) * n);
          if(u) std::memset(u + b * m, 0, sizeof(double) * m);
          if(f) f[b] = 0;
          if(iters) iters[b] = 0;
          if(status) status[b] = st;
          if(act) std::memset(act + b * m, 0, static_cast<size_t>(m));
          if(active_list)
            for(int k = 0; k < n; ++k) active_list[b * n + k] = -1;
          if(nactive) nactive[b] = 0;
          if(L_out) std::memcpy(L_out + b * static_cast<long>(n) * n, Gs.data(), sizeof(double) * n * n);
          if(flops) flops[b] = solver.flops();
          if(margin) margin[b] = solver.minMargin();
          continue;
        }
        if(x) std::memcpy(x + b * n, solver.solution(), sizeof(double) * n);
        if(u) std::memcpy(u + b * m, solver.multipliers(), sizeof(double) * m);
        if(f) f[b] = solver.objectiveValue();
        if(iters) iters[b] = solver.iterations();
        if(status) status[b] = st;
        if(act)
        {
          const auto & as = solver.activeSet();
          for(int i = 0; i < m; ++i) act[b * m + i] = static_cast<signed char>(as[static_cast<size_t>(i)]);
        }
        const auto & al = solver.activeList();
        if(active_list)
        {
          for(int k = 0; k < n; ++k) active_list[b * n + k] = k < static_cast<int>(al.size()) ? al[static_cast<size_t>(k)] : -1;
        }
        if(nactive) nactive[b] = static_cast<int>(al.size());
        if(L_out) std::memcpy(L_out + b * static_cast<long>(n) * n, Gs.data(), sizeof(double) * n * n);
        if(flops) flops[b] = solver.flops();
        if(margin) margin[b] = solver.minMargin();
      }
    }
    int w = worst.load();
    while(localWorst > w && !worst.compare_exchange_weak(w, localWorst)) {}
  };

  if(nthreads == 1)
    worker();
  else
  {
    std::vector<std::thread> th;
    for(int t = 0; t < nthreads; ++t) th.emplace_back(worker);
    for(auto & t : th) t.join();
  }
  return worst.load();
}

#define GI_BATCH_PARAMS                                                                                                          \
  int n, int mc, int nb, long batch, const double *G, long sG, int ldg, const double *a, long sa, const double *C, long sC,     \
      int ldc, const double *bl, long sbl, const double *bu, long sbu, const double *xl, long sxl, const double *xu, long sxu,  \
      int max_iter, double big_bnd, double *x, double *u, double *f, int *iters, int *status, signed char *act,                 \
      int *active_list, int *nactive, double *L_out, double *flops, double *margin, int nthreads
#define GI_BATCH_ARGS                                                                                                        \
  n, mc, nb, batch, G, sG, ldg, a, sa, C, sC, ldc, bl, sbl, bu, sbu, xl, sxl, xu, sxu, max_iter, big_bnd, x, u, f, iters,   \
      status, act, active_list, nactive, L_out, flops, margin, nthreads

/** GoldfarbIdnaniSolver::solve over a batch (see solve_batch_impl). */
int gi_oracle_solve_batch(GI_BATCH_PARAMS)
{
  return solve_batch_impl(nullptr, 0, 0, 0, GI_BATCH_ARGS);
}

/** experimental::GoldfarbIdnaniSolver::solve over a batch: `as` (nullable) = initial active-set guess,
 * int8 x (mc + nb) per instance, `sas` elements apart; warm_start = SolverOptions::warmStart_. */
int gi_oracle_solve_batch_warm(const signed char * as, long sas, int warm_start, GI_BATCH_PARAMS)
{
  return solve_batch_impl(as, sas, 1, warm_start, GI_BATCH_ARGS);
}

/** Single solve with the per-iteration event trace (debugging aid for parity work).
 * events: up to max_events rows of 7 doubles {it, p, status, l, kind, t1, t2}. Returns status;
 * *n_events receives the number of events.
 */
int gi_oracle_solve_trace(int n,
                          int mc,
                          int nb,
                          const double * G,
                          int ldg,
                          const double * a,
                          const double * C,
                          int ldc,
                          const double * bl,
                          const double * bu,
                          const double * xl,
                          const double * xu,
                          int max_iter,
                          double big_bnd,
                          double * x,
                          double * u,
                          double * f,
                          int * iters,
                          double * events,
                          int max_events,
                          int * n_events,
                          double * J_out,
                          double * R_out)
{
  GIOracle solver(n, mc, nb > 0);
  SolverOptions opt;
  opt.maxIter = max_iter;
  opt.bigBnd = big_bnd;
  solver.options(opt);
  solver.instrument(true);
  std::vector<double> Gs(static_cast<size_t>(n) * n);
  for(int j = 0; j < n; ++j) std::memcpy(Gs.data() + static_cast<size_t>(j) * n, G + static_cast<size_t>(j) * ldg, sizeof(double) * n);
  int st = solver.solve(Gs.data(), n, a, C, ldc, bl, bu, nb ? xl : nullptr, nb ? xu : nullptr);
  const int m = mc + nb;
  if(x) std::memcpy(x, solver.solution(), sizeof(double) * n);
  if(u) std::memcpy(u, solver.multipliers(), sizeof(double) * m);
  if(f) *f = solver.objectiveValue();
  if(iters) *iters = solver.iterations();
  const auto & tr = solver.trace();
  int ne = std::min<int>(static_cast<int>(tr.size()), max_events);
  for(int e = 0; e < ne; ++e)
  {
    double * row = events + 7 * e;
    row[0] = tr[e].it;
    row[1] = tr[e].p;
    row[2] = tr[e].status;
    row[3] = tr[e].l;
    row[4] = tr[e].kind;
    row[5] = tr[e].t1;
    row[6] = tr[e].t2;
  }
  if(n_events) *n_events = static_cast<int>(tr.size());
  if(J_out) std::memcpy(J_out, solver.J(), sizeof(double) * n * n);
  if(R_out) std::memcpy(R_out, solver.R(), sizeof(double) * n * n);
  return st;
}

// ---- ActiveSet handle API (for the golden sequences of tests/ActiveSetTest.cpp:60-133) ----
void * gi_oracle_as_create(int nCstr, int nBnd)
{
  return new ActiveSet(nCstr, nBnd);
}
void gi_oracle_as_destroy(void * h)
{
  delete static_cast<ActiveSet *>(h);
}
void gi_oracle_as_activate(void * h, int idx, int status)
{
  static_cast<ActiveSet *>(h)->activate(idx, static_cast<ActivationStatus>(status));
}
void gi_oracle_as_deactivate(void * h, int activeIdx)
{
  static_cast<ActiveSet *>(h)->deactivate(activeIdx);
}
void gi_oracle_as_reset(void * h)
{
  static_cast<ActiveSet *>(h)->reset();
}
/** status_out[nbAll], active_out[nbAll] (first return-value entries valid), counters[8] =
 * {nbActiveCstr, eq, ineq, lowerIneq, upperIneq, bound, lowerBound, upperBound}. Returns nbActiveCstr. */
int gi_oracle_as_query(void * h, signed char * status_out, int * active_out, int * counters)
{
  ActiveSet * A = static_cast<ActiveSet *>(h);
  for(int i = 0; i < A->nbAll(); ++i) status_out[i] = static_cast<signed char>(A->activationStatus(i));
  for(int k = 0; k < A->nbActiveCstr(); ++k) active_out[k] = (*A)[k];
  counters[0] = A->nbActiveCstr();
  counters[1] = A->nbActiveEquality();
  counters[2] = A->nbActiveInequality();
  counters[3] = A->nbActiveLowerInequality();
  counters[4] = A->nbActiveUpperInequality();
  counters[5] = A->nbActiveBound();
  counters[6] = A->nbActiveLowerBound();
  counters[7] = A->nbActiveUpperBound();
  return A->nbActiveCstr();
}

double gi_oracle_dot4(int len, const double * a, const double * b)
{
  return dot4(len, a, 1, b, 1);
}
double gi_oracle_dot32(int len, const double * a, const double * b)
{
  return dot32(len, a, b);
}
void gi_oracle_givens(double p, double q, double * out3)
{
  makeGivens(p, q, out3[0], out3[1], out3[2]);
}

} // extern "C"
