// Seeded batch generator of random strictly-convex QPs with a planted solution.
//
// Host-side test support (the counterpart of the reference's test-support library that is
// compiled into libjrl-qp: src/test/randomProblems.cpp:15-251, include/jrl-qp/test/randomMatrices.h,
// src/test/problems.cpp:9-37, include/jrl-qp/test/problems.h:110-115). It produces the synthetic
// inputs for the parity tests and for bench.py; it is not on the solve path.
//
// What is restated, step by step (same numbering as randomProblems.cpp):
//   1. A (nObj = nVar, full rank => i.i.d. N(0,1), randomMatrices.h:157,200-210) and the matrix Ca of
//      strongly active constraint normals (equalities, strongly active inequalities, then the active
//      bounds as rows of the identity on the first variables); a vector [u; v] with
//      [A^T Ca^T] [u; v] = 0 from a column-pivoted Householder QR: v ~ U[-1,1] on the columns pivoted
//      last, u = -R1^-1 R2 v (randomProblems.cpp:49-72).
//   2. single-sided inequalities: multipliers made non-negative by flipping rows (:98-111).
//   3./4. inactive (and weakly active) constraint normals (:113-160).
//   5. planted x ~ U[-1,1]^n; b = A x - u_A; bounds and slacks |U[-1,1]| placed so that the KKT
//      conditions hold at x (:162-225).
//   6. Fisher-Yates shuffles of the inequality rows and of the variables (:227-248).
//   QP form: G = A^T A, a = -A^T b, C = [E; C] with equalities FIRST, l = [f; l], u = [f; u].
// Deviations (documented in DESIGN.md): the reference draws from an unseeded std::mt19937 /
// std::rand (3rd-party/effolkronium/random.hpp:118-130); here every instance k uses a counter-based
// stream seeded with (seed + k), so batches are reproducible and shardable across ranks. Weakly active
// double-sided inequalities set l = u_old consistently (the reference leaves l on the wrong side,
// :190-194; it never exercises that branch). nSharedRank = 0 and strictlyFeasible = false only.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace
{

struct Rng
{
  std::uint64_t s[4];
  bool haveSpare = false;
  double spare = 0;
  static std::uint64_t splitmix(std::uint64_t & x)
  {
    std::uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  explicit Rng(std::uint64_t seed)
  {
    for(auto & v : s) v = splitmix(seed);
  }
  static std::uint64_t rotl(std::uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  std::uint64_t next()
  { // xoshiro256**
    std::uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return r;
  }
  double unit() { return static_cast<double>(next() >> 11) * 0x1.0p-53; } // [0,1)
  double uniform() { return 2.0 * unit() - 1.0; } // [-1,1)
  double normal()
  {
    if(haveSpare)
    {
      haveSpare = false;
      return spare;
    }
    double u1, u2;
    do
    {
      u1 = unit();
    } while(u1 <= 0.0);
    u2 = unit();
    double rad = std::sqrt(-2.0 * std::log(u1));
    double ang = 6.283185307179586476925 * u2;
    spare = rad * std::sin(ang);
    haveSpare = true;
    return rad * std::cos(ang);
  }
  int randint(int lo, int hi) // inclusive
  {
    std::uint64_t span = static_cast<std::uint64_t>(hi - lo) + 1;
    return lo + static_cast<int>(next() % span);
  }
};

struct Characs
{
  int nVar, nEq, nIneq, nStrongActIneq, nWeakActIneq, nStrongActBounds, nWeakActBounds;
  bool bounds, doubleSided;
};

// Row-major dense helper
struct Mat
{
  int r = 0, c = 0;
  std::vector<double> v;
  Mat() = default;
  Mat(int r_, int c_) : r(r_), c(c_), v(static_cast<size_t>(r_) * c_, 0.0) {}
  double & operator()(int i, int j) { return v[static_cast<size_t>(i) * c + j]; }
  double operator()(int i, int j) const { return v[static_cast<size_t>(i) * c + j]; }
  double * row(int i) { return v.data() + static_cast<size_t>(i) * c; }
  const double * row(int i) const { return v.data() + static_cast<size_t>(i) * c; }
};

// Haar-distributed orthogonal matrix (randomMatrices.h:62-127 builds it from Householder
// reflections of random unit vectors; the distribution is the same as the Q factor, with positive
// diagonal of R, of a Gaussian matrix, which is what is computed here).
Mat randOrtho(int k, Rng & g)
{
  Mat Q(k, k);
  for(auto & e : Q.v) e = g.normal();
  // modified Gram-Schmidt on rows, twice for orthogonality
  for(int i = 0; i < k; ++i)
  {
    for(int pass = 0; pass < 2; ++pass)
      for(int j = 0; j < i; ++j)
      {
        double dot = 0;
        for(int t = 0; t < k; ++t) dot += Q(i, t) * Q(j, t);
        for(int t = 0; t < k; ++t) Q(i, t) -= dot * Q(j, t);
      }
    double nrm = 0;
    for(int t = 0; t < k; ++t) nrm += Q(i, t) * Q(i, t);
    nrm = std::sqrt(nrm);
    for(int t = 0; t < k; ++t) Q(i, t) /= nrm;
  }
  return Q;
}

// Solve [A^T Ca^T] mult = 0 as in randomProblems.cpp:49-72 (column-pivoted Householder QR).
// M is n x cols (row-major), overwritten. mult has size cols.
void nullVector(Mat & M, std::vector<double> & mult, Rng & g)
{
  const int n = M.r, cols = M.c;
  std::vector<int> perm(static_cast<size_t>(cols));
  for(int j = 0; j < cols; ++j) perm[static_cast<size_t>(j)] = j;
  std::vector<double> hv(static_cast<size_t>(n));
  for(int k = 0; k < n; ++k)
  {
    // pivot: remaining column with the largest norm
    int best = k;
    double bestNorm = -1;
    for(int j = k; j < cols; ++j)
    {
      double s = 0;
      for(int i = k; i < n; ++i) s += M(i, j) * M(i, j);
      if(s > bestNorm)
      {
        bestNorm = s;
        best = j;
      }
    }
    if(best != k)
    {
      for(int i = 0; i < n; ++i) std::swap(M(i, k), M(i, best));
      std::swap(perm[static_cast<size_t>(k)], perm[static_cast<size_t>(best)]);
    }
    // Householder on column k, rows k..n-1
    double alpha = M(k, k);
    double tail = 0;
    for(int i = k + 1; i < n; ++i) tail += M(i, k) * M(i, k);
    if(tail == 0.0) continue;
    double beta = -std::copysign(std::sqrt(alpha * alpha + tail), alpha);
    double tau = (beta - alpha) / beta;
    double inv = 1.0 / (alpha - beta);
    hv[static_cast<size_t>(k)] = 1.0;
    for(int i = k + 1; i < n; ++i) hv[static_cast<size_t>(i)] = M(i, k) * inv;
    M(k, k) = beta;
    for(int i = k + 1; i < n; ++i) M(i, k) = 0;
    for(int j = k + 1; j < cols; ++j)
    {
      double s = 0;
      for(int i = k; i < n; ++i) s += hv[static_cast<size_t>(i)] * M(i, j);
      s *= tau;
      for(int i = k; i < n; ++i) M(i, j) -= s * hv[static_cast<size_t>(i)];
    }
  }
  // v random on the trailing (cols - n) pivoted columns; u = -R1^-1 R2 v
  const int nv = cols - n;
  std::vector<double> red(static_cast<size_t>(cols));
  for(int j = 0; j < nv; ++j) red[static_cast<size_t>(n + j)] = g.uniform();
  for(int i = 0; i < n; ++i)
  {
    double s = 0;
    for(int j = 0; j < nv; ++j) s += M(i, n + j) * red[static_cast<size_t>(n + j)];
    red[static_cast<size_t>(i)] = -s;
  }
  for(int i = n - 1; i >= 0; --i)
  {
    double s = red[static_cast<size_t>(i)];
    for(int j = i + 1; j < n; ++j) s -= M(i, j) * red[static_cast<size_t>(j)];
    red[static_cast<size_t>(i)] = s / M(i, i);
  }
  mult.assign(static_cast<size_t>(cols), 0.0);
  for(int j = 0; j < cols; ++j) mult[static_cast<size_t>(perm[static_cast<size_t>(j)])] = red[static_cast<size_t>(j)];
}

struct Out
{
  double *G, *a, *C, *bl, *bu, *xl, *xu, *x, *lambda;
};

void generateOne(const Characs & ch, std::uint64_t seed, const Out & o)
{
  Rng g(seed);
  const int n = ch.nVar, nEq = ch.nEq, nIneq = ch.nIneq;
  const int nsi = ch.nStrongActIneq, nwi = ch.nWeakActIneq, nsb = ch.nStrongActBounds, nwb = ch.nWeakActBounds;
  const int nstrong = nEq + nsi + nsb;
  const double inf = std::numeric_limits<double>::infinity();

  // 1 -
  Mat A(n, n);
  for(auto & e : A.v) e = g.normal();
  Mat Ca(nstrong, n);
  for(auto & e : Ca.v) e = g.normal();
  for(int i = 0; i < nsb; ++i)
  {
    double * row = Ca.row(nEq + nsi + i);
    for(int j = 0; j < n; ++j) row[j] = (i == j) ? 1.0 : 0.0;
  }
  std::vector<double> mult(static_cast<size_t>(n + nstrong), 0.0);
  if(nstrong > 0)
  {
    Mat M(n, n + nstrong);
    for(int i = 0; i < n; ++i)
    {
      for(int j = 0; j < n; ++j) M(i, j) = A(j, i);
      for(int j = 0; j < nstrong; ++j) M(i, n + j) = Ca(j, i);
    }
    nullVector(M, mult, g);
  }

  // 2 -
  if(!ch.doubleSided)
  {
    for(int i = 0; i < nsi; ++i)
    {
      double & mu = mult[static_cast<size_t>(n + nEq + i)];
      if(mu < 0)
      {
        mu = -mu;
        double * row = Ca.row(nEq + i);
        for(int j = 0; j < n; ++j) row[j] = -row[j];
      }
    }
  }

  // 3 -
  std::vector<double> x(static_cast<size_t>(n));
  Mat E(nEq, n), Cm(nIneq, n);
  std::vector<double> fE(static_cast<size_t>(nEq)), l(static_cast<size_t>(nIneq)), u(static_cast<size_t>(nIneq));
  std::vector<double> lamIneq(static_cast<size_t>(nIneq), 0.0), lamBnd(static_cast<size_t>(n), 0.0);
  for(int i = 0; i < nEq; ++i) std::memcpy(E.row(i), Ca.row(i), sizeof(double) * n);
  for(int i = 0; i < nsi; ++i)
  {
    std::memcpy(Cm.row(i), Ca.row(nEq + i), sizeof(double) * n);
    lamIneq[static_cast<size_t>(i)] = mult[static_cast<size_t>(n + nEq + i)];
  }
  for(int i = 0; i < nsb; ++i) lamBnd[static_cast<size_t>(i)] = mult[static_cast<size_t>(n + nEq + nsi + i)];

  // 4 -
  if(nwi > 0 && nstrong > 0)
  {
    int k = std::max(nwi, nstrong);
    Mat Q = randOrtho(k, g);
    // Q1 = Q.topRows(nwi) (k = nstrong) or Q.leftCols(nstrong) (k = nwi): nwi x nstrong either way
    for(int i = 0; i < nwi; ++i)
      for(int j = 0; j < n; ++j)
      {
        double s = 0;
        for(int t = 0; t < nstrong; ++t) s += Q(i, t) * Ca(t, j);
        Cm(nsi + i, j) = s;
      }
  }
  else
  {
    for(int i = 0; i < nwi; ++i)
      for(int j = 0; j < n; ++j) Cm(nsi + i, j) = g.normal();
  }
  for(int i = nsi + nwi; i < nIneq; ++i)
    for(int j = 0; j < n; ++j) Cm(i, j) = g.normal();

  // 5 -
  for(auto & e : x) e = g.uniform();
  std::vector<double> b(static_cast<size_t>(n));
  for(int i = 0; i < n; ++i)
  {
    double s = 0;
    for(int j = 0; j < n; ++j) s += A(i, j) * x[static_cast<size_t>(j)];
    b[static_cast<size_t>(i)] = s - mult[static_cast<size_t>(i)];
  }
  auto rowDot = [&](const double * row)
  {
    double s = 0;
    for(int j = 0; j < n; ++j) s += row[j] * x[static_cast<size_t>(j)];
    return s;
  };
  for(int i = 0; i < nEq; ++i) fE[static_cast<size_t>(i)] = rowDot(E.row(i));
  for(int i = 0; i < nIneq; ++i) u[static_cast<size_t>(i)] = rowDot(Cm.row(i));
  if(ch.doubleSided)
  {
    l = u;
    std::vector<double> rl(static_cast<size_t>(nIneq)), ru(static_cast<size_t>(nIneq));
    for(auto & e : rl) e = std::abs(g.uniform());
    for(auto & e : ru) e = std::abs(g.uniform());
    for(int i = 0; i < nsi; ++i)
    {
      if(lamIneq[static_cast<size_t>(i)] > 0)
        l[static_cast<size_t>(i)] -= rl[static_cast<size_t>(i)];
      else
        u[static_cast<size_t>(i)] += ru[static_cast<size_t>(i)];
    }
    for(int i = nsi; i < nsi + nwi; ++i)
    {
      size_t k = static_cast<size_t>(i);
      if(rl[k] > ru[k])
        l[k] -= rl[k]; // active at its upper bound
      else
      { // flip the row: active at its lower bound
        double * row = Cm.row(i);
        for(int j = 0; j < n; ++j) row[j] = -row[j];
        l[k] = -u[k];
        u[k] = l[k] + ru[k];
      }
    }
    for(int i = nsi + nwi; i < nIneq; ++i)
    {
      l[static_cast<size_t>(i)] -= rl[static_cast<size_t>(i)];
      u[static_cast<size_t>(i)] += ru[static_cast<size_t>(i)];
    }
  }
  else
  {
    for(auto & e : l) e = -inf;
    for(int i = nsi + nwi; i < nIneq; ++i) u[static_cast<size_t>(i)] += std::abs(g.uniform());
  }
  std::vector<double> xl, xu;
  if(ch.bounds)
  {
    std::vector<double> r(static_cast<size_t>(n));
    for(auto & e : r) e = g.uniform();
    xl = x;
    xu = x;
    for(int i = 0; i < nsb; ++i)
    {
      size_t k = static_cast<size_t>(i);
      if(lamBnd[k] > 0)
        xl[k] -= std::abs(r[k]);
      else
        xu[k] += std::abs(r[k]);
    }
    for(int i = nsb; i < nsb + nwb; ++i)
    {
      size_t k = static_cast<size_t>(i);
      if(r[k] > 0)
        xl[k] -= r[k];
      else
        xu[k] -= r[k];
    }
    for(int i = nsb + nwb; i < n; ++i) xl[static_cast<size_t>(i)] -= std::abs(g.uniform());
    for(int i = nsb + nwb; i < n; ++i) xu[static_cast<size_t>(i)] += std::abs(g.uniform());
  }

  // 6 -
  for(int i = nIneq - 1; i > 0; --i)
  {
    int j = g.randint(0, i);
    if(i == j) continue;
    std::swap_ranges(Cm.row(i), Cm.row(i) + n, Cm.row(j));
    std::swap(u[static_cast<size_t>(i)], u[static_cast<size_t>(j)]);
    std::swap(l[static_cast<size_t>(i)], l[static_cast<size_t>(j)]);
    std::swap(lamIneq[static_cast<size_t>(i)], lamIneq[static_cast<size_t>(j)]);
  }
  if(ch.bounds)
  {
    for(int i = n - 1; i > 0; --i)
    {
      int j = g.randint(0, i);
      if(i == j) continue;
      for(int r = 0; r < n; ++r) std::swap(A(r, i), A(r, j));
      for(int r = 0; r < nIneq; ++r) std::swap(Cm(r, i), Cm(r, j));
      for(int r = 0; r < nEq; ++r) std::swap(E(r, i), E(r, j));
      std::swap(xl[static_cast<size_t>(i)], xl[static_cast<size_t>(j)]);
      std::swap(xu[static_cast<size_t>(i)], xu[static_cast<size_t>(j)]);
      std::swap(lamBnd[static_cast<size_t>(i)], lamBnd[static_cast<size_t>(j)]);
      std::swap(x[static_cast<size_t>(i)], x[static_cast<size_t>(j)]);
    }
  }

  // QP form: G = A^T A (exactly symmetric), a = -A^T b  (problems.h:110-115)
  for(int i = 0; i < n; ++i)
    for(int j = 0; j <= i; ++j)
    {
      double s = 0;
      for(int r = 0; r < n; ++r) s += A(r, i) * A(r, j);
      o.G[static_cast<size_t>(i) * n + j] = s;
      o.G[static_cast<size_t>(j) * n + i] = s;
    }
  for(int i = 0; i < n; ++i)
  {
    double s = 0;
    for(int r = 0; r < n; ++r) s += A(r, i) * b[static_cast<size_t>(r)];
    o.a[i] = -s;
  }
  const int mc = nEq + nIneq;
  for(int i = 0; i < nEq; ++i)
  {
    std::memcpy(o.C + static_cast<size_t>(i) * n, E.row(i), sizeof(double) * n);
    o.bl[i] = fE[static_cast<size_t>(i)];
    o.bu[i] = fE[static_cast<size_t>(i)];
  }
  for(int i = 0; i < nIneq; ++i)
  {
    std::memcpy(o.C + static_cast<size_t>(nEq + i) * n, Cm.row(i), sizeof(double) * n);
    o.bl[nEq + i] = l[static_cast<size_t>(i)];
    o.bu[nEq + i] = u[static_cast<size_t>(i)];
  }
  if(ch.bounds)
  {
    std::memcpy(o.xl, xl.data(), sizeof(double) * n);
    std::memcpy(o.xu, xu.data(), sizeof(double) * n);
  }
  if(o.x) std::memcpy(o.x, x.data(), sizeof(double) * n);
  if(o.lambda)
  {
    for(int i = 0; i < nEq; ++i) o.lambda[i] = mult[static_cast<size_t>(n + i)];
    for(int i = 0; i < nIneq; ++i) o.lambda[nEq + i] = lamIneq[static_cast<size_t>(i)];
    if(ch.bounds)
      for(int i = 0; i < n; ++i) o.lambda[mc + i] = lamBnd[static_cast<size_t>(i)];
  }
}

} // namespace

extern "C"
{

/** Generate `batch` problems; instance k uses the stream seeded with seed + first_index + k.
 * Dense outputs (row-major over the batch): G[batch][n][n] (symmetric), a[batch][n],
 * C[batch][mc][n] with mc = nEq + nIneq (row i = normal of constraint i, i.e. column i of the
 * reference's n x mc column-major C; equalities first), bl/bu[batch][mc], xl/xu[batch][n] (only if
 * bounds), planted x[batch][n] and multipliers lambda[batch][mc + (bounds ? n : 0)] (nullable).
 * Returns 0, or -1 if the characteristics are inconsistent (ProblemCharacteristics::check,
 * randomProblems.cpp:253-266).
 */
int jrlqp_ts_random_problems(int nVar,
                             int nEq,
                             int nIneq,
                             int nStrongActIneq,
                             int nWeakActIneq,
                             int nStrongActBounds,
                             int nWeakActBounds,
                             int bounds,
                             int doubleSided,
                             unsigned long long seed,
                             long first_index,
                             long batch,
                             double * G,
                             double * a,
                             double * C,
                             double * bl,
                             double * bu,
                             double * xl,
                             double * xu,
                             double * x,
                             double * lambda,
                             int nthreads)
{
  Characs ch{nVar, nEq, nIneq, nStrongActIneq, nWeakActIneq, nStrongActBounds, nWeakActBounds, bounds != 0, doubleSided != 0};
  if(nVar < 0 || nEq < 0 || nIneq < 0 || nStrongActIneq < 0 || nWeakActIneq < 0 || nStrongActBounds < 0 || nWeakActBounds < 0) return -1;
  if(nVar < nEq || nStrongActIneq + nWeakActIneq > nIneq) return -1;
  if(bounds ? (nStrongActBounds + nWeakActBounds > nVar) : (nStrongActBounds != 0 || nWeakActBounds != 0)) return -1;
  if(nEq + nStrongActIneq + nStrongActBounds > nVar) return -1;
  const int n = nVar, mc = nEq + nIneq, m = mc + (bounds ? n : 0);
  if(nthreads < 1) nthreads = 1;
  std::atomic<long> next(0);
  auto worker = [&]()
  {
    for(;;)
    {
      long k = next.fetch_add(1);
      if(k >= batch) break;
      Out o;
      o.G = G + k * static_cast<long>(n) * n;
      o.a = a + k * n;
      o.C = C + k * static_cast<long>(mc) * n;
      o.bl = bl + k * mc;
      o.bu = bu + k * mc;
      o.xl = bounds ? xl + k * n : nullptr;
      o.xu = bounds ? xu + k * n : nullptr;
      o.x = x ? x + k * n : nullptr;
      o.lambda = lambda ? lambda + k * m : nullptr;
      generateOne(ch, seed + static_cast<unsigned long long>(first_index + k), o);
    }
  };
  if(nthreads == 1)
    worker();
  else
  {
    std::vector<std::thread> th;
    for(int t = 0; t < nthreads; ++t) th.emplace_back(worker);
    for(auto & t : th) t.join();
  }
  return 0;
}

} // extern "C"
