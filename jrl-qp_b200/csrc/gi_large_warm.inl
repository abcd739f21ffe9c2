// Member functions of GiLarge<T, WARM> (included inside the struct, gi_large.cuh): warm-start capable
// initialisation, experimental::GoldfarbIdnaniSolver::init_ (src/experimental/GoldfarbIdnaniSolver.cpp:66-111)
// and the functions it calls (:306-486). Canonical arithmetic of oracle/warm_oracle.cpp, bit for bit.
//   B = L^-1 N in Bw (n x q column-major), Householder QR in place (R copied to the packed Rp at the
//   end, essential parts of the reflectors stay below the diagonal of Bw), J = L^-T Q.

// processInitialActiveSet (src/experimental/GoldfarbIdnaniSolver.cpp:306-381). Returns the status.
__device__ int warm_active_set(long long b)
{
  const bool use_as = P.as_in != nullptr && P.warm_start != 0;
  const signed char * as = use_as ? P.as_in + b * P.s_as : nullptr;
  const double big = P.big_bnd;
  for(int c = tid; c < m; c += T)
  {
    int s = ST_INACTIVE;
    if(c >= mc)
    {
      const int i = c - mc;
      const double lo = xl[i], up = xu[i];
      if(lo == up)
        s = ST_FIXED;
      else if(use_as)
      {
        const int g = as[c];
        if((g == ST_LOWER_BOUND && !(lo < -big)) || (g == ST_UPPER_BOUND && !(up > big))) s = g;
      }
    }
    else
    {
      const double lo = bl[c], up = bu[c];
      if(lo == up)
        s = ST_EQUALITY;
      else if(use_as)
      {
        const int g = as[c];
        if((g == ST_LOWER && !(lo < -big)) || (g == ST_UPPER && !(up > big)) || g == ST_EQUALITY) s = g;
      }
    }
    stat[c] = (signed char)s;
  }
  sync();
  // ordered active list: bounds first, then general constraints (activation order of the reference)
  if(warp == 0)
  {
    int cnt = 0, neq = 0;
    for(int base = 0; base < m; base += 32)
    {
      const int o = base + lane;
      const int c = o < nb ? mc + o : o - nb;
      const int sv = o < m ? stat[c] : ST_INACTIVE;
      const unsigned act = __ballot_sync(JRLQP_FULL, sv != ST_INACTIVE);
      neq += __popc(__ballot_sync(JRLQP_FULL, sv == ST_EQUALITY || sv == ST_FIXED));
      const int pos = cnt + __popc(act & ((1u << lane) - 1u));
      if(sv != ST_INACTIVE && pos < n) alist[pos] = c;
      cnt += __popc(act);
    }
    if(lane == 0)
    {
      iscr[4] = cnt;
      iscr[5] = neq;
    }
  }
  sync();
  int cnt = iscr[4];
  const int neq = iscr[5];
  if(cnt > n)
  {
    if(neq > n) return TS_OVERCONSTRAINED_PROBLEM;
    // walking the activation order backwards, every non-equality entry is dropped until n remain
    if(tid == 0)
    {
      int excess = cnt - n;
      for(int o = m - 1; o >= 0 && excess > 0; --o)
      {
        const int c = o < nb ? mc + o : o - nb;
        const int sv = stat[c];
        if(sv != ST_INACTIVE && sv != ST_EQUALITY && sv != ST_FIXED)
        {
          stat[c] = ST_INACTIVE;
          --excess;
        }
      }
      int pos = 0;
      for(int o = 0; o < m; ++o)
      {
        const int c = o < nb ? mc + o : o - nb;
        if(stat[c] != ST_INACTIVE) alist[pos++] = c;
      }
    }
    cnt = n;
    sync();
  }
  q = cnt;
  return TS_SUCCESS;
}

// Active normals, b_act (initializeComputationData, :383-418) and B = L^-1 N: one WARP per active column
// k. B(r,k) = (N(r,k) - dot4_{j<r}(L(r,j), B(j,k))) / L(r,r) is sequential in r; rows are processed in
// blocks of 32 (lane = row): the part of every chain that only involves rows above the block is
// accumulated by all the lanes in parallel (coalesced reads of L), then the 32 rows of the block are
// finished one after the other, the new entry being broadcast to the lanes below it. Every chain still
// receives its terms in ascending j: same bits as the sequential evaluation.
__device__ void warm_B(long long b, const double * __restrict__ Lw)
{
  for(int k = warp; k < q; k += NW)
  {
    double * Bk = Bw + (long long)k * ldl;
    const int ci = alist[k];
    const int sv = stat[ci];
    const bool general = ci < mc;
    const bool neg = general ? sv == ST_UPPER : sv == ST_UPPER_BOUND;
    const double * cg = general ? P.C + b * P.sC + (long long)ci * P.ldc : nullptr;
    const int pb = ci - mc;
    if(lane == 0) bact[k] = general ? (neg ? -bu[ci] : bl[ci]) : (neg ? -xu[pb] : xl[pb]);
#pragma unroll 1
    for(int r0 = 0; r0 < n; r0 += 32)
    {
      const int r = r0 + lane;
      const int rc = min(r, n - 1);
      double nr;
      if(general)
      {
        const double v = cg[rc];
        nr = neg ? -v : v;
      }
      else
        nr = rc == pb ? (neg ? -1.0 : 1.0) : 0.0;
      const double lrr = ldiag[rc];
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const double * Lr = Lw + rc;
      __syncwarp(); // entries of the previous block (written by single lanes) are visible
#pragma unroll 2
      for(int j = 0; j < r0; j += 4)
      {
        const double l0 = Lr[(long long)j * ldl], l1 = Lr[(long long)(j + 1) * ldl], l2 = Lr[(long long)(j + 2) * ldl], l3 = Lr[(long long)(j + 3) * ldl];
        const double2 b01 = *reinterpret_cast<const double2 *>(Bk + j);
        const double2 b23 = *reinterpret_cast<const double2 *>(Bk + j + 2);
        a0 = fma(l0, b01.x, a0);
        a1 = fma(l1, b01.y, a1);
        a2 = fma(l2, b23.x, a2);
        a3 = fma(l3, b23.y, a3);
      }
      const int nblk = min(32, n - r0);
#if JRLQP_WARMB_FAST
      // The 32 rows of the block are finished one after the other: what is serial is one quotient, one shuffle and one fma per
      // row. The entries of L a group of four rows needs do not depend on the chain: they are loaded one group AHEAD (one L2
      // round trip per four rows off the serial path instead of one per row on it), and the quotient comes from the stored
      // reciprocal of the diagonal with its proof of correct rounding (fp64_exact.cuh; the stock division when it declines).
      const double rlr = rinv[rc];
      double l0 = 0, l1 = 0, l2 = 0, l3 = 0;
      if(r < n)
      {
        if(lane > 0) l0 = Lr[(long long)r0 * ldl];
        if(lane > 1 && 1 < nblk) l1 = Lr[(long long)(r0 + 1) * ldl];
        if(lane > 2 && 2 < nblk) l2 = Lr[(long long)(r0 + 2) * ldl];
        if(lane > 3 && 3 < nblk) l3 = Lr[(long long)(r0 + 3) * ldl];
      }
#pragma unroll 1
      for(int jj = 0; jj < nblk; jj += 4)
      {
        double m0 = 0, m1 = 0, m2 = 0, m3 = 0; // the next group
        if(r < n && jj + 4 < nblk)
        {
          if(lane > jj + 4) m0 = Lr[(long long)(r0 + jj + 4) * ldl];
          if(lane > jj + 5 && jj + 5 < nblk) m1 = Lr[(long long)(r0 + jj + 5) * ldl];
          if(lane > jj + 6 && jj + 6 < nblk) m2 = Lr[(long long)(r0 + jj + 6) * ldl];
          if(lane > jj + 7 && jj + 7 < nblk) m3 = Lr[(long long)(r0 + jj + 7) * ldl];
        }
        // r0 is a multiple of 4: column r0 + jj + u feeds chain u
        {
          const double num = nr - ((a0 + a1) + (a2 + a3));
          bool okd;
          double val = div_rcp(num, lrr, rlr, okd);
          if(!okd) val = num / lrr;
          const double bj = __shfl_sync(JRLQP_FULL, val, jj);
          if(lane == jj) Bk[r] = val;
          if(lane > jj && r < n) a0 = fma(l0, bj, a0);
        }
        if(jj + 1 < nblk)
        {
          const double num = nr - ((a0 + a1) + (a2 + a3));
          bool okd;
          double val = div_rcp(num, lrr, rlr, okd);
          if(!okd) val = num / lrr;
          const double bj = __shfl_sync(JRLQP_FULL, val, jj + 1);
          if(lane == jj + 1) Bk[r] = val;
          if(lane > jj + 1 && r < n) a1 = fma(l1, bj, a1);
        }
        if(jj + 2 < nblk)
        {
          const double num = nr - ((a0 + a1) + (a2 + a3));
          bool okd;
          double val = div_rcp(num, lrr, rlr, okd);
          if(!okd) val = num / lrr;
          const double bj = __shfl_sync(JRLQP_FULL, val, jj + 2);
          if(lane == jj + 2) Bk[r] = val;
          if(lane > jj + 2 && r < n) a2 = fma(l2, bj, a2);
        }
        if(jj + 3 < nblk)
        {
          const double num = nr - ((a0 + a1) + (a2 + a3));
          bool okd;
          double val = div_rcp(num, lrr, rlr, okd);
          if(!okd) val = num / lrr;
          const double bj = __shfl_sync(JRLQP_FULL, val, jj + 3);
          if(lane == jj + 3) Bk[r] = val;
          if(lane > jj + 3 && r < n) a3 = fma(l3, bj, a3);
        }
        l0 = m0;
        l1 = m1;
        l2 = m2;
        l3 = m3;
      }
#else
#pragma unroll 1
      for(int jj = 0; jj < nblk; jj += 4)
      {
        // r0 is a multiple of 4: column r0 + jj + u feeds chain u
        {
          const double val = (nr - ((a0 + a1) + (a2 + a3))) / lrr;
          const double bj = __shfl_sync(JRLQP_FULL, val, jj);
          if(lane == jj) Bk[r] = val;
          if(lane > jj && r < n) a0 = fma(Lr[(long long)(r0 + jj) * ldl], bj, a0);
        }
        if(jj + 1 < nblk)
        {
          const double val = (nr - ((a0 + a1) + (a2 + a3))) / lrr;
          const double bj = __shfl_sync(JRLQP_FULL, val, jj + 1);
          if(lane == jj + 1) Bk[r] = val;
          if(lane > jj + 1 && r < n) a1 = fma(Lr[(long long)(r0 + jj + 1) * ldl], bj, a1);
        }
        if(jj + 2 < nblk)
        {
          const double val = (nr - ((a0 + a1) + (a2 + a3))) / lrr;
          const double bj = __shfl_sync(JRLQP_FULL, val, jj + 2);
          if(lane == jj + 2) Bk[r] = val;
          if(lane > jj + 2 && r < n) a2 = fma(Lr[(long long)(r0 + jj + 2) * ldl], bj, a2);
        }
        if(jj + 3 < nblk)
        {
          const double val = (nr - ((a0 + a1) + (a2 + a3))) / lrr;
          const double bj = __shfl_sync(JRLQP_FULL, val, jj + 3);
          if(lane == jj + 3) Bk[r] = val;
          if(lane > jj + 3 && r < n) a3 = fma(Lr[(long long)(r0 + jj + 3) * ldl], bj, a3);
        }
      }
#endif
    }
  }
}

// Householder QR of B (n x q, Bw) in place, unblocked, then R -> packed Rp
__device__ void warm_qr()
{
#pragma unroll 1
  for(int k = 0; k < q; ++k)
  {
    const int len = n - k - 1;
    double * Bk = Bw + (long long)k * ldl;
    double * ess = Bk + k + 1;
    const double c0 = Bk[k];
    double tsq = 0.0;
    for(int t = lane; t < len; t += 32) // dot32 order, every warp redundantly
    {
      const double e = ess[t];
      tsq = fma(e, e, tsq);
    }
    tsq = warp_sum32(tsq);
    double tau, beta;
    const bool degenerate = tsq <= 2.2250738585072014e-308;
    if(degenerate)
    {
      tau = 0.0;
      beta = c0;
    }
    else
    {
      beta = sqrt(fma(c0, c0, tsq));
      if(c0 >= 0.0) beta = -beta;
      tau = (beta - c0) / beta;
    }
    sync(); // everybody has read column k before it is rewritten
    {
      const double den = c0 - beta;
      for(int t = tid; t < len; t += T) ess[t] = degenerate ? 0.0 : ess[t] / den;
    }
    if(tid == 0)
    {
      Bk[k] = beta;
      hco[k] = tau;
    }
    sync();
    if(len == 0)
    {
      for(int j = k + 1 + tid; j < q; j += T)
      {
        double * top = Bw + (long long)j * ldl + k;
        *top = *top * (1.0 - tau);
      }
    }
    else if(tau != 0.0)
    {
      // tmp_j = dot4(ess, bottom_j) + top_j: 4 lanes per column (lane t = chain t), 8 columns per warp
      const int t4 = lane & 3, sub = lane >> 2;
      for(int j0 = k + 1 + 8 * warp; j0 < q; j0 += 8 * NW)
      {
        const int j = j0 + sub;
        const double * bot = Bw + (long long)min(j, q - 1) * ldl + k + 1;
        double acc = 0.0;
        if(j < q)
          for(int t = t4; t < len; t += 4) acc = fma(ess[t], bot[t], acc);
        const double a1 = __shfl_down_sync(JRLQP_FULL, acc, 1);
        const double s01 = acc + a1; // valid where t4 is even
        const double s23 = __shfl_down_sync(JRLQP_FULL, s01, 2);
        if(t4 == 0 && j < q) wv[j] = (s01 + s23) + bot[-1];
      }
      sync();
      // top_j = fma(-tau, tmp_j, top_j) ; bottom(i,j) = fma(-(tau ess_i), tmp_j, bottom(i,j)): warp = column, lanes = rows
      for(int j = k + 1 + warp; j < q; j += NW)
      {
        double * col = Bw + (long long)j * ldl + k;
        const double tmp = wv[j];
        if(lane == 0) col[0] = fma(-tau, tmp, col[0]);
        for(int t = lane; t < len; t += 32) col[1 + t] = fma(-(tau * ess[t]), tmp, col[1 + t]);
      }
    }
    sync();
  }
  for(int k = warp; k < q; k += NW)
    for(int i = lane; i <= k; i += 32) Rp[colR(k) + i] = Bw[(long long)k * ldl + i];
  sync();
}

// J = J Q with the rows of J taken through shared memory, R = srows (8 or 16) at a time. The reflectors act on the rows of J
// independently, and a row needs ALL its entries twice per reflector (inner product, then update): with thread = row on the
// column-major workspace (warm_JQ below) every reflector streams J twice from L2 / HBM with a few values per thread in flight
// — 60 % of a warm-started n = 387 solve (profiles/r5o_*: long_scoreboard, issue slots 8 % busy). Here a slab of R rows (all
// columns: R n doubles, 25 KB at n = 387) is loaded once, every reflector is applied to it on chip, and it is stored once.
// Inner product of a row: four lanes, lane c runs chain c of the canonical dot4 (entries t = c, c + 4, ... ascending), the
// chains are folded (a0 + a1) + (a2 + a3) by two shuffles; update: all threads, one entry each. Same operations per entry in
// the same order as warm_JQ: same bits. The essential part of the next reflector is staged by the warps that have no inner
// product to run.
__device__ void warm_JQ_slab()
{
  const int R = srows, rs = R == 16 ? 4 : 3;
  double * S = slab; // S[c * R + r] = J(r0 + r, c)
  const int ne = (n + 1) & ~1;
  double * E0 = slab + (long long)n * R; // two buffers for ess_k
  double * TT = E0 + 2 * ne; // tau * tmp of every row of the slab
  const int ndot = 4 * R; // threads of the inner products (whole warps: R = 8 / 16)
  for(int r0 = 0; r0 < n; r0 += R)
  {
    const int nr = min(R, n - r0);
    // ---- load the slab (8 consecutive threads read 64 contiguous bytes of a column), four values in flight per thread
    {
      const int tot = n << rs;
      for(int e0 = tid; e0 < tot; e0 += 4 * T)
      {
        double v[4];
#pragma unroll
        for(int u = 0; u < 4; ++u)
        {
          const int e = e0 + u * T;
          const int c = e >> rs, r = e & (R - 1);
          v[u] = (e < tot && r < nr) ? Jc[(long long)c * ldl + r0 + r] : 0.0;
        }
#pragma unroll
        for(int u = 0; u < 4; ++u)
          if(e0 + u * T < tot) S[e0 + u * T] = v[u];
      }
    }
    // first reflector with something to do: stage its essential part
    int k = 0;
    while(k < q && !(n - k - 1 == 0 || hco[k] != 0.0)) ++k;
    int buf = 0;
    if(k < q)
    {
      const int len = n - k - 1;
      const double * ess = Bw + (long long)k * ldl + k + 1;
      for(int t = tid; t < len; t += T) E0[t] = ess[t];
    }
    sync();
#pragma unroll 1
    while(k < q)
    {
      const int len = n - k - 1;
      const double tau = hco[k];
      int kn = k + 1; // the next reflector with something to do
      while(kn < q && !(n - kn - 1 == 0 || hco[kn] != 0.0)) ++kn;
      const double * E = E0 + buf * ne;
      if(len == 0)
      {
        if(tid < nr) S[(k << rs) + tid] = S[(k << rs) + tid] * (1.0 - tau);
      }
      else
      {
        if(tid < ndot)
        {
          const int r = tid >> 2, c = tid & 3;
          const double * Sr = S + ((long long)(k + 1) << rs) + r;
          double a = 0.0;
#pragma unroll 4
          for(int t = c; t < len; t += 4) a = fma(Sr[t << rs], E[t], a);
          a = a + __shfl_xor_sync(JRLQP_FULL, a, 1);
          a = a + __shfl_xor_sync(JRLQP_FULL, a, 2);
          const double top = S[(k << rs) + r];
          const double tmp = a + top;
          __syncwarp();
          if(c == 0)
          {
            S[(k << rs) + r] = fma(-tau, tmp, top);
            TT[r] = tau * tmp;
          }
        }
        else if(kn < q)
        {
          // the other warps stage ess of the next reflector meanwhile
          const int lenn = n - kn - 1;
          const double * essn = Bw + (long long)kn * ldl + kn + 1;
          double * En = E0 + (buf ^ 1) * ne;
          for(int t = tid - ndot; t < lenn; t += T - ndot) En[t] = essn[t];
        }
        sync();
        // update: entry (t2, r) <- fma(-tt_r, ess[t2], entry)
        {
          const int tot = len << rs;
          double * Su = S + ((long long)(k + 1) << rs);
          for(int e = tid; e < tot; e += T)
          {
            const int t2 = e >> rs, r = e & (R - 1);
            Su[e] = fma(-TT[r], E[t2], Su[e]);
          }
        }
      }
      sync();
      if(len == 0 && kn < q)
      {
        // (k = n - 1 is the last reflector there can be: nothing follows; kept for completeness)
        const int lenn = n - kn - 1;
        const double * essn = Bw + (long long)kn * ldl + kn + 1;
        double * En = E0 + (buf ^ 1) * ne;
        for(int t = tid; t < lenn; t += T) En[t] = essn[t];
        sync();
      }
      buf ^= 1;
      k = kn;
    }
    // ---- store the slab
    {
      const int tot = n << rs;
      for(int e = tid; e < tot; e += T)
      {
        const int c = e >> rs, r = e & (R - 1);
        if(r < nr) Jc[(long long)c * ldl + r0 + r] = S[e];
      }
    }
    sync();
  }
}

// J = J Q (HouseholderSequence::applyThisOnTheRight), thread = row of the column-major J
__device__ void warm_JQ()
{
  if(srows > 0)
  {
    warm_JQ_slab();
    return;
  }
  for(int row = tid; row < n; row += T)
  {
    double * Ji = Jc + row;
#pragma unroll 1
    for(int k = 0; k < q; ++k)
    {
      const int len = n - k - 1;
      const double * ess = Bw + (long long)k * ldl + k + 1;
      const double tau = hco[k];
      double * Jk = Ji + (long long)k * ldl;
      if(len == 0)
        *Jk = *Jk * (1.0 - tau);
      else if(tau != 0.0)
      {
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        const double * Jt = Jk + ldl;
        int t = 0;
#pragma unroll 2
        for(; t + 3 < len; t += 4)
        {
          const double j0 = Jt[(long long)t * ldl], j1 = Jt[(long long)(t + 1) * ldl], j2 = Jt[(long long)(t + 2) * ldl], j3 = Jt[(long long)(t + 3) * ldl];
          a0 = fma(j0, ess[t], a0);
          a1 = fma(j1, ess[t + 1], a1);
          a2 = fma(j2, ess[t + 2], a2);
          a3 = fma(j3, ess[t + 3], a3);
        }
        if(t < len) a0 = fma(Jt[(long long)t * ldl], ess[t], a0);
        if(t + 1 < len) a1 = fma(Jt[(long long)(t + 1) * ldl], ess[t + 1], a1);
        if(t + 2 < len) a2 = fma(Jt[(long long)(t + 2) * ldl], ess[t + 2], a2);
        const double tmp = ((a0 + a1) + (a2 + a3)) + *Jk;
        *Jk = fma(-tau, tmp, *Jk);
        const double tt = tau * tmp;
        double * Jw = Jk + ldl;
#pragma unroll 4
        for(int t2 = 0; t2 < len; ++t2) Jw[(long long)t2 * ldl] = fma(-tt, ess[t2], Jw[(long long)t2 * ldl]);
      }
    }
  }
  sync();
}

// initializePrimalDualPoints (src/experimental/GoldfarbIdnaniSolver.cpp:461-486)
__device__ void warm_primal_dual(const double * ab)
{
  for(int i = tid; i < n; i += T) cv[i] = __ldg(ab + i);
  sync();
  // alpha = J^T a, thread = column
  for(int j = tid; j < n; j += T)
  {
    const double * Jj = Jc + (long long)j * ldl;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int i = 0;
#pragma unroll 2
    for(; i + 3 < n; i += 4)
    {
      const double2 p0 = *reinterpret_cast<const double2 *>(Jj + i);
      const double2 p1 = *reinterpret_cast<const double2 *>(Jj + i + 2);
      a0 = fma(p0.x, cv[i], a0);
      a1 = fma(p0.y, cv[i + 1], a1);
      a2 = fma(p1.x, cv[i + 2], a2);
      a3 = fma(p1.y, cv[i + 3], a3);
    }
    if(i < n) a0 = fma(Jj[i], cv[i], a0);
    if(i + 1 < n) a1 = fma(Jj[i + 1], cv[i + 1], a1);
    if(i + 2 < n) a2 = fma(Jj[i + 2], cv[i + 2], a2);
    alp[j] = (a0 + a1) + (a2 + a3);
  }
  // beta = R^-T b_act on warp 0 (column-oriented forward substitution, true division), into zs
  if(warp == 0)
  {
    for(int k = lane; k < q; k += 32) wv[k] = bact[k];
    __syncwarp();
#pragma unroll 1
    for(int k = 0; k < q; ++k)
    {
      const double bk = wv[k] / Rp[colR(k) + k];
      __syncwarp();
      if(lane == 0) zs[k] = bk;
      for(int i = k + 1 + lane; i < q; i += 32) wv[i] = fma(-bk, Rp[colR(i) + k], wv[i]);
      __syncwarp();
    }
  }
  sync();
  // x = J1 beta - J2 alpha2, thread = row ; d = alpha1 + beta (right-hand side of u)
  for(int i = tid; i < n; i += T)
  {
    const double * Ji = Jc + i;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int c = 0;
#pragma unroll 2
    for(; c + 3 < q; c += 4)
    {
      const double j0 = Ji[(long long)c * ldl], j1 = Ji[(long long)(c + 1) * ldl], j2 = Ji[(long long)(c + 2) * ldl], j3 = Ji[(long long)(c + 3) * ldl];
      a0 = fma(j0, zs[c], a0);
      a1 = fma(j1, zs[c + 1], a1);
      a2 = fma(j2, zs[c + 2], a2);
      a3 = fma(j3, zs[c + 3], a3);
    }
    if(c < q) a0 = fma(Ji[(long long)c * ldl], zs[c], a0);
    if(c + 1 < q) a1 = fma(Ji[(long long)(c + 1) * ldl], zs[c + 1], a1);
    if(c + 2 < q) a2 = fma(Ji[(long long)(c + 2) * ldl], zs[c + 2], a2);
    const double s1 = (a0 + a1) + (a2 + a3);
    a0 = a1 = a2 = a3 = 0;
    c = q;
#pragma unroll 2
    for(; c + 3 < n; c += 4)
    {
      const double j0 = Ji[(long long)c * ldl], j1 = Ji[(long long)(c + 1) * ldl], j2 = Ji[(long long)(c + 2) * ldl], j3 = Ji[(long long)(c + 3) * ldl];
      a0 = fma(j0, alp[c], a0);
      a1 = fma(j1, alp[c + 1], a1);
      a2 = fma(j2, alp[c + 2], a2);
      a3 = fma(j3, alp[c + 3], a3);
    }
    if(c < n) a0 = fma(Ji[(long long)c * ldl], alp[c], a0);
    if(c + 1 < n) a1 = fma(Ji[(long long)(c + 1) * ldl], alp[c + 1], a1);
    if(c + 2 < n) a2 = fma(Ji[(long long)(c + 2) * ldl], alp[c + 2], a2);
    const double s2 = (a0 + a1) + (a2 + a3);
    xs[i] = s1 - s2;
  }
  sync(); // xs (and the reads of zs) complete before ds is rewritten
  for(int k = tid; k < q; k += T) ds[k] = alp[k] + zs[k];
  sync();
  // u = R^-1 (alpha1 + beta)
  if(warp == 0)
  {
    back_substitution();
    __syncwarp();
    for(int k = lane; k < q; k += 32) us[k] = rs[k];
  }
  // f = beta.(0.5 beta + alpha1) - 0.5 |alpha2|^2 (dot32 order), every warp redundantly
  {
    double s1 = 0.0, s2 = 0.0;
    for(int k = lane; k < q; k += 32)
    {
      const double bk = zs[k];
      s1 = fma(bk, fma(0.5, bk, alp[k]), s1);
    }
    for(int k = lane; k < n - q; k += 32)
    {
      const double ak = alp[q + k];
      s2 = fma(ak, ak, s2);
    }
    f = warp_sum32(s1) - 0.5 * warp_sum32(s2);
  }
  sync();
}

// experimental init_: returns the termination status (SUCCESS: ready for the main loop)
__device__ int init_warm(long long b, int & it)
{
  const double * __restrict__ ab = P.a + b * P.sa;
  int st = warm_active_set(b);
  if(st != TS_SUCCESS) return st;
  if(P.pre != nullptr)
  {
    // factor shared by the batch (see init()): B = L^-1 N from the shared L, J copied
    if(!load_prefactor(b)) return TS_NON_POS_HESSIAN;
    copy_pre_J(0);
    warm_B(b, pre_L());
    sync();
  }
  else
  {
    if(!cholesky(b)) return TS_NON_POS_HESSIAN;
    build_J(0);
    warm_B(b, Lw); // reads L: before J takes over its storage
    sync();
    transpose_J();
  }
  warm_qr();
  warm_JQ();
  for(int c = tid; c < m; c += T) eqf[c] = 0;
  warm_primal_dual(ab);

  // constraints activated with a negative multiplier are dropped, most negative first (:83-108)
#pragma unroll 1
  for(;;)
  {
    double bu_ = -1e-14;
    int bl_ = JRLQP_NONE;
    for(int l = lane; l < q; l += 32)
    {
      const int sv = stat[alist[l]];
      const double ul = us[l];
      if(ul < bu_ && sv != ST_FIXED && sv != ST_EQUALITY)
      {
        bu_ = ul;
        bl_ = l;
      }
    }
#pragma unroll
    for(int off = 16; off >= 1; off >>= 1)
    {
      const double ou = __shfl_xor_sync(JRLQP_FULL, bu_, off);
      const int ol = __shfl_xor_sync(JRLQP_FULL, bl_, off);
      if(ou < bu_ || (ou == bu_ && ol < bl_))
      {
        bu_ = ou;
        bl_ = ol;
      }
    }
    const int lmin = __shfl_sync(JRLQP_FULL, bl_, 0);
    if(lmin == JRLQP_NONE) break;
    ++it;
    sync();
    // b_act.segment(lmin, q-1-lmin) = b_act.tail(q-1-lmin)
    for(int k = lmin + tid; k + 1 < q; k += T) wv[k] = bact[k + 1];
    sync();
    for(int k = lmin + tid; k + 1 < q; k += T) bact[k] = wv[k];
    remove_constraint(lmin);
    warm_primal_dual(ab);
  }
  return TS_SUCCESS;
}
