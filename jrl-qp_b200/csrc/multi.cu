// Multi-GPU entry points of the C-ABI (include/jrlqp_b200.h, jrlqp_multi_*): ONE handle, one host batch, every GPU
// of the box. QPs are independent, so the batch is cut into contiguous shards [lo_k, hi_k) (sizes differ by at most
// one: the rule of jrl-qp_b200/sharding.py), shard k is solved on device k by its own jrlqp_solver (own streams, own
// device staging), driven by a persistent host thread per device; results land directly in the caller's arrays
// (disjoint ranges), so the "gather" costs no extra copy. No collective, no NCCL: nothing crosses QPs.
// Arrays shared by the batch (stride 0) are uploaded — and, for the large-n kernel, factorised — once per device.
//
// Also here: jrlqp_measure_host_link, the platform probe behind the end-to-end numbers (concurrent pinned
// host -> device / device -> host copies over any subset of the GPUs).
#include "jrlqp_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace
{

struct Job
{
  const jrlqp_problem * pb = nullptr;
  const jrlqp_result * res = nullptr;
  bool warm = false;
};

struct Worker
{
  int device = 0;
  jrlqp_solver * solver = nullptr;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  bool has_job = false, done = false, quit = false;
  jrlqp_problem pb{};
  jrlqp_result res{};
  bool warm = false;
  int rc = 0;
  double seconds = 0.0; // wall time of the last call on this device
};

void shard_range(long long batch, int k, int g, long long & lo, long long & hi)
{
  const long long base = batch / g, rem = batch % g;
  lo = k * base + std::min<long long>(k, rem);
  hi = lo + base + (k < rem ? 1 : 0);
}

} // namespace

struct jrlqp_multi
{
  int n = 0, mc = 0, nb = 0, m = 0;
  long long capacity = 0;
  long long dev_capacity = 0; // capacity of every per-device solver
  std::vector<Worker *> workers;
  // Load balancing: shard k of the next call gets a share weight[k] / sum of the batch. The weights start equal and,
  // when balancing is on, follow the throughput every device achieved in the previous calls (QPs per second of its
  // shard): on a box whose GPUs do not see the same host-link bandwidth (profiles/r02n_multi_e2e.txt: 23 GB/s for four
  // of the GPUs against 35 GB/s for the other four when all eight copy at once) equal shards leave the fast links idle
  // while the slow ones finish.
  std::vector<double> weight;
  bool balancing = true;
  std::string err;
};

// contiguous shards proportional to the weights (equal weights: sizes differ by at most one, the rule of sharding.py)
static void weighted_shards(const jrlqp_multi * mh, long long batch, std::vector<long long> & bounds)
{
  const int g = (int)mh->workers.size();
  bounds.assign((size_t)g + 1, 0);
  bool equal = true;
  for(int k = 1; k < g; ++k) equal = equal && mh->weight[(size_t)k] == mh->weight[0];
  if(equal)
  {
    for(int k = 0; k < g; ++k)
    {
      long long lo, hi;
      shard_range(batch, k, g, lo, hi);
      bounds[(size_t)k] = lo;
      bounds[(size_t)k + 1] = hi;
    }
    return;
  }
  double sum = 0.0;
  for(double w : mh->weight) sum += w;
  double acc = 0.0;
  for(int k = 0; k < g; ++k)
  {
    acc += mh->weight[(size_t)k];
    long long hi = k == g - 1 ? batch : (long long)((double)batch * (acc / sum) + 0.5);
    hi = std::max(hi, bounds[(size_t)k]);
    hi = std::min(hi, std::min(batch, bounds[(size_t)k] + mh->dev_capacity)); // never above a device's capacity
    bounds[(size_t)k + 1] = hi;
  }
  // whatever the clamps left over goes to the devices that still have room, from the last one backwards
  long long rest = batch - bounds[(size_t)g];
  for(int k = g - 1; k >= 0 && rest > 0; --k)
  {
    const long long room = mh->dev_capacity - (bounds[(size_t)k + 1] - bounds[(size_t)k]);
    const long long add = std::min(room, rest);
    for(int j = k + 1; j <= g; ++j) bounds[(size_t)j] += add;
    rest -= add;
  }
}

static void worker_main(Worker * w)
{
  cudaSetDevice(w->device);
  for(;;)
  {
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->has_job || w->quit; });
    if(w->quit) return;
    w->has_job = false;
    lk.unlock();
    int rc = JRLQP_SUCCESS;
    const auto t0 = std::chrono::steady_clock::now();
    if(w->pb.batch > 0) rc = w->warm ? jrlqp_solve_batch_warm_host(w->solver, &w->pb, &w->res) : jrlqp_solve_batch_host(w->solver, &w->pb, &w->res);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    lk.lock();
    w->seconds = dt;
    w->rc = rc;
    w->done = true;
    lk.unlock();
    w->cv.notify_all();
  }
}

extern "C"
{

int jrlqp_multi_create(jrlqp_multi ** out, int32_t n, int32_t mc, int32_t use_bounds, int64_t batch_capacity, const int32_t * devices, int32_t n_devices)
{
  if(!out) return JRLQP_ERR_ARG;
  *out = nullptr;
  if(n < 1 || mc < 0 || batch_capacity < 0) return JRLQP_ERR_ARG;
  jrlqp_multi * mh = new jrlqp_multi();
  *out = mh; // returned even on failure so that jrlqp_multi_last_error is readable
  mh->n = n;
  mh->mc = mc;
  mh->nb = use_bounds ? n : 0;
  mh->m = mc + mh->nb;
  mh->capacity = batch_capacity;
  int ndev = 0;
  if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
  {
    mh->err = "no CUDA device";
    return JRLQP_ERR_CUDA;
  }
  std::vector<int> devs;
  if(devices && n_devices > 0)
    devs.assign(devices, devices + n_devices);
  else
  {
    const int g = n_devices > 0 ? std::min<int>(n_devices, ndev) : ndev;
    for(int d = 0; d < g; ++d) devs.push_back(d);
  }
  const int g = (int)devs.size();
  // every device can take up to 1.5 x its equal share (load balancing), never more than the whole batch
  mh->dev_capacity = std::min<long long>(batch_capacity, ((batch_capacity + g - 1) / g * 3 + 1) / 2);
  mh->weight.assign((size_t)g, 1.0);
  for(int k = 0; k < g; ++k)
  {
    const long long cap = mh->dev_capacity;
    Worker * w = new Worker();
    w->device = devs[k];
    mh->workers.push_back(w);
    const int rc = jrlqp_create(&w->solver, n, mc, use_bounds, cap, devs[k]);
    if(rc != JRLQP_OK)
    {
      mh->err = std::string("device ") + std::to_string(devs[k]) + ": " + (w->solver ? jrlqp_last_error(w->solver) : "jrlqp_create failed");
      return rc;
    }
  }
  for(Worker * w : mh->workers) w->th = std::thread(worker_main, w);
  return JRLQP_OK;
}

int jrlqp_multi_destroy(jrlqp_multi * mh)
{
  if(!mh) return JRLQP_OK;
  for(Worker * w : mh->workers)
  {
    if(w->th.joinable())
    {
      {
        std::lock_guard<std::mutex> lk(w->mu);
        w->quit = true;
      }
      w->cv.notify_all();
      w->th.join();
    }
    if(w->solver) jrlqp_destroy(w->solver);
    delete w;
  }
  delete mh;
  return JRLQP_OK;
}

int jrlqp_multi_device_count(const jrlqp_multi * mh)
{
  return mh ? (int)mh->workers.size() : 0;
}

int jrlqp_multi_device(const jrlqp_multi * mh, int32_t k)
{
  if(!mh || k < 0 || k >= (int)mh->workers.size()) return -1;
  return mh->workers[(size_t)k]->device;
}

jrlqp_solver * jrlqp_multi_solver(jrlqp_multi * mh, int32_t k)
{
  if(!mh || k < 0 || k >= (int)mh->workers.size()) return nullptr;
  return mh->workers[(size_t)k]->solver;
}

int jrlqp_multi_shard(const jrlqp_multi * mh, int64_t batch, int32_t k, int64_t * begin, int64_t * end)
{
  if(!mh || !begin || !end || k < 0 || k >= (int)mh->workers.size() || batch < 0) return JRLQP_ERR_ARG;
  std::vector<long long> bounds;
  weighted_shards(mh, batch, bounds);
  *begin = bounds[(size_t)k];
  *end = bounds[(size_t)k + 1];
  return JRLQP_OK;
}

int jrlqp_multi_set_balancing(jrlqp_multi * mh, int32_t on)
{
  if(!mh) return JRLQP_ERR_ARG;
  mh->balancing = on != 0;
  if(!mh->balancing) std::fill(mh->weight.begin(), mh->weight.end(), 1.0);
  return JRLQP_OK;
}

int jrlqp_multi_get_weights(const jrlqp_multi * mh, double * weights)
{
  if(!mh || !weights) return JRLQP_ERR_ARG;
  double sum = 0.0;
  for(double w : mh->weight) sum += w;
  for(size_t k = 0; k < mh->weight.size(); ++k) weights[k] = mh->weight[k] / sum;
  return JRLQP_OK;
}

int jrlqp_multi_set_options(jrlqp_multi * mh, const jrlqp_options * opt)
{
  if(!mh || !opt) return JRLQP_ERR_ARG;
  for(Worker * w : mh->workers)
  {
    const int rc = jrlqp_set_options(w->solver, opt);
    if(rc != JRLQP_OK) return rc;
  }
  return JRLQP_OK;
}

const char * jrlqp_multi_last_error(const jrlqp_multi * mh)
{
  return mh ? mh->err.c_str() : "null handle";
}

static int multi_solve(jrlqp_multi * mh, const jrlqp_problem * pb, const jrlqp_result * res, bool warm)
{
  if(!mh || !pb || !res || !res->x) return JRLQP_ERR_ARG;
  if(pb->batch < 0) return JRLQP_ERR_ARG;
  if(pb->batch > mh->capacity) return JRLQP_ERR_CAPACITY;
  const int g = (int)mh->workers.size();
  const long long n = mh->n, m = mh->m;
  std::vector<long long> bounds;
  weighted_shards(mh, pb->batch, bounds);
  // scatter: shard k = a view of the caller's arrays (pointer + offset; shared arrays are passed as they are)
  for(int k = 0; k < g; ++k)
  {
    Worker * w = mh->workers[(size_t)k];
    const long long lo = bounds[(size_t)k], hi = bounds[(size_t)k + 1];
    jrlqp_problem sp = *pb;
    sp.batch = hi - lo;
    auto off = [&](const double * p, long long stride) { return p ? p + lo * stride : nullptr; };
    sp.G = off(pb->G, pb->G_stride);
    sp.a = off(pb->a, pb->a_stride);
    sp.C = off(pb->C, pb->C_stride);
    sp.bl = off(pb->bl, pb->bl_stride);
    sp.bu = off(pb->bu, pb->bu_stride);
    sp.xl = off(pb->xl, pb->xl_stride);
    sp.xu = off(pb->xu, pb->xu_stride);
    sp.as_in = pb->as_in ? pb->as_in + lo * pb->as_stride : nullptr;
    jrlqp_result sr{};
    sr.x = res->x + lo * n;
    sr.u = res->u ? res->u + lo * m : nullptr;
    sr.f = res->f ? res->f + lo : nullptr;
    sr.iterations = res->iterations ? res->iterations + lo : nullptr;
    sr.status = res->status ? res->status + lo : nullptr;
    sr.active_set = res->active_set ? res->active_set + lo * m : nullptr;
    sr.active_list = res->active_list ? res->active_list + lo * n : nullptr;
    sr.n_active = res->n_active ? res->n_active + lo : nullptr;
    sr.L = res->L ? res->L + lo * n * n : nullptr;
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->pb = sp;
      w->res = sr;
      w->warm = warm;
      w->done = false;
      w->has_job = true;
    }
    w->cv.notify_all();
  }
  // gather: the results are already in place; wait for every device and reduce the return value
  int worst = 0, err = 0;
  for(int k = 0; k < g; ++k)
  {
    Worker * w = mh->workers[(size_t)k];
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->done; });
    if(w->rc < 0)
    {
      if(err == 0)
      {
        err = w->rc;
        mh->err = std::string("device ") + std::to_string(w->device) + ": " + jrlqp_last_error(w->solver);
      }
    }
    else
      worst = std::max(worst, w->rc);
  }
  // load balancing: next call's shares follow the throughput each device just achieved (smoothed), when the shards were
  // large enough for the measurement to mean something
  if(mh->balancing && err == 0 && g > 1 && pb->batch >= 4096ll * g)
  {
    double tot = 0.0;
    std::vector<double> rate((size_t)g, 0.0);
    bool valid = true;
    for(int k = 0; k < g; ++k)
    {
      const long long cnt = bounds[(size_t)k + 1] - bounds[(size_t)k];
      const double dt = mh->workers[(size_t)k]->seconds;
      valid = valid && cnt > 0 && dt > 0.0;
      rate[(size_t)k] = valid ? (double)cnt / dt : 0.0;
      tot += rate[(size_t)k];
    }
    if(valid)
    {
      double wsum = 0.0;
      for(double w : mh->weight) wsum += w;
      for(int k = 0; k < g; ++k) mh->weight[(size_t)k] = 0.5 * mh->weight[(size_t)k] / wsum + 0.5 * rate[(size_t)k] / tot;
      // shares sum to one and none exceeds 1.5 / g (the capacity of a device): what is cut off goes to the others
      const double cap_share = 1.5 / g;
      for(int pass = 0; pass < g; ++pass)
      {
        double excess = 0.0, free_sum = 0.0;
        for(double & w : mh->weight)
        {
          if(w > cap_share)
          {
            excess += w - cap_share;
            w = cap_share;
          }
          else if(w < cap_share)
            free_sum += w;
        }
        if(excess <= 0.0 || free_sum <= 0.0) break;
        for(double & w : mh->weight)
          if(w < cap_share) w += excess * w / free_sum;
      }
    }
  }
  return err != 0 ? err : worst;
}

int jrlqp_multi_solve_batch_host(jrlqp_multi * mh, const jrlqp_problem * pb, const jrlqp_result * res)
{
  return multi_solve(mh, pb, res, false);
}

int jrlqp_multi_solve_batch_warm_host(jrlqp_multi * mh, const jrlqp_problem * pb, const jrlqp_result * res)
{
  return multi_solve(mh, pb, res, true);
}

// ---------------------------------------------------------------------------------------------
// Platform probe: aggregate bandwidth of concurrent pinned-host <-> device copies over `n_devices` GPUs
// (direction: 0 host -> device, 1 device -> host, 2 both at once). Every device copies `bytes` per repetition
// from / to its own pinned buffer on its own stream, `reps` times back to back; the wall time between a common
// start and the last completion gives the aggregate GB/s (returned; per-device GB/s in per_device[] if not NULL).
// ---------------------------------------------------------------------------------------------
double jrlqp_measure_host_link(const int32_t * devices, int32_t n_devices, int64_t bytes, int32_t reps, int32_t direction, double * per_device)
{
  if(n_devices < 1 || bytes < 1 || reps < 1) return -1.0;
  struct Dev
  {
    int dev;
    void *h = nullptr, *d = nullptr, *h2 = nullptr, *d2 = nullptr;
    cudaStream_t s = nullptr, s2 = nullptr;
    double seconds = 0.0;
  };
  std::vector<Dev> dv((size_t)n_devices);
  bool ok = true;
  for(int k = 0; k < n_devices; ++k)
  {
    Dev & x = dv[(size_t)k];
    x.dev = devices ? devices[k] : k;
    ok = ok && cudaSetDevice(x.dev) == cudaSuccess;
    ok = ok && cudaHostAlloc(&x.h, (size_t)bytes, cudaHostAllocPortable) == cudaSuccess;
    ok = ok && cudaMalloc(&x.d, (size_t)bytes) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) == cudaSuccess;
    if(direction == 2)
    {
      ok = ok && cudaHostAlloc(&x.h2, (size_t)bytes, cudaHostAllocPortable) == cudaSuccess;
      ok = ok && cudaMalloc(&x.d2, (size_t)bytes) == cudaSuccess;
      ok = ok && cudaStreamCreateWithFlags(&x.s2, cudaStreamNonBlocking) == cudaSuccess;
    }
    if(ok) std::memset(x.h, 1, (size_t)bytes); // touch the pages
  }
  double agg = -1.0;
  if(ok)
  {
    auto run = [&](Dev & x, int nrep)
    {
      cudaSetDevice(x.dev);
      for(int r = 0; r < nrep; ++r)
      {
        if(direction == 0 || direction == 2) cudaMemcpyAsync(x.d, x.h, (size_t)bytes, cudaMemcpyHostToDevice, x.s);
        if(direction == 1) cudaMemcpyAsync(x.h, x.d, (size_t)bytes, cudaMemcpyDeviceToHost, x.s);
        if(direction == 2) cudaMemcpyAsync(x.h2, x.d2, (size_t)bytes, cudaMemcpyDeviceToHost, x.s2);
      }
      cudaStreamSynchronize(x.s);
      if(x.s2) cudaStreamSynchronize(x.s2);
    };
    for(Dev & x : dv) run(x, 1); // warm-up
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for(Dev & x : dv)
      th.emplace_back(
          [&, px = &x]
          {
            run(*px, reps);
            px->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
          });
    for(auto & t : th) t.join();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double per = (double)bytes * reps * (direction == 2 ? 2.0 : 1.0);
    agg = per * n_devices / wall / 1e9;
    if(per_device)
      for(int k = 0; k < n_devices; ++k) per_device[k] = per / dv[(size_t)k].seconds / 1e9;
  }
  for(Dev & x : dv)
  {
    cudaSetDevice(x.dev);
    if(x.s) cudaStreamDestroy(x.s);
    if(x.s2) cudaStreamDestroy(x.s2);
    if(x.d) cudaFree(x.d);
    if(x.d2) cudaFree(x.d2);
    if(x.h) cudaFreeHost(x.h);
    if(x.h2) cudaFreeHost(x.h2);
  }
  return agg;
}

} // extern "C"
