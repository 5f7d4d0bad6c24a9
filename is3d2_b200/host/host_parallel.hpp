// Small std::thread helpers shared by the host-side ingestion (surface.cpp) and result writers (emission.cpp).
#pragma once

#include <cstdlib>
#include <thread>
#include <vector>

namespace is3dhost {

// number of worker threads for `work_items` independent items: hardware threads, overridable with `env_name`
inline int host_threads(size_t work_items, const char *env_name)
{
  unsigned hw = std::thread::hardware_concurrency();
  if (const char *v = getenv(env_name)) hw = (unsigned)atoi(v);
  size_t t = hw ? hw : 1;
  if (t > work_items) t = work_items;
  if (t > 256) t = 256;
  return (int)(t < 1 ? 1 : t);
}

// fn(t) on nthreads threads (the caller's thread is worker 0)
template <class Fn>
void parallel_for(int nthreads, Fn fn)
{
  if (nthreads <= 1) { fn(0); return; }
  std::vector<std::thread> pool;
  pool.reserve(nthreads - 1);
  for (int t = 1; t < nthreads; t++) pool.emplace_back(fn, t);
  fn(0);
  for (auto &th : pool) th.join();
}

}  // namespace is3dhost
