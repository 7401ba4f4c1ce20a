// lkgpu_comm_selftest.cpp -- CPU-only check of lkgpu::ShardComm (tests/test_host_comm.py launches WORLD_SIZE of these):
// a ticket queue drained by 3 threads per process hands out every index exactly once, and all-gathers of ragged
// contributions arrive complete and in rank order on every rank.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <thread>

#include "lkgpu_comm.hpp"

int main() {
  try {
    auto comm = lkgpu::ShardComm::from_env();
    if (!comm) { printf("{\"error\": \"WORLD_SIZE <= 1\"}\n"); return 2; }
    // failure injection (tests/test_host_comm.py): this rank leaves before the exchange; the others must fail, not hang
    if (const char* die = getenv("LKGPU_COMM_SELFTEST_DIE"); die && atoi(die) == comm->rank()) {
      printf("{\"rank\": %d, \"left_early\": true}\n", comm->rank());
      return 3;
    }
    const int total = 50;
    std::vector<double> taken;
    std::mutex mu;
    std::vector<std::thread> pool;
    std::exception_ptr failure;
    for (int w = 0; w < 3; ++w)
      pool.emplace_back([&]() {
        try {
          for (long long t = comm->next_ticket(7); t < total; t = comm->next_ticket(7)) {
            std::lock_guard<std::mutex> lk(mu);
            taken.push_back((double)t);
            std::this_thread::sleep_for(std::chrono::milliseconds(1 + comm->rank()));
          }
        } catch (...) {
          std::lock_guard<std::mutex> lk(mu);
          if (!failure) failure = std::current_exception();
        }
      });
    for (auto& t : pool) t.join();
    if (failure) std::rethrow_exception(failure);
    std::vector<long long> counts;
    std::vector<double> all = comm->allgather(taken, &counts);
    std::sort(all.begin(), all.end());
    bool ok = (int)all.size() == total;
    for (int i = 0; ok && i < total; ++i) ok = all[i] == (double)i;
    // ragged gather: rank r contributes r + 1 copies of r
    std::vector<double> mine(comm->rank() + 1, (double)comm->rank());
    std::vector<double> g = comm->allgather(mine, &counts);
    size_t pos = 0;
    for (int r = 0; ok && r < comm->world(); ++r) {
      ok = counts[r] == r + 1;
      for (int q = 0; ok && q <= r; ++q) ok = g[pos++] == (double)r;
    }
    // a second key starts from 0 again
    const long long t2 = comm->next_ticket(8);
    comm->barrier();
    printf("{\"rank\": %d, \"world\": %d, \"ok\": %s, \"taken\": %zu, \"ticket_key8\": %lld}\n", comm->rank(), comm->world(),
           ok ? "true" : "false", taken.size(), t2);
    return ok ? 0 : 1;
  } catch (const std::exception& e) {
    printf("{\"error\": \"%s\"}\n", e.what());
    return 1;
  }
}
