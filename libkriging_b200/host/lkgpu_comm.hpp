// lkgpu_comm.hpp -- exchange between the C++ host processes of one sharded fit (one process per GPU).
//
// The path shards along multistart rows: every process runs whole L-BFGS-B starts on its own GPU
// and the only data that crosses processes is (a) which start a free worker takes next and (b) one row per start --
// objective value, success flag, evaluation count, gamma -- for the reference's argmin (Kriging.cpp:2097-2110).
// That is a few hundred bytes per fit, so the exchange is a plain TCP star around rank 0 (no MPI, no NCCL, no
// GPU traffic): rank 0 runs a small server thread, every rank (rank 0 included) is a client of it.
//   next_ticket(key)     shared counter per key: the dynamic start queue (a straggling start does not hold idle
//                        the processes that finished their share)
//   allgather(rows)      every rank's doubles, concatenated in rank order, on every rank
//   barrier()
// Rendezvous from the launcher's environment (torchrun's names): RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT
// (+ LKGPU_COMM_PORT_OFFSET, default 17, so that a torch.distributed store on MASTER_PORT is left alone).
#pragma once
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace lkgpu {

class ShardComm {
 public:
  ShardComm(int rank, int world, const std::string& addr, int port, double timeout_s = 120.0);
  ~ShardComm();
  ShardComm(const ShardComm&) = delete;
  ShardComm& operator=(const ShardComm&) = delete;
  // nullptr when WORLD_SIZE is absent or 1
  static std::unique_ptr<ShardComm> from_env();

  int rank() const { return m_rank; }
  int world() const { return m_world; }
  // 0, 1, 2, ... in the order the requests reach rank 0; thread-safe
  long long next_ticket(long long key);
  // counts[r] = number of doubles rank r contributed (optional)
  std::vector<double> allgather(const std::vector<double>& mine, std::vector<long long>* counts = nullptr);
  void barrier();

 private:
  void serve();
  int m_rank, m_world;
  int m_fd = -1;         // this rank's connection to the server
  int m_listen_fd = -1;  // rank 0 only
  std::thread m_server;
  std::mutex m_mutex;    // one request / reply at a time per process
  volatile bool m_stop = false;
};

}  // namespace lkgpu
