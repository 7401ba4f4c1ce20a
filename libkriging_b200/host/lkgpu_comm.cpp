// lkgpu_comm.cpp -- see lkgpu_comm.hpp.  POSIX sockets only.
#include "lkgpu_comm.hpp"

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>

namespace lkgpu {
namespace {

enum Op : int32_t { OP_HELLO = 1, OP_TICKET = 2, OP_GATHER = 3 };
struct Header {
  int32_t op, rank;
  int64_t key, count;  // count: doubles that follow (OP_GATHER)
};

void write_all(int fd, const void* buf, size_t len) {
  const char* p = static_cast<const char*>(buf);
  while (len > 0) {
    const ssize_t w = ::send(fd, p, len, MSG_NOSIGNAL);
    if (w <= 0) throw std::runtime_error("lkgpu::ShardComm: connection lost (send)");
    p += w;
    len -= (size_t)w;
  }
}
bool read_all(int fd, void* buf, size_t len) {  // false: peer closed before the first byte
  char* p = static_cast<char*>(buf);
  bool first = true;
  while (len > 0) {
    const ssize_t r = ::recv(fd, p, len, 0);
    if (r == 0 && first) return false;
    if (r <= 0) throw std::runtime_error("lkgpu::ShardComm: connection lost (recv)");
    first = false;
    p += r;
    len -= (size_t)r;
  }
  return true;
}
double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
void set_nodelay(int fd) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
}

}  // namespace

std::unique_ptr<ShardComm> ShardComm::from_env() {
  const char* ws = getenv("WORLD_SIZE");
  const int world = ws ? atoi(ws) : 1;
  if (world <= 1) return nullptr;
  const char* rk = getenv("RANK");
  const char* addr = getenv("MASTER_ADDR");
  const char* port = getenv("MASTER_PORT");
  const char* off = getenv("LKGPU_COMM_PORT_OFFSET");
  if (!rk || !port) throw std::runtime_error("lkgpu::ShardComm: WORLD_SIZE > 1 needs RANK and MASTER_PORT");
  return std::make_unique<ShardComm>(atoi(rk), world, addr ? addr : "127.0.0.1", atoi(port) + (off ? atoi(off) : 17));
}

ShardComm::ShardComm(int rank, int world, const std::string& addr, int port, double timeout_s)
    : m_rank(rank), m_world(world) {
  if (rank < 0 || rank >= world) throw std::runtime_error("lkgpu::ShardComm: rank out of range");
  if (rank == 0) {
    m_listen_fd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (m_listen_fd < 0) throw std::runtime_error("lkgpu::ShardComm: socket() failed");
    int one = 1;
    setsockopt(m_listen_fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in sa;
    memset(&sa, 0, sizeof(sa));
    sa.sin_family = AF_INET;
    sa.sin_addr.s_addr = htonl(INADDR_ANY);
    sa.sin_port = htons((uint16_t)port);
    if (::bind(m_listen_fd, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) != 0 || ::listen(m_listen_fd, world + 4) != 0) {
      ::close(m_listen_fd);
      throw std::runtime_error("lkgpu::ShardComm: cannot listen on port " + std::to_string(port));
    }
    m_server = std::thread([this]() { serve(); });
  }
  // every rank (0 included) connects as a client; the server may come up later than we do
  addrinfo hints, *res = nullptr;
  memset(&hints, 0, sizeof(hints));
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  const std::string host = rank == 0 ? "127.0.0.1" : addr;
  if (getaddrinfo(host.c_str(), std::to_string(port).c_str(), &hints, &res) != 0 || !res)
    throw std::runtime_error("lkgpu::ShardComm: cannot resolve " + host);
  const double t_end = now_s() + timeout_s;
  while (true) {
    m_fd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (m_fd >= 0 && ::connect(m_fd, res->ai_addr, res->ai_addrlen) == 0) break;
    if (m_fd >= 0) ::close(m_fd);
    m_fd = -1;
    if (now_s() > t_end) {
      freeaddrinfo(res);
      throw std::runtime_error("lkgpu::ShardComm: rank 0 not reachable at " + host + ":" + std::to_string(port));
    }
    usleep(50 * 1000);
  }
  freeaddrinfo(res);
  set_nodelay(m_fd);
  Header h{OP_HELLO, m_rank, 0, 0};
  write_all(m_fd, &h, sizeof(h));
  barrier();  // returns once all ranks are connected
}

ShardComm::~ShardComm() {
  if (m_fd >= 0) ::close(m_fd);
  m_stop = true;
  if (m_server.joinable()) m_server.join();
  if (m_listen_fd >= 0) ::close(m_listen_fd);
}

long long ShardComm::next_ticket(long long key) {
  std::lock_guard<std::mutex> lk(m_mutex);
  Header h{OP_TICKET, m_rank, key, 0};
  write_all(m_fd, &h, sizeof(h));
  int64_t t = -1;
  if (!read_all(m_fd, &t, sizeof(t))) throw std::runtime_error("lkgpu::ShardComm: server closed the connection");
  return t;
}

std::vector<double> ShardComm::allgather(const std::vector<double>& mine, std::vector<long long>* counts) {
  std::lock_guard<std::mutex> lk(m_mutex);
  Header h{OP_GATHER, m_rank, 0, (int64_t)mine.size()};
  write_all(m_fd, &h, sizeof(h));
  if (!mine.empty()) write_all(m_fd, mine.data(), mine.size() * sizeof(double));
  std::vector<int64_t> cnt(m_world);
  if (!read_all(m_fd, cnt.data(), cnt.size() * sizeof(int64_t)))
    throw std::runtime_error("lkgpu::ShardComm: server closed the connection");
  size_t total = 0;
  for (int64_t c : cnt) total += (size_t)c;
  std::vector<double> all(total);
  if (total > 0) read_all(m_fd, all.data(), total * sizeof(double));
  if (counts) counts->assign(cnt.begin(), cnt.end());
  return all;
}

void ShardComm::barrier() { allgather({}); }

// ---- rank 0's server thread --------------------------------------------------------------------------------
void ShardComm::serve() {
  std::vector<int> fds;                 // accepted connections
  std::vector<int> rank_of;             // rank behind each connection (-1 until HELLO)
  std::map<int64_t, int64_t> counters;  // ticket counters by key
  std::vector<std::vector<double>> parts(m_world);
  std::vector<bool> have(m_world, false);
  std::vector<int> fd_of_rank(m_world, -1);
  int n_have = 0, n_closed = 0;
  bool lost = false;  // a peer went away: no later all-gather can complete, the others must fail instead of waiting
  try {
    while (true) {
      std::vector<pollfd> pf;
      pf.push_back({m_listen_fd, POLLIN, 0});
      for (int fd : fds) pf.push_back({fd, (short)(fd >= 0 ? POLLIN : 0), 0});
      const int r = ::poll(pf.data(), pf.size(), 100);
      if (r < 0) break;
      if (m_stop && (n_closed >= (int)fds.size())) break;
      if (r == 0) continue;
      if (pf[0].revents & POLLIN) {
        const int c = ::accept(m_listen_fd, nullptr, nullptr);
        if (c >= 0) {
          set_nodelay(c);
          fds.push_back(c);
          rank_of.push_back(-1);
        }
      }
      for (size_t i = 0; i + 1 < pf.size() && i < fds.size(); ++i) {
        if (fds[i] < 0 || !(pf[i + 1].revents & (POLLIN | POLLHUP | POLLERR))) continue;
        Header h;
        if (!read_all(fds[i], &h, sizeof(h))) {  // peer closed
          ::close(fds[i]);
          fds[i] = -1;
          ++n_closed;
          lost = true;
          if (n_have > 0) throw std::runtime_error("a process left during an all-gather");
          continue;
        }
        if (h.op == OP_HELLO) {
          if (h.rank < 0 || h.rank >= m_world) throw std::runtime_error("bad rank in HELLO");
          rank_of[i] = h.rank;
          fd_of_rank[h.rank] = fds[i];
        } else if (h.op == OP_TICKET) {
          const int64_t t = counters[h.key]++;
          write_all(fds[i], &t, sizeof(t));
        } else if (h.op == OP_GATHER) {
          const int rk = rank_of[i];
          if (rk < 0 || have[rk]) throw std::runtime_error("unexpected GATHER");
          if (lost) throw std::runtime_error("all-gather after a process left");
          parts[rk].resize((size_t)h.count);
          if (h.count > 0) read_all(fds[i], parts[rk].data(), (size_t)h.count * sizeof(double));
          have[rk] = true;
          if (++n_have == m_world) {
            std::vector<int64_t> cnt(m_world);
            std::vector<double> all;
            for (int q = 0; q < m_world; ++q) {
              cnt[q] = (int64_t)parts[q].size();
              all.insert(all.end(), parts[q].begin(), parts[q].end());
            }
            for (int q = 0; q < m_world; ++q) {
              write_all(fd_of_rank[q], cnt.data(), cnt.size() * sizeof(int64_t));
              if (!all.empty()) write_all(fd_of_rank[q], all.data(), all.size() * sizeof(double));
              have[q] = false;
              parts[q].clear();
            }
            n_have = 0;
          }
        }
      }
    }
  } catch (const std::exception&) {
    // a broken connection, or a process that left before the exchange was over (it failed: the normal exit is after the
    // last barrier), ends the server; the clients see their sockets close and raise instead of waiting for ever
  }
  for (int fd : fds)
    if (fd >= 0) ::close(fd);
}

}  // namespace lkgpu
