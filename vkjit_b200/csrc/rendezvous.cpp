// rendezvous.cpp — the one exchange a multi-GPU job needs before its first kernel, without torch / MPI:
// rank 0's 128-byte NCCL unique id to everyone, and every rank's 64-byte cudaIpc mailbox handle to everyone.
//
// One process per GPU, started by ANY launcher that sets RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR and MASTER_PORT
// (torchrun, an mpirun wrapper, a shell loop).  Rank 0 listens on MASTER_ADDR : ($VKJIT_RDZV_PORT or MASTER_PORT + 1
// — MASTER_PORT itself belongs to torch's own store when a torch process is around), the others connect (retrying
// until rank 0 is up), send {rank, blob}, and get {root blob, all blobs in rank order} back once EVERY rank has
// arrived — the reply doubles as a barrier.  Plain blocking sockets, one node (cudaIpc handles do not travel further).
// The reference has no counterpart (single device, backend/vulkan/device.rs:162-200).
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.h"
#include "dist.h"

namespace vkjit {
namespace dist {
namespace {

bool send_all(int fd, const void* p, size_t n) {
  const char* c = (const char*)p;
  while (n) {
    const ssize_t k = ::send(fd, c, n, MSG_NOSIGNAL);
    if (k <= 0) { if (k < 0 && errno == EINTR) continue; return false; }
    c += k; n -= (size_t)k;
  }
  return true;
}
bool recv_all(int fd, void* p, size_t n) {
  char* c = (char*)p;
  while (n) {
    const ssize_t k = ::recv(fd, c, n, 0);
    if (k <= 0) { if (k < 0 && errno == EINTR) continue; return false; }
    c += k; n -= (size_t)k;
  }
  return true;
}
void set_timeouts(int fd, double secs) {
  timeval tv;
  tv.tv_sec = (time_t)secs;
  tv.tv_usec = (suseconds_t)((secs - (double)tv.tv_sec) * 1e6);
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
  setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
}
struct Fd {
  int fd = -1;
  ~Fd() { if (fd >= 0) ::close(fd); }
};
constexpr uint32_t kHello = 0x564B5244u;  // "VKRD"

}  // namespace

void rendezvous(int rank, int world, const char* addr, int port, const void* mine, size_t blob_bytes, void* root_blob,
                size_t root_bytes, void* all_out, double timeout_s) {
  if (world < 1 || rank < 0 || rank >= world) fail(VKJIT_ERR_INVALID, "rendezvous: bad rank/world");
  if (world == 1) { memcpy(all_out, mine, blob_bytes); return; }
  const uint64_t t_end = now_ns() + (uint64_t)(timeout_s * 1e9);
  sockaddr_in sa;
  memset(&sa, 0, sizeof sa);
  sa.sin_family = AF_INET;
  sa.sin_port = htons((uint16_t)port);
  if (inet_pton(AF_INET, addr, &sa.sin_addr) != 1) {
    hostent* he = gethostbyname(addr);
    if (!he || he->h_addrtype != AF_INET) fail(VKJIT_ERR_DIST, std::string("rendezvous: cannot resolve ") + addr);
    memcpy(&sa.sin_addr, he->h_addr_list[0], sizeof sa.sin_addr);
  }
  const size_t reply_bytes = root_bytes + (size_t)world * blob_bytes;
  if (rank == 0) {
    Fd ls;
    ls.fd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (ls.fd < 0) fail(VKJIT_ERR_DIST, "rendezvous: socket() failed");
    int one = 1;
    setsockopt(ls.fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in any = sa;
    any.sin_addr.s_addr = htonl(INADDR_ANY);
    if (::bind(ls.fd, (sockaddr*)&any, sizeof any) != 0)
      fail(VKJIT_ERR_DIST, "rendezvous: cannot bind port " + std::to_string(port) + " (" + strerror(errno) + "); set VKJIT_RDZV_PORT");
    if (::listen(ls.fd, world) != 0) fail(VKJIT_ERR_DIST, "rendezvous: listen() failed");
    set_timeouts(ls.fd, timeout_s);
    std::vector<char> reply(reply_bytes);
    memcpy(reply.data(), root_blob, root_bytes);
    memcpy(reply.data() + root_bytes, mine, blob_bytes);
    std::vector<Fd> peers((size_t)world);
    int arrived = 1;
    while (arrived < world) {
      if (now_ns() > t_end) fail(VKJIT_ERR_DIST, "rendezvous: only " + std::to_string(arrived) + " of " + std::to_string(world) + " ranks arrived");
      const int c = ::accept(ls.fd, nullptr, nullptr);
      if (c < 0) { if (errno == EINTR || errno == EAGAIN || errno == EWOULDBLOCK) continue; fail(VKJIT_ERR_DIST, "rendezvous: accept() failed"); }
      set_timeouts(c, 5.0);  // the hello arrives at once or not at all: a silent stray connection must not stall the job
      uint32_t hdr[3];
      std::vector<char> blob(blob_bytes);
      if (!recv_all(c, hdr, sizeof hdr) || hdr[0] != kHello || (int)hdr[2] != world || hdr[1] == 0 || (int)hdr[1] >= world ||
          peers[hdr[1]].fd >= 0 || !recv_all(c, blob.data(), blob_bytes)) {
        ::close(c);  // a stray connection, a rank of another job, or a duplicate: ignore
        continue;
      }
      set_timeouts(c, timeout_s);
      peers[hdr[1]].fd = c;
      memcpy(reply.data() + root_bytes + (size_t)hdr[1] * blob_bytes, blob.data(), blob_bytes);
      ++arrived;
    }
    for (int r = 1; r < world; ++r)
      if (!send_all(peers[(size_t)r].fd, reply.data(), reply_bytes)) fail(VKJIT_ERR_DIST, "rendezvous: reply to rank " + std::to_string(r) + " failed");
    memcpy(all_out, reply.data() + root_bytes, (size_t)world * blob_bytes);
    // wait until every rank has read the reply (they close their end): nobody leaves before all have the data
    for (int r = 1; r < world; ++r) { char b; (void)::recv(peers[(size_t)r].fd, &b, 1, 0); }
    return;
  }
  for (;;) {
    Fd s;
    s.fd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (s.fd < 0) fail(VKJIT_ERR_DIST, "rendezvous: socket() failed");
    if (::connect(s.fd, (sockaddr*)&sa, sizeof sa) == 0) {
      set_timeouts(s.fd, timeout_s);
      const uint32_t hdr[3] = {kHello, (uint32_t)rank, (uint32_t)world};
      std::vector<char> reply(reply_bytes);
      if (send_all(s.fd, hdr, sizeof hdr) && send_all(s.fd, mine, blob_bytes) && recv_all(s.fd, reply.data(), reply_bytes)) {
        memcpy(root_blob, reply.data(), root_bytes);
        memcpy(all_out, reply.data() + root_bytes, (size_t)world * blob_bytes);
        return;
      }
    }
    if (now_ns() > t_end) fail(VKJIT_ERR_DIST, std::string("rendezvous: rank 0 not reachable at ") + addr + ":" + std::to_string(port));
    ::usleep(20000);  // rank 0 is not listening yet (or dropped us): retry
  }
}

namespace {
int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
}  // namespace

void rendezvous_endpoint(std::string& addr, int& port) {
  const char* a = getenv("MASTER_ADDR");
  addr = (a && *a) ? a : "127.0.0.1";
  port = env_int("VKJIT_RDZV_PORT", env_int("MASTER_PORT", 29500) + 1);
}

}  // namespace dist
}  // namespace vkjit
