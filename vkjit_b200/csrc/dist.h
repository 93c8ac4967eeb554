// dist.h — multi-GPU plumbing: one process per GPU, NCCL over NVLink for the only real exchange
// step on the path (combining per-GPU reduction partials).  The reference is single-device
// (libs/vkjit-core/src/backend/vulkan/device.rs:162-200); this is new.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include "prims.h"

namespace vkjit {
namespace dist {

bool active();
int rank();
int world();
void unique_id(void* out128);
void init(int rank, int world, const void* id128);
void shutdown();
// in-place all-reduce of `count` 4-byte elements on the backend stream; result replicated
void allreduce(void* buf, uint32_t ty, int red, size_t count);
// Fused path: peer-mapped mailboxes (cudaIpc).  mailbox_handle allocates the local mailbox and
// returns its 64-byte IPC handle; mailbox_open maps every rank's mailbox (world x 64 bytes, in rank
// order).  Once open, reductions combine inside the reduce kernel over NVLink instead of NCCL.
void mailbox_handle(void* out64);
void mailbox_open(const void* handles, int world);
bool p2p_enabled();
void set_p2p(bool on);  // switch between the fused mailbox path and NCCL (all ranks must agree)
prims::Mailbox next_mailbox();  // bumps the collective sequence number

// Torch-free bring-up (rendezvous.cpp): exchange over one TCP connection per rank to rank 0.  `mine` (blob_bytes) of
// every rank lands in all_out (world x blob_bytes, rank order) on every rank; root_blob (root_bytes) is rank 0's on
// entry and everyone's on return.  Returns only once all ranks have arrived.
void rendezvous(int rank, int world, const char* addr, int port, const void* mine, size_t blob_bytes, void* root_blob,
                size_t root_bytes, void* all_out, double timeout_s);
// init_env(): RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT (VKJIT_RDZV_PORT) from the environment:
// Backend::init(LOCAL_RANK) if needed, rendezvous, NCCL communicator, peer mailboxes.
void init_env();
void rendezvous_endpoint(std::string& addr, int& port);

// contiguous shards in units of 4 lanes (16-byte aligned); the last rank takes the ragged tail
void shard_range(size_t n, int rank, int world, size_t& lo, size_t& hi);

}  // namespace dist
}  // namespace vkjit
