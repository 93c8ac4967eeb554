// dist.h — multi-GPU plumbing: one process per GPU, NCCL over NVLink for the only real exchange
// step on the path (combining per-GPU reduction partials).  The reference is single-device
// (libs/vkjit-core/src/backend/vulkan/device.rs:162-200); this is new.
#pragma once
#include <cstddef>
#include <cstdint>

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
// contiguous shards in units of 4 lanes (16-byte aligned); the last rank takes the ragged tail
void shard_range(size_t n, int rank, int world, size_t& lo, size_t& hi);

}  // namespace dist
}  // namespace vkjit
