// prims.h — hand-written sm_100a runtime primitives (no reference counterpart; the reference
// has no reductions, scans or compaction — SURVEY.md §2a, Appendix A.3).
#pragma once
#include <cstddef>
#include <cstdint>

namespace vkjit {
namespace prims {

// Scratch the primitives need between launches; owned by the backend, zero-initialised once.
struct Scratch {
  void* partials = nullptr;       // reduce: one 4-byte partial per CTA
  unsigned int* ticket = nullptr; // reduce: last-CTA election (self-resetting)
  uint64_t* tile_state = nullptr; // scan: look-back status, one 128-byte slot per tile (slot 0 reserved)
  size_t tile_state_words = 0;
};

// Peer-memory mailbox for the fused reduce + all-reduce (one process per GPU, NVLink P2P).
// Every GPU owns kMailSlots x kMailRanks 64-bit words {seq:32 | value:32}; rank r publishes its
// partial of collective number `seq` into word [seq % kMailSlots][r] of EVERY GPU's mailbox with
// one system-scope store per peer, then combines the `world` words of its own mailbox in rank
// order (deterministic, identical on every rank).  Value and sequence number travel in one
// 64-bit store, so no fence is needed.
constexpr int kMailSlots = 64;
constexpr int kMailRanks = 8;
struct Mailbox {
  uint64_t* const* peers = nullptr;  // DEVICE table: mailbox base of every rank (kept out of the kernel
                                     // parameter so no thread needs a local copy to index it)
  uint64_t* local = nullptr;         // this rank's mailbox
  int rank = 0, world = 1;
  uint32_t seq = 0;            // >= 1; identical on all ranks for the same collective
  uint32_t* fault = nullptr;   // host-mapped word: receives `seq` when the watchdog below expires (checked at every sync)
  uint64_t timeout_ns = 0;     // how long an exchange waits for a peer ($VKJIT_DIST_TIMEOUT_S, default 120 s); 0 = forever
};

constexpr int kReduceThreads = 512;
constexpr int kReduceMaxCtas = 2048;
constexpr int kScanThreads = 1024;
constexpr int kScanMinTile = 12288;  // look-back tile (smallest tile of the kernels in scan.cu); sizes the status array

// number of 8-byte words `tile_state` must hold for n lanes
constexpr int kStatusWordsPerTile = 16;  // one 128-byte line of 64-bit words per tile (kStatusStride in scan_common.cuh)
size_t scan_state_words(size_t n, size_t tile = kScanMinTile);

// out[0] = reduce(in[0..n)).  ty: VKJIT_TY_{U32,I32,F32}; red: VKJIT_RED_*.  One launch:
// vectorised grid-stride partials -> warp shuffle -> shared-memory tree -> last CTA folds the
// per-CTA partials in a fixed order (deterministic for a given n and grid).
// `mailbox` (optional, multi-GPU): the last CTA also performs the all-reduce — it publishes the per-GPU
// partial into every peer's mailbox over NVLink and folds all ranks' partials in rank order.
// `flags`: kReduceWaitFirst = plain stream order (the default); 0 = the streaming phase may overlap the tail of the
// kernel launched right before on the stream (programmatic dependent launch) — only when the caller can prove that
// `in` is not written by a kernel that may still be running (see capi.cpp: vkjit_reduce).
constexpr uint32_t kReduceWaitFirst = 1u;
void reduce(int red, uint32_t ty, const void* in, size_t n, void* out, const Scratch& sc, int sm_count, void* stream,
            const Mailbox* mailbox = nullptr, uint32_t flags = kReduceWaitFirst);
// $VKJIT_REDUCE_TRACE=1: every reduce launch records eight %globaltimer stamps (kReduceTraceWords per launch, ring of
// kReduceTraceLaunches launches): [0] CTA 0 enters, then in the LAST CTA (the one that folds): [1] its streaming phase
// done, [2] previous kernel complete (griddepcontrol.wait returned), [3] ticket taken, [4] partials folded,
// [5] peer exchange done, [6] launch number, [7] world.  Copies the ring to `out` (host); returns launches recorded.
constexpr int kReduceTraceWords = 8, kReduceTraceLaunches = 4096;
size_t reduce_trace_dump(unsigned long long* out, size_t cap_words, void* stream);
// Stand-alone exchange: out[0] = combine over ranks of out[0] (used when a rank's shard is empty).
void p2p_allreduce(int red, uint32_t ty, void* out, const Mailbox& mailbox, void* stream);

// Single-pass decoupled look-back prefix sum (mod 2^32) over u32 words.
// `initial` (optional, device): prefix carried in from outside — the sharded scan passes the sum of the lower
// ranks' totals, which becomes the predecessor of tile 0.
void prefix_sum(const uint32_t* in, uint32_t* out, size_t n, bool exclusive, const Scratch& sc, int sm_count, void* stream,
                const uint32_t* initial = nullptr);
// Sharded scan support: exclusive scan over ranks of one u32 per GPU through the peer mailboxes, and the two
// tiny kernels of the NCCL path (one-hot vector of totals -> all-reduce -> prefix of this rank).
void p2p_exscan_u32(const uint32_t* mine, uint32_t* out, const Mailbox& mailbox, void* stream);
void p2p_exscan_total_u32(const uint32_t* mine, uint32_t* out2, const Mailbox& mailbox, void* stream);  // out2 = {below, total}
void one_hot_u32(const uint32_t* mine, int rank, int world, uint32_t* v, void* stream);
void prefix_of_rank_u32(const uint32_t* v, int rank, uint32_t* out, void* stream);

// Stream compaction: lanes whose mask word is non-zero, stable order.  values == nullptr writes the
// lane index (+ *index_base when given: the global index of lane 0 of a sharded mask).  *count_out (device)
// receives the number of selected lanes.
void compress(const uint32_t* mask, const uint32_t* values, uint32_t* out, uint32_t* count_out, size_t n,
              const Scratch& sc, int sm_count, void* stream, const uint32_t* index_base = nullptr);

// out[i] = value for i in [0, n)
void fill_u32(uint32_t* out, uint32_t value, size_t n, void* stream);

}  // namespace prims
}  // namespace vkjit
