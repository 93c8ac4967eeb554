/*
 * vkjit_b200.h — C ABI of the B200-native execution backend for vkjit.
 *
 * This is the drop-in boundary: exactly the calls a vkjit
 * front-end (vkjit-rust `Var`/`eval!`, vkjit-python `Var`/`eval`) makes on
 * `vkjit_core::Ir` and, below it, on `Backend`/`Array`.  Every entry point cites
 * the reference interface it replaces (paths relative to the reference tree,
 * libs/vkjit-core/src/...).  Plain pointers and sizes only; no C++ or torch
 * types cross this line.
 *
 * Conventions
 *   - every function returns a vkjit_status (0 = OK); the reference panics on
 *     every misuse, this ABI returns a status and keeps a thread-local message
 *     readable through vkjit_last_error().
 *   - a vkjit_var is the reference's `VarId` (index into `Ir.vars`,
 *     internal.rs:79-80).  Handles are reference counted exactly like the
 *     reference: constructors return a var with ref_count == 1 and bump the
 *     count of every dependency (internal.rs:186-209); front-ends call
 *     vkjit_inc_ref / vkjit_dec_ref from Clone / Drop
 *     (vkjit-rust/src/types.rs:128-140).
 *   - all scalar element types occupy 4 bytes in device memory, Bool included
 *     (vartype.rs:45-64, crevice std140).
 *   - trace construction never needs a device.  vkjit_init / array upload /
 *     eval / read need a B200; without one they fail with VKJIT_ERR_NO_DEVICE.
 *     There is no CPU fallback.
 */
#ifndef VKJIT_B200_H
#define VKJIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VKJIT_B200_ABI_VERSION 1

typedef int32_t vkjit_status;
enum {
  VKJIT_OK = 0,
  VKJIT_ERR_INVALID = 1,     /* bad handle / argument (reference: index panic)              */
  VKJIT_ERR_TYPE = 2,        /* type mismatch (reference: assert internal.rs:232, :447)     */
  VKJIT_ERR_SIZE = 3,        /* kernel size mismatch / size-less schedule (internal.rs:697-706, :1202) */
  VKJIT_ERR_UNSUPPORTED = 4, /* reference: unimplemented!()                                  */
  VKJIT_ERR_NO_DEVICE = 5,   /* no usable CUDA device / backend not initialised             */
  VKJIT_ERR_CUDA = 6,        /* CUDA driver/runtime error                                   */
  VKJIT_ERR_COMPILE = 7,     /* NVRTC rejected generated code                               */
  VKJIT_ERR_DIST = 8         /* NCCL / multi-GPU error                                      */
};

/* VarType (vartype.rs:24-33).  Numeric order == the reference's derive(Ord)
 * promotion order Struct < Void < Bool < U32 < I32 < F32.  Struct types are
 * interned per Ir by vkjit_type_struct and compare below every scalar. */
typedef uint32_t vkjit_type;
enum {
  VKJIT_TY_VOID = 1,
  VKJIT_TY_BOOL = 2,
  VKJIT_TY_U32 = 3,
  VKJIT_TY_I32 = 4,
  VKJIT_TY_F32 = 5,
  VKJIT_TY_STRUCT_BASE = 16
};

typedef uint32_t vkjit_var; /* VarId, internal.rs:79-80 */
typedef struct vkjit_ir vkjit_ir; /* one `Ir`, internal.rs:126-131 */

/* Binary ops.  0..9 are the reference's `Bop` (internal.rs:31-43); >= 16 are
 * extensions with no reference implementation. */
enum {
  VKJIT_BOP_ADD = 0, VKJIT_BOP_SUB = 1, VKJIT_BOP_MUL = 2, VKJIT_BOP_DIV = 3,
  VKJIT_BOP_LT = 4, VKJIT_BOP_GT = 5, VKJIT_BOP_EQ = 6, VKJIT_BOP_LEQ = 7,
  VKJIT_BOP_GEQ = 8, VKJIT_BOP_NEQ = 9,
  VKJIT_BOP_AND = 16, VKJIT_BOP_OR = 17, VKJIT_BOP_XOR = 18,
  VKJIT_BOP_SHL = 19, VKJIT_BOP_SHR = 20, VKJIT_BOP_MIN = 21, VKJIT_BOP_MAX = 22
};
/* Unary ops (all extensions). */
enum {
  VKJIT_UOP_NEG = 0, VKJIT_UOP_ABS = 1, VKJIT_UOP_NOT = 2, VKJIT_UOP_SQRT = 3,
  VKJIT_UOP_EXP = 4, VKJIT_UOP_LOG = 5, VKJIT_UOP_SIN = 6, VKJIT_UOP_COS = 7
};
/* Horizontal reductions (extension). */
enum { VKJIT_RED_SUM = 0, VKJIT_RED_MIN = 1, VKJIT_RED_MAX = 2 };

/* ------------------------------------------------------------------ */
/* Lifecycle — Backend::create (backend/mod.rs:20), Device::create      */
/* (backend/vulkan/device.rs:74-225), Device::drop (device.rs:228-240). */
/* ------------------------------------------------------------------ */
/* Bind this process to CUDA device `device` (-1: use $LOCAL_RANK, else 0),
 * create the backend stream and the cudaMallocAsync pool.  Idempotent. */
vkjit_status vkjit_init(int32_t device);
vkjit_status vkjit_shutdown(void);
/* 1 if vkjit_init succeeded in this process. */
int32_t vkjit_is_initialized(void);
const char* vkjit_last_error(void);
uint32_t vkjit_abi_version(void);
/* The CUstream/cudaStream_t every launch and copy is ordered on. */
vkjit_status vkjit_stream(void** out_stream);
/* CUDA device ordinal the backend was initialised on. */
vkjit_status vkjit_device(int32_t* out_device);
/* cudaStreamSynchronize on the backend stream (reference: vkWaitForFences +
 * vkDeviceWaitIdle inside every execute, backend/vulkan/mod.rs:185-190; here
 * eval is asynchronous and only readback/sync waits). */
vkjit_status vkjit_sync(void);
/* Pinned host staging for front-ends that want zero-copy-like readback. */
vkjit_status vkjit_host_alloc(size_t bytes, void** out_ptr);
vkjit_status vkjit_host_free(void* ptr);

/* ------------------------------------------------------------------ */
/* Ir lifetime — Ir::new (internal.rs:168-182).                         */
/* ------------------------------------------------------------------ */
vkjit_status vkjit_ir_create(vkjit_ir** out_ir);
vkjit_status vkjit_ir_destroy(vkjit_ir* ir);

/* ------------------------------------------------------------------ */
/* Types — VarType::Struct (vartype.rs:24-27).                          */
/* ------------------------------------------------------------------ */
vkjit_status vkjit_type_struct(vkjit_ir* ir, const vkjit_type* elems, size_t n, vkjit_type* out_ty);
vkjit_status vkjit_type_struct_len(vkjit_ir* ir, vkjit_type ty, size_t* out_n);
vkjit_status vkjit_type_struct_elem(vkjit_ir* ir, vkjit_type ty, size_t i, vkjit_type* out_elem);

/* ------------------------------------------------------------------ */
/* Trace constructors — one per `Ir` method.                            */
/* ------------------------------------------------------------------ */
/* Ir::const_f32/i32/u32/bool, internal.rs:301-312 */
vkjit_status vkjit_const_f32(vkjit_ir* ir, float v, vkjit_var* out);
vkjit_status vkjit_const_i32(vkjit_ir* ir, int32_t v, vkjit_var* out);
vkjit_status vkjit_const_u32(vkjit_ir* ir, uint32_t v, vkjit_var* out);
vkjit_status vkjit_const_bool(vkjit_ir* ir, int32_t v, vkjit_var* out);
/* Ir::array_f32/i32/u32, internal.rs:313-348 -> Backend::create_array_from_slice
 * (backend/mod.rs:21, vulkan/mod.rs:56-73).  `data` is HOST memory holding n
 * 4-byte elements; it is copied to a fresh device array (H2D on the backend
 * stream; the call returns after the host buffer may be reused). */
vkjit_status vkjit_array_f32(vkjit_ir* ir, const float* data, size_t n, vkjit_var* out);
vkjit_status vkjit_array_i32(vkjit_ir* ir, const int32_t* data, size_t n, vkjit_var* out);
vkjit_status vkjit_array_u32(vkjit_ir* ir, const uint32_t* data, size_t n, vkjit_var* out);
/* Extension: Bool array from n 4-byte words (0 / non-0). */
vkjit_status vkjit_array_bool(vkjit_ir* ir, const uint32_t* data, size_t n, vkjit_var* out);
/* Extension: uninitialised device array of n elements — Backend::create_array
 * (backend/mod.rs:22) exposed as a Binding var. */
vkjit_status vkjit_array_empty(vkjit_ir* ir, vkjit_type ty, size_t n, vkjit_var* out);
/* Extension: zero-copy view of FOREIGN device memory (n 4-byte elements at device_ptr) as a Binding var.
 * Not owned: never freed here; the caller keeps it alive while the var lives and orders its producer before
 * the backend stream (vkjit_stream).  Pointers that are not 16-byte aligned select the scalar kernel variant. */
vkjit_status vkjit_array_wrap_device(vkjit_ir* ir, vkjit_type ty, uint64_t device_ptr, size_t n, vkjit_var* out);
/* Ir::arange, internal.rs:235-237 */
/* Same, with an owner: release(ctx) is called exactly once when the view's array is dropped (the var is freed or
 * re-evaluated), after the library has synchronised its stream and released the Ir lock. */
vkjit_status vkjit_array_wrap_device_owned(vkjit_ir* ir, vkjit_type ty, uint64_t device_ptr, size_t n,
                                           void (*release)(void*), void* ctx, vkjit_var* out);
/* DLPack ("dltensor" ABI, DLManagedTensor*).  to_dlpack: zero-copy export of an evaluated var — the tensor holds
 * one reference on the var until its deleter runs (the Ir must outlive it); the backend stream is synchronised
 * first.  from_dlpack: zero-copy import of a 1-D contiguous f32/i32/u32 CUDA tensor; on success the library owns
 * the tensor and calls its deleter when the var's array is dropped; on failure the caller keeps ownership.
 * vkjit_dlpack_delete runs a tensor's deleter (for an exported tensor nobody consumed). */
vkjit_status vkjit_var_to_dlpack(vkjit_ir* ir, vkjit_var id, void** out_managed_tensor);
vkjit_status vkjit_var_from_dlpack(vkjit_ir* ir, void* managed_tensor, vkjit_var* out);
void vkjit_dlpack_delete(void* managed_tensor);
vkjit_status vkjit_arange(vkjit_ir* ir, vkjit_type ty, size_t n, vkjit_var* out);
/* Ir::linspace, internal.rs:238-246 (endpoint excluded) */
vkjit_status vkjit_linspace(vkjit_ir* ir, vkjit_type ty, vkjit_var start, vkjit_var stop, size_t n, vkjit_var* out);
/* Ir::zeros / Ir::ones, internal.rs:247-282 */
vkjit_status vkjit_zeros(vkjit_ir* ir, vkjit_type ty, vkjit_var* out);
vkjit_status vkjit_ones(vkjit_ir* ir, vkjit_type ty, vkjit_var* out);
/* Ir::cast, internal.rs:283-290 (returns `src` itself, no new ref, when the type already matches) */
vkjit_status vkjit_cast(vkjit_ir* ir, vkjit_var src, vkjit_type ty, vkjit_var* out);
/* Ir::{add,sub,mul,div,lt,gt,eq,leq,geq,neq} via bop!, internal.rs:146-166, :218-227;
 * also the extension kinds (VKJIT_BOP_AND ...). */
vkjit_status vkjit_bop(vkjit_ir* ir, int32_t kind, vkjit_var lhs, vkjit_var rhs, vkjit_var* out);
/* Extension unary ops. */
vkjit_status vkjit_uop(vkjit_ir* ir, int32_t kind, vkjit_var src, vkjit_var* out);
/* Extension: reinterpret the 32 payload bits as another 4-byte scalar type. */
vkjit_status vkjit_bitcast(vkjit_ir* ir, vkjit_var src, vkjit_type ty, vkjit_var* out);
/* Ir::select, internal.rs:229-234 (asserts lhs type == rhs type) */
vkjit_status vkjit_select(vkjit_ir* ir, vkjit_var cond, vkjit_var lhs, vkjit_var rhs, vkjit_var* out);
/* Ir::struct_init / getattr / setattr, internal.rs:291-300, :349-367 */
vkjit_status vkjit_struct_init(vkjit_ir* ir, const vkjit_var* elems, size_t n, vkjit_var* out);
vkjit_status vkjit_getattr(vkjit_ir* ir, vkjit_var src, size_t idx, vkjit_var* out);
vkjit_status vkjit_setattr(vkjit_ir* ir, vkjit_var dst, vkjit_var src, size_t idx, vkjit_var* out);
/* Ir::gather, internal.rs:368-378.  has_active == 0: unmasked. */
vkjit_status vkjit_gather(vkjit_ir* ir, vkjit_var src, vkjit_var idx, int32_t has_active, vkjit_var active, vkjit_var* out);
/* Ir::scatter, internal.rs:379-400 (argument order src, dst, idx, active as in the reference) */
vkjit_status vkjit_scatter(vkjit_ir* ir, vkjit_var src, vkjit_var dst, vkjit_var idx, int32_t has_active, vkjit_var active, vkjit_var* out);
/* Extension: dst[idx[i]] += src[i] atomically. */
vkjit_status vkjit_scatter_add(vkjit_ir* ir, vkjit_var src, vkjit_var dst, vkjit_var idx, int32_t has_active, vkjit_var active, vkjit_var* out);

/* ------------------------------------------------------------------ */
/* Introspection / lifetime.                                            */
/* ------------------------------------------------------------------ */
/* Var::ty, internal.rs:115-119 */
vkjit_status vkjit_var_type(vkjit_ir* ir, vkjit_var id, vkjit_type* out_ty);
/* `ir.vars[id].ref_count` (white-box field the reference's own test reads, test.rs:204-206) */
vkjit_status vkjit_var_ref_count(vkjit_ir* ir, vkjit_var id, uint32_t* out);
/* `ir.vars.len()` and `ir.arrays.len()` (test.rs:204, :206) */
vkjit_status vkjit_var_count(vkjit_ir* ir, size_t* out);
/* `ir.vars[id].deps` / `.side_effects` (internal.rs:108-109; what DepIterator / SeIterator walk, iterators.rs:6-61): up to
 * `cap` dependency ids into deps[], the total count into *out_ndeps; *out_side_effect = the scatter target when
 * *out_has_side_effect != 0 (a var has at most one, internal.rs:396). */
vkjit_status vkjit_var_deps(vkjit_ir* ir, vkjit_var id, vkjit_var* deps, size_t cap, size_t* out_ndeps,
                            int32_t* out_has_side_effect, vkjit_var* out_side_effect);
vkjit_status vkjit_array_count(vkjit_ir* ir, size_t* out);
/* Ir::is_buffer, internal.rs:401-403 */
vkjit_status vkjit_is_buffer(vkjit_ir* ir, vkjit_var id, int32_t* out);
/* Array::size / stride and Array::device_address (backend/mod.rs:8-12) of a buffer var */
vkjit_status vkjit_var_size(vkjit_ir* ir, vkjit_var id, size_t* out_elems);
vkjit_status vkjit_var_device_ptr(vkjit_ir* ir, vkjit_var id, uint64_t* out_ptr);
/* Ir::inc_ref_count / dec_ref_count, internal.rs:450-469 */
vkjit_status vkjit_inc_ref(vkjit_ir* ir, vkjit_var id);
vkjit_status vkjit_dec_ref(vkjit_ir* ir, vkjit_var id);
/* `format!("{:#?}", ir)` (vkjit-rust functions.rs:54-56, vkjit-python functions.rs:36-39).
 * Writes a NUL-terminated dump into buf (truncating) and the full length into out_len. */
vkjit_status vkjit_ir_repr(vkjit_ir* ir, char* buf, size_t cap, size_t* out_len);
/* Debug of one var (vkjit-rust types.rs:115-126) */
vkjit_status vkjit_var_repr(vkjit_ir* ir, vkjit_var id, char* buf, size_t cap, size_t* out_len);

/* ------------------------------------------------------------------ */
/* Execute — Ir::schedule / Ir::eval / Ir::as_slice.                    */
/* ------------------------------------------------------------------ */
/* Ir::schedule, internal.rs:476-481 */
vkjit_status vkjit_schedule(vkjit_ir* ir, const vkjit_var* ids, size_t n);
/* Ir::eval, internal.rs:482-525: schedule, compile (kernel cache keyed by trace
 * hash), launch, rewrite every root into a Binding owning its output, clear the
 * schedule.  Asynchronous on the backend stream. */
vkjit_status vkjit_eval(vkjit_ir* ir, const vkjit_var* ids, size_t n);
/* Ir::as_slice<T>, internal.rs:443-449 (+ Array::map, vulkan/mod.rs:29-34).
 * `ty` must equal the var's type (the reference asserts TypeId equality);
 * copies min(bytes, size) bytes device->host and synchronises. */
vkjit_status vkjit_read(vkjit_ir* ir, vkjit_var id, vkjit_type ty, void* dst, size_t bytes);

/* ------------------------------------------------------------------ */
/* Runtime primitives (extensions; hand-written CUDA).   */
/* ------------------------------------------------------------------ */
/* reduce_{sum,min,max}: returns a 1-element Binding of the same type.  A buffer operand runs the
 * hand-written reduction; an UNEVALUATED operand is reduced by one generated kernel that fuses the
 * trace with the reduction (the operand is not materialised and stays unevaluated).  Returns a 1-element Binding
 * of the same type.  With vkjit_dist_init active and a sharded operand the
 * per-GPU partial is combined across ranks (result replicated).  Reductions issued back to back overlap on the
 * device (programmatic dependent launch: the fold / exchange of one runs under the streaming phase of the next)
 * whenever the backend can prove the operand is complete; the observable order is that of the stream. */
vkjit_status vkjit_reduce(vkjit_ir* ir, int32_t red, vkjit_var id, vkjit_var* out);
/* prefix sum of a U32/I32 var (mod 2^32); exclusive != 0 -> exclusive scan.  A sharded operand (multi-GPU) is
 * scanned over the GLOBAL range: per-rank totals are exchanged and each rank's result is its slice.
 * prefix_sum / compress / compress_values: an UNEVALUATED operand is evaluated inside the primitive's kernel (one
 * generated kernel: trace body + look-back scan); like reduce, the primitives never turn an operand into a buffer. */
vkjit_status vkjit_prefix_sum(vkjit_ir* ir, vkjit_var id, int32_t exclusive, vkjit_var* out);
/* compress(mask): stable list of lane indices whose mask is set (U32[count]).
 * Sharded mask (multi-GPU): every rank compacts its shard; the result is a ragged sharded array (rank r holds the
 * global elements [vkjit_var_shard_base(out), + its size)), indices are GLOBAL lane numbers and *out_count is the
 * GLOBAL count (replicated).  All ranks must call it, also those whose shard is empty. */
vkjit_status vkjit_compress(vkjit_ir* ir, vkjit_var mask, vkjit_var* out_indices, size_t* out_count);
/* compress(values, mask): values of the selected lanes, stable order. */
vkjit_status vkjit_compress_values(vkjit_ir* ir, vkjit_var values, vkjit_var mask, vkjit_var* out_values, size_t* out_count);

/* ------------------------------------------------------------------ */
/* Multi-GPU (new; the reference is single-device, device.rs:162-200).  */
/* One process per GPU.  Contiguous 1-D shards; only elementwise traces */
/* and reductions shard.                                */
/* ------------------------------------------------------------------ */
/* 128-byte NCCL unique id, created on rank 0 and shipped by the host. */
vkjit_status vkjit_dist_unique_id(void* out_id128);
vkjit_status vkjit_dist_init(int32_t rank, int32_t world, const void* id128);
/* Fused reduce + all-reduce over NVLink peer memory (optional; NCCL is used until it is set up).
 * vkjit_dist_mailbox_handle: allocate this rank's mailbox, return its 64-byte cudaIpc handle.
 * vkjit_dist_mailbox_open: map all ranks' mailboxes (world x 64 bytes, rank order).  Afterwards the
 * last CTA of every sharded reduction publishes the per-GPU partial into every peer's mailbox and
 * combines in rank order (deterministic); $VKJIT_DIST=nccl keeps the NCCL path. */
vkjit_status vkjit_dist_mailbox_handle(void* out_handle64);
vkjit_status vkjit_dist_mailbox_open(const void* handles, int32_t world);
/* Switch between the fused mailbox path (1) and NCCL (0); every rank must make the same call. */
/* Multi-GPU WITHOUT torch / MPI (one process per GPU, any launcher): reads RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR,
 * MASTER_PORT from the environment, calls vkjit_init(LOCAL_RANK) if that has not happened, exchanges rank 0's NCCL id
 * and every rank's mailbox handle over TCP (rank 0 listens on MASTER_ADDR : $VKJIT_RDZV_PORT, default MASTER_PORT + 1;
 * $VKJIT_RDZV_TIMEOUT_S, default 120), then does what vkjit_dist_init + vkjit_dist_mailbox_open do.  A vkjit-rust or C
 * caller shards with this one call; the front-end's single global `Ir` (vkjit-rust/src/lib.rs:9-11) stays as it is. */
vkjit_status vkjit_dist_init_env(void);
/* Debug / test: the rendezvous alone (no device, no NCCL): blob64 of every rank -> out_all (world x 64 bytes, rank
 * order); root128 is rank 0's on entry and everyone's on return. */
vkjit_status vkjit_debug_rendezvous(int32_t rank, int32_t world, const char* addr, int32_t port, const void* blob64,
                                    void* root128, void* out_all, double timeout_s);
vkjit_status vkjit_dist_set_p2p(int32_t on);
vkjit_status vkjit_dist_shutdown(void);
vkjit_status vkjit_dist_info(int32_t* out_rank, int32_t* out_world);
/* Shard [lo, hi) of a global 1-D range of n lanes owned by `rank` of `world`
 * (multiples of 4 lanes so every shard stays 16-byte aligned). Pure function. */
vkjit_status vkjit_shard_range(size_t n, int32_t rank, int32_t world, size_t* out_lo, size_t* out_hi);
/* arange over the GLOBAL range [0, n): this rank's var has hi-lo lanes and
 * lane i holds lo + i.  Vars derived from it are "sharded". */
vkjit_status vkjit_arange_sharded(vkjit_ir* ir, vkjit_type ty, size_t n, vkjit_var* out);
/* Upload this rank's slice of a replicated HOST array of n elements. */
vkjit_status vkjit_array_sharded(vkjit_ir* ir, vkjit_type ty, const void* data, size_t n, vkjit_var* out);
/* Upload this rank's slice directly: `data` holds the n_local elements this rank owns. */
vkjit_status vkjit_array_shard_local(vkjit_ir* ir, vkjit_type ty, const void* data, size_t n_local, vkjit_var* out);
vkjit_status vkjit_var_is_sharded(vkjit_ir* ir, vkjit_var id, int32_t* out);
/* Global index of the first element this rank holds of a sharded var: set for vkjit_arange_sharded,
 * vkjit_array_sharded and the results of a sharded compress (0 otherwise). */
vkjit_status vkjit_var_shard_base(vkjit_ir* ir, vkjit_var id, uint64_t* out);

/* ------------------------------------------------------------------ */
/* Counters (new).                                                      */
/* ------------------------------------------------------------------ */
typedef struct vkjit_stats_t {
  uint64_t cache_hits;      /* evals served from the kernel cache            */
  uint64_t cache_misses;    /* evals that ran NVRTC                          */
  uint64_t trace_launches;  /* NVRTC-generated kernels launched              */
  uint64_t prim_launches;   /* hand-written primitive kernels launched       */
  uint64_t last_compile_ns; /* codegen + NVRTC + module load of the last miss */
  uint64_t last_eval_ns;    /* host time spent inside the last vkjit_eval    */
  uint64_t bytes_h2d;
  uint64_t bytes_d2h;
  uint64_t pool_bytes_live; /* bytes currently handed out by the pool        */
  uint64_t collectives;     /* cross-GPU combines issued                     */
  uint64_t disk_hits;       /* kernel-cache misses served from $VKJIT_CACHE_DIR (no NVRTC) */
} vkjit_stats_t;
vkjit_status vkjit_stats(vkjit_stats_t* out);
vkjit_status vkjit_stats_reset(void);
/* Drop every cached kernel (forces recompilation; used by the compile-latency bench). */
vkjit_status vkjit_cache_clear(void);

/* Debug: generated CUDA C for the given roots without launching (needs no
 * device).  compile bit 0: additionally run NVRTC for sm_100a and report its verdict (out_cubin_bytes
 * receives the cubin size, 0 if not compiled); compile bit 1: generate the shared-memory-privatised
 * scatter_add variant that large launches use (bit 2: the 2-CTA cluster variant, bit 3: the bin-range pass variant). */
vkjit_status vkjit_debug_codegen(vkjit_ir* ir, const vkjit_var* ids, size_t n, int32_t compile,
                                 char* buf, size_t cap, size_t* out_len, size_t* out_cubin_bytes);
/* Debug: average host time of the per-eval trace walk + hash (the cache-hit critical path) and the
 * number of trace nodes; needs no device.  reps = 0: time ONE walk, in whatever cache state the construction
 * of the trace left its vars (what a fresh eval pays). */
vkjit_status vkjit_debug_walk_ns(vkjit_ir* ir, const vkjit_var* ids, size_t n, uint32_t reps, uint64_t* out_ns, uint32_t* out_nodes);
/* Debug: the bookkeeping of Ir::eval (internal.rs:482-525) WITHOUT compiling or launching anything — schedule, walk,
 * every root rewritten into a Binding (a foreign view of the right size with NO memory behind it), the consumed trace
 * released, schedule cleared.  Lets the ref-count contract of eval be tested without a device; never evaluate or read
 * anything that depends on such a root. */
vkjit_status vkjit_debug_eval_bookkeeping(vkjit_ir* ir, const vkjit_var* ids, size_t n);
/* Same for the fused trace -> reduce kernel of vkjit_reduce(red, id). */
vkjit_status vkjit_debug_codegen_reduce(vkjit_ir* ir, vkjit_var id, int32_t red, int32_t compile,
                                        char* buf, size_t cap, size_t* out_len, size_t* out_cubin_bytes);
/* Same for the fused trace -> scan kernel that vkjit_prefix_sum / vkjit_compress / vkjit_compress_values launch when
 * an operand is unevaluated.  mode: 0 exclusive sum, 1 inclusive sum, 2 compress -> indices (ids = {mask}),
 * 3 compress -> values (ids = {mask, values}). */
vkjit_status vkjit_debug_codegen_scan(vkjit_ir* ir, const vkjit_var* ids, size_t n, int32_t mode, int32_t compile,
                                      char* buf, size_t cap, size_t* out_len, size_t* out_cubin_bytes);
/* Debug ($VKJIT_REDUCE_TRACE=1 in the environment before the first reduction): per-launch %globaltimer stamps of the
 * hand-written reduce kernel, 8 words per launch, oldest launch first: [0] CTA 0 enters; in the CTA that folds:
 * [1] its streaming phase done, [2] previous kernel on the stream complete, [3] last ticket taken, [4] partials folded,
 * [5] peer exchange done / result written; [6] launch number; [7] world size.  Synchronises; clears the ring. */
vkjit_status vkjit_debug_reduce_trace(uint64_t* out, size_t cap_words, size_t* out_launches);

#ifdef __cplusplus
}
#endif
#endif /* VKJIT_B200_H */
