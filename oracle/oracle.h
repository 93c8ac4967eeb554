/*
 * oracle.h — C API of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * The oracle is a CPU restatement of the reference's algorithm for the hot path
 * (trace IR -> per-lane evaluation -> readback); see oracle.cpp for the
 * file:line map.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product
 * (vkjit_b200/) never links, imports or executes anything in this directory.
 *
 * The function set mirrors include/vkjit_b200.h one to one (prefix orc_
 * instead of vkjit_, same argument meaning, same status codes) minus
 * everything that only makes sense with a device (init, stream, pinned
 * memory, NCCL, kernel cache), so the parity tests can drive both sides
 * through the same host code.
 */
#ifndef VKJIT_ORACLE_H
#define VKJIT_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ir orc_ir;

const char* orc_last_error(void);
/* worker threads used by eval and the eager primitives (default 1) */
int32_t orc_set_threads(int32_t n);

int32_t orc_ir_create(orc_ir** out);
int32_t orc_ir_destroy(orc_ir* ir);

int32_t orc_type_struct(orc_ir* ir, const uint32_t* elems, size_t n, uint32_t* out_ty);
int32_t orc_type_struct_len(orc_ir* ir, uint32_t ty, size_t* out_n);
int32_t orc_type_struct_elem(orc_ir* ir, uint32_t ty, size_t i, uint32_t* out_elem);

int32_t orc_const_f32(orc_ir* ir, float v, uint32_t* out);
int32_t orc_const_i32(orc_ir* ir, int32_t v, uint32_t* out);
int32_t orc_const_u32(orc_ir* ir, uint32_t v, uint32_t* out);
int32_t orc_const_bool(orc_ir* ir, int32_t v, uint32_t* out);
int32_t orc_array_f32(orc_ir* ir, const float* data, size_t n, uint32_t* out);
int32_t orc_array_i32(orc_ir* ir, const int32_t* data, size_t n, uint32_t* out);
int32_t orc_array_u32(orc_ir* ir, const uint32_t* data, size_t n, uint32_t* out);
int32_t orc_array_bool(orc_ir* ir, const uint32_t* data, size_t n, uint32_t* out);
int32_t orc_array_empty(orc_ir* ir, uint32_t ty, size_t n, uint32_t* out);
int32_t orc_arange(orc_ir* ir, uint32_t ty, size_t n, uint32_t* out);
int32_t orc_linspace(orc_ir* ir, uint32_t ty, uint32_t start, uint32_t stop, size_t n, uint32_t* out);
int32_t orc_zeros(orc_ir* ir, uint32_t ty, uint32_t* out);
int32_t orc_ones(orc_ir* ir, uint32_t ty, uint32_t* out);
int32_t orc_cast(orc_ir* ir, uint32_t src, uint32_t ty, uint32_t* out);
int32_t orc_bop(orc_ir* ir, int32_t kind, uint32_t lhs, uint32_t rhs, uint32_t* out);
int32_t orc_uop(orc_ir* ir, int32_t kind, uint32_t src, uint32_t* out);
int32_t orc_bitcast(orc_ir* ir, uint32_t src, uint32_t ty, uint32_t* out);
int32_t orc_select(orc_ir* ir, uint32_t cond, uint32_t lhs, uint32_t rhs, uint32_t* out);
int32_t orc_struct_init(orc_ir* ir, const uint32_t* elems, size_t n, uint32_t* out);
int32_t orc_getattr(orc_ir* ir, uint32_t src, size_t idx, uint32_t* out);
int32_t orc_setattr(orc_ir* ir, uint32_t dst, uint32_t src, size_t idx, uint32_t* out);
int32_t orc_gather(orc_ir* ir, uint32_t src, uint32_t idx, int32_t has_active, uint32_t active, uint32_t* out);
int32_t orc_scatter(orc_ir* ir, uint32_t src, uint32_t dst, uint32_t idx, int32_t has_active, uint32_t active, uint32_t* out);
int32_t orc_scatter_add(orc_ir* ir, uint32_t src, uint32_t dst, uint32_t idx, int32_t has_active, uint32_t active, uint32_t* out);

int32_t orc_var_type(orc_ir* ir, uint32_t id, uint32_t* out_ty);
int32_t orc_var_ref_count(orc_ir* ir, uint32_t id, uint32_t* out);
int32_t orc_var_count(orc_ir* ir, size_t* out);
int32_t orc_array_count(orc_ir* ir, size_t* out);
int32_t orc_is_buffer(orc_ir* ir, uint32_t id, int32_t* out);
int32_t orc_var_size(orc_ir* ir, uint32_t id, size_t* out_elems);
int32_t orc_inc_ref(orc_ir* ir, uint32_t id);
int32_t orc_dec_ref(orc_ir* ir, uint32_t id);
int32_t orc_ir_repr(orc_ir* ir, char* buf, size_t cap, size_t* out_len);
int32_t orc_var_repr(orc_ir* ir, uint32_t id, char* buf, size_t cap, size_t* out_len);

int32_t orc_schedule(orc_ir* ir, const uint32_t* ids, size_t n);
int32_t orc_eval(orc_ir* ir, const uint32_t* ids, size_t n);
int32_t orc_read(orc_ir* ir, uint32_t id, uint32_t ty, void* dst, size_t bytes);
/* host pointer to the words of a buffer var (valid until the var dies) */
int32_t orc_var_host_ptr(orc_ir* ir, uint32_t id, void** out_ptr);

int32_t orc_reduce(orc_ir* ir, int32_t red, uint32_t id, uint32_t* out);
int32_t orc_prefix_sum(orc_ir* ir, uint32_t id, int32_t exclusive, uint32_t* out);
int32_t orc_compress(orc_ir* ir, uint32_t mask, uint32_t* out_indices, size_t* out_count);
int32_t orc_compress_values(orc_ir* ir, uint32_t values, uint32_t mask, uint32_t* out_values, size_t* out_count);

/* Sharded constructors restated on one host: rank/world are explicit. */
int32_t orc_shard_range(size_t n, int32_t rank, int32_t world, size_t* out_lo, size_t* out_hi);
int32_t orc_arange_shard(orc_ir* ir, uint32_t ty, size_t n, int32_t rank, int32_t world, uint32_t* out);

/* The stateless input generator of SURVEY.md §8d, so host-side tests and the
 * CPU baseline fill arrays with exactly the words the device generates.
 * kind: 0 = raw u32 hash, 1 = f32 uniform [0,1), 2 = f32 uniform [-1,1),
 *       3 = u32 hash & 0xFFFF, 4 = bool (hash & 1) as 0/1 word */
int32_t orc_fill_hash(void* dst, size_t n, uint64_t first_lane, uint32_t seed, int32_t kind);

#ifdef __cplusplus
}
#endif
#endif
