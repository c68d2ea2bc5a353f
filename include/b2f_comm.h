/* libb2f_comm.so -- the training path's only collective behind a C ABI.
 *
 * Replaces what nn.DataParallelTable(1, true, true) does with nccl.torch once per step in the reference
 * (util.lua:27-48: flattenParams = true, usenccl = true; train.lua:480 `model:backward`, :494-496
 * `model:syncParameters()`): ONE sum-all-reduce of the flattened fp32 gradient.  Here every rank (one process per
 * GPU) owns a full replica, so the reduce-to-GPU-1 + parameter broadcast pair of the reference collapses into an
 * all-reduce, after which every rank takes the same Adam step.
 *
 * A separate library so that libb2f_cuda.so keeps no NCCL dependency; NCCL itself is dlopen'ed ("libnccl.so.2") at
 * b2f_comm_unique_id / b2f_comm_init, so a host without NCCL can still load the symbols.  Same conventions as b2f.h:
 * plain pointers and sizes, int status (0 = ok, negative = B2F_E*, positive = ncclResult_t + 1000), explicit stream,
 * thread-local last error.  Rendezvous is the host's business: rank 0 calls b2f_comm_unique_id and ships the 128 bytes
 * to the other ranks by whatever it has (the Lua host: its `threads`/socket channel; the Python mirror:
 * torch.distributed's store).                                                                                    */
#ifndef B2F_COMM_H
#define B2F_COMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B2F_COMM_API
#else
#define B2F_COMM_API __attribute__((visibility("default")))
#endif

#define B2F_COMM_ID_BYTES 128
typedef struct b2f_comm* b2f_comm_t;
typedef void* b2f_comm_stream_t; /* cudaStream_t */

B2F_COMM_API int b2f_comm_abi_version(void);
B2F_COMM_API const char* b2f_comm_last_error(void);
/* ncclGetUniqueId: fills B2F_COMM_ID_BYTES bytes (rank 0 only).                                              */
B2F_COMM_API int b2f_comm_unique_id(void* id_out);
/* ncclCommInitRank on the calling thread's current device.  Collective: every rank calls it with the same id. */
B2F_COMM_API int b2f_comm_init(b2f_comm_t* comm, const void* id, int world, int rank);
B2F_COMM_API int b2f_comm_world(b2f_comm_t comm, int* world, int* rank);
/* In-place sum over all ranks of buf[0 .. count), fp32, enqueued on `stream` (asynchronous: returns when the
 * collective is enqueued).  With world == 1 it is a no-op.  Buckets = several calls on disjoint ranges.        */
B2F_COMM_API int b2f_comm_allreduce_sum_f32(b2f_comm_t comm, float* buf, size_t count, b2f_comm_stream_t stream);
/* Same for float64 (the logged loss scalars, train.lua:497-513).                                              */
B2F_COMM_API int b2f_comm_allreduce_sum_f64(b2f_comm_t comm, double* buf, size_t count, b2f_comm_stream_t stream);
B2F_COMM_API int b2f_comm_destroy(b2f_comm_t comm);

#ifdef __cplusplus
}
#endif
#endif
