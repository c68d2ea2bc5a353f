/* Stand-in for cutorch's THCGeneral.h / THCTensor.h: the fields and accessors the reference's
 * sampler glue reads (BilinearSamplerBHWD.cu:121-150).  See lua.h in this directory. */
#ifndef B2F_REF_SHIM_THCGENERAL_H
#define B2F_REF_SHIM_THCGENERAL_H
#include <cuda_runtime.h>
#include <stdexcept>

typedef struct THCState { cudaStream_t stream; } THCState;
typedef struct THCudaTensor {
    float* data;
    long   size[4];
    long   stride[4];
} THCudaTensor;

static inline cudaStream_t THCState_getCurrentStream(THCState* s) { return s->stream; }
static inline float* THCudaTensor_data(THCState*, THCudaTensor* t) { return t->data; }
static inline long   THCudaTensor_size(THCState*, THCudaTensor* t, int d) { return t->size[d]; }
static inline long   THCudaTensor_stride(THCState*, THCudaTensor* t, int d) { return t->stride[d]; }
static inline void   THError(const char* msg) { throw std::runtime_error(msg); }
#endif
