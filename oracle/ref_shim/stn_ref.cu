/* oracle/_ref/libstn_ref.so -- the REFERENCE's own sampler (Lua-C glue + CUDA kernels), compiled
 * unmodified from /root/reference/extras/stnbhwd/{utils.c,BilinearSamplerBHWD.cu} in the include order of
 * extras/stnbhwd/init.cu:1-6, against the stand-in Torch7 headers of this directory, for sm_100a (the
 * reference's own flags are -arch=sm_30, which nvcc 12.9 no longer accepts).  TEST INFRASTRUCTURE ONLY:
 * loaded by tests/ (GPU parity pin for rows a4/a5 of SURVEY.md section 8) and never by the product.
 * Nothing of the reference is copied into this repository; the sources are #included by path. */
#include "luaT.h"
#include "THC.h"
#include "utils.c"
#include "BilinearSamplerBHWD.cu"

static void fill(THCudaTensor* t, float* p, long d0, long d1, long d2, long d3) {
    t->data = p;
    t->size[0] = d0; t->size[1] = d1; t->size[2] = d2; t->size[3] = d3;
    t->stride[3] = 1; t->stride[2] = d3; t->stride[1] = d2 * d3; t->stride[0] = d1 * d2 * d3;
}

#define REF_EXPORT extern "C" __attribute__((visibility("default")))

/* img (B,H,W,C), grid (B,Hg,Wg,2), out (B,Hg,Wg,C): contiguous device buffers. */
REF_EXPORT int stn_ref_update_output(float* img, float* grid, float* out,
                                     int B, int H, int W, int C, int Hg, int Wg, void* stream) {
    THCState st; st.stream = (cudaStream_t)stream;
    THCudaTensor ti, tg, to;
    fill(&ti, img, B, H, W, C); fill(&tg, grid, B, Hg, Wg, 2); fill(&to, out, B, Hg, Wg, C);
    lua_State L = {}; L.thc = &st; L.slot[2] = &ti; L.slot[3] = &tg; L.slot[4] = &to;
    try { cunn_BilinearSamplerBHWD_updateOutput(&L); } catch (const std::exception&) { return -1; }
    return 0;
}

/* gradImg / gradGrid must be zero-filled by the caller, as BilinearSamplerBHWD.lua:99-102 does. */
REF_EXPORT int stn_ref_update_grad_input(float* img, float* grid, float* gradImg, float* gradGrid,
                                         float* gradOut, int B, int H, int W, int C, int Hg, int Wg,
                                         int only_grid, void* stream) {
    THCState st; st.stream = (cudaStream_t)stream;
    THCudaTensor ti, tg, tgi, tgg, tgo;
    fill(&ti, img, B, H, W, C); fill(&tg, grid, B, Hg, Wg, 2); fill(&tgi, gradImg, B, H, W, C);
    fill(&tgg, gradGrid, B, Hg, Wg, 2); fill(&tgo, gradOut, B, Hg, Wg, C);
    lua_State L = {}; L.thc = &st; L.slot[2] = &ti; L.slot[3] = &tg; L.slot[4] = &tgi; L.slot[5] = &tgg;
    L.slot[6] = &tgo;
    try {
        if (only_grid) cunn_BilinearSamplerBHWD_updateGradInputOnlyGrid(&L);
        else cunn_BilinearSamplerBHWD_updateGradInput(&L);
    } catch (const std::exception&) { return -1; }
    return 0;
}
