/* Stand-in for the tiny part of the Lua C API that the reference's sampler glue touches
 * (extras/stnbhwd/utils.c:3-11, BilinearSamplerBHWD.cu:118-158, 313-435).  TEST INFRASTRUCTURE ONLY:
 * it exists so that the reference's own, unmodified CUDA source can be compiled from where it lies
 * under /root/reference into oracle/_ref/libstn_ref.so and be used as the parity pin of the sampler.
 * A "Lua state" here is just the argument slots of one call plus the THC state. */
#ifndef B2F_REF_SHIM_LUA_H
#define B2F_REF_SHIM_LUA_H
#include <stddef.h>
#include <stdio.h>

struct THCState;
typedef struct lua_State {
    void* slot[8];            /* slot[i] = userdata at stack index i (1-based, as luaT_checkudata sees it) */
    struct THCState* thc;     /* what cutorch.getState() returns */
} lua_State;

typedef int (*lua_CFunction)(lua_State*);
typedef struct luaL_Reg { const char* name; lua_CFunction func; } luaL_Reg;

/* utils.c: lua_getglobal(L,"cutorch"); lua_getfield(L,-1,"getState"); lua_call(L,0,1);
 *          state = lua_touserdata(L,-1); lua_pop(L,2); */
static inline void  lua_getglobal(lua_State*, const char*) {}
static inline void  lua_getfield(lua_State*, int, const char*) {}
static inline void  lua_call(lua_State*, int, int) {}
static inline void* lua_touserdata(lua_State* L, int) { return (void*)L->thc; }
static inline void  lua_pop(lua_State*, int) {}
static inline void  lua_newtable(lua_State*) {}
#endif
