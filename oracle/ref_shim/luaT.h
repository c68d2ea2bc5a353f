/* Stand-in for luaT (torch7/lib/luaT): typed userdata access.  See lua.h in this directory. */
#ifndef B2F_REF_SHIM_LUAT_H
#define B2F_REF_SHIM_LUAT_H
#include "lua.h"
#define LUA_EXTERNC extern "C"
#define DLL_EXPORT
static inline void* luaT_checkudata(lua_State* L, int idx, const char*) { return L->slot[idx]; }
static inline void  luaT_pushmetatable(lua_State*, const char*) {}
static inline void  luaT_registeratname(lua_State*, const luaL_Reg*, const char*) {}
#endif
