/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Minimal stand-in for the Lua 5.2 C API headers, declaring exactly the 28
 * entry points the reference's engine/NQ/fisheye.c calls (fisheye.c:1222-1264,
 * 1545-1651, 1659-1913).  They are implemented in lua_shim.cpp on top of the
 * repo's own Lua-subset evaluator (blinky_b200/csrc/minilua), because no Lua
 * exists in the build image.  Used only to compile the UNMODIFIED reference
 * fisheye.c into oracle/_ref/ as the parity oracle.
 */
#ifndef BLINKY_ORACLE_LUA_SHIM_H
#define BLINKY_ORACLE_LUA_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lua_State lua_State;
typedef double lua_Number;
typedef ptrdiff_t lua_Integer;
typedef int (*lua_CFunction)(lua_State *L);

#define LUA_MULTRET (-1)
#define LUA_REGISTRYINDEX (-1001000)

#define LUA_OK 0
#define LUA_YIELD 1
#define LUA_ERRRUN 2
#define LUA_ERRSYNTAX 3
#define LUA_ERRMEM 4
#define LUA_ERRGCMM 5
#define LUA_ERRERR 6
#define LUA_ERRFILE 7

lua_State *luaL_newstate(void);
void luaL_openlibs(lua_State *L);
void lua_close(lua_State *L);

int luaL_loadbuffer(lua_State *L, const char *buff, size_t sz, const char *name);
int luaL_loadfile(lua_State *L, const char *filename);
int lua_pcall(lua_State *L, int nargs, int nresults, int errfunc);
void lua_call(lua_State *L, int nargs, int nresults);

int lua_gettop(lua_State *L);
void lua_pop(lua_State *L, int n);

void lua_pushnil(lua_State *L);
void lua_pushnumber(lua_State *L, lua_Number n);
void lua_pushinteger(lua_State *L, lua_Integer n);
void lua_pushcfunction(lua_State *L, lua_CFunction f);

void lua_getglobal(lua_State *L, const char *name);
void lua_setglobal(lua_State *L, const char *name);
void lua_rawgeti(lua_State *L, int idx, int n);
size_t lua_rawlen(lua_State *L, int idx);
int lua_next(lua_State *L, int idx);
int luaL_ref(lua_State *L, int t);

int lua_isnil(lua_State *L, int idx);
int lua_isnumber(lua_State *L, int idx);
int lua_isstring(lua_State *L, int idx);
int lua_isfunction(lua_State *L, int idx);
int lua_istable(lua_State *L, int idx);

lua_Number lua_tonumber(lua_State *L, int idx);
lua_Integer lua_tointeger(lua_State *L, int idx);
const char *lua_tostring(lua_State *L, int idx);
lua_Number luaL_checknumber(lua_State *L, int arg);

#ifdef __cplusplus
}
#endif
#endif
