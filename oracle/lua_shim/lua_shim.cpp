// TEST INFRASTRUCTURE — not part of the product.
//
// Lua 5.2 C-API subset (see lua.h here) implemented over the repo's minilua
// evaluator, so that the reference's own engine/NQ/fisheye.c can be compiled
// unmodified into oracle/_ref/.  Behaviour follows the Lua 5.2 reference manual
// §4.8 for each entry point; stack indices are relative to the running
// C function's frame exactly as in Lua.
#include "lua.h"

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "minilua.h"

using minilua::LuaError;
using minilua::Table;
using minilua::Type;
using minilua::Value;
using minilua::ValueList;

struct lua_State {
    minilua::State S;
    std::vector<Value> stack;
    size_t base = 0;  // first slot of the current C frame
    std::vector<Value> registry;
    std::vector<std::string *> cstr_keepalive;
    struct CFn {
        lua_State *L;
        lua_CFunction f;
    };
    std::vector<CFn *> cfns;
    ~lua_State() {
        for (auto *c : cfns) delete c;
        for (auto *s : cstr_keepalive) delete s;
    }
};

namespace {

Value &at(lua_State *L, int idx) {
    static Value nil_sentinel;
    if (idx > 0) {
        size_t p = L->base + static_cast<size_t>(idx) - 1;
        if (p >= L->stack.size()) {
            nil_sentinel = Value();
            return nil_sentinel;
        }
        return L->stack[p];
    }
    if (idx < 0 && idx > LUA_REGISTRYINDEX) {
        size_t p = L->stack.size() - static_cast<size_t>(-idx);
        return L->stack[p];
    }
    fprintf(stderr, "lua_shim: unsupported stack index %d\n", idx);
    abort();
}

void push(lua_State *L, const Value &v) { L->stack.push_back(v); }

void trampoline(minilua::State &, const Value *args, int nargs, ValueList &out, void *ud) {
    lua_State::CFn *c = static_cast<lua_State::CFn *>(ud);
    lua_State *L = c->L;
    size_t saved_base = L->base;
    size_t frame = L->stack.size();
    L->base = frame;
    for (int i = 0; i < nargs; ++i) L->stack.push_back(args[i]);
    int nret;
    try {
        nret = c->f(L);
    } catch (...) {
        L->stack.resize(frame);
        L->base = saved_base;
        throw;
    }
    size_t top = L->stack.size();
    for (size_t i = top - static_cast<size_t>(nret); i < top; ++i) out.push_back(L->stack[i]);
    L->stack.resize(frame);
    L->base = saved_base;
}

int do_call(lua_State *L, int nargs, int nresults, bool protect) {
    size_t fpos = L->stack.size() - static_cast<size_t>(nargs) - 1;
    Value fn = L->stack[fpos];
    std::vector<Value> args(L->stack.begin() + static_cast<long>(fpos) + 1, L->stack.end());
    L->stack.resize(fpos);
    ValueList out;
    try {
        L->S.call(fn, args.data(), nargs, out);
    } catch (LuaError &e) {
        if (!protect) {
            // Lua: "PANIC: unprotected error in call to Lua API"
            fprintf(stderr, "PANIC: unprotected error in call to Lua API (%s)\n", e.what());
            abort();
        }
        L->stack.resize(fpos);
        push(L, L->S.new_string(e.what()));
        return LUA_ERRRUN;
    }
    int n = out.size();
    if (nresults == LUA_MULTRET) {
        for (int i = 0; i < n; ++i) push(L, out[i]);
    } else {
        for (int i = 0; i < nresults; ++i) push(L, i < n ? out[i] : Value());
    }
    return LUA_OK;
}

}  // namespace

extern "C" {

lua_State *luaL_newstate(void) { return new lua_State(); }
void luaL_openlibs(lua_State *) {}  // minilua opens its libraries on construction
void lua_close(lua_State *L) { delete L; }

int luaL_loadbuffer(lua_State *L, const char *buff, size_t sz, const char *name) {
    try {
        push(L, L->S.load(std::string(buff, sz), name ? name : "?"));
        return LUA_OK;
    } catch (LuaError &e) {
        push(L, L->S.new_string(e.what()));
        return LUA_ERRSYNTAX;
    }
}

int luaL_loadfile(lua_State *L, const char *filename) {
    FILE *f = fopen(filename, "rb");
    if (!f) {
        push(L, L->S.new_string(std::string("cannot open ") + filename));
        return LUA_ERRFILE;
    }
    fclose(f);
    try {
        push(L, L->S.load_file(filename));
        return LUA_OK;
    } catch (LuaError &e) {
        push(L, L->S.new_string(e.what()));
        return LUA_ERRSYNTAX;
    }
}

int lua_pcall(lua_State *L, int nargs, int nresults, int) { return do_call(L, nargs, nresults, true); }
void lua_call(lua_State *L, int nargs, int nresults) { do_call(L, nargs, nresults, false); }

int lua_gettop(lua_State *L) { return static_cast<int>(L->stack.size() - L->base); }
void lua_pop(lua_State *L, int n) { L->stack.resize(L->stack.size() - static_cast<size_t>(n)); }

void lua_pushnil(lua_State *L) { push(L, Value()); }
void lua_pushnumber(lua_State *L, lua_Number n) { push(L, Value(n)); }
void lua_pushinteger(lua_State *L, lua_Integer n) { push(L, Value(static_cast<double>(n))); }
void lua_pushcfunction(lua_State *L, lua_CFunction f) {
    auto *c = new lua_State::CFn{L, f};
    L->cfns.push_back(c);
    push(L, L->S.new_cfunction(trampoline, c, "cfunction"));
}

void lua_getglobal(lua_State *L, const char *name) { push(L, L->S.get_global(name)); }
void lua_setglobal(lua_State *L, const char *name) {
    L->S.set_global(name, L->stack.back());
    L->stack.pop_back();
}

void lua_rawgeti(lua_State *L, int idx, int n) {
    if (idx == LUA_REGISTRYINDEX) {
        push(L, (n >= 0 && static_cast<size_t>(n) < L->registry.size()) ? L->registry[static_cast<size_t>(n)] : Value());
        return;
    }
    Value t = at(L, idx);
    push(L, t.is_table() ? static_cast<Table *>(t.obj())->get_int(n) : Value());
}

size_t lua_rawlen(lua_State *L, int idx) {
    const Value &v = at(L, idx);
    if (v.is_table()) return static_cast<size_t>(static_cast<Table *>(v.obj())->length());
    if (v.is_string()) return v.str().size();
    return 0;
}

int lua_next(lua_State *L, int idx) {
    Value tv = at(L, idx);  // resolve idx while the key is still on the stack, as Lua does
    Value key = L->stack.back();
    L->stack.pop_back();
    Table *t = static_cast<Table *>(tv.obj());
    size_t pos = 0;
    if (!key.is_nil()) {
        bool found = false;
        if (key.is_number()) {
            double d = key.num();
            if (d >= 1 && d <= static_cast<double>(t->arr.size()) && d == static_cast<double>(static_cast<size_t>(d))) {
                pos = static_cast<size_t>(d);
                found = true;
            }
        }
        if (!found) {
            for (size_t j = 0; j < t->hash_order.size(); ++j)
                if (t->hash_order[j].raw_equals(key)) {
                    pos = t->arr.size() + j + 1;
                    found = true;
                    break;
                }
        }
        if (!found) {
            fprintf(stderr, "lua_shim: invalid key to lua_next\n");
            abort();
        }
    }
    Value k, v;
    if (!t->next(&pos, &k, &v)) return 0;
    push(L, k);
    push(L, v);
    return 1;
}

int luaL_ref(lua_State *L, int t) {
    if (t != LUA_REGISTRYINDEX) {
        fprintf(stderr, "lua_shim: luaL_ref only supports the registry\n");
        abort();
    }
    Value v = L->stack.back();
    L->stack.pop_back();
    if (v.is_nil()) return -1;  // LUA_REFNIL
    if (L->registry.empty()) L->registry.push_back(Value());  // refs start at 1
    L->registry.push_back(v);
    return static_cast<int>(L->registry.size()) - 1;
}

int lua_isnil(lua_State *L, int idx) { return at(L, idx).is_nil(); }
int lua_isnumber(lua_State *L, int idx) {
    double d;
    return at(L, idx).to_number(&d);
}
int lua_isstring(lua_State *L, int idx) {
    const Value &v = at(L, idx);
    return v.is_string() || v.is_number();
}
int lua_isfunction(lua_State *L, int idx) { return at(L, idx).is_function(); }
int lua_istable(lua_State *L, int idx) { return at(L, idx).is_table(); }

lua_Number lua_tonumber(lua_State *L, int idx) {
    double d = 0;
    if (!at(L, idx).to_number(&d)) return 0;
    return d;
}

lua_Integer lua_tointeger(lua_State *L, int idx) {
    double d = 0;
    if (!at(L, idx).to_number(&d)) return 0;
    return static_cast<lua_Integer>(d);  // lua_number2integer: C cast
}

const char *lua_tostring(lua_State *L, int idx) {
    Value &v = at(L, idx);
    if (v.is_number()) v = L->S.new_string(minilua::State::tostring(v));  // converts in place, as Lua does
    if (!v.is_string()) return nullptr;
    // the reference uses the pointer after popping the value (fisheye.c:1090-1102
    // pops later, :1671-1672 before): keep a stable copy alive for the State's life.
    std::string *keep = new std::string(v.str());
    L->cstr_keepalive.push_back(keep);
    return keep->c_str();
}

lua_Number luaL_checknumber(lua_State *L, int arg) {
    double d;
    if (!at(L, arg).to_number(&d)) {
        throw LuaError(std::string("bad argument #") + std::to_string(arg) + " (number expected, got " +
                       minilua::State::type_name(at(L, arg)) + ")");
    }
    return d;
}

}  // extern "C"
