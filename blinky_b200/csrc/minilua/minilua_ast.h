// minilua — internal AST definitions, shared by the evaluator (minilua.cpp) and the
// Lua -> C/CUDA transpiler (../lua_transpile.cpp).  Not part of the public API.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "minilua.h"

namespace minilua {


struct VarInfo {
    int slot = 0;
    bool captured = false;
    std::string name;  // source name, for diagnostics and generated code
};

enum class EK : uint8_t {
    Nil, True, False, Number, String, Vararg,
    Local, Upval, Global,
    Index, Call, Method, Function,
    Add, Sub, Mul, Div, Mod, Pow, Concat,
    Eq, Ne, Lt, Le, Gt, Ge,
    And, Or, Not, Neg, Len,
    Table, Paren,
};

struct Expr {
    EK k = EK::Nil;
    int line = 0;
    double num = 0;          // Number
    int id = 0;              // String: kstr id; Global: global id; Upval: index; Method: kstr id
    VarInfo *var = nullptr;  // Local
    Expr *l = nullptr;       // lhs / object / callee / operand
    Expr *r = nullptr;       // rhs / key
    std::vector<Expr *> list;    // call args / table positional values
    std::vector<Expr *> keys;    // table: explicit keys (parallel to vals)
    std::vector<Expr *> vals;    // table: values for explicit keys
    Proto *proto = nullptr;  // Function
};

enum class SK : uint8_t {
    Local, Assign, Call, Do, While, Repeat, If, NumFor, GenFor, Return, Break, LocalFunction,
    Goto, Label,   // Lua 5.2 `goto name` / `::name::` (Stmt::label)
};

struct Block;

struct Stmt {
    SK k = SK::Do;
    int line = 0;
    std::vector<VarInfo *> vars;   // Local / GenFor / NumFor(1) / LocalFunction(1)
    std::vector<Expr *> targets;   // Assign
    std::vector<Expr *> exprs;     // Local / Assign / Return / GenFor explist / NumFor(start,limit[,step])
    Expr *e = nullptr;             // Call / While cond / Repeat cond / LocalFunction fn
    Block *body = nullptr;         // Do / While / Repeat / NumFor / GenFor
    std::vector<Expr *> conds;     // If
    std::vector<Block *> blocks;   // If (conds.size() or conds.size()+1 entries)
    std::string label;             // Goto / Label
};

struct Block {
    std::vector<Stmt *> stmts;
};

struct UpvalDesc {
    std::string name;
    bool from_parent_local = false;
    VarInfo *var = nullptr;  // when from_parent_local
    int index = 0;           // else parent's upvalue index
};

struct Proto {
    int nparams = 0;
    bool is_vararg = false;
    int nslots = 0;
    std::vector<VarInfo *> params;
    std::vector<UpvalDesc> upvals;
    Block *body = nullptr;
    std::string name;
    int line = 0;
};

struct Chunk {
    std::string name;
    std::vector<std::unique_ptr<Expr>> exprs;
    std::vector<std::unique_ptr<Stmt>> stmts;
    std::vector<std::unique_ptr<Block>> blocks;
    std::vector<std::unique_ptr<Proto>> protos;
    std::vector<std::unique_ptr<VarInfo>> vars;
    Proto *main = nullptr;
};


}  // namespace minilua
