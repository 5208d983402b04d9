// minilua — a small Lua 5.2-subset evaluator for Blinky globe/lens scripts.
//
// Why it exists: the reference evaluates lens/globe scripts with an external
// liblua 5.2 (engine/NQ/fisheye.c:278-280, 1222-1265, 1545-1651; linked at
// engine/Makefile:834-841).  No Lua is available where this code is built, and
// the product must JIT-evaluate `lens_inverse` / `lens_forward` / `globe_plate`
// on the host when it builds a lensmap.  This is that evaluator.
//
// Scope: the language subset the globe/lens script API needs and then some —
// locals with lexical scoping, closures/upvalues, multiple assignment and
// multiple returns (with last-call expansion), numeric and generic `for`,
// `while`, `repeat..until`, `if/elseif/else`, `break`, `goto`-less blocks,
// tables (array + hash part), strings as values, the `math`, `table`, `string`
// (minimal) libraries and the usual base functions.  No metatables,
// coroutines, or `goto`.
//
// Numeric semantics follow Lua 5.2 exactly: every number is an IEEE double,
// every arithmetic node is evaluated on its own (this file's translation unit
// must be compiled with -ffp-contract=off), `^` is libm pow(), `%` is
// a - floor(a/b)*b, and math.* forwards to libm.
//
// Threading: a State is single-threaded.  Compiled code (Chunk/Proto) is
// immutable and may be shared by many States; State::clone() deep-copies the
// mutable part (globals, tables, closures, upvalue cells) so that independent
// worker threads can evaluate the same lens concurrently.
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace minilua {

class State;
class Value;
struct Proto;
struct Chunk;
struct Universe;

struct LuaError : public std::runtime_error {
    explicit LuaError(const std::string &m, bool pos = false) : std::runtime_error(m), positioned(pos) {}
    bool positioned;  // message already carries "chunk:line:"
};

enum class Type : uint8_t {
    Nil = 0,
    Boolean,
    Number,
    String,
    Table,
    Function,
    Box,  // internal: heap cell for a local captured by a closure
};

// Heap object base with an intrusive, NON-atomic reference count.  Objects
// never cross States (see State::clone), so no atomics are required.
struct Object {
    int rc = 0;
    // every object is linked into its owning State's heap list so that
    // reference cycles can be found (State::collect_cycles) and so that
    // nothing outlives the State.
    int gc_refs = 0;
    bool gc_mark = false;
    Object *gc_prev = nullptr, *gc_next = nullptr;
    State *owner = nullptr;
    virtual ~Object();
};

// Small vector with inline storage, used for argument and result lists.
template <typename T, int N>
class SmallVec {
public:
    SmallVec() : data_(reinterpret_cast<T *>(inline_)), size_(0), cap_(N) {}
    ~SmallVec() {
        clear();
        if (data_ != reinterpret_cast<T *>(inline_)) ::operator delete(data_);
    }
    SmallVec(const SmallVec &) = delete;
    SmallVec &operator=(const SmallVec &) = delete;
    int size() const { return size_; }
    bool empty() const { return size_ == 0; }
    T &operator[](int i) { return data_[i]; }
    const T &operator[](int i) const { return data_[i]; }
    T *data() { return data_; }
    const T *data() const { return data_; }
    void clear() {
        for (int i = 0; i < size_; ++i) data_[i].~T();
        size_ = 0;
    }
    void push_back(const T &v) {
        if (size_ == cap_) grow();
        new (&data_[size_++]) T(v);
    }
    void pop_back() { data_[--size_].~T(); }
    void resize_down(int n) {
        while (size_ > n) pop_back();
    }
    T &back() { return data_[size_ - 1]; }

private:
    void grow() {
        int ncap = cap_ * 2;
        T *nd = static_cast<T *>(::operator new(sizeof(T) * ncap));
        for (int i = 0; i < size_; ++i) {
            new (&nd[i]) T(data_[i]);
            data_[i].~T();
        }
        if (data_ != reinterpret_cast<T *>(inline_)) ::operator delete(data_);
        data_ = nd;
        cap_ = ncap;
    }
    alignas(T) unsigned char inline_[sizeof(T) * N];
    T *data_;
    int size_, cap_;
};

class Value {
public:
    Value() : t_(Type::Nil) { u_.n = 0; }
    Value(double n) : t_(Type::Number) { u_.n = n; }
    static Value boolean(bool b) {
        Value v;
        v.t_ = Type::Boolean;
        v.u_.b = b;
        return v;
    }
    static Value object(Type t, Object *o) {
        Value v;
        v.t_ = t;
        v.u_.o = o;
        ++o->rc;
        return v;
    }
    Value(const Value &o) : t_(o.t_), u_(o.u_) {
        if (t_ >= Type::String) ++u_.o->rc;
    }
    Value(Value &&o) noexcept : t_(o.t_), u_(o.u_) { o.t_ = Type::Nil; }
    Value &operator=(const Value &o) {
        if (o.t_ >= Type::String) ++o.u_.o->rc;
        release();
        t_ = o.t_;
        u_ = o.u_;
        return *this;
    }
    Value &operator=(Value &&o) noexcept {
        if (this != &o) {
            release();
            t_ = o.t_;
            u_ = o.u_;
            o.t_ = Type::Nil;
        }
        return *this;
    }
    ~Value() { release(); }

    Type type() const { return t_; }
    bool is_nil() const { return t_ == Type::Nil; }
    bool is_number() const { return t_ == Type::Number; }
    bool is_string() const { return t_ == Type::String; }
    bool is_table() const { return t_ == Type::Table; }
    bool is_function() const { return t_ == Type::Function; }
    bool is_boolean() const { return t_ == Type::Boolean; }
    bool truthy() const { return !(t_ == Type::Nil || (t_ == Type::Boolean && !u_.b)); }
    double num() const { return u_.n; }
    bool boolean_value() const { return u_.b; }
    Object *obj() const { return u_.o; }

    // string payload (only valid when is_string())
    const std::string &str() const;

    // Lua's lua_isnumber()/lua_tonumber(): numbers, and strings that parse as
    // numbers.  Returns false when not convertible.
    bool to_number(double *out) const;

    bool raw_equals(const Value &o) const;

private:
    void release() {
        if (t_ >= Type::String) {
            if (--u_.o->rc == 0) delete u_.o;
        }
        t_ = Type::Nil;
    }
    Type t_;
    union {
        double n;
        bool b;
        Object *o;
    } u_;
};

using ValueList = SmallVec<Value, 6>;

// C function callable from scripts.  Pushes its results onto `out`.
using CFunction = void (*)(State &L, const Value *args, int nargs, ValueList &out, void *ud);

struct Str : Object {
    std::string s;
    explicit Str(std::string v) : s(std::move(v)) {}
};

struct ValueHash {
    size_t operator()(const Value &v) const;
};
struct ValueEq {
    bool operator()(const Value &a, const Value &b) const { return a.raw_equals(b); }
};

struct Table : Object {
    std::vector<Value> arr;  // keys 1..arr.size()
    std::unordered_map<Value, Value, ValueHash, ValueEq> hash;
    // insertion-ordered key list of the hash part so iteration is deterministic
    std::vector<Value> hash_order;
    Value meta;  // metatable (nil or a table): setmetatable / getmetatable

    Value get(const Value &k) const;
    Value get_int(int64_t i) const;
    Value get_str(const std::string &k) const;
    void set(const Value &k, const Value &v);  // throws on nil/NaN key
    void set_int(int64_t i, const Value &v);
    int64_t length() const;  // border, as lua_rawlen
    // lua_next-style iteration: `pos` starts at 0; returns false at the end.
    bool next(size_t *pos, Value *k, Value *v) const;

private:
    void migrate();
};

struct Box : Object {
    Value v;
};

struct Function : Object {
    // C function
    CFunction cfn = nullptr;
    // optional numeric fast paths (math.*): one or two numbers in, one number out.  The
    // evaluator calls these directly when every argument already is a number, skipping
    // the argument/result lists; semantics are identical to going through `cfn`.
    double (*fast1)(double) = nullptr;
    double (*fast2)(double, double) = nullptr;
    void *ud = nullptr;
    const char *cname = nullptr;
    // Lua closure
    const Proto *proto = nullptr;
    std::shared_ptr<const Chunk> chunk;  // keeps the code alive
    std::vector<Box *> upvals;           // owned references (rc held)
    ~Function() override;
};

// Process-wide (or per family of cloned States) registry giving stable small
// integer ids to global names and string constants, so compiled code never
// holds pointers into a particular State.
struct Universe {
    std::mutex mu;
    std::unordered_map<std::string, int> global_ids;
    std::vector<std::string> global_names;
    std::unordered_map<std::string, int> kstr_ids;
    std::vector<std::string> kstrs;
    int global_id(const std::string &name);
    int kstr_id(const std::string &s);
};

class State {
public:
    State();
    explicit State(std::shared_ptr<Universe> u);
    ~State();
    State(const State &) = delete;
    State &operator=(const State &) = delete;

    // Deep copy of all mutable state; compiled code is shared.
    std::unique_ptr<State> clone() const;

    // --- globals -------------------------------------------------------
    Value get_global(const std::string &name);
    void set_global(const std::string &name, const Value &v);
    void register_function(const std::string &name, CFunction f, void *ud = nullptr);

    // --- loading / calling ---------------------------------------------
    // Compile `src` to a function value (throws LuaError on syntax errors).
    Value load(const std::string &src, const std::string &chunkname);
    Value load_file(const std::string &path);  // throws LuaError if unreadable
    // Compile and run, discarding results.
    void run(const std::string &src, const std::string &chunkname);
    // Call fn(args...) collecting all results (throws LuaError on runtime errors).
    void call(const Value &fn, const Value *args, int nargs, ValueList &out);

    // --- helpers -------------------------------------------------------
    Value new_string(const std::string &s);
    Value new_table();
    Value new_cfunction(CFunction f, void *ud, const char *name);
    static std::string tostring(const Value &v);
    static const char *type_name(const Value &v);

    // where print() output goes (default: stdout)
    using PrintSink = void (*)(const char *text, void *ud);
    void set_print_sink(PrintSink s, void *ud) {
        print_sink_ = s;
        print_ud_ = ud;
    }
    void emit_print(const std::string &s);

    std::shared_ptr<Universe> universe() const { return uni_; }

    // --- interpreter internals (public for the evaluator's free functions)
    Value &global_slot(int id) {
        if (id >= static_cast<int>(globals_.size())) globals_.resize(id + 64);
        return globals_[id];
    }
    const Value &kstr(int id);
    Value *stack_alloc(int n);
    void stack_free(int n);
    int depth = 0;
    // heap: allocation helpers link the object into this State's heap list
    Table *alloc_table();
    Box *alloc_box();
    Function *alloc_function();
    Str *alloc_str(std::string s);
    void untrack(Object *o);
    bool gc_pending() const { return gc_pending_; }
    // trial-deletion cycle collector; only call at points where every live
    // object is referenced by a Value (the evaluator does so on function entry)
    void collect_cycles();
    size_t heap_objects() const { return gc_count_; }

private:
    void open_libs();
    void track(Object *o);
    Object *gc_head_ = nullptr;
    size_t gc_count_ = 0;
    size_t gc_threshold_ = 1u << 16;
    bool gc_pending_ = false;
    std::shared_ptr<Universe> uni_;
    std::vector<Value> globals_;
    std::vector<Value> kstr_;
    std::vector<Value> stack_;
    size_t stack_top_ = 0;
    PrintSink print_sink_ = nullptr;
    void *print_ud_ = nullptr;
};

}  // namespace minilua
