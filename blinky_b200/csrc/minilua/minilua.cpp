// minilua — implementation.  See minilua.h for scope and rationale.
//
// Structure: lexer -> recursive-descent parser producing an AST whose
// variables are already resolved to (frame slot | upvalue index | global id)
// -> a switch-based tree evaluator.  Compile this file with
// -ffp-contract=off so that no a*b+c is fused (Lua evaluates each arithmetic
// op on its own; lensmap parity depends on it).
#include "minilua.h"
#include "minilua_ast.h"

#include <cerrno>
#include <cinttypes>
#include <algorithm>
#include <cmath>
#include <ctime>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace minilua {

// ---------------------------------------------------------------------------
// Value helpers
// ---------------------------------------------------------------------------

const std::string &Value::str() const { return static_cast<Str *>(u_.o)->s; }

static bool str_to_number(const std::string &s, double *out) {
    const char *p = s.c_str();
    while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\f' || *p == '\v') ++p;
    if (*p == '\0') return false;
    // reject things strtod accepts but Lua does not ("inf", "nan")
    const char *q = p;
    if (*q == '+' || *q == '-') ++q;
    if (!((*q >= '0' && *q <= '9') || *q == '.')) return false;
    char *end = nullptr;
    double v = strtod(p, &end);
    if (end == p) return false;
    while (*end == ' ' || *end == '\t' || *end == '\n' || *end == '\r' || *end == '\f' || *end == '\v') ++end;
    if (*end != '\0') return false;
    *out = v;
    return true;
}

bool Value::to_number(double *out) const {
    if (t_ == Type::Number) {
        *out = u_.n;
        return true;
    }
    if (t_ == Type::String) return str_to_number(str(), out);
    return false;
}

bool Value::raw_equals(const Value &o) const {
    if (t_ != o.t_) return false;
    switch (t_) {
        case Type::Nil: return true;
        case Type::Boolean: return u_.b == o.u_.b;
        case Type::Number: return u_.n == o.u_.n;
        case Type::String: return u_.o == o.u_.o || str() == o.str();
        default: return u_.o == o.u_.o;
    }
}

size_t ValueHash::operator()(const Value &v) const {
    switch (v.type()) {
        case Type::Nil: return 0;
        case Type::Boolean: return v.boolean_value() ? 1 : 2;
        case Type::Number: {
            double d = v.num();
            if (d == 0) d = 0;  // +0/-0 hash alike
            uint64_t b;
            memcpy(&b, &d, sizeof b);
            return std::hash<uint64_t>()(b);
        }
        case Type::String: return std::hash<std::string>()(v.str());
        default: return std::hash<const void *>()(v.obj());
    }
}

Object::~Object() {
    if (owner) owner->untrack(this);
}

Function::~Function() {
    for (Box *b : upvals)
        if (b && --b->rc == 0) delete b;
}

// ---------------------------------------------------------------------------
// Table
// ---------------------------------------------------------------------------

static inline bool as_array_index(const Value &k, int64_t *idx) {
    if (!k.is_number()) return false;
    double d = k.num();
    if (d >= 1 && d < 9.0e15 && d == std::floor(d)) {
        *idx = static_cast<int64_t>(d);
        return true;
    }
    return false;
}

Value Table::get_int(int64_t i) const {
    if (i >= 1 && static_cast<size_t>(i) <= arr.size()) return arr[static_cast<size_t>(i - 1)];
    if (hash.empty()) return Value();
    auto it = hash.find(Value(static_cast<double>(i)));
    return it == hash.end() ? Value() : it->second;
}

Value Table::get(const Value &k) const {
    int64_t i;
    if (as_array_index(k, &i)) return get_int(i);
    if (k.is_nil() || hash.empty()) return Value();
    auto it = hash.find(k);
    return it == hash.end() ? Value() : it->second;
}

Value Table::get_str(const std::string &k) const {
    if (hash.empty()) return Value();
    for (auto &kv : hash)
        if (kv.first.is_string() && kv.first.str() == k) return kv.second;
    return Value();
}

void Table::migrate() {
    // move keys arr.size()+1, +2, ... out of the hash part while they exist
    while (!hash.empty()) {
        Value k(static_cast<double>(arr.size() + 1));
        auto it = hash.find(k);
        if (it == hash.end()) break;
        arr.push_back(it->second);
        hash.erase(it);
        for (size_t j = 0; j < hash_order.size(); ++j)
            if (hash_order[j].raw_equals(k)) {
                hash_order.erase(hash_order.begin() + static_cast<long>(j));
                break;
            }
    }
}

void Table::set_int(int64_t i, const Value &v) {
    if (i >= 1 && static_cast<size_t>(i) <= arr.size()) {
        arr[static_cast<size_t>(i - 1)] = v;
        if (v.is_nil() && static_cast<size_t>(i) == arr.size()) {
            while (!arr.empty() && arr.back().is_nil()) arr.pop_back();
        }
        return;
    }
    if (i >= 1 && static_cast<size_t>(i) == arr.size() + 1) {
        if (v.is_nil()) return;
        arr.push_back(v);
        migrate();
        return;
    }
    Value k(static_cast<double>(i));
    auto it = hash.find(k);
    if (v.is_nil()) {
        if (it != hash.end()) {
            hash.erase(it);
            for (size_t j = 0; j < hash_order.size(); ++j)
                if (hash_order[j].raw_equals(k)) {
                    hash_order.erase(hash_order.begin() + static_cast<long>(j));
                    break;
                }
        }
        return;
    }
    if (it == hash.end()) {
        hash.emplace(k, v);
        hash_order.push_back(k);
    } else {
        it->second = v;
    }
}

void Table::set(const Value &k, const Value &v) {
    int64_t i;
    if (as_array_index(k, &i)) {
        set_int(i, v);
        return;
    }
    if (k.is_nil()) throw LuaError("table index is nil");
    if (k.is_number() && k.num() != k.num()) throw LuaError("table index is NaN");
    auto it = hash.find(k);
    if (v.is_nil()) {
        if (it != hash.end()) {
            hash.erase(it);
            for (size_t j = 0; j < hash_order.size(); ++j)
                if (hash_order[j].raw_equals(k)) {
                    hash_order.erase(hash_order.begin() + static_cast<long>(j));
                    break;
                }
        }
        return;
    }
    if (it == hash.end()) {
        hash.emplace(k, v);
        hash_order.push_back(k);
    } else {
        it->second = v;
    }
}

int64_t Table::length() const { return static_cast<int64_t>(arr.size()); }

bool Table::next(size_t *pos, Value *k, Value *v) const {
    size_t p = *pos;
    while (p < arr.size()) {
        if (!arr[p].is_nil()) {
            *k = Value(static_cast<double>(p + 1));
            *v = arr[p];
            *pos = p + 1;
            return true;
        }
        ++p;
    }
    size_t h = p - arr.size();
    if (h < hash_order.size()) {
        *k = hash_order[h];
        *v = hash.find(hash_order[h])->second;
        *pos = p + 1;
        return true;
    }
    return false;
}

// ---------------------------------------------------------------------------
// Universe
// ---------------------------------------------------------------------------

int Universe::global_id(const std::string &name) {
    std::lock_guard<std::mutex> g(mu);
    auto it = global_ids.find(name);
    if (it != global_ids.end()) return it->second;
    int id = static_cast<int>(global_names.size());
    global_names.push_back(name);
    global_ids.emplace(name, id);
    return id;
}

int Universe::kstr_id(const std::string &s) {
    std::lock_guard<std::mutex> g(mu);
    auto it = kstr_ids.find(s);
    if (it != kstr_ids.end()) return it->second;
    int id = static_cast<int>(kstrs.size());
    kstrs.push_back(s);
    kstr_ids.emplace(s, id);
    return id;
}

// ---------------------------------------------------------------------------
// Lexer
// ---------------------------------------------------------------------------

enum Tok {
    T_EOF = 256, T_NAME, T_NUMBER, T_STRING,
    T_AND, T_BREAK, T_DO, T_ELSE, T_ELSEIF, T_END, T_FALSE, T_FOR, T_FUNCTION, T_GOTO, T_IF, T_IN,
    T_LOCAL, T_NIL, T_NOT, T_OR, T_REPEAT, T_RETURN, T_THEN, T_TRUE, T_UNTIL, T_WHILE,
    T_CONCAT, T_DOTS, T_EQ, T_GE, T_LE, T_NE, T_DBCOLON,
};

static const struct {
    const char *w;
    int t;
} kKeywords[] = {
    {"and", T_AND}, {"break", T_BREAK}, {"do", T_DO}, {"else", T_ELSE}, {"elseif", T_ELSEIF},
    {"end", T_END}, {"false", T_FALSE}, {"for", T_FOR}, {"function", T_FUNCTION}, {"goto", T_GOTO},
    {"if", T_IF}, {"in", T_IN}, {"local", T_LOCAL}, {"nil", T_NIL}, {"not", T_NOT}, {"or", T_OR},
    {"repeat", T_REPEAT}, {"return", T_RETURN}, {"then", T_THEN}, {"true", T_TRUE},
    {"until", T_UNTIL}, {"while", T_WHILE},
};

struct Token {
    int t = T_EOF;
    double num = 0;
    std::string s;
    int line = 1;
};

class Lexer {
public:
    Lexer(const std::string &src, const std::string &chunk) : s_(src), chunk_(chunk) {
        // skip a leading '#' line (shebang), as luaL_loadfile does
        if (!s_.empty() && s_[0] == '#')
            while (p_ < s_.size() && s_[p_] != '\n') ++p_;
    }

    [[noreturn]] void error(const std::string &msg, int line) const {
        std::ostringstream o;
        o << chunk_ << ":" << line << ": " << msg;
        throw LuaError(o.str(), true);
    }

    Token next() {
        Token tk;
        for (;;) {
            if (p_ >= s_.size()) {
                tk.t = T_EOF;
                tk.line = line_;
                return tk;
            }
            char c = s_[p_];
            if (c == '\n') {
                ++line_;
                ++p_;
                continue;
            }
            if (c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v') {
                ++p_;
                continue;
            }
            if (c == '-' && peek(1) == '-') {
                p_ += 2;
                if (peek(0) == '[') {
                    int lvl = long_bracket_level();
                    if (lvl >= 0) {
                        std::string dummy;
                        read_long_string(lvl, &dummy, "comment");
                        continue;
                    }
                }
                while (p_ < s_.size() && s_[p_] != '\n') ++p_;
                continue;
            }
            break;
        }
        tk.line = line_;
        char c = s_[p_];
        if (isalpha(static_cast<unsigned char>(c)) || c == '_') {
            size_t b = p_;
            while (p_ < s_.size() && (isalnum(static_cast<unsigned char>(s_[p_])) || s_[p_] == '_')) ++p_;
            tk.s = s_.substr(b, p_ - b);
            tk.t = T_NAME;
            for (auto &kw : kKeywords)
                if (tk.s == kw.w) {
                    tk.t = kw.t;
                    break;
                }
            return tk;
        }
        if (isdigit(static_cast<unsigned char>(c)) || (c == '.' && isdigit(static_cast<unsigned char>(peek(1))))) {
            read_number(&tk);
            return tk;
        }
        switch (c) {
            case '"':
            case '\'': read_string(c, &tk); return tk;
            case '[': {
                int lvl = long_bracket_level();
                if (lvl >= 0) {
                    read_long_string(lvl, &tk.s, "string");
                    tk.t = T_STRING;
                    return tk;
                }
                ++p_;
                tk.t = '[';
                return tk;
            }
            case '=':
                if (peek(1) == '=') { p_ += 2; tk.t = T_EQ; } else { ++p_; tk.t = '='; }
                return tk;
            case '<':
                if (peek(1) == '=') { p_ += 2; tk.t = T_LE; } else { ++p_; tk.t = '<'; }
                return tk;
            case '>':
                if (peek(1) == '=') { p_ += 2; tk.t = T_GE; } else { ++p_; tk.t = '>'; }
                return tk;
            case '~':
                if (peek(1) == '=') { p_ += 2; tk.t = T_NE; return tk; }
                error("unexpected symbol near '~'", line_);
            case ':':
                if (peek(1) == ':') { p_ += 2; tk.t = T_DBCOLON; } else { ++p_; tk.t = ':'; }
                return tk;
            case '.':
                if (peek(1) == '.') {
                    if (peek(2) == '.') { p_ += 3; tk.t = T_DOTS; } else { p_ += 2; tk.t = T_CONCAT; }
                } else { ++p_; tk.t = '.'; }
                return tk;
            default:
                ++p_;
                tk.t = static_cast<unsigned char>(c);
                return tk;
        }
    }

private:
    char peek(size_t o) const { return p_ + o < s_.size() ? s_[p_ + o] : '\0'; }

    // at '[': returns level if "[" "="* "[" follows, else -1 (does not consume)
    int long_bracket_level() const {
        size_t q = p_ + 1;
        int lvl = 0;
        while (q < s_.size() && s_[q] == '=') { ++lvl; ++q; }
        if (q < s_.size() && s_[q] == '[') return lvl;
        return -1;
    }

    void read_long_string(int lvl, std::string *out, const char *what) {
        int start = line_;
        p_ += static_cast<size_t>(lvl) + 2;
        if (peek(0) == '\r') ++p_;
        if (peek(0) == '\n') { ++p_; ++line_; }
        for (;;) {
            if (p_ >= s_.size()) error(std::string("unfinished long ") + what, start);
            char c = s_[p_];
            if (c == ']') {
                size_t q = p_ + 1;
                int l2 = 0;
                while (q < s_.size() && s_[q] == '=') { ++l2; ++q; }
                if (l2 == lvl && q < s_.size() && s_[q] == ']') {
                    p_ = q + 1;
                    return;
                }
                out->push_back(c);
                ++p_;
            } else {
                if (c == '\n') ++line_;
                out->push_back(c);
                ++p_;
            }
        }
    }

    void read_number(Token *tk) {
        size_t b = p_;
        bool hex = (s_[p_] == '0' && (peek(1) == 'x' || peek(1) == 'X'));
        if (hex) p_ += 2;
        for (;;) {
            char c = peek(0);
            if (hex ? isxdigit(static_cast<unsigned char>(c)) : isdigit(static_cast<unsigned char>(c))) { ++p_; continue; }
            if (c == '.') { ++p_; continue; }
            if ((!hex && (c == 'e' || c == 'E')) || (hex && (c == 'p' || c == 'P'))) {
                ++p_;
                if (peek(0) == '+' || peek(0) == '-') ++p_;
                continue;
            }
            break;
        }
        // trailing alnum glued to a number is malformed
        while (isalnum(static_cast<unsigned char>(peek(0))) || peek(0) == '_') ++p_;
        std::string txt = s_.substr(b, p_ - b);
        char *end = nullptr;
        double v = strtod(txt.c_str(), &end);
        if (end == txt.c_str() || *end != '\0') error("malformed number near '" + txt + "'", line_);
        tk->t = T_NUMBER;
        tk->num = v;
    }

    void read_string(char q, Token *tk) {
        int start = line_;
        ++p_;
        std::string out;
        for (;;) {
            if (p_ >= s_.size()) error("unfinished string", start);
            char c = s_[p_];
            if (c == q) { ++p_; break; }
            if (c == '\n') error("unfinished string", start);
            if (c == '\\') {
                ++p_;
                char e = peek(0);
                switch (e) {
                    case 'n': out.push_back('\n'); ++p_; break;
                    case 't': out.push_back('\t'); ++p_; break;
                    case 'r': out.push_back('\r'); ++p_; break;
                    case 'a': out.push_back('\a'); ++p_; break;
                    case 'b': out.push_back('\b'); ++p_; break;
                    case 'f': out.push_back('\f'); ++p_; break;
                    case 'v': out.push_back('\v'); ++p_; break;
                    case '\\': out.push_back('\\'); ++p_; break;
                    case '"': out.push_back('"'); ++p_; break;
                    case '\'': out.push_back('\''); ++p_; break;
                    case '\n': out.push_back('\n'); ++line_; ++p_; break;
                    case 'x': {
                        ++p_;
                        int v = 0;
                        for (int i = 0; i < 2; ++i) {
                            char h = peek(0);
                            if (!isxdigit(static_cast<unsigned char>(h))) error("hexadecimal digit expected", line_);
                            v = v * 16 + (isdigit(static_cast<unsigned char>(h)) ? h - '0' : (tolower(h) - 'a' + 10));
                            ++p_;
                        }
                        out.push_back(static_cast<char>(v));
                        break;
                    }
                    case 'z': {
                        ++p_;
                        while (p_ < s_.size() && isspace(static_cast<unsigned char>(s_[p_]))) {
                            if (s_[p_] == '\n') ++line_;
                            ++p_;
                        }
                        break;
                    }
                    default: {
                        if (isdigit(static_cast<unsigned char>(e))) {
                            int v = 0;
                            for (int i = 0; i < 3 && isdigit(static_cast<unsigned char>(peek(0))); ++i) {
                                v = v * 10 + (peek(0) - '0');
                                ++p_;
                            }
                            if (v > 255) error("decimal escape too large", line_);
                            out.push_back(static_cast<char>(v));
                        } else {
                            error("invalid escape sequence", line_);
                        }
                    }
                }
                continue;
            }
            out.push_back(c);
            ++p_;
        }
        tk->t = T_STRING;
        tk->s = out;
    }

    const std::string &s_;
    std::string chunk_;
    size_t p_ = 0;
    int line_ = 1;
};

// ---------------------------------------------------------------------------
// Parser
// ---------------------------------------------------------------------------

struct FuncState {
    FuncState *parent = nullptr;
    Proto *proto = nullptr;
    struct Active {
        std::string name;
        VarInfo *var;
    };
    std::vector<Active> actives;           // visible locals, innermost last
    std::vector<size_t> block_starts;      // actives.size() at block entry
    std::vector<int> slot_starts;          // nslot at block entry
    int nslot = 0;
    std::vector<std::string> upval_names;
    int loop_depth = 0;
};

class Parser {
public:
    Parser(const std::string &src, const std::string &chunkname, Universe *uni)
        : lex_(src, chunkname), uni_(uni), chunk_(std::make_shared<Chunk>()) {
        chunk_->name = chunkname;
        advance();
    }

    std::shared_ptr<Chunk> parse_chunk() {
        FuncState fs;
        Proto *p = new_proto();
        p->is_vararg = true;
        p->name = "main chunk";
        fs.proto = p;
        fs_ = &fs;
        open_block();
        p->body = parse_block();
        close_block();
        if (tok_.t != T_EOF) err_expected("<eof>");
        p->nslots = max_slots_[p];
        chunk_->main = p;
        fs_ = nullptr;
        return chunk_;
    }

private:
    // --- arena helpers ---
    Expr *new_expr(EK k, int line) {
        chunk_->exprs.emplace_back(new Expr());
        Expr *e = chunk_->exprs.back().get();
        e->k = k;
        e->line = line;
        return e;
    }
    Stmt *new_stmt(SK k, int line) {
        chunk_->stmts.emplace_back(new Stmt());
        Stmt *s = chunk_->stmts.back().get();
        s->k = k;
        s->line = line;
        return s;
    }
    Block *new_block() {
        chunk_->blocks.emplace_back(new Block());
        return chunk_->blocks.back().get();
    }
    Proto *new_proto() {
        chunk_->protos.emplace_back(new Proto());
        return chunk_->protos.back().get();
    }
    VarInfo *new_var() {
        chunk_->vars.emplace_back(new VarInfo());
        return chunk_->vars.back().get();
    }

    // --- token helpers ---
    void advance() {
        if (has_ahead_) {
            tok_ = ahead_;
            has_ahead_ = false;
        } else {
            tok_ = lex_.next();
        }
    }
    const Token &lookahead() {
        if (!has_ahead_) {
            ahead_ = lex_.next();
            has_ahead_ = true;
        }
        return ahead_;
    }
    std::string tok_text(const Token &t) const {
        switch (t.t) {
            case T_EOF: return "<eof>";
            case T_NAME: return t.s;
            case T_STRING: return t.s;
            case T_NUMBER: {
                char b[64];
                snprintf(b, sizeof b, "%.14g", t.num);
                return b;
            }
            case T_CONCAT: return "..";
            case T_DOTS: return "...";
            case T_EQ: return "==";
            case T_GE: return ">=";
            case T_LE: return "<=";
            case T_NE: return "~=";
            case T_DBCOLON: return "::";
            default:
                for (auto &kw : kKeywords)
                    if (kw.t == t.t) return kw.w;
                return std::string(1, static_cast<char>(t.t));
        }
    }
    [[noreturn]] void err_expected(const std::string &what) {
        lex_.error("'" + what + "' expected near '" + tok_text(tok_) + "'", tok_.line);
    }
    [[noreturn]] void err(const std::string &msg) { lex_.error(msg + " near '" + tok_text(tok_) + "'", tok_.line); }
    bool check(int t) const { return tok_.t == t; }
    bool accept(int t) {
        if (tok_.t == t) {
            advance();
            return true;
        }
        return false;
    }
    void expect(int t, const char *what) {
        if (tok_.t != t) err_expected(what);
        advance();
    }
    void expect_match(int t, const char *what, const char *opener, int line) {
        if (tok_.t != t) {
            if (line == tok_.line) err_expected(what);
            std::ostringstream o;
            o << "'" << what << "' expected (to close '" << opener << "' at line " << line << ") near '" << tok_text(tok_) << "'";
            lex_.error(o.str(), tok_.line);
        }
        advance();
    }
    std::string expect_name() {
        if (tok_.t != T_NAME) err_expected("<name>");
        std::string s = tok_.s;
        advance();
        return s;
    }

    // --- scopes ---
    void open_block() {
        fs_->block_starts.push_back(fs_->actives.size());
        fs_->slot_starts.push_back(fs_->nslot);
    }
    void close_block() {
        fs_->actives.resize(fs_->block_starts.back());
        fs_->nslot = fs_->slot_starts.back();
        fs_->block_starts.pop_back();
        fs_->slot_starts.pop_back();
    }
    VarInfo *declare_local(const std::string &name) {
        VarInfo *v = new_var();
        v->name = name;
        v->slot = fs_->nslot++;
        int &mx = max_slots_[fs_->proto];
        if (fs_->nslot > mx) mx = fs_->nslot;
        fs_->actives.push_back({name, v});
        return v;
    }
    // reserve a slot now, make the name visible later (Lua: `local x = x`)
    VarInfo *reserve_local() {
        VarInfo *v = new_var();
        v->slot = fs_->nslot++;
        int &mx = max_slots_[fs_->proto];
        if (fs_->nslot > mx) mx = fs_->nslot;
        return v;
    }
    void activate_local(const std::string &name, VarInfo *v) {
        v->name = name;
        fs_->actives.push_back({name, v});
    }

    static VarInfo *find_local(FuncState *fs, const std::string &name) {
        for (size_t i = fs->actives.size(); i-- > 0;)
            if (fs->actives[i].name == name) return fs->actives[i].var;
        return nullptr;
    }
    // returns upvalue index in fs, or -1 when `name` is not a local of any enclosing function
    int find_upval(FuncState *fs, const std::string &name) {
        for (size_t i = 0; i < fs->upval_names.size(); ++i)
            if (fs->upval_names[i] == name) return static_cast<int>(i);
        if (!fs->parent) return -1;
        UpvalDesc d;
        if (VarInfo *v = find_local(fs->parent, name)) {
            v->captured = true;
            d.from_parent_local = true;
            d.var = v;
        } else {
            int pi = find_upval(fs->parent, name);
            if (pi < 0) return -1;
            d.from_parent_local = false;
            d.index = pi;
        }
        d.name = name;
        fs->proto->upvals.push_back(d);
        fs->upval_names.push_back(name);
        return static_cast<int>(fs->upval_names.size()) - 1;
    }
    Expr *resolve_name(const std::string &name, int line) {
        if (VarInfo *v = find_local(fs_, name)) {
            Expr *e = new_expr(EK::Local, line);
            e->var = v;
            return e;
        }
        int ui = find_upval(fs_, name);
        if (ui >= 0) {
            Expr *e = new_expr(EK::Upval, line);
            e->id = ui;
            return e;
        }
        Expr *e = new_expr(EK::Global, line);
        e->id = uni_->global_id(name);
        return e;
    }
    Expr *string_const(const std::string &s, int line) {
        Expr *e = new_expr(EK::String, line);
        e->id = uni_->kstr_id(s);
        return e;
    }

    // --- statements ---
    bool block_follow() const {
        switch (tok_.t) {
            case T_ELSE: case T_ELSEIF: case T_END: case T_EOF: case T_UNTIL: return true;
            default: return false;
        }
    }

    Block *parse_block() {
        Block *b = new_block();
        while (!block_follow()) {
            if (check(T_RETURN)) {
                b->stmts.push_back(parse_return());
                break;
            }
            Stmt *s = parse_statement();
            if (s) b->stmts.push_back(s);
        }
        return b;
    }

    Stmt *parse_return() {
        int line = tok_.line;
        advance();
        Stmt *s = new_stmt(SK::Return, line);
        if (!block_follow() && !check(';')) parse_exprlist(&s->exprs);
        accept(';');
        return s;
    }

    Stmt *parse_statement() {
        int line = tok_.line;
        switch (tok_.t) {
            case ';': advance(); return nullptr;
            case T_IF: return parse_if();
            case T_WHILE: {
                advance();
                Stmt *s = new_stmt(SK::While, line);
                s->e = parse_expr();
                expect(T_DO, "do");
                ++fs_->loop_depth;
                open_block();
                s->body = parse_block();
                close_block();
                --fs_->loop_depth;
                expect_match(T_END, "end", "while", line);
                return s;
            }
            case T_DO: {
                advance();
                Stmt *s = new_stmt(SK::Do, line);
                open_block();
                s->body = parse_block();
                close_block();
                expect_match(T_END, "end", "do", line);
                return s;
            }
            case T_FOR: return parse_for();
            case T_REPEAT: {
                advance();
                Stmt *s = new_stmt(SK::Repeat, line);
                ++fs_->loop_depth;
                open_block();
                s->body = parse_block();
                expect_match(T_UNTIL, "until", "repeat", line);
                s->e = parse_expr();  // sees the body's locals
                close_block();
                --fs_->loop_depth;
                return s;
            }
            case T_FUNCTION: return parse_function_stat();
            case T_LOCAL: {
                advance();
                if (accept(T_FUNCTION)) {
                    Stmt *s = new_stmt(SK::LocalFunction, line);
                    std::string name = expect_name();
                    VarInfo *v = declare_local(name);  // visible inside its own body
                    s->vars.push_back(v);
                    s->e = parse_function_body(false, name, line);
                    return s;
                }
                Stmt *s = new_stmt(SK::Local, line);
                std::vector<std::string> names;
                do {
                    names.push_back(expect_name());
                } while (accept(','));
                for (size_t i = 0; i < names.size(); ++i) s->vars.push_back(reserve_local());
                if (accept('=')) parse_exprlist(&s->exprs);
                for (size_t i = 0; i < names.size(); ++i) activate_local(names[i], s->vars[i]);
                return s;
            }
            case T_RETURN: return parse_return();
            case T_BREAK: {
                advance();
                if (fs_->loop_depth == 0) lex_.error("<break> at line " + std::to_string(line) + " not inside a loop", line);
                return new_stmt(SK::Break, line);
            }
            case T_GOTO: {   // Lua 5.2 (the reference links liblua 5.2): resolved when executed, see exec_block
                advance();
                Stmt *s = new_stmt(SK::Goto, line);
                s->label = expect_name();
                return s;
            }
            case T_DBCOLON: {
                advance();
                Stmt *s = new_stmt(SK::Label, line);
                s->label = expect_name();
                if (!accept(T_DBCOLON)) err("'::' expected");
                return s;
            }
            default: return parse_expr_stat();
        }
    }

    Stmt *parse_if() {
        int line = tok_.line;
        Stmt *s = new_stmt(SK::If, line);
        advance();  // if
        for (;;) {
            s->conds.push_back(parse_expr());
            expect(T_THEN, "then");
            open_block();
            s->blocks.push_back(parse_block());
            close_block();
            if (accept(T_ELSEIF)) continue;
            if (accept(T_ELSE)) {
                open_block();
                s->blocks.push_back(parse_block());
                close_block();
            }
            expect_match(T_END, "end", "if", line);
            break;
        }
        return s;
    }

    Stmt *parse_for() {
        int line = tok_.line;
        advance();  // for
        std::string n1 = expect_name();
        if (check('=')) {
            advance();
            Stmt *s = new_stmt(SK::NumFor, line);
            s->exprs.push_back(parse_expr());
            expect(',', ",");
            s->exprs.push_back(parse_expr());
            if (accept(',')) s->exprs.push_back(parse_expr());
            expect(T_DO, "do");
            ++fs_->loop_depth;
            open_block();
            s->vars.push_back(declare_local(n1));
            open_block();
            s->body = parse_block();
            close_block();
            close_block();
            --fs_->loop_depth;
            expect_match(T_END, "end", "for", line);
            return s;
        }
        if (check(',') || check(T_IN)) {
            Stmt *s = new_stmt(SK::GenFor, line);
            std::vector<std::string> names{n1};
            while (accept(',')) names.push_back(expect_name());
            expect(T_IN, "in");
            parse_exprlist(&s->exprs);
            expect(T_DO, "do");
            ++fs_->loop_depth;
            open_block();
            for (auto &n : names) s->vars.push_back(declare_local(n));
            open_block();
            s->body = parse_block();
            close_block();
            close_block();
            --fs_->loop_depth;
            expect_match(T_END, "end", "for", line);
            return s;
        }
        err_expected("=' or 'in");
    }

    Stmt *parse_function_stat() {
        int line = tok_.line;
        advance();  // function
        std::string name = expect_name();
        Expr *target = resolve_name(name, line);
        std::string full = name;
        bool method = false;
        while (check('.') || check(':')) {
            bool colon = check(':');
            advance();
            std::string key = expect_name();
            Expr *idx = new_expr(EK::Index, line);
            idx->l = target;
            idx->r = string_const(key, line);
            target = idx;
            full += (colon ? ":" : ".") + key;
            if (colon) {
                method = true;
                break;
            }
        }
        Stmt *s = new_stmt(SK::Assign, line);
        s->targets.push_back(target);
        s->exprs.push_back(parse_function_body(method, full, line));
        return s;
    }

    Stmt *parse_expr_stat() {
        int line = tok_.line;
        Expr *e = parse_suffixed();
        if (check('=') || check(',')) {
            Stmt *s = new_stmt(SK::Assign, line);
            check_assignable(e);
            s->targets.push_back(e);
            while (accept(',')) {
                Expr *t = parse_suffixed();
                check_assignable(t);
                s->targets.push_back(t);
            }
            expect('=', "=");
            parse_exprlist(&s->exprs);
            return s;
        }
        if (e->k != EK::Call && e->k != EK::Method) err("syntax error");
        Stmt *s = new_stmt(SK::Call, line);
        s->e = e;
        return s;
    }

    void check_assignable(Expr *e) {
        if (e->k != EK::Local && e->k != EK::Upval && e->k != EK::Global && e->k != EK::Index) err("syntax error");
    }

    // --- expressions ---
    void parse_exprlist(std::vector<Expr *> *out) {
        out->push_back(parse_expr());
        while (accept(',')) out->push_back(parse_expr());
    }

    Expr *parse_function_body(bool method, const std::string &name, int line) {
        FuncState nfs;
        nfs.parent = fs_;
        Proto *p = new_proto();
        p->name = name;
        p->line = line;
        nfs.proto = p;
        FuncState *saved = fs_;
        fs_ = &nfs;
        open_block();
        if (method) p->params.push_back(declare_local("self"));
        expect('(', "(");
        if (!check(')')) {
            do {
                if (check(T_DOTS)) {
                    advance();
                    p->is_vararg = true;
                    break;
                }
                p->params.push_back(declare_local(expect_name()));
            } while (accept(','));
        }
        expect(')', ")");
        p->nparams = static_cast<int>(p->params.size());
        p->body = parse_block();
        expect_match(T_END, "end", "function", line);
        close_block();
        p->nslots = max_slots_[p];
        fs_ = saved;
        Expr *e = new_expr(EK::Function, line);
        e->proto = p;
        return e;
    }

    Expr *parse_primary() {
        int line = tok_.line;
        if (check(T_NAME)) {
            std::string n = tok_.s;
            advance();
            return resolve_name(n, line);
        }
        if (accept('(')) {
            Expr *inner = parse_expr();
            expect_match(')', ")", "(", line);
            Expr *e = new_expr(EK::Paren, line);
            e->l = inner;
            return e;
        }
        err("unexpected symbol");
    }

    void parse_call_args(Expr *call) {
        int line = tok_.line;
        if (check(T_STRING)) {
            call->list.push_back(string_const(tok_.s, line));
            advance();
            return;
        }
        if (check('{')) {
            call->list.push_back(parse_table());
            return;
        }
        if (!check('(')) err("function arguments expected");
        advance();
        if (!check(')')) parse_exprlist(&call->list);
        expect_match(')', ")", "(", line);
    }

    Expr *parse_suffixed() {
        Expr *e = parse_primary();
        for (;;) {
            int line = tok_.line;
            switch (tok_.t) {
                case '.': {
                    advance();
                    Expr *idx = new_expr(EK::Index, line);
                    idx->l = e;
                    idx->r = string_const(expect_name(), line);
                    e = idx;
                    break;
                }
                case '[': {
                    advance();
                    Expr *idx = new_expr(EK::Index, line);
                    idx->l = e;
                    idx->r = parse_expr();
                    expect(']', "]");
                    e = idx;
                    break;
                }
                case ':': {
                    advance();
                    Expr *m = new_expr(EK::Method, line);
                    m->l = e;
                    m->id = uni_->kstr_id(expect_name());
                    parse_call_args(m);
                    e = m;
                    break;
                }
                case '(': case T_STRING: case '{': {
                    Expr *c = new_expr(EK::Call, line);
                    c->l = e;
                    parse_call_args(c);
                    e = c;
                    break;
                }
                default: return e;
            }
        }
    }

    Expr *parse_table() {
        int line = tok_.line;
        expect('{', "{");
        Expr *t = new_expr(EK::Table, line);
        while (!check('}')) {
            if (check('[')) {
                advance();
                Expr *k = parse_expr();
                expect(']', "]");
                expect('=', "=");
                t->keys.push_back(k);
                t->vals.push_back(parse_expr());
            } else if (check(T_NAME) && lookahead().t == '=') {
                Expr *k = string_const(tok_.s, tok_.line);
                advance();
                advance();
                t->keys.push_back(k);
                t->vals.push_back(parse_expr());
            } else {
                t->list.push_back(parse_expr());
            }
            if (!accept(',') && !accept(';')) break;
        }
        expect_match('}', "}", "{", line);
        return t;
    }

    Expr *parse_simple() {
        int line = tok_.line;
        switch (tok_.t) {
            case T_NUMBER: {
                Expr *e = new_expr(EK::Number, line);
                e->num = tok_.num;
                advance();
                return e;
            }
            case T_STRING: {
                Expr *e = string_const(tok_.s, line);
                advance();
                return e;
            }
            case T_NIL: advance(); return new_expr(EK::Nil, line);
            case T_TRUE: advance(); return new_expr(EK::True, line);
            case T_FALSE: advance(); return new_expr(EK::False, line);
            case T_DOTS:
                if (!fs_->proto->is_vararg) err("cannot use '...' outside a vararg function");
                advance();
                return new_expr(EK::Vararg, line);
            case '{': return parse_table();
            case T_FUNCTION: advance(); return parse_function_body(false, "anonymous", line);
            default: return parse_suffixed();
        }
    }

    // operator precedence (Lua 5.2 manual §3.4.7): left, right binding powers
    static bool binop_info(int t, EK *k, int *lp, int *rp) {
        switch (t) {
            case T_OR: *k = EK::Or; *lp = 1; *rp = 1; return true;
            case T_AND: *k = EK::And; *lp = 2; *rp = 2; return true;
            case '<': *k = EK::Lt; *lp = 3; *rp = 3; return true;
            case '>': *k = EK::Gt; *lp = 3; *rp = 3; return true;
            case T_LE: *k = EK::Le; *lp = 3; *rp = 3; return true;
            case T_GE: *k = EK::Ge; *lp = 3; *rp = 3; return true;
            case T_NE: *k = EK::Ne; *lp = 3; *rp = 3; return true;
            case T_EQ: *k = EK::Eq; *lp = 3; *rp = 3; return true;
            case T_CONCAT: *k = EK::Concat; *lp = 5; *rp = 4; return true;  // right assoc
            case '+': *k = EK::Add; *lp = 6; *rp = 6; return true;
            case '-': *k = EK::Sub; *lp = 6; *rp = 6; return true;
            case '*': *k = EK::Mul; *lp = 7; *rp = 7; return true;
            case '/': *k = EK::Div; *lp = 7; *rp = 7; return true;
            case '%': *k = EK::Mod; *lp = 7; *rp = 7; return true;
            case '^': *k = EK::Pow; *lp = 10; *rp = 9; return true;  // right assoc
            default: return false;
        }
    }
    static const int kUnaryPriority = 8;

    Expr *parse_subexpr(int limit) {
        Expr *e;
        int line = tok_.line;
        if (check(T_NOT) || check('-') || check('#')) {
            EK k = check(T_NOT) ? EK::Not : (check('-') ? EK::Neg : EK::Len);
            advance();
            Expr *operand = parse_subexpr(kUnaryPriority);
            if (k == EK::Neg && operand->k == EK::Number) {
                operand->num = -operand->num;  // constant fold, same as luac
                e = operand;
            } else {
                e = new_expr(k, line);
                e->l = operand;
            }
        } else {
            e = parse_simple();
        }
        EK k;
        int lp, rp;
        while (binop_info(tok_.t, &k, &lp, &rp) && lp > limit) {
            int oline = tok_.line;
            advance();
            Expr *rhs = parse_subexpr(rp);
            Expr *b = new_expr(k, oline);
            b->l = e;
            b->r = rhs;
            e = b;
        }
        return e;
    }

    Expr *parse_expr() { return parse_subexpr(0); }

    Lexer lex_;
    Universe *uni_;
    std::shared_ptr<Chunk> chunk_;
    Token tok_, ahead_;
    bool has_ahead_ = false;
    FuncState *fs_ = nullptr;
    std::unordered_map<Proto *, int> max_slots_;
};

// ---------------------------------------------------------------------------
// Evaluator
// ---------------------------------------------------------------------------

namespace {

struct Frame {
    Value *base;
    Function *fn;
    const Value *varargs;
    int nvarargs;
    const Chunk *chunk;
    const Stmt *pending_goto = nullptr;   // the `goto` being unwound (Flow::Goto)
};

// Goto: a `goto` looking for its label — every enclosing block is searched on the way out (a label is visible in the
// block that holds it and the blocks nested in it, which is where Lua 5.2 lets a goto sit); loops and ifs pass it on
enum class Flow { Normal, Break, Return, Goto };

[[noreturn]] void rt_error(const Frame &f, int line, const std::string &msg) {
    std::ostringstream o;
    o << (f.chunk ? f.chunk->name : std::string("?")) << ":" << line << ": " << msg;
    throw LuaError(o.str(), true);
}

void call_value(State &L, const Value &fn, const Value *args, int nargs, ValueList &out, const Frame *caller, int line);
void eval(State &L, Frame &f, const Expr *e, Value &out);
void eval_multi(State &L, Frame &f, const Expr *e, ValueList &out);
Flow exec_block(State &L, Frame &f, const Block *b, ValueList &ret);

inline Value &slot_ref(Frame &f, const VarInfo *v) {
    Value &s = f.base[v->slot];
    if (s.type() == Type::Box) return static_cast<Box *>(s.obj())->v;
    return s;
}

inline void declare_slot(State &L, Frame &f, const VarInfo *v, const Value &val) {
    if (v->captured) {
        Box *b = L.alloc_box();
        b->v = val;
        f.base[v->slot] = Value::object(Type::Box, b);
    } else {
        f.base[v->slot] = val;
    }
}

const char *arith_name(EK k) {
    (void)k;
    return "perform arithmetic on";
}

bool coerce_num(const Value &v, double *out) { return v.to_number(out); }

inline double do_arith(EK k, double a, double b) {
    switch (k) {
        case EK::Add: return a + b;
        case EK::Sub: return a - b;
        case EK::Mul: return a * b;
        case EK::Div: return a / b;
        case EK::Mod: return a - std::floor(a / b) * b;  // luai_nummod (Lua 5.2 luaconf.h)
        case EK::Pow: return std::pow(a, b);
        default: return 0;
    }
}

std::string number_to_string(double d) {
    char b[64];
    snprintf(b, sizeof b, "%.14g", d);
    return b;
}

// ---- metatables: only tables carry one; every use sits on a path that would otherwise raise an error or
// return nil, so programs without metatables (all shipped lenses) run exactly as before
Value metamethod(const Value &v, const char *name) {
    if (!v.is_table()) return Value();
    const Value &m = static_cast<const Table *>(v.obj())->meta;
    if (!m.is_table()) return Value();
    return static_cast<const Table *>(m.obj())->get_str(name);
}
// the handler of a binary event: the left operand's, else the right one's
Value binary_handler(const Value &a, const Value &b, const char *name) {
    Value h = metamethod(a, name);
    return h.is_nil() ? metamethod(b, name) : h;
}
bool call_binary(State &L, const Frame &f, int line, const Value &h, const Value &a, const Value &b, Value &out) {
    if (h.is_nil()) return false;
    const Value args[2] = {a, b};
    ValueList rets;
    call_value(L, h, args, 2, rets, &f, line);
    out = rets.size() > 0 ? rets[0] : Value();
    return true;
}

bool less_than(State &L, const Frame &f, int line, const Value &a, const Value &b) {
    if (a.is_number() && b.is_number()) return a.num() < b.num();
    if (a.is_string() && b.is_string()) return a.str() < b.str();
    Value r;
    if (call_binary(L, f, line, binary_handler(a, b, "__lt"), a, b, r)) return r.truthy();
    if (a.type() == b.type())
        rt_error(f, line, std::string("attempt to compare two ") + State::type_name(a) + " values");
    rt_error(f, line, std::string("attempt to compare ") + State::type_name(a) + " with " + State::type_name(b));
}

bool less_equal(State &L, const Frame &f, int line, const Value &a, const Value &b) {
    if (a.is_number() && b.is_number()) return a.num() <= b.num();
    if (a.is_string() && b.is_string()) return a.str() <= b.str();
    Value r;
    if (call_binary(L, f, line, binary_handler(a, b, "__le"), a, b, r)) return r.truthy();
    if (call_binary(L, f, line, binary_handler(b, a, "__lt"), b, a, r)) return !r.truthy();   // a <= b  ==  not (b < a)
    if (a.type() == b.type())
        rt_error(f, line, std::string("attempt to compare two ") + State::type_name(a) + " values");
    rt_error(f, line, std::string("attempt to compare ") + State::type_name(a) + " with " + State::type_name(b));
}

std::string describe(const Expr *e, const Frame &f) {
    // best-effort variable naming for error messages, like Lua's varinfo
    (void)f;
    switch (e->k) {
        case EK::Global: return "global";
        case EK::Local: return "local";
        case EK::Upval: return "upvalue";
        case EK::Index: return "field";
        default: return "";
    }
}

void index_value(State &L, Frame &f, const Expr *e, const Value &obj, const Value &key, Value &out) {
    if (obj.is_table()) {
        const Table *t = static_cast<Table *>(obj.obj());
        out = t->get(key);
        if (!out.is_nil() || !t->meta.is_table()) return;
        // __index: a table to look in next (chains), or a function(table, key)
        Value cur = obj;
        for (int hops = 0; hops < 100; ++hops) {
            Value h = metamethod(cur, "__index");
            if (h.is_nil()) return;
            if (h.is_function()) {
                const Value args[2] = {cur, key};
                ValueList rets;
                call_value(L, h, args, 2, rets, &f, e->line);
                out = rets.size() > 0 ? rets[0] : Value();
                return;
            }
            if (!h.is_table()) rt_error(f, e->line, std::string("attempt to index a ") + State::type_name(h) + " value");
            out = static_cast<Table *>(h.obj())->get(key);
            if (!out.is_nil()) return;
            cur = h;
        }
        rt_error(f, e->line, "loop in gettable");
    }
    if (obj.is_string()) {  // ("x"):len() style access goes through the string library
        Value lib = L.get_global("string");
        if (lib.is_table()) {
            out = static_cast<Table *>(lib.obj())->get(key);
            return;
        }
    }
    std::string what = describe(e->l, f);
    rt_error(f, e->line, "attempt to index " + (what.empty() ? std::string("a ") : what + " (a ") +
                             State::type_name(obj) + " value" + (what.empty() ? "" : ")"));
}

void make_closure(State &L, Frame &f, const Expr *e, Value &out) {
    Function *fn = L.alloc_function();
    fn->proto = e->proto;
    // share ownership of the code with the defining closure
    fn->chunk = f.fn ? f.fn->chunk : nullptr;
    fn->upvals.reserve(e->proto->upvals.size());
    for (const UpvalDesc &d : e->proto->upvals) {
        Box *b;
        if (d.from_parent_local) {
            Value &s = f.base[d.var->slot];
            if (s.type() != Type::Box) {
                // captured variable that has not been boxed yet (declared before
                // capture analysis marked it) — box it in place now
                Box *nb = L.alloc_box();
                nb->v = s;
                s = Value::object(Type::Box, nb);
            }
            b = static_cast<Box *>(s.obj());
        } else {
            b = f.fn->upvals[static_cast<size_t>(d.index)];
        }
        ++b->rc;
        fn->upvals.push_back(b);
    }
    out = Value::object(Type::Function, fn);
}

// If `e` names a variable, returns the variable's storage (no copy, no refcount traffic);
// otherwise evaluates into `tmp` and returns that.
inline const Value *operand(State &L, Frame &f, const Expr *e, Value &tmp) {
    switch (e->k) {
        case EK::Local: return &slot_ref(f, e->var);
        case EK::Upval: return &f.fn->upvals[static_cast<size_t>(e->id)]->v;
        case EK::Global: return &L.global_slot(e->id);
        default: eval(L, f, e, tmp); return &tmp;
    }
}

void eval_args(State &L, Frame &f, const std::vector<Expr *> &list, ValueList &out) {
    size_t n = list.size();
    for (size_t i = 0; i < n; ++i) {
        const Expr *a = list[i];
        if (i + 1 == n && (a->k == EK::Call || a->k == EK::Method || a->k == EK::Vararg)) {
            eval_multi(L, f, a, out);
        } else {
            Value v;
            eval(L, f, a, v);
            out.push_back(v);
        }
    }
}

void eval_call(State &L, Frame &f, const Expr *e, ValueList &out) {
    Value fn;
    ValueList args;
    if (e->k == EK::Method) {
        Value obj;
        eval(L, f, e->l, obj);
        index_value(L, f, e, obj, L.kstr(e->id), fn);
        args.push_back(obj);
    } else {
        eval(L, f, e->l, fn);
    }
    eval_args(L, f, e->list, args);
    if (!fn.is_function() && !metamethod(fn, "__call").is_nil()) {   // callable table: handler(table, args...)
        ValueList with_self;
        with_self.push_back(fn);
        for (int i = 0; i < args.size(); ++i) with_self.push_back(args[i]);
        call_value(L, metamethod(fn, "__call"), with_self.data(), with_self.size(), out, &f, e->line);
        return;
    }
    if (!fn.is_function()) {
        std::string what;
        const Expr *c = e->l;
        if (e->k == EK::Method) {
            what = "method '" + L.kstr(e->id).str() + "'";
        } else if (c->k == EK::Global) {
            what = "global '" + L.universe()->global_names[static_cast<size_t>(c->id)] + "'";
        } else if (c->k == EK::Index && c->r->k == EK::String) {
            what = "field '" + L.kstr(c->r->id).str() + "'";
        } else if (c->k == EK::Local) {
            what = "local";
        }
        rt_error(f, e->line, "attempt to call " + (what.empty() ? std::string("a ") : what + " (a ") +
                                 State::type_name(fn) + " value" + (what.empty() ? "" : ")"));
    }
    call_value(L, fn, args.data(), args.size(), out, &f, e->line);
}

void eval_multi(State &L, Frame &f, const Expr *e, ValueList &out) {
    switch (e->k) {
        case EK::Call:
        case EK::Method: eval_call(L, f, e, out); return;
        case EK::Vararg:
            for (int i = 0; i < f.nvarargs; ++i) out.push_back(f.varargs[i]);
            return;
        default: {
            Value v;
            eval(L, f, e, v);
            out.push_back(v);
        }
    }
}

void eval(State &L, Frame &f, const Expr *e, Value &out) {
    switch (e->k) {
        case EK::Nil: out = Value(); return;
        case EK::True: out = Value::boolean(true); return;
        case EK::False: out = Value::boolean(false); return;
        case EK::Number: out = Value(e->num); return;
        case EK::String: out = L.kstr(e->id); return;
        case EK::Vararg: out = f.nvarargs > 0 ? f.varargs[0] : Value(); return;
        case EK::Local: out = slot_ref(f, e->var); return;
        case EK::Upval: out = f.fn->upvals[static_cast<size_t>(e->id)]->v; return;
        case EK::Global: out = L.global_slot(e->id); return;
        case EK::Paren: eval(L, f, e->l, out); return;
        case EK::Index: {
            Value tobj, key;
            const Value *po = operand(L, f, e->l, tobj);
            if (po->is_table()) {
                Table *t = static_cast<Table *>(po->obj());
                Value keep = Value::object(Type::Table, t);  // the key expression may drop the last reference
                eval(L, f, e->r, key);
                out = t->get(key);
                if (out.is_nil() && t->meta.is_table()) index_value(L, f, e, keep, key, out);   // __index
                return;
            }
            Value obj = *po;
            eval(L, f, e->r, key);
            index_value(L, f, e, obj, key, out);
            return;
        }
        case EK::Call: {
            // numeric fast path: math.* style C functions on number arguments
            const size_t na = e->list.size();
            if (na == 1 || na == 2) {
                Value tf;
                const Value *pf = operand(L, f, e->l, tf);
                if (pf->is_function()) {
                    const Function *fn = static_cast<const Function *>(pf->obj());
                    const Expr *a0 = e->list[0];
                    if (na == 1 && fn->fast1 && a0->k != EK::Call && a0->k != EK::Method && a0->k != EK::Vararg) {
                        double (*fp)(double) = fn->fast1;
                        Value ta;
                        const Value *pa = operand(L, f, a0, ta);
                        if (pa->is_number()) {
                            out = Value(fp(pa->num()));
                            return;
                        }
                        // not a number: take the general route with the value already computed
                        Value fv = *pf, av = *pa;
                        ValueList rets;
                        call_value(L, fv, &av, 1, rets, &f, e->line);
                        if (rets.size() > 0) out = rets[0]; else out = Value();
                        return;
                    }
                    const Expr *a1 = na == 2 ? e->list[1] : nullptr;
                    if (na == 2 && fn->fast2 && a1->k != EK::Call && a1->k != EK::Method && a1->k != EK::Vararg) {
                        double (*fp)(double, double) = fn->fast2;
                        Value fv = *pf;  // arguments may reassign the callee variable
                        Value ta, tb;
                        const Value *pa = operand(L, f, a0, ta);
                        Value av = *pa;
                        const Value *pb = operand(L, f, a1, tb);
                        if (av.is_number() && pb->is_number()) {
                            out = Value(fp(av.num(), pb->num()));
                            return;
                        }
                        Value args2[2] = {av, *pb};
                        ValueList rets;
                        call_value(L, fv, args2, 2, rets, &f, e->line);
                        if (rets.size() > 0) out = rets[0]; else out = Value();
                        return;
                    }
                }
            }
            ValueList rets;
            eval_call(L, f, e, rets);
            if (rets.size() > 0) out = rets[0]; else out = Value();
            return;
        }
        case EK::Method: {
            ValueList rets;
            eval_call(L, f, e, rets);
            if (rets.size() > 0) out = rets[0]; else out = Value();
            return;
        }
        case EK::Function: make_closure(L, f, e, out); return;
        case EK::Add: case EK::Sub: case EK::Mul: case EK::Div: case EK::Mod: case EK::Pow: {
            Value ta, tb;
            // left operand first, then right (evaluation order is observable through calls)
            const Value *pa = operand(L, f, e->l, ta);
            double x0 = 0;
            const bool a_num = pa->is_number();
            if (a_num) x0 = pa->num();  // read now: the right operand may assign the variable
            else if (pa != &ta) ta = *pa;
            const Value *pb = operand(L, f, e->r, tb);
            if (a_num && pb->is_number()) {
                out = Value(do_arith(e->k, x0, pb->num()));
                return;
            }
            Value a = a_num ? Value(x0) : ta, b = *pb;
            double x, y;
            if (a.is_table() || b.is_table()) {
                static const char *const kEvent[] = {"__add", "__sub", "__mul", "__div", "__mod", "__pow"};
                const char *ev = e->k == EK::Add ? kEvent[0] : e->k == EK::Sub ? kEvent[1] : e->k == EK::Mul ? kEvent[2]
                                 : e->k == EK::Div ? kEvent[3] : e->k == EK::Mod ? kEvent[4] : kEvent[5];
                if (call_binary(L, f, e->line, binary_handler(a, b, ev), a, b, out)) return;
            }
            if (!coerce_num(a, &x)) {
                std::string what = describe(e->l, f);
                rt_error(f, e->line, std::string("attempt to ") + arith_name(e->k) + " a " + State::type_name(a) + " value" +
                                         (what.empty() ? "" : " (" + what + ")"));
            }
            if (!coerce_num(b, &y)) {
                std::string what = describe(e->r, f);
                rt_error(f, e->line, std::string("attempt to ") + arith_name(e->k) + " a " + State::type_name(b) + " value" +
                                         (what.empty() ? "" : " (" + what + ")"));
            }
            out = Value(do_arith(e->k, x, y));
            return;
        }
        case EK::Concat: {
            Value a, b;
            eval(L, f, e->l, a);
            eval(L, f, e->r, b);
            if ((a.is_table() || b.is_table()) && call_binary(L, f, e->line, binary_handler(a, b, "__concat"), a, b, out)) return;
            if (!(a.is_string() || a.is_number()))
                rt_error(f, e->line, std::string("attempt to concatenate a ") + State::type_name(a) + " value");
            if (!(b.is_string() || b.is_number()))
                rt_error(f, e->line, std::string("attempt to concatenate a ") + State::type_name(b) + " value");
            std::string s = a.is_string() ? a.str() : number_to_string(a.num());
            s += b.is_string() ? b.str() : number_to_string(b.num());
            out = L.new_string(s);
            return;
        }
        case EK::Eq: case EK::Ne: {
            Value a, b;
            eval(L, f, e->l, a);
            eval(L, f, e->r, b);
            bool eq = a.raw_equals(b);
            if (!eq && a.is_table() && b.is_table()) {   // __eq: only for two tables (5.2: handlers of either)
                Value r;
                if (call_binary(L, f, e->line, binary_handler(a, b, "__eq"), a, b, r)) eq = r.truthy();
            }
            out = Value::boolean(e->k == EK::Eq ? eq : !eq);
            return;
        }
        case EK::Lt: case EK::Le: case EK::Gt: case EK::Ge: {
            Value a, b;
            {
                Value ta, tb;
                const Value *pa = operand(L, f, e->l, ta);
                if (pa->is_number()) {
                    const double x0 = pa->num();
                    const Value *pb = operand(L, f, e->r, tb);
                    if (pb->is_number()) {
                        const double y0 = pb->num();
                        bool r;
                        switch (e->k) {
                            case EK::Lt: r = x0 < y0; break;
                            case EK::Le: r = x0 <= y0; break;
                            case EK::Gt: r = y0 < x0; break;
                            default: r = y0 <= x0; break;
                        }
                        out = Value::boolean(r);
                        return;
                    }
                    a = Value(x0);
                    b = *pb;
                } else {
                    a = *pa;
                    eval(L, f, e->r, b);
                }
            }
            bool r;
            switch (e->k) {
                case EK::Lt: r = less_than(L, f, e->line, a, b); break;
                case EK::Le: r = less_equal(L, f, e->line, a, b); break;
                case EK::Gt: r = less_than(L, f, e->line, b, a); break;
                default: r = less_equal(L, f, e->line, b, a); break;
            }
            out = Value::boolean(r);
            return;
        }
        case EK::And: {
            eval(L, f, e->l, out);
            if (out.truthy()) eval(L, f, e->r, out);
            return;
        }
        case EK::Or: {
            eval(L, f, e->l, out);
            if (!out.truthy()) eval(L, f, e->r, out);
            return;
        }
        case EK::Not: {
            Value a;
            eval(L, f, e->l, a);
            out = Value::boolean(!a.truthy());
            return;
        }
        case EK::Neg: {
            Value a;
            eval(L, f, e->l, a);
            double x;
            if (a.is_table() && call_binary(L, f, e->line, metamethod(a, "__unm"), a, a, out)) return;
            if (!coerce_num(a, &x))
                rt_error(f, e->line, std::string("attempt to perform arithmetic on a ") + State::type_name(a) + " value");
            out = Value(-x);
            return;
        }
        case EK::Len: {
            Value a;
            eval(L, f, e->l, a);
            if (a.is_string()) out = Value(static_cast<double>(a.str().size()));
            else if (a.is_table() && call_binary(L, f, e->line, metamethod(a, "__len"), a, a, out)) return;
            else if (a.is_table()) out = Value(static_cast<double>(static_cast<Table *>(a.obj())->length()));
            else rt_error(f, e->line, std::string("attempt to get length of a ") + State::type_name(a) + " value");
            return;
        }
        case EK::Table: {
            Table *t = L.alloc_table();
            Value tv = Value::object(Type::Table, t);
            for (size_t i = 0; i < e->keys.size(); ++i) {
                Value k, v;
                eval(L, f, e->keys[i], k);
                eval(L, f, e->vals[i], v);
                try {
                    t->set(k, v);
                } catch (LuaError &err) {
                    rt_error(f, e->line, err.what());
                }
            }
            if (!e->list.empty()) {
                const size_t n = e->list.size();
                const Expr *last = e->list[n - 1];
                const bool multi = last->k == EK::Call || last->k == EK::Method || last->k == EK::Vararg;
                t->arr.reserve(n);
                int64_t at = 1;
                for (size_t i = 0; i + (multi ? 1 : 0) < n; ++i) {
                    Value v;
                    eval(L, f, e->list[i], v);
                    t->set_int(at++, v);
                }
                if (multi) {
                    ValueList vals;
                    eval_multi(L, f, last, vals);
                    for (int i = 0; i < vals.size(); ++i) t->set_int(at++, vals[i]);
                }
            }
            out = tv;
            return;
        }
    }
}

void assign_to(State &L, Frame &f, const Expr *target, const Value &v) {
    switch (target->k) {
        case EK::Local: slot_ref(f, target->var) = v; return;
        case EK::Upval: f.fn->upvals[static_cast<size_t>(target->id)]->v = v; return;
        case EK::Global: L.global_slot(target->id) = v; return;
        case EK::Index: {
            Value obj, key;
            eval(L, f, target->l, obj);
            eval(L, f, target->r, key);
            if (!obj.is_table()) {
                std::string what = describe(target->l, f);
                rt_error(f, target->line, "attempt to index " + (what.empty() ? std::string("a ") : what + " (a ") +
                                              State::type_name(obj) + " value" + (what.empty() ? "" : ")"));
            }
            // __newindex: consulted only when the key is absent from the table itself
            for (int hops = 0; hops < 100 && obj.is_table(); ++hops) {
                Table *t = static_cast<Table *>(obj.obj());
                if (!t->meta.is_table() || !t->get(key).is_nil()) break;
                Value h = metamethod(obj, "__newindex");
                if (h.is_nil()) break;
                if (h.is_function()) {
                    const Value args[3] = {obj, key, v};
                    ValueList rets;
                    call_value(L, h, args, 3, rets, &f, target->line);
                    return;
                }
                obj = h;   // a table: the assignment goes there (and may meet its metatable)
            }
            if (!obj.is_table()) rt_error(f, target->line, std::string("attempt to index a ") + State::type_name(obj) + " value");
            try {
                static_cast<Table *>(obj.obj())->set(key, v);
            } catch (LuaError &err) {
                rt_error(f, target->line, err.what());
            }
            return;
        }
        default: rt_error(f, target->line, "cannot assign");
    }
}

Flow exec_stmt(State &L, Frame &f, const Stmt *s, ValueList &ret) {
    switch (s->k) {
        case SK::Local: {
            if (s->vars.size() == 1 && s->exprs.size() == 1) {
                Value v;
                eval(L, f, s->exprs[0], v);
                declare_slot(L, f, s->vars[0], v);
                return Flow::Normal;
            }
            ValueList vals;
            eval_args(L, f, s->exprs, vals);
            for (size_t i = 0; i < s->vars.size(); ++i)
                declare_slot(L, f, s->vars[i], static_cast<int>(i) < vals.size() ? vals[static_cast<int>(i)] : Value());
            return Flow::Normal;
        }
        case SK::Assign: {
            if (s->targets.size() == 1 && s->exprs.size() == 1) {
                Value v;
                eval(L, f, s->exprs[0], v);
                assign_to(L, f, s->targets[0], v);
                return Flow::Normal;
            }
            ValueList vals;
            eval_args(L, f, s->exprs, vals);
            // Lua leaves the order of multiple assignment undefined; the real VM
            // stores right-to-left, which we mirror.
            for (size_t i = s->targets.size(); i-- > 0;)
                assign_to(L, f, s->targets[i], static_cast<int>(i) < vals.size() ? vals[static_cast<int>(i)] : Value());
            return Flow::Normal;
        }
        case SK::Call: {
            ValueList rets;
            eval_call(L, f, s->e, rets);
            return Flow::Normal;
        }
        case SK::Goto:
            f.pending_goto = s;
            return Flow::Goto;
        case SK::Label: return Flow::Normal;
        case SK::Do: return exec_block(L, f, s->body, ret);
        case SK::While: {
            for (;;) {
                Value c;
                eval(L, f, s->e, c);
                if (!c.truthy()) break;
                Flow fl = exec_block(L, f, s->body, ret);
                if (fl == Flow::Break) break;
                if (fl == Flow::Return || fl == Flow::Goto) return fl;
            }
            return Flow::Normal;
        }
        case SK::Repeat: {
            for (;;) {
                Flow fl = exec_block(L, f, s->body, ret);
                if (fl == Flow::Break) break;
                if (fl == Flow::Return || fl == Flow::Goto) return fl;
                Value c;
                eval(L, f, s->e, c);
                if (c.truthy()) break;
            }
            return Flow::Normal;
        }
        case SK::If: {
            for (size_t i = 0; i < s->conds.size(); ++i) {
                Value c;
                eval(L, f, s->conds[i], c);
                if (c.truthy()) return exec_block(L, f, s->blocks[i], ret);
            }
            if (s->blocks.size() > s->conds.size()) return exec_block(L, f, s->blocks.back(), ret);
            return Flow::Normal;
        }
        case SK::NumFor: {
            Value a, b, c;
            eval(L, f, s->exprs[0], a);
            eval(L, f, s->exprs[1], b);
            double start, limit, step = 1;
            if (!coerce_num(a, &start)) rt_error(f, s->line, "'for' initial value must be a number");
            if (!coerce_num(b, &limit)) rt_error(f, s->line, "'for' limit must be a number");
            if (s->exprs.size() > 2) {
                eval(L, f, s->exprs[2], c);
                if (!coerce_num(c, &step)) rt_error(f, s->line, "'for' step must be a number");
            }
            // Lua 5.2 lvm.c OP_FORPREP/OP_FORLOOP: idx = init - step; loop { idx += step; test }
            double idx = start - step;
            for (;;) {
                idx = idx + step;
                if (step > 0 ? !(idx <= limit) : !(limit <= idx)) break;
                declare_slot(L, f, s->vars[0], Value(idx));
                Flow fl = exec_block(L, f, s->body, ret);
                if (fl == Flow::Break) break;
                if (fl == Flow::Return || fl == Flow::Goto) return fl;
            }
            return Flow::Normal;
        }
        case SK::GenFor: {
            ValueList init;
            eval_args(L, f, s->exprs, init);
            Value fn = init.size() > 0 ? init[0] : Value();
            Value st = init.size() > 1 ? init[1] : Value();
            Value ctl = init.size() > 2 ? init[2] : Value();
            for (;;) {
                Value args[2] = {st, ctl};
                ValueList rets;
                if (!fn.is_function())
                    rt_error(f, s->line, std::string("attempt to call a ") + State::type_name(fn) + " value");
                call_value(L, fn, args, 2, rets, &f, s->line);
                if (rets.size() == 0 || rets[0].is_nil()) break;
                ctl = rets[0];
                for (size_t i = 0; i < s->vars.size(); ++i)
                    declare_slot(L, f, s->vars[i], static_cast<int>(i) < rets.size() ? rets[static_cast<int>(i)] : Value());
                Flow fl = exec_block(L, f, s->body, ret);
                if (fl == Flow::Break) break;
                if (fl == Flow::Return || fl == Flow::Goto) return fl;
            }
            return Flow::Normal;
        }
        case SK::Return: {
            eval_args(L, f, s->exprs, ret);
            return Flow::Return;
        }
        case SK::Break: return Flow::Break;
        case SK::LocalFunction: {
            declare_slot(L, f, s->vars[0], Value());
            Value fn;
            eval(L, f, s->e, fn);
            slot_ref(f, s->vars[0]) = fn;
            return Flow::Normal;
        }
    }
    return Flow::Normal;
}

Flow exec_block(State &L, Frame &f, const Block *b, ValueList &ret) {
    for (size_t i = 0; i < b->stmts.size(); ++i) {
        Flow fl = exec_stmt(L, f, b->stmts[i], ret);
        if (fl == Flow::Goto) {
            // is the label in this block?  then execution continues behind it (backward jumps included)
            size_t at = b->stmts.size();
            for (size_t k = 0; k < b->stmts.size(); ++k)
                if (b->stmts[k]->k == SK::Label && b->stmts[k]->label == f.pending_goto->label) at = k;
            if (at == b->stmts.size()) return fl;   // not here: an enclosing block's
            f.pending_goto = nullptr;
            i = at;                                  // (the loop's ++i steps over the label itself)
            continue;
        }
        if (fl != Flow::Normal) return fl;
    }
    return Flow::Normal;
}

struct StackGuard {
    State &L;
    int n;
    StackGuard(State &l, int c) : L(l), n(c) {}
    ~StackGuard() {
        L.stack_free(n);
        --L.depth;
    }
};

void call_value(State &L, const Value &fnv, const Value *args, int nargs, ValueList &out, const Frame *caller, int line) {
    if (!fnv.is_function()) {
        std::string msg = std::string("attempt to call a ") + State::type_name(fnv) + " value";
        if (caller) rt_error(*caller, line, msg);
        throw LuaError(msg);
    }
    Function *fn = static_cast<Function *>(fnv.obj());
    if (fn->cfn) {
        try {
            fn->cfn(L, args, nargs, out, fn->ud);
        } catch (LuaError &e) {
            // errors raised by the C function itself get the calling line, as
            // luaL_error/luaL_argerror do
            if (caller && !e.positioned) rt_error(*caller, line, e.what());
            throw;
        }
        return;
    }
    const Proto *p = fn->proto;
    if (L.depth >= 195) {
        if (caller) rt_error(*caller, line, "stack overflow");
        throw LuaError("stack overflow");
    }
    ++L.depth;
    Value keep = fnv;  // the callee may overwrite the variable holding itself
    if (L.gc_pending()) L.collect_cycles();
    Frame f;
    f.base = L.stack_alloc(p->nslots);
    StackGuard guard(L, p->nslots);
    f.fn = fn;
    f.chunk = fn->chunk.get();
    for (int i = 0; i < p->nparams; ++i)
        declare_slot(L, f, p->params[static_cast<size_t>(i)], i < nargs ? args[i] : Value());
    if (p->is_vararg && nargs > p->nparams) {
        f.varargs = args + p->nparams;
        f.nvarargs = nargs - p->nparams;
    } else {
        f.varargs = nullptr;
        f.nvarargs = 0;
    }
    if (exec_block(L, f, p->body, out) == Flow::Goto) rt_error(f, f.pending_goto->line, "no visible label '" + f.pending_goto->label + "' for goto");
}

}  // namespace

// ---------------------------------------------------------------------------
// Standard library
// ---------------------------------------------------------------------------

namespace {

[[noreturn]] void arg_error(int i, const char *fname, const std::string &msg) {
    std::ostringstream o;
    o << "bad argument #" << i << " to '" << fname << "' (" << msg << ")";
    throw LuaError(o.str());
}

double check_number(const Value *args, int nargs, int i, const char *fname) {
    double d;
    if (i > nargs) arg_error(i, fname, "number expected, got no value");
    if (!args[i - 1].to_number(&d)) arg_error(i, fname, std::string("number expected, got ") + State::type_name(args[i - 1]));
    return d;
}

double opt_number(const Value *args, int nargs, int i, const char *fname, double def) {
    if (i > nargs || args[i - 1].is_nil()) return def;
    return check_number(args, nargs, i, fname);
}

Table *check_table(const Value *args, int nargs, int i, const char *fname) {
    if (i > nargs) arg_error(i, fname, "table expected, got no value");
    if (!args[i - 1].is_table()) arg_error(i, fname, std::string("table expected, got ") + State::type_name(args[i - 1]));
    return static_cast<Table *>(args[i - 1].obj());
}

std::string check_string(const Value *args, int nargs, int i, const char *fname) {
    if (i > nargs) arg_error(i, fname, "string expected, got no value");
    if (args[i - 1].is_string()) return args[i - 1].str();
    if (args[i - 1].is_number()) return number_to_string(args[i - 1].num());
    arg_error(i, fname, std::string("string expected, got ") + State::type_name(args[i - 1]));
}

#define MATH1(NAME, EXPR)                                                                  \
    void m_##NAME(State &, const Value *a, int n, ValueList &out, void *) {                \
        double x = check_number(a, n, 1, #NAME);                                           \
        out.push_back(Value(EXPR));                                                        \
    }

MATH1(abs, std::fabs(x))
MATH1(acos, std::acos(x))
MATH1(asin, std::asin(x))
MATH1(atan, std::atan(x))
MATH1(ceil, std::ceil(x))
MATH1(cos, std::cos(x))
MATH1(cosh, std::cosh(x))
MATH1(deg, x / (M_PI / 180.0))
MATH1(exp, std::exp(x))
MATH1(floor, std::floor(x))
MATH1(log10, std::log10(x))
MATH1(rad, x * (M_PI / 180.0))
MATH1(sin, std::sin(x))
MATH1(sinh, std::sinh(x))
MATH1(sqrt, std::sqrt(x))
MATH1(tan, std::tan(x))
MATH1(tanh, std::tanh(x))
#undef MATH1

void m_atan2(State &, const Value *a, int n, ValueList &out, void *) {
    out.push_back(Value(std::atan2(check_number(a, n, 1, "atan2"), check_number(a, n, 2, "atan2"))));
}
void m_pow(State &, const Value *a, int n, ValueList &out, void *) {
    out.push_back(Value(std::pow(check_number(a, n, 1, "pow"), check_number(a, n, 2, "pow"))));
}
void m_fmod(State &, const Value *a, int n, ValueList &out, void *) {
    out.push_back(Value(std::fmod(check_number(a, n, 1, "fmod"), check_number(a, n, 2, "fmod"))));
}
void m_ldexp(State &, const Value *a, int n, ValueList &out, void *) {
    out.push_back(Value(std::ldexp(check_number(a, n, 1, "ldexp"), static_cast<int>(check_number(a, n, 2, "ldexp")))));
}
void m_frexp(State &, const Value *a, int n, ValueList &out, void *) {
    int e;
    double m = std::frexp(check_number(a, n, 1, "frexp"), &e);
    out.push_back(Value(m));
    out.push_back(Value(static_cast<double>(e)));
}
void m_log(State &, const Value *a, int n, ValueList &out, void *) {
    // Lua 5.2 lmathlib.c math_log: natural log, or log(x)/log(base) (log10 for base 10)
    double x = check_number(a, n, 1, "log");
    double r;
    if (n < 2 || a[1].is_nil()) {
        r = std::log(x);
    } else {
        double base = check_number(a, n, 2, "log");
        if (base == 10.0) r = std::log10(x); else r = std::log(x) / std::log(base);
    }
    out.push_back(Value(r));
}
void m_modf(State &, const Value *a, int n, ValueList &out, void *) {
    double ip;
    double fp = std::modf(check_number(a, n, 1, "modf"), &ip);
    out.push_back(Value(ip));
    out.push_back(Value(fp));
}
void m_max(State &, const Value *a, int n, ValueList &out, void *) {
    double m = check_number(a, n, 1, "max");
    for (int i = 2; i <= n; ++i) {
        double d = check_number(a, n, i, "max");
        if (d > m) m = d;
    }
    out.push_back(Value(m));
}
void m_min(State &, const Value *a, int n, ValueList &out, void *) {
    double m = check_number(a, n, 1, "min");
    for (int i = 2; i <= n; ++i) {
        double d = check_number(a, n, i, "min");
        if (d < m) m = d;
    }
    out.push_back(Value(m));
}

// deterministic xorshift64* so that scripts using math.random are reproducible
struct Rng {
    uint64_t s = 0x9E3779B97F4A7C15ull;
    double next() {
        s ^= s >> 12;
        s ^= s << 25;
        s ^= s >> 27;
        uint64_t r = s * 0x2545F4914F6CDD1Dull;
        return static_cast<double>(r >> 11) / 9007199254740992.0;
    }
};
void m_random(State &, const Value *a, int n, ValueList &out, void *ud) {
    Rng *rng = static_cast<Rng *>(ud);
    double r = rng->next();
    if (n == 0) {
        out.push_back(Value(r));
    } else if (n == 1) {
        double u = check_number(a, n, 1, "random");
        if (!(1.0 <= u)) arg_error(1, "random", "interval is empty");
        out.push_back(Value(std::floor(r * u) + 1.0));
    } else {
        double l = check_number(a, n, 1, "random"), u = check_number(a, n, 2, "random");
        if (!(l <= u)) arg_error(2, "random", "interval is empty");
        out.push_back(Value(std::floor(r * (u - l + 1)) + l));
    }
}
void m_randomseed(State &, const Value *a, int n, ValueList &, void *ud) {
    Rng *rng = static_cast<Rng *>(ud);
    double d = check_number(a, n, 1, "randomseed");
    uint64_t b;
    memcpy(&b, &d, sizeof b);
    rng->s = b ? b : 0x9E3779B97F4A7C15ull;
}

void b_tostring(State &L, const Value *a, int n, ValueList &out, void *);
void b_print(State &L, const Value *a, int n, ValueList &, void *) {
    std::string line;
    for (int i = 0; i < n; ++i) {
        if (i) line += "\t";
        if (a[i].is_table()) {   // may carry __tostring
            ValueList s;
            b_tostring(L, a + i, 1, s, nullptr);
            line += s[0].str();
        } else {
            line += State::tostring(a[i]);
        }
    }
    line += "\n";
    L.emit_print(line);
}
void b_type(State &L, const Value *a, int n, ValueList &out, void *) {
    if (n < 1) arg_error(1, "type", "value expected");
    out.push_back(L.new_string(State::type_name(a[0])));
}
void b_tostring(State &L, const Value *a, int n, ValueList &out, void *) {
    if (n < 1) arg_error(1, "tostring", "value expected");
    Value h = metamethod(a[0], "__tostring");
    if (!h.is_nil()) {
        ValueList rets;
        L.call(h, a, 1, rets);
        if (rets.size() < 1 || !(rets[0].is_string() || rets[0].is_number())) throw LuaError("'__tostring' must return a string");
        out.push_back(rets[0].is_string() ? rets[0] : L.new_string(number_to_string(rets[0].num())));
        return;
    }
    out.push_back(L.new_string(State::tostring(a[0])));
}
void b_setmetatable(State &, const Value *a, int n, ValueList &out, void *) {
    Table *t = check_table(a, n, 1, "setmetatable");
    if (n < 2 || !(a[1].is_nil() || a[1].is_table())) arg_error(2, "setmetatable", "nil or table expected");
    if (!metamethod(a[0], "__metatable").is_nil()) throw LuaError("cannot change a protected metatable");
    t->meta = a[1];
    out.push_back(a[0]);
}
void b_getmetatable(State &L, const Value *a, int n, ValueList &out, void *) {
    if (n < 1) arg_error(1, "getmetatable", "value expected");
    if (a[0].is_string()) {   // strings share one metatable whose __index is the string library
        Value mt = L.new_table();
        static_cast<Table *>(mt.obj())->set(L.new_string("__index"), L.get_global("string"));
        out.push_back(mt);
        return;
    }
    if (!a[0].is_table()) { out.push_back(Value()); return; }
    Value prot = metamethod(a[0], "__metatable");
    out.push_back(prot.is_nil() ? static_cast<Table *>(a[0].obj())->meta : prot);
}
void b_tonumber(State &, const Value *a, int n, ValueList &out, void *) {
    if (n < 1) arg_error(1, "tonumber", "value expected");
    if (n >= 2 && !a[1].is_nil()) {
        int base = static_cast<int>(check_number(a, n, 2, "tonumber"));
        std::string s = check_string(a, n, 1, "tonumber");
        if (base < 2 || base > 36) arg_error(2, "tonumber", "base out of range");
        const char *p = s.c_str();
        while (isspace(static_cast<unsigned char>(*p))) ++p;
        bool neg = false;
        if (*p == '-') { neg = true; ++p; } else if (*p == '+') ++p;
        if (!isalnum(static_cast<unsigned char>(*p))) { out.push_back(Value()); return; }
        double v = 0;
        for (; isalnum(static_cast<unsigned char>(*p)); ++p) {
            int d = isdigit(static_cast<unsigned char>(*p)) ? *p - '0' : toupper(*p) - 'A' + 10;
            if (d >= base) { out.push_back(Value()); return; }
            v = v * base + d;
        }
        while (isspace(static_cast<unsigned char>(*p))) ++p;
        if (*p) { out.push_back(Value()); return; }
        out.push_back(Value(neg ? -v : v));
        return;
    }
    double d;
    if (a[0].to_number(&d)) out.push_back(Value(d)); else out.push_back(Value());
}
void b_error(State &, const Value *a, int n, ValueList &, void *) {
    throw LuaError(n >= 1 ? State::tostring(a[0]) : std::string("nil"));
}
void b_assert(State &, const Value *a, int n, ValueList &out, void *) {
    if (n < 1 || !a[0].truthy()) throw LuaError(n >= 2 ? State::tostring(a[1]) : std::string("assertion failed!"));
    for (int i = 0; i < n; ++i) out.push_back(a[i]);
}
void b_select(State &, const Value *a, int n, ValueList &out, void *) {
    if (n >= 1 && a[0].is_string() && a[0].str() == "#") {
        out.push_back(Value(static_cast<double>(n - 1)));
        return;
    }
    double d = check_number(a, n, 1, "select");
    long i = static_cast<long>(d);
    if (i < 0) i = n + i; else if (i > n) i = n;
    if (i < 1) arg_error(1, "select", "index out of range");
    for (long k = i; k < n; ++k) out.push_back(a[k]);
}
void b_next(State &, const Value *a, int n, ValueList &out, void *) {
    Table *t = check_table(a, n, 1, "next");
    size_t pos = 0;
    if (n >= 2 && !a[1].is_nil()) {
        // locate the position after key a[1]
        int64_t idx;
        if (as_array_index(a[1], &idx) && static_cast<size_t>(idx) <= t->arr.size()) {
            pos = static_cast<size_t>(idx);
        } else {
            bool found = false;
            for (size_t j = 0; j < t->hash_order.size(); ++j)
                if (t->hash_order[j].raw_equals(a[1])) {
                    pos = t->arr.size() + j + 1;
                    found = true;
                    break;
                }
            if (!found) throw LuaError("invalid key to 'next'");
        }
    }
    Value k, v;
    if (t->next(&pos, &k, &v)) {
        out.push_back(k);
        out.push_back(v);
    } else {
        out.push_back(Value());
    }
}
void b_pairs(State &L, const Value *a, int n, ValueList &out, void *) {
    check_table(a, n, 1, "pairs");
    out.push_back(L.get_global("next"));
    out.push_back(a[0]);
    out.push_back(Value());
}
void ipairs_iter(State &, const Value *a, int n, ValueList &out, void *) {
    Table *t = check_table(a, n, 1, "ipairs");
    double i = check_number(a, n, 2, "ipairs") + 1;
    Value v = t->get_int(static_cast<int64_t>(i));
    if (v.is_nil()) {
        out.push_back(Value());
    } else {
        out.push_back(Value(i));
        out.push_back(v);
    }
}
void b_ipairs(State &L, const Value *a, int n, ValueList &out, void *) {
    check_table(a, n, 1, "ipairs");
    out.push_back(L.new_cfunction(ipairs_iter, nullptr, "ipairs_iter"));
    out.push_back(a[0]);
    out.push_back(Value(0.0));
}
void b_pcall(State &L, const Value *a, int n, ValueList &out, void *) {
    if (n < 1) arg_error(1, "pcall", "value expected");
    ValueList rets;
    int saved_depth = L.depth;
    try {
        L.call(a[0], a + 1, n - 1, rets);
        out.push_back(Value::boolean(true));
        for (int i = 0; i < rets.size(); ++i) out.push_back(rets[i]);
    } catch (LuaError &e) {
        L.depth = saved_depth;
        out.push_back(Value::boolean(false));
        out.push_back(L.new_string(e.what()));
    }
}
void b_xpcall(State &L, const Value *a, int n, ValueList &out, void *) {
    if (n < 2) arg_error(2, "xpcall", "value expected");
    ValueList rets;
    int saved_depth = L.depth;
    try {
        L.call(a[0], a + 2, n - 2, rets);
        out.push_back(Value::boolean(true));
        for (int i = 0; i < rets.size(); ++i) out.push_back(rets[i]);
    } catch (LuaError &e) {
        L.depth = saved_depth;
        Value msg = L.new_string(e.what());
        ValueList h;
        L.call(a[1], &msg, 1, h);   // the message handler's results replace the message
        out.push_back(Value::boolean(false));
        for (int i = 0; i < h.size(); ++i) out.push_back(h[i]);
    }
}
void b_collectgarbage(State &, const Value *a, int n, ValueList &out, void *) {
    // memory is reference counted (plus cycle collection at State teardown): nothing to do, report 0 KB
    const std::string opt = (n >= 1 && a[0].is_string()) ? a[0].str() : std::string("collect");
    if (opt == "isrunning") out.push_back(Value::boolean(true));
    else out.push_back(Value(0.0));
}
void b_load(State &L, const Value *a, int n, ValueList &out, void *) {
    // load(string [, chunkname]): compiled in the current State's global environment
    if (n < 1 || !a[0].is_string()) arg_error(1, "load", "string expected (functions as chunk readers are not supported)");
    const std::string name = (n >= 2 && a[1].is_string()) ? a[1].str() : "=(load)";
    try {
        out.push_back(L.load(a[0].str(), name));
    } catch (LuaError &e) {
        out.push_back(Value());
        out.push_back(L.new_string(e.what()));
    }
}
void os_time(State &, const Value *, int, ValueList &out, void *) { out.push_back(Value(static_cast<double>(time(nullptr)))); }
void os_clock(State &, const Value *, int, ValueList &out, void *) { out.push_back(Value(static_cast<double>(clock()) / CLOCKS_PER_SEC)); }
void os_getenv(State &L, const Value *a, int n, ValueList &out, void *) {
    const char *v = getenv(check_string(a, n, 1, "getenv").c_str());
    out.push_back(v ? L.new_string(v) : Value());
}
void os_date(State &L, const Value *a, int n, ValueList &out, void *) {
    std::string fmt = n >= 1 && a[0].is_string() ? a[0].str() : "%c";
    time_t t = n >= 2 ? static_cast<time_t>(check_number(a, n, 2, "date")) : time(nullptr);
    struct tm tmv;
    if (!fmt.empty() && fmt[0] == '!') { gmtime_r(&t, &tmv); fmt.erase(0, 1); } else localtime_r(&t, &tmv);
    char buf[256];
    out.push_back(L.new_string(std::string(buf, strftime(buf, sizeof buf, fmt.c_str(), &tmv))));
}
void io_write(State &L, const Value *a, int n, ValueList &, void *) {
    // goes where print() goes, without separators or the newline
    std::string s;
    for (int i = 1; i <= n; ++i) s += check_string(a, n, i, "write");
    L.emit_print(s);
}
void b_rawequal(State &, const Value *a, int n, ValueList &out, void *) {
    out.push_back(Value::boolean(n >= 2 && a[0].raw_equals(a[1])));
}
void b_rawlen(State &, const Value *a, int n, ValueList &out, void *) {
    if (n >= 1 && a[0].is_table()) out.push_back(Value(static_cast<double>(static_cast<Table *>(a[0].obj())->length())));
    else if (n >= 1 && a[0].is_string()) out.push_back(Value(static_cast<double>(a[0].str().size())));
    else arg_error(1, "rawlen", "table or string expected");
}
void b_rawget(State &, const Value *a, int n, ValueList &out, void *) {
    Table *t = check_table(a, n, 1, "rawget");
    out.push_back(n >= 2 ? t->get(a[1]) : Value());
}
void b_rawset(State &, const Value *a, int n, ValueList &out, void *) {
    Table *t = check_table(a, n, 1, "rawset");
    if (n < 3) arg_error(3, "rawset", "value expected");
    t->set(a[1], a[2]);
    out.push_back(a[0]);
}

void t_unpack(State &, const Value *a, int n, ValueList &out, void *) {
    Table *t = check_table(a, n, 1, "unpack");
    int64_t i = static_cast<int64_t>(opt_number(a, n, 2, "unpack", 1));
    int64_t e = (n >= 3 && !a[2].is_nil()) ? static_cast<int64_t>(check_number(a, n, 3, "unpack")) : t->length();
    if (e - i >= 1000000) throw LuaError("too many results to unpack");
    for (; i <= e; ++i) out.push_back(t->get_int(i));
}
void t_insert(State &, const Value *a, int n, ValueList &, void *) {
    Table *t = check_table(a, n, 1, "insert");
    int64_t e = t->length() + 1;
    if (n == 2) {
        t->set_int(e, a[1]);
    } else if (n == 3) {
        int64_t pos = static_cast<int64_t>(check_number(a, n, 2, "insert"));
        if (pos < 1 || pos > e) arg_error(2, "insert", "position out of bounds");
        for (int64_t i = e; i > pos; --i) t->set_int(i, t->get_int(i - 1));
        t->set_int(pos, a[2]);
    } else {
        throw LuaError("wrong number of arguments to 'insert'");
    }
}
void t_remove(State &, const Value *a, int n, ValueList &out, void *) {
    Table *t = check_table(a, n, 1, "remove");
    int64_t size = t->length();
    int64_t pos = static_cast<int64_t>(opt_number(a, n, 2, "remove", static_cast<double>(size)));
    if (n >= 2 && size + 1 != pos && (pos < 1 || pos > size + 1)) arg_error(2, "remove", "position out of bounds");
    if (size == 0 && n < 2) { out.push_back(Value()); return; }
    out.push_back(t->get_int(pos));
    for (; pos < size; ++pos) t->set_int(pos, t->get_int(pos + 1));
    t->set_int(pos, Value());
}
void t_concat(State &L, const Value *a, int n, ValueList &out, void *) {
    Table *t = check_table(a, n, 1, "concat");
    std::string sep = (n >= 2 && !a[1].is_nil()) ? check_string(a, n, 2, "concat") : std::string();
    int64_t i = static_cast<int64_t>(opt_number(a, n, 3, "concat", 1));
    int64_t e = (n >= 4 && !a[3].is_nil()) ? static_cast<int64_t>(check_number(a, n, 4, "concat")) : t->length();
    std::string s;
    for (; i <= e; ++i) {
        Value v = t->get_int(i);
        if (v.is_string()) s += v.str();
        else if (v.is_number()) s += number_to_string(v.num());
        else throw LuaError("invalid value (at index " + std::to_string(i) + ") in table for 'concat'");
        if (i != e) s += sep;
    }
    out.push_back(L.new_string(s));
}
// table.sort (ltablib.c sort): ascending by `<` on numbers / strings, or by the caller's order
// function.  A merge sort: an inconsistent order function gives some permutation, never a crash.
bool sort_less(State &L, const Value &cmp, const Value &x, const Value &y) {
    if (cmp.is_function()) {
        Value args[2] = {x, y};
        ValueList r;
        L.call(cmp, args, 2, r);
        return r.size() > 0 && r[0].truthy();
    }
    if (x.is_number() && y.is_number()) return x.num() < y.num();
    if (x.is_string() && y.is_string()) return x.str() < y.str();
    throw LuaError(std::string("attempt to compare ") + State::type_name(x) + " with " + State::type_name(y));
}
void t_sort(State &L, const Value *a, int n, ValueList &, void *) {
    Table *t = check_table(a, n, 1, "sort");
    Value cmp = n >= 2 ? a[1] : Value();
    if (!cmp.is_nil() && !cmp.is_function()) arg_error(2, "sort", "function expected");
    const int64_t size = t->length();
    std::vector<Value> v, tmp;
    for (int64_t i = 1; i <= size; ++i) v.push_back(t->get_int(i));
    tmp.resize(v.size());
    for (size_t width = 1; width < v.size(); width *= 2) {
        for (size_t lo = 0; lo < v.size(); lo += 2 * width) {
            const size_t mid = std::min(lo + width, v.size()), hi = std::min(lo + 2 * width, v.size());
            size_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = sort_less(L, cmp, v[j], v[i]) ? v[j++] : v[i++];
            while (i < mid) tmp[k++] = v[i++];
            while (j < hi) tmp[k++] = v[j++];
        }
        v.swap(tmp);
    }
    for (int64_t i = 1; i <= size; ++i) t->set_int(i, v[static_cast<size_t>(i - 1)]);
}
void t_pack(State &L, const Value *a, int n, ValueList &out, void *) {
    Value tv = L.new_table();
    Table *t = static_cast<Table *>(tv.obj());
    for (int i = 0; i < n; ++i) t->set_int(i + 1, a[i]);
    t->set(L.new_string("n"), Value(static_cast<double>(n)));
    out.push_back(tv);
}

void s_len(State &, const Value *a, int n, ValueList &out, void *) {
    out.push_back(Value(static_cast<double>(check_string(a, n, 1, "len").size())));
}
void s_sub(State &L, const Value *a, int n, ValueList &out, void *) {
    std::string s = check_string(a, n, 1, "sub");
    long l = static_cast<long>(s.size());
    long i = static_cast<long>(opt_number(a, n, 2, "sub", 1));
    long j = static_cast<long>(opt_number(a, n, 3, "sub", -1));
    if (i < 0) i = std::max(l + i + 1, 1L); else if (i == 0) i = 1;
    if (j < 0) j = l + j + 1; else if (j > l) j = l;
    out.push_back(L.new_string(i <= j ? s.substr(static_cast<size_t>(i - 1), static_cast<size_t>(j - i + 1)) : std::string()));
}
void s_rep(State &L, const Value *a, int n, ValueList &out, void *) {
    std::string s = check_string(a, n, 1, "rep");
    long c = static_cast<long>(check_number(a, n, 2, "rep"));
    std::string sep = (n >= 3 && !a[2].is_nil()) ? check_string(a, n, 3, "rep") : std::string();
    std::string r;
    if (c > 0 && (s.size() + sep.size()) * static_cast<size_t>(c) > (64u << 20)) throw LuaError("resulting string too large");
    for (long i = 0; i < c; ++i) {
        r += s;
        if (i + 1 < c) r += sep;
    }
    out.push_back(L.new_string(r));
}
void s_upper(State &L, const Value *a, int n, ValueList &out, void *) {
    std::string s = check_string(a, n, 1, "upper");
    for (auto &c : s) c = static_cast<char>(toupper(static_cast<unsigned char>(c)));
    out.push_back(L.new_string(s));
}
void s_lower(State &L, const Value *a, int n, ValueList &out, void *) {
    std::string s = check_string(a, n, 1, "lower");
    for (auto &c : s) c = static_cast<char>(tolower(static_cast<unsigned char>(c)));
    out.push_back(L.new_string(s));
}
void s_byte(State &, const Value *a, int n, ValueList &out, void *) {
    std::string s = check_string(a, n, 1, "byte");
    long l = static_cast<long>(s.size());
    long i = static_cast<long>(opt_number(a, n, 2, "byte", 1));
    long j = static_cast<long>(opt_number(a, n, 3, "byte", static_cast<double>(i)));
    if (i < 0) i = std::max(l + i + 1, 1L); else if (i == 0) i = 1;
    if (j < 0) j = l + j + 1; else if (j > l) j = l;
    for (long k = i; k <= j; ++k) out.push_back(Value(static_cast<double>(static_cast<unsigned char>(s[static_cast<size_t>(k - 1)]))));
}
void s_char(State &L, const Value *a, int n, ValueList &out, void *) {
    std::string s;
    for (int i = 1; i <= n; ++i) s.push_back(static_cast<char>(static_cast<int>(check_number(a, n, i, "char"))));
    out.push_back(L.new_string(s));
}
// ---- Lua patterns (reference manual 6.4.1): character classes %a %c %d %g %l %p %s %u %w %x and their
// complements, sets [...], the quantifiers * + - ?, anchors ^ $, captures ( ) and position captures (),
// back-references %1-%9, %bxy and the frontier %f[set].  A backtracking matcher over byte strings.
struct PatMatch {
    const char *src, *src_end, *pat_end;
    int level = 0, depth = 0;
    struct { const char *init; long len; } cap[32];
    static constexpr long kUnfinished = -1, kPosition = -2;

    [[noreturn]] static void bad(const char *m) { throw LuaError(m); }

    static bool in_class(int c, int cl) {
        bool r;
        switch (tolower(cl)) {
            case 'a': r = isalpha(c); break;
            case 'c': r = iscntrl(c); break;
            case 'd': r = isdigit(c); break;
            case 'g': r = isgraph(c); break;
            case 'l': r = islower(c); break;
            case 'p': r = ispunct(c); break;
            case 's': r = isspace(c); break;
            case 'u': r = isupper(c); break;
            case 'w': r = isalnum(c); break;
            case 'x': r = isxdigit(c); break;
            default: return cl == c;
        }
        return isupper(cl) ? !r : r;
    }
    // end of the single-character item that starts at p
    const char *item_end(const char *p) const {
        if (p == pat_end) bad("malformed pattern (ends with '%')");
        const unsigned char c = static_cast<unsigned char>(*p++);
        if (c == '%') {
            if (p == pat_end) bad("malformed pattern (ends with '%')");
            return p + 1;
        }
        if (c == '[') {
            if (p < pat_end && *p == '^') ++p;
            do {   // the first character may be ']'
                if (p == pat_end) bad("malformed pattern (missing ']')");
                if (*p++ == '%' && p < pat_end) ++p;
            } while (p == pat_end || *p != ']');
            return p + 1;
        }
        return p;
    }
    bool in_set(int c, const char *p, const char *set_end) const {   // p at '[', set_end at ']'
        bool neg = false;
        if (p[1] == '^') { neg = true; ++p; }
        while (++p < set_end) {
            if (*p == '%') {
                ++p;
                if (in_class(c, static_cast<unsigned char>(*p))) return !neg;
            } else if (p[1] == '-' && p + 2 < set_end) {
                if (static_cast<unsigned char>(p[0]) <= c && c <= static_cast<unsigned char>(p[2])) return !neg;
                p += 2;
            } else if (static_cast<unsigned char>(*p) == c) {
                return !neg;
            }
        }
        return neg;
    }
    bool single(const char *s, const char *p, const char *ep) const {
        if (s >= src_end) return false;
        const int c = static_cast<unsigned char>(*s);
        switch (*p) {
            case '.': return true;
            case '%': return in_class(c, static_cast<unsigned char>(p[1]));
            case '[': return in_set(c, p, ep - 1);
            default: return static_cast<unsigned char>(*p) == c;
        }
    }
    const char *match(const char *s, const char *p) {
        if (++depth > 200) bad("pattern too complex");
        const char *r = do_match(s, p);
        --depth;
        return r;
    }
    const char *do_match(const char *s, const char *p) {
        for (;;) {
            if (p == pat_end) return s;
            switch (*p) {
                case '(':
                    if (p + 1 < pat_end && p[1] == ')') return open_capture(s, p + 2, kPosition);
                    return open_capture(s, p + 1, kUnfinished);
                case ')': return close_capture(s, p + 1);
                case '$':
                    if (p + 1 == pat_end) return s == src_end ? s : nullptr;
                    break;
                case '%':
                    if (p + 1 < pat_end && p[1] == 'b') return balance(s, p + 2);
                    if (p + 1 < pat_end && p[1] == 'f') {
                        p += 2;
                        if (p == pat_end || *p != '[') bad("missing '[' after '%f' in pattern");
                        const char *ep = item_end(p);
                        const int prev = s == src ? 0 : static_cast<unsigned char>(s[-1]);
                        const int cur = s < src_end ? static_cast<unsigned char>(*s) : 0;
                        if (!in_set(prev, p, ep - 1) && in_set(cur, p, ep - 1)) { p = ep; continue; }
                        return nullptr;
                    }
                    if (p + 1 < pat_end && isdigit(static_cast<unsigned char>(p[1]))) {
                        const int l = p[1] - '1';
                        if (l < 0 || l >= level || cap[l].len == kUnfinished) bad("invalid capture index in pattern");
                        const size_t len = static_cast<size_t>(cap[l].len);
                        if (static_cast<size_t>(src_end - s) >= len && memcmp(cap[l].init, s, len) == 0) { s += len; p += 2; continue; }
                        return nullptr;
                    }
                    break;
                default: break;
            }
            const char *ep = item_end(p);
            const char q = ep < pat_end ? *ep : '\0';
            if (q == '?') {
                if (single(s, p, ep)) {
                    if (const char *r = match(s + 1, ep + 1)) return r;
                }
                p = ep + 1;
                continue;
            }
            if (q == '+') return single(s, p, ep) ? max_expand(s + 1, p, ep) : nullptr;
            if (q == '*') return max_expand(s, p, ep);
            if (q == '-') {
                for (;;) {
                    if (const char *r = match(s, ep + 1)) return r;
                    if (single(s, p, ep)) ++s; else return nullptr;
                }
            }
            if (!single(s, p, ep)) return nullptr;
            ++s;
            p = ep;
        }
    }
    const char *max_expand(const char *s, const char *p, const char *ep) {
        long i = 0;
        while (single(s + i, p, ep)) ++i;
        for (; i >= 0; --i)
            if (const char *r = match(s + i, ep + 1)) return r;
        return nullptr;
    }
    const char *open_capture(const char *s, const char *p, long what) {
        if (level >= 32) bad("too many captures");
        cap[level].init = s;
        cap[level].len = what;
        ++level;
        const char *r = match(s, p);
        if (!r) --level;
        return r;
    }
    const char *close_capture(const char *s, const char *p) {
        int l = level - 1;
        while (l >= 0 && cap[l].len != kUnfinished) --l;
        if (l < 0) bad("invalid pattern capture");
        cap[l].len = s - cap[l].init;
        const char *r = match(s, p);
        if (!r) cap[l].len = kUnfinished;
        return r;
    }
    const char *balance(const char *s, const char *p) {
        if (p + 1 >= pat_end) bad("malformed pattern (missing arguments to '%b')");
        if (s >= src_end || *s != *p) return nullptr;
        const int open = *p, close = p[1];
        int cont = 1;
        for (const char *t = s + 1; t < src_end; ++t) {
            if (*t == close) {
                if (--cont == 0) return match(t + 1, p + 2);
            } else if (*t == open) {
                ++cont;
            }
        }
        return nullptr;
    }
    void push_capture(State &L, int i, const char *s, const char *e, ValueList &out) const {
        if (i >= level) {
            if (i == 0) out.push_back(L.new_string(std::string(s, e)));   // no explicit captures: the whole match
            else bad("invalid capture index");
            return;
        }
        if (cap[i].len == kUnfinished) bad("unfinished capture");
        if (cap[i].len == kPosition) out.push_back(Value(static_cast<double>(cap[i].init - src + 1)));
        else out.push_back(L.new_string(std::string(cap[i].init, static_cast<size_t>(cap[i].len))));
    }
};

// string.find / string.match
void str_find_aux(State &L, const Value *a, int n, ValueList &out, bool find, const char *fname) {
    const std::string s = check_string(a, n, 1, fname), pat = check_string(a, n, 2, fname);
    long init = static_cast<long>(opt_number(a, n, 3, fname, 1));
    const long ls = static_cast<long>(s.size());
    if (init < 0) init = std::max(ls + init + 1, 1L); else if (init == 0) init = 1;
    if (init > ls + 1) { out.push_back(Value()); return; }
    const bool plain = find && n >= 4 && a[3].truthy();
    if (find && (plain || pat.find_first_of("^$*+?.([%-") == std::string::npos)) {
        const size_t at = s.find(pat, static_cast<size_t>(init - 1));
        if (at == std::string::npos) { out.push_back(Value()); return; }
        out.push_back(Value(static_cast<double>(at + 1)));
        out.push_back(Value(static_cast<double>(at + pat.size())));
        return;
    }
    PatMatch m;
    m.src = s.data();
    m.src_end = s.data() + s.size();
    const char *p = pat.data();
    m.pat_end = pat.data() + pat.size();
    const bool anchor = !pat.empty() && *p == '^';
    if (anchor) ++p;
    const char *s1 = s.data() + init - 1;
    do {
        m.level = 0;
        m.depth = 0;
        if (const char *e = m.match(s1, p)) {
            if (find) {
                out.push_back(Value(static_cast<double>(s1 - s.data() + 1)));
                out.push_back(Value(static_cast<double>(e - s.data())));
                for (int i = 0; i < m.level; ++i) m.push_capture(L, i, s1, e, out);
            } else {
                const int nc = m.level == 0 ? 1 : m.level;
                for (int i = 0; i < nc; ++i) m.push_capture(L, i, s1, e, out);
            }
            return;
        }
    } while (s1++ < m.src_end && !anchor);
    out.push_back(Value());
}
void s_find(State &L, const Value *a, int n, ValueList &out, void *) { str_find_aux(L, a, n, out, true, "find"); }
void s_match(State &L, const Value *a, int n, ValueList &out, void *) { str_find_aux(L, a, n, out, false, "match"); }
void s_reverse(State &L, const Value *a, int n, ValueList &out, void *) {
    std::string s = check_string(a, n, 1, "reverse");
    std::reverse(s.begin(), s.end());
    out.push_back(L.new_string(s));
}

// string.gmatch and string.gsub sit on string.find in Lua itself (run once per State)
const char *kStringPrelude = R"LUA(
function string.gmatch(s, p)
  local pos, done = 1, false
  return function()
    if done then return nil end
    local r = table.pack(string.find(s, p, pos))
    local st, e = r[1], r[2]
    if st == nil then done = true return nil end
    if e >= st then pos = e + 1 else pos = st + 1 end
    if pos > #s + 1 then done = true end
    if r.n > 2 then return table.unpack(r, 3, r.n) end
    return string.sub(s, st, e)
  end
end
function string.gsub(s, p, repl, max_n)
  local tr = type(repl)
  if tr ~= "string" and tr ~= "number" and tr ~= "table" and tr ~= "function" then
    error("bad argument #3 to 'gsub' (string/function/table expected)")
  end
  local out, pos, count, anchor = {}, 1, 0, string.sub(p, 1, 1) == "^"
  while max_n == nil or count < max_n do
    local r = table.pack(string.find(s, p, pos))
    local st, e = r[1], r[2]
    if st == nil then break end
    count = count + 1
    out[#out + 1] = string.sub(s, pos, st - 1)
    local whole = string.sub(s, st, e)
    local rv
    if tr == "string" or tr == "number" then
      local rs, acc, i = tostring(repl), {}, 1
      while i <= #rs do
        local c = string.sub(rs, i, i)
        if c ~= "%" then
          acc[#acc + 1] = c
        else
          i = i + 1
          local d = string.sub(rs, i, i)
          if d == "%" then
            acc[#acc + 1] = "%"
          elseif d == "0" then
            acc[#acc + 1] = whole
          elseif d >= "1" and d <= "9" and #d == 1 then
            local k = tonumber(d)
            if k == 1 and r.n == 2 then
              acc[#acc + 1] = whole
            elseif k + 2 > r.n then
              error("invalid capture index %" .. d .. " in replacement string")
            else
              acc[#acc + 1] = tostring(r[k + 2])
            end
          else
            error("invalid use of '%' in replacement string")
          end
        end
        i = i + 1
      end
      rv = table.concat(acc)
    elseif tr == "table" then
      if r.n > 2 then rv = repl[r[3]] else rv = repl[whole] end
    else
      if r.n > 2 then rv = repl(table.unpack(r, 3, r.n)) else rv = repl(whole) end
    end
    if rv == nil or rv == false then
      rv = whole
    elseif type(rv) ~= "string" and type(rv) ~= "number" then
      error("invalid replacement value (a " .. type(rv) .. ")")
    end
    out[#out + 1] = tostring(rv)
    if e >= st then
      pos = e + 1
    else
      out[#out + 1] = string.sub(s, st, st)
      pos = st + 1
    end
    if pos > #s + 1 or anchor then break end
  end
  out[#out + 1] = string.sub(s, pos)
  return table.concat(out), count
end
)LUA";

void s_format(State &L, const Value *a, int n, ValueList &out, void *) {
    std::string fmt = check_string(a, n, 1, "format");
    std::string r;
    int arg = 1;
    for (size_t i = 0; i < fmt.size(); ++i) {
        if (fmt[i] != '%') { r.push_back(fmt[i]); continue; }
        ++i;
        if (i >= fmt.size()) throw LuaError("invalid option '%' to 'format'");
        if (fmt[i] == '%') { r.push_back('%'); continue; }
        std::string spec = "%";
        while (i < fmt.size() && strchr("-+ #0", fmt[i])) spec.push_back(fmt[i++]);
        while (i < fmt.size() && isdigit(static_cast<unsigned char>(fmt[i]))) spec.push_back(fmt[i++]);
        if (i < fmt.size() && fmt[i] == '.') {
            spec.push_back(fmt[i++]);
            while (i < fmt.size() && isdigit(static_cast<unsigned char>(fmt[i]))) spec.push_back(fmt[i++]);
        }
        if (i >= fmt.size() || spec.size() > 20) throw LuaError("invalid format string to 'format'");
        char conv = fmt[i];
        ++arg;
        char buf[512];
        switch (conv) {
            case 'c': r.push_back(static_cast<char>(static_cast<int>(check_number(a, n, arg, "format")))); break;
            case 'd': case 'i': {
                spec += PRId64;
                snprintf(buf, sizeof buf, spec.c_str(), static_cast<int64_t>(check_number(a, n, arg, "format")));
                r += buf;
                break;
            }
            case 'o': case 'u': case 'x': case 'X': {
                spec += "ll";
                spec.push_back(conv);
                snprintf(buf, sizeof buf, spec.c_str(), static_cast<unsigned long long>(static_cast<int64_t>(check_number(a, n, arg, "format"))));
                r += buf;
                break;
            }
            case 'e': case 'E': case 'f': case 'g': case 'G': case 'a': case 'A': {
                spec.push_back(conv);
                snprintf(buf, sizeof buf, spec.c_str(), check_number(a, n, arg, "format"));
                r += buf;
                break;
            }
            case 's': {
                if (arg > n) arg_error(arg, "format", "no value");
                std::string s = State::tostring(a[arg - 1]);
                spec.push_back('s');
                if (spec == "%s") {
                    r += s;
                } else {
                    std::vector<char> big(s.size() + 128);
                    snprintf(big.data(), big.size(), spec.c_str(), s.c_str());
                    r += big.data();
                }
                break;
            }
            case 'q': {
                std::string s = check_string(a, n, arg, "format");
                r.push_back('"');
                for (char c : s) {
                    if (c == '"' || c == '\\') { r.push_back('\\'); r.push_back(c); }
                    else if (c == '\n') { r += "\\n"; }
                    else if (c == '\0') { r += "\\0"; }
                    else r.push_back(c);
                }
                r.push_back('"');
                break;
            }
            default: throw LuaError(std::string("invalid option '%") + conv + "' to 'format'");
        }
    }
    out.push_back(L.new_string(r));
}

}  // namespace

// ---------------------------------------------------------------------------
// State
// ---------------------------------------------------------------------------

static const int kStackValues = 1 << 16;

State::State() : State(std::make_shared<Universe>()) {}

State::State(std::shared_ptr<Universe> u) : uni_(std::move(u)) {
    stack_.resize(kStackValues);
    open_libs();
}

State::~State() {
    // break reference cycles (tables/closures referring to each other) so the
    // intrusive refcounts can reach zero: clear containers first.
    for (auto &g : globals_) g = Value();
    for (auto &s : stack_) s = Value();
    for (auto &k : kstr_) k = Value();
    collect_cycles();  // nothing is rooted any more: frees every remaining cycle
    // anything still linked is held by a Value outside this State (host code
    // keeping a Value past the State's lifetime); detach so its later release
    // does not touch freed memory.
    while (gc_head_) untrack(gc_head_);
}

Value *State::stack_alloc(int n) {
    if (stack_top_ + static_cast<size_t>(n) > stack_.size()) throw LuaError("stack overflow");
    Value *b = &stack_[stack_top_];
    stack_top_ += static_cast<size_t>(n);
    return b;
}

void State::stack_free(int n) {
    for (int i = 0; i < n; ++i) stack_[--stack_top_] = Value();
}

const Value &State::kstr(int id) {
    if (id >= static_cast<int>(kstr_.size())) kstr_.resize(static_cast<size_t>(id) + 32);
    Value &v = kstr_[static_cast<size_t>(id)];
    if (v.is_nil()) {
        std::string s;
        {
            std::lock_guard<std::mutex> g(uni_->mu);
            s = uni_->kstrs[static_cast<size_t>(id)];
        }
        v = new_string(s);
    }
    return v;
}

void State::track(Object *o) {
    o->owner = this;
    o->gc_prev = nullptr;
    o->gc_next = gc_head_;
    if (gc_head_) gc_head_->gc_prev = o;
    gc_head_ = o;
    if (++gc_count_ > gc_threshold_) gc_pending_ = true;
}

void State::untrack(Object *o) {
    if (o->gc_prev) o->gc_prev->gc_next = o->gc_next; else gc_head_ = o->gc_next;
    if (o->gc_next) o->gc_next->gc_prev = o->gc_prev;
    o->gc_prev = o->gc_next = nullptr;
    o->owner = nullptr;
    --gc_count_;
}

Table *State::alloc_table() {
    Table *t = new Table();
    track(t);
    return t;
}
Box *State::alloc_box() {
    Box *b = new Box();
    track(b);
    return b;
}
Function *State::alloc_function() {
    Function *f = new Function();
    track(f);
    return f;
}
Str *State::alloc_str(std::string s) {
    Str *o = new Str(std::move(s));
    track(o);
    return o;
}

namespace {
template <typename F>
void for_each_child(Object *o, F &&fn) {
    if (Table *t = dynamic_cast<Table *>(o)) {
        for (const Value &v : t->arr)
            if (v.type() >= Type::String) fn(v.obj());
        for (auto &kv : t->hash) {
            if (kv.first.type() >= Type::String) fn(kv.first.obj());
            if (kv.second.type() >= Type::String) fn(kv.second.obj());
        }
        for (const Value &v : t->hash_order)
            if (v.type() >= Type::String) fn(v.obj());
        if (t->meta.type() >= Type::String) fn(t->meta.obj());
    } else if (Function *f = dynamic_cast<Function *>(o)) {
        for (Box *b : f->upvals)
            if (b) fn(b);
    } else if (Box *b = dynamic_cast<Box *>(o)) {
        if (b->v.type() >= Type::String) fn(b->v.obj());
    }
}
}  // namespace

void State::collect_cycles() {
    gc_pending_ = false;
    // CPython-style trial deletion: an object whose refcount is not fully
    // explained by references from other heap objects is externally rooted.
    for (Object *o = gc_head_; o; o = o->gc_next) {
        o->gc_refs = o->rc;
        o->gc_mark = false;
    }
    for (Object *o = gc_head_; o; o = o->gc_next)
        for_each_child(o, [](Object *c) { --c->gc_refs; });
    std::vector<Object *> work;
    for (Object *o = gc_head_; o; o = o->gc_next)
        if (o->gc_refs > 0) {
            o->gc_mark = true;
            work.push_back(o);
        }
    while (!work.empty()) {
        Object *o = work.back();
        work.pop_back();
        for_each_child(o, [&work](Object *c) {
            if (!c->gc_mark) {
                c->gc_mark = true;
                work.push_back(c);
            }
        });
    }
    std::vector<Object *> garbage;
    for (Object *o = gc_head_; o; o = o->gc_next)
        if (!o->gc_mark) {
            ++o->rc;  // pin while we tear the cycle apart
            garbage.push_back(o);
        }
    for (Object *o : garbage) {
        if (Table *t = dynamic_cast<Table *>(o)) {
            t->arr.clear();
            t->hash.clear();
            t->hash_order.clear();
            t->meta = Value();
        } else if (Function *f = dynamic_cast<Function *>(o)) {
            std::vector<Box *> ups;
            ups.swap(f->upvals);
            for (Box *b : ups)
                if (b && --b->rc == 0) delete b;
        } else if (Box *b = dynamic_cast<Box *>(o)) {
            b->v = Value();
        }
    }
    for (Object *o : garbage)
        if (--o->rc == 0) delete o;
    gc_threshold_ = gc_count_ * 2 > (1u << 16) ? gc_count_ * 2 : (1u << 16);
}

Value State::new_string(const std::string &s) { return Value::object(Type::String, alloc_str(s)); }
Value State::new_table() { return Value::object(Type::Table, alloc_table()); }
Value State::new_cfunction(CFunction f, void *ud, const char *name) {
    Function *fn = alloc_function();
    fn->cfn = f;
    fn->ud = ud;
    fn->cname = name;
    return Value::object(Type::Function, fn);
}

const char *State::type_name(const Value &v) {
    switch (v.type()) {
        case Type::Nil: return "nil";
        case Type::Boolean: return "boolean";
        case Type::Number: return "number";
        case Type::String: return "string";
        case Type::Table: return "table";
        case Type::Function: return "function";
        default: return "userdata";
    }
}

std::string State::tostring(const Value &v) {
    switch (v.type()) {
        case Type::Nil: return "nil";
        case Type::Boolean: return v.boolean_value() ? "true" : "false";
        case Type::Number: return number_to_string(v.num());
        case Type::String: return v.str();
        default: {
            char b[64];
            snprintf(b, sizeof b, "%s: %p", type_name(v), static_cast<void *>(v.obj()));
            return b;
        }
    }
}

void State::emit_print(const std::string &s) {
    if (print_sink_) print_sink_(s.c_str(), print_ud_);
    else fputs(s.c_str(), stdout);
}

Value State::get_global(const std::string &name) { return global_slot(uni_->global_id(name)); }
void State::set_global(const std::string &name, const Value &v) { global_slot(uni_->global_id(name)) = v; }
void State::register_function(const std::string &name, CFunction f, void *ud) {
    set_global(name, new_cfunction(f, ud, nullptr));
}

Value State::load(const std::string &src, const std::string &chunkname) {
    Parser p(src, chunkname, uni_.get());
    std::shared_ptr<Chunk> chunk = p.parse_chunk();
    Function *fn = alloc_function();
    fn->proto = chunk->main;
    fn->chunk = chunk;
    return Value::object(Type::Function, fn);
}

Value State::load_file(const std::string &path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw LuaError("cannot open " + path);
    std::stringstream ss;
    ss << in.rdbuf();
    // chunk name: file name without directories, like luaL_loadfile's "@path" shortened
    std::string name = path;
    size_t slash = name.find_last_of('/');
    if (slash != std::string::npos) name = name.substr(slash + 1);
    return load(ss.str(), name);
}

void State::run(const std::string &src, const std::string &chunkname) {
    Value fn = load(src, chunkname);
    ValueList out;
    call(fn, nullptr, 0, out);
}

void State::call(const Value &fn, const Value *args, int nargs, ValueList &out) {
    int saved_depth = depth;
    size_t saved_top = stack_top_;
    try {
        call_value(*this, fn, args, nargs, out, nullptr, 0);
    } catch (...) {
        depth = saved_depth;
        // StackGuard already unwound the frames; make sure the top is consistent
        while (stack_top_ > saved_top) stack_[--stack_top_] = Value();
        throw;
    }
}

static void reg(State &L, Table *t, const char *name, CFunction f, void *ud = nullptr) {
    t->set(L.new_string(name), L.new_cfunction(f, ud, name));
}
static void reg1(State &L, Table *t, const char *name, CFunction f, double (*fast)(double)) {
    Value v = L.new_cfunction(f, nullptr, name);
    static_cast<Function *>(v.obj())->fast1 = fast;
    t->set(L.new_string(name), v);
}
static void reg2(State &L, Table *t, const char *name, CFunction f, double (*fast)(double, double)) {
    Value v = L.new_cfunction(f, nullptr, name);
    static_cast<Function *>(v.obj())->fast2 = fast;
    t->set(L.new_string(name), v);
}
// the numeric cores of the math1 wrappers above (must compute exactly what they compute)
static double f_abs(double x) { return std::fabs(x); }
static double f_acos(double x) { return std::acos(x); }
static double f_asin(double x) { return std::asin(x); }
static double f_atan(double x) { return std::atan(x); }
static double f_ceil(double x) { return std::ceil(x); }
static double f_cos(double x) { return std::cos(x); }
static double f_cosh(double x) { return std::cosh(x); }
static double f_exp(double x) { return std::exp(x); }
static double f_floor(double x) { return std::floor(x); }
static double f_log(double x) { return std::log(x); }
static double f_log10(double x) { return std::log10(x); }
static double f_sin(double x) { return std::sin(x); }
static double f_sinh(double x) { return std::sinh(x); }
static double f_sqrt(double x) { return std::sqrt(x); }
static double f_tan(double x) { return std::tan(x); }
static double f_tanh(double x) { return std::tanh(x); }
static double f_atan2(double y, double x) { return std::atan2(y, x); }
static double f_pow(double x, double y) { return std::pow(x, y); }
static double f_fmod(double x, double y) { return std::fmod(x, y); }

void State::open_libs() {
    register_function("print", b_print);
    register_function("type", b_type);
    register_function("tostring", b_tostring);
    register_function("tonumber", b_tonumber);
    register_function("error", b_error);
    register_function("assert", b_assert);
    register_function("select", b_select);
    register_function("next", b_next);
    register_function("pairs", b_pairs);
    register_function("ipairs", b_ipairs);
    register_function("pcall", b_pcall);
    register_function("xpcall", b_xpcall);
    register_function("collectgarbage", b_collectgarbage);
    register_function("load", b_load);
    register_function("loadstring", b_load);  // 5.1 name
    register_function("setmetatable", b_setmetatable);
    register_function("getmetatable", b_getmetatable);
    register_function("rawequal", b_rawequal);
    register_function("rawlen", b_rawlen);
    register_function("rawget", b_rawget);
    register_function("rawset", b_rawset);
    set_global("_VERSION", new_string("Lua 5.2"));

    Value mv = new_table();
    Table *m = static_cast<Table *>(mv.obj());
    reg1(*this, m, "abs", m_abs, f_abs);
    reg1(*this, m, "acos", m_acos, f_acos);
    reg1(*this, m, "asin", m_asin, f_asin);
    reg1(*this, m, "atan", m_atan, f_atan);
    reg2(*this, m, "atan2", m_atan2, f_atan2);
    reg1(*this, m, "ceil", m_ceil, f_ceil);
    reg1(*this, m, "cos", m_cos, f_cos);
    reg1(*this, m, "cosh", m_cosh, f_cosh);
    reg(*this, m, "deg", m_deg);
    reg1(*this, m, "exp", m_exp, f_exp);
    reg1(*this, m, "floor", m_floor, f_floor);
    reg2(*this, m, "fmod", m_fmod, f_fmod);
    reg(*this, m, "frexp", m_frexp);
    reg(*this, m, "ldexp", m_ldexp);
    reg1(*this, m, "log", m_log, f_log);  // one-argument form; log(x, base) has 2 args and takes the general route
    reg1(*this, m, "log10", m_log10, f_log10);
    reg(*this, m, "max", m_max);
    reg(*this, m, "min", m_min);
    reg(*this, m, "modf", m_modf);
    reg2(*this, m, "pow", m_pow, f_pow);
    reg(*this, m, "rad", m_rad);
    reg1(*this, m, "sin", m_sin, f_sin);
    reg1(*this, m, "sinh", m_sinh, f_sinh);
    reg1(*this, m, "sqrt", m_sqrt, f_sqrt);
    reg1(*this, m, "tan", m_tan, f_tan);
    reg1(*this, m, "tanh", m_tanh, f_tanh);
    // one RNG per State, leaked deliberately small (freed with the process)
    Rng *rng = new Rng();
    reg(*this, m, "random", m_random, rng);
    reg(*this, m, "randomseed", m_randomseed, rng);
    m->set(new_string("pi"), Value(M_PI));
    m->set(new_string("huge"), Value(HUGE_VAL));
    set_global("math", mv);

    Value tv = new_table();
    Table *t = static_cast<Table *>(tv.obj());
    reg(*this, t, "unpack", t_unpack);
    reg(*this, t, "insert", t_insert);
    reg(*this, t, "remove", t_remove);
    reg(*this, t, "concat", t_concat);
    reg(*this, t, "sort", t_sort);
    reg(*this, t, "pack", t_pack);
    set_global("table", tv);
    set_global("unpack", t->get(new_string("unpack")));  // 5.1 alias, harmless

    Value sv = new_table();
    Table *s = static_cast<Table *>(sv.obj());
    reg(*this, s, "len", s_len);
    reg(*this, s, "sub", s_sub);
    reg(*this, s, "rep", s_rep);
    reg(*this, s, "upper", s_upper);
    reg(*this, s, "lower", s_lower);
    reg(*this, s, "byte", s_byte);
    reg(*this, s, "char", s_char);
    reg(*this, s, "format", s_format);
    reg(*this, s, "find", s_find);
    reg(*this, s, "match", s_match);
    reg(*this, s, "reverse", s_reverse);
    set_global("string", sv);

    // the corners of os / io a lens script might touch (clock for timing prints, write for progress dots); no file access
    Value ov = new_table();
    Table *o = static_cast<Table *>(ov.obj());
    reg(*this, o, "time", os_time);
    reg(*this, o, "clock", os_clock);
    reg(*this, o, "date", os_date);
    reg(*this, o, "getenv", os_getenv);
    set_global("os", ov);
    Value iv = new_table();
    reg(*this, static_cast<Table *>(iv.obj()), "write", io_write);
    set_global("io", iv);
    run(kStringPrelude, "=string");
}

// ---------------------------------------------------------------------------
// clone
// ---------------------------------------------------------------------------

namespace {

struct Cloner {
    State &dst;
    explicit Cloner(State &d) : dst(d) {}
    std::unordered_map<const Object *, Object *> map;

    Value clone(const Value &v) {
        switch (v.type()) {
            case Type::Nil:
            case Type::Boolean:
            case Type::Number: return v;
            case Type::String: {
                auto it = map.find(v.obj());
                if (it != map.end()) return Value::object(Type::String, it->second);
                Str *s = dst.alloc_str(v.str());
                map.emplace(v.obj(), s);
                return Value::object(Type::String, s);
            }
            case Type::Table: {
                auto it = map.find(v.obj());
                if (it != map.end()) return Value::object(Type::Table, it->second);
                const Table *src = static_cast<const Table *>(v.obj());
                Table *t = dst.alloc_table();
                Value tv = Value::object(Type::Table, t);
                map.emplace(v.obj(), t);
                t->arr.reserve(src->arr.size());
                for (const Value &e : src->arr) t->arr.push_back(clone(e));
                for (const Value &k : src->hash_order) {
                    Value nk = clone(k);
                    t->hash.emplace(nk, clone(src->hash.find(k)->second));
                    t->hash_order.push_back(nk);
                }
                t->meta = clone(src->meta);
                return tv;
            }
            case Type::Function: {
                auto it = map.find(v.obj());
                if (it != map.end()) return Value::object(Type::Function, it->second);
                const Function *src = static_cast<const Function *>(v.obj());
                Function *fn = dst.alloc_function();
                Value fv = Value::object(Type::Function, fn);
                map.emplace(v.obj(), fn);
                fn->cfn = src->cfn;
                fn->fast1 = src->fast1;
                fn->fast2 = src->fast2;
                fn->ud = src->ud;
                fn->cname = src->cname;
                fn->proto = src->proto;
                fn->chunk = src->chunk;
                for (Box *b : src->upvals) {
                    Box *nb = clone_box(b);
                    ++nb->rc;
                    fn->upvals.push_back(nb);
                }
                return fv;
            }
            case Type::Box: {
                Box *nb = clone_box(static_cast<Box *>(v.obj()));
                return Value::object(Type::Box, nb);
            }
        }
        return Value();
    }

    Box *clone_box(const Box *b) {
        auto it = map.find(b);
        if (it != map.end()) return static_cast<Box *>(it->second);
        Box *nb = dst.alloc_box();
        map.emplace(b, nb);
        nb->v = clone(b->v);
        return nb;
    }
};

}  // namespace

std::unique_ptr<State> State::clone() const {
    std::unique_ptr<State> s(new State(uni_));
    // the fresh State's own math.random pair (bound to its private RNG)
    Value own_math = s->get_global("math");
    Value own_random, own_seed;
    if (own_math.is_table()) {
        own_random = static_cast<Table *>(own_math.obj())->get_str("random");
        own_seed = static_cast<Table *>(own_math.obj())->get_str("randomseed");
    }
    Cloner c(*s);
    s->globals_.resize(globals_.size());
    for (size_t i = 0; i < globals_.size(); ++i) s->globals_[i] = c.clone(globals_[i]);
    s->print_sink_ = print_sink_;
    s->print_ud_ = print_ud_;
    // math.random's RNG userdata must not be shared between threads: re-point
    // the cloned `math` table at the new State's private pair, carrying over
    // the source generator's current state.
    Value m = s->get_global("math");
    Value src_math = const_cast<State *>(this)->get_global("math");
    if (m.is_table() && src_math.is_table() && own_random.is_function()) {
        Value src_random = static_cast<Table *>(src_math.obj())->get_str("random");
        if (src_random.is_function() && static_cast<Function *>(src_random.obj())->cfn == m_random) {
            *static_cast<Rng *>(static_cast<Function *>(own_random.obj())->ud) =
                *static_cast<Rng *>(static_cast<Function *>(src_random.obj())->ud);
            static_cast<Table *>(m.obj())->set(s->new_string("random"), own_random);
            static_cast<Table *>(m.obj())->set(s->new_string("randomseed"), own_seed);
        }
    }
    return s;
}

}  // namespace minilua
