"""The Lua-subset evaluator (blinky_b200/csrc/minilua) exercised like a `lua` binary.
Expected values follow the Lua 5.2 reference manual."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "blinky_b200", "csrc", "minilua", "minilua")


@pytest.fixture(scope="module")
def lua(bb):
    if not os.path.exists(EXE):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.check_call(["make", "minilua"], cwd=os.path.join(ROOT, "blinky_b200"), env=env)

    def run(code: str):
        r = subprocess.run([EXE, "-e", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
        return r.returncode, r.stdout, r.stderr

    return run


def ok(lua, code, expect_out=None):
    rc, out, err = lua(code)
    assert rc == 0, err
    if expect_out is not None:
        assert out == expect_out
    return out


def test_assignment_and_multiple_returns(lua):
    ok(lua, """
local a, b, c = 1, 2
assert(a == 1 and b == 2 and c == nil)
local function f(...) return ... end
assert(select('#', f(1,2,3)) == 3)
local t = {f(1,2,3), f(4,5,6)}
assert(#t == 4 and t[1]==1 and t[2]==4 and t[3]==5 and t[4]==6)
assert(select('#', (f(1,2,3))) == 1)
local x, y = 1, 2
x, y = y, x
assert(x == 2 and y == 1)
local function two() return 1, 2 end
local p, q, r = two()
assert(p == 1 and q == 2 and r == nil)
local m = math.max(two())
assert(m == 2)
""")


def test_closures_and_scoping(lua):
    ok(lua, """
local function counter() local n = 0; return function() n = n + 1; return n end end
local c1, c2 = counter(), counter()
assert(c1()==1 and c1()==2 and c2()==1)
local fs = {}
for i=1,3 do fs[i] = function() return i end end
assert(fs[1]()==1 and fs[2]()==2 and fs[3]()==3)
local v = 1
do local v = 2; assert(v == 2) end
assert(v == 1)
local w = 1
local w = w + 1
assert(w == 2)
local function fib(n) if n < 2 then return n end return fib(n-1)+fib(n-2) end
assert(fib(20) == 6765)
-- a global defined later is visible at call time
function g1() return g2() + 1 end
function g2() return 41 end
assert(g1() == 42)
-- repeat..until sees the body's locals
local i = 0
repeat local k = i; i = i + 1 until k >= 4
assert(i == 5)
""")


def test_loops(lua):
    ok(lua, """
local s = 0
for x=1,2,0.25 do s = s + x end
assert(s == 7.5)
for x=10,1,-3 do s = s + 1 end
assert(s == 11.5)
for x=1,0 do error('never') end
local i = 0
while true do i = i + 1; if i > 10 then break end end
assert(i == 11)
local sum = 0
for k,v in ipairs({10,20,30}) do sum = sum + k*v end
assert(sum == 140)
local cnt = 0
for k,v in pairs({a=1,b=2,3}) do cnt = cnt + 1 end
assert(cnt == 3)
""")


def test_goto_and_labels(lua):
    """Lua 5.2 goto (the reference links liblua 5.2, so user lenses may use it): the continue idiom, backward
    jumps, jumping out of nested loops and blocks, labels at the end of repeat bodies, and a goto without label"""
    ok(lua, """
local s = 0
for i = 1, 10 do
  if i % 2 == 0 then goto continue end
  s = s + i
  ::continue::
end
assert(s == 25)
local n = 0
::top::
n = n + 1
if n < 5 then goto top end
assert(n == 5)
local hit
for i = 1, 3 do
  for j = 1, 3 do
    if i * j == 4 then hit = i * 10 + j goto out end
  end
end
::out::
assert(hit == 22)
local function f(x)
  do
    if x > 1 then goto big end
    return "small"
  end
  ::big::
  return "big"
end
assert(f(0) == "small" and f(5) == "big")
local k, skipped = 0, 0
repeat
  k = k + 1
  if k == 2 then goto skip end
  skipped = skipped + 1
  ::skip::
until k >= 3
assert(k == 3 and skipped == 2)
local i = 0
while true do
  i = i + 1
  if i > 3 then goto done end
end
::done::
assert(i == 4)
local okc, err = pcall(function() goto nowhere end)
assert(not okc and type(err) == "string")
""")
    rc, out, err = lua("goto nowhere")
    assert rc != 0 and "no visible label 'nowhere'" in err


def test_arithmetic_follows_ieee_and_lua(lua):
    ok(lua, """
assert(2^10 == 1024 and 7 % 3 == 1 and -7 % 3 == 2 and 7 % -3 == -2)
assert(2^-1 == 0.5 and -2^2 == -4 and 2^3^2 == 512)
assert(1 .. 2 == "12" and "10" + 5 == 15 and "0x10" + 0 == 16)
assert(.25 == 0.25 and 1.e-10 == 1e-10 and 0x10 == 16 and 3 == 3.0)
local nan = 0/0
assert(nan ~= nan and not (nan < 0) and not (nan >= 0))
assert(1/0 == math.huge and -1/0 == -math.huge)
assert(tostring(1/0) == "inf" and tostring(10/2) == "5" and tostring(0.1) == "0.1")
assert(tostring(2^53) == "9.007199254741e+15")
assert(0.1 + 0.2 ~= 0.3)
assert(19.73920880217871723738 == 19.739208802178716)
assert(math.pi == 3.14159265358979323846)
local ip, fp = math.modf(3.75); assert(ip == 3 and fp == 0.75)
ip, fp = math.modf(-3.75); assert(ip == -3 and fp == -0.75)
assert(math.floor(-0.5) == -1 and math.ceil(-0.5) == 0 and math.abs(-3) == 3)
assert(math.fmod(7, 3) == 1 and math.fmod(-7, 3) == -1)
assert(math.sqrt(2)*math.sqrt(2) ~= 2)
""")


def test_logic_and_comparison(lua):
    ok(lua, """
assert(not nil == true and (nil or 5) == 5 and (false and 1) == false and (1 and 2) == 2)
assert(1 < 2 and "a" < "b" and 2 >= 2 and not (1 == "1"))
assert(nil == nil and nil ~= false)
local t = {}
assert(t == t and {} ~= {})
""")


def test_tables(lua):
    ok(lua, """
local cols = {3,3}
local r = math.modf(0.5)
assert(cols[r+1] == 3)            -- float key 1.0 hits the array part
local t = {}
t[1.0] = "x"; t[2] = "y"
assert(#t == 2 and t[1] == "x")
t = {1,2,3,nil}
assert(#t == 3)
t = {n=1, [2]="b", "a"}
assert(t.n == 1 and t[1] == "a" and t[2] == "b" and #t == 2)
t = {{1,2},{3,4},}
assert(t[2][1] == 3)
t = {}
t[3] = 'c'; t[2] = 'b'; t[1] = 'a'
assert(#t == 3)
assert(math.max(table.unpack({3,9,2})) == 9)
table.insert(t, 'd'); assert(#t == 4 and t[4] == 'd')
assert(table.remove(t) == 'd' and #t == 3)
assert(table.concat({1,2,3}, ",") == "1,2,3")
""")


def test_strings(lua):
    ok(lua, """
assert(("abc"):len() == 3 and ("abc"):upper() == "ABC" and #"abc" == 3)
assert(string.format("%d %5.2f %s %g", 3, 1.5, "z", 0.1) == "3  1.50 z 0.1")
assert(string.sub("hello", 2, 4) == "ell" and string.rep("ab", 3) == "ababab")
local ls = [[
hello]]
assert(ls == "hello")
assert("a\\n" == "a\\10" and '\\65' == "A" and "\\x41" == "A")
--[[ multi
line ]] assert(true)
--[==[ another ]==]
assert(tonumber("12") == 12 and tonumber("z") == nil and tonumber("ff", 16) == 255)
""")


def test_string_patterns(lua):
    """string.find / match / gmatch / gsub / reverse with Lua patterns (reference manual 6.4.1)"""
    ok(lua, """
assert(select('#', string.find("hello world", "o w")) == 2)
local a, b = string.find("hello world", "o w") assert(a == 5 and b == 7)
a, b = string.find("hello world", "l+") assert(a == 3 and b == 4)
assert(string.find("hello", "xyz") == nil)
a, b = string.find("a.b", ".", 1, true) assert(a == 2 and b == 2)
local s, e, k, v = string.find("key = value", "(%w+)%s*=%s*(%w+)") assert(s == 1 and e == 11 and k == "key" and v == "value")
local y, m, d = string.match("2024-09-23", "(%d+)-(%d+)-(%d+)") assert(y == "2024" and m == "09" and d == "23")
assert(string.match("  trim  ", "^%s*(.-)%s*$") == "trim")
local p1, p2 = string.match("hello", "()ll()") assert(p1 == 3 and p2 == 5)
assert(string.match("THE (quick) fox", "%((%a+)%)") == "quick")
assert(string.match("f(a(b)c)d", "%b()") == "(a(b)c)")
assert(string.match("THE END", "%f[%a]%a+", 5) == "END")
assert(string.match("0x1F", "^0[xX](%x+)$") == "1F")
assert(string.match("aaa", "a-b") == nil and string.match("aaab", "a-b") == "aaab")
assert(string.match("x=1;x=1", "(x=%d);%1") == "x=1")
local r, n = string.gsub("hello world", "o", "0") assert(r == "hell0 w0rld" and n == 2)
r, n = string.gsub("hello world", "(%w+)", "<%1>") assert(r == "<hello> <world>" and n == 2)
r, n = string.gsub("abc", "", "-") assert(r == "-a-b-c-" and n == 4)
r, n = string.gsub("hello", "l", {l = "L"}) assert(r == "heLLo" and n == 2)
r, n = string.gsub("1 2 3", "%d", function(d) return d * 2 end, 2) assert(r == "2 4 3" and n == 2)
r, n = string.gsub("abc", "^a", "X") assert(r == "Xbc" and n == 1)
r, n = string.gsub("x = $y", "%$(%w+)", "%%%1") assert(r == "x = %y" and n == 1)
r, n = string.gsub("abc", "%w", "%0%0") assert(r == "aabbcc" and n == 3)
local t = {}
for k, v in string.gmatch("a=1, b=2", "(%w+)=(%w+)") do t[#t + 1] = k .. v end
assert(table.concat(t, ",") == "a1,b2")
t = {}
for w in string.gmatch("one two", "%a+") do t[#t + 1] = w end
assert(#t == 2 and t[2] == "two")
assert(("abc"):reverse() == "cba" and ("abc"):find("c") == 3)
assert(not pcall(string.find, "a", "[a"))
assert(not pcall(string.gsub, "a", "a", "%2"))
""")


def test_metatables(lua):
    """setmetatable / getmetatable and the events a math-heavy script would use: __index (table chain and function),
    __newindex, arithmetic, __unm, __eq, __lt, __le, __len, __concat, __call, __tostring, protected metatables"""
    out = ok(lua, """
local V = {}
V.__index = V
local function vec(x, y, z) return setmetatable({x = x, y = y, z = z}, V) end
V.__add = function(a, b) return vec(a.x + b.x, a.y + b.y, a.z + b.z) end
V.__sub = function(a, b) return vec(a.x - b.x, a.y - b.y, a.z - b.z) end
V.__mul = function(a, b)
  if type(a) == "number" then return vec(a * b.x, a * b.y, a * b.z) end
  if type(b) == "number" then return vec(a.x * b, a.y * b, a.z * b) end
  return a.x * b.x + a.y * b.y + a.z * b.z
end
V.__div = function(a, b) return vec(a.x / b, a.y / b, a.z / b) end
V.__unm = function(a) return vec(-a.x, -a.y, -a.z) end
V.__eq = function(a, b) return a.x == b.x and a.y == b.y and a.z == b.z end
V.__lt = function(a, b) return a:len() < b:len() end
V.__le = function(a, b) return a:len() <= b:len() end
V.__len = function(a) return 3 end
V.__tostring = function(a) return "(" .. a.x .. "," .. a.y .. "," .. a.z .. ")" end
V.__concat = function(a, b) return tostring(a) .. tostring(b) end
V.__call = function(a, k) return a[k] end
function V.len(a) return math.sqrt(a * a) end
local a, b = vec(1, 2, 3), vec(4, 5, 6)
assert(tostring(a + b) == "(5,7,9)" and tostring(b - a) == "(3,3,3)" and a * b == 32)
assert(tostring(2 * a) == "(2,4,6)" and tostring(a * 2) == "(2,4,6)" and tostring(-a) == "(-1,-2,-3)" and tostring(b / 2) == "(2,2.5,3)")
assert(a == vec(1, 2, 3) and a ~= b and a < b and a <= b and not (a > b) and b >= a)
assert(#a == 3 and a .. b == "(1,2,3)(4,5,6)" and a("y") == 2 and math.abs(a:len() - 3.7416573867739) < 1e-12)
local log = {}
local t = setmetatable({}, {__index = function(t, k) return k .. "!" end,
                            __newindex = function(t, k, v) rawset(t, k, v * 2) log[#log + 1] = k end})
t.a = 5
t.a = 7   -- present now: plain assignment
assert(t.a == 7 and t.zzz == "zzz!" and rawget(t, "zzz") == nil and #log == 1)
local store = {}
local proxy = setmetatable({}, {__newindex = store, __index = store})
proxy.k = 1
assert(rawget(proxy, "k") == nil and store.k == 1 and proxy.k == 1)
local Base = {hello = function() return "base" end}
Base.__index = Base
local Derived = setmetatable({}, Base)
Derived.__index = Derived
local obj = setmetatable({}, Derived)
assert(obj.hello() == "base" and getmetatable(obj) == Derived and getmetatable("x").__index == string)
local p = setmetatable({}, {__metatable = "locked"})
assert(getmetatable(p) == "locked" and not pcall(setmetatable, p, {}))
assert(not pcall(function() return {} + 1 end))
assert(setmetatable(obj, nil) == obj and getmetatable(obj) == nil and obj.hello == nil)
print(a, 1)
""")
    assert out == "(1,2,3)\t1\n"


def test_base_library_corners(lua):
    """xpcall, load, collectgarbage and the harmless corners of os / io (the reference opens all libraries,
    fisheye.c:1228; file access stays out)"""
    out = ok(lua, """
local okc, m = xpcall(function() error("x") end, function(msg) return "handled: " .. msg end)
assert(okc == false and m:find("handled: ", 1, true) == 1 and m:find("x", 1, true))
local ok2, v = xpcall(function(a, b) return a + b end, print, 1, 2)
assert(ok2 == true and v == 3)
assert(collectgarbage("count") == 0 and collectgarbage() == 0)
local f = load("return 1 + 2")
assert(f() == 3)
local g, err = load("syntax error here")
assert(g == nil and type(err) == "string")
assert(type(os.time()) == "number" and type(os.clock()) == "number" and os.date("!%Y", 0) == "1970")
assert(os.getenv("BLINKY_NO_SUCH_VARIABLE") == nil)
assert(io.open == nil and os.remove == nil and os.execute == nil)
io.write("a", 1, "b")
""")
    assert out == "a1b"


def test_errors_are_caught_and_positioned(lua):
    out = ok(lua, """
local ok, err = pcall(function() local z = nil; return z + 1 end)
assert(not ok)
print(err)
ok, err = pcall(function() error("boom") end)
print(ok, err)
ok, err = pcall(function() local t = nil; return t.x end)
print(err)
ok, err = pcall(function() undefined_function() end)
print(err)
""")
    lines = out.strip().split("\n")
    assert "attempt to perform arithmetic on a nil value" in lines[0] and ":2:" in lines[0]
    assert lines[1].startswith("false") and "boom" in lines[1]
    assert "attempt to index" in lines[2]
    assert "attempt to call" in lines[3] and "undefined_function" in lines[3]


def test_syntax_errors_fail_to_load(lua):
    rc, out, err = lua("x = = 1")
    assert rc != 0 and "unexpected symbol" in err
    rc, out, err = lua("for i=1,2 do")
    assert rc != 0 and "'end' expected" in err
    rc, out, err = lua("goto done")
    assert rc != 0


def test_stack_overflow_is_an_error_not_a_crash(lua):
    out = ok(lua, "local function f() return 1 + f() end print(pcall(f))")
    assert out.startswith("false") and "stack overflow" in out


def test_cycles_are_collected(lua):
    # closures that capture themselves form reference cycles; memory must stay bounded
    ok(lua, """
for i = 1, 300000 do
  local function rec(n) if n == 0 then return 0 end return rec(n-1) end
  rec(1)
  local t = {}; t.self = t
end
print("done")
""", "done\n")


def test_table_sort(lua):
    ok(lua, """
local t = {3,1,2,9,-4,7.5}
table.sort(t); print(table.concat(t, ","))
table.sort(t, function(a,b) return a > b end); print(table.concat(t, ","))
local w = {"pear","apple","fig"}; table.sort(w); print(table.concat(w, " "))
print(pcall(table.sort, {1,"x"}))
local big = {}
for i = 1, 1000 do big[i] = (i * 7919) % 1013 end
table.sort(big)
local sorted = true
for i = 2, #big do if big[i-1] > big[i] then sorted = false end end
print(sorted, #big)
""", "-4,1,2,3,7.5,9\n9,7.5,3,2,1,-4\napple fig pear\nfalse\tattempt to compare string with number\ntrue\t1000\n")
