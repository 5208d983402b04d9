"""Randomised pinning of the Lua -> C++ lens translator: seeded random lens functions over the
supported subset (arithmetic, math.*, comparisons, and/or/not, if/elseif, numeric for, while,
repeat, local arrays, helper functions with several results, script-level constants and state,
`return nil`) are run through the interpreter and through the g++-compiled translation; every
result must have the same bit pattern.  Same libm on both sides, so any difference is a translator
(or interpreter) bug, not rounding."""
import ctypes
import random
import struct

import numpy as np
import pytest

from test_transpile import WRAP, _compile_host

UNARY = ["sin", "cos", "atan", "tanh", "abs", "floor", "ceil", "exp1", "sqrtabs", "logabs", "tan", "asinc", "sinh1"]
BINARY = ["+", "-", "*", "/", "atan2", "max", "min", "fmodnz", "powint", "mod"]


class Gen:
    def __init__(self, seed):
        self.r = random.Random(seed)
        self.lines = []
        self.nvars = 0
        self.indent = 1
        self.helpers = []  # (name, nargs, nres)
        self.predicates = []  # (name, nargs): helper functions returning a boolean
        self.arrays = []   # (name, size)

    def emit(self, s):
        self.lines.append("  " * self.indent + s)

    def const(self):
        c = self.r.choice([0.5, 1, 2, 3, -1.25, 0.1, 7, 1e-3, 2.75, self.r.uniform(-3, 3)])
        return repr(float(c)) if self.r.random() < 0.7 else repr(c)

    def expr(self, names, depth=0):
        r = self.r
        if depth > 3 or r.random() < 0.25:
            k = r.random()
            if k < 0.6 and names:
                return r.choice(names)
            if k < 0.7 and self.arrays:
                a, n = r.choice(self.arrays)
                return f"{a}[{r.randint(1, n)}]"
            if k < 0.8:
                return r.choice(["pi", "K1", "K2", "tab[2]", "#tab"])
            return self.const()
        if r.random() < 0.45:
            f = r.choice(UNARY)
            a = self.expr(names, depth + 1)
            return {
                "exp1": f"exp(min({a}, 3))", "sqrtabs": f"sqrt(abs({a}))", "logabs": f"log(abs({a}) + 0.5)",
                "asinc": f"asin(max(-1, min(1, {a})))", "sinh1": f"sinh(max(-2, min(2, {a})))",
            }.get(f, f"{f}({a})")
        op = r.choice(BINARY)
        a, b = self.expr(names, depth + 1), self.expr(names, depth + 1)
        if op in "+-*/":
            return f"({a} {op} {b})"
        if op == "fmodnz":
            return f"math.fmod({a}, abs({b}) + 0.75)"
        if op == "mod":
            return f"({a} % (abs({b}) + 1.5))"
        if op == "powint":
            return f"({a}) ^ {r.choice([2, 3, 0.5, -1])}" if r.random() < 0.7 else f"pow(abs({a}) + 0.1, {b})"
        return f"{op}({a}, {b})"

    def cond(self, names, depth=0):
        r = self.r
        if depth < 2 and r.random() < 0.3:
            j = r.choice(["and", "or"])
            return f"({self.cond(names, depth + 1)} {j} {self.cond(names, depth + 1)})"
        if depth < 2 and r.random() < 0.15:
            return f"not ({self.cond(names, depth + 1)})"
        if self.predicates and r.random() < 0.2:
            name, nargs = r.choice(self.predicates)
            return f"{name}({', '.join(self.expr(names, 2) for _ in range(nargs))})"
        return f"{self.expr(names, 2)} {r.choice(['<', '<=', '>', '>=', '==', '~='])} {self.expr(names, 2)}"

    def new_var(self):
        self.nvars += 1
        return f"v{self.nvars}"

    def block(self, names, depth=0, budget=6):
        r = self.r
        names = list(names)
        for _ in range(r.randint(2, budget)):
            k = r.random()
            if k < 0.35:
                v = self.new_var()
                self.emit(f"local {v} = {self.expr(names)}")
                names.append(v)
            elif k < 0.5 and names:
                self.emit(f"{r.choice(names)} = {self.expr(names)}")
            elif k < 0.62 and depth < 2:
                self.emit(f"if {self.cond(names)} then")
                self.indent += 1
                self.block(names, depth + 1, 3)
                self.indent -= 1
                if r.random() < 0.5:
                    self.emit(f"elseif {self.cond(names)} then")
                    self.indent += 1
                    self.block(names, depth + 1, 2)
                    self.indent -= 1
                if r.random() < 0.6:
                    self.emit("else")
                    self.indent += 1
                    self.block(names, depth + 1, 2)
                    self.indent -= 1
                self.emit("end")
            elif k < 0.72 and depth < 2 and names:
                acc = r.choice(names)
                i = self.new_var()
                step = r.choice(["", ", 2", ", 0.5"])
                self.emit(f"for {i} = 1, {r.randint(2, 6)}{step} do")
                self.indent += 1
                self.emit(f"{acc} = {acc} * 0.5 + {self.expr(names + [i], 2)}")
                if self.arrays and r.random() < 0.5:
                    a, n = r.choice(self.arrays)
                    self.emit(f"{a}[{r.randint(1, n)}] = {self.expr(names + [i], 2)}")
                if r.random() < 0.2:
                    self.emit(f"if {self.cond(names + [i])} then break end")
                self.indent -= 1
                self.emit("end")
            elif k < 0.78 and depth < 2 and names:
                c = self.new_var()
                acc = r.choice(names)
                self.emit(f"local {c} = 0")
                if r.random() < 0.5:
                    self.emit(f"while {c} < {r.randint(1, 4)} do")
                    self.indent += 1
                    self.emit(f"{acc} = {self.expr(names + [c], 2)}")
                    self.emit(f"{c} = {c} + 1")
                    self.indent -= 1
                    self.emit("end")
                else:
                    self.emit("repeat")
                    self.indent += 1
                    self.emit(f"{acc} = {acc} / 2 + {self.expr(names + [c], 3)}")
                    self.emit(f"{c} = {c} + 1")
                    self.indent -= 1
                    self.emit(f"until {c} >= {r.randint(1, 3)} or {acc} > 50")
                names.append(c)
            elif k < 0.86 and depth == 0:
                a = f"arr{len(self.arrays)}"
                n = r.randint(2, 4)
                self.emit(f"local {a} = {{{', '.join(self.expr(names, 2) for _ in range(n))}}}")
                self.arrays.append((a, n))
            elif k < 0.95 and self.helpers:
                name, nargs, nres = r.choice(self.helpers)
                outs = [self.new_var() for _ in range(nres)]
                self.emit(f"local {', '.join(outs)} = {name}({', '.join(self.expr(names, 2) for _ in range(nargs))})")
                names += outs
            elif names:
                self.emit(f"counter = counter + {self.expr(names, 3)}")
                self.emit(f"if memo ~= {names[0]} then memo = {names[0]} end")
                names += ["counter", "memo"]
        return names

    def program(self, forward=False):
        r = self.r
        head = ["local min, max, floor, ceil = math.min, math.max, math.floor, math.ceil", "local K1 = 0.625", "local K2 = sqrt(2) / 3", "local tab = {0.25, -1.5, 3, 0.125}", "local counter, memo = 0.5"]
        for h in range(r.randint(0, 2)):
            nargs, nres = r.randint(1, 3), r.randint(1, 3)
            args = [f"a{i}" for i in range(nargs)]
            self.lines, self.indent = [], 1
            saved_arrays, self.arrays = self.arrays, []
            names = self.block(args, 1, 3)
            if r.random() < 0.3:
                self.emit(f"if {self.cond(names)} then return {', '.join(self.const() for _ in range(nres))} end")
            self.emit(f"return {', '.join(self.expr(names, 2) for _ in range(nres))}")
            self.arrays = saved_arrays
            head += [f"local function h{h}({', '.join(args)})"] + self.lines + ["end"]
            self.helpers.append((f"h{h}", nargs, nres))
        if r.random() < 0.5:  # a predicate: helper returning a boolean
            args = ["a0", "a1"][: r.randint(1, 2)]
            head += [f"local function pred0({', '.join(args)})", f"  return {self.cond(args)}", "end"]
            self.predicates.append(("pred0", len(args)))
        self.lines, self.indent = [], 1
        if forward:  # lens_forward(x, y, z) -> x, y
            names = self.block(["x", "y", "z"], 0, 6)
            if r.random() < 0.4:
                self.emit(f"if {self.cond(names)} then return nil end")
            if r.random() < 0.4:
                lat, lon = self.new_var(), self.new_var()
                self.emit(f"local {lat}, {lon} = ray_to_latlon(x, y, z)")
                names += [lat, lon]
            self.emit(f"return {self.expr(names, 2)}, {self.expr(names, 2)}")
            return "\n".join(head + ["function lens_forward(x, y, z)"] + self.lines + ["end"])
        if r.random() < 0.5:
            self.emit(f"if {self.cond(['x', 'y'])} then return nil end")
        names = self.block(["x", "y"], 0, 7)
        if r.random() < 0.3:
            self.emit(f"if {self.cond(names)} then return nil end")
        if r.random() < 0.3:
            self.emit(f"return latlon_to_ray({self.expr(names, 2)}, {self.expr(names, 2)})")
        else:
            self.emit(f"return {', '.join(self.expr(names, 2) for _ in range(3))}")
        return "\n".join(head + ["function lens_inverse(x, y)"] + self.lines + ["end"])


@pytest.mark.parametrize("seed", range(40))
def test_random_lens_translation_is_bit_identical(host, tmp_path, seed):
    src = Gen(seed).program()
    host.command("f_globe cube")
    host.load_lens(f"fuzz{seed}", src)
    try:
        cpp = host.lens_source()
    except Exception as e:  # noqa: BLE001 — the generator must stay inside the subset
        pytest.fail(f"generated lens was refused: {e}\n{src}")
    lib = _compile_host(cpp, str(tmp_path / f"fuzz{seed}"), WRAP)
    out = (ctypes.c_double * 8)()
    flag = ctypes.c_uint()
    rng = np.random.default_rng(seed)
    pts = [(0.0, 0.0), (1.0, -1.0), (-0.5, 2.0)] + [tuple(rng.uniform(-3, 3, 2)) for _ in range(150)]
    for x, y in pts:
        host.load_lens(f"fuzz{seed}", src)  # script-level state back to its initial values, as on the device
        st, ray = host.lens_inverse(x, y)
        assert st in (0, 1), (seed, st, host.log[-300:], src)
        st2 = lib.lt_eval(x, y, None, 0, out, ctypes.byref(flag))
        assert st == st2, (seed, x, y, src)
        if st == 1:
            assert struct.pack("3d", *ray) == struct.pack("3d", out[0], out[1], out[2]), (seed, x, y, ray, list(out[:3]), src)


@pytest.mark.parametrize("seed", range(40))
def test_random_lens_error_bounds_are_sound(host, tmp_path, seed):
    """the same programs against a build whose libm results are off by up to 3 x 2^20 ulp (bounds
    scaled alike, see test_transpile.perturbed): wherever no risk flag is raised, the branch taken,
    the nil/values status and the float32 ray must be the exact ones"""
    from test_transpile import f32bits, perturbed

    src = Gen(seed).program()
    host.command("f_globe cube")
    host.load_lens(f"fuzz{seed}", src)
    lib = _compile_host(perturbed(host.lens_source(), 1 << 20), str(tmp_path / f"pert{seed}"), WRAP)
    out = (ctypes.c_double * 8)()
    flag = ctypes.c_uint()
    rng = np.random.default_rng(1000 + seed)
    decided = 0
    pts = [tuple(rng.uniform(-3, 3, 2)) for _ in range(250)]
    for x, y in pts:
        host.load_lens(f"fuzz{seed}", src)
        st, ray = host.lens_inverse(x, y)
        st2 = lib.lt_eval(x, y, None, 0, out, ctypes.byref(flag))
        if flag.value:
            continue
        decided += 1
        assert st == st2, (seed, x, y, st, st2, src)
        if st == 1:
            assert f32bits(ray) == f32bits([out[0], out[1], out[2]]), (seed, x, y, ray, list(out[:6]), src)


@pytest.mark.parametrize("seed", range(100, 120))
def test_random_lens_forward_translation_is_bit_identical(host, tmp_path, seed):
    from test_transpile import WRAP_FWD

    src = Gen(seed).program(forward=True)
    host.command("f_globe cube")
    host.load_lens(f"fwd{seed}", src)
    lib = _compile_host(host.lens_source(forward=True), str(tmp_path / f"fwd{seed}"), WRAP_FWD)
    out = (ctypes.c_double * 4)()
    flag = ctypes.c_uint()
    rng = np.random.default_rng(seed)
    rays = rng.normal(size=(120, 3))
    rays = (rays / np.linalg.norm(rays, axis=1, keepdims=True)).astype(np.float32).astype(np.float64)
    for rx, ry, rz in rays:
        host.load_lens(f"fwd{seed}", src)
        st, xy = host.lens_forward(rx, ry, rz)
        assert st in (0, 1), (seed, st, host.log[-300:], src)
        st2 = lib.lt_eval_fwd(rx, ry, rz, out, ctypes.byref(flag))
        assert st == st2, (seed, rx, ry, rz, src)
        if st == 1:
            assert struct.pack("2d", *xy) == struct.pack("2d", out[0], out[1]), (seed, rx, ry, rz, xy, list(out[:2]), src)
