"""CPU check of the product's dedicated field squaring (csrc/fe25519.cuh: fe_sq_wide) without a GPU: the inline-PTX
carry-chain blocks are parsed out of the source and executed by a small interpreter (mad.lo/hi with .cc / madc, add /
addc), the C++ call sequence around them is restated here, and the 512-bit result is compared with Python integers.
Catches operand-numbering and column-layout mistakes before GPU time is spent; the -m gpu test
test_gpu_primitives.py::test_field_ops_match_bigints (op 4) runs the real kernel."""
import os
import random
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "dusk-blindbidproof_b200", "csrc", "fe25519.cuh")).read()
M32 = (1 << 32) - 1


class Ptr:
    def __init__(self, arr, off=0):
        self.arr, self.off = arr, off

    def __getitem__(self, i):
        return self.arr[self.off + i]

    def __setitem__(self, i, v):
        self.arr[self.off + i] = v


def parse_asm_at(pos):
    """pos = index of 'asm(' in SRC -> (instructions, operands[(constraint, C expression)])"""
    depth, k = 0, pos + 3
    while True:
        if SRC[k] == "(":
            depth += 1
        elif SRC[k] == ")":
            depth -= 1
            if depth == 0:
                break
        k += 1
    body = SRC[pos + 4:k]
    parts, cur, in_str = [], "", False
    for ch in body:
        if ch == '"':
            in_str = not in_str
        if ch == ":" and not in_str:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    tmpl = "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', parts[0])).replace("\\n", "").replace("\\t", "")
    ops = [(m.group(1), m.group(2).strip()) for part in parts[1:3] for m in re.finditer(r'"([+=]?&?r)"\(([^)]*)\)', part)]
    return [x.strip() for x in tmpl.split(";") if x.strip()], ops


def run_asm(pos, env):
    instrs, ops = parse_asm_at(pos)

    def rd(tok):
        tok = tok.strip()
        return (eval(ops[int(tok[1:])][1], {}, env) & M32) if tok.startswith("%") else int(tok)

    def wr(tok, v):
        kind, expr = ops[int(tok.strip()[1:])]
        assert kind in ("+r", "=r", "+&r", "=&r"), "write to an input operand"
        m = re.match(r"(\w+)\[(\d+)\]$", expr)
        if m:
            env[m.group(1)][int(m.group(2))] = v & M32
        else:
            env[expr] = v & M32

    cf = 0
    for ins in instrs:
        op, rest = ins.split(None, 1)
        a = [x.strip() for x in rest.split(",")]
        f = op.split(".")
        if f[0] in ("mad", "madc"):
            p = rd(a[1]) * rd(a[2])
            t = ((p >> 32) if "hi" in f else (p & M32)) + rd(a[3]) + (cf if f[0] == "madc" else 0)
        elif f[0] in ("add", "addc"):
            t = rd(a[1]) + rd(a[2]) + (cf if f[0] == "addc" else 0)
        else:
            raise AssertionError("unhandled instruction " + ins)
        wr(a[0], t)
        if "cc" in f:
            cf = t >> 32


def helper(name, c, xs, y):
    pos = SRC.index("asm(", SRC.index(name + "(uint32_t *c"))
    env = {"c": c, "y": y}
    env.update({"x%d" % i: x for i, x in enumerate(xs)})
    run_asm(pos, env)
    return env.get("cy")


def mulN(c, xs, y):
    for t, x in enumerate(xs):
        c[2 * t], c[2 * t + 1] = (x * y) & M32, (x * y) >> 32


def fe_sq_wide(a):
    """restates the body of fe_sq_wide() call by call"""
    body0 = SRC.index("BBP_DEV void fe_sq_wide")
    body = SRC[body0:SRC.index("\n}\n", body0)]
    calls = re.findall(r"(?:(ev|od)\[(\d+)\] = )?(mul4|mul3|mad\d_\w+)\((ev|od)(?: \+ (\d+))?, ([^;]*)\);", body)
    assert len(calls) == 13, "fe_sq_wide no longer has the 13 chain calls this test restates"
    acc = {"ev": [None] * 14, "od": [None] * 14}
    for ret_arr, ret_idx, fn, arr, off, args in calls:
        vals = [eval(x, {}, {"a": a}) for x in args.split(",")]
        dst = Ptr(acc[arr], int(off or 0))
        if fn in ("mul4", "mul3"):
            mulN(dst, vals[:-1], vals[-1])
        else:
            cy = helper(fn, dst, vals[:-1], vals[-1])
            if ret_arr:
                acc[ret_arr][int(ret_idx)] = cy
    ev, od = acc["ev"], acc["od"]
    blocks = [body0 + m.start() for m in re.finditer(r"asm\(", body)]
    assert len(blocks) == 2
    s = [None] * 15
    s[1] = od[0]
    run_asm(blocks[0], {"ev": ev, "od": od, "s": s})
    t = [None] * 16
    t[1] = (s[1] << 1) & M32
    for k in range(2, 15):
        t[k] = ((s[k] << 1) | (s[k - 1] >> 31)) & M32
    t[15] = s[14] >> 31
    r = [None] * 16
    run_asm(blocks[1], {"a": list(a), "t": t, "r": r})
    return r


def test_fe_sq_wide_ptx_blocks_square_correctly():
    rnd = random.Random(7)
    cases = [[M32] * 8, [0] * 8, [M32, 0] * 4, [0, M32] * 4, [1] + [0] * 7, [0] * 7 + [M32]]
    cases += [[rnd.choice([0, 1, M32, rnd.getrandbits(32)]) for _ in range(8)] for _ in range(1500)]
    for a in cases:
        A = sum(x << (32 * i) for i, x in enumerate(a))
        r = fe_sq_wide(a)
        assert None not in r
        assert sum(x << (32 * i) for i, x in enumerate(r)) == A * A, a
