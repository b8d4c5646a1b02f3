"""A small ConstraintSystem recorder in Python, shaped like bulletproofs' r1cs API (what src/gadgets.rs is written against):
`multiply(l, r)` allocates a multiplier, evaluates both linear combinations on the assignment and contributes the two
constraints l - L_i = 0, r - R_i = 0; `constrain(lc)` contributes lc = 0. `flatten()` yields the bbp_cs arrays."""
from orc import L_ORDER, le

COMMITTED, LEFT, RIGHT, OUT, ONE = range(5)


class LC:
    def __init__(self, terms=None):
        self.terms = list(terms or [])          # [(kind, index, coefficient int)]

    @staticmethod
    def var(kind, idx):
        return LC([(kind, idx, 1)])

    @staticmethod
    def const(c):
        return LC([(ONE, 0, c % L_ORDER)])

    def __add__(self, o):
        return LC(self.terms + o.terms)

    def __sub__(self, o):
        return LC(self.terms + [(k, i, (-c) % L_ORDER) for k, i, c in o.terms])

    def scale(self, s):
        return LC([(k, i, c * s % L_ORDER) for k, i, c in self.terms])


class Recorder:
    def __init__(self, values):
        self.v = [x % L_ORDER for x in values]   # committed values
        self.aL, self.aR, self.aO = [], [], []
        self.cons = []

    def committed(self, i):
        return LC.var(COMMITTED, i)

    def eval(self, lc):
        acc = 0
        for k, i, c in lc.terms:
            val = {COMMITTED: lambda: self.v[i], LEFT: lambda: self.aL[i], RIGHT: lambda: self.aR[i], OUT: lambda: self.aO[i], ONE: lambda: 1}[k]()
            acc = (acc + c * val) % L_ORDER
        return acc

    def multiply(self, left, right):
        i = len(self.aL)
        lv, rv = self.eval(left), self.eval(right)
        self.aL.append(lv); self.aR.append(rv); self.aO.append(lv * rv % L_ORDER)
        l, r, o = LC.var(LEFT, i), LC.var(RIGHT, i), LC.var(OUT, i)
        self.cons.append(left - l)
        self.cons.append(right - r)
        return l, r, o

    def constrain(self, lc):
        self.cons.append(lc)

    def flatten(self):
        con_ptr, term_var, coeff = [0], [], b""
        for lc in self.cons:
            for k, i, c in lc.terms:
                term_var.append((k << 28) | i)
                coeff += le(c)
            con_ptr.append(len(term_var))
        return dict(n_mul=len(self.aL), m=len(self.v), con_ptr=con_ptr, term_var=term_var, term_coeff=coeff)

    def witness(self):
        pack = lambda xs: b"".join(le(x) for x in xs)
        return pack(self.aL), pack(self.aR), pack(self.aO), pack(self.v)


def example_circuit(seed, n_extra=5):
    """A circuit with non +-1 coefficients, constants, committed variables in several constraints and a multiplier output
    feeding many constraints: proves knowledge of (x, y, w) with  3 x y + 5 = w,  (x + 2 y)^2 = s  public, a chain of
    cubes, and a weighted sum."""
    import random
    rnd = random.Random(seed)
    x, y = rnd.getrandbits(200), rnd.getrandbits(250)
    w = (3 * x * y + 5) % L_ORDER
    cs = Recorder([x, y, w])
    X, Y, W = cs.committed(0), cs.committed(1), cs.committed(2)
    _, _, xy = cs.multiply(X, Y)
    cs.constrain(xy.scale(3) + LC.const(5) - W)
    t = X + Y.scale(2)
    _, _, sq = cs.multiply(t, t)
    s_pub = pow((x + 2 * y) % L_ORDER, 2, L_ORDER)
    cs.constrain(sq - LC.const(s_pub))
    cur, acc, acc_val = sq, LC(), 0
    for k in range(n_extra):
        coeff = rnd.getrandbits(252) % L_ORDER
        _, _, c2 = cs.multiply(cur + LC.const(k + 7), cur.scale(coeff) - X)
        acc = acc + c2.scale(k + 2) + sq            # sq feeds every round
        acc_val = (acc_val + (k + 2) * cs.aO[-1] + cs.aO[1]) % L_ORDER
        cur = c2
    cs.constrain(acc - LC.const(acc_val))
    return cs


def random_circuit(seed, n_mul, n_commit, n_free):
    """a random satisfiable circuit: n_mul multipliers whose inputs are random linear combinations (random 252-bit
    coefficients, small ones, +-1 and constants mixed) of everything allocated so far, plus n_free extra constraints that are
    satisfied by construction (a random combination minus its own value)."""
    import random
    rnd = random.Random(seed)
    cs = Recorder([rnd.getrandbits(rnd.choice([1, 64, 252])) for _ in range(n_commit)])
    pool = [cs.committed(i) for i in range(n_commit)]

    def coeff():
        k = rnd.random()
        if k < 0.35:
            return 1
        if k < 0.5:
            return L_ORDER - 1
        if k < 0.7:
            return rnd.randrange(2, 1000)
        return rnd.getrandbits(252) % L_ORDER

    def random_lc():
        lc = LC.const(rnd.getrandbits(200)) if rnd.random() < 0.4 or not pool else LC()
        for _ in range(rnd.randrange(1, 5)):
            if pool:
                lc = lc + rnd.choice(pool).scale(coeff())
        return lc

    for _ in range(n_mul):
        l, r, o = cs.multiply(random_lc(), random_lc())
        pool += [l, r, o]
    for _ in range(n_free):
        lc = random_lc()
        cs.constrain(lc - LC.const(cs.eval(lc)))
    return cs
