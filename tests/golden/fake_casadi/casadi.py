"""A tiny tracing stand-in for the parts of the `casadi` API the reference's three Mpc classes use.

Purpose (tests/golden/make_golden.py): CasADi/IPOPT cannot be installed offline, but the reference's *problem
definition* is ordinary Python that only needs Opti/MX objects to trace through.  Executing the UNMODIFIED
reference sources against this stand-in yields the exact NLP they pose — decision-variable layout, objective,
equality defects, inequality rows, parameters — as an expression graph that can be evaluated (with forward-mode
derivatives) at arbitrary points.  `Opti.solve()` hands that NLP to scipy's SLSQP, so `Mpc.perform_mpc` of the
reference runs end to end with a substitute solver.  Nothing here is shipped or imported by the product.

Covered API: Opti.variable/parameter/subject_to/bounded/minimize/solver/set_initial/set_value/solve, sol.value,
MX indexing / item assignment / arithmetic / .T / .shape, exp, log, sin, cos, mtimes, vertcat, Function.
"""
import numpy as np

__version__ = "fake-3.6.3"


# ------------------------------------------------------------------------------------------------------------
# scalar expression nodes
# ------------------------------------------------------------------------------------------------------------
class Node:
    __slots__ = ("op", "args", "val", "name")
    _count = 0

    def __init__(self, op, args=(), val=None, name=None):
        self.op, self.args, self.val, self.name = op, args, val, name
        Node._count += 1


def _n(x):
    if isinstance(x, Node):
        return x
    return Node("const", (), float(x))


def _bin(op, a, b):
    return Node(op, (_n(a), _n(b)))


class Dual:
    """value + gradient (forward mode over all decision variables at once)."""
    __slots__ = ("v", "g")

    def __init__(self, v, g):
        self.v, self.g = v, g


def evaluate(roots, env, nvar=0):
    """Evaluate a list of root Nodes.  env maps leaf Node -> float (parameters) or Dual (variables).
    Returns list of Dual (if nvar) or floats.  Iterative post-order with memoisation."""
    memo = {}
    zero = np.zeros(nvar) if nvar else None

    def leaf(nd):
        if nd.op == "const":
            return Dual(nd.val, zero) if nvar else nd.val
        v = env[nd]
        if nvar and not isinstance(v, Dual):
            return Dual(float(v), zero)
        return v

    out = []
    for root in roots:
        stack = [(root, False)]
        while stack:
            nd, done = stack.pop()
            if id(nd) in memo:
                continue
            if nd.op in ("const", "sym"):
                memo[id(nd)] = leaf(nd)
                continue
            if not done:
                stack.append((nd, True))
                for a in nd.args:
                    if id(a) not in memo:
                        stack.append((a, False))
                continue
            a = [memo[id(x)] for x in nd.args]
            memo[id(nd)] = _apply(nd.op, a, nvar)
        out.append(memo[id(root)])
    return out


def _apply(op, a, nvar):
    if not nvar:
        x = a[0]
        y = a[1] if len(a) > 1 else None
        if op == "add": return x + y
        if op == "sub": return x - y
        if op == "mul": return x * y
        if op == "div": return x / y
        if op == "pow": return x ** y
        if op == "neg": return -x
        if op == "exp": return np.exp(x)
        if op == "log": return np.log(x)
        if op == "sin": return np.sin(x)
        if op == "cos": return np.cos(x)
        raise NotImplementedError(op)
    x = a[0]
    y = a[1] if len(a) > 1 else None
    if op == "add": return Dual(x.v + y.v, x.g + y.g)
    if op == "sub": return Dual(x.v - y.v, x.g - y.g)
    if op == "mul": return Dual(x.v * y.v, x.g * y.v + y.g * x.v)
    if op == "div": return Dual(x.v / y.v, (x.g * y.v - y.g * x.v) / (y.v * y.v))
    if op == "pow":
        v = x.v ** y.v
        g = y.v * x.v ** (y.v - 1) * x.g
        if np.any(y.g != 0):
            g = g + v * np.log(x.v) * y.g
        return Dual(v, g)
    if op == "neg": return Dual(-x.v, -x.g)
    if op == "exp":
        e = np.exp(x.v)
        return Dual(e, e * x.g)
    if op == "log": return Dual(np.log(x.v), x.g / x.v)
    if op == "sin": return Dual(np.sin(x.v), np.cos(x.v) * x.g)
    if op == "cos": return Dual(np.cos(x.v), -np.sin(x.v) * x.g)
    raise NotImplementedError(op)


def substitute(roots, mapping):
    """Copy of the graph below `roots` with leaf symbols replaced according to mapping {Node: Node}."""
    memo = {}

    def rec(nd):
        k = id(nd)
        if k in memo:
            return memo[k]
        if nd in mapping:
            r = mapping[nd]
        elif nd.op in ("const", "sym"):
            r = nd
        else:
            r = Node(nd.op, tuple(rec(a) for a in nd.args))
        memo[k] = r
        return r

    return [rec(r) for r in roots]


# ------------------------------------------------------------------------------------------------------------
# MX: a 2-D array of scalar nodes (column-major flattening like CasADi)
# ------------------------------------------------------------------------------------------------------------
class MX:
    __array_priority__ = 1000.0

    def __init__(self, arr):
        a = np.empty(np.shape(arr), dtype=object)
        flat_src = np.asarray(arr, dtype=object).reshape(-1)
        a.reshape(-1)[:] = [_n(v) for v in flat_src]
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(-1, 1)
        self.a = a

    # -- construction helpers
    @staticmethod
    def _wrap(x, like=None):
        if isinstance(x, MX):
            return x
        if isinstance(x, Constraint):
            raise TypeError("constraint used as expression")
        arr = np.asarray(x, dtype=float)
        if arr.ndim == 0:
            return MX(np.array([[float(arr)]], dtype=object))
        if arr.ndim == 1:
            arr = arr.reshape(-1, 1)
        return MX(arr.astype(object))

    @property
    def shape(self):
        return self.a.shape

    @property
    def T(self):
        return MX(self.a.T.copy())

    def nodes(self):
        """column-major list of scalar nodes"""
        return list(self.a.reshape(-1, order="F"))

    # -- indexing
    def _key(self, key):
        if not isinstance(key, tuple):
            # single index on a vector (column or row)
            if self.a.shape[1] == 1:
                return (key, 0)
            if self.a.shape[0] == 1:
                return (0, key)
            raise IndexError("linear indexing of a matrix is not supported by the stand-in")
        return key

    def __getitem__(self, key):
        r, c = self._key(key)
        sub = self.a[r, c]
        if isinstance(sub, Node):
            return MX(np.array([[sub]], dtype=object))
        if sub.ndim == 1:
            # keep column orientation for X[:, k] and P[a:b]; row orientation for X[0, :]
            if isinstance(r, slice) or isinstance(r, (list, np.ndarray)):
                sub = sub.reshape(-1, 1)
            else:
                sub = sub.reshape(1, -1)
        return MX(sub.copy())

    def __setitem__(self, key, value):
        r, c = self._key(key)
        v = MX._wrap(value)
        tgt = self.a[r, c]
        if isinstance(tgt, Node):
            self.a[r, c] = v.a.reshape(-1)[0]
        else:
            self.a[r, c] = v.a.reshape(np.shape(tgt))

    # -- arithmetic (elementwise with scalar broadcasting)
    def _ew(self, other, op, swap=False):
        o = MX._wrap(other)
        A, B = (o.a, self.a) if swap else (self.a, o.a)
        if A.shape != B.shape:
            if A.size == 1:
                A = np.broadcast_to(A, B.shape)
            elif B.size == 1:
                B = np.broadcast_to(B, A.shape)
            else:
                raise ValueError(f"shape mismatch {A.shape} vs {B.shape}")
        out = np.empty(A.shape, dtype=object)
        for idx in np.ndindex(A.shape):
            out[idx] = Node(op, (A[idx], B[idx]))
        return MX(out)

    def __add__(self, o): return self._ew(o, "add")
    def __radd__(self, o): return self._ew(o, "add", True)
    def __sub__(self, o): return self._ew(o, "sub")
    def __rsub__(self, o): return self._ew(o, "sub", True)
    def __mul__(self, o): return self._ew(o, "mul")
    def __rmul__(self, o): return self._ew(o, "mul", True)
    def __truediv__(self, o): return self._ew(o, "div")
    def __rtruediv__(self, o): return self._ew(o, "div", True)
    def __pow__(self, o): return self._ew(o, "pow")

    def __neg__(self):
        out = np.empty(self.a.shape, dtype=object)
        for idx in np.ndindex(self.a.shape):
            out[idx] = Node("neg", (self.a[idx],))
        return MX(out)

    def __eq__(self, o):  # noqa: PLR0124 — builds an equality constraint like casadi
        return Constraint("eq", self - MX._wrap(o))

    __hash__ = None

    def _unary(self, op):
        out = np.empty(self.a.shape, dtype=object)
        for idx in np.ndindex(self.a.shape):
            out[idx] = Node(op, (self.a[idx],))
        return MX(out)


class Constraint:
    def __init__(self, kind, expr, lo=None, hi=None):
        self.kind, self.expr, self.lo, self.hi = kind, expr, lo, hi


def exp(x): return MX._wrap(x)._unary("exp")
def log(x): return MX._wrap(x)._unary("log")
def sin(x): return MX._wrap(x)._unary("sin")
def cos(x): return MX._wrap(x)._unary("cos")


def mtimes(a, b):
    A, B = MX._wrap(a).a, MX._wrap(b).a
    if A.size == 1 or B.size == 1:
        return MX._wrap(a) * MX._wrap(b)
    if A.shape[1] != B.shape[0]:
        raise ValueError(f"mtimes shape mismatch {A.shape} x {B.shape}")
    out = np.empty((A.shape[0], B.shape[1]), dtype=object)
    for i in range(A.shape[0]):
        for j in range(B.shape[1]):
            acc = None
            for k in range(A.shape[1]):
                x, y = A[i, k], B[k, j]
                # CasADi's sparse numeric matrices drop structural zeros; keep the same terms
                if (x.op == "const" and x.val == 0.0) or (y.op == "const" and y.val == 0.0):
                    continue
                t = Node("mul", (x, y))
                acc = t if acc is None else Node("add", (acc, t))
            out[i, j] = acc if acc is not None else _n(0.0)
    return MX(out)


def vertcat(*args):
    cols = []
    for a in args:
        m = MX._wrap(a)
        cols.append(m.a if m.a.shape[1] == 1 else m.a.reshape(-1, 1))
    return MX(np.vstack(cols))


class Function:
    def __init__(self, name, ins, outs, in_names=None, out_names=None):
        self.name, self.ins, self.outs = name, [MX._wrap(i) for i in ins], [MX._wrap(o) for o in outs]

    def __call__(self, *args):
        mapping = {}
        for sym, arg in zip(self.ins, args):
            for s, a in zip(sym.nodes(), MX._wrap(arg).nodes()):
                mapping[s] = a
        res = []
        for o in self.outs:
            new = substitute(o.nodes(), mapping)
            res.append(MX(np.array(new, dtype=object).reshape(o.shape, order="F")))
        return res[0] if len(res) == 1 else res


# ------------------------------------------------------------------------------------------------------------
# Opti
# ------------------------------------------------------------------------------------------------------------
class OptiSol:
    def __init__(self, opti, x, stats):
        self.opti, self.x, self._stats = opti, x, stats

    def value(self, expr):
        m = MX._wrap(expr)
        env = self.opti._env(self.x)
        vals = evaluate(m.nodes(), env)
        out = np.array(vals, dtype=float).reshape(m.shape, order="F")
        if out.shape == (1, 1):
            return float(out[0, 0])
        if out.shape[1] == 1:
            return out[:, 0]
        return out

    def stats(self):
        return self._stats


class Opti:
    def __init__(self):
        self.vars = []       # scalar symbol nodes, in creation order (column-major per call)
        self.pars = []
        self.par_val = {}
        self.init = {}
        self.objective = None
        self.constraints = []
        self.solver_name = None
        self.solver_opts = {}

    def _sym(self, store, n, m, tag):
        arr = np.empty((n, m), dtype=object)
        for j in range(m):
            for i in range(n):
                nd = Node("sym", (), None, f"{tag}{len(store)}")
                store.append(nd)
                arr[i, j] = nd
        return MX(arr)

    def variable(self, n=1, m=1):
        return self._sym(self.vars, n, m, "x")

    def parameter(self, n=1, m=1):
        return self._sym(self.pars, n, m, "p")

    def bounded(self, lo, expr, hi):
        return Constraint("ineq", MX._wrap(expr), float(lo), float(hi))

    def subject_to(self, c):
        if not isinstance(c, Constraint):
            raise TypeError("subject_to expects a constraint")
        self.constraints.append(c)

    def minimize(self, expr):
        self.objective = MX._wrap(expr)

    def solver(self, name, opts=None):
        self.solver_name, self.solver_opts = name, dict(opts or {})

    def set_initial(self, var, value):
        v = np.asarray(value, dtype=float)
        m = MX._wrap(var)
        vals = np.broadcast_to(v.reshape(m.shape) if v.size == m.a.size else v, m.shape)
        for nd, val in zip(m.nodes(), vals.reshape(-1, order="F")):
            self.init[nd] = float(val)

    def set_value(self, par, value):
        m = MX._wrap(par)
        if isinstance(value, MX):
            vals = [nd.val for nd in value.nodes()]
        else:
            vals = np.asarray(value, dtype=float).reshape(-1, order="F")
        if len(vals) != m.a.size:
            raise ValueError("set_value: size mismatch")
        for nd, val in zip(m.nodes(), vals):
            self.par_val[nd] = float(val)

    # -- numeric access used by the golden generator
    def _env(self, x, dual=False):
        n = len(self.vars)
        env = {}
        for i, nd in enumerate(self.vars):
            if dual:
                g = np.zeros(n); g[i] = 1.0
                env[nd] = Dual(float(x[i]), g)
            else:
                env[nd] = float(x[i])
        for nd in self.pars:
            env[nd] = self.par_val.get(nd, np.nan)
        return env

    def x0(self):
        return np.array([self.init.get(nd, 0.0) for nd in self.vars])

    def eq_rows(self):
        rows = []
        for c in self.constraints:
            if c.kind == "eq":
                rows.extend(c.expr.nodes())
        return rows

    def ineq_rows(self):
        rows, lo, hi = [], [], []
        for c in self.constraints:
            if c.kind == "ineq":
                nd = c.expr.nodes()
                rows.extend(nd); lo.extend([c.lo] * len(nd)); hi.extend([c.hi] * len(nd))
        return rows, np.array(lo), np.array(hi)

    def eval_f(self, x, grad=False):
        if grad:
            d = evaluate(self.objective.nodes(), self._env(x, True), len(self.vars))[0]
            return d.v, d.g
        return evaluate(self.objective.nodes(), self._env(x))[0]

    def eval_g(self, x, jac=False):
        rows = self.eq_rows()
        if jac:
            d = evaluate(rows, self._env(x, True), len(self.vars))
            return np.array([t.v for t in d]), np.array([t.g for t in d])
        return np.array(evaluate(rows, self._env(x)))

    def eval_h(self, x, jac=False):
        rows, lo, hi = self.ineq_rows()
        if jac:
            d = evaluate(rows, self._env(x, True), len(self.vars))
            return np.array([t.v for t in d]), np.array([t.g for t in d]), lo, hi
        return np.array(evaluate(rows, self._env(x))), lo, hi

    def solve(self):
        """Substitute solver: scipy SLSQP on the traced NLP (structurally absent variables are pinned at their
        start value).  Raises RuntimeError on failure like Opti.solve()."""
        from scipy.optimize import minimize  # noqa: PLC0415
        x0 = self.x0()
        n = len(x0)
        # variables that appear nowhere (e.g. the 5 symbols of get_system_function, the overwritten X[:,0])
        _, g0 = self.eval_f(x0, True)
        _, Jg = self.eval_g(x0, True)
        _, Jh, lo, hi = self.eval_h(x0, True)
        used = (np.abs(Jg).sum(0) + np.abs(Jh).sum(0) + np.abs(g0)) > 0
        # a second probe point in case a derivative vanishes at the start point
        xp = x0 + 0.37
        _, g1 = self.eval_f(xp, True)
        _, Jg1 = self.eval_g(xp, True)
        used |= (np.abs(Jg1).sum(0) + np.abs(g1)) > 0
        idx = np.where(used)[0]

        def full(z):
            x = x0.copy(); x[idx] = z
            return x

        cons = [{"type": "eq", "fun": lambda z: self.eval_g(full(z)), "jac": lambda z: self.eval_g(full(z), True)[1][:, idx]}]
        # inequality rows that are single variables become bounds; anything else a general inequality
        blo, bhi = np.full(n, -np.inf), np.full(n, np.inf)
        rows, lo, hi = self.ineq_rows()
        var_index = {id(nd): i for i, nd in enumerate(self.vars)}
        general = []
        for r, l, h in zip(rows, lo, hi):
            if r.op == "sym" and id(r) in var_index:
                i = var_index[id(r)]
                blo[i], bhi[i] = max(blo[i], l), min(bhi[i], h)
            else:
                general.append((r, l, h))
        if general:
            raise NotImplementedError("general inequalities are not needed for the reference's problems")
        res = minimize(lambda z: self.eval_f(full(z), True)[0], x0[idx], jac=lambda z: self.eval_f(full(z), True)[1][idx],
                       method="SLSQP", bounds=list(zip(blo[idx], bhi[idx])), constraints=cons,
                       options={"maxiter": 1000, "ftol": 1e-14})
        # status 8 ("positive directional derivative for linesearch") is how SLSQP stops when it cannot improve a
        # converged point any further; accept it when the point is feasible
        feasible = np.abs(self.eval_g(full(res.x))).max() <= 1e-9
        if not (res.success or (res.status == 8 and feasible)):
            raise RuntimeError(f"Error in Opti::solve: substitute solver failed: {res.message}")
        return OptiSol(self, full(res.x), {"success": True, "return_status": "SLSQP: " + str(res.message), "iter_count": res.nit})
