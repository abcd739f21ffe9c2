"""QPS reader and the Maros-Meszaros bookkeeping of the reference's test-suite (host-side test support).

Mirrors `jrl::qp::test::QPSReader` (tests/QPSReader.h:17-117, tests/QPSReader.cpp:162-480): same section handling,
same conventions, same error situations (reported as QPSError with the `(line N, section S)` context the reference
puts in its exception text), and `QPSPbData` / the selection rules of the "Test Suite" test-case
(tests/QPSProblems.h:7-17, tests/GoldfarbIdnaniSolverTest.cpp:246-307).

Conventions restated from the reference:
  * a line starting with a blank is a data line of the current section; a line starting with a section keyword
    (NAME ROWS COLUMNS RHS RANGES BOUNDS QUADOBJ ENDATA, case-insensitive) opens that section; anything else is skipped;
  * ROWS: `E|L|G|N name`; the first N row is the objective, a second one is an error; the others are numbered in order;
  * COLUMNS / RHS / RANGES: `name  row value [row value]`; an RHS on the objective row is MINUS the objective constant;
    a second RHS / RANGES set name is an error;
  * row limits: E: l = u = rhs; L: l = -inf, u = rhs; G: l = rhs, u = +inf; RANGES R on a row: E: u += R if R >= 0 else
    l += R; L: l = u - |R|; G: u = l + |R|;
  * BOUNDS: default 0 <= x < +inf; LO, UP, FX (xl = xu, flags hasFixedVariables), FR, MI (xl = -inf), PL (xu = +inf);
  * QUADOBJ: `col row value [row value]` sets G(row, col); fullObjMat mirrors the strict lower triangle upwards;
  * useBounds = any finite bound.
Output matrices follow the reference's QPProblem: C is nbCstr x nbVar (row = constraint), so a solver call passes
C.T exactly as the reference's tests pass `pb.C.transpose()`.
"""
import dataclasses
import math

import numpy as np

_SECTIONS = ("name", "rows", "columns", "rhs", "ranges", "bounds", "quadobj", "endata")


class QPSError(RuntimeError):
    pass


@dataclasses.dataclass
class QPProblem:
    G: np.ndarray
    a: np.ndarray
    C: np.ndarray
    l: np.ndarray  # noqa: E741 - the reference's field name
    u: np.ndarray
    xl: np.ndarray
    xu: np.ndarray
    objCst: float = 0.0
    name: str = ""


@dataclasses.dataclass
class ProblemProperties:
    nbVar: int
    nbCstr: int
    nbEq: int
    useBounds: bool
    hasFixedVariables: bool


@dataclasses.dataclass
class QPSPbData:
    name: str
    fstar: float
    cond: float
    nbCstr: int
    nbVar: int


class QPSReader:
    def __init__(self, fullObjMat=False):
        self.fullObjMat = bool(fullObjMat)
        self.bigBnd = math.inf

    # ---- helpers ---------------------------------------------------------------------------------
    def _fail(self, message):
        raise QPSError(f"{message} (line {self._line}, section {self._section})")

    def _number(self, tok, what):
        try:
            return float(tok.replace("d", "e").replace("D", "E"))
        except ValueError:
            self._fail(f"Failed to read {what} value")

    def _value_line(self, line):
        """`name  key value [key value]` -> (name, [(key, value), ...])"""
        t = line.split()
        if len(t) < 3:
            self._fail("Failed to read first " + ("name" if len(t) < 2 else "value"))
        pairs = [(t[1], self._number(t[2], "first"))]
        if len(t) >= 4:
            if len(t) < 5:
                self._fail("Failed to read second value")
            pairs.append((t[3], self._number(t[4], "second")))
        return t[0], pairs

    def _row(self, name):
        try:
            return self._rows[name]
        except KeyError:
            self._fail(f"Unknown row {name}")

    def _col(self, name):
        try:
            return self._cols[name]
        except KeyError:
            self._fail(f"Unknown column {name}")

    # ---- parsing ---------------------------------------------------------------------------------
    def read(self, filename):
        self._rows, self._cols = {}, {}  # name -> (index, type) ; name -> index
        self._line, self._section = 0, "other"
        name = ""
        rhs_name = range_name = None
        obj_read = False
        n_rows = 0
        Cv, Gv, av, bv, rv, xv = [], [], [], [], [], []
        obj_cst = 0.0
        try:
            fh = open(filename)
        except OSError:
            self._fail(f"Unable to open file {filename}")
        with fh:
            for raw in fh:
                self._line += 1
                line = raw.rstrip("\r\n")
                if not line:
                    continue
                if line[0] != " ":
                    head = line.split()
                    key = head[0].lower() if head else ""
                    if key in _SECTIONS:
                        self._section = key
                        if key == "name":
                            if len(head) < 2:
                                self._fail("Failed to read name")
                            name = head[1]
                    continue  # any other non-indented line is ignored, as in the reference
                sec = self._section
                if sec == "name":
                    self._fail("We shouldn't be in a NAME section")
                elif sec == "endata":
                    self._fail("We shouldn't be in a ENDATA section")
                elif sec == "rows":
                    t = line.split()
                    if len(t) < 1:
                        self._fail("Failed to read row type")
                    ty = t[0][0].lower()  # the reference reads ONE character
                    rest = ([t[0][1:]] if len(t[0]) > 1 else []) + t[1:]
                    if ty not in "elgn":
                        self._fail("Unknown row type")
                    if not rest:
                        self._fail("Failed to read row name")
                    if rest[0] in self._rows:
                        self._fail("Duplicate row name")
                    if ty == "n":
                        if obj_read:
                            self._fail("We don't handle \"no restriction\" rows")
                        obj_read = True
                        self._rows[rest[0]] = (-1, "n")
                    else:
                        self._rows[rest[0]] = (n_rows, ty)
                        n_rows += 1
                elif sec == "columns":
                    col, pairs = self._value_line(line)
                    c = self._cols.setdefault(col, len(self._cols))
                    for rname, val in pairs:
                        r, ty = self._row(rname)
                        if ty == "n":
                            av.append((c, val))
                        else:
                            Cv.append((r, c, val))
                elif sec == "rhs":
                    nm, pairs = self._value_line(line)
                    if rhs_name is None:
                        rhs_name = nm
                    elif rhs_name != nm:
                        self._fail("Attempting to use different RHS name. I don't know what this means")
                    for rname, val in pairs:
                        r, ty = self._row(rname)
                        if ty == "n":
                            obj_cst = -val  # the rhs is on the other side of the objective row
                        else:
                            bv.append((r, val, ty))
                elif sec == "ranges":
                    nm, pairs = self._value_line(line)
                    if range_name is None:
                        range_name = nm
                    elif range_name != nm:
                        self._fail("Attempting to use different range name. I don't know what this means")
                    for rname, val in pairs:
                        r, ty = self._row(rname)
                        if ty == "n":
                            self._fail("Attempting to add range on a N row")
                        rv.append((r, val, ty))
                elif sec == "bounds":
                    t = line.split()
                    if not t:
                        self._fail("Unable to read bound type")
                    ty = t[0]
                    if ty not in ("LO", "UP", "FX", "FR", "MI", "PL"):
                        self._fail("Unknown bound type")
                    if len(t) < 2:
                        self._fail("Unable to read bound name")
                    if ty == "FR":
                        if len(t) < 3:
                            self._fail("Unable to read column name")
                        xv.append((self._col(t[2]), math.inf, ty))
                    else:
                        if len(t) < 3:
                            self._fail("Failed to read first name")
                        if len(t) < 4:
                            self._fail("Failed to read first value")
                        xv.append((self._col(t[2]), self._number(t[3], "first"), ty))
                elif sec == "quadobj":
                    col, pairs = self._value_line(line)
                    c = self._col(col)
                    for rname, val in pairs:
                        Gv.append((self._col(rname), c, val))

        # ---- populate (tests/QPSReader.cpp:186-307, same order of application)
        n, big = len(self._cols), self.bigBnd
        G, a = np.zeros((n, n)), np.zeros(n)
        Cm, lo, up = np.zeros((n_rows, n)), np.zeros(n_rows), np.zeros(n_rows)
        xl, xu = np.zeros(n), np.full(n, big)
        for r, c, v in Gv:
            G[r, c] = v
        if self.fullObjMat:
            G = np.tril(G) + np.tril(G, -1).T
        for c, v in av:
            a[c] = v
        for r, c, v in Cv:
            Cm[r, c] = v
        nb_eq = 0
        for i, ty in self._rows.values():
            if ty == "e":
                nb_eq += 1
            elif ty == "l":
                lo[i] = -big
            elif ty == "g":
                up[i] = big
        for i, v, ty in bv:
            if ty == "e":
                lo[i] = up[i] = v
            elif ty == "l":
                lo[i], up[i] = -big, v
            else:
                lo[i], up[i] = v, big
        for i, v, ty in rv:
            if ty == "e":
                if v >= 0:
                    up[i] += v
                else:
                    lo[i] += v
            elif ty == "l":
                lo[i] = up[i] - abs(v)
            else:
                up[i] = lo[i] + abs(v)
        fixed = False
        for i, v, ty in xv:
            if ty == "LO":
                xl[i] = v
            elif ty == "UP":
                xu[i] = v
            elif ty == "FX":
                xl[i] = xu[i] = v
                fixed = True
            elif ty == "FR":
                xl[i], xu[i] = -big, big
            elif ty == "MI":
                xl[i] = -big
            else:
                xu[i] = big
        use_bounds = bool((xl > -big).any() or (xu < big).any())
        qp = QPProblem(G, a, Cm, lo, up, xl, xu, obj_cst, name)
        return qp, ProblemProperties(n, n_rows, nb_eq, use_bounds, fixed)


def write_qps(filename, qp, name="QP"):
    """Inverse of QPSReader.read for a dense QPProblem (test support: round trips, fixture generation). Rows with
    l == u become E rows, one-sided rows L / G rows, two-sided rows an L row plus a RANGES entry; 17 significant digits."""
    n, m = qp.a.size, qp.l.size
    big = math.inf
    f = lambda v: repr(float(v))  # noqa: E731
    rows = []
    for i in range(m):
        lo, up = qp.l[i], qp.u[i]
        if lo == up:
            rows.append(("E", lo, None))
        elif lo == -big and up < big:
            rows.append(("L", up, None))
        elif up == big and lo > -big:
            rows.append(("G", lo, None))
        elif lo > -big and up < big:
            rows.append(("L", up, up - lo))
        else:
            raise ValueError("a free row cannot be written (the reference reader rejects extra N rows)")
    with open(filename, "w") as fh:
        fh.write(f"NAME          {name}\nROWS\n N  obj\n")
        for i, (ty, _, _) in enumerate(rows):
            fh.write(f" {ty}  r{i}\n")
        fh.write("COLUMNS\n")
        for j in range(n):
            wrote = False
            if qp.a[j] != 0.0:
                fh.write(f"    x{j}  obj  {f(qp.a[j])}\n")
                wrote = True
            for i in range(m):
                if qp.C[i, j] != 0.0:
                    fh.write(f"    x{j}  r{i}  {f(qp.C[i, j])}\n")
                    wrote = True
            if not wrote:  # a column must appear to exist
                fh.write(f"    x{j}  obj  0.0\n")
        fh.write("RHS\n")
        if qp.objCst != 0.0:
            fh.write(f"    rhs  obj  {f(-qp.objCst)}\n")
        for i, (_, b, _) in enumerate(rows):
            if b != 0.0:
                fh.write(f"    rhs  r{i}  {f(b)}\n")
        if any(r is not None for _, _, r in rows):
            fh.write("RANGES\n")
            for i, (_, _, r) in enumerate(rows):
                if r is not None:
                    fh.write(f"    rng  r{i}  {f(r)}\n")
        fh.write("BOUNDS\n")
        for j in range(n):
            lo, up = qp.xl[j], qp.xu[j]
            if lo == up:
                fh.write(f" FX bnd  x{j}  {f(lo)}\n")
                continue
            if lo == -big and up == big:
                fh.write(f" FR bnd  x{j}\n")
                continue
            if lo == -big:
                fh.write(f" MI bnd  x{j}  0.0\n")
            elif lo != 0.0:
                fh.write(f" LO bnd  x{j}  {f(lo)}\n")
            if up != big:
                fh.write(f" UP bnd  x{j}  {f(up)}\n")
        fh.write("QUADOBJ\n")
        for j in range(n):
            for i in range(j, n):
                if qp.G[i, j] != 0.0:
                    fh.write(f"    x{j}  x{i}  {f(qp.G[i, j])}\n")
        fh.write("ENDATA\n")


# Rows of the Maros-Meszaros table (tests/QPSProblems.h:22-160: name, f*, estimated cond(G), nbCstr, nbVar) for the
# problems whose QPS files this repository carries under tests/golden/qps/ (the repository data set itself is not in
# the reference checkout; these small problems were re-entered from their published definitions, see the README there).
INF = math.inf
marosMeszarosPbList = [
    QPSPbData("hs21", -9.9960000e+01, 100, 1, 2),
    QPSPbData("hs35", 1.1111111e-01, 16.3937, 1, 3),
    QPSPbData("hs35mod", 2.5000000e-01, 16.3937, 1, 3),
    QPSPbData("hs76", -4.6818182e+00, 16.3937, 3, 4),
    QPSPbData("qptest", 4.3718750e+00, 1.6612, 2, 2),
    QPSPbData("tame", 0.0000000e+00, 1.1568581e+17, 1, 2),
    QPSPbData("zecevic2", -4.1250000e+00, INF, 2, 2),
]


def suite_action(pb):
    """Selection rules of the reference's "Test Suite" loop (tests/GoldfarbIdnaniSolverTest.cpp:253-275):
    'skip' (ill-conditioned or too large), 'non_pos_hessian' (cond == Inf: the solve must return NON_POS_HESSIAN)
    or 'solve' (SUCCESS, testKKT, objective + objCst == f* to 1e-6)."""
    if 1e8 < pb.cond < INF:
        return "skip"
    if pb.nbVar > 500 or pb.nbCstr > 1000:
        return "skip"
    return "non_pos_hessian" if pb.cond == INF else "solve"


def suite_max_iter(pb):
    return max(50, 10 * max(pb.nbCstr, pb.nbVar))  # tests/GoldfarbIdnaniSolverTest.cpp:292
